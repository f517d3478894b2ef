"""GPU check: front end (processImage) against the oracle over long sequences; at the first id mismatch, the oracle's
RANSAC inputs of that frame are replayed through lvbk_ransac and saved for offline analysis."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from larvio_b200.config import Config
from larvio_b200 import synth, api, harness
from oracle.frontend import ImageProcessorOracle

NF = int(os.environ.get('NF', '124')); IDS = [int(x) for x in os.environ.get('SEQS', '0,1').split(',')]; S = len(IDS)
cfg = Config.load('configs/euroc_mono.yaml')
seqs = [synth.make_sequence(cfg.raw, s, NF) for s in IDS]
b = api.Batch(cfg, n_seq=S)
fes = [ImageProcessorOracle(cfg.raw) for _ in range(S)]
calls = []
for fe in fes:
    orig = fe._ransac
    def wrap(p1, p2, orig=orig, fe=fe):
        m = orig(p1, p2)
        calls.append((fe, p1.copy(), p2.copy(), None if m is None else m.copy()))
        return m
    fe._ransac = wrap
feed = harness.ImuFeeder(seqs, stride=2048)
imu_o = [[] for _ in range(S)]; k = [0] * S
bad = 0; n_r = 0; n_bad_r = 0
for j in range(NF):
    feed.push_until(j)
    del calls[:]
    msgs = []
    for s in range(S):
        k2 = synth.imu_window(seqs[s], k[s], seqs[s].img_t[j]); imu_o[s].extend(seqs[s].imu[k[s]:k2].tolist()); k[s] = k2
        msgs.append(fes[s].process_image(seqs[s].images[j], seqs[s].img_t[j], np.array(imu_o[s]).reshape(-1, 7)))
    # replay every RANSAC input of this frame through the CUDA kernel
    for (fe, p1, p2, m) in calls:
        if m is None or len(p1) < 15:
            continue
        n_r += 1
        gm = b.k_ransac([p1.astype(np.float32)], [p2.astype(np.float32)])[0]
        if not np.array_equal(gm.astype(bool), m.astype(bool)):
            n_bad_r += 1
            si = fes.index(fe)
            np.savez('gpurun_out/ransac_mismatch_%d_%d_%d.npz' % (IDS[si], j, n_bad_r), p1=p1, p2=p2, cv=m, gpu=gm)
            print('frame', j, 'seq', IDS[si], 'RANSAC mask mismatch: n', len(p1), 'cv inliers', int(m.astype(bool).sum()), 'gpu inliers', int(gm.astype(bool).sum()), flush=True)
    imgs = np.stack([seqs[s].images[j] for s in range(S)]); t_img = np.array([seqs[s].img_t[j] for s in range(S)])
    feat, out_n, has = b.process_images(imgs, t_img, feed.buf, feed.n)
    for s in range(S):
        if bool(has[s]) != (msgs[s] is not None):
            print('frame', j, 'seq', IDS[s], 'publish mismatch'); bad += 1; continue
        if msgs[s] is None:
            continue
        g = feat[s, :out_n[s]]
        if len(g) != len(msgs[s].ids) or not np.array_equal(g['id'], msgs[s].ids):
            a = set(g['id'].tolist()); o = set(msgs[s].ids.tolist())
            print('frame', j, 'seq', IDS[s], 'ID MISMATCH gpu-only', sorted(a - o)[:10], 'oracle-only', sorted(o - a)[:10], flush=True)
            bad += 1
    if bad:
        break
print(json.dumps(dict(frames=j + 1, id_mismatch_frames=bad, ransac_calls=n_r, ransac_mismatch=n_bad_r)))
