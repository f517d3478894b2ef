"""LK campaign against cv2.calcOpticalFlowPyrLK (run via gpurun): the TMA-staged kernel must be bit-identical in
positions and status on (a) CLAHE'd synthetic frames, (b) high-contrast frames that force the slow (chain-replay)
paths, (c) large initial-flow errors that force the search tile to be re-staged.  Prints one JSON line per case with
the kernel's own path counters (lvb_get_stats [10..13])."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from larvio_b200.config import Config
from larvio_b200 import synth, api

cfg = Config.load('configs/euroc_mono.yaml')
NS = 4
seqs = [synth.make_sequence(cfg.raw, s, 4) for s in range(NS)]
b = api.Batch(cfg, n_seq=NS)
cl = cv2.createCLAHE(3.0, (8, 8))
crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
rng = np.random.default_rng(0)
M = 416


def points(img, m):
    p = cv2.goodFeaturesToTrack(img, m - 40, 0.01, 10).reshape(-1, 2)
    extra = np.array([[0.3, 0.2], [751.0, 479.0], [5.5, 470.2], [745.1, 3.9], [-3.0, 100.0], [760.0, 200.0], [10.4, 10.6], [741.2, 469.9]], np.float32)
    p = np.concatenate([p, extra])
    if len(p) < m:
        p = np.concatenate([p, rng.uniform([0, 0], [751, 479], (m - len(p), 2)).astype(np.float32)])
    return p[:m].astype(np.float32)


def run(name, A, B, sigma):
    P = np.stack([points(A[s], M) for s in range(NS)])
    init = (P + rng.normal(0, sigma, P.shape)).astype(np.float32)
    s0 = b.stats()
    out, st = b.k_lk(A, B, P, init)
    s1 = b.stats()
    rep = dict(case=name, sigma=sigma, status_mismatch=0, pos_mismatch=0, n_ok=0, max_diff=0.0)
    for s in range(NS):
        ref, rst, _ = cv2.calcOpticalFlowPyrLK(A[s], B[s], P[s].reshape(-1, 1, 2), init[s].reshape(-1, 1, 2).copy(), winSize=(21, 21),
                                               maxLevel=2, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        ref = ref.reshape(-1, 2); rst = rst.reshape(-1)
        rep['status_mismatch'] += int((rst != st[s]).sum())
        ok = (rst == 1) & (st[s] == 1)
        rep['n_ok'] += int(ok.sum())
        neq = (ref[ok] != out[s][ok]).any(1)
        rep['pos_mismatch'] += int(neq.sum())
        if ok.any():
            rep['max_diff'] = max(rep['max_diff'], float(np.abs(ref[ok] - out[s][ok]).max()))
    d = [int(s1[k] - s0[k]) for k in (10, 11, 12, 13)]
    rep.update(iterations=d[0], slow_iterations=d[1], slow_setups=d[2], restages=d[3], points=NS * M)
    print(json.dumps(rep), flush=True)
    return rep


bad = 0
# (a) the bench's own image statistics
A = np.stack([cl.apply(seqs[s].images[0]) for s in range(NS)]); B = np.stack([cl.apply(seqs[s].images[1]) for s in range(NS)])
for sg in (0.5, 1.5, 4.0, 9.0):
    r = run('synthetic', A, B, sg); bad += r['status_mismatch'] + r['pos_mismatch']
# (b) high contrast: binary noise blurred a little, shifted by a sub-pixel affine warp
hc = []
for s in range(NS):
    base = (rng.integers(0, 2, (480, 752)) * 255).astype(np.uint8)
    base = cv2.GaussianBlur(base, (0, 0), 0.8 + 0.3 * s)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    Mw = np.array([[1.0, 0.002, 1.3 + s], [-0.002, 1.0, -0.7]], np.float64)
    hc.append((base, cv2.warpAffine(base, Mw, (752, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)))
A = np.stack([h[0] for h in hc]); B = np.stack([h[1] for h in hc])
for sg in (1.0, 3.0):
    r = run('high_contrast', A, B, sg); bad += r['status_mismatch'] + r['pos_mismatch']
# (c) frames two apart (larger true motion) with a poor prediction
A = np.stack([cl.apply(seqs[s].images[0]) for s in range(NS)]); B = np.stack([cl.apply(seqs[s].images[3]) for s in range(NS)])
for sg in (2.0, 12.0):
    r = run('far_frames', A, B, sg); bad += r['status_mismatch'] + r['pos_mismatch']
print('LK_CAMPAIGN_MISMATCHES', bad)
sys.exit(1 if bad else 0)
