"""GPU bring-up check: CLAHE/pyramid/blur bit-exactness and LK agreement vs cv2 (run via gpurun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from larvio_b200.config import Config
from larvio_b200 import synth, api

cfg = Config.load('configs/euroc_mono.yaml')
seqs = [synth.make_sequence(cfg.raw, s, 3) for s in range(2)]
b = api.Batch(cfg, n_seq=2)
imgs = np.stack([seqs[0].images[0], seqs[1].images[0], seqs[0].images[1], seqs[1].images[1]])
clahe, l1, l2, blur = b.k_pyramid(imgs)
cl = cv2.createCLAHE(3.0, (8, 8))
res = {}
for i in range(4):
    ref = cl.apply(imgs[i])
    r1 = cv2.pyrDown(ref); r2 = cv2.pyrDown(r1)
    rb = cv2.GaussianBlur(ref, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    res[f'img{i}'] = dict(clahe=int((ref != clahe[i]).sum()), l1=int((r1 != l1[i]).sum()), l2=int((r2 != l2[i]).sum()),
                          blur=int((rb != blur[i]).sum()))
print(json.dumps(res))
# LK
A = np.stack([cl.apply(seqs[s].images[0]) for s in range(2)])
Bn = np.stack([cl.apply(seqs[s].images[1]) for s in range(2)])
rng = np.random.default_rng(0)
P = []
for s in range(2):
    p = cv2.goodFeaturesToTrack(A[s], 200, 0.01, 20).reshape(-1, 2)
    extra = np.array([[0.3, 0.2], [751.0, 479.0], [5.5, 470.2], [745.1, 3.9], [-3.0, 100.0], [760.0, 200.0]], np.float32)
    p = np.concatenate([p, extra])[:208]
    if len(p) < 208:
        p = np.concatenate([p, rng.uniform(0, 470, (208 - len(p), 2)).astype(np.float32)])
    P.append(p)
P = np.stack(P).astype(np.float32)
init = P + rng.normal(0, 1.5, P.shape).astype(np.float32)
out, st = b.k_lk(A, Bn, P, init)
crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
for s in range(2):
    ref, rst, _ = cv2.calcOpticalFlowPyrLK(A[s], Bn[s], P[s].reshape(-1, 1, 2), init[s].reshape(-1, 1, 2).copy(),
                                           winSize=(21, 21), maxLevel=2, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    ref = ref.reshape(-1, 2); rst = rst.reshape(-1)
    both = (rst == 1) & (st[s] == 1)
    d = np.abs(ref[both] - out[s][both]).max(axis=1)
    print(json.dumps(dict(seq=s, status_mismatch=int((rst != st[s]).sum()), n_ok=int(both.sum()),
                          max_diff=float(d.max()), mean_diff=float(d.mean()), bit_identical=int((d == 0).sum()),
                          over_1e3=int((d > 1e-3).sum()))))
    bad = np.where(rst != st[s])[0]
    for i in bad[:10]:
        print('status mismatch', i, P[s][i], ref[i], out[s][i], rst[i], st[s][i])
print('launches', b.launches)
