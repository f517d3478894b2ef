"""GPU bring-up check 2: ORB / detector / undistort / RANSAC stage parity and the full front end
against the oracle on short sequences (run via gpurun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from larvio_b200.config import Config
from larvio_b200 import synth, api
from oracle.orb import OrbOracle
from oracle.frontend import ImageProcessorOracle

NF = int(os.environ.get('NF', '16'))
cfg = Config.load('configs/euroc_mono.yaml')
S = 2
seqs = [synth.make_sequence(cfg.raw, s, NF) for s in range(S)]
b = api.Batch(cfg, n_seq=S)
cl = cv2.createCLAHE(3.0, (8, 8))
rng = np.random.default_rng(1)

# ---- ORB
img = cl.apply(seqs[0].images[0])
pts = cv2.goodFeaturesToTrack(img, 200, 0.01, 20).reshape(-1, 2)
pts = pts + rng.uniform(-0.5, 0.5, pts.shape).astype(np.float32)
pts = np.concatenate([pts, np.array([[0.2, 0.3], [751, 479], [3.4, 476.5], [748.2, 2.2]], np.float32)])[:200]
ang, desc = b.k_orb(img[None], pts[None])
o = OrbOracle(img)
ra = np.array([o.ic_angle(p) for p in pts], np.float32)
rd = o.compute(pts)
print(json.dumps(dict(stage='orb', angle_mismatch=int((ra != ang[0]).sum()), desc_mismatch_rows=int((rd != desc[0]).any(axis=1).sum()),
                      desc_bits=int(np.unpackbits(rd ^ desc[0]).sum()))))

# ---- detector
imgs = np.stack([cl.apply(seqs[s].images[1]) for s in range(S)])
det, eig = b.k_detect(imgs, None, 200, return_eig=True)
for s in range(S):
    ref = cv2.goodFeaturesToTrack(imgs[s], 200, 0.01, 20).reshape(-1, 2)
    re = cv2.cornerMinEigenVal(imgs[s], 3, ksize=3)
    same = len(ref) == len(det[s]) and np.array_equal(ref, det[s])
    print(json.dumps(dict(stage='detect', seq=s, n_ref=len(ref), n_gpu=len(det[s]), identical=bool(same),
                          eig_mismatch=int((re != eig[s]).sum()), eig_maxdiff=float(np.abs(re - eig[s]).max()))))
mask = np.full(imgs.shape, 255, np.uint8)
for s in range(S):
    for p in cv2.goodFeaturesToTrack(imgs[s], 120, 0.01, 20).reshape(-1, 2):
        x, y = int(round(p[0])), int(round(p[1]))
        mask[s, max(y - 20, 0):min(y + 20, 479) + 1, max(x - 20, 0):min(x + 20, 751) + 1] = 0
det = b.k_detect(imgs, mask, [80, 57])
for s, want in enumerate([80, 57]):
    ref = cv2.goodFeaturesToTrack(imgs[s], want, 0.01, 20, mask=mask[s]).reshape(-1, 2)
    print(json.dumps(dict(stage='detect_masked', seq=s, n_ref=len(ref), n_gpu=len(det[s]),
                          identical=bool(len(ref) == len(det[s]) and np.array_equal(ref, det[s])))))

# ---- undistort
K = np.array([[cfg['intrinsics']['fx'], 0, cfg['intrinsics']['cx']], [0, cfg['intrinsics']['fy'], cfg['intrinsics']['cy']], [0, 0, 1.0]])
D = np.array([cfg['distortion_coeffs'][k] for k in ('k1', 'k2', 'p1', 'p2')])
p = rng.uniform([0, 0], [752, 480], (500, 2)).astype(np.float32)
for to_px in (0, 1):
    ref = cv2.undistortPoints(p.reshape(-1, 1, 2), K, D, R=np.eye(3), P=(K if to_px else np.eye(3))).reshape(-1, 2)
    out = b.k_undistort(p, bool(to_px))
    print(json.dumps(dict(stage='undistort', to_pixels=to_px, mismatch=int((ref != out).any(axis=1).sum()),
                          maxdiff=float(np.abs(ref - out).max()))))

# ---- RANSAC
P1, P2, REF = [], [], []
for trial in range(120):
    n = int(rng.integers(3, 220)) if trial % 4 else int(rng.integers(3, 20))
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(2, 9, n)], 1)
    rv = rng.normal(0, 0.03, 3); t = rng.normal(0, 0.08, 3)
    R, _ = cv2.Rodrigues(rv)
    x1 = (K @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    X2 = (R @ X.T).T + t; x2 = (K @ X2.T).T; x2 = x2[:, :2] / x2[:, 2:]
    x1 += rng.normal(0, 0.15, x1.shape); x2 += rng.normal(0, 0.15, x2.shape)
    no = int(n * rng.uniform(0, 0.3)); oi = rng.choice(n, no, replace=False)
    x2[oi] += rng.uniform(-15, 15, (no, 2))
    p1 = x1.astype(np.float32); p2 = x2.astype(np.float32)
    m = None
    if n >= 7:
        _, m = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
    REF.append(np.ones(n, np.uint8) if m is None else m.reshape(-1))
    P1.append(p1); P2.append(p2)
masks = b.k_ransac(P1, P2)
bad = [(i, len(P1[i]), int(REF[i].sum()), int(masks[i].sum())) for i in range(len(P1)) if not np.array_equal(REF[i], masks[i])]
print(json.dumps(dict(stage='ransac', trials=len(P1), mismatches=len(bad), detail=bad[:12])))

# ---- full front end vs oracle
oracles = [ImageProcessorOracle(cfg.raw) for _ in range(S)]
k = [0] * S
tot = dict(frames=0, msg_frames=0, id_mismatch_frames=0, n_mismatch_frames=0, has_mismatch=0, max_uv_diff=0.0, max_vel_diff=0.0)
first_bad = None
t0 = time.time()
for j in range(NF):
    rows = []
    for s in range(S):
        k[s] = synth.imu_window(seqs[s], k[s], seqs[s].img_t[j])
        rows.append(seqs[s].imu[:k[s]])
    imu, n_imu = api.Batch.pack_imu(rows)
    imgs = np.stack([seqs[s].images[j] for s in range(S)])
    t_img = np.array([seqs[s].img_t[j] for s in range(S)])
    feat, out_n, has = b.process_images(imgs, t_img, imu, n_imu)
    for s in range(S):
        msg = oracles[s].process_image(seqs[s].images[j], seqs[s].img_t[j], rows[s])
        tot['frames'] += 1
        if (msg is not None) != bool(has[s]):
            tot['has_mismatch'] += 1
            if first_bad is None: first_bad = (j, s, 'has', msg is not None, int(has[s]))
            continue
        if msg is None:
            continue
        tot['msg_frames'] += 1
        n = int(out_n[s])
        if n != len(msg.ids):
            tot['n_mismatch_frames'] += 1
            if first_bad is None: first_bad = (j, s, 'n', len(msg.ids), n, {a: (None if v is None else int(np.sum(v))) for a, v in oracles[s].trace.items() if a.endswith(('fwd', 'rev', 'desc', 'ransac'))})
            continue
        g = feat[s, :n]
        if not np.array_equal(g['id'], msg.ids):
            tot['id_mismatch_frames'] += 1
            if first_bad is None: first_bad = (j, s, 'ids')
            continue
        uv = np.stack([g['u'], g['v'], g['u_init'], g['v_init']], 1)
        vel = np.stack([g['u_vel'], g['v_vel'], g['u_init_vel'], g['v_init_vel']], 1)
        tot['max_uv_diff'] = max(tot['max_uv_diff'], float(np.abs(uv - msg.data[:, :4]).max()))
        tot['max_vel_diff'] = max(tot['max_vel_diff'], float(np.abs(vel - msg.data[:, 4:]).max()))
tot['first_bad'] = first_bad
tot['stage'] = 'frontend'
tot['sec'] = time.time() - t0
print(json.dumps(tot, default=str))
print('launches', b.launches)
