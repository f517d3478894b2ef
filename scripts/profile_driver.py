"""Fork-free driver for ncu: `gen` writes a synthetic batch to disk (multiprocess, NOT under ncu),
`run` replays it through lvb_step in a single process (the process ncu attaches to)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from larvio_b200.config import Config          # noqa: E402
from larvio_b200 import synth                  # noqa: E402

S = int(os.environ.get("S", "64")); NF = int(os.environ.get("NF", "14"))
PATH = os.environ.get("BATCH_NPZ", "/tmp/lvb_batch.npz")
cfg = Config.load(os.path.join(ROOT, "configs", "euroc_mono.yaml"), max_features_in_one_grid=0, sw_size=30)

if sys.argv[1] == "gen":
    sys.path.insert(0, ROOT)
    import bench
    seqs = bench.generate(cfg.raw, list(range(S)), NF, os.cpu_count() or 1)
    np.savez(PATH, images=np.stack([s.images for s in seqs]), img_t=np.stack([s.img_t for s in seqs]),
             imu=np.stack([s.imu for s in seqs]), q=np.stack([s.gt_q[0] for s in seqs]), p=np.stack([s.gt_p[0] for s in seqs]),
             v=np.stack([s.gt_v[0] for s in seqs]))
    print("generated", S, NF)
else:
    from larvio_b200 import api
    d = np.load(PATH)
    images, img_t, imu_all = d["images"], d["img_t"], d["imu"]
    b = api.Batch(cfg, n_seq=S)
    for s in range(S):
        b.set_initial_state(s, img_t[s, 0], d["q"][s], d["p"][s], d["v"][s], np.zeros(3), np.zeros(3))
    imu = np.zeros((S, 96), api.IMU_DTYPE); n_imu = np.zeros(S, np.int32); k = np.zeros(S, np.int64)
    # ncu --profile-from-start off: only frames [PF, PF+PN) are profiled (driver-API range, independent of the runtime in use)
    import ctypes
    PF = int(os.environ.get("PF", "-1")); PN = int(os.environ.get("PN", "2"))
    cuda = ctypes.CDLL("libcuda.so.1")
    for j in range(NF):
        if j == PF:
            b.synchronize(); cuda.cuProfilerStart()
        if j == PF + PN:
            b.synchronize(); cuda.cuProfilerStop()
        for s in range(S):
            k2 = int(k[s])
            while k2 < imu_all.shape[1] and imu_all[s, k2, 0] - img_t[s, j] < 0.05:
                k2 += 1
            r = imu_all[s, int(k[s]):k2]; n = int(n_imu[s]); m = len(r)
            imu["t"][s, n:n + m] = r[:, 0]; imu["gyro"][s, n:n + m] = r[:, 1:4]; imu["acc"][s, n:n + m] = r[:, 4:7]
            n_imu[s] = n + m; k[s] = k2
        b.step(np.ascontiguousarray(images[:, j]), img_t[:, j], imu, n_imu)
    print("ran", NF, "frames,", b.launches, "launches")
