"""TEST INFRASTRUCTURE: the oracle pipeline against the WHOLE compiled reference pipeline (oracle/_ref/larvio_ref_main, `make ref_main`)
on self-starting synthetic sequences, over filter configurations and seeds.  Build container only.
    python scripts/ref_pipeline_campaign.py [n_frames]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from larvio_b200.config import Config               # noqa: E402
from larvio_b200 import synth                       # noqa: E402
import ref_runner as rr                             # noqa: E402

Y = os.path.join(ROOT, "configs", "euroc_mono.yaml")
NF = int(sys.argv[1]) if len(sys.argv) > 1 else 150
CASES = [
    ("hybrid 1-D, seed 14", dict(), 14),
    ("hybrid 1-D, seed 15", dict(), 15),
    ("pure MSCKF, sw 30", dict(max_features_in_one_grid=0, sw_size=30), 16),
    ("hybrid 3-D", dict(feature_idp_dim=3), 17),
    ("Schmidt 1-D, oldest-pose pruning", dict(use_schmidt=1, sw_size=12, translation_threshold=0.02), 18),
    ("Schmidt 3-D, oldest-pose pruning", dict(use_schmidt=1, sw_size=12, translation_threshold=0.02, feature_idp_dim=3), 19),
    ("hybrid + IMU-intrinsic calibration", dict(calib_imu_instrinsic=1), 20),
    ("pure MSCKF, no FEJ / extrinsics / td", dict(max_features_in_one_grid=0, if_FEJ=0, estimate_extrin=0, estimate_td=0), 21),
    ("400 tracks, 50-pose window, 4x5 grid", dict(max_features_num=400, sw_size=50, aug_grid_rows=4, aug_grid_cols=5, min_distance=14), 22),
]
for name, ov, sid in CASES:
    cfg = Config.load(Y, **ov)
    seq = synth.make_sequence(cfg.raw, sid, NF, static_until=1.4)
    with tempfile.TemporaryDirectory() as td:
        mav = rr.write_mav(td, seq)
        t0 = time.time()
        try:
            ref = rr.run_reference_pipeline(cfg.raw, mav)
            t1 = time.time()
            w = rr.compare_odometry(rr.run_oracle_pipeline(cfg.raw, mav), ref)
            print("%-40s %3d publications, %d map-point lists | t %.0e R %.1e p %.1e v %.1e pts %.1e | reference %.0fs oracle %.0fs" % (
                name, w["n"], w["n_lists"], w["t"], w["R"], w["p"], w["v"], w["pts"], t1 - t0, time.time() - t1), flush=True)
        except Exception as e:
            print("%-40s %s" % (name, str(e)[-400:]), flush=True)
