"""TEST INFRASTRUCTURE: parameter sweep of oracle/backend.py against the compiled reference filter (oracle/_ref/larvio_ref): settings the
fixtures do not visit (publishing rate, td, track length, window size, ZUPT switches, feature thresholds).  Build container only.
    python scripts/ref_sweep.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np                                  # noqa: E402
from larvio_b200.config import Config               # noqa: E402
from larvio_b200 import synth                       # noqa: E402
import ref_runner as rr                             # noqa: E402

Y = os.path.join(ROOT, "configs", "euroc_mono.yaml")
M = dict(max_features_in_one_grid=0)
CASES = [
    ("publish every frame (20 Hz)", dict(M, sw_size=16, pub_frequency=20), 30, 70, {}),
    ("publish at 5 Hz (40 IMU samples per update)", dict(M, sw_size=12, pub_frequency=5), 31, 120, {}),
    ("publish at 2.5 Hz (80 IMU samples per update)", dict(M, sw_size=8, pub_frequency=2.5), 32, 160, {}),
    ("td 5 ms, estimated", dict(M, sw_size=12, td=0.005), 33, 60, {}),
    ("td -8 ms, not estimated", dict(M, sw_size=12, td=-0.008, estimate_td=0), 34, 60, {}),
    ("max_track_len 10", dict(M, sw_size=14, max_track_len=10), 35, 70, {}),
    ("max_track_len 3, least_observation_number 2", dict(M, sw_size=12, max_track_len=3, least_observation_number=2), 36, 60, {}),
    ("sw_size 5", dict(M, sw_size=5), 37, 50, {}),
    ("sw_size 40", dict(M, sw_size=40), 38, 110, {}),
    ("ZUPT switched off, static start", dict(M, sw_size=12, if_ZUPT_valid=0), 39, 50, dict(static_until=1.0)),
    ("ZUPT threshold 0.02 (fires while moving)", dict(M, sw_size=12, zupt_max_feature_dis=0.02), 40, 70, {}),
    ("feature_translation_threshold 0.05", dict(M, sw_size=12, feature_translation_threshold=0.05), 41, 60, {}),
    ("hybrid, 2 features per cell, 3x3 grid", dict(sw_size=16, max_features_in_one_grid=2, aug_grid_rows=3, aug_grid_cols=3), 42, 130, {}),
    ("hybrid 3-D, max_track_len 4", dict(sw_size=16, feature_idp_dim=3, max_track_len=4), 43, 130, {}),
    ("hybrid, rotation_threshold 0 (oldest-pose pruning)", dict(sw_size=12, rotation_threshold=0.0), 44, 130, {}),
    ("hybrid, tracking_rate_threshold 1.5 (oldest-pose pruning)", dict(sw_size=12, tracking_rate_threshold=1.5), 45, 130, {}),
    ("Schmidt 3-D, tracking_rate_threshold 1.5", dict(sw_size=12, tracking_rate_threshold=1.5, use_schmidt=1, feature_idp_dim=3), 46, 150, {}),
    ("calibrating IMU intrinsics, no FEJ", dict(M, sw_size=12, calib_imu_instrinsic=1, if_FEJ=0), 47, 60, {}),
]
for name, ov, sid, nf, kw in (CASES if len(sys.argv) < 2 else [c for c in CASES if any(a in c[0] for a in sys.argv[1:])]):
    cfg = Config.load(Y, **ov)
    seq = synth.make_sequence(cfg.raw, sid, nf, **kw)
    calls = rr.record_calls(cfg.raw, seq, nf)
    j0 = calls[0]["frame"]
    init = (seq.img_t[j0], seq.gt_q[j0], seq.gt_p[j0], seq.gt_v[j0], np.zeros(3), np.zeros(3))
    try:
        b = rr.run_reference_on_calls(cfg.raw, calls, init, False)
        a = rr.run_oracle_on_calls(cfg.raw, calls, init, False)
        w = rr.compare_runs(a, b)
        oks = [r for r in b if r["ok"]]
        print("%-58s %3d updates, dim <= %3d, slam <= %2d, nui <= %d | q %.1e p %.1e v %.1e P %.1e" % (
            name, len(oks), max(r["P"].shape[0] for r in oks), max(len(r["slam_ids"]) for r in oks), max(len(r["nui_ids"]) for r in oks),
            w["q"], w["p"], w["v"], w["P"]), flush=True)
    except Exception as e:
        print("%-58s DIFFERS / FAILED: %s" % (name, str(e)[-300:]), flush=True)
