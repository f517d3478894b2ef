"""TEST INFRASTRUCTURE: oracle/backend.py against the compiled reference (oracle/_ref/larvio_ref, `make ref`) over the filter
configurations the GPU parity tests use.  Run in the build container (needs /root/reference):
    python scripts/ref_campaign.py [case ...]
Prints, per case, the largest deviation of state and covariance over all processFeatures calls."""
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
E = dict(max_features_num=400, sw_size=50, aug_grid_rows=4, aug_grid_cols=5, min_distance=14)
CASES = {
    # name: (config overrides, sequence id, frames, synth kwargs, static_init)
    "msckf_sw12": (dict(max_features_in_one_grid=0, sw_size=12), 0, 60, {}, False),
    "msckf_sw30": (dict(max_features_in_one_grid=0, sw_size=30), 10, 90, {}, False),
    "hybrid_1d": (dict(sw_size=16), 0, 130, {}, False),
    "hybrid_3d": (dict(sw_size=16, feature_idp_dim=3), 0, 130, {}, False),
    "calib_msckf": (dict(sw_size=16, max_features_in_one_grid=0, calib_imu_instrinsic=1), 0, 70, {}, False),
    "config_d": (dict(sw_size=16, calib_imu_instrinsic=1), 0, 124, {}, False),
    "calib_3d": (dict(sw_size=16, calib_imu_instrinsic=1, feature_idp_dim=3), 0, 124, {}, False),
    "zupt": (dict(max_features_in_one_grid=0, sw_size=12), 5, 40, dict(static_until=1.0), False),
    "self_start": (dict(max_features_in_one_grid=0, sw_size=16), 3, 56, dict(static_until=1.4), True),
    "config_e": (E, 0, 132, {}, False),
    "no_fej_no_calib": (dict(max_features_in_one_grid=0, sw_size=12, if_FEJ=0, estimate_extrin=0, estimate_td=0), 1, 50, {}, False),
    # translation_threshold 0.02: the two newest poses are never "redundant", so findRedundantImuStates removes the OLDEST ones
    # (larvio.cpp:2292-2297): marginalised poses carry real measurements, old anchors are handed over / become nuisance states
    "msckf_oldest": (dict(max_features_in_one_grid=0, sw_size=12, translation_threshold=0.02), 2, 70, {}, False),
    "hybrid_1d_oldest": (dict(sw_size=12, translation_threshold=0.02), 0, 130, {}, False),
    "hybrid_3d_oldest": (dict(sw_size=12, translation_threshold=0.02, feature_idp_dim=3), 0, 130, {}, False),
    "schmidt_1d_oldest": (dict(sw_size=12, translation_threshold=0.02, use_schmidt=1), 0, 150, {}, False),
    "schmidt_3d_oldest": (dict(sw_size=12, translation_threshold=0.02, use_schmidt=1, feature_idp_dim=3), 0, 150, {}, False),
    "schmidt_1d": (dict(sw_size=16, use_schmidt=1), 0, 130, {}, False),
    "schmidt_3d": (dict(sw_size=16, use_schmidt=1, feature_idp_dim=3), 0, 130, {}, False),
}


def run_case(name):
    ov, sid, nf, kw, static_init = CASES[name]
    cfg = Config.load(Y, **ov)
    t0 = time.time()
    seq = synth.make_sequence(cfg.raw, sid, nf, **kw)
    calls = rr.record_calls(cfg.raw, seq, nf)
    t1 = time.time()
    j0 = calls[0]["frame"]
    init = None if static_init else (seq.img_t[j0], seq.gt_q[j0], seq.gt_p[j0], seq.gt_v[j0], np.zeros(3), np.zeros(3))
    b = rr.run_reference_on_calls(cfg.raw, calls, init, static_init)
    t2 = time.time()
    try:
        a = rr.run_oracle_on_calls(cfg.raw, calls, init, static_init)
    except NotImplementedError as e:
        print("%-16s reference only (%s): %d calls, %d ok, max dim %d, max nuisance %d" % (
            name, e, len(b), sum(r["ok"] for r in b), max(r["P"].shape[0] for r in b if r["ok"]), max(len(r["nui_ids"]) for r in b if r["ok"])))
        return
    w = rr.compare_runs(a, b)
    oks = [r for r in b if r["ok"]]
    print("%-16s calls %3d ok %3d  max dim %3d  max slam %2d  max nui %d | q %.1e p %.1e v %.1e bg %.1e ba %.1e ext %.1e td %.1e P %.1e | synth+fe %.0fs ref %.1fs oracle %.0fs" % (
        name, len(b), len(oks), max(r["P"].shape[0] for r in oks), max(len(r["slam_ids"]) for r in oks), max(len(r["nui_ids"]) for r in oks), w["q"], w["p"], w["v"], w["bg"], w["ba"],
        w["ext"], w["td"], w["P"], t1 - t0, t2 - t1, time.time() - t2), flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
