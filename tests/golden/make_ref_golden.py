"""Generate tests/golden/ref_<case>.npz: golden vectors produced by the REFERENCE ITSELF (run in the build container, where
/root/reference exists; committed with its output).

For every case: a synthetic sequence -> the front-end oracle's feature messages + the IMU samples between them (the stream
app/larvioMain.cpp:87-117 hands LarVio::processFeatures) -> oracle/_ref/larvio_ref, i.e. the reference's own src/larvio.cpp +
src/StaticInitializer.cpp compiled unmodified against the stand-in headers of oracle/ref_shim/ (`make ref`).  A fixture holds the
recorded calls and, per call, the reference's state, extrinsics, td, IMU intrinsics, bookkeeping (state dimension, window size,
SLAM feature ids, nuisance states, IMU samples left) and a fingerprint of the covariance (P z for a fixed z, diag P, ||P||_F; the
full P of the last call).  tests/test_cpu.py checks oracle/backend.py and the compiled oracle against them; tests/test_gpu.py
checks the CUDA back end against them on the GPU box, which has no /root/reference.
Usage: make ref && python tests/golden/make_ref_golden.py [case ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from larvio_b200.config import Config          # noqa: E402
from larvio_b200 import synth                  # noqa: E402
import ref_runner as rr                        # noqa: E402

Y = os.path.join(ROOT, "configs", "euroc_mono.yaml")
# name: (config overrides, sequence id, frames, synth kwargs [+ "freeze": (a, b): images a..b-1 repeat image a, which the filter sees as a
#        standstill; "jump": (a, b): images a..b-1 show the last image], static initialiser instead of an injected state)
CASES = {
    "msckf_sw30": (dict(max_features_in_one_grid=0, sw_size=30), 10, 90, {}, False),                       # BASELINE configs[1]/[2]
    "msckf_oldest": (dict(max_features_in_one_grid=0, sw_size=12, translation_threshold=0.02), 2, 60, {}, False),
    "hybrid_1d_oldest": (dict(sw_size=12, translation_threshold=0.02), 0, 124, {}, False),
    "hybrid_3d": (dict(sw_size=16, feature_idp_dim=3), 0, 124, {}, False),
    "config_d": (dict(sw_size=16, calib_imu_instrinsic=1), 0, 120, {}, False),                              # BASELINE configs[3]
    "zupt": (dict(max_features_in_one_grid=0, sw_size=12), 5, 40, dict(static_until=1.0), False),
    "self_start": (dict(max_features_in_one_grid=0, sw_size=16), 3, 56, dict(static_until=1.4), True),
    "no_fej_no_calib": (dict(max_features_in_one_grid=0, sw_size=12, if_FEJ=0, estimate_extrin=0, estimate_td=0), 1, 50, {}, False),
    "calib_3d": (dict(sw_size=16, calib_imu_instrinsic=1, feature_idp_dim=3), 0, 116, {}, False),
    # a standstill in the middle of a hybrid run: checkZUPT drops every SLAM feature from the state (larvio.cpp:2770-2782) and blocks
    # promotions for 5 s; self start
    "hybrid_zupt": (dict(sw_size=16), 26, 176, dict(static_until=1.4, freeze=(150, 166)), True),
    # the scene jumps during the standstill: the static initialiser restarts its count (StaticInitializer.cpp:46-71) and starts later
    "self_start_jump": (dict(max_features_in_one_grid=0, sw_size=12), 61, 70, dict(static_until=2.2, jump=(8, 10)), True),
    "schmidt_1d_oldest": (dict(sw_size=12, translation_threshold=0.02, use_schmidt=1), 0, 150, {}, False),
    "schmidt_3d_oldest": (dict(sw_size=12, translation_threshold=0.02, use_schmidt=1, feature_idp_dim=3), 0, 150, {}, False),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        ov, sid, nf, kw, static_init = CASES[name]
        cfg = Config.load(Y, **ov)
        kw = dict(kw); freeze = kw.pop("freeze", None); jump = kw.pop("jump", None)
        seq = synth.make_sequence(cfg.raw, sid, nf, **kw)
        if freeze:
            seq.images = seq.images.copy(); seq.images[freeze[0]:freeze[1]] = seq.images[freeze[0]]
        if jump:                                                   # images jump[0]..jump[1]-1 show the END of the sequence
            seq.images = seq.images.copy(); seq.images[jump[0]:jump[1]] = seq.images[-1]
        calls = rr.record_calls(cfg.raw, seq, nf)
        j0 = calls[0]["frame"]
        init = None if static_init else (float(seq.img_t[j0]), seq.gt_q[j0], seq.gt_p[j0], seq.gt_v[j0], np.zeros(3), np.zeros(3))
        ref = rr.run_reference_on_calls(cfg.raw, calls, init, static_init)
        fx = rr.pack_fixture(ov, init, static_init, calls, ref)
        path = os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name)
        np.savez_compressed(path, **fx)
        oks = [r for r in ref if r["ok"]]
        print("%-20s %3d calls, %3d updates, max dim %3d, max SLAM %2d, max nuisance %d -> %s (%d KB)" % (
            name, len(ref), len(oks), max(r["P"].shape[0] for r in oks), max(len(r["slam_ids"]) for r in oks),
            max(len(r["nui_ids"]) for r in oks), os.path.relpath(path, ROOT), os.path.getsize(path) // 1024), flush=True)
