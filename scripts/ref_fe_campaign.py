"""TEST INFRASTRUCTURE: oracle/frontend.py against the compiled reference front end (oracle/_ref/larvio_ref_fe, `make ref_fe`) over
more sequences and settings than the committed fixtures hold.  Build container only.
    python scripts/ref_fe_campaign.py [n_frames]
Prints per case the number of published frames, frames with differing ids, and the largest difference of the message columns."""
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
NF = int(sys.argv[1]) if len(sys.argv) > 1 else 120
CASES = [
    ("seed 4", dict(max_features_in_one_grid=0), 4, NF, {}),
    ("seed 5", dict(max_features_in_one_grid=0), 5, NF, {}),
    ("seed 6 pub 20 Hz", dict(max_features_in_one_grid=0, pub_frequency=20), 6, NF, {}),
    ("seed 7 pub 5 Hz", dict(max_features_in_one_grid=0, pub_frequency=5), 7, NF, {}),
    ("seed 8 no CLAHE", dict(max_features_in_one_grid=0, flag_equalize=0), 8, NF, {}),
    ("seed 9 equidistant", dict(max_features_in_one_grid=0, distortion_model="equidistant"), 9, NF // 2, {}),
    ("seed 10 3 pyramid levels", dict(max_features_in_one_grid=0, pyramid_levels=3), 10, NF // 2, {}),
    ("seed 11 patch 15", dict(max_features_in_one_grid=0, patch_size=15), 11, NF // 2, {}),
    ("seed 12 static start", dict(max_features_in_one_grid=0), 12, NF // 2, dict(static_until=1.2)),
    ("seed 13 image noise 3", dict(max_features_in_one_grid=0), 13, NF // 2, dict(image_noise=3.0)),
]
for name, ov, sid, nf, kw in CASES:
    cfg = Config.load(Y, **ov)
    seq = synth.make_sequence(cfg.raw, sid, nf, **kw)
    t0 = time.time()
    try:
        ref = rr.run_reference_frontend(cfg.raw, seq, nf)
    except Exception as e:
        print("%-26s reference run failed: %s" % (name, str(e)[-300:]), flush=True); continue
    t1 = time.time()
    calls = {c["frame"]: c for c in rr.record_calls(cfg.raw, seq, nf)}
    try:
        n_pub, bad, worst = rr.compare_fe([calls.get(j) for j in range(nf)], ref)
        print("%-26s %3d frames, %3d published, %d with differing ids, max column difference %.3g   (reference %.0fs, oracle %.0fs)" % (
            name, nf, n_pub, bad, worst, t1 - t0, time.time() - t1), flush=True)
    except AssertionError as e:
        print("%-26s %s" % (name, e), flush=True)
