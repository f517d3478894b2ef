"""TEST INFRASTRUCTURE: oracle/frontend.py against the compiled reference front end (oracle/_ref/larvio_ref_fe) on doctored image / IMU
streams that push the FIRST / SECOND / OTHER image state machine and the track bookkeeping through their rare branches.  Build container only.
    python scripts/ref_fe_edge_cases.py"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2          # noqa: E402
import numpy as np  # noqa: E402
from larvio_b200.config import Config
from larvio_b200 import synth
import ref_runner as rr
cfg = Config.load(os.path.join(ROOT, "configs", "euroc_mono.yaml"), max_features_in_one_grid=0)
def case(name, sid, nf, edit):
    seq = copy.copy(synth.make_sequence(cfg.raw, sid, nf)); seq.images = seq.images.copy(); seq.imu = seq.imu.copy()
    edit(seq)
    try:
        ref = rr.run_reference_frontend(cfg.raw, seq, nf)
        calls = {c["frame"]: c for c in rr.record_calls(cfg.raw, seq, nf)}
        n_pub, bad, worst = rr.compare_fe([calls.get(j) for j in range(nf)], ref)
        sizes = [len(m["ids"]) for m in ref if m is not None]
        print("%-52s %3d frames, %3d published (%d..%d features), %d with differing ids, max column difference %.3g" % (name, nf, n_pub, min(sizes) if sizes else 0, max(sizes) if sizes else 0, bad, worst), flush=True)
    except AssertionError as e:
        print("%-52s %s" % (name, e), flush=True)
def grey_start(s): s.images[0:3] = 117
def blackout_after_second(s): s.images[2:5] = 117
def every_seventh(s): s.images[6::7] = 117
def blurred(s):
    for j in range(len(s.images)): s.images[j] = cv2.GaussianBlur(s.images[j], (0, 0), 9.0)
def late_imu(s): s.imu = s.imu[s.imu[:, 0] > s.img_t[3] + 0.001]
def dark_half(s): s.images[:, :, :376] = 20
def flicker(s):
    for j in range(0, len(s.images), 2): s.images[j] = (s.images[j].astype(np.int32) * 6 // 10 + 30).astype(np.uint8)
case("three grey frames first (initializeFirstFrame fails)", 50, 40, grey_start)
case("blackout right after the second image", 51, 40, blackout_after_second)
case("every seventh frame grey", 52, 60, every_seventh)
case("sigma-9 blurred images (few corners)", 53, 40, blurred)
case("IMU buffer starts after the fourth image", 54, 40, late_imu)
case("left half dark", 55, 40, dark_half)
case("brightness flicker on every other frame", 56, 40, flicker)
