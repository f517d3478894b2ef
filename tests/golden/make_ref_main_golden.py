"""Generate tests/golden/ref_main_hybrid_selfstart.txt: the odometry and map points the REFERENCE's whole per-frame pipeline
publishes (build container only; committed with its output).

oracle/_ref/larvio_ref_main = src/image_processor.cpp + ORBDescriptor.cpp + larvio.cpp + StaticInitializer.cpp +
FlexibleInitializer.cpp compiled unmodified (`make ref_main`) behind the loop of app/larvioMain.cpp:87-117, on an on-disk EuRoC
ASL directory written from the synthetic generator: euroc.yaml defaults (hybrid filter, 5x6 SLAM grid), 150 frames, a 1.4-s
standstill so that the reference's own static initialiser starts the filter.  Lines: `ODO t R(9) p(3) v(3)` per publication,
`PTS S|A n id x y z ...` after every 10th one.  Also kept: the two files LarVio itself writes (msckf_2_state.txt, msckf_2_takeoff.txt).  tests/test_cpu.py runs the oracle pipeline against them, tests/test_gpu.py the drop-in
facade (larvio_b200/bin/larvio_shim_demo: the same loop on the shim classes over the CUDA library).
Usage: make ref_main && python tests/golden/make_ref_main_golden.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from larvio_b200.config import Config          # noqa: E402
from larvio_b200 import synth                  # noqa: E402
import ref_runner as rr                        # noqa: E402

SPEC = dict(seq=3, frames=150, static_until=1.4)

if __name__ == "__main__":
    cfg = Config.load(os.path.join(ROOT, "configs", "euroc_mono.yaml"))
    seq = synth.make_sequence(cfg.raw, SPEC["seq"], SPEC["frames"], static_until=SPEC["static_until"])
    with tempfile.TemporaryDirectory() as td:
        txt, state_log, takeoff_log = rr.run_reference_pipeline(cfg.raw, rr.write_mav(td, seq), with_logs=True)
    path = os.path.join(ROOT, "tests", "golden", "ref_main_hybrid_selfstart.txt")
    open(path, "w").write(txt + "\n")
    # the two files the reference's LarVio writes itself (larvio.cpp:388, 420-453): the on-disk OUTPUT format of SURVEY 8(f-4)
    open(os.path.join(ROOT, "tests", "golden", "ref_main_msckf_2_state.txt"), "w").write(state_log)
    open(os.path.join(ROOT, "tests", "golden", "ref_main_msckf_2_takeoff.txt"), "w").write(takeoff_log)
    odo, pts = rr.parse_odometry_lines(txt)
    print("%d odometry lines, %d map-point lists -> %s (%d KB)" % (len(odo), len(pts), os.path.relpath(path, ROOT), os.path.getsize(path) // 1024))
