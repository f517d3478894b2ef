"""Generate tests/golden/ref_fe_<case>.npz: the feature messages the REFERENCE's own front end publishes (run in the build
container, where /root/reference exists; committed with its output).

oracle/_ref/larvio_ref_fe = /root/reference/src/image_processor.cpp + src/ORBDescriptor.cpp compiled unmodified against the
stand-in headers of oracle/ref_shim/ (`make ref_fe`); every OpenCV function it calls is executed by the cv2 module of this image
through oracle/cv_server.py.  A fixture holds, per frame, whether processImage returned true and the MonoCameraMeasurement it
filled (ids in their order + the eight columns), plus a hash of the synthetic images.  tests/test_cpu.py checks oracle/frontend.py
against them bit for bit; tests/test_gpu.py checks the CUDA front end (lvb_process_images) on the GPU box.
Usage: make ref_fe && python tests/golden/make_ref_fe_golden.py [case ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_runner as rr                        # noqa: E402

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(rr.FE_CASES)):
        cfg, seq, nf = rr.fe_case_sequence(name)
        msgs = rr.run_reference_frontend(cfg.raw, seq, nf)
        path = os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name)
        np.savez_compressed(path, **rr.pack_fe_fixture(name, seq, msgs))
        pub = [m for m in msgs if m is not None]
        print("%-18s %3d frames, %3d messages, %3d..%3d features, ids up to %d -> %s (%d KB)" % (
            name, nf, len(pub), min(len(m["ids"]) for m in pub), max(len(m["ids"]) for m in pub),
            max(int(m["ids"].max()) for m in pub if len(m["ids"])), os.path.relpath(path, ROOT), os.path.getsize(path) // 1024), flush=True)
