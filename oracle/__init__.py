"""CPU oracle for the LARVIO per-frame hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product (``larvio_b200``)
never does: it fails loudly when its CUDA library is missing.

Parity status: **pinned**.  The reference ships no tests, fixtures or golden vectors, so the pins are:
* front end — the very OpenCV functions the reference calls, executed by cv2 4.13 (pinned in this image), with the glue of
  image_processor.cpp restated around them (frontend.py, orb.py, lk_exact.py, ransac.py: each checked against cv2 itself), and the
  glue itself pinned by the reference's OWN image_processor.cpp + ORBDescriptor.cpp: `make ref_fe` compiles them unmodified against
  oracle/ref_shim/opencv2/lvb_cv.hpp (OpenCV functions forwarded to cv2 through oracle/cv_server.py); tests/golden/ref_fe_*.npz hold
  the messages it published, frontend.py reproduces them bit for bit (tests/test_cpu.py);
* back end — golden vectors produced by the reference's OWN filter: `make ref` compiles /root/reference/src/larvio.cpp,
  StaticInitializer.cpp and FlexibleInitializer.cpp unmodified against the stand-in headers of oracle/ref_shim/ (Eigen / boost /
  OpenCV-core subsets written for this purpose; the real libraries are not in the image) into oracle/_ref/larvio_ref, driven by
  oracle/ref_driver.cpp.  tests/golden/make_ref_golden.py records its answers in tests/golden/ref_*.npz; backend.py,
  backend_c.cpp and the CUDA filter are tested against them (tests/test_cpu.py, tests/test_gpu.py).  The pin is of the reference's
  logic; its linear algebra runs on the stand-in (rounding differs from real Eigen, nothing else).
tests/golden/oracle_seq.npz (tests/golden/make_golden.py) additionally pins the synthetic generator and the front-end oracle.

backend_c.cpp / orb_c.cpp are compiled twins of backend.py (pure MSCKF scope) and orb.py,
pinned to them by tests/test_cpu.py and used by bench.py's CPU legs so that the CPU arm
times compiled code like the reference; they are built into oracle/_build/ by `make`.
"""
