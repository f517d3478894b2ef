"""CPU oracle for the LARVIO per-frame hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product (``larvio_b200``)
never does: it fails loudly when its CUDA library is missing.

Parity status: **parity unpinned** by the reference's own tests — the reference ships no
tests, fixtures or golden vectors and cannot be compiled here (no
Eigen / OpenCV C++ / SuiteSparse / Boost).  The front-end oracle therefore calls the
very OpenCV functions the reference calls (cv2 4.13, pinned in this image) and restates
the glue of image_processor.cpp around them; the back-end oracle is a numpy float64
restatement of larvio.cpp / feature.hpp.  Golden vectors under tests/golden/ are
generated from this oracle by tests/golden/make_golden.py.

backend_c.cpp / orb_c.cpp are compiled twins of backend.py (pure MSCKF scope) and orb.py,
pinned to them by tests/test_cpu.py and used by bench.py's CPU legs so that the CPU arm
times compiled code like the reference; they are built into oracle/_build/ by `make`.
"""
