#!/bin/bash
# One short gpurun call for the Schmidt nuisance states on the device: per-call diagnostic tables of both Schmidt fixtures, then
# all reference-fixture tests (the non-Schmidt ones must still pass) and the loud-failure test.
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 100 python scripts/gpu_check_fixture.py schmidt_1d_oldest schmidt_3d_oldest > gpurun_out/${TAG}_schmidt_diag.txt 2>&1; echo "diag rc=$?"
tail -30 gpurun_out/${TAG}_schmidt_diag.txt
timeout 150 python -m pytest tests/test_gpu.py -q -m gpu -k "compiled_reference or unsupported" > gpurun_out/${TAG}_fixture_tests.txt 2>&1; echo "fixture tests rc=$?" | tee -a gpurun_out/${TAG}_fixture_tests.txt
tail -15 gpurun_out/${TAG}_fixture_tests.txt
