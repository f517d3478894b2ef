#!/bin/bash
# One short gpurun call: (1) the CUDA filter against the reference-made fixtures (tests/golden/ref_*.npz), (2) the GPU tests that
# exercise what changed with them (persistent out-of-range grid cells, the latched bFirstFeatures gate, map points).
# usage: bash scripts/gpu_verify_fixtures.sh <tag>
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu.py -x -q -m gpu -k "compiled_reference" > gpurun_out/${TAG}_fixture_tests.txt 2>&1; echo "fixture tests rc=$?" | tee -a gpurun_out/${TAG}_fixture_tests.txt
tail -5 gpurun_out/${TAG}_fixture_tests.txt
timeout 230 python -m pytest tests/test_gpu.py -q -m gpu -k "hybrid_slam_features_match_oracle or self_start_with or shim_facade" > gpurun_out/${TAG}_touched_tests.txt 2>&1; echo "touched tests rc=$?" | tee -a gpurun_out/${TAG}_touched_tests.txt
tail -5 gpurun_out/${TAG}_touched_tests.txt
