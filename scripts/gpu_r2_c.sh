#!/bin/bash
# Round-2 GPU call C: LK v2 (aligned 48-wide TMA tiles, uniform chains).  Gate: tiny case under compute-sanitizer, then the
# campaign; only if both pass, the -m gpu suite and the steady-state A/B of the staged back-end variants.
set -u
mkdir -p gpurun_out
cat > /tmp/lk_tiny.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, cv2
from larvio_b200.config import Config
from larvio_b200 import synth, api
cfg = Config.load('configs/euroc_mono.yaml')
sq = synth.make_sequence(cfg.raw, 0, 2)
b = api.Batch(cfg, n_seq=1)
cl = cv2.createCLAHE(3.0, (8, 8))
A = cl.apply(sq.images[0])[None]; B = cl.apply(sq.images[1])[None]
P = cv2.goodFeaturesToTrack(A[0], 8, 0.01, 20).reshape(1, -1, 2).astype(np.float32)
P = np.concatenate([P, np.array([[[0.3, 0.2], [751.0, 479.0], [5.5, 470.2], [745.1, 3.9]]], np.float32)], 1)
out, st = b.k_lk(A, B, P, P.copy())
print('tiny lk ok', st.sum(), out[0, :2])
PY
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/lk_tiny.py > gpurun_out/r2c_sanitizer_tma.txt 2>&1; rc=$?; echo "sanitizer(tma) rc=$rc"
grep -E "Illegal|Invalid|at 0x|fe_lk|ERROR SUMMARY|tiny lk" gpurun_out/r2c_sanitizer_tma.txt | head -12
timeout 600 python scripts/gpu_check_lk.py > gpurun_out/r2c_lk_tma.txt 2>&1; rc2=$?; echo "campaign(tma) rc=$rc2"; tail -11 gpurun_out/r2c_lk_tma.txt | cut -c1-330
if [ $rc -ne 0 ] || [ $rc2 -ne 0 ]; then
  LVB_DEBUG_LK_NOTMA=1 timeout 600 python scripts/gpu_check_lk.py > gpurun_out/r2c_lk_notma.txt 2>&1; echo "campaign(no tma) rc=$?"; tail -11 gpurun_out/r2c_lk_notma.txt | cut -c1-330
  exit 1
fi
timeout 1700 python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r2c_pytest.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.txt
grep -E "passed|failed|PASSED|FAILED|XPASS|XFAIL|ERROR|^[0-9.]+s " gpurun_out/r2c_pytest.txt | tail -60
VARIANTS="none chol_blocked qr_lean gemm_dmma trsm_wide graph chol_blocked,qr_lean,gemm_dmma,trsm_wide,graph" STEPS=30 STREAMS=4 SKIP_PARITY=1 bash scripts/gpu_ab_variants.sh
