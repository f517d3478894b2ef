#!/bin/bash
# Round-2 GPU call A: LK campaign (new TMA kernel) under a watchdog, the whole -m gpu suite, then the A/B of the staged variants
# at steady state (pre-rolled bench).   gpurun --timeout 2400 -- 'bash scripts/gpu_r2_a.sh'
set -u
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2; } > gpurun_out/r2a_env.txt 2>&1
timeout 600 python scripts/gpu_check_lk.py > gpurun_out/r2a_lk.txt 2>&1; echo "lk campaign rc=$?" | tee -a gpurun_out/r2a_lk.txt
tail -12 gpurun_out/r2a_lk.txt
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.txt
grep -E "passed|failed|PASSED|FAILED|XPASS|XFAIL|ERROR" gpurun_out/r2a_pytest.txt | tail -50
VARIANTS="none chol_blocked qr_lean gemm_dmma trsm_wide graph chol_blocked,qr_lean,gemm_dmma,trsm_wide chol_blocked,qr_lean,gemm_dmma,trsm_wide,graph" STEPS=40 STREAMS=4 SKIP_PARITY=1 bash scripts/gpu_ab_variants.sh
