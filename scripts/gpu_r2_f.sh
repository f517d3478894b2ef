#!/bin/bash
# Round-2 GPU call F: the -m gpu suite on the final kernels, the headline bench line, workload D, and the ncu evidence for profiles/.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2f_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/r2f_pytest.txt | cut -c1-200
[ $rc -ne 0 ] && exit 1
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%s: value %.0f e2e %.0f ms/step %.3f launches/step %.1f ekf_ms %s" % (d["config"]["workload"][:12], d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"] / d["steps"], d["ekf_update_ms"]["per_sequence_ms"]))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"])[:16]:
        print("  %-28s %5.1f%% %7.1f us x %d" % (k, 100 * v["share"], 1e3 * v["ms_per_launch"], v["launches"]))
    print(" ", d["steady_state"], d["cpu_baseline"], d["roofline"])
except Exception as e:
    print("bench parse failed", e)
PY
}
timeout 900 python bench.py --steps 60 --warmup 6 > gpurun_out/r2f_bench.log 2>&1; echo "bench C rc=$?"; tail -1 gpurun_out/r2f_bench.log > gpurun_out/r2f_bench_C.json; summ gpurun_out/r2f_bench_C.json
timeout 900 python bench.py --workload D --steps 20 --warmup 4 --profile-steps 4 --cpu-frames 0 > gpurun_out/r2f_bench_D.log 2>&1; echo "bench D rc=$?"; tail -1 gpurun_out/r2f_bench_D.log > gpurun_out/r2f_bench_D.json; summ gpurun_out/r2f_bench_D.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/r2f_bench_ref.log | cut -c1-900
KERNELS="lk_kernel ransac_kernel be_qr_kernel be_gemm_kernel orb_gate_kernel be_propagate_kernel be_stack_kernel be_feature_kernel select_kernel corner_kernel be_add_obs_kernel be_chol_kernel clahe_apply_kernel blur7_kernel pyrdown_kernel" KEEP="lk_kernel be_gemm_kernel" timeout 1500 bash scripts/gpu_profile.sh r2f > gpurun_out/r2f_profile.log 2>&1; echo "profile rc=$?"; tail -4 gpurun_out/r2f_profile.log
