#!/bin/bash
# Round-2 GPU call E: the -m gpu suite on the new ORB gate / add_obs hash / propagate / RANSAC-round kernels, the headline bench line,
# and one bench line each for BASELINE configs[3] and configs[4] per GPU (workloads D and E).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2e_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/r2e_pytest.txt | cut -c1-200
[ $rc -ne 0 ] && exit 1
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%s: value %.0f e2e %.0f ms/step %.3f launches/step %.1f ekf_ms %s" % (d["config"]["workload"][:12], d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"] / d["steps"], d["ekf_update_ms"]["per_sequence_ms"]))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"])[:14]:
        print("  %-28s %5.1f%% %7.1f us x %d" % (k, 100 * v["share"], 1e3 * v["ms_per_launch"], v["launches"]))
    print(" ", d["steady_state"], d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
}
timeout 900 python bench.py --steps 60 --warmup 6 > gpurun_out/r2e_bench.log 2>&1; echo "bench C rc=$?"; tail -1 gpurun_out/r2e_bench.log > gpurun_out/r2e_bench_C.json; summ gpurun_out/r2e_bench_C.json
timeout 900 python bench.py --workload D --steps 20 --warmup 4 --profile-steps 4 --cpu-frames 0 > gpurun_out/r2e_bench_D.log 2>&1; echo "bench D rc=$?"; tail -1 gpurun_out/r2e_bench_D.log > gpurun_out/r2e_bench_D.json; summ gpurun_out/r2e_bench_D.json
timeout 1200 python bench.py --workload E --steps 16 --warmup 4 --profile-steps 4 --cpu-frames 0 > gpurun_out/r2e_bench_E.log 2>&1; echo "bench E rc=$?"; tail -1 gpurun_out/r2e_bench_E.log > gpurun_out/r2e_bench_E.json; summ gpurun_out/r2e_bench_E.json
