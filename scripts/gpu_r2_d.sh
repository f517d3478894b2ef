#!/bin/bash
# Round-2 GPU call D: gate (LK campaign), the whole -m gpu suite, the default bench line, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/gpu_check_lk.py > gpurun_out/r2d_lk.txt 2>&1; rc=$?; echo "campaign rc=$rc"; tail -2 gpurun_out/r2d_lk.txt | cut -c1-300
[ $rc -ne 0 ] && exit 1
timeout 2400 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.txt
grep -E "passed|failed|PASSED|FAILED|XPASS|XFAIL|ERROR|^[0-9.]+s " gpurun_out/r2d_pytest.txt | tail -70 | cut -c1-220
timeout 900 python bench.py --steps 60 --warmup 6 > gpurun_out/r2d_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2d_bench.log > gpurun_out/r2d_bench.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_bench.json"))
    print("value %.0f e2e %.0f ms/step %.3f launches/step %.1f cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"] / d["steps"], d["cpu_baseline"]))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"])[:24]:
        print("  %-28s %5.1f%% %7.1f us x %d" % (k, 100 * v["share"], 1e3 * v["ms_per_launch"], v["launches"]))
    print(d["steady_state"], d["lk_paths"], d["roofline"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 1500 bash scripts/gpu_profile.sh r2 > gpurun_out/r2d_profile.log 2>&1; echo "profile rc=$?"; tail -5 gpurun_out/r2d_profile.log
