#!/bin/bash
# round-2 call L: QR by row blocks in shared memory, in-place reflectors - back-end parity subset, then the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "backend_matches or fused_step or baseline_config or hybrid_slam or 3d_inverse or config_e_capacity or zupt or golden" 2>&1 | tail -4
timeout 600 python bench.py --cpu-frames 4 > gpurun_out/r2l_bench_1gpu_configC.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
b = json.loads([l for l in open("gpurun_out/r2l_bench_1gpu_configC.json") if l.startswith("{")][-1])
print("value %.0f e2e %.0f ms/step %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"]), "ekf", b["ekf_update_ms"]["per_sequence_ms"], "launches", b.get("gpu_launches"))
for k, v in sorted(b["kernels"].items(), key=lambda kv: -kv[1]["share"])[:12]:
    print("  %-28s %6.1f us x %3d  share %.3f" % (k, v["ms_per_launch"] * 1e3, v["launches"], v["share"]))
PY
