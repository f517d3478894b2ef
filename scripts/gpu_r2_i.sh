#!/bin/bash
# round-2 call I: full GPU suite (incl. 3-D inverse depth), bench both arms with the compiled CPU oracle
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2i_pytest.txt; cat gpurun_out/r2i_pytest.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2i_bench_reference_arm.json 2> gpurun_out/r2i_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2i_bench_1gpu_configC.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2i_bench_reference_arm.json", "gpurun_out/r2i_bench_1gpu_configC.json"):
    try:
        b = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.0f e2e %.0f ms/step %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"]), "cpu", b.get("cpu_baseline", {}).get("value"), b.get("cpu_baseline", {}).get("fe_ms_per_frame"), b.get("cpu_baseline", {}).get("be_ms_per_update"), "launches", b.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/r2i_bench.err
