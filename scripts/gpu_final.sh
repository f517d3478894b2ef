#!/bin/bash
# Round-end evidence in one gpurun call: full GPU suite, both bench arms (workload C), ncu launch list + captures, workloads D / E.
# usage: bash scripts/gpu_final.sh <tag>       (outputs land in gpurun_out/<tag>_*; copy what is to be judged into profiles/)
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench_1gpu_configC.json 2> gpurun_out/${TAG}_bench.err; echo "bench C rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err; echo "reference arm rc=$?"
KERNELS="lk_kernel be_qr_kernel be_gemm_kernel be_feature_kernel be_propagate_kernel" KEEP="lk_kernel be_qr_kernel" timeout 900 bash scripts/gpu_profile.sh ${TAG} 2>&1 | tail -12
mkdir -p gpurun_out/profiles_${TAG}; ls gpurun_out/profiles_${TAG}
timeout 500 python bench.py --workload D --cpu-frames 4 > gpurun_out/${TAG}_bench_1gpu_configD.json 2> gpurun_out/${TAG}_benchD.err; echo "bench D rc=$?"
timeout 600 python bench.py --workload E --cpu-frames 4 > gpurun_out/${TAG}_bench_1gpu_configE.json 2> gpurun_out/${TAG}_benchE.err; echo "bench E rc=$?"
python - <<PY
import json
for w in ("configC", "reference_arm", "configD", "configE"):
    f = "gpurun_out/${TAG}_bench_" + ("1gpu_" if w != "reference_arm" else "") + w + ".json"
    try:
        b = json.loads([l for l in open(f) if l.startswith("{")][-1])
        cb = b.get("cpu_baseline", {})
        print(w, "value %.0f e2e %.0f ms/step %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"]), "cpu %.0f (fe %.1f ms, be %.2f ms)" % (cb.get("value", 0), cb.get("fe_ms_per_frame", 0), cb.get("be_ms_per_update", 0)), "roofline", (b.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(w, "unreadable:", e)
PY
