#!/bin/bash
# A/B of the staged kernel variants in ONE gpurun call: parity gate first, then a short bench per variant.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_ab_variants.sh'
# Output: gpurun_out/ab_<variant>.json (bench JSON lines) and gpurun_out/ab_summary.txt.
set -u
VARIANTS=${VARIANTS:-"none chol_blocked qr_lean gemm_dmma trsm_wide graph chol_blocked,qr_lean,gemm_dmma,trsm_wide chol_blocked,qr_lean,gemm_dmma,trsm_wide,graph"}
STEPS=${STEPS:-60}
STREAMS=${STREAMS:-4}
mkdir -p gpurun_out
if [ -z "${SKIP_PARITY:-}" ]; then
  python -m pytest tests/test_gpu.py -q -k "staged_kernel_variants" -rA 2>&1 | tail -15 > gpurun_out/ab_parity.txt
  cat gpurun_out/ab_parity.txt
fi
: > gpurun_out/ab_summary.txt
for V in $VARIANTS; do
  if [ "$V" = none ]; then unset LVB_EXPERIMENT; else export LVB_EXPERIMENT=$V; fi
  TAG=$(echo $V | tr ',' '+')
  python bench.py --steps $STEPS --warmup 6 --cpu-frames 0 --streams $STREAMS > gpurun_out/ab_$TAG.log 2>&1
  tail -1 gpurun_out/ab_$TAG.log > gpurun_out/ab_$TAG.json
  python - "$TAG" <<'PY' | tee -a gpurun_out/ab_summary.txt
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % tag))
    k = d["kernels"]
    pick = {n: round(v["ms_per_launch"] * 1e3, 1) for n, v in k.items() if any(x in n for x in ("chol", "qr", "gemm", "trsm", "lk_"))}
    pick["d"] = round(d["steady_state"]["state_dim_mean"], 1)
    print("%-60s value %8.0f  e2e %8.0f  ms/step %.3f  us/launch %s" % (tag, d["value"], d["e2e"]["value"], d["ms_per_step"], pick))
except Exception as e:
    print("%-60s FAILED (%s)" % (tag, e))
PY
done
# graph mode is about host overhead: see whether more sub-batches scale with it
for NS in 8; do
  LVB_EXPERIMENT=graph python bench.py --steps $STEPS --warmup 6 --cpu-frames 0 --streams $NS > gpurun_out/ab_graph_s$NS.log 2>&1
  tail -1 gpurun_out/ab_graph_s$NS.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph streams=$NS value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))" | tee -a gpurun_out/ab_summary.txt
done
