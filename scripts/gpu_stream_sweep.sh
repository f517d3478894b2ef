#!/bin/bash
# sub-batch (stream) count sweep of the bench workload; the CPU leg is skipped (--cpu-frames 0), inputs come from the bench cache
mkdir -p gpurun_out
for NS in ${SWEEP:-4 6 8 3 5}; do
  timeout 300 python bench.py --streams $NS --cpu-frames 0 --steps 90 > gpurun_out/sweep_s$NS.json 2> gpurun_out/sweep_s$NS.err
  python - <<PY
import json
try:
    b = json.loads([l for l in open("gpurun_out/sweep_s$NS.json") if l.startswith("{")][-1])
    print("streams $NS: value %.0f e2e %.0f ms/step %.3f" % (b["value"], b["e2e"]["value"], b["ms_per_step"]))
except Exception as e:
    print("streams $NS: unreadable", e)
PY
done
