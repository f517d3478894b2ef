#!/bin/bash
# Round-2 GPU call G (2 GPUs): smoke(), lvbm_* over two real devices, and the 2-GPU bench line under torchrun.
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python - <<'PY' 2>&1 | tail -3
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from larvio_b200.config import Config
from larvio_b200 import synth, api, harness
cfg = Config.load('configs/euroc_mono.yaml', max_features_in_one_grid=0, sw_size=12)
sq = [synth.make_sequence(cfg.raw, s % 2, 12) for s in range(4)]
one = api.Batch(cfg, n_seq=4); two = api.MultiBatch(cfg, 4, [0, 1])
f1 = harness.ImuFeeder(sq); f2 = harness.ImuFeeder(sq)
for s in range(4):
    a = (sq[s].img_t[0], sq[s].gt_q[0], sq[s].gt_p[0], sq[s].gt_v[0], np.zeros(3), np.zeros(3))
    one.set_initial_state(s, *a); two.set_initial_state(s, *a)
for j in range(12):
    f1.push_until(j); f2.push_until(j)
    imgs = np.stack([sq[s].images[j] for s in range(4)]); t = np.array([sq[s].img_t[j] for s in range(4)])
    p1 = one.step(imgs, t, f1.buf, f1.n); p2 = two.step(imgs, t, f2.buf, f2.n)
    assert np.array_equal(p1, p2)
print("lvbm over 2 GPUs bit-identical to one handle:", bool(np.array_equal(one.get_states(), two.get_states())))
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2g_bench_2gpu.log 2>&1; echo "bench 2gpu rc=$?"
tail -1 gpurun_out/r2g_bench_2gpu.log > gpurun_out/r2g_bench_2gpu.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2g_bench_2gpu.json"))
    print("2 GPUs: value %.0f e2e %.0f ms/step %.3f n_gpus %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["n_gpus"]))
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/r2g_bench_2gpu.log | cut -c1-300
