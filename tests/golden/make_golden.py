"""Generate tests/golden/*.npz from the CPU oracle (run in the build container; committed with its output).

The reference ships no golden vectors; these fixtures pin (a) the synthetic generator,
(b) the front-end oracle's feature messages and (c) the back-end oracle's filter states so that the GPU
box — which has neither /root/reference nor necessarily the same CPU ISA — checks against bytes produced
here.  Usage: python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from larvio_b200.config import Config          # noqa: E402
from larvio_b200 import synth                  # noqa: E402
from oracle_runner import run_oracle           # noqa: E402

NF = 14
cfg = Config.load(os.path.join(ROOT, "configs", "euroc_mono.yaml"), max_features_in_one_grid=0, sw_size=12)
out = {}
for s in range(2):
    seq = synth.make_sequence(cfg.raw, s, NF)
    out["img_sha_%d" % s] = np.frombuffer(hashlib.sha256(seq.images.tobytes()).digest(), np.uint8)
    out["imu_%d" % s] = seq.imu
    recs = run_oracle(cfg.raw, seq, NF)
    for r in recs:
        j = r["frame"]
        if r["msg"] is not None:
            out["ids_%d_%d" % (s, j)] = r["msg"].ids
            out["data_%d_%d" % (s, j)] = r["msg"].data
        if r["ok"]:
            out["state_%d_%d" % (s, j)] = np.concatenate([r["q"], r["p"], r["v"], r["bg"], r["ba"]])
            out["Pdiag_%d_%d" % (s, j)] = np.diag(r["P"]).copy()
            out["Pfro_%d_%d" % (s, j)] = np.array([np.linalg.norm(r["P"]), r["P"].shape[0]])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_seq.npz"), **out)
print("wrote", len(out), "arrays")
