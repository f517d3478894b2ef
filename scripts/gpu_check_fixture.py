"""Diagnostic (GPU box): replay reference fixtures through the CUDA filter and print, per processFeatures call, the state
dimension of both sides and the deviations - without stopping at the first difference.
usage: python scripts/gpu_check_fixture.py <case> [<case> ...]   (cases: tests/golden/ref_<case>.npz)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_runner as rr          # noqa: E402
import test_gpu as tg            # noqa: E402


def table(run, ref):
    first_bad = None
    for i, (x, y) in enumerate(zip(run, ref)):
        if not y["ok"] or not x["ok"]:
            if x["ok"] != y["ok"]:
                print("call %3d: ok %s vs reference %s" % (i, x["ok"], y["ok"]))
            continue
        d = y["dim"]
        same = x["P"].shape[0] == d
        pz = float(np.linalg.norm(x["P"] @ rr.fingerprint_vector(d) - y["Pz"]) / np.linalg.norm(y["Pz"])) if same else -1.0
        dp = float(np.abs(x["p"] - y["p"]).max())
        bad = (not same) or pz > 1e-8 or dp > 1e-8
        if bad and first_bad is None:
            first_bad = i
        if bad or i % 10 == 0:
            print("call %3d: dim %3d / %3d  win %2d / %2d  slam ref %2d  nui ref %d  Pz %.2e  p %.2e  stable %s/%s active %d/%d%s" % (
                i, x["P"].shape[0], d, x["n_win"], y["n_win"], len(y["slam_ids"]), y["n_nui"], pz, dp, sorted(x["stable"]), sorted(y["stable"]),
                len(x["active"]), len(y["active"]), "   <-- differs" if bad else ""))
        if first_bad is not None and i > first_bad + 3:
            break
    print("first differing call:", first_bad)


if __name__ == "__main__":
    orig = rr.compare_with_fixture
    for name in sys.argv[1:]:
        print("==== %s" % name)
        grabbed = {}

        def grab(run, ref):
            grabbed["run"], grabbed["ref"] = run, ref
            return dict(n=0, q=0, p=0, v=0, bg=0, ba=0, ext=0, td=0, Pz=0, Pdiag=0, P=0, pts=0, n_pts=0)
        rr.compare_with_fixture = grab
        try:
            tg._drive_fixture(name)
            table(grabbed["run"], grabbed["ref"])
        except Exception as e:        # a capacity / CUDA error of the library: report and go on
            print("FAILED:", repr(e)[:600])
        finally:
            rr.compare_with_fixture = orig
