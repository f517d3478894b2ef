"""The CPU oracle driven like the reference driver drives LARVIO (app/larvioMain.cpp:87-117): IMU windowing, processImage,
processFeatures when it returned true.  TEST INFRASTRUCTURE (tests/, tests/golden/make_golden.py) - never imported by the package."""
import numpy as np

from larvio_b200 import synth


def run_oracle(cfg_raw, seq, n_frames, init_from_truth=True):
    """CPU oracle over one sequence. Returns per-frame dicts (msg ids/data, state after processFeatures)."""
    from oracle.frontend import ImageProcessorOracle
    from oracle.backend import LarVioOracle
    fe = ImageProcessorOracle(cfg_raw); be = LarVioOracle(cfg_raw)
    imu = []; k = 0; out = []
    for j in range(n_frames):
        k2 = synth.imu_window(seq, k, seq.img_t[j]); imu.extend(seq.imu[k:k2].tolist()); k = k2
        msg = fe.process_image(seq.images[j], seq.img_t[j], np.array(imu).reshape(-1, 7))
        rec = dict(frame=j, msg=msg, ok=False)
        if msg is not None:
            if init_from_truth and not be.is_gravity_set:
                be.set_initial_state(seq.img_t[j], seq.gt_q[j], seq.gt_p[j], seq.gt_v[j], np.zeros(3), np.zeros(3))
            rec["ok"] = be.process_features(msg, imu)
            if rec["ok"]:
                s = be.imu_state
                rec.update(q=s.q.copy(), p=s.p.copy(), v=s.v.copy(), bg=s.bg.copy(), ba=s.ba.copy(), P=be.P.copy(), n_win=len(be.aug),
                           t=float(seq.img_t[j]), n_slam=len(getattr(be, "feature_states", [])), dim=be.P.shape[0],
                           pos_err=float(np.linalg.norm(s.p - seq.gt_p[j])))
        out.append(rec)
    return out
