"""Driver loop of the reference (app/larvioMain.cpp:87-117) for a batch of sequences, shared by tests,
smoke() and bench.py: IMU windowing (:98), processImage, processFeatures when it returned true."""
from __future__ import annotations

import numpy as np

from . import synth


class ImuFeeder:
    """Caller-side IMU buffers (one per sequence) in the layout the C ABI takes."""

    def __init__(self, seqs, stride=96):
        from . import api
        self.seqs = seqs
        self.S = len(seqs)
        self.k = [0] * self.S
        self.buf = np.zeros((self.S, stride), api.IMU_DTYPE)
        self.n = np.zeros(self.S, np.int32)

    def push_until(self, j):
        """Append every sample with t_imu - t_img[j] < 0.05 (larvioMain.cpp:98-102)."""
        for s, sq in enumerate(self.seqs):
            k2 = synth.imu_window(sq, self.k[s], sq.img_t[j])
            r = sq.imu[self.k[s]:k2]
            n, m = int(self.n[s]), len(r)
            self.buf["t"][s, n:n + m] = r[:, 0]; self.buf["gyro"][s, n:n + m] = r[:, 1:4]; self.buf["acc"][s, n:n + m] = r[:, 4:7]
            self.n[s] = n + m
            self.k[s] = k2

    def rows(self, s):
        n = int(self.n[s])
        return np.concatenate([self.buf["t"][s, :n, None], self.buf["gyro"][s, :n], self.buf["acc"][s, :n]], 1)


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
