"""Driver-side helpers of the reference loop (app/larvioMain.cpp:87-117) for a batch of sequences, shared by tests,
smoke() and bench.py: the caller-owned IMU buffers and their windowing (:98).  Nothing here touches the oracle."""
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
