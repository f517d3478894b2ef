"""EuRoC-shaped synthetic sequences for parity tests and the benchmark.

Not part of the hot path and not the oracle: this is the *input generator* shared by
``bench.py``, ``tests/`` and the CPU baseline so that all arms see identical bytes.
Per sequence ``s`` the seed is ``1234 + s``.  A camera with the EuRoC cam0 intrinsics,
radtan distortion and T_cam_imu moves on a smooth 6-DoF Lissajous inside a textured
box room; images are ray-cast (752x480 u8 @ 20 Hz), the IMU is sampled from the
analytic trajectory at 200 Hz with bias + white noise.  The lens model used for
rendering is the same 5-iteration radtan inverse the front end applies when it
undistorts, so scene geometry and measurements are consistent.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

GRAVITY = np.array([0.0, 0.0, -9.81])

ROOM_MIN = np.array([-5.0, -4.0, -1.6])
ROOM_MAX = np.array([5.0, 4.0, 2.0])
TEXELS_PER_M = 100.0


def _so3_exp(v: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(v)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def _so3_log(R: np.ndarray) -> np.ndarray:
    c = max(-1.0, min(1.0, (np.trace(R) - 1) / 2))
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    if th < 1e-9:
        return w
    return w * th / np.sin(th)


def rot_to_quat_xyzw(R: np.ndarray) -> np.ndarray:
    """Hamilton quaternion [x y z w] of a rotation matrix (body->world)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(1 + t) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, s / 4])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = s / 4
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def _make_textures(seed: int = 7) -> Tuple[np.ndarray, np.ndarray]:
    """Six wall textures packed in one float32 atlas; returns (atlas, per-wall [x0,y0,w,h])."""
    import cv2
    rng = np.random.default_rng(seed)
    ext = ROOM_MAX - ROOM_MIN
    # wall index = 2*axis + (positive side); its texture axes are the two other axes
    dims = []
    for a in range(3):
        o = [i for i in range(3) if i != a]
        dims.append((int(ext[o[0]] * TEXELS_PER_M), int(ext[o[1]] * TEXELS_PER_M)))
    W = max(d[0] for d in dims) + 8
    H = sum(d[1] + 8 for d in dims) * 2
    atlas = np.full((H, W), 128.0, np.float32)
    rects = np.zeros((6, 4), np.int32)
    y = 4
    for a in range(3):
        for side in range(2):
            w, h = dims[a]
            n1 = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), 1.6)
            n2 = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), 7.0)
            n3 = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), 30.0)
            tex = 128 + 34 * n1 / n1.std() + 22 * n2 / n2.std() + 18 * n3 / n3.std()
            # high-contrast rectangles: strong Shi-Tomasi corners spread over the wall
            nblob = int(w * h / 2600)
            bx = rng.integers(4, w - 20, nblob)
            by = rng.integers(4, h - 20, nblob)
            bw = rng.integers(5, 14, nblob)
            bh = rng.integers(5, 14, nblob)
            bv = rng.choice([28.0, 60.0, 200.0, 232.0], nblob)
            for i in range(nblob):
                tex[by[i]:by[i] + bh[i], bx[i]:bx[i] + bw[i]] = bv[i]
            tex = cv2.GaussianBlur(tex, (0, 0), 0.9)
            atlas[y:y + h, 4:4 + w] = tex
            # replicate a 4-texel apron so bilinear taps at wall edges stay in-wall
            atlas[y - 4:y, 4:4 + w] = tex[0:1]
            atlas[y + h:y + h + 4, 4:4 + w] = tex[-1:]
            atlas[y - 4:y + h + 4, 0:4] = atlas[y - 4:y + h + 4, 4:5]
            atlas[y - 4:y + h + 4, 4 + w:8 + w] = atlas[y - 4:y + h + 4, 3 + w:4 + w]
            rects[2 * a + side] = (4, y, w, h)
            y += h + 8
    return np.clip(atlas[:y + 4], 0, 255), rects


@dataclass
class Sequence:
    """One synthetic sequence: images (uint8 [F,H,W]), stamps, IMU rows [t,wx,wy,wz,ax,ay,az], truth."""
    images: np.ndarray
    img_t: np.ndarray
    imu: np.ndarray
    gt_t: np.ndarray
    gt_p: np.ndarray       # body position in world, at image times
    gt_q: np.ndarray       # body->world Hamilton [x y z w], at image times
    gt_v: np.ndarray
    gyro_bias: np.ndarray
    acc_bias: np.ndarray


class Trajectory:
    def __init__(self, seed: int, R_b2c: np.ndarray, static_until: float = None):
        self.static_until = static_until
        rng = np.random.default_rng(seed)
        self.amp_p = np.array([1.3, 1.1, 0.35]) * rng.uniform(0.8, 1.2, 3)
        self.f_p = np.array([0.13, 0.17, 0.23]) * rng.uniform(0.85, 1.15, 3)
        self.ph_p = rng.uniform(0, 2 * np.pi, 3)
        self.amp_r = np.array([0.10, 0.12, 0.22]) * rng.uniform(0.7, 1.2, 3)
        self.f_r = np.array([0.31, 0.23, 0.11]) * rng.uniform(0.85, 1.15, 3)
        self.ph_r = rng.uniform(0, 2 * np.pi, 3)
        yaw = rng.uniform(-np.pi, np.pi)
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        R_wc0 = Rz @ np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]])
        self.R_wb0 = R_wc0 @ R_b2c
        self.center = np.array([0.0, 0.0, 0.1]) + rng.uniform(-0.3, 0.3, 3)

    def _gate(self, t):
        """0 while the platform stands still (t < static_until), smooth ramp to 1 over the next second."""
        if self.static_until is None:
            return 1.0
        x = min(max((t - self.static_until) / 1.0, 0.0), 1.0)
        return x * x * x * (x * (6 * x - 15) + 10)

    def p(self, t):
        g = self._gate(t)
        if self.static_until is None:
            return self.center + self.amp_p * np.sin(2 * np.pi * self.f_p * t + self.ph_p)
        s0 = np.sin(2 * np.pi * self.f_p * self.static_until + self.ph_p)
        return self.center + g * self.amp_p * (np.sin(2 * np.pi * self.f_p * t + self.ph_p) - s0)

    def v(self, t):
        if self.static_until is not None:
            h = 1e-4
            return (self.p(t + h) - self.p(t - h)) / (2 * h)
        w = 2 * np.pi * self.f_p
        return self.amp_p * w * np.cos(w * t + self.ph_p)

    def a(self, t):
        if self.static_until is not None:
            h = 1e-3
            return (self.p(t + h) - 2 * self.p(t) + self.p(t - h)) / (h * h)
        w = 2 * np.pi * self.f_p
        return -self.amp_p * w * w * np.sin(w * t + self.ph_p)

    def R(self, t):
        if self.static_until is not None:
            s0 = np.sin(2 * np.pi * self.f_r * self.static_until + self.ph_r)
            th = self._gate(t) * self.amp_r * (np.sin(2 * np.pi * self.f_r * t + self.ph_r) - s0)
        else:
            th = self.amp_r * np.sin(2 * np.pi * self.f_r * t + self.ph_r)
        return self.R_wb0 @ _so3_exp(th)

    def omega_body(self, t, h=1e-5):
        return _so3_log(self.R(t - h).T @ self.R(t + h)) / (2 * h)


class Renderer:
    """Ray-casts the box room through the radtan lens; one instance per process."""

    def __init__(self, cfg_raw: dict):
        import cv2
        self.W, self.H = int(cfg_raw["resolution_width"]), int(cfg_raw["resolution_height"])
        it, dc = cfg_raw["intrinsics"], cfg_raw["distortion_coeffs"]
        K = np.array([[it["fx"], 0, it["cx"]], [0, it["fy"], it["cy"]], [0, 0, 1.0]])
        D = np.array([dc["k1"], dc["k2"], dc["p1"], dc["p2"]])
        u, v = np.meshgrid(np.arange(self.W, dtype=np.float64), np.arange(self.H, dtype=np.float64))
        pts = np.stack([u.ravel(), v.ravel()], 1).reshape(-1, 1, 2)
        if cfg_raw.get("distortion_model", "radtan") == "equidistant":
            n = cv2.fisheye.undistortPoints(pts, K, D)
        else:
            n = cv2.undistortPoints(pts, K, D)
        n = n.reshape(-1, 2)
        self.rays_c = np.concatenate([n, np.ones((n.shape[0], 1))], 1)   # [HW,3]
        T = np.array(cfg_raw["T_cam_imu"]["data"]).reshape(4, 4)
        self.R_b2c = T[:3, :3]
        self.t_c_b = -T[:3, :3].T @ T[:3, 3]
        self.atlas, self.rects = _make_textures()

    def render(self, R_wb: np.ndarray, p_wb: np.ndarray, rng=None, noise_sigma=0.8) -> np.ndarray:
        import cv2
        R_wc = R_wb @ self.R_b2c.T
        c = p_wb + R_wb @ self.t_c_b
        d = self.rays_c @ R_wc.T                                     # [HW,3] world ray dirs
        with np.errstate(divide="ignore", invalid="ignore"):
            plane = np.where(d > 0, ROOM_MAX, ROOM_MIN)
            t = (plane - c) / d
        t[~np.isfinite(t)] = np.inf
        t[t <= 0] = np.inf
        axis = np.argmin(t, 1)
        tt = t[np.arange(t.shape[0]), axis]
        hit = c + d * tt[:, None]
        wall = 2 * axis + (d[np.arange(d.shape[0]), axis] > 0)
        o0 = np.array([1, 0, 0])[axis]
        o1 = np.array([2, 2, 1])[axis]
        idx = np.arange(hit.shape[0])
        r = self.rects[wall]
        mx = r[:, 0] + (hit[idx, o0] - ROOM_MIN[o0]) * TEXELS_PER_M - 0.5
        my = r[:, 1] + (hit[idx, o1] - ROOM_MIN[o1]) * TEXELS_PER_M - 0.5
        img = cv2.remap(self.atlas, mx.astype(np.float32).reshape(self.H, self.W),
                        my.astype(np.float32).reshape(self.H, self.W), cv2.INTER_LINEAR)
        if rng is not None and noise_sigma > 0:
            img = img + rng.standard_normal(img.shape).astype(np.float32) * noise_sigma
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)


_RENDERERS = {}


def _renderer(cfg_raw):
    key = id(cfg_raw)
    if key not in _RENDERERS:
        _RENDERERS.clear()
        _RENDERERS[key] = Renderer(cfg_raw)
    return _RENDERERS[key]


def make_sequence(cfg_raw: dict, seq_index: int, n_frames: int, t0: float = 0.05,
                  imu_noise: bool = True, image_noise: float = 0.8, static_until: float = None) -> Sequence:
    """Generate sequence ``seq_index`` (seed 1234+seq_index) with ``n_frames`` images."""
    rend = _renderer(cfg_raw)
    seed = 1234 + seq_index
    rng = np.random.default_rng(seed)
    traj = Trajectory(seed, rend.R_b2c, static_until)
    img_rate, imu_rate = float(cfg_raw["img_rate"]), float(cfg_raw["imu_rate"])
    img_t = t0 + np.arange(n_frames) / img_rate
    n_imu = int(round((img_t[-1] + 0.1) * imu_rate)) + 1
    imu_t = np.arange(n_imu) / imu_rate
    bg = rng.normal(0, 2e-3, 3) if imu_noise else np.zeros(3)
    ba = rng.normal(0, 2e-2, 3) if imu_noise else np.zeros(3)
    imu = np.zeros((n_imu, 7))
    sg = 1.7e-4 * np.sqrt(imu_rate) if imu_noise else 0.0
    sa = 2.0e-3 * np.sqrt(imu_rate) if imu_noise else 0.0
    for k, t in enumerate(imu_t):
        R = traj.R(t)
        w = traj.omega_body(t) + bg + rng.normal(0, 1, 3) * sg
        f = R.T @ (traj.a(t) - GRAVITY) + ba + rng.normal(0, 1, 3) * sa
        imu[k] = (t, *w, *f)
    images = np.zeros((n_frames, rend.H, rend.W), np.uint8)
    gt_p = np.zeros((n_frames, 3)); gt_q = np.zeros((n_frames, 4)); gt_v = np.zeros((n_frames, 3))
    for j, t in enumerate(img_t):
        R = traj.R(t)
        images[j] = rend.render(R, traj.p(t), rng, image_noise)
        gt_p[j] = traj.p(t); gt_q[j] = rot_to_quat_xyzw(R); gt_v[j] = traj.v(t)
    return Sequence(images, img_t, imu, img_t.copy(), gt_p, gt_q, gt_v, bg, ba)


def imu_window(seq: Sequence, k_start: int, t_img: float) -> int:
    """Index one past the last IMU row with ``t_imu - t_img < 0.05`` (app/larvioMain.cpp:98)."""
    k = k_start
    n = seq.imu.shape[0]
    while k < n and seq.imu[k, 0] - t_img < 0.05:
        k += 1
    return k
