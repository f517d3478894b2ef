"""numpy restatement of the reference's steered-BRIEF descriptor at arbitrary points.

Follows /root/reference/src/ORBDescriptor.cpp: ctor + umax :287-332, computeOrbDescriptor
:334-382, computeDescriptors :385-416, initializeLayerAndPyramid :418-483 (level 0 only —
level 1 is never sampled), IC_Angle :486-513; Hamming distance
include/ORB/ORBDescriptor.h:44-60.  OpenCV pieces (copyMakeBorder, GaussianBlur,
fastAtan2) are the real cv2 4.13 functions.
"""
import os
import numpy as np
import cv2

HALF_PATCH = 15
BORDER = 32          # max(edgeThreshold=31, ceil(15*sqrt2)=22, 4) + 1   (ORBDescriptor.cpp:419-421)
_PATTERN = np.load(os.path.join(os.path.dirname(__file__), "orb_pattern.npy")).astype(np.int32)
FACTOR_PI = np.float32(np.pi / np.float32(180.0))   # (float)(CV_PI/180.f)


def _umax():
    # ORBDescriptor.cpp:316-329
    u = [0] * (HALF_PATCH + 1)
    vmax = int(np.floor(HALF_PATCH * np.sqrt(np.float32(2.0)) / 2 + 1))
    vmin = int(np.ceil(HALF_PATCH * np.sqrt(np.float32(2.0)) / 2))
    hp2 = float(HALF_PATCH * HALF_PATCH)
    for v in range(vmax + 1):
        u[v] = int(np.rint(np.sqrt(hp2 - v * v)))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while u[v0] == u[v0 + 1]:
            v0 += 1
        u[v] = v0
        v0 += 1
    return u


UMAX = _umax()
# offsets of the radius-15 disc (rows v = -15..15, |u| <= umax[|v|])
_DISC_V = np.concatenate([np.full(2 * UMAX[abs(v)] + 1, v) for v in range(-HALF_PATCH, HALF_PATCH + 1)]).astype(np.int64)
_DISC_U = np.concatenate([np.arange(-UMAX[abs(v)], UMAX[abs(v)] + 1) for v in range(-HALF_PATCH, HALF_PATCH + 1)]).astype(np.int64)


# The compiled twin (oracle/orb_c.cpp): bench.py's CPU legs switch it on with ``use_compiled(True)`` so that they time compiled
# code; the numpy path below stays the default and is what the parity tests use (the two are pinned bit for bit in tests/test_cpu.py).
_C = None
_USE_C = False
_UMAX32 = None
_PATTERN32 = None


def use_compiled(on: bool = True):
    """Route ``OrbOracle.compute`` / ``hamming_rows`` through oracle/_build/liboracle_orb.so (built on demand with g++)."""
    global _C, _USE_C, _UMAX32, _PATTERN32
    if on and _C is None:
        import ctypes
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        lib = os.path.join(here, "_build", "liboracle_orb.so"); src = os.path.join(here, "orb_c.cpp")
        if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(lib), exist_ok=True)
            subprocess.run(["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", lib, src], check=True)
        _C = ctypes.CDLL(lib)
        _C.orbc_compute.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _C.orbc_hamming_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _UMAX32 = np.array(UMAX, np.int32); _PATTERN32 = np.ascontiguousarray(_PATTERN, np.int32)
    _USE_C = bool(on)


class OrbOracle:
    """One instance per image, like ``new ORBdescriptor(curr_pyramid_[0], 2, levels)``."""

    def __init__(self, image: np.ndarray):
        self.raw = cv2.copyMakeBorder(image, BORDER, BORDER, BORDER, BORDER, cv2.BORDER_REFLECT_101)
        self.blur = self.raw.copy()
        # blur applied to the interior ROI only; the 32-px frame stays unblurred (App. C-3)
        self.blur[BORDER:-BORDER, BORDER:-BORDER] = cv2.GaussianBlur(
            image, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)

    def ic_angle(self, pt) -> np.float32:
        cx = int(np.rint(np.float32(pt[0]))) + BORDER
        cy = int(np.rint(np.float32(pt[1]))) + BORDER
        img = self.raw.astype(np.int64)
        m10 = 0
        m01 = 0
        us = np.arange(-HALF_PATCH, HALF_PATCH + 1)
        m10 += int((us * img[cy, cx - HALF_PATCH:cx + HALF_PATCH + 1]).sum())
        for v in range(1, HALF_PATCH + 1):
            d = UMAX[v]
            u = np.arange(-d, d + 1)
            plus = img[cy + v, cx - d:cx + d + 1]
            minus = img[cy - v, cx - d:cx + d + 1]
            m01 += v * int((plus - minus).sum())
            m10 += int((u * (plus + minus)).sum())
        return np.float32(cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10))))

    def descriptor(self, pt, angle_deg: np.float32) -> np.ndarray:
        ang = np.float32(angle_deg) * FACTOR_PI
        a = np.float32(np.cos(np.float64(ang)))
        b = np.float32(np.sin(np.float64(ang)))
        cx = int(np.rint(np.float32(pt[0]))) + BORDER
        cy = int(np.rint(np.float32(pt[1]))) + BORDER
        px = _PATTERN[:, [0, 2]].astype(np.float32)   # [256,2] x of both taps
        py = _PATTERN[:, [1, 3]].astype(np.float32)
        x = px * a - py * b                           # float32, separately rounded products
        y = px * b + py * a
        ix = np.rint(x).astype(np.int32)
        iy = np.rint(y).astype(np.int32)
        val = self.blur[cy + iy, cx + ix].astype(np.int32)
        bits = (val[:, 0] < val[:, 1]).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    def compute_loop(self, pts: np.ndarray) -> np.ndarray:
        """Literal per-point restatement (slow); ``compute`` is the vectorised equivalent."""
        out = np.zeros((len(pts), 32), np.uint8)
        for i, p in enumerate(pts):
            out[i] = self.descriptor(p, self.ic_angle(p))
        return out

    def angles(self, pts: np.ndarray) -> np.ndarray:
        pts = np.asarray(pts, np.float32).reshape(-1, 2)
        cx = np.rint(pts[:, 0]).astype(np.int64) + BORDER
        cy = np.rint(pts[:, 1]).astype(np.int64) + BORDER
        val = self.raw[cy[:, None] + _DISC_V[None, :], cx[:, None] + _DISC_U[None, :]].astype(np.int64)
        m10 = (val * _DISC_U[None, :]).sum(1)
        m01 = (val * _DISC_V[None, :]).sum(1)
        y = m01.astype(np.float32); x = m10.astype(np.float32)
        return np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)

    def compute(self, pts: np.ndarray) -> np.ndarray:
        pts = np.asarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        if n == 0:
            return np.zeros((0, 32), np.uint8)
        if _USE_C:
            pts = np.ascontiguousarray(pts); out = np.zeros((n, 32), np.uint8)
            assert self.raw.flags.c_contiguous and self.blur.flags.c_contiguous
            _C.orbc_compute(self.raw.ctypes.data, self.blur.ctypes.data, self.raw.shape[1], BORDER, pts.ctypes.data, n,
                            _PATTERN32.ctypes.data, _UMAX32.ctypes.data, out.ctypes.data, None)
            return out
        ang = (self.angles(pts) * FACTOR_PI).astype(np.float32)
        a = np.cos(ang.astype(np.float64)).astype(np.float32)[:, None, None]
        b = np.sin(ang.astype(np.float64)).astype(np.float32)[:, None, None]
        cx = (np.rint(pts[:, 0]).astype(np.int64) + BORDER)[:, None, None]
        cy = (np.rint(pts[:, 1]).astype(np.int64) + BORDER)[:, None, None]
        px = _PATTERN[:, [0, 2]].astype(np.float32)[None]
        py = _PATTERN[:, [1, 3]].astype(np.float32)[None]
        x = (px * a).astype(np.float32) - (py * b).astype(np.float32)
        y = (px * b).astype(np.float32) + (py * a).astype(np.float32)
        ix = np.rint(x).astype(np.int64); iy = np.rint(y).astype(np.int64)
        val = self.blur[cy + iy, cx + ix].astype(np.int32)
        bits = (val[:, :, 0] < val[:, :, 1]).astype(np.uint8)
        return np.packbits(bits.reshape(n, 32, 8), axis=2, bitorder="little").reshape(n, 32)


def hamming(a: np.ndarray, b: np.ndarray) -> int:
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def hamming_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    if _USE_C and len(a):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8); out = np.zeros(len(a), np.int32)
        _C.orbc_hamming_rows(a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data)
        return out
    return np.unpackbits(np.bitwise_xor(a, b), axis=1).sum(1).astype(np.int32)
