"""Restatement of cv::findFundamentalMat(FM_RANSAC, 1.0, 0.99) (OpenCV 4.x calib3d:
ptsetreg.cpp RANSACPointSetRegistrator::run/getSubset, fundam.cpp FMEstimatorCallback) —
TEST INFRASTRUCTURE.  OpenCV is not vendored under /root/reference; this file
restates its published algorithm and is pinned against cv2 4.13 by tests/test_oracle_ransac.py
(inlier masks identical on randomized inputs).  The CUDA kernel follows this restatement.
Call sites in the reference: image_processor.cpp:498-500, 755-757, 968-970.
"""
import numpy as np

RNG_COEFF = 4164903690


class CvRNG:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _collinear(pts, count):
    """haveCollinearPoints(): only the last point is tested against all earlier pairs."""
    i = count - 1
    for j in range(i):
        dx1 = float(pts[j][0]) - float(pts[i][0]); dy1 = float(pts[j][1]) - float(pts[i][1])
        for k in range(j):
            dx2 = float(pts[k][0]) - float(pts[i][0]); dy2 = float(pts[k][1]) - float(pts[i][1])
            if abs(dx2 * dy1 - dy2 * dx1) <= 1.1920929e-07 * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def get_subset(m1, m2, rng, max_attempts=10000, model_points=7):
    count = len(m1)
    for _ in range(max_attempts):
        idx = []
        for i in range(model_points):
            v = rng.uniform(0, count)
            while v in idx:
                v = rng.uniform(0, count)
            idx.append(v)
        if not _collinear(m1[idx], model_points) and not _collinear(m2[idx], model_points):
            return idx
    return None


def solve_cubic(c):
    a0, a1, a2, a3 = (float(x) for x in c)
    if a0 == 0:
        if a1 == 0:
            if a2 == 0:
                return []
            return [-a3 / a2]
        d = a2 * a2 - 4 * a1 * a3
        if d >= 0:
            d = np.sqrt(d)
            q1 = (-a2 + d) * 0.5; q2 = (a2 + d) * -0.5
            if abs(q1) > abs(q2):
                x0 = q1 / a1; x1 = a3 / q1
            else:
                x0 = q2 / a1; x1 = a3 / q2
            return [x0, x1] if d > 0 else [x0]
        return []
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0
    Q = (a1 * a1 - 3 * a2) * (1. / 9)
    R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54)
    Qc = Q * Q * Q
    d = Qc - R * R
    if d > 0:
        theta = np.arccos(R / np.sqrt(Qc)); sq = np.sqrt(Q)
        t0 = -2 * sq; t1 = theta * (1. / 3); t2 = a1 * (1. / 3)
        return [t0 * np.cos(t1) - t2, t0 * np.cos(t1 + 2. * np.pi / 3) - t2, t0 * np.cos(t1 + 4. * np.pi / 3) - t2]
    if d == 0:
        if R >= 0:
            x0 = -2 * R ** (1. / 3) - a1 / 3; x1 = R ** (1. / 3) - a1 / 3
        else:
            x0 = 2 * (-R) ** (1. / 3) - a1 / 3; x1 = -(-R) ** (1. / 3) - a1 / 3
        return [x0] if x0 == x1 else [x0, x1]
    d = np.sqrt(-d)
    e = (d + abs(R)) ** (1. / 3)
    if R > 0:
        e = -e
    return [(e + Q / e) - a1 * (1. / 3)]


# OpenCV's JacobiSVD (core/src/lapack.cpp JacobiSVDImpl_) fills the two null-space rows of Vt from
# +-1/9 sign vectors drawn from cv::RNG(0x12345678), Gram-Schmidt'ed against the singular vectors:
# row 7 = unit projection of R1 onto null(A), row 8 = unit projection of R2 made orthogonal to row 7.
# The order of the (up to three) candidate F matrices — which decides ties between equally good
# models of one sample — depends on this basis, so it is part of the restatement.
_NULL_R1 = np.array([-1, -1, 1, -1, -1, -1, -1, 1, 1], np.float64) / 9.0
_NULL_R2 = np.array([1, -1, 1, 1, 1, 1, 1, -1, 1], np.float64) / 9.0


def opencv_null_basis(n1, n2):
    """n1, n2: any basis of the 2-D null space -> (Vt[7], Vt[8]) as cv::SVDecomp(FULL_UV) returns them."""
    e1 = n1 / np.linalg.norm(n1)
    e2 = n2 - (e1 @ n2) * e1
    e2 = e2 / np.linalg.norm(e2)
    p1 = (e1 @ _NULL_R1) * e1 + (e2 @ _NULL_R1) * e2
    f1 = p1 / np.linalg.norm(p1)
    p2 = (e1 @ _NULL_R2) * e1 + (e2 @ _NULL_R2) * e2
    p2 = p2 - (f1 @ p2) * f1
    f2 = p2 / np.linalg.norm(p2)
    return f1, f2


def run_7point(m1, m2):
    """calib3d/src/fundam.cpp run7Point as shipped in the OpenCV 4.13 of this container: the seven pairs are
    Hartley-normalised (centroid to the origin, mean distance sqrt(2)) before the 7x9 system is formed, and every
    candidate F is mapped back with T2^T F T1 and rescaled to F(3,3) = 1.  (OpenCV <= 4.5 solved the raw-pixel
    system; the candidate SET is the same, their ORDER - which decides ties - is not.)"""
    m1d = np.asarray(m1, np.float32).astype(np.float64); m2d = np.asarray(m2, np.float32).astype(np.float64)
    m1c = np.zeros(2); m2c = np.zeros(2)
    for i in range(7):
        m1c = m1c + m1d[i]; m2c = m2c + m2d[i]
    t = 1. / 7
    m1c = m1c * t; m2c = m2c * t
    scale1 = 0.0; scale2 = 0.0
    for i in range(7):
        scale1 += np.sqrt((m1d[i][0] - m1c[0]) * (m1d[i][0] - m1c[0]) + (m1d[i][1] - m1c[1]) * (m1d[i][1] - m1c[1]))
        scale2 += np.sqrt((m2d[i][0] - m2c[0]) * (m2d[i][0] - m2c[0]) + (m2d[i][1] - m2c[1]) * (m2d[i][1] - m2c[1]))
    scale1 *= t; scale2 *= t
    if scale1 < 1.1920928955078125e-07 or scale2 < 1.1920928955078125e-07:
        return []
    scale1 = np.sqrt(2.) / scale1; scale2 = np.sqrt(2.) / scale2
    A = np.zeros((7, 9))
    for i in range(7):
        x0 = (m1d[i][0] - m1c[0]) * scale1; y0 = (m1d[i][1] - m1c[1]) * scale1
        x1 = (m2d[i][0] - m2c[0]) * scale2; y1 = (m2d[i][1] - m2c[1]) * scale2
        A[i] = [x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1]
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    f1, f2 = opencv_null_basis(Vt[7], Vt[8])
    f1 = f1 - f2
    t0 = f2[4] * f2[8] - f2[5] * f2[7]; t1 = f2[3] * f2[8] - f2[5] * f2[6]; t2 = f2[3] * f2[7] - f2[4] * f2[6]
    c = np.zeros(4)
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2
    c[2] = (f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
            f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
            f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
            f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]))
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6]
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2
    c[1] = (f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
            f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
            f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
            f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]))
    roots = solve_cubic(c)
    Fs = []
    for r in roots:
        lam, mu = r, 1.0
        s = f1[8] * r + f2[8]
        F = np.zeros(9)
        if abs(s) > 2.220446049250313e-16:
            mu = 1. / s; lam *= mu; F[8] = 1.0
        else:
            F[8] = 0.0
        F[:8] = f1[:8] * lam + f2[:8] * mu
        # de-normalise: T2^T F T1 with T = [[s, 0, -s cx], [0, s, -s cy], [0, 0, 1]], then F(3,3) = 1
        M = np.zeros(9)
        for j in range(3):
            M[j] = scale2 * F[j]; M[3 + j] = scale2 * F[3 + j]
            M[6 + j] = (-scale2 * m2c[0]) * F[j] + (-scale2 * m2c[1]) * F[3 + j] + F[6 + j]
        G = np.zeros(9)
        for i in range(3):
            G[3 * i] = M[3 * i] * scale1; G[3 * i + 1] = M[3 * i + 1] * scale1
            G[3 * i + 2] = M[3 * i] * (-scale1 * m1c[0]) + M[3 * i + 1] * (-scale1 * m1c[1]) + M[3 * i + 2]
        if abs(G[8]) > 1.1920928955078125e-07:
            G = G * (1. / G[8])
        Fs.append(G)
    return Fs


def compute_error(m1, m2, F):
    x1 = m1[:, 0].astype(np.float64); y1 = m1[:, 1].astype(np.float64)
    x2 = m2[:, 0].astype(np.float64); y2 = m2[:, 1].astype(np.float64)
    a = F[0] * x1 + F[1] * y1 + F[2]; b = F[3] * x1 + F[4] * y1 + F[5]; c = F[6] * x1 + F[7] * y1 + F[8]
    s2 = 1. / (a * a + b * b); d2 = x2 * a + y2 * b + c
    a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8]
    s1 = 1. / (a * a + b * b); d1 = x1 * a + y1 * b + c
    return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(np.float32)


def update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.), 1.); ep = min(max(ep, 0.), 1.)
    num = max(1. - p, 2.2250738585072014e-308)
    denom = 1. - (1. - ep) ** model_points
    if denom < 2.2250738585072014e-308:
        return 0
    num = np.log(num); denom = np.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def _lmeds_mask(m1, m2, confidence, max_iters):
    """LMeDSPointSetRegistrator::run — what findFundamentalMat(FM_RANSAC) uses for 8..14 points."""
    count = len(m1)
    rng = CvRNG()
    niters = max(update_num_iters(confidence, 0.45, 7, max_iters), 3)
    min_median = np.inf
    best = None
    for it in range(niters):
        idx = get_subset(m1, m2, rng, max_attempts=1000)
        if idx is None:
            if it == 0:
                return None
            break
        for F in run_7point(m1[idx], m2[idx]):
            err = compute_error(m1, m2, F)
            med = float(np.sort(err)[count // 2])
            if med < min_median:
                min_median = med; best = F
    if best is None:
        return None
    sigma = 2.5 * 1.4826 * (1 + 5. / (count - 7)) * np.sqrt(min_median)
    sigma = max(sigma, 0.001)
    err = compute_error(m1, m2, best)
    return (err <= np.float32(sigma * sigma)).astype(np.uint8)


def find_fundamental_ransac_mask(p1, p2, threshold=1.0, confidence=0.99, max_iters=1000):
    """Returns the inlier mask (uint8) or None exactly where cv2 returns mask None."""
    m1 = np.asarray(p1, np.float32).reshape(-1, 2); m2 = np.asarray(p2, np.float32).reshape(-1, 2)
    count = len(m1)
    if count < 7:
        return None
    if count == 7:
        return np.ones(7, np.uint8) if len(run_7point(m1, m2)) > 0 else None
    if count < 15:
        return _lmeds_mask(m1, m2, confidence, max_iters)
    rng = CvRNG()
    niters = max(max_iters, 1)
    best_mask = None; max_good = 0
    thr2 = np.float32(threshold * threshold)
    it = 0
    while it < niters:
        idx = get_subset(m1, m2, rng)
        if idx is None:
            if it == 0:
                return None
            break
        for F in run_7point(m1[idx], m2[idx]):
            err = compute_error(m1, m2, F)
            mask = (err <= thr2).astype(np.uint8)
            good = int(mask.sum())
            if good > max(max_good, 6):
                best_mask = mask; max_good = good
                niters = update_num_iters(confidence, (count - good) / count, 7, niters)
        it += 1
    return best_mask
