"""Scalar restatement of cv::calcOpticalFlowPyrLK (OpenCV 4.x video/src/lkpyramid.cpp, LKTrackerInvoker)
that replays the SSE (CV_SIMD128) accumulation order so results are BIT-identical to cv2 on x86 —
TEST INFRASTRUCTURE.  The CUDA kernel fe_lk.cu follows the same order.  Pinned by tests/test_cpu.py.

Accumulation order (per level):
  A11/A12/A22: four float lanes; lane j adds, row by row, pixels x = j, 4+j, 8+j, 12+j; pixels 16..20 go to a
               scalar tail accumulator; total = tail + ((l0 + l2) + (l1 + l3)).
  b1/b2:       per row and per 8-pixel chunk c (x0 = 8c) four accumulators receive
               float(int(d[x0+p]*g[x0+p] + d[x0+4+p]*g[x0+4+p])), p = 0..3; scalar tail for 16..20;
               total = tail + ((q0 + q2) + (q1 + q3)).

Status: this follows the call the REFERENCE makes - err = noArray() (image_processor.cpp:368-377).  The Python binding always
requests `err`, and with err present OpenCV re-checks the converged point at level 0 and clears status when it lies more
than a window outside the image.  A 45 360-point campaign (corners, random and border points, 0.3 / 2 / 8 px initial
errors, 36 image pairs) gives bit-identical positions wherever both report success and exactly that status difference on
5 points, all >= 10 px outside the image - where the reference's own in-image gate (:571-578) or the 1-px
forward/backward gate (:630-642) discards the point in either case, so tracks and ids are unaffected.
"""
import numpy as np
import numba as nb

f32 = np.float32
W_BITS = 14


@nb.njit(cache=True)
def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


@nb.njit(cache=True)
def _lk_point(levelsA, derivA, levelsB, prev_pt, next_pt, max_level, max_iter, eps2, min_eig_thr):
    win = 21
    half = f32(10.0)
    status = True
    nx = next_pt[0]; ny = next_pt[1]
    Iw = np.zeros((win, win), np.int32); Ixw = np.zeros((win, win), np.int32); Iyw = np.zeros((win, win), np.int32)
    for level in range(max_level, -1, -1):
        I = levelsA[level]; dI = derivA[level]; J = levelsB[level]
        rows = I.shape[0] - 2 * 24; cols = I.shape[1] - 2 * 24     # padded by 24
        scale = f32(1.0) / f32(1 << level)
        px = f32(prev_pt[0] * scale); py = f32(prev_pt[1] * scale)
        if level == max_level:
            nx = f32(nx * scale); ny = f32(ny * scale)
        else:
            nx = f32(nx * f32(2.0)); ny = f32(ny * f32(2.0))
        px = f32(px - half); py = f32(py - half)
        ipx = int(np.floor(px)); ipy = int(np.floor(py))
        if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
            if level == 0:
                status = False
            continue
        a = f32(px - f32(ipx)); b = f32(py - f32(ipy))
        iw00 = int(np.rint(f32(f32(f32(1.0) - a) * f32(f32(1.0) - b)) * f32(1 << W_BITS)))
        iw01 = int(np.rint(f32(a * f32(f32(1.0) - b)) * f32(1 << W_BITS)))
        iw10 = int(np.rint(f32(f32(f32(1.0) - a) * b) * f32(1 << W_BITS)))
        iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
        lA11 = np.zeros(4, f32); lA12 = np.zeros(4, f32); lA22 = np.zeros(4, f32)
        tA11 = f32(0.0); tA12 = f32(0.0); tA22 = f32(0.0)
        for y in range(win):
            for x in range(win):
                r = 24 + ipy + y; c = 24 + ipx + x
                ival = _descale(int(I[r, c]) * iw00 + int(I[r, c + 1]) * iw01 + int(I[r + 1, c]) * iw10 + int(I[r + 1, c + 1]) * iw11, W_BITS - 5)
                ixv = _descale(int(dI[r, c, 0]) * iw00 + int(dI[r, c + 1, 0]) * iw01 + int(dI[r + 1, c, 0]) * iw10 + int(dI[r + 1, c + 1, 0]) * iw11, W_BITS)
                iyv = _descale(int(dI[r, c, 1]) * iw00 + int(dI[r, c + 1, 1]) * iw01 + int(dI[r + 1, c, 1]) * iw10 + int(dI[r + 1, c + 1, 1]) * iw11, W_BITS)
                Iw[y, x] = ival; Ixw[y, x] = ixv; Iyw[y, x] = iyv
                if x < 16:
                    j = x & 3
                    lA11[j] = f32(f32(f32(ixv) * f32(ixv)) + lA11[j])
                    lA12[j] = f32(f32(f32(ixv) * f32(iyv)) + lA12[j])
                    lA22[j] = f32(f32(f32(iyv) * f32(iyv)) + lA22[j])
                else:
                    tA11 = f32(tA11 + f32(ixv * ixv)); tA12 = f32(tA12 + f32(ixv * iyv)); tA22 = f32(tA22 + f32(iyv * iyv))
        iA11 = f32(tA11 + f32(f32(lA11[0] + lA11[2]) + f32(lA11[1] + lA11[3])))
        iA12 = f32(tA12 + f32(f32(lA12[0] + lA12[2]) + f32(lA12[1] + lA12[3])))
        iA22 = f32(tA22 + f32(f32(lA22[0] + lA22[2]) + f32(lA22[1] + lA22[3])))
        FLT_SCALE = f32(1.0) / f32(1 << 20)
        A11 = f32(iA11 * FLT_SCALE); A12 = f32(iA12 * FLT_SCALE); A22 = f32(iA22 * FLT_SCALE)
        D = f32(f32(A11 * A22) - f32(A12 * A12))
        dif = f32(A11 - A22)
        rad = f32(f32(dif * dif) + f32(f32(f32(4.0) * A12) * A12))
        min_eig = f32(f32(f32(A22 + A11) - f32(np.sqrt(rad))) / f32(2 * win * win))
        if np.float64(min_eig) < min_eig_thr or D < f32(1.1920929e-07):
            if level == 0:
                status = False
            continue
        D = f32(f32(1.0) / D)
        npx = f32(nx - half); npy = f32(ny - half)
        pdx = f32(0.0); pdy = f32(0.0)
        for j in range(max_iter):
            inx = int(np.floor(npx)); iny = int(np.floor(npy))
            if inx < -win or inx >= cols or iny < -win or iny >= rows:
                if level == 0:
                    status = False
                break
            a = f32(npx - f32(inx)); b = f32(npy - f32(iny))
            iw00 = int(np.rint(f32(f32(f32(1.0) - a) * f32(f32(1.0) - b)) * f32(1 << W_BITS)))
            iw01 = int(np.rint(f32(a * f32(f32(1.0) - b)) * f32(1 << W_BITS)))
            iw10 = int(np.rint(f32(f32(f32(1.0) - a) * b) * f32(1 << W_BITS)))
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
            q1 = np.zeros(4, f32); q2 = np.zeros(4, f32)      # accumulators for b1 / b2: [p0, p1, p2, p3]
            t1 = f32(0.0); t2 = f32(0.0)
            dd = np.zeros(win, np.int32)
            for y in range(win):
                for x in range(win):
                    r = 24 + iny + y; c = 24 + inx + x
                    dd[x] = _descale(int(J[r, c]) * iw00 + int(J[r, c + 1]) * iw01 + int(J[r + 1, c]) * iw10 + int(J[r + 1, c + 1]) * iw11, W_BITS - 5) - Iw[y, x]
                for ch in range(2):
                    x0 = 8 * ch
                    for p in range(4):
                        s1 = dd[x0 + p] * Ixw[y, x0 + p] + dd[x0 + 4 + p] * Ixw[y, x0 + 4 + p]
                        s2 = dd[x0 + p] * Iyw[y, x0 + p] + dd[x0 + 4 + p] * Iyw[y, x0 + 4 + p]
                        q1[p] = f32(q1[p] + f32(s1)); q2[p] = f32(q2[p] + f32(s2))
                for x in range(16, win):
                    t1 = f32(t1 + f32(dd[x] * Ixw[y, x])); t2 = f32(t2 + f32(dd[x] * Iyw[y, x]))
            # (qb0 + qb1) lanes: [b1(p0)+b1(p2), b2(p0)+b2(p2), b1(p1)+b1(p3), b2(p1)+b2(p3)] ; reduce = lane0 + lane1 of the recombined
            ib1 = f32(t1 + f32(f32(q1[0] + q1[2]) + f32(q1[1] + q1[3])))
            ib2 = f32(t2 + f32(f32(q2[0] + q2[2]) + f32(q2[1] + q2[3])))
            b1 = f32(ib1 * FLT_SCALE); b2 = f32(ib2 * FLT_SCALE)
            dx = f32(f32(f32(A12 * b2) - f32(A22 * b1)) * D)
            dy = f32(f32(f32(A12 * b1) - f32(A11 * b2)) * D)
            npx = f32(npx + dx); npy = f32(npy + dy)
            nx = f32(npx + half); ny = f32(npy + half)
            if np.float64(dx) * np.float64(dx) + np.float64(dy) * np.float64(dy) <= eps2:
                break
            if j > 0 and np.float64(abs(f32(dx + pdx))) < 0.01 and np.float64(abs(f32(dy + pdy))) < 0.01:
                nx = f32(nx - f32(dx * f32(0.5))); ny = f32(ny - f32(dy * f32(0.5)))
                break
            pdx = dx; pdy = dy
    return nx, ny, status


def build_pyramid(img, levels=2, pad=24):
    import cv2
    out = []; der = []
    cur = img
    for l in range(levels + 1):
        if l > 0:
            cur = cv2.pyrDown(cur)
        p = cv2.copyMakeBorder(cur, pad, pad, pad, pad, cv2.BORDER_REFLECT_101)
        dx = cv2.Scharr(cur, cv2.CV_16S, 1, 0); dy = cv2.Scharr(cur, cv2.CV_16S, 0, 1)
        d = np.zeros((p.shape[0], p.shape[1], 2), np.int16)
        d[pad:-pad, pad:-pad, 0] = dx; d[pad:-pad, pad:-pad, 1] = dy
        out.append(p); der.append(d)
    return out, der


def calc_optical_flow_pyr_lk(imgA, imgB, prev_pts, init_pts, max_level=2, max_iter=30, eps=0.01):
    from numba.typed import List
    pa, da = build_pyramid(imgA, max_level); pb, _ = build_pyramid(imgB, max_level)
    LA = List(); DA = List(); LB = List()
    for l in range(max_level + 1):
        LA.append(pa[l]); DA.append(da[l]); LB.append(pb[l])
    out = np.zeros((len(prev_pts), 2), np.float32); st = np.zeros(len(prev_pts), np.uint8)
    for i in range(len(prev_pts)):
        x, y, s = _lk_point(LA, DA, LB, prev_pts[i].astype(np.float32), init_pts[i].astype(np.float32), max_level, max_iter,
                            float(eps) * float(eps), 1e-4)
        out[i] = (x, y); st[i] = 1 if s else 0
    return out, st
