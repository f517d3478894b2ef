"""TEST INFRASTRUCTURE (oracle/): OpenCV calls on behalf of oracle/_ref/larvio_ref_fe.

The reference's front end (src/image_processor.cpp, src/ORBDescriptor.cpp) is compiled unmodified against the stand-in headers
of oracle/ref_shim/ (`make ref_fe`).  Those headers contain no image-processing code: every OpenCV FUNCTION the reference calls
(CLAHE, buildOpticalFlowPyramid, calcOpticalFlowPyrLK, goodFeaturesToTrack, undistortPoints, findFundamentalMat, GaussianBlur,
copyMakeBorder, resize, Rodrigues, fastAtan2) is forwarded over a pipe to this process, which executes it with the cv2 module of
this image (OpenCV 4.13, the same binary the front-end oracle is pinned against).  Protocol: little-endian frames
[int32 opcode][int32 n_args] then per argument [int32 dtype code][int32 ndim][int32 dims...][raw bytes]; the reply is
[int32 n_results] + arrays in the same encoding.  Never imported by the product."""
import struct
import sys

import cv2
import numpy as np

DT = {0: np.uint8, 1: np.int32, 2: np.float32, 3: np.float64}
CODE = {np.dtype(v): k for k, v in DT.items()}
inp = sys.stdin.buffer
out = sys.stdout.buffer
sys.stdout = sys.stderr          # anything printed by accident must not corrupt the pipe


def read_exact(n):
    b = inp.read(n)
    if len(b) != n:
        raise EOFError
    return b


def read_array():
    dt, nd = struct.unpack("<ii", read_exact(8))
    dims = struct.unpack("<%di" % nd, read_exact(4 * nd)) if nd else ()
    a = np.frombuffer(read_exact(int(np.prod(dims, dtype=np.int64)) * np.dtype(DT[dt]).itemsize), DT[dt]).reshape(dims)
    return a.copy()


def write_arrays(arrs):
    out.write(struct.pack("<i", len(arrs)))
    for a in arrs:
        a = np.ascontiguousarray(a)
        out.write(struct.pack("<ii", CODE[a.dtype], a.ndim))
        if a.ndim:
            out.write(struct.pack("<%di" % a.ndim, *a.shape))
        out.write(a.tobytes())
    out.flush()


images = {}          # handle -> image an optical-flow pyramid was built from (cv2 rebuilds the identical pyramid inside calcOpticalFlowPyrLK)


def op_clahe(a):
    img, clip, tiles = a
    return [cv2.createCLAHE(float(clip[0]), (int(tiles[0]), int(tiles[1]))).apply(img)]


def op_build_pyramid(a):
    # returns the level images (what the reference reads with at<>/rows/cols); the handle keeps the source for LK
    img, win, levels, handle = a
    n, pyr = cv2.buildOpticalFlowPyramid(img, (int(win[0]), int(win[1])), int(levels[0]), None, True, cv2.BORDER_REFLECT_101, cv2.BORDER_CONSTANT, False)
    images[int(handle[0])] = img
    for k in [k for k in images if k < int(handle[0]) - 8]:
        del images[k]
    return [np.array([n], np.int32)] + [np.ascontiguousarray(pyr[2 * l]) for l in range(n + 1)]


def op_lk(a):
    hp, hc, prev_pts, next_pts, win, levels, crit, flags = a
    crit_t = (int(crit[0]), int(crit[1]), float(crit[2]))
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(images[int(hp[0])], images[int(hc[0])], prev_pts.reshape(-1, 1, 2).astype(np.float32),
                                          next_pts.reshape(-1, 1, 2).astype(np.float32).copy(), winSize=(int(win[0]), int(win[1])),
                                          maxLevel=int(levels[0]), criteria=crit_t, flags=int(flags[0]))
    return [nxt.reshape(-1, 2).astype(np.float32), st.reshape(-1).astype(np.uint8)]


def op_gftt(a):
    img, params, mask = a
    maxc, q, md = int(params[0]), float(params[1]), float(params[2])
    m = mask if mask.size else None
    pts = cv2.goodFeaturesToTrack(img, maxc, q, md, mask=m)
    return [np.zeros((0, 2), np.float32) if pts is None else pts.reshape(-1, 2).astype(np.float32)]


def op_undistort(a):
    pts, K, dist, R, Kn, fisheye = a
    if int(fisheye[0]):
        o = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2).astype(np.float32), K, dist, R=R, P=Kn)
    else:
        o = cv2.undistortPoints(pts.reshape(-1, 1, 2).astype(np.float32), K, dist, R=R, P=Kn)
    return [o.reshape(-1, 2).astype(np.float32)]


def op_fundamental(a):
    p1, p2, params = a
    F, mask = cv2.findFundamentalMat(p1.reshape(-1, 2), p2.reshape(-1, 2), int(params[0]), float(params[1]), float(params[2]))
    if mask is None:
        return [np.zeros(0, np.uint8)]
    return [mask.reshape(-1).astype(np.uint8)]


def op_gaussian(a):
    img, k, sig, border = a
    return [cv2.GaussianBlur(img, (int(k[0]), int(k[1])), float(sig[0]), sigmaY=float(sig[1]), borderType=int(border[0]))]


def op_make_border(a):
    img, b, border = a
    return [cv2.copyMakeBorder(img, int(b[0]), int(b[1]), int(b[2]), int(b[3]), int(border[0]))]


def op_resize(a):
    img, sz, interp = a
    return [cv2.resize(img, (int(sz[0]), int(sz[1])), interpolation=int(interp[0]))]


def op_rodrigues(a):
    r, _ = cv2.Rodrigues(a[0].reshape(3, 1))
    return [r]


def op_fast_atan2(a):
    return [np.array([cv2.fastAtan2(float(a[0][0]), float(a[0][1]))], np.float32)]


OPS = {1: op_clahe, 2: op_build_pyramid, 3: op_lk, 4: op_gftt, 5: op_undistort, 6: op_fundamental, 7: op_gaussian, 8: op_make_border,
       9: op_resize, 10: op_rodrigues, 11: op_fast_atan2}

if __name__ == "__main__":
    cv2.setNumThreads(1)
    try:
        while True:
            op, n = struct.unpack("<ii", read_exact(8))
            args = [read_array() for _ in range(n)]
            write_arrays(OPS[op](args))
    except EOFError:
        pass
