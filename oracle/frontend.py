"""CPU oracle of ``ImageProcessor`` (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates /root/reference/src/image_processor.cpp line by line around the same OpenCV
calls (cv2 4.13): processImage :130-219, integrateImuData :222-263,
predictFeatureTracking :266-293, createImagePyramids :318-334, initializeFirstFrame
:337-352, initializeFirstFeatures :355-537, trackFeatures :540-811, trackNewFeatures
:813-1002, findNewFeaturesToBeTracked :1005-1037, undistortPoints :1040-1072,
getFeatureMsg :1076-1128, publish :1170-1172 (bookkeeping only; the debug overlay is
GUI-only and dropped).  Float32/float64 typing of every intermediate follows the C++.
"""
from __future__ import annotations

import numpy as np
import cv2

from .orb import OrbOracle, hamming, hamming_rows

f32 = np.float32


def _mm33_f32(A, B):
    """cv::Matx33f product: s += a(i,k)*b(k,j) in float, k ascending."""
    C = np.zeros((3, 3), f32)
    for i in range(3):
        for j in range(3):
            s = f32(0)
            for k in range(3):
                s = f32(s + f32(A[i, k] * B[k, j]))
            C[i, j] = s
    return C


def _inv33_f32(a):
    """cv::Matx33f::inv() closed form (Matx_FastInvOp<float,3,3>)."""
    a = a.astype(f32)
    d = f32(f32(f32(a[0, 0] * f32(f32(a[1, 1] * a[2, 2]) - f32(a[2, 1] * a[1, 2])))
                - f32(a[0, 1] * f32(f32(a[1, 0] * a[2, 2]) - f32(a[2, 0] * a[1, 2]))))
            + f32(a[0, 2] * f32(f32(a[1, 0] * a[2, 1]) - f32(a[2, 0] * a[1, 1]))))
    d = f32(f32(1) / d)
    b = np.zeros((3, 3), f32)
    b[0, 0] = f32(f32(f32(a[1, 1] * a[2, 2]) - f32(a[1, 2] * a[2, 1])) * d)
    b[0, 1] = f32(f32(f32(a[0, 2] * a[2, 1]) - f32(a[0, 1] * a[2, 2])) * d)
    b[0, 2] = f32(f32(f32(a[0, 1] * a[1, 2]) - f32(a[0, 2] * a[1, 1])) * d)
    b[1, 0] = f32(f32(f32(a[1, 2] * a[2, 0]) - f32(a[1, 0] * a[2, 2])) * d)
    b[1, 1] = f32(f32(f32(a[0, 0] * a[2, 2]) - f32(a[0, 2] * a[2, 0])) * d)
    b[1, 2] = f32(f32(f32(a[0, 2] * a[1, 0]) - f32(a[0, 0] * a[1, 2])) * d)
    b[2, 0] = f32(f32(f32(a[1, 0] * a[2, 1]) - f32(a[1, 1] * a[2, 0])) * d)
    b[2, 1] = f32(f32(f32(a[0, 1] * a[2, 0]) - f32(a[0, 0] * a[2, 1])) * d)
    b[2, 2] = f32(f32(f32(a[0, 0] * a[1, 1]) - f32(a[0, 1] * a[1, 0])) * d)
    return b


class FeatureMsg:
    """MonoCameraMeasurement (include/larvio/feature_msg.h:15-56) as arrays."""

    def __init__(self, t):
        self.t = float(t)
        self.ids = np.zeros(0, np.uint64)
        # columns: u v u_init v_init u_vel v_vel u_init_vel v_init_vel
        self.data = np.zeros((0, 8), np.float64)


class ImageProcessorOracle:
    FIRST_IMAGE, SECOND_IMAGE, OTHER_IMAGES = 1, 2, 3

    def __init__(self, cfg_raw: dict):
        r = cfg_raw
        self.patch_size = int(r["patch_size"]); self.pyramid_levels = int(r["pyramid_levels"])
        self.max_iteration = int(r["max_iteration"]); self.track_precision = float(r["track_precision"])
        self.max_features_num = int(r["max_features_num"]); self.min_distance = int(r["min_distance"])
        self.flag_equalize = bool(int(r["flag_equalize"]))
        self.pub_frequency = int(r["pub_frequency"]) if float(r["pub_frequency"]).is_integer() else r["pub_frequency"]
        self.model = r.get("distortion_model", "radtan")
        it, dc = r["intrinsics"], r["distortion_coeffs"]
        self.intr = np.array([it["fx"], it["fy"], it["cx"], it["cy"]], np.float64)
        self.dist = np.array([dc["k1"], dc["k2"], dc["p1"], dc["p2"]], np.float64)
        self.K = np.array([[self.intr[0], 0, self.intr[2]], [0, self.intr[1], self.intr[3]], [0, 0, 1.0]])
        T = np.array(r["T_cam_imu"]["data"], np.float64).reshape(4, 4)
        self.R_cam_imu = T[:3, :3].T.copy()           # image_processor.cpp:93
        self.image_state = self.FIRST_IMAGE
        self.next_feature_id = 0
        self.bFirstImg = False
        self.pub_counter = 0
        self.prev_img = None; self.curr_img = None
        self.prev_orb = None; self.curr_orb = None
        self.prev_pts = np.zeros((0, 2), f32); self.curr_pts = np.zeros((0, 2), f32)
        self.pts_ids = []; self.pts_lifetime = []
        self.init_pts = np.zeros((0, 2), f32)
        self.descs = np.zeros((0, 32), np.uint8)
        self.new_pts = np.zeros((0, 2), f32)
        self.last_pub_time = 0.0; self.curr_img_time = 0.0; self.prev_img_time = 0.0
        self.R_p2c = np.eye(3, dtype=f32)
        self.clahe = cv2.createCLAHE(3.0, (8, 8))
        self.trace = {}        # per-frame stage survivors, for parity debugging

    # ---- image_processor.cpp:130-219
    def process_image(self, image: np.ndarray, t_img: float, imu: np.ndarray):
        """imu: rows [t, wx, wy, wz, ax, ay, az] currently in the caller's buffer. Returns FeatureMsg or None."""
        if not self.bFirstImg:
            if len(imu) > 0 and imu[0, 0] - t_img <= 0.0:
                self.bFirstImg = True
            else:
                return None
        self.trace = {}
        self.curr_img = self.clahe.apply(image) if self.flag_equalize else image   # :318-334
        self.curr_orb = OrbOracle(self.curr_img)                                    # :150
        self.curr_img_time = float(t_img)
        msg = None
        if self.image_state == self.FIRST_IMAGE:
            if self._initialize_first_frame():
                self.image_state = self.SECOND_IMAGE
        elif self.image_state == self.SECOND_IMAGE:
            if not self._initialize_first_features(imu):
                self.image_state = self.FIRST_IMAGE
            else:
                if self.curr_img_time - self.last_pub_time >= 0.9 * (1.0 / self.pub_frequency):
                    self._find_new_features()
                    msg = self._get_feature_msg()
                    self._publish()
                self.image_state = self.OTHER_IMAGES
        else:
            self._integrate_imu(imu)
            self._track_features()
            self._track_new_features()
            if self.curr_img_time - self.last_pub_time >= 0.9 * (1.0 / self.pub_frequency):
                self._find_new_features()
                msg = self._get_feature_msg()
                self._publish()
        self.prev_img = self.curr_img
        self.prev_orb = self.curr_orb
        self.prev_pts = self.curr_pts
        self.curr_pts = np.zeros((0, 2), f32)
        self.prev_img_time = self.curr_img_time
        return msg

    # ---- :222-263
    def _integrate_imu(self, imu):
        t = imu[:, 0] if len(imu) else np.zeros(0)
        b = 0
        while b < len(t) and t[b] - self.prev_img_time < -0.0049:
            b += 1
        e = b
        while e < len(t) and t[e] - self.curr_img_time < 0.0049:
            e += 1
        mean = np.zeros(3, f32)
        for k in range(b, e):
            mean = (mean + imu[k, 1:4].astype(f32)).astype(f32)
        if e - b > 0:
            mean = (mean * f32(f32(1.0) / f32(e - b))).astype(f32)
        Rt = self.R_cam_imu.T
        md = mean.astype(np.float64)
        cam = np.array([Rt[i, 0] * md[0] + Rt[i, 1] * md[1] + Rt[i, 2] * md[2] for i in range(3)]).astype(f32)
        dtime = self.curr_img_time - self.prev_img_time
        rvec = (cam.astype(np.float64) * dtime).astype(f32)
        R, _ = cv2.Rodrigues(rvec.reshape(3, 1))
        self.R_p2c = R.astype(f32).T.copy()

    # ---- :266-293
    def _predict(self, pts):
        if len(pts) == 0:
            return np.zeros((0, 2), f32)
        K = np.array([[self.intr[0], 0, self.intr[2]], [0, self.intr[1], self.intr[3]], [0, 0, 1]], np.float64).astype(f32)
        H = _mm33_f32(_mm33_f32(K, self.R_p2c), _inv33_f32(K))
        # H * (x, y, 1) in float32, products and sums rounded one by one like the Matx/Point3f arithmetic (vectorised over points:
        # numpy float32 array ops round after every operation, exactly like the scalar loop they replace)
        x = np.ascontiguousarray(pts[:, 0], f32); y = np.ascontiguousarray(pts[:, 1], f32)
        q = []
        for r in range(3):
            sacc = H[r, 0] * x
            sacc = sacc + H[r, 1] * y
            sacc = sacc + f32(H[r, 2] * f32(1.0))
            q.append(sacc.astype(f32))
        out = np.zeros((len(pts), 2), f32)
        out[:, 0] = q[0] / q[2]; out[:, 1] = q[1] / q[2]
        return out

    def _lk(self, img_a, img_b, pts_a, init_b):
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, self.max_iteration, self.track_precision)
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(
            img_a, img_b, pts_a.reshape(-1, 1, 2).astype(f32), init_b.reshape(-1, 1, 2).astype(f32).copy(),
            winSize=(self.patch_size, self.patch_size), maxLevel=self.pyramid_levels, criteria=crit,
            flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        return nxt.reshape(-1, 2), st.reshape(-1).astype(np.uint8)

    def _in_image(self, pts, st):
        h, w = self.curr_img.shape
        st = st.copy()
        if len(pts):
            out = (pts[:, 1] < 0) | (pts[:, 1] > h - 1) | (pts[:, 0] < 0) | (pts[:, 0] > w - 1)
            st[out & (st != 0)] = 0
        return st

    def _reverse_check(self, curr_in, prev_in):
        back, st = self._lk(self.curr_img, self.prev_img, curr_in, prev_in.copy())
        h, w = self.prev_img.shape
        if len(back):
            out = (back[:, 1] < 0) | (back[:, 1] > h - 1) | (back[:, 0] < 0) | (back[:, 0] > w - 1)
            d = (back - prev_in).astype(f32)                                  # Point2f difference
            dis = np.sqrt(d[:, 0].astype(np.float64) * d[:, 0].astype(np.float64) + d[:, 1].astype(np.float64) * d[:, 1].astype(np.float64)).astype(f32)
            st[(st != 0) & (out | (dis > 1))] = 0
        return st

    def _undistort(self, pts, to_pixels: bool):
        if len(pts) == 0:
            return np.zeros((0, 2), f32)
        P = self.K if to_pixels else np.eye(3)
        src = pts.reshape(-1, 1, 2).astype(f32)
        if self.model == "equidistant":
            out = cv2.fisheye.undistortPoints(src, self.K, self.dist, R=np.eye(3), P=P)
        else:
            out = cv2.undistortPoints(src, self.K, self.dist, R=np.eye(3), P=P)
        return out.reshape(-1, 2).astype(f32)

    def _ransac(self, p1, p2):
        """cv::findFundamentalMat(FM_RANSAC, 1.0, 0.99); returns a marker vector or None (→ keep all)."""
        if len(p1) < 7:
            return None
        _, mask = cv2.findFundamentalMat(p1.astype(f32), p2.astype(f32), cv2.FM_RANSAC, 1.0, 0.99)
        if mask is None:
            return None
        return mask.reshape(-1).astype(np.uint8)

    @staticmethod
    def _keep(arr, markers):
        if markers is None or len(markers) != len(arr):   # image_processor.h:219-223
            return arr
        m = np.asarray(markers) != 0
        if isinstance(arr, list):
            return [a for a, k in zip(arr, m) if k]
        return arr[m]

    def _clear_tracks(self):
        self.prev_pts = np.zeros((0, 2), f32); self.curr_pts = np.zeros((0, 2), f32)
        self.pts_ids = []; self.pts_lifetime = []
        self.init_pts = np.zeros((0, 2), f32); self.descs = np.zeros((0, 32), np.uint8)

    # ---- :337-352
    def _initialize_first_frame(self):
        p = cv2.goodFeaturesToTrack(self.curr_img, self.max_features_num, 0.01, self.min_distance)
        self.new_pts = np.zeros((0, 2), f32) if p is None else p.reshape(-1, 2).astype(f32)
        self.last_pub_time = self.curr_img_time
        return len(self.new_pts) > 20

    # ---- shared LK<->LK<->ORB<->RANSAC chain of :355-537 and :813-1002
    def _new_feature_chain(self, min_after_each: int, min_after_lk: int):
        curr = self._predict(self.new_pts)
        nxt, st = self._lk(self.prev_img, self.curr_img, self.new_pts, curr)
        st = self._in_image(nxt, st)
        prev1 = self._keep(self.new_pts, st); curr1 = self._keep(nxt, st)
        self.trace["new_fwd"] = st.copy()
        if len(prev1) < min_after_lk or len(prev1) <= 0:
            return None
        rst = self._reverse_check(curr1, prev1)
        prev2 = self._keep(prev1, rst); curr2 = self._keep(curr1, rst)
        self.trace["new_rev"] = rst.copy()
        if len(prev2) < min_after_lk or len(prev2) <= 0:
            return None
        dprev = self.prev_orb.compute(prev2); dcurr = self.curr_orb.compute(curr2)
        dis = hamming_rows(dprev, dcurr)
        dm = (dis <= 58).astype(np.uint8)
        self.trace["new_desc"] = dm.copy(); self.trace["new_hamming"] = dis
        prev3 = self._keep(prev2, dm); curr3 = self._keep(curr2, dm); desc3 = dprev[dm != 0]
        if len(prev3) < 20:
            return None
        up = self._undistort(prev3, True); uc = self._undistort(curr3, True)
        rm = self._ransac(up, uc)
        self.trace["new_ransac"] = None if rm is None else rm.copy()
        prev4 = self._keep(prev3, rm); curr4 = self._keep(curr3, rm); desc4 = self._keep(desc3, rm)
        if len(curr4) < min_after_each or len(curr4) <= 0:
            return None
        return prev4, curr4, desc4

    # ---- :355-537
    def _initialize_first_features(self, imu):
        self._integrate_imu(imu)
        res = self._new_feature_chain(20, 20)
        if res is None:
            return False
        prev4, curr4, desc4 = res
        n = len(prev4)
        self.prev_pts = prev4.copy(); self.curr_pts = curr4.copy()
        self.init_pts = np.full((n, 2), -1, f32)
        self.pts_ids = list(range(self.next_feature_id, self.next_feature_id + n)); self.next_feature_id += n
        self.pts_lifetime = [2] * n
        self.descs = desc4.copy()
        self.new_pts = np.zeros((0, 2), f32)
        return True

    # ---- :540-811
    def _track_features(self):
        if len(self.prev_pts) == 0:
            return
        pred = self._predict(self.prev_pts)
        nxt, st = self._lk(self.prev_img, self.curr_img, self.prev_pts, pred)
        st = self._in_image(nxt, st)
        self.trace["trk_fwd"] = st.copy(); self.trace["trk_fwd_pts"] = nxt.copy()
        ids = self._keep(self.pts_ids, st); life = self._keep(self.pts_lifetime, st)
        prev1 = self._keep(self.prev_pts, st); curr1 = self._keep(nxt, st)
        init1 = self._keep(self.init_pts, st); desc1 = self._keep(self.descs, st)
        if len(curr1) == 0:
            self._clear_tracks(); return
        rst = self._reverse_check(curr1, prev1)
        self.trace["trk_rev"] = rst.copy()
        ids = self._keep(ids, rst); life = self._keep(life, rst)
        prev2 = self._keep(prev1, rst); curr2 = self._keep(curr1, rst)
        init2 = self._keep(init1, rst); desc2 = self._keep(desc1, rst)
        if len(curr2) == 0:
            self._clear_tracks(); return
        dcurr = self.curr_orb.compute(curr2)
        dis = hamming_rows(desc2, dcurr)
        dm = (dis <= 58).astype(np.uint8)
        self.trace["trk_desc"] = dm.copy(); self.trace["trk_hamming"] = dis
        ids = self._keep(ids, dm); life = self._keep(life, dm)
        prev3 = self._keep(prev2, dm); curr3 = self._keep(curr2, dm)
        init3 = self._keep(init2, dm); desc3 = self._keep(desc2, dm)
        if len(prev3) == 0:
            self._clear_tracks(); return
        up = self._undistort(prev3, True); uc = self._undistort(curr3, True)
        rm = self._ransac(up, uc)
        self.trace["trk_ransac"] = None if rm is None else rm.copy()
        ids = self._keep(ids, rm); life = self._keep(life, rm)
        prev4 = self._keep(prev3, rm); curr4 = self._keep(curr3, rm)
        init4 = self._keep(init3, rm); desc4 = self._keep(desc3, rm)
        if len(curr4) == 0:
            self._clear_tracks(); return
        self.prev_pts = prev4.copy(); self.curr_pts = curr4.copy()
        self.pts_ids = list(ids); self.pts_lifetime = [l + 1 for l in life]
        self.init_pts = init4.copy(); self.descs = desc4.copy()

    # ---- :813-1002
    def _track_new_features(self):
        if len(self.new_pts) <= 0:
            return
        res = self._new_feature_chain(1, 1)
        if res is None:
            return
        prev4, curr4, desc4 = res
        n = len(prev4)
        self.prev_pts = np.concatenate([self.prev_pts.reshape(-1, 2), prev4]).astype(f32)
        self.curr_pts = np.concatenate([self.curr_pts.reshape(-1, 2), curr4]).astype(f32)
        self.pts_ids = list(self.pts_ids) + list(range(self.next_feature_id, self.next_feature_id + n))
        self.next_feature_id += n
        self.pts_lifetime = list(self.pts_lifetime) + [2] * n
        self.init_pts = np.concatenate([self.init_pts.reshape(-1, 2), prev4]).astype(f32)
        self.descs = np.concatenate([self.descs.reshape(-1, 32), desc4]).astype(np.uint8)
        self.new_pts = np.zeros((0, 2), f32)

    # ---- :1005-1037
    def _find_new_features(self):
        h, w = self.curr_img.shape
        mask = np.full((h, w), 255, np.uint8)
        md = self.min_distance
        for p in self.curr_pts:
            ry = int(np.floor(abs(float(p[1])) + 0.5) * (1 if p[1] >= 0 else -1))   # C round(): half away from zero
            rx = int(np.floor(abs(float(p[0])) + 0.5) * (1 if p[0] >= 0 else -1))
            r0 = max(ry - md, 0); r1 = min(ry + md, h - 1)
            c0 = max(rx - md, 0); c1 = min(rx + md, w - 1)
            mask[r0:r1 + 1, c0:c1 + 1] = 0
        self.new_pts = np.zeros((0, 2), f32)
        want = self.max_features_num - len(self.curr_pts)
        if want > 0:
            p = cv2.goodFeaturesToTrack(self.curr_img, want, 0.01, self.min_distance, mask=mask)
            if p is not None:
                self.new_pts = p.reshape(-1, 2).astype(f32)
        self.trace["mask"] = mask

    # ---- :1076-1128
    def _get_feature_msg(self):
        msg = FeatureMsg(self.curr_img_time)
        cu = self._undistort(self.curr_pts, False)
        iu = self._undistort(self.init_pts, False)
        pu = self._undistort(self.prev_pts, False)
        dt_1 = self.curr_img_time - self.prev_img_time
        prev_is_last = self.prev_img_time == self.last_pub_time
        dt_2 = dt_1 if prev_is_last else self.prev_img_time - self.last_pub_time
        n = len(self.pts_ids)
        msg.ids = np.array(self.pts_ids, np.uint64)
        d = np.zeros((n, 8), np.float64)
        for i in range(n):
            d[i, 0] = cu[i, 0]; d[i, 1] = cu[i, 1]
            d[i, 4] = np.float64(f32(cu[i, 0] - pu[i, 0])) / dt_1
            d[i, 5] = np.float64(f32(cu[i, 1] - pu[i, 1])) / dt_1
            if self.init_pts[i, 0] == -1 and self.init_pts[i, 1] == -1:
                d[i, 2] = -1; d[i, 3] = -1
            else:
                d[i, 2] = iu[i, 0]; d[i, 3] = iu[i, 1]
                self.init_pts[i] = (-1, -1)
                a = cu[i] if prev_is_last else pu[i]
                d[i, 6] = np.float64(f32(a[0] - iu[i, 0])) / dt_2
                d[i, 7] = np.float64(f32(a[1] - iu[i, 1])) / dt_2
        msg.data = d
        return msg

    # ---- :1131-1175 (bookkeeping part)
    def _publish(self):
        self.last_pub_time = self.curr_img_time
        self.pub_counter += 1
