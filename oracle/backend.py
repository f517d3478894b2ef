"""CPU oracle of ``LarVio`` (TEST INFRASTRUCTURE — see oracle/__init__.py).

numpy float64 restatement of /root/reference/src/larvio.cpp - pure MSCKF and hybrid (1-D / 3-D inverse-depth EKF-SLAM features,
Schmidt nuisance states), FEJ, online extrinsics / td / IMU intrinsics, ZUPT - pinned call by call against the reference's own
larvio.cpp compiled unmodified (oracle/_ref, tests/golden/ref_*.npz, tests/test_cpu.py::test_backend_oracle_matches_the_compiled_reference):
processFeatures :363-461, batchImuProcessing :464-517, processModel :520-578, predictNewState
:581-649, calPhi :3475-3530, stateAugmentation :720-801, addFeatureObservations :804-856,
measurementJacobian_msckf :859-921, featureJacobian_msckf :924-981, measurementUpdate_msckf
:1420-1602, measurementUpdate_hybrid :1605-1862 (empty SLAM blocks), gatingTest :1865-1880,
removeLostFeatures :1883-2256, findRedundantImuStates :2259-2307, pruneImuStateBuffer :2310-2641,
checkZUPT :2751-2788 (detection only); Feature::{cost,jacobian,generateInitialGuess,checkMotion,
initializePosition,initializePosition_AssignAnchor} include/larvio/feature.hpp:252-721;
quaternion helpers include/larvio/math_utils.hpp:26-231.
Third-party pieces are replaced per SURVEY.md §8c: SPQR thin QR -> numpy.linalg.qr, JacobiSVD
left null space -> numpy.linalg.svd, LDLT solves -> numpy.linalg.solve, boost chi^2 quantile ->
scipy.stats.chi2.ppf(0.05, k).  The update is invariant to these basis choices up to rounding.
The initialisers (larvio.cpp:375-391) are out of scope: ``set_initial_state`` injects what they
would leave behind.
"""
from __future__ import annotations

import numpy as np
from scipy.stats import chi2

GRAVITY = np.array([0.0, 0.0, -9.81])


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def quat_to_rot(q):
    """Eigen Quaterniond(w,x,y,z).toRotationMatrix() for q = [x y z w]."""
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    """Eigen Quaterniond(Matrix3d).coeffs() -> [x y z w]."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t; q[1] = (R[0, 2] - R[2, 0]) * t; q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def quat_mul(a, b):
    """Hamilton product of Eigen quaternions given as [x y z w]."""
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def small_angle_quat(dtheta):
    dq = dtheta / 2.0
    n2 = float(dq @ dq)
    q = np.zeros(4)
    if n2 <= 1:
        q[:3] = dq; q[3] = np.sqrt(1 - n2)
    else:
        q[:3] = dq; q[3] = 1
        q = q / np.sqrt(1 + n2)
    return q


class ImuState:
    def __init__(self):
        self.id = 0; self.time = 0.0; self.dt = 0.0
        self.q = np.array([0, 0, 0, 1.0]); self.p = np.zeros(3); self.v = np.zeros(3)
        self.bg = np.zeros(3); self.ba = np.zeros(3)
        self.R_imu_cam0 = np.eye(3); self.t_cam0_imu = np.zeros(3)

    def copy(self):
        o = ImuState()
        o.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        return o


class AugState:
    def __init__(self, sid):
        self.id = sid; self.time = 0.0; self.dt = 0.0
        self.q = np.array([0, 0, 0, 1.0]); self.p = np.zeros(3); self.p_FEJ = np.zeros(3)
        self.R_imu_cam0 = np.eye(3); self.t_cam0_imu = np.zeros(3)
        self.q_cam = np.array([0, 0, 0, 1.0]); self.p_cam = np.zeros(3)


class Feature:
    huber_epsilon = 0.01; estimation_precision = 5e-7; initial_damping = 1e-3
    outer_max = 10; inner_max = 10

    def __init__(self, fid, translation_threshold):
        self.id = fid
        self.obs = {}        # state_id -> (2,)
        self.obs_vel = {}
        self.position = np.zeros(3)
        self.is_initialized = False
        self.id_anchor = -1
        self.invDepth = 0.0
        self.obs_anchor = np.zeros(3)
        self.invParam = np.zeros(3)        # (x/z, y/z, 1/z) in the anchor camera frame (feature.hpp:231, 542, 711, 880)
        self.in_state = False
        self.ekf_feature = False
        self.totalObsNum = 0
        self.position_FEJ = np.zeros(3)
        self.translation_threshold = translation_threshold

    @staticmethod
    def _cost(R, t, x, z):
        h = R @ np.array([x[0], x[1], 1.0]) + x[2] * t
        zh = np.array([h[0] / h[2], h[1] / h[2]])
        d = zh - z
        return float(d @ d)

    def _jacobian(self, R, t, x, z):
        h = R @ np.array([x[0], x[1], 1.0]) + x[2] * t
        W = np.zeros((3, 3)); W[:, :2] = R[:, :2]; W[:, 2] = t
        J = np.zeros((2, 3))
        J[0] = 1 / h[2] * W[0] - h[0] / (h[2] * h[2]) * W[2]
        J[1] = 1 / h[2] * W[1] - h[1] / (h[2] * h[2]) * W[2]
        r = np.array([h[0] / h[2], h[1] / h[2]]) - z
        e = np.linalg.norm(r)
        w = 1.0 if e <= self.huber_epsilon else np.sqrt(2.0 * self.huber_epsilon / e)
        return J, r, w

    @staticmethod
    def _initial_guess(R, t, z1, z2):
        m = R @ np.array([z1[0], z1[1], 1.0])
        A = np.array([m[0] - z2[0] * m[2], m[1] - z2[1] * m[2]])
        b = np.array([z2[0] * t[2] - t[0], z2[1] * t[2] - t[1]])
        depth = (A @ b) / (A @ A)
        return np.array([z1[0] * depth, z1[1] * depth, depth])

    def check_motion(self, aug, if_tracked):
        ids = sorted(self.obs.keys())
        first = ids[0]
        last = ids[-2] if if_tracked else ids[-1]
        Rf = quat_to_rot(aug[first].q_cam); tf = aug[first].p_cam
        tl = aug[last].p_cam
        d = np.array([self.obs[first][0], self.obs[first][1], 1.0])
        d = d / np.linalg.norm(d)
        d = Rf @ d
        tr = tl - tf
        par = float(tr @ d)
        orth = tr - par * d
        return np.linalg.norm(orth) > self.translation_threshold

    def initialize_inv_param(self, aug, curr_id):
        """initializeInvParamPosition (feature.hpp:723-890): like initializePosition but always starts from the
        two-view guess and marks the feature as a potential EKF-SLAM feature."""
        ok = self.initialize_position(aug, curr_id, force_guess=True)
        if ok:
            self.ekf_feature = True
        return ok

    def initialize_position(self, aug, curr_id=None, force_guess=False):
        """initializePosition (curr_id given: skip the current camera) / _AssignAnchor (curr_id None)."""
        poses = []; meas = []; cam_ids = []
        for sid in sorted(self.obs.keys()):
            if sid not in aug:
                continue
            if curr_id is not None and sid == curr_id:
                continue
            meas.append(self.obs[sid].copy())
            poses.append((quat_to_rot(aug[sid].q_cam), aug[sid].p_cam.copy()))
            cam_ids.append(sid)
        Rl, tl = poses[-1]
        rel = []
        for (R, t) in poses:
            # pose.inverse() * T_c_w_last
            Ri = R.T
            rel.append((Ri @ Rl, Ri @ (tl - t)))
        if not self.is_initialized or force_guess:
            init = self._initial_guess(rel[0][0], rel[0][1], meas[-1], meas[0])
        else:
            init = Rl.T @ (self.position - tl)
        sol = np.array([init[0] / init[2], init[1] / init[2], 1.0 / init[2]])
        lam = self.initial_damping
        inner = 0; outer = 0
        reduced = False
        delta_norm = 0.0
        total = sum(self._cost(R, t, sol, z) for (R, t), z in zip(rel, meas))
        while True:
            A = np.zeros((3, 3)); b = np.zeros(3)
            for (R, t), z in zip(rel, meas):
                J, r, w = self._jacobian(R, t, sol, z)
                if w == 1:
                    A += J.T @ J; b += J.T @ r
                else:
                    A += w * w * (J.T @ J); b += w * w * (J.T @ r)
            while True:
                delta = np.linalg.solve(A + lam * np.eye(3), b)
                new_sol = sol - delta
                delta_norm = np.linalg.norm(delta)
                new_cost = sum(self._cost(R, t, new_sol, z) for (R, t), z in zip(rel, meas))
                if new_cost < total:
                    reduced = True; sol = new_sol; total = new_cost
                    lam = lam / 10 if lam / 10 > 1e-10 else 1e-10
                else:
                    reduced = False
                    lam = lam * 10 if lam * 10 < 1e12 else 1e12
                cont = (inner < self.inner_max) and (not reduced)
                inner += 1
                if not cont:
                    break
            inner = 0
            cont = (outer < self.outer_max) and (delta_norm > self.estimation_precision)
            outer += 1
            if not cont:
                break
        final = np.array([sol[0] / sol[2], sol[1] / sol[2], 1.0 / sol[2]])
        valid = True
        for (R, t) in rel:
            pos = R @ final + t
            if pos[2] <= 0:
                valid = False
                break
        n = len(rel)
        if total / (2 * n * n) > 4.7673e-04:
            valid = False
        if valid:
            if not self.is_initialized:
                self.position_FEJ = self.position.copy()      # feature.hpp:538-539 (the PREVIOUS estimate, literal)
            self.is_initialized = True
            self.position = Rl @ final + tl
            self.id_anchor = cam_ids[-1]
            self.invDepth = 1 / final[2]
            self.obs_anchor = np.array([final[0] * self.invDepth, final[1] * self.invDepth, 1.0])
            self.invParam = np.array([final[0] / final[2], final[1] / final[2], 1 / final[2]])
        return valid


class LarVioOracle:
    def __init__(self, cfg_raw: dict):
        r = cfg_raw
        self.features_rate = float(r["pub_frequency"]); self.imu_rate = float(r["imu_rate"])
        self.imu_img_timeTh = 1 / (2 * self.imu_rate)
        self.rotation_threshold = float(r["rotation_threshold"])
        self.translation_threshold = float(r["translation_threshold"])
        self.tracking_rate_threshold = float(r["tracking_rate_threshold"])
        self.max_track_len = int(r["max_track_len"])
        self.feature_translation_threshold = float(r["feature_translation_threshold"])
        self.td = float(r["td"])
        self.estimate_td = bool(int(r["estimate_td"])); self.estimate_extrin = bool(int(r["estimate_extrin"]))
        self.gyro_noise = float(r["noise_gyro"]) ** 2; self.acc_noise = float(r["noise_acc"]) ** 2
        self.gyro_bias_noise = float(r["noise_gyro_bias"]) ** 2; self.acc_bias_noise = float(r["noise_acc_bias"]) ** 2
        self.feature_noise = float(r["noise_feature"]) ** 2
        self.calib_imu = bool(int(r["calib_imu_instrinsic"]))
        self.LEG = 46 if self.calib_imu else 22                    # larvio.cpp:158-161
        self.Ma = np.eye(3); self.Tg = np.eye(3); self.As = np.zeros((3, 3))      # :129-131
        # T1/A1/M1 = strictly lower entries (1,0) (2,0) (2,1); T2/A2/M2 = diagonals; T3/A3 = upper (0,1) (0,2) (1,2)  (:132-155)
        self.imu_intr = np.concatenate([[0, 0, 0], [1, 1, 1], [0, 0, 0], np.zeros(9), [0, 0, 0], [1, 1, 1]]).astype(np.float64)
        P = np.zeros((self.LEG, self.LEG))
        P[0:3, 0:3] = np.eye(3) * float(r["initial_covariance_orientation"])
        P[3:6, 3:6] = np.eye(3) * float(r["initial_covariance_velocity"])
        P[6:9, 6:9] = np.eye(3) * float(r["initial_covariance_position"])
        P[9:12, 9:12] = np.eye(3) * float(r["initial_covariance_gyro_bias"])
        P[12:15, 12:15] = np.eye(3) * float(r["initial_covariance_acc_bias"])
        if self.estimate_extrin:
            P[15:18, 15:18] = np.eye(3) * float(r["initial_covariance_extrin_rot"])
            P[18:21, 18:21] = np.eye(3) * float(r["initial_covariance_extrin_trans"])
        if self.estimate_td:
            P[21, 21] = 4e-6
        if self.calib_imu:
            P[22:46, 22:46] = 1e-4 * np.eye(24)                   # :183-186
        self.P = P
        T = np.array(r["T_cam_imu"]["data"], np.float64).reshape(4, 4)
        self.imu_state = ImuState()
        self.imu_state.R_imu_cam0 = T[:3, :3].copy()              # larvio.cpp:189-202
        self.imu_state.t_cam0_imu = -T[:3, :3].T @ T[:3, 3]
        self.sw_size = int(r["sw_size"]); self.if_FEJ_config = bool(int(r["if_FEJ"]))
        self.least_obs = int(r["least_observation_number"])
        self.if_ZUPT_valid = bool(int(r["if_ZUPT_valid"])); self.zupt_max_feature_dis = float(r["zupt_max_feature_dis"])
        self.zupt_noise_v = float(r["zupt_noise_v"]) ** 2; self.zupt_noise_p = float(r["zupt_noise_p"]) ** 2
        self.zupt_noise_q = float(r["zupt_noise_q"]) ** 2
        self.max_features = max(int(r["max_features_in_one_grid"]), 0)
        self.grid_rows = int(r["aug_grid_rows"]); self.grid_cols = int(r["aug_grid_cols"])
        self.feature_idp_dim = int(r["feature_idp_dim"])
        self.use_schmidt = bool(int(r["use_schmidt"]))
        self.hybrid = self.max_features * self.grid_rows * self.grid_cols != 0
        if self.feature_idp_dim not in (1, 3):
            self.feature_idp_dim = 3                                   # larvio.cpp:270-274
        self.idp = self.feature_idp_dim
        # use_schmidt (larvio.cpp:277): poses that leave the window while they anchor SLAM features stay behind the feature block
        # as nuisance states (:2351-2358, :2569-2613): ids in covariance order, their frozen pose, the features they anchor
        self.nui_ids = []; self.nui_states = {}; self.nui_features = {}
        it = r["intrinsics"]
        fx, fy, cx, cy = float(it["fx"]), float(it["fy"]), float(it["cx"]), float(it["cy"])
        U, V = int(r["resolution_width"]), int(r["resolution_height"])
        self.x_min = -cx / fx; self.y_min = -cy / fy; self.x_max = (U - cx) / fx; self.y_max = (V - cy) / fy
        if self.grid_rows * self.grid_cols != 0:
            self.grid_width = (self.x_max - self.x_min) / self.grid_cols; self.grid_height = (self.y_max - self.y_min) / self.grid_rows
        else:
            self.grid_width = self.x_max - self.x_min; self.grid_height = self.y_max - self.y_min
        self.grid_map = {i: [] for i in range(self.grid_rows * self.grid_cols)}
        self.feature_states = []          # ids of EKF-SLAM features, in state order
        self.lost_slam_features = {}      # id -> position at removal (larvio.h:208), cleared by the getter
        self.active_slam_features = {}    # id -> latest position while in the state (larvio.h:212), cleared by the getter
        self.last_ZUPT_time = 0.0
        Qc = np.zeros((12, 12))
        Qc[0:3, 0:3] = np.eye(3) * self.gyro_noise; Qc[3:6, 3:6] = np.eye(3) * self.acc_noise
        Qc[6:9, 6:9] = np.eye(3) * self.gyro_bias_noise; Qc[9:12, 9:12] = np.eye(3) * self.acc_bias_noise
        self.Qc = Qc
        self.if_FEJ = False; self.if_ZUPT = False; self.bFirstFeatures = False
        self.is_gravity_set = False
        self.chi2 = {i: float(chi2.ppf(0.05, i)) for i in range(1, 100)}
        self.aug = {}                     # id -> AugState (iteration = sorted ids)
        self.map_server = {}              # id -> Feature
        self.FEJ_now = self.imu_state.copy(); self.FEJ_old = self.imu_state.copy()
        self.imu_old = self.imu_state.copy()
        self.m_gyro_old = None; self.m_acc_old = None
        self.next_state_id = 0
        self.tracking_rate = 0.0
        self.coarse_feature_dis = []
        self.take_off_stamp = 0.0
        self.stats = {}                   # per-call counters (rows, r, d, ...) for flop accounting
        self.zupt_events = 0

    # what FlexibleInitializer::tryIncInit leaves behind (larvio.cpp:376-386)
    def set_initial_state(self, t, q_xyzw, p, v, bg, ba):
        s = self.imu_state
        s.time = float(t); s.q = np.array(q_xyzw, float); s.p = np.array(p, float); s.v = np.array(v, float)
        s.bg = np.array(bg, float); s.ba = np.array(ba, float)
        self.is_gravity_set = True
        self.bFirstFeatures = True        # tryIncInit is only reached behind the gate of :366-372, and the initialising call goes on (:391)
        self.take_off_stamp = s.time
        self.last_ZUPT_time = s.time
        self.FEJ_now = s.copy()

    # ---- larvio.cpp:363-461.  imu: list of rows [t,w(3),a(3)] (mutated like the reference)
    def process_features(self, msg, imu: list) -> bool:
        if not self.bFirstFeatures:
            if len(imu) > 0 and imu[0][0] - msg.t - self.td <= 0.0:
                self.bFirstFeatures = True
            else:
                return False
        if not self.is_gravity_set:
            return False
        self.stats = {}
        self._batch_imu(msg.t + self.td, imu)
        self._add_observations(msg)
        self._augment()
        if self.if_ZUPT_valid:
            self.if_ZUPT = self._check_zupt()
        self._remove_lost_features()
        self._prune()
        if self.if_FEJ_config and not self.if_FEJ and self.imu_state.time - self.take_off_stamp >= 0:
            self.if_FEJ = True
        for fid in self.feature_states:                    # :455-458  active_slam_features[fid] = map_server[fid]
            self.active_slam_features[fid] = self.map_server[fid].position.copy()
        return True

    # ---- :2719-2733: both getters hand their map over and clear it
    def get_stable_map_points(self):
        out = self.lost_slam_features; self.lost_slam_features = {}
        return out

    def get_active_map_points(self):
        out = self.active_slam_features; self.active_slam_features = {}
        return out

    # ---- :464-517
    def _batch_imu(self, time_bound, imu):
        used = 0
        dt = 0.0
        for row in imu:
            t = row[0]
            if t <= self.imu_state.time:
                used += 1
                continue
            if t - time_bound > self.imu_img_timeTh:
                break
            dt = t - time_bound
            g = np.array(row[1:4], float); a = np.array(row[4:7], float)
            if self.m_gyro_old is None:
                self.m_gyro_old = g.copy(); self.m_acc_old = a.copy()
            self._process_model(t, g, a)
            used += 1
            self.m_gyro_old = g; self.m_acc_old = a
        self.imu_state.id = self.next_state_id
        self.next_state_id += 1
        self.imu_state.dt = dt
        del imu[:used]
        self.stats["n_imu"] = used

    # ---- :520-578
    def _process_model(self, time, m_gyro, m_acc):
        s = self.imu_state
        f = m_acc - s.ba; acc = self.Ma @ f
        w = m_gyro - self.As @ acc - s.bg; gyro = self.Tg @ w
        f_old = self.m_acc_old - s.ba; acc_old = self.Ma @ f_old
        w_old = self.m_gyro_old - self.As @ acc_old - s.bg; gyro_old = self.Tg @ w_old
        dtime = time - s.time
        self._predict_new_state(dtime, gyro, acc)
        Phi = self._cal_phi(dtime, f, w, acc, gyro, f_old, w_old, acc_old, gyro_old)
        C = quat_to_rot(self.imu_old.q)
        L = self.LEG
        G = np.zeros((L, 12))
        G[0:3, 0:3] = -C; G[3:6, 3:6] = -C; G[9:12, 6:9] = np.eye(3); G[12:15, 9:12] = np.eye(3)
        Q = Phi @ G @ self.Qc @ G.T @ Phi.T * dtime
        P = self.P
        P[:L, :L] = Phi @ P[:L, :L] @ Phi.T + Q
        if P.shape[0] > L:
            P[:L, L:] = Phi @ P[:L, L:]
            P[L:, :L] = P[L:, :L] @ Phi.T
        self.P = (P + P.T) / 2.0
        s.time = time
        self.FEJ_now.time = time

    # ---- :581-649
    def _predict_new_state(self, dt, gyro, acc):
        gn = np.linalg.norm(gyro)
        Om = np.zeros((4, 4))
        Om[:3, :3] = -skew(gyro); Om[:3, 3] = gyro; Om[3, :3] = -gyro
        self.imu_old = self.imu_state.copy()
        s = self.imu_state
        q, v, p = s.q, s.v, s.p
        if gn > 1e-5:
            dq_dt = (np.cos(gn * dt * 0.5) * np.eye(4) + 1 / gn * np.sin(gn * dt * 0.5) * Om) @ q
            dq_dt2 = (np.cos(gn * dt * 0.25) * np.eye(4) + 1 / gn * np.sin(gn * dt * 0.25) * Om) @ q
        else:
            dq_dt = (np.eye(4) + 0.5 * dt * Om) * np.cos(gn * dt * 0.5) @ q
            dq_dt2 = (np.eye(4) + 0.25 * dt * Om) * np.cos(gn * dt * 0.25) @ q
        dR = quat_to_rot(dq_dt); dR2 = quat_to_rot(dq_dt2)
        k1v = quat_to_rot(q) @ acc + GRAVITY; k1p = v
        k1_v = v + k1v * dt / 2
        k2v = dR2 @ acc + GRAVITY; k2p = k1_v
        k2_v = v + k2v * dt / 2
        k3v = dR2 @ acc + GRAVITY; k3p = k2_v
        k3_v = v + k3v * dt
        k4v = dR @ acc + GRAVITY; k4p = k3_v
        s.q = dq_dt / np.linalg.norm(dq_dt)
        s.v = v + dt / 6 * (k1v + 2 * k2v + 2 * k3v + k4v)
        s.p = p + dt / 6 * (k1p + 2 * k2p + 2 * k3p + k4p)
        self.FEJ_old = self.FEJ_now.copy()
        self.FEJ_now = s.copy()

    # ---- :3475-3530
    def _cal_phi(self, dtime, f, w, acc, gyro, f_old, w_old, acc_old, gyro_old):
        axis = dtime * (gyro_old + gyro) / 2 + dtime * dtime * np.cross(gyro_old, gyro) / 12
        Ah = skew(axis)
        C = quat_to_rot(self.imu_old.q)
        L = self.LEG
        Phi = np.eye(L)
        TA = self.Tg @ self.As
        if self.if_FEJ:
            vk, pk, vk1, pk1 = self.FEJ_old.v, self.FEJ_old.p, self.FEJ_now.v, self.FEJ_now.p
        else:
            vk, pk, vk1, pk1 = self.imu_old.v, self.imu_old.p, self.imu_state.v, self.imu_state.p
        g = GRAVITY
        I3 = np.eye(3)
        Phi[0:3, 9:12] = -0.5 * C @ (2 * I3 + Ah) * dtime @ self.Tg
        Phi[0:3, 12:15] = 0.5 * C @ (2 * I3 + Ah) * dtime @ TA @ self.Ma
        Phi[3:6, 0:3] = -skew(vk1 - vk - g * dtime)
        Phi[3:6, 9:12] = (skew(-pk1 + pk + vk1 * dtime - 0.5 * g * dtime * dtime) @ C +
                          skew(-0.5 * pk1 + 0.5 * pk + 0.5 * vk1 * dtime - g * dtime * dtime / 6) @ C @ Ah)
        Phi[3:6, 12:15] = -0.5 * C @ (2 * I3 + Ah) * dtime @ self.Ma - Phi[3:6, 9:12] @ TA @ self.Ma
        Phi[6:9, 0:3] = -skew(pk1 - pk - vk * dtime - 0.5 * g * dtime * dtime)
        Phi[6:9, 3:6] = I3 * dtime
        Phi[6:9, 9:12] = (-dtime * dtime * dtime * skew(g) @ C / 6 +
                          dtime * skew(pk1 - pk - g * dtime * dtime / 6) @ C @ Ah / 4)
        Phi[6:9, 12:15] = -C @ (3 * I3 + Ah) * dtime * dtime / 6 @ self.Ma - Phi[6:9, 9:12] @ TA @ self.Ma
        if not self.calib_imu:
            return Phi
        # ---- IMU-intrinsic columns (:3532-3797).  For a 3-vector x the selectors are
        #   Lo(x): (1,0)=x0 (2,1)=x0 (2,2)=x1    Di(x) = diag(x)    Up(x): (0,0)=x1 (0,1)=x2 (1,2)=x2
        def Lo(x):
            m = np.zeros((3, 3)); m[1, 0] = x[0]; m[2, 1] = x[0]; m[2, 2] = x[1]; return m

        def Di(x):
            return np.diag(x)

        def Up(x):
            m = np.zeros((3, 3)); m[0, 0] = x[1]; m[0, 1] = x[2]; m[1, 2] = x[2]; return m
        f_mid = (f + f_old) / 2; acc_mid = (acc + acc_old) / 2
        w_mid = (w_old + w) / 2 + dtime * np.cross(w_old, w) / 12
        R_mid = I3 + 0.5 * Ah; R_kp1 = I3 + Ah
        S_mid = skew(R_mid @ acc_mid); S_kp1 = skew(R_kp1 @ acc)
        Tg = self.Tg
        # (column, selector, samples at k / k+1/2 / k+1, left factor, sign of the q block, direct-v term?)
        groups = [(22, Lo, (w_old, w_mid, w), I3, +1, False), (25, Di, (w_old, w_mid, w), I3, +1, False),
                  (28, Up, (w_old, w_mid, w), I3, +1, False),
                  (31, Lo, (acc_old, acc_mid, acc), Tg, -1, False), (34, Di, (acc_old, acc_mid, acc), Tg, -1, False),
                  (37, Up, (acc_old, acc_mid, acc), Tg, -1, False),
                  (40, Lo, (f_old, f_mid, f), TA, -1, True), (43, Di, (f_old, f_mid, f), TA, -1, True)]
        for col, sel, (xk, xh, xp), Lf, sgn, direct in groups:
            kq1 = Lf @ sel(xk); kq2 = R_mid @ Lf @ sel(xh); kq4 = R_kp1 @ Lf @ sel(xp)
            Rq = dtime * (kq1 + 4 * kq2 + kq4) / 6
            Phi[0:3, col:col + 3] = sgn * C @ Rq
            if not direct:
                kv1 = np.zeros((3, 3))
                kv2 = S_mid * dtime @ kq1 / 2
                kv3 = S_mid * dtime @ kq2 / 2
                kv4 = S_kp1 @ Rq
            else:
                kv1 = sel(xk)
                kv2 = R_mid @ sel(xh) + S_mid * dtime @ kq1 / 2
                kv3 = R_mid @ sel(xh) + S_mid * dtime @ kq2 / 2
                kv4 = R_kp1 @ sel(xp) + S_kp1 @ Rq
            fR = dtime * (kv1 + 2 * kv2 + 2 * kv3 + kv4) / 6
            vs = +1 if direct else -sgn            # Phi_v: -C f for T, +C f for A, +C v for M
            Phi[3:6, col:col + 3] = vs * C @ fR
            kp1 = np.zeros((3, 3)); kp2 = dtime * kv1 / 2; kp3 = dtime * kv2 / 2; kp4 = fR
            Phi[6:9, col:col + 3] = vs * C @ (dtime * (kp1 + 2 * kp2 + 2 * kp3 + kp4) / 6)
        return Phi

    # ---- :1497-1507, 1713-1723, 2858-2868 + updateImuMx :3803-3847
    def _inject_imu_intrinsics(self, dx):
        if not self.calib_imu:
            return
        self.imu_intr = self.imu_intr + dx[22:46]
        T1, T2, T3, A1, A2, A3, M1, M2 = [self.imu_intr[3 * i:3 * i + 3] for i in range(8)]
        self.Tg = np.array([[T2[0], T3[0], T3[1]], [T1[0], T2[1], T3[2]], [T1[1], T1[2], T2[2]]])
        self.As = np.array([[A2[0], A3[0], A3[1]], [A1[0], A2[1], A3[2]], [A1[1], A1[2], A2[2]]])
        Ma = self.Ma.copy()                                      # the upper triangle of Ma is never written (:3839-3844)
        Ma[0, 0] = M2[0]; Ma[1, 0] = M1[0]; Ma[1, 1] = M2[1]; Ma[2, 0] = M1[1]; Ma[2, 1] = M1[2]; Ma[2, 2] = M2[2]
        self.Ma = Ma

    # ---- :720-801
    def _augment(self):
        s = self.imu_state
        a = AugState(s.id)
        a.time = s.time; a.dt = s.dt; a.q = s.q.copy(); a.p = s.p.copy(); a.p_FEJ = self.FEJ_now.p.copy()
        a.R_imu_cam0 = s.R_imu_cam0.copy(); a.t_cam0_imu = s.t_cam0_imu.copy()
        R_b2w = quat_to_rot(s.q)
        R_w2c = s.R_imu_cam0 @ R_b2w.T
        a.q_cam = rot_to_quat(R_w2c.T)
        a.p_cam = s.p + R_b2w @ s.t_cam0_imu
        self.aug[s.id] = a
        P = self.P
        d = P.shape[0]
        sel = [0, 1, 2, 6, 7, 8]
        P12 = P[sel, :]
        P11 = P12[:, sel]
        nf = self.idp * len(self.feature_states) + 6 * len(self.nui_ids)
        pe = d - nf                        # end of the pose block; SLAM features and nuisance states follow (larvio.cpp:768-793)
        order = list(range(pe)) + list(range(d, d + 6)) + list(range(pe, d))
        Pn = np.zeros((d + 6, d + 6))
        Pn[:d, :d] = P
        Pn[d:, :d] = P12; Pn[:d, d:] = P12.T; Pn[d:, d:] = P11
        Pn = Pn[np.ix_(order, order)]
        self.P = (Pn + Pn.T) / 2.0

    # ---- :804-856
    def _add_observations(self, msg):
        sid = self.imu_state.id
        curr_num = len(self.map_server)
        tracked = 0
        dt = self.imu_state.dt
        for fid, f in zip(msg.ids, msg.data):
            fid = int(fid)
            u, v, u_init, v_init, u_vel, v_vel, u_init_vel, v_init_vel = f
            if fid not in self.map_server:
                ft = Feature(fid, self.feature_translation_threshold)
                self.map_server[fid] = ft
                ft.obs[sid] = np.array([u + u_vel * dt, v + v_vel * dt]); ft.obs_vel[sid] = np.array([u_vel, v_vel])
                ft.totalObsNum += 1
                if not (u_init == -1 and v_init == -1) and (sid - 1) in self.aug:
                    dt_ = self.aug[sid - 1].dt
                    ft.obs[sid - 1] = np.array([u_init + u_init_vel * dt_, v_init + v_init_vel * dt_])
                    ft.obs_vel[sid - 1] = np.array([u_init_vel, v_init_vel])
                    ft.totalObsNum += 1
            else:
                ft = self.map_server[fid]
                ft.obs[sid] = np.array([u + u_vel * dt, v + v_vel * dt]); ft.obs_vel[sid] = np.array([u_vel, v_vel])
                ft.totalObsNum += 1
                tracked += 1
                if self.if_ZUPT_valid and (sid - 1) in ft.obs:
                    self.coarse_feature_dis.append(float(np.linalg.norm(np.array([u, v]) - ft.obs[sid - 1])))
        with np.errstate(divide="ignore", invalid="ignore"):
            self.tracking_rate = float(np.float64(tracked) / np.float64(curr_num))

    # ---- checkZUPT :2751-2788 (pure MSCKF: no SLAM features to drop)
    def _check_zupt(self):
        d = self.coarse_feature_dis
        self.coarse_feature_dis = []
        if len(d) < 20:
            return False
        d = sorted(d)
        if d[-9] < self.zupt_max_feature_dis:
            self.zupt_events += 1
            if self.feature_states:                       # :2770-2782
                nf = self.idp * len(self.feature_states)
                self.P = self.P[:-nf, :-nf]
                for fid in self.feature_states:
                    ft = self.map_server[fid]
                    ft.is_initialized = False; ft.ekf_feature = False; ft.in_state = False
                self.feature_states = []
            self._zupt_update()
            self.last_ZUPT_time = self.imu_state.time
            return True
        return False

    # ---- measurementUpdate_ZUPT_vpq :2791-2962
    def _zupt_update(self):
        N = len(self.aug)
        d = self.P.shape[1]
        L = self.LEG
        H = np.zeros((9, d))
        H[0:3, 3:6] = np.eye(3)
        H[3:6, L + 6 * N - 3:L + 6 * N] = np.eye(3)          # (SLAM features were dropped: d == L + 6N)
        H[3:6, L + 6 * N - 9:L + 6 * N - 6] = -np.eye(3)
        H[6:9, L + 6 * N - 6:L + 6 * N - 3] = -0.5 * np.eye(3)
        H[6:9, L + 6 * N - 12:L + 6 * N - 9] = 0.5 * np.eye(3)
        sid = self.imu_state.id
        cur, prv = self.aug[sid], self.aug[sid - 1]
        r = np.zeros(9)
        r[0:3] = -self.imu_state.v
        r[3:6] = -(cur.p - prv.p)
        qp_conj = np.array([-prv.q[0], -prv.q[1], -prv.q[2], prv.q[3]])
        r[6:9] = quat_mul(cur.q, qp_conj)[:3]
        Rz = np.diag([self.zupt_noise_v] * 3 + [self.zupt_noise_p] * 3 + [self.zupt_noise_q] * 3)
        self._update(H, r, "zupt", Rz)

    # ---- :859-921
    def _meas_jacobian(self, sid, ft):
        a = self.aug[sid]
        R_b2c = a.R_imu_cam0; t_c_b = a.t_cam0_imu
        R_b2w = quat_to_rot(a.q); R_w2b = R_b2w.T
        R_w2c = R_b2c @ R_w2b
        t_c_w = a.p + R_b2w @ t_c_b
        p_w = ft.position
        z = ft.obs[sid]
        p_c = R_w2c @ (p_w - t_c_w)
        p_bf_w = (p_w - a.p_FEJ) if self.if_FEJ else (p_w - a.p)
        dz = np.zeros((2, 3))
        dz[0, 0] = 1 / p_c[2]; dz[1, 1] = 1 / p_c[2]
        dz[0, 2] = -p_c[0] / (p_c[2] * p_c[2]); dz[1, 2] = -p_c[1] / (p_c[2] * p_c[2])
        dxb = np.zeros((3, 6)); dxb[:, :3] = R_w2c @ skew(p_bf_w); dxb[:, 3:] = -R_w2c
        dxe = np.zeros((3, 6)); dxe[:, :3] = R_w2c @ skew(p_bf_w) @ R_b2w - R_b2c @ skew(t_c_b); dxe[:, 3:] = -R_b2c
        H_x = dz @ dxb; H_e = dz @ dxe; H_f = dz @ R_w2c
        r = z - np.array([p_c[0] / p_c[2], p_c[1] / p_c[2]])
        return H_x, H_e, H_f, r

    # ---- :924-981
    def _feature_jacobian(self, ft, state_ids):
        valid = [sid for sid in state_ids if sid in ft.obs]
        rows = 2 * len(valid)
        d = self.P.shape[1]
        Hx = np.zeros((rows, d)); Hf = np.zeros((rows, 3)); r = np.zeros(rows)
        order = sorted(self.aug.keys())
        k = 0
        for sid in valid:
            H_xi, H_ei, H_fi, r_i = self._meas_jacobian(sid, ft)
            cntr = order.index(sid)
            Hx[k:k + 2, self.LEG + 6 * cntr:self.LEG + 6 * cntr + 6] = H_xi
            Hx[k:k + 2, 15:21] = H_ei
            if self.estimate_td:
                Hx[k:k + 2, 21] = ft.obs_vel[sid]
            Hf[k:k + 2] = H_fi
            r[k:k + 2] = r_i
            k += 2
        U, _, _ = np.linalg.svd(Hf, full_matrices=True)
        A = U[:, 3:]
        return A.T @ Hx, A.T @ r

    # ---- :1865-1880
    def _gating(self, H, r, dof):
        S = H @ self.P @ H.T + self.feature_noise * np.eye(H.shape[0])
        gamma = float(r @ np.linalg.solve(S, r))
        self.stats.setdefault("gates", []).append((dof, gamma, self.chi2.get(dof, 0.0)))      # diagnostics only
        return gamma < self.chi2.get(dof, 0.0)

    def _compress(self, H, r, cols):
        if H.shape[0] == 0 or H.shape[0] <= H.shape[1]:
            return H, r
        Q, R = np.linalg.qr(H, mode="reduced")
        return R[:cols], (Q.T @ r)[:cols]

    # ---- removeLostFeatures :1883-2256
    def _remove_lost_features(self):
        sid_now = self.imu_state.id
        invalid = []; msckf_ids = []; lost_ids = []
        ekf_new_ids = []; ekf_lost = []; ekf_ids = []
        for fid in sorted(self.map_server.keys()):
            ft = self.map_server[fid]
            if ft.in_state:
                (ekf_ids if sid_now in ft.obs else ekf_lost).append(fid)
        self.stats["n_ekf_lost"] = len(ekf_lost)
        self._rm_lost_features_cov(ekf_lost)
        if self.use_schmidt:
            self._rm_useless_nuisance()                             # :1920-1921
        self._update_grid_map()
        for fid in sorted(self.map_server.keys()):
            ft = self.map_server[fid]
            if ft.in_state:
                continue
            tracked_now = sid_now in ft.obs
            if not tracked_now:
                if len(ft.obs) < self.least_obs:
                    invalid.append(fid); continue
                if not ft.is_initialized:
                    if not ft.check_motion(self.aug, tracked_now):
                        invalid.append(fid); continue
                    if not ft.initialize_position(self.aug, sid_now):
                        invalid.append(fid); continue
                msckf_ids.append(fid); lost_ids.append(fid)
            else:
                if not (len(ft.obs) >= self.max_track_len):
                    continue
                code = self._grid_code(ft.obs[sid_now]) if self.hybrid else 0
                if (self.hybrid and len(self.grid_map.get(code, [])) < self.max_features and
                        self.imu_state.time - self.last_ZUPT_time > 5 and
                        (len(self.feature_states) + len(ekf_new_ids)) < self.max_features * self.grid_rows * self.grid_cols):
                    if not ft.ekf_feature:
                        ft.is_initialized = False
                        if ft.check_motion(self.aug, tracked_now):
                            ft.initialize_inv_param(self.aug, sid_now)
                    if not ft.is_initialized:
                        continue
                    ekf_new_ids.append(fid)
                    self.grid_map.setdefault(code, []).append(fid)
                else:
                    if not ft.is_initialized:
                        if ft.check_motion(self.aug, tracked_now):
                            ft.initialize_position(self.aug, sid_now)
                    if not ft.is_initialized:
                        continue
                    msckf_ids.append(fid); lost_ids.append(fid)
        for fid in invalid:
            del self.map_server[fid]
        self.stats["n_msckf_features"] = len(msckf_ids); self.stats["n_ekf_new"] = len(ekf_new_ids); self.stats["n_ekf"] = len(ekf_ids)
        if len(msckf_ids) == 0 and len(ekf_new_ids) == 0 and len(ekf_ids) == 0:
            return
        if not self.if_ZUPT:
            d = self.P.shape[1]
            # ---- new EKF-SLAM features (:2019-2125)
            for fid in ekf_new_ids:
                self.map_server[fid].in_state = True
                self.feature_states.append(fid)
            blocks = []
            for fid in list(ekf_new_ids):
                ft = self.map_server[fid]
                sids = sorted(ft.obs.keys())
                Hj, rj = self._feature_jacobian_ekf_new(ft, sids)
                Hm, rm = self._feature_jacobian(ft, sids)
                if self._gating(Hm, rm, 2 * len(sids) - 3):
                    blocks.append((fid, Hj, rj))
                else:
                    ft.in_state = False
            # drop the features that failed the gate: their state columns disappear
            kept = [b[0] for b in blocks]
            old_fs = list(self.feature_states)
            n_old = len(old_fs) - len(ekf_new_ids)
            self.feature_states = old_fs[:n_old] + kept
            idp = self.idp
            if kept:
                keep_cols = list(range(d)) + [d + idp * ekf_new_ids.index(fid) + c for fid in kept for c in range(idp)]
                H_new = np.concatenate([b[1][:, keep_cols] for b in blocks]); r_new = np.concatenate([b[2] for b in blocks])
                n_new = idp * len(kept)                         # new state COLUMNS (sz_new)
                Hf = H_new[:, d:]
                U_, _, _ = np.linalg.svd(Hf, full_matrices=True)
                Vn = U_[:, n_new:]
                Q, _ = np.linalg.qr(Hf, mode="complete")
                W = np.concatenate([Vn, Q[:, :n_new]], 1)
                H_new = W.T @ H_new; r_new = W.T @ r_new
            else:
                H_new = np.zeros((0, d)); r_new = np.zeros(0); n_new = 0
            # ---- in-state EKF-SLAM features (:2127-2177)
            Hs = []; rs = []
            for fid in ekf_ids:
                Hj, rj = self._feature_jacobian_ekf(self.map_server[fid])
                if self._gating(Hj, rj, 2):
                    Hs.append(Hj); rs.append(rj)
            H_ekf = np.concatenate(Hs) if Hs else np.zeros((0, d)); r_ekf = np.concatenate(rs) if rs else np.zeros(0)
            H_ekf, r_ekf = self._compress(H_ekf, r_ekf, d)
            # ---- MSCKF features (:2179-2233)
            cols = self.LEG + 6 * len(self.aug)
            Hs = []; rs = []; m_hist = []
            for fid in msckf_ids:
                ft = self.map_server[fid]
                sids = sorted(ft.obs.keys())
                Hj, rj = self._feature_jacobian(ft, sids)
                if self._gating(Hj, rj, 2 * len(sids) - 3):
                    Hs.append(Hj[:, :cols]); rs.append(rj); m_hist.append(len(sids))
            H = np.concatenate(Hs) if Hs else np.zeros((0, cols)); r = np.concatenate(rs) if rs else np.zeros(0)
            self.stats["rows_msckf"] = H.shape[0]; self.stats["m_hist"] = m_hist
            H, r = self._compress(H, r, cols)
            H_msckf = np.zeros((H.shape[0], d)); H_msckf[:, :H.shape[1]] = H
            self._update_hybrid(H_new, r_new, n_new, H_ekf, r_ekf, H_msckf, r)
        else:
            for fid in msckf_ids:
                self.map_server[fid].is_initialized = False
        for fid in lost_ids:
            del self.map_server[fid]

    # ---- measurementUpdate_hybrid :1605-1862 (no Schmidt)
    def _update_hybrid(self, H_new, r_new, n_new, H_ekf, r_ekf, H_msckf, r_msckf):
        d = self.P.shape[1]
        sz_r = len(r_new) + len(r_ekf) + len(r_msckf)
        if sz_r == 0:
            return
        k = H_new.shape[0] - n_new
        H_o = np.concatenate([H_msckf, H_ekf, H_new[:k, :d]]); r_o = np.concatenate([r_msckf, r_ekf, r_new[:k]])
        H_1 = H_new[k:, :d]; H_2 = H_new[k:, d:]; r_1 = r_new[k:]
        P = self.P
        S = H_o @ P @ H_o.T + self.feature_noise * np.eye(H_o.shape[0])
        K = np.linalg.solve(S, H_o @ P).T if H_o.shape[0] else np.zeros((d, 0))
        dx_leg = K @ r_o if H_o.shape[0] else np.zeros(d)
        self.stats.setdefault("updates", []).append(dict(tag="hybrid", r=H_o.shape[0], d=d, n_new=n_new))
        if n_new:
            h2 = np.diag(H_2).copy()           # Eigen LDLT reads the lower triangle of the triangular factor (App. C-13)
            HH = H_1 / h2[:, None]
            dx_new = -HH @ dx_leg + r_1 / h2
            dx = np.concatenate([dx_leg, dx_new])
        else:
            dx = dx_leg
        s = self.imu_state
        s.q = quat_mul(small_angle_quat(dx[0:3]), s.q)
        s.v = s.v + dx[3:6]; s.p = s.p + dx[6:9]; s.bg = s.bg + dx[9:12]; s.ba = s.ba + dx[12:15]
        s.R_imu_cam0 = s.R_imu_cam0 @ quat_to_rot(small_angle_quat(dx[15:18])).T
        s.t_cam0_imu = s.t_cam0_imu + dx[18:21]
        self.td += dx[21]
        self._inject_imu_intrinsics(dx)
        for i, sid in enumerate(sorted(self.aug.keys())):
            a = self.aug[sid]
            da = dx[self.LEG + 6 * i:self.LEG + 6 * i + 6]
            a.q = quat_mul(small_angle_quat(da[0:3]), a.q)
            a.p = a.p + da[3:6]
            R_b2w = quat_to_rot(a.q)
            a.q_cam = rot_to_quat(R_b2w @ s.R_imu_cam0.T)
            a.p_cam = a.p + R_b2w @ s.t_cam0_imu
        base = self.LEG + 6 * len(self.aug)
        num_old = len(self.feature_states) - n_new // self.idp
        self._update_feature_states(dx, base, d, num_old)
        nn = 6 * len(self.nui_ids) if self.use_schmidt else 0
        if H_o.shape[0]:
            I_KH = np.eye(d) - K @ H_o
            if nn:                                               # :1805-1814
                n0 = base + self.idp * num_old
                P_nui = P[n0:n0 + nn, n0:n0 + nn].copy()
                P = I_KH @ P
                P[n0:n0 + nn, n0:n0 + nn] = P_nui
            else:
                P = I_KH @ P
            P = (P + P.T) / 2.0
        if n_new:
            nHHP = -HH @ P
            H22 = H_2.T @ H_2                                    # :1823-1825 (the FULL triangular factor here, unlike HH)
            P22 = -nHHP @ HH.T + self.feature_noise * np.linalg.solve(H22, np.eye(n_new))
            Pn = np.zeros((d + n_new, d + n_new))
            Pn[:d, :d] = P; Pn[d:, :d] = nHHP; Pn[:d, d:] = nHHP.T; Pn[d:, d:] = P22
            if nn:                                               # :1832-1845: the new columns go in FRONT of the nuisance block
                order = list(range(d - nn)) + list(range(d, d + n_new)) + list(range(d - nn, d))
                Pn = Pn[np.ix_(order, order)]
            P = (Pn + Pn.T) / 2.0
        self.P = P

    # ---- :1420-1602 / :1605-1862 with empty SLAM blocks
    def _update(self, H, r, tag, Rn=None):
        if H.shape[0] == 0 or r.shape[0] == 0:
            return
        P = self.P
        S = H @ P @ H.T + (self.feature_noise * np.eye(H.shape[0]) if Rn is None else Rn)
        Kt = np.linalg.solve(S, H @ P)
        K = Kt.T
        dx = K @ r
        self.stats.setdefault("updates", []).append(dict(tag=tag, r=H.shape[0], d=P.shape[0]))
        s = self.imu_state
        s.q = quat_mul(small_angle_quat(dx[0:3]), s.q)
        s.v = s.v + dx[3:6]; s.p = s.p + dx[6:9]; s.bg = s.bg + dx[9:12]; s.ba = s.ba + dx[12:15]
        dqe = small_angle_quat(dx[15:18])
        s.R_imu_cam0 = s.R_imu_cam0 @ quat_to_rot(dqe).T
        s.t_cam0_imu = s.t_cam0_imu + dx[18:21]
        self.td += dx[21]
        self._inject_imu_intrinsics(dx)
        for i, sid in enumerate(sorted(self.aug.keys())):
            a = self.aug[sid]
            da = dx[self.LEG + 6 * i:self.LEG + 6 * i + 6]
            a.q = quat_mul(small_angle_quat(da[0:3]), a.q)
            a.p = a.p + da[3:6]
            R_b2w = quat_to_rot(a.q)
            a.q_cam = rot_to_quat(R_b2w @ s.R_imu_cam0.T)
            a.p_cam = a.p + R_b2w @ s.t_cam0_imu
        self._update_feature_states(dx, self.LEG + 6 * len(self.aug))
        I_KH = np.eye(K.shape[0]) - K @ H
        if self.use_schmidt and self.nui_ids:          # :1579-1589 / :2940-2950: the nuisance block keeps its prior
            n0 = self.LEG + 6 * len(self.aug) + self.idp * len(self.feature_states); n1 = n0 + 6 * len(self.nui_ids)
            P_nui = P[n0:n1, n0:n1].copy()
            P = I_KH @ P
            P[n0:n1, n0:n1] = P_nui
        else:
            P = I_KH @ P
        self.P = (P + P.T) / 2.0

    # inverse-depth correction of the in-state features + recomputed world positions (:1536-1575, :1752-1801)
    # new_at / num_old: in measurementUpdate_hybrid the corrections of the features added by this update sit behind the old
    # covariance (:1777-1793), i.e. behind the nuisance block when there is one
    def _update_feature_states(self, dx, base, new_at=None, num_old=None):
        for i, fid in enumerate(self.feature_states):
            ft = self.map_server[fid]
            a = self._anchor_state(ft)
            at = base + self.idp * i if (num_old is None or i < num_old) else new_at + self.idp * (i - num_old)
            if self.idp == 3:
                ft.invParam = ft.invParam + dx[at:at + 3]
                p_c = np.array([ft.invParam[0] / ft.invParam[2], ft.invParam[1] / ft.invParam[2], 1 / ft.invParam[2]])
            else:
                ft.invDepth += dx[at]
                p_c = np.array([ft.obs_anchor[0] / ft.invDepth, ft.obs_anchor[1] / ft.invDepth, 1 / ft.invDepth])
            ft.position = quat_to_rot(a.q_cam) @ p_c + a.p_cam

    # the anchor pose of an in-state feature: a window state, or with use_schmidt a nuisance state (:1000-1010, :1543-1549)
    def _anchor_state(self, ft):
        return self.aug[ft.id_anchor] if ft.id_anchor in self.aug else self.nui_states[ft.id_anchor]

    # ---- measurementJacobian_ekf_3didp :984-1114
    def _meas_jacobian_3didp(self, sid, ft):
        k = self.aug[sid]; a = self._anchor_state(ft)
        nui = ft.id_anchor not in self.aug                       # :999-1010: a nuisance anchor is used as frozen (no FEJ, its own camera pose)
        R_b2c = k.R_imu_cam0; t_c_b = k.t_cam0_imu
        R_bk2w = quat_to_rot(k.q); R_w2bk = R_bk2w.T
        R_w2ck = R_b2c @ R_w2bk; t_ck_w = k.p + R_bk2w @ t_c_b
        R_ba2w = quat_to_rot(a.q); R_w2ba = R_ba2w.T
        R_w2ca = quat_to_rot(a.q_cam).T if nui else R_b2c @ R_w2ba           # :1032-1038
        f_ca = ft.invParam
        if self.if_FEJ and not nui:
            p_ca = R_b2c @ (R_w2ba @ (ft.position_FEJ - a.p_FEJ) - t_c_b)
        else:
            p_ca = np.array([f_ca[0] / f_ca[2], f_ca[1] / f_ca[2], 1 / f_ca[2]])
        p_w = ft.position
        z = ft.obs[sid]
        p_ck = R_w2ck @ (p_w - t_ck_w)
        r = z - np.array([p_ck[0] / p_ck[2], p_ck[1] / p_ck[2]])
        if sid == ft.id_anchor:                                  # :1065-1073
            H_f = np.zeros((2, 3)); H_f[0, 0] = 1; H_f[1, 1] = 1
            return H_f, np.zeros((2, 6)), np.zeros((2, 6)), np.zeros((2, 6)), r
        J_k = np.zeros((2, 3))
        J_k[0, 0] = 1 / p_ck[2]; J_k[1, 1] = 1 / p_ck[2]
        J_k[0, 2] = -p_ck[0] / (p_ck[2] * p_ck[2]); J_k[1, 2] = -p_ck[1] / (p_ck[2] * p_ck[2])
        J_p = R_w2ck @ R_w2ca.T
        p_baf_w = (ft.position_FEJ - a.p_FEJ) if (self.if_FEJ and not nui) else (p_w - a.p)
        p_bkf_w = (ft.position_FEJ - k.p_FEJ) if self.if_FEJ else (p_w - k.p)
        J_xa = np.zeros((3, 6)); J_xa[:, :3] = -R_w2ck @ skew(p_baf_w); J_xa[:, 3:] = R_w2ck
        J_xk = np.zeros((3, 6)); J_xk[:, :3] = R_w2ck @ skew(p_bkf_w); J_xk[:, 3:] = -R_w2ck
        J_e = np.zeros((3, 6))
        Sk = skew(R_w2bk @ p_bkf_w - t_c_b)
        Mx = R_w2bk @ R_w2ba.T @ skew(R_b2c.T @ p_ca)
        J_e[:, :3] = R_b2c @ (Sk - Mx); J_e[:, 3:] = R_b2c @ (R_w2bk @ R_w2ba.T - np.eye(3))
        J_f = np.eye(3)
        J_f[0, 2] = -f_ca[0] / f_ca[2]; J_f[1, 2] = -f_ca[1] / f_ca[2]; J_f[2, 2] = -1 / f_ca[2]
        J_f = J_f / f_ca[2]
        return J_k @ J_p @ J_f, J_k @ J_xa, J_k @ J_xk, J_k @ J_e, r

    def _meas_jacobian_idp(self, sid, ft):
        """(H_f [2 x idp], H_a, H_x, H_e, r) of one observation of an in-state feature."""
        if self.idp == 3:
            return self._meas_jacobian_3didp(sid, ft)
        H_f, H_a, H_x, H_e, r = self._meas_jacobian_1didp(sid, ft)
        return H_f.reshape(2, 1), H_a, H_x, H_e, r

    # ---- measurementJacobian_ekf_1didp :1117-1244
    def _meas_jacobian_1didp(self, sid, ft):
        k = self.aug[sid]; a = self._anchor_state(ft)
        nui = ft.id_anchor not in self.aug                       # :1132-1143
        R_b2c = k.R_imu_cam0; t_c_b = k.t_cam0_imu
        f_an = ft.obs_anchor
        R_bk2w = quat_to_rot(k.q); R_w2bk = R_bk2w.T
        R_w2ck = R_b2c @ R_w2bk; t_ck_w = k.p + R_bk2w @ t_c_b
        R_ba2w = quat_to_rot(a.q); R_w2ba = R_ba2w.T
        R_w2ca = quat_to_rot(a.q_cam).T if nui else R_b2c @ R_w2ba           # :1168-1174
        if self.if_FEJ and not nui:
            p_ca = R_b2c @ (R_w2ba @ (ft.position_FEJ - a.p_FEJ) - t_c_b)
        else:
            p_ca = np.array([f_an[0] / ft.invDepth, f_an[1] / ft.invDepth, 1 / ft.invDepth])
        p_w = ft.position
        z = ft.obs[sid]
        p_ck = R_w2ck @ (p_w - t_ck_w)
        r = z - np.array([p_ck[0] / p_ck[2], p_ck[1] / p_ck[2]])
        J_k = np.zeros((2, 3))
        J_k[0, 0] = 1 / p_ck[2]; J_k[1, 1] = 1 / p_ck[2]
        J_k[0, 2] = -p_ck[0] / (p_ck[2] * p_ck[2]); J_k[1, 2] = -p_ck[1] / (p_ck[2] * p_ck[2])
        J_d = R_w2ck @ R_w2ca.T @ f_an
        p_baf_w = (ft.position_FEJ - a.p_FEJ) if (self.if_FEJ and not nui) else (p_w - a.p)
        p_bkf_w = (ft.position_FEJ - k.p_FEJ) if self.if_FEJ else (p_w - k.p)
        J_xa = np.zeros((3, 6)); J_xa[:, :3] = -R_w2ck @ skew(p_baf_w); J_xa[:, 3:] = R_w2ck
        J_xk = np.zeros((3, 6)); J_xk[:, :3] = R_w2ck @ skew(p_bkf_w); J_xk[:, 3:] = -R_w2ck
        J_e = np.zeros((3, 6))
        Sk = skew(R_w2bk @ p_bkf_w - t_c_b)
        Mx = R_w2bk @ R_w2ba.T @ skew(R_b2c.T @ p_ca)
        J_e[:, :3] = R_b2c @ (Sk - Mx); J_e[:, 3:] = R_b2c @ (R_w2bk @ R_w2ba.T - np.eye(3))
        J_rho = -1 / (ft.invDepth * ft.invDepth)
        return (J_k @ J_d * J_rho), J_k @ J_xa, J_k @ J_xk, J_k @ J_e, r

    # ---- featureJacobian_ekf_new :1247-1338 (columns: current state + one per feature in feature_states)
    def _feature_jacobian_ekf_new(self, ft, state_ids):
        idp = self.idp
        # the anchor's own observation is not used with 1-D inverse depth (:1260-1262)
        valid = [sid for sid in state_ids if sid in ft.obs and not (idp == 1 and sid == ft.id_anchor)]
        ncol = self.LEG + 6 * len(self.aug) + idp * len(self.feature_states) + 6 * len(self.nui_ids)
        H = np.zeros((2 * len(valid), ncol)); r = np.zeros(2 * len(valid))
        order = sorted(self.aug.keys())
        a_idx = self.LEG + 6 * order.index(ft.id_anchor)
        # the features this call adds are not in the covariance yet; their columns follow it, i.e. the nuisance block (:1291-1300)
        d = self.P.shape[0]
        num_old = len(self.feature_states) - (ncol - d) // idp
        f_idx = d + idp * (self.feature_states.index(ft.id) - num_old)
        k = 0
        for sid in valid:
            H_f, H_a, H_x, H_e, r_i = self._meas_jacobian_idp(sid, ft)
            H[k:k + 2, f_idx:f_idx + idp] = H_f
            H[k:k + 2, a_idx:a_idx + 6] = H_a
            c = self.LEG + 6 * order.index(sid)
            H[k:k + 2, c:c + 6] = H_x
            H[k:k + 2, 15:21] = H_e
            if self.estimate_td:
                H[k:k + 2, 21] = ft.obs_vel[sid]
            r[k:k + 2] = r_i
            k += 2
        return H, r

    # ---- featureJacobian_ekf :1341-1417
    def _feature_jacobian_ekf(self, ft):
        sid = self.imu_state.id
        order = sorted(self.aug.keys())
        H = np.zeros((2, self.P.shape[1]))
        H_f, H_a, H_x, H_e, r = self._meas_jacobian_idp(sid, ft)
        f_idx = self.LEG + 6 * len(self.aug) + self.idp * self.feature_states.index(ft.id)
        H[:, f_idx:f_idx + self.idp] = H_f
        if self.use_schmidt and ft.id_anchor in self.nui_ids:     # :1351-1366: the anchor's columns are in the nuisance block
            num_new = (self.LEG + 6 * len(self.aug) + self.idp * len(self.feature_states) + 6 * len(self.nui_ids) - self.P.shape[0]) // self.idp
            a_idx = (self.LEG + 6 * len(self.aug) + self.idp * (len(self.feature_states) - num_new) + 6 * self.nui_ids.index(ft.id_anchor))
        else:
            a_idx = self.LEG + 6 * order.index(ft.id_anchor)
        H[:, a_idx:a_idx + 6] = H_a
        c = self.LEG + 6 * order.index(sid)
        H[:, c:c + 6] = H_x
        H[:, 15:21] = H_e
        if self.estimate_td:
            H[:, 21] = ft.obs_vel[sid]
        return H, r

    # ---- rmLostFeaturesCov :3296-3348
    def _rm_lost_features_cov(self, lost_ids):
        for fid in lost_ids:
            seq = self.feature_states.index(fid)
            i0 = self.LEG + 6 * len(self.aug) + self.idp * seq
            keep = [i for i in range(self.P.shape[0]) if not (i0 <= i < i0 + self.idp)]
            self.P = self.P[np.ix_(keep, keep)]
            self.feature_states.pop(seq)
            if self.use_schmidt:                                                     # :3329-3339
                an = self.map_server[fid].id_anchor
                if an in self.nui_ids:
                    self.nui_features[an].remove(fid)
            self.lost_slam_features[fid] = self.map_server[fid].position.copy()      # :3342
            del self.map_server[fid]

    # ---- rmUselessNuisanceState :3850-3895: nuisance states that anchor no feature any more leave the covariance
    def _rm_useless_nuisance(self):
        for nid in [i for i in self.nui_ids if len(self.nui_features.get(i, [])) == 0]:
            seq = self.nui_ids.index(nid)
            n0 = self.LEG + 6 * len(self.aug) + self.idp * len(self.feature_states) + 6 * seq
            keep = [i for i in range(self.P.shape[0]) if not (n0 <= i < n0 + 6)]
            self.P = self.P[np.ix_(keep, keep)]
            self.nui_ids.pop(seq); del self.nui_states[nid]; self.nui_features.pop(nid, None)

    # ---- updateGridMap :3351-3370
    def _grid_code(self, xy):
        row = int((xy[1] - self.y_min) / self.grid_height); col = int((xy[0] - self.x_min) / self.grid_width)
        return row * self.grid_cols + col

    def _update_grid_map(self):
        if self.grid_rows * self.grid_cols == 0:
            return
        # :3355-3357 clears the rows*cols cells only: a cell whose code falls OUTSIDE that range (observations beyond the image
        # border, e.g. row == grid_rows) is created by operator[] below or at :1972/:1988 and never emptied again, so it stays
        # "occupied" for the rest of the run (pinned against the compiled reference, tests/golden/ref_*.npz)
        for i in range(self.grid_rows * self.grid_cols):
            self.grid_map[i] = []
        for fid in self.feature_states:
            code = self._grid_code(self.map_server[fid].obs[self.imu_state.id])
            self.grid_map.setdefault(code, []).append(fid)

    # ---- updateFeatureCov_3didp :2965-3122.  Literal restatement INCLUDING the reference's slip at :3000 and :3066: the
    # "new" pose and its column block are looked up with old_state_id, so H_x_new overwrites H_x_old in the old block and
    # the Jacobian never touches the new anchor's columns.
    def _update_feature_cov_3didp(self, ft, old_id, new_id):
        N = len(self.aug)
        p_w = ft.position
        R_b2c = self.imu_state.R_imu_cam0; t_c_b = self.imu_state.t_cam0_imu
        o = self.aug[old_id]
        R_b2w_old = quat_to_rot(o.q); R_c2w_old = quat_to_rot(o.q_cam)
        if self.if_FEJ:
            p_old = R_b2c @ (R_b2w_old.T @ (ft.position_FEJ - o.p_FEJ) - t_c_b)
        else:
            p_old = R_c2w_old.T @ (p_w - o.p_cam)
        n = self.aug[old_id]                                       # sic (:3000)
        R_b2w_new = quat_to_rot(n.q); R_w2b_new = R_b2w_new.T
        R_w2c_new = quat_to_rot(n.q_cam).T
        inv_new = ft.invParam
        if self.if_FEJ:
            p_bf_old = ft.position_FEJ - o.p_FEJ; p_bf_new = ft.position_FEJ - n.p_FEJ
        else:
            p_bf_old = p_w - o.p; p_bf_new = p_w - n.p
        J_fp_new = np.eye(3)
        J_fp_new[0, 2] = -inv_new[0]; J_fp_new[1, 2] = -inv_new[1]; J_fp_new[2, 2] = -inv_new[2]
        J_fp_new = inv_new[2] * J_fp_new
        J_p = R_w2c_new @ R_c2w_old
        J_x_old = np.zeros((3, 6)); J_x_old[:, :3] = -R_w2c_new @ skew(p_bf_old); J_x_old[:, 3:] = R_w2c_new
        J_x_new = np.zeros((3, 6)); J_x_new[:, :3] = R_w2c_new @ skew(p_bf_new); J_x_new[:, 3:] = -R_w2c_new
        J_e = np.zeros((3, 6))
        Sk = skew(R_w2b_new @ p_bf_new - t_c_b)
        Mx = R_w2b_new @ R_b2w_old @ skew(R_b2c.T @ p_old)
        J_e[:, :3] = R_b2c @ (Sk - Mx); J_e[:, 3:] = R_b2c @ (R_w2b_new @ R_b2w_old - np.eye(3))
        J_pf_old = np.eye(3)
        J_pf_old[0, 2] = -p_old[0]; J_pf_old[1, 2] = -p_old[1]; J_pf_old[2, 2] = -p_old[2]
        J_pf_old = p_old[2] * J_pf_old
        H_f_new = J_fp_new @ J_p @ J_pf_old
        H_x_old = J_fp_new @ J_x_old; H_x_new = J_fp_new @ J_x_new; H_e = J_fp_new @ J_e
        J = np.zeros((3, self.P.shape[1]))
        order = sorted(self.aug.keys())
        oc = order.index(old_id); nc = order.index(old_id)         # sic (:3066)
        fc = self.feature_states.index(ft.id)
        fi = self.LEG + 6 * N + 3 * fc
        J[:, fi:fi + 3] = H_f_new
        J[:, self.LEG + 6 * oc:self.LEG + 6 * oc + 6] = H_x_old
        J[:, self.LEG + 6 * nc:self.LEG + 6 * nc + 6] = H_x_new
        J[:, 15:21] = H_e
        Pfl = J @ self.P
        Pff = Pfl @ J.T
        P = self.P
        left = Pfl[:, :fi].copy(); right = Pfl[:, fi + 3:].copy()
        P[fi:fi + 3, fi:fi + 3] = Pff
        P[fi:fi + 3, :fi] = left; P[:fi, fi:fi + 3] = left.T
        P[fi:fi + 3, fi + 3:] = right; P[fi + 3:, fi:fi + 3] = right.T
        self.P = (P + P.T) / 2.0

    # ---- updateFeatureCov_1didp :3125-3293
    def _update_feature_cov_1didp(self, ft, old_id, new_id):
        N = len(self.aug)
        p_w = ft.position
        R_b2c = self.imu_state.R_imu_cam0; t_c_b = self.imu_state.t_cam0_imu
        o = self.aug[old_id]; n = self.aug[new_id]
        R_b2w_old = quat_to_rot(o.q); R_c2w_old = quat_to_rot(o.q_cam)
        if self.if_FEJ:
            p_old = R_b2c @ (R_b2w_old.T @ (ft.position_FEJ - o.p_FEJ) - t_c_b)
        else:
            p_old = R_c2w_old.T @ (p_w - o.p_cam)
        p_old_ = R_c2w_old.T @ (p_w - o.p_cam)
        invDepth_old = 1 / p_old_[2]
        f_old = np.array([p_old_[0] / p_old_[2], p_old_[1] / p_old_[2], 1.0])
        R_b2w_new = quat_to_rot(n.q); R_w2b_new = R_b2w_new.T
        R_w2c_new = quat_to_rot(n.q_cam).T
        invDepth_new = ft.invDepth
        if self.if_FEJ:
            p_bf_old = ft.position_FEJ - o.p_FEJ; p_bf_new = ft.position_FEJ - n.p_FEJ
        else:
            p_bf_old = p_w - o.p; p_bf_new = p_w - n.p
        J_rho_d_new = -invDepth_new * invDepth_new
        J_d = (R_w2c_new @ R_c2w_old @ f_old)[2]
        J_theta_old = (-R_w2c_new @ skew(p_bf_old))[2]; J_p_old = R_w2c_new[2]
        J_theta_new = (R_w2c_new @ skew(p_bf_new))[2]; J_p_new = (-R_w2c_new)[2]
        Sk = skew(R_w2b_new @ p_bf_new - t_c_b)
        Mx = R_w2b_new @ R_b2w_old @ skew(R_b2c.T @ p_old)
        J_e_theta = (R_b2c @ (Sk - Mx))[2]; J_e_p = (R_b2c @ (R_w2b_new @ R_b2w_old - np.eye(3)))[2]
        J_d_rho_old = -1 / (invDepth_old * invDepth_old)
        J = np.zeros(self.P.shape[1])
        order = sorted(self.aug.keys())
        oc = order.index(old_id); nc = order.index(new_id); fc = self.feature_states.index(ft.id)
        fi = self.LEG + 6 * N + fc
        J[fi] = J_rho_d_new * J_d * J_d_rho_old
        J[self.LEG + 6 * oc:self.LEG + 6 * oc + 3] = J_rho_d_new * J_theta_old
        J[self.LEG + 6 * oc + 3:self.LEG + 6 * oc + 6] = J_rho_d_new * J_p_old
        J[self.LEG + 6 * nc:self.LEG + 6 * nc + 3] = J_rho_d_new * J_theta_new
        J[self.LEG + 6 * nc + 3:self.LEG + 6 * nc + 6] = J_rho_d_new * J_p_new
        J[15:18] = J_rho_d_new * J_e_theta; J[18:21] = J_rho_d_new * J_e_p
        Pfl = J @ self.P
        Pff = float(Pfl @ J)
        P = self.P
        P[fi, :] = Pfl; P[:, fi] = Pfl
        P[fi, fi] = Pff
        self.P = (P + P.T) / 2.0

    # ---- :2259-2307
    def _find_redundant(self):
        ids = sorted(self.aug.keys())
        key_i = len(ids) - 4
        st_i = key_i + 1
        first_i = 0
        key = self.aug[ids[key_i]]
        key_R = quat_to_rot(key.q_cam)
        rm = []
        for _ in range(2):
            a = self.aug[ids[st_i]]
            Rr = quat_to_rot(a.q_cam).T
            dist = np.linalg.norm(a.p_cam - key.p_cam)
            M = Rr @ key_R
            # Eigen AngleAxisd(R).angle(): via quaternion, angle = 2*atan2(|vec|, |w|)
            qq = rot_to_quat(M)
            n = np.linalg.norm(qq[:3])
            angle = 2 * np.arctan2(n, abs(qq[3]))
            if angle < self.rotation_threshold and dist < self.translation_threshold and self.tracking_rate > self.tracking_rate_threshold:
                rm.append(ids[st_i]); st_i += 1
            else:
                rm.append(ids[first_i]); first_i += 1
                st_i -= 2
        return sorted(rm)

    # ---- :2310-2641 (pure MSCKF)
    def _prune(self):
        if not self.if_ZUPT:
            if len(self.aug) < self.sw_size:
                return
            rm_ids = self._find_redundant()
        else:
            rm_ids = [self.imu_state.id - 1]
        rows = 0
        used = []
        new_nui = []
        sid_now = self.imu_state.id
        for fid in sorted(self.map_server.keys()):
            ft = self.map_server[fid]
            involved = [sid for sid in rm_ids if sid in ft.obs]
            if len(involved) == 0:
                continue
            if ft.in_state:
                if ft.id_anchor in involved:                       # :2345-2405
                    if self.use_schmidt and self.imu_state.id - ft.id_anchor > 2:      # :2351-2358: a mature anchor becomes a nuisance state
                        self.nui_features.setdefault(ft.id_anchor, []).append(fid)
                        if ft.id_anchor not in new_nui:
                            new_nui.append(ft.id_anchor)
                        continue
                    if self.idp == 3:                              # the newest state becomes the anchor (:2361-2378)
                        new_id = self.imu_state.id
                        a = self.aug[new_id]
                        p_new = quat_to_rot(a.q_cam).T @ (ft.position - a.p_cam)
                        ft.invParam = np.array([p_new[0] / p_new[2], p_new[1] / p_new[2], 1 / p_new[2]])
                        self._update_feature_cov_3didp(ft, ft.id_anchor, new_id)
                    else:
                        new_id = self._new_anchor_id(ft, involved)
                        a = self.aug[new_id]
                        p_new = quat_to_rot(a.q_cam).T @ (ft.position - a.p_cam)
                        ft.invDepth = 1 / p_new[2]
                        ft.obs_anchor = np.array([p_new[0] / p_new[2], p_new[1] / p_new[2], ft.obs_anchor[2]])
                        self._update_feature_cov_1didp(ft, ft.id_anchor, new_id)
                    ft.id_anchor = new_id
                    self.stats["anchor_changes"] = self.stats.get("anchor_changes", 0) + 1
                continue
            if ft.is_initialized and ft.id_anchor in involved:     # :2407-2460
                if self.idp == 3:
                    new_id = self.imu_state.id
                    a = self.aug[new_id]
                    p_new = quat_to_rot(a.q_cam).T @ (ft.position - a.p_cam)
                    ft.invParam = np.array([p_new[0] / p_new[2], p_new[1] / p_new[2], 1 / p_new[2]])
                else:
                    new_id = self._new_anchor_id(ft, involved)
                    a = self.aug[new_id]
                    R_c2w = quat_to_rot(a.q_cam)
                    p_new = R_c2w.T @ (ft.position - a.p_cam)
                    ft.invDepth = 1 / p_new[2]
                    ft.obs_anchor = np.array([ft.obs[new_id][0], ft.obs[new_id][1], ft.obs_anchor[2]])
                ft.id_anchor = new_id
            if not self.if_ZUPT and not ft.ekf_feature and len(involved) > 1:
                tracked = sid_now in ft.obs
                if not ft.is_initialized:
                    if not ft.check_motion(self.aug, tracked):
                        continue
                    if not ft.initialize_position(self.aug, None):
                        continue
                used.append(fid)
                rows += 2 * len(involved) - 3
        self.stats["prune_used"] = len(used)
        if not self.if_ZUPT and len(used) != 0:
            Hs = []; rs = []
            for fid in sorted(self.map_server.keys()):
                ft = self.map_server[fid]
                involved = [sid for sid in rm_ids if sid in ft.obs]
                if fid in used:
                    Hj, rj = self._feature_jacobian(ft, involved)
                    if self._gating(Hj, rj, 2 * len(involved) - 3):
                        Hs.append(Hj); rs.append(rj)
                for sid in involved:
                    del ft.obs[sid]
            H = np.concatenate(Hs) if Hs else np.zeros((0, self.P.shape[1]))
            r = np.concatenate(rs) if rs else np.zeros(0)
            if H.shape[0] > 0:
                cols = self.LEG + 6 * len(self.aug)
                if H.shape[0] > H.shape[1]:
                    H, r = self._compress(H, r, cols)
                self._update(H, r, "msckf")
        else:
            for fid in sorted(self.map_server.keys()):
                ft = self.map_server[fid]
                for sid in [sid for sid in rm_ids if sid in ft.obs]:
                    del ft.obs[sid]
        for sid in rm_ids:
            order = sorted(self.aug.keys())
            seq = order.index(sid)
            a0 = self.LEG + 6 * seq
            keep = [i for i in range(self.P.shape[0]) if not (a0 <= i < a0 + 6)]
            if self.use_schmidt and sid in new_nui:                # :2569-2613: the pose block moves behind everything else
                perm = keep + list(range(a0, a0 + 6))
                self.P = self.P[np.ix_(perm, perm)]
                self.nui_ids.append(sid); self.nui_states[sid] = self.aug[sid]
            else:
                self.P = self.P[np.ix_(keep, keep)]
            del self.aug[sid]

    # ---- getNewAnchorId :3412-3472
    def _new_anchor_id(self, ft, involved):
        order = sorted(self.aug.keys())
        size = len(order)
        if size <= 2:
            return order[-1]
        best = None; min_dis = 99999.0
        for sid in order[:size - 2]:
            if sid not in ft.obs or sid in involved:
                continue
            a = self.aug[sid]
            p_new = quat_to_rot(a.q_cam).T @ (ft.position - a.p_cam)
            dis = float(np.linalg.norm(np.array([p_new[0] / p_new[2], p_new[1] / p_new[2]]) - ft.obs[sid]))
            if min_dis > dis:
                min_dis = dis; best = sid
        return best if best is not None else order[-1]
