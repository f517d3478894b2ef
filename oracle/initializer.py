"""TEST INFRASTRUCTURE (see oracle/__init__.py): CPU restatement of the reference's inclinometer initialiser,
src/StaticInitializer.cpp:13-161, as FlexibleInitializer::tryIncInit (src/FlexibleInitializer.cpp:10-26) drives it.
Pinned by tests/golden/ref_self_start.npz: the reference's own StaticInitializer.cpp (compiled unmodified, oracle/_ref) starts the
filter at the same call with the same state; also checked against the synthetic truth."""
import numpy as np


class StaticInitializerOracle:
    def __init__(self, cfg_raw: dict):
        self.max_feature_dis = float(cfg_raw["zupt_max_feature_dis"])                              # larvio.cpp:343-344
        self.static_num = int(float(cfg_raw["static_duration"]) * float(cfg_raw["pub_frequency"]))  # :223-224
        self.td = float(cfg_raw["td"])
        self.counter = 0
        self.init_features = {}
        self.lower_time_bound = 0.0

    # ---- tryIncInit :13-75; msg_ids / msg_uv are the message's feature ids and (u, v); imu rows [t, w(3), a(3)]
    def try_inc_init(self, msg_ids, msg_uv, t_msg, imu):
        if self.counter == 0:
            self.counter += 1
            self.init_features = {int(i): np.array(p, np.float64) for i, p in zip(msg_ids, msg_uv)}
            self.lower_time_bound = t_msg + self.td
            return None
        curr = {}
        dis = []
        for i, p in zip(msg_ids, msg_uv):
            p = np.array(p, np.float64)
            curr[int(i)] = p
            if int(i) in self.init_features:
                d = p - self.init_features[int(i)]
                dis.append(float(np.sqrt(d[0] * d[0] + d[1] * d[1])))
        if len(dis) < 20:
            self.counter = 0
            return None
        dis.sort()
        max_dis = dis[len(dis) - 19]
        if max_dis < self.max_feature_dis:
            self.counter += 1
            self.init_features = curr
            if self.counter < self.static_num:
                return None
        else:
            self.counter = 0
            return None
        return self._initialize(t_msg + self.td, np.asarray(imu, np.float64).reshape(-1, 7))

    # ---- initializeGravityAndBias :78-124 + assignInitialState :127-158 (Ma = Tg = I, As = 0)
    def _initialize(self, time_bound, imu):
        sw = np.zeros(3); sa = np.zeros(3); used = 0; last_t = 0.0
        for row in imu:
            if row[0] < self.lower_time_bound:
                continue
            if row[0] > time_bound:
                break
            sw = sw + row[1:4]; sa = sa + row[4:7]; used += 1; last_t = row[0]
        bg = sw / used
        g_imu = sa / used
        gn = np.linalg.norm(g_imu)
        v0 = g_imu / np.sqrt(g_imu @ g_imu); v1 = np.array([-0.0, -0.0, gn]); v1 = v1 / np.sqrt(v1 @ v1)
        c = float(v1 @ v0)
        axis = np.cross(v0, v1)
        s = np.sqrt((1.0 + c) * 2.0)
        q = np.concatenate([axis * (1.0 / s), [s * 0.5]])                 # Eigen coeffs order x y z w
        useful = 0
        for row in imu:
            if row[0] > last_t:
                break
            useful += 1
        if useful >= len(imu):
            useful -= 1
        return dict(t=last_t, q=q, p=np.zeros(3), v=np.zeros(3), bg=bg, ba=np.zeros(3), gyro_old=imu[useful, 1:4].copy(),
                    acc_old=imu[useful, 4:7].copy(), n_consumed=useful, used=used)
