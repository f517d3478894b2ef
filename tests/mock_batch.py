"""An oracle-backed stand-in for larvio_b200.api.Batch, used ONLY to execute the GPU tests' Python harness on a box
without a GPU (tests/test_cpu.py::test_gpu_test_harness_runs_against_the_mock): it answers the Batch calls the harness
makes with a second, independent oracle instance, so harness bugs (indexing, buffer strides, empty messages) surface
on CPU instead of costing a GPU run.  It is test infrastructure and never part of the product path."""
import numpy as np


class MockBatch:
    def __init__(self, cfg, n_seq, device=0):
        from larvio_b200 import api
        from oracle.frontend import ImageProcessorOracle
        from oracle.backend import LarVioOracle
        self.api = api
        self.S = n_seq
        self.cap = ((int(cfg.raw["max_features_num"]) + 31) // 32) * 32
        self.fe = [ImageProcessorOracle(cfg.raw) for _ in range(n_seq)]
        self.be = [LarVioOracle(cfg.raw) for _ in range(n_seq)]

    @staticmethod
    def _rows(buf, n, s):
        k = int(n[s])
        return np.concatenate([buf["t"][s, :k, None], buf["gyro"][s, :k], buf["acc"][s, :k]], 1).reshape(-1, 7)

    def set_initial_state(self, s, t, q, p, v, bg, ba):
        self.be[s].set_initial_state(t, np.array(q, float), np.array(p, float), np.array(v, float), np.array(bg, float), np.array(ba, float))

    def _consume(self, s, msg, buf, n):
        rows = self._rows(buf, n, s).tolist()
        ok = self.be[s].process_features(msg, rows)
        k = len(rows)
        arr = np.array(rows).reshape(-1, 7)
        buf["t"][s, :k] = arr[:, 0]; buf["gyro"][s, :k] = arr[:, 1:4]; buf["acc"][s, :k] = arr[:, 4:7]
        n[s] = k
        return bool(ok)

    def step(self, images, t_img, buf, n, images_on_device=False):
        ok = np.zeros(self.S, np.uint8)
        for s in range(self.S):
            msg = self.fe[s].process_image(images[s], float(t_img[s]), self._rows(buf, n, s))
            if msg is not None:
                ok[s] = self._consume(s, msg, buf, n)
        return ok

    def process_images(self, images, t_img, buf, n):
        feat = np.zeros((self.S, self.cap), self.api.FEATURE_DTYPE); out_n = np.zeros(self.S, np.int32); has = np.zeros(self.S, np.uint8)
        self._msgs = [None] * self.S
        for s in range(self.S):
            msg = self.fe[s].process_image(images[s], float(t_img[s]), self._rows(buf, n, s))
            self._msgs[s] = msg
            if msg is None:
                continue
            k = len(msg.ids); has[s] = 1; out_n[s] = k; feat["id"][s, :k] = msg.ids
            for c, name in enumerate(["u", "v", "u_init", "v_init", "u_vel", "v_vel", "u_init_vel", "v_init_vel"]):
                feat[name][s, :k] = msg.data[:, c]
        return feat, out_n, has

    def process_features(self, valid, t_msg, feat, n_feat, buf, n):
        from oracle.frontend import FeatureMsg
        ok = np.zeros(self.S, np.uint8)
        for s in range(self.S):
            if not valid[s]:
                continue
            k = int(n_feat[s])
            msg = FeatureMsg(float(t_msg[s]))
            msg.ids = feat["id"][s, :k].copy()
            msg.data = np.stack([feat[name][s, :k] for name in ["u", "v", "u_init", "v_init", "u_vel", "v_vel", "u_init_vel", "v_init_vel"]], 1)
            ok[s] = self._consume(s, msg, buf, n)
        return ok

    def get_state(self, s):
        o = self.be[s].imu_state
        return dict(t=o.time, q=o.q.copy(), p=o.p.copy(), v=o.v.copy(), bg=o.bg.copy(), ba=o.ba.copy())

    def get_covariance(self, s):
        return self.be[s].P.copy()

    def get_calibration(self, s):
        b = self.be[s]; o = b.imu_state
        return dict(R_imu_cam0=o.R_imu_cam0.copy(), t_cam0_imu=o.t_cam0_imu.copy(), td=b.td, Tg=b.Tg.copy(), As=b.As.copy(), Ma=b.Ma.copy())

    def get_window(self, s, cap=64):
        a = self.be[s].aug
        return np.array([np.concatenate([a[k].q, a[k].p]) for k in sorted(a)]).reshape(-1, 7)

    def get_points(self, s, which, cap=512):
        m = self.be[s].get_active_map_points() if which else self.be[s].get_stable_map_points()
        return {int(k): np.array(v) for k, v in m.items()}

    def close(self):
        pass
