"""Parser for the reference's OpenCV-FileStorage YAML dialect (``%YAML:1.0``).

The reference reads ONE yaml file twice with ``cv::FileStorage``: front-end keys at
image_processor.cpp:51-92, back-end keys at larvio.cpp:65-277.
This module parses the same file without OpenCV (the C++ host side has its own parser,
``csrc/lvb_config.cpp``; both must accept config/euroc.yaml unchanged) and turns it
into the flat ``LvbConfig`` POD that crosses the C ABI (include/larvio_b200.h).
"""
from __future__ import annotations

import ctypes
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List


def _scalar(tok: str) -> Any:
    tok = tok.strip()
    if len(tok) >= 2 and tok[0] == '"' and tok[-1] == '"':
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


def _strip_comment(raw: str) -> str:
    """Drop a trailing ``# comment`` unless the ``#`` sits inside a quoted string."""
    inq = False
    for j, ch in enumerate(raw):
        if ch == '"':
            inq = not inq
        elif ch == '#' and not inq:
            return raw[:j].rstrip()
    return raw.rstrip()


def parse_opencv_yaml(text: str) -> Dict[str, Any]:
    """Tiny subset parser: scalars, one level of nested maps, ``!!opencv-matrix``."""
    out: Dict[str, Any] = {}
    lines = text.splitlines()
    i = 0
    cur_map = None
    while i < len(lines):
        raw = lines[i]
        i += 1
        line = _strip_comment(raw)
        if not line.strip() or line.startswith('%YAML') or line.strip() == '---':
            continue
        indent = len(line) - len(line.lstrip())
        m = re.match(r'\s*([A-Za-z_][A-Za-z0-9_]*)\s*:\s*(.*)$', line)
        if not m:
            continue
        key, val = m.group(1), m.group(2).strip()
        if indent == 0:
            cur_map = None
            if val.startswith('!!opencv-matrix'):
                mat: Dict[str, Any] = {}
                # gather until the data list closes
                buf = ''
                while i < len(lines):
                    l2 = lines[i].split('#', 1)[0]
                    i += 1
                    mm = re.match(r'\s*(rows|cols|dt)\s*:\s*(\S+)', l2)
                    if mm:
                        mat[mm.group(1)] = _scalar(mm.group(2))
                        continue
                    buf += ' ' + l2.strip()
                    if ']' in l2:
                        break
                nums = re.findall(r'[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?', buf[buf.find('['):])
                mat['data'] = [float(x) for x in nums]
                assert len(mat['data']) == mat['rows'] * mat['cols'], key
                out[key] = mat
            elif val == '':
                cur_map = {}
                out[key] = cur_map
            else:
                out[key] = _scalar(val)
        else:
            if cur_map is not None:
                cur_map[key] = _scalar(val)
    return out


class LvbConfig(ctypes.Structure):
    """Mirror of ``struct LvbConfig`` in include/larvio_b200.h (keep in sync)."""
    _fields_ = [
        # camera
        ("width", ctypes.c_int), ("height", ctypes.c_int),
        ("distortion_model", ctypes.c_int),  # 0 radtan, 1 equidistant
        ("_pad0", ctypes.c_int),
        ("fx", ctypes.c_double), ("fy", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double),
        ("dist", ctypes.c_double * 4),
        ("T_cam_imu", ctypes.c_double * 16),
        # front end
        ("pyramid_levels", ctypes.c_int), ("patch_size", ctypes.c_int),
        ("max_iteration", ctypes.c_int), ("max_features_num", ctypes.c_int),
        ("min_distance", ctypes.c_int), ("flag_equalize", ctypes.c_int),
        ("track_precision", ctypes.c_double),
        ("pub_frequency", ctypes.c_double), ("img_rate", ctypes.c_double),
        # back end
        ("imu_rate", ctypes.c_double),
        ("rotation_threshold", ctypes.c_double), ("translation_threshold", ctypes.c_double),
        ("tracking_rate_threshold", ctypes.c_double),
        ("feature_translation_threshold", ctypes.c_double),
        ("td", ctypes.c_double),
        ("noise_gyro", ctypes.c_double), ("noise_acc", ctypes.c_double),
        ("noise_gyro_bias", ctypes.c_double), ("noise_acc_bias", ctypes.c_double),
        ("noise_feature", ctypes.c_double),
        ("cov_orientation", ctypes.c_double), ("cov_velocity", ctypes.c_double),
        ("cov_position", ctypes.c_double), ("cov_gyro_bias", ctypes.c_double),
        ("cov_acc_bias", ctypes.c_double), ("cov_extrin_rot", ctypes.c_double),
        ("cov_extrin_trans", ctypes.c_double),
        ("zupt_max_feature_dis", ctypes.c_double),
        ("zupt_noise_v", ctypes.c_double), ("zupt_noise_p", ctypes.c_double), ("zupt_noise_q", ctypes.c_double),
        ("static_duration", ctypes.c_double),
        ("max_track_len", ctypes.c_int), ("sw_size", ctypes.c_int),
        ("least_observation_number", ctypes.c_int),
        ("if_FEJ", ctypes.c_int), ("estimate_extrin", ctypes.c_int), ("estimate_td", ctypes.c_int),
        ("calib_imu_instrinsic", ctypes.c_int), ("if_ZUPT_valid", ctypes.c_int),
        ("max_features_in_one_grid", ctypes.c_int), ("aug_grid_rows", ctypes.c_int),
        ("aug_grid_cols", ctypes.c_int), ("feature_idp_dim", ctypes.c_int),
        ("use_schmidt", ctypes.c_int), ("_pad1", ctypes.c_int),
    ]


_DIRECT = [
    "pyramid_levels", "patch_size", "max_iteration", "max_features_num", "min_distance",
    "flag_equalize", "track_precision", "pub_frequency", "img_rate", "imu_rate",
    "rotation_threshold", "translation_threshold", "tracking_rate_threshold",
    "feature_translation_threshold", "td", "noise_gyro", "noise_acc", "noise_gyro_bias",
    "noise_acc_bias", "noise_feature", "zupt_max_feature_dis", "zupt_noise_v", "zupt_noise_p",
    "zupt_noise_q", "static_duration", "max_track_len", "sw_size", "least_observation_number",
    "if_FEJ", "estimate_extrin", "estimate_td", "calib_imu_instrinsic", "if_ZUPT_valid",
    "max_features_in_one_grid", "aug_grid_rows", "aug_grid_cols", "feature_idp_dim", "use_schmidt",
]
_RENAMED = {
    "cov_orientation": "initial_covariance_orientation",
    "cov_velocity": "initial_covariance_velocity",
    "cov_position": "initial_covariance_position",
    "cov_gyro_bias": "initial_covariance_gyro_bias",
    "cov_acc_bias": "initial_covariance_acc_bias",
    "cov_extrin_rot": "initial_covariance_extrin_rot",
    "cov_extrin_trans": "initial_covariance_extrin_trans",
}


@dataclass
class Config:
    """Python view of one parsed config file; ``raw`` keeps every key as read."""
    raw: Dict[str, Any] = field(default_factory=dict)

    @staticmethod
    def load(path: str, **overrides) -> "Config":
        with open(path, "r") as f:
            raw = parse_opencv_yaml(f.read())
        raw.update(overrides)
        return Config(raw)

    def __getitem__(self, k):
        return self.raw[k]

    def get(self, k, d=None):
        return self.raw.get(k, d)

    def with_overrides(self, **kw) -> "Config":
        r = dict(self.raw)
        r.update(kw)
        return Config(r)

    def to_struct(self) -> LvbConfig:
        c = LvbConfig()
        r = self.raw
        c.width = int(r["resolution_width"])
        c.height = int(r["resolution_height"])
        c.distortion_model = 1 if r.get("distortion_model", "radtan") == "equidistant" else 0
        intr, dist = r["intrinsics"], r["distortion_coeffs"]
        c.fx, c.fy, c.cx, c.cy = (float(intr[k]) for k in ("fx", "fy", "cx", "cy"))
        for i, k in enumerate(("k1", "k2", "p1", "p2")):
            c.dist[i] = float(dist[k])
        for i, v in enumerate(r["T_cam_imu"]["data"]):
            c.T_cam_imu[i] = v
        for k in _DIRECT:
            setattr(c, k, type(getattr(c, k))(r[k]))
        for k, src in _RENAMED.items():
            setattr(c, k, float(r[src]))
        return c
