"""On-disk formats either side of the hot path, host side:

* EuRoC ASL ingest as the reference's replay driver reads it: `cam0/data.csv` (`#timestamp [ns],filename`) and
  `imu0/data.csv` (`#timestamp [ns],w_x,w_y,w_z,a_x,a_y,a_z`), include/utils/DataReader.hpp:31-120, first image/IMU
  alignment :123-165, and the replay loop's "IMU up to 50 ms past the image" rule, app/larvioMain.cpp:87-102;
* the trajectory log `msckf_2_state.txt` the estimator appends one line to per published frame (larvio.cpp:318, 420-453)
  and `msckf_2_takeoff.txt` (:319, 388).

Images are decoded with cv2 (host); everything after the decode runs through the C ABI."""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np


def load_image_list(csv_path: str) -> List[Tuple[float, str]]:
    """loadImageList (DataReader.hpp:31-60): header line skipped, stamp = 1e-9 * integer ns, second column = file name.
    The reference's `while(!eof) getline` loop also emits one bogus record for the empty line after the last newline;
    it is not reproduced."""
    out = []
    with open(csv_path) as f:
        f.readline()
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            parts = line.split(",")
            out.append((1e-9 * int(parts[0]), parts[1].strip() if len(parts) > 1 else ""))
    return out


def load_imu_file(csv_path: str) -> np.ndarray:
    """loadImuFile (DataReader.hpp:68-120) -> rows [t, w_x, w_y, w_z, a_x, a_y, a_z]."""
    rows = []
    with open(csv_path) as f:
        f.readline()
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            p = line.split(",")
            rows.append([1e-9 * int(p[0])] + [float(x) for x in p[1:7]])
    return np.array(rows, np.float64).reshape(-1, 7)


def find_first_align(imu: np.ndarray, imgs: List[Tuple[float, str]]):
    """findFirstAlign (DataReader.hpp:123-165): (image index, IMU index) of the first pair with EQUAL stamps, or None."""
    imu_t0, img_t0 = imu[0, 0], imgs[0][0]
    if imu_t0 > img_t0:
        for i in range(1, len(imgs)):
            if imu_t0 <= imgs[i][0]:
                hit = np.nonzero(imu[:, 0] == imgs[i][0])[0]
                return (i, int(hit[0])) if len(hit) else None
        return None
    if imu_t0 < img_t0:
        hit = np.nonzero(imu[1:, 0] == img_t0)[0]
        return (0, int(hit[0]) + 1) if len(hit) else None
    return (0, 0)


class Replay:
    """The replay driver's data loop (app/larvioMain.cpp:62-102) for one EuRoC sequence directory: iterates
    (t_img, image u8 [H,W], new IMU rows) with the 0.05 s look-ahead rule."""

    def __init__(self, mav_dir: str):
        import cv2
        self._cv2 = cv2
        self.dir = mav_dir
        self.imgs = load_image_list(os.path.join(mav_dir, "cam0", "data.csv"))
        self.imu = load_imu_file(os.path.join(mav_dir, "imu0", "data.csv"))
        al = find_first_align(self.imu, self.imgs)
        if al is None:
            raise ValueError("no image/IMU pair with equal stamps (findFirstAlign failed)")
        self.imgs = self.imgs[al[0]:]
        self.imu = self.imu[al[1]:]

    def __iter__(self):
        k = 0
        for t, name in self.imgs:
            img = self._cv2.imread(os.path.join(self.dir, "cam0", "data", name), 0)
            k2 = k
            while k2 < len(self.imu) and self.imu[k2, 0] - t < 0.05:
                k2 += 1
            yield t, img, self.imu[k:k2]
            k = k2


def _fmt(x: float) -> str:
    """C++ ostream default formatting of a double (precision 6, %g)."""
    return "%g" % x


def state_line(t_rel, q_xyzw, v, p, bg, ba, R_imu_cam0, t_cam0_imu) -> str:
    """One line of msckf_2_state.txt (larvio.cpp:420-453): time since take-off, q (w x y z), v, p, bg, ba,
    q_bc = Quaterniond(R_imu_cam0) (w x y z), t_cam0_imu — default stream precision."""
    from .synth import rot_to_quat_xyzw
    qbc = rot_to_quat_xyzw(np.asarray(R_imu_cam0, np.float64).reshape(3, 3))
    vals = [t_rel, q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2], *v, *p, *bg, *ba, qbc[3], qbc[0], qbc[1], qbc[2], *t_cam0_imu]
    return " ".join(_fmt(float(x)) for x in vals)


class TrajectoryLog:
    """msckf_2_state.txt + msckf_2_takeoff.txt of one sequence, fed from Batch.get_state / get_calibration."""

    def __init__(self, output_dir: str):
        os.makedirs(output_dir, exist_ok=True)
        self.f_state = open(os.path.join(output_dir, "msckf_2_state.txt"), "w")
        self.f_takeoff = open(os.path.join(output_dir, "msckf_2_takeoff.txt"), "w")
        self.take_off = None

    def set_take_off(self, t: float):
        self.take_off = float(t)
        self.f_takeoff.write("%.9f\n" % t)          # fixed << setprecision(9) (larvio.cpp:388)
        self.f_takeoff.flush()

    def append(self, state: dict, calib: dict):
        if self.take_off is None:
            raise ValueError("take-off stamp not set")
        self.f_state.write(state_line(state["t"] - self.take_off, state["q"], state["v"], state["p"], state["bg"], state["ba"],
                                      calib["R_imu_cam0"], calib["t_cam0_imu"]) + "\n")

    def close(self):
        self.f_state.close(); self.f_takeoff.close()


def read_state_log(path: str) -> np.ndarray:
    """msckf_2_state.txt -> [n, 24]."""
    return np.loadtxt(path, ndmin=2)
