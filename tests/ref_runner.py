"""TEST INFRASTRUCTURE: record the stream app/larvioMain.cpp:87-117 feeds LarVio::processFeatures (feature messages of the
front-end oracle + the IMU samples between them) and replay it through oracle/_ref/larvio_ref - the reference's own
src/larvio.cpp + StaticInitializer.cpp compiled unmodified against oracle/ref_shim/ (Makefile target `ref`; only possible
where /root/reference exists, i.e. in the build container).  tests/golden/make_ref_golden.py turns the replies into the
committed fixtures tests/golden/ref_*.npz, which travel to the GPU box.  Never imported by the package."""
import os
import subprocess
import tempfile

import numpy as np

from larvio_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "larvio_ref")

_SCALARS = """sw_size max_track_len least_observation_number if_FEJ estimate_extrin estimate_td calib_imu_instrinsic feature_idp_dim
use_schmidt max_features_in_one_grid aug_grid_rows aug_grid_cols rotation_threshold translation_threshold tracking_rate_threshold
feature_translation_threshold position_std_threshold reset_fej_threshold noise_gyro noise_acc noise_gyro_bias noise_acc_bias
noise_feature initial_covariance_orientation initial_covariance_velocity initial_covariance_position initial_covariance_gyro_bias
initial_covariance_acc_bias initial_covariance_extrin_rot initial_covariance_extrin_trans if_ZUPT_valid zupt_max_feature_dis
zupt_noise_v zupt_noise_p zupt_noise_q static_duration imu_rate pub_frequency resolution_width resolution_height td
fast_threshold patch_size pyramid_levels max_iteration track_precision ransac_threshold max_features_num min_distance flag_equalize
img_rate""".split()


def write_reference_yaml(cfg_raw: dict, path: str, output_dir: str):
    """cfg_raw -> a settings file in the reference's own format (config/euroc.yaml), every key larvio.cpp:58-277 reads."""
    lines = ["%YAML:1.0", "", 'output_dir: "%s"' % output_dir]
    for k in _SCALARS:
        lines.append("%s: %s" % (k, repr(cfg_raw[k]) if isinstance(cfg_raw[k], float) else cfg_raw[k]))
    lines.append('distortion_model: "%s"' % cfg_raw["distortion_model"])
    it = cfg_raw["intrinsics"]
    lines.append("intrinsics:")
    for k in ("fx", "fy", "cx", "cy"):
        lines.append("   %s: %r" % (k, float(it[k])))
    lines.append("distortion_coeffs:")
    for k in ("k1", "k2", "p1", "p2"):
        lines.append("   %s: %r" % (k, float(cfg_raw["distortion_coeffs"][k])))
    T = np.asarray(cfg_raw["T_cam_imu"]["data"], np.float64).reshape(4, 4)
    lines += ["T_cam_imu: !!opencv-matrix", "   rows: 4", "   cols: 4", "   dt: d", "   data:"]
    rows = [", ".join(repr(float(v)) for v in T[i]) for i in range(4)]
    lines.append("    [" + ",\n     ".join(rows) + "]")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


REF_FE_BIN = os.path.join(ROOT, "oracle", "_ref", "larvio_ref_fe")


def run_reference_frontend(cfg_raw, seq, n_frames):
    """The reference's own ImageProcessor (oracle/_ref/larvio_ref_fe, `make ref_fe`; its OpenCV calls run in cv2 through
    oracle/cv_server.py) over the images and IMU rows of one sequence -> per frame None or dict(t, ids, data[n][8])."""
    if not os.path.exists(REF_FE_BIN):
        raise FileNotFoundError(REF_FE_BIN + " (build it with `make ref_fe`; needs /root/reference)")
    import sys
    with tempfile.TemporaryDirectory() as td:
        ypath = os.path.join(td, "cfg.yaml"); ipath = os.path.join(td, "in.bin"); opath = os.path.join(td, "out.bin")
        write_reference_yaml(cfg_raw, ypath, td + "/")
        imgs = np.ascontiguousarray(seq.images[:n_frames], np.uint8)
        with open(ipath, "wb") as f:
            np.array([n_frames, imgs.shape[1], imgs.shape[2], len(seq.imu)], np.float64).tofile(f)
            np.asarray(seq.img_t[:n_frames], np.float64).tofile(f)
            np.ascontiguousarray(seq.imu, np.float64).tofile(f)
            imgs.tofile(f)
        env = dict(os.environ, LVB_CV_SERVER=os.path.join(ROOT, "oracle", "cv_server.py"), LVB_CV_SERVER_PYTHON=sys.executable)
        r = subprocess.run([REF_FE_BIN, ypath, ipath, opath], capture_output=True, text=True, timeout=3600, env=env)
        if r.returncode != 0:
            raise RuntimeError("larvio_ref_fe failed (%d): %s" % (r.returncode, (r.stderr or r.stdout)[-3000:]))
        o = np.fromfile(opath, np.float64)
    k = 0
    out = []
    for _ in range(n_frames):
        has = bool(o[k]); k += 1
        if not has:
            out.append(None); continue
        t = float(o[k]); n = int(o[k + 1]); k += 2
        rows = o[k:k + 9 * n].reshape(n, 9); k += 9 * n
        out.append(dict(t=t, ids=rows[:, 0].astype(np.uint64), data=rows[:, 1:].copy()))
    assert k == len(o)
    return out


def record_calls(cfg_raw, seq, n_frames):
    """Front-end oracle over the images of one sequence -> the list of processFeatures calls [(t_msg, new imu rows, ids, data)]
    plus the frame index of each call."""
    from oracle.frontend import ImageProcessorOracle
    fe = ImageProcessorOracle(cfg_raw)
    imu = []
    pending = []
    k = 0
    calls = []
    for j in range(n_frames):
        k2 = synth.imu_window(seq, k, seq.img_t[j])
        new = seq.imu[k:k2]
        imu.extend(new.tolist()); pending.extend(new.tolist()); k = k2
        msg = fe.process_image(seq.images[j], seq.img_t[j], np.array(imu).reshape(-1, 7))
        if msg is not None:
            calls.append(dict(frame=j, t=float(msg.t), imu=np.array(pending, np.float64).reshape(-1, 7),
                              ids=np.asarray(msg.ids, np.uint64).copy(), data=np.asarray(msg.data, np.float64).copy()))
            pending = []
        # the front end only trims ITS view of the buffer; the filter consumes the caller's copy (larvio.cpp:510-512)
    return calls


def run_oracle_on_calls(cfg_raw, calls, init=None, static_init=False, compiled=False):
    """oracle/backend.py (compiled=True: oracle/backend_c.cpp) over recorded calls.  init = (t, q, p, v, bg, ba) forces the initial
    state before the first call; static_init=True uses oracle/initializer.py the way tests/oracle_runner.py does for self-starting runs."""
    from oracle.frontend import FeatureMsg
    if compiled:
        from oracle.backend_c import LarVioOracleC
        be = LarVioOracleC(cfg_raw); be.bFirstFeatures = False
    else:
        from oracle.backend import LarVioOracle
        be = LarVioOracle(cfg_raw)
    si = None
    if static_init:
        from oracle.initializer import StaticInitializerOracle
        si = StaticInitializerOracle(cfg_raw)
    imu = []
    out = []
    for c in calls:
        imu.extend(c["imu"].tolist())
        msg = FeatureMsg(c["t"]); msg.ids = c["ids"]; msg.data = c["data"]
        if init is not None and not be.is_gravity_set:
            be.set_initial_state(*init)
        if si is not None and not be.is_gravity_set:
            # larvio.cpp:365-389: the bFirstFeatures gate precedes the initialiser
            if not be.bFirstFeatures:
                if len(imu) > 0 and imu[0][0] - msg.t - be.td <= 0.0:
                    be.bFirstFeatures = True
                else:
                    out.append(dict(ok=False)); continue
            r = si.try_inc_init(msg.ids, msg.data[:, 0:2], msg.t, imu)
            if r is None:
                out.append(dict(ok=False)); continue
            be.set_initial_state(r["t"], r["q"], r["p"], r["v"], r["bg"], r["ba"])
            be.m_gyro_old = r["gyro_old"]; be.m_acc_old = r["acc_old"]
            del imu[:r["n_consumed"]]
        ok = be.process_features(msg, imu)
        rec = dict(ok=bool(ok))
        if ok:
            s = be.imu_state
            rec.update(t=s.time, q=s.q.copy(), p=s.p.copy(), v=s.v.copy(), bg=s.bg.copy(), ba=s.ba.copy(),
                       R_imu_cam0=s.R_imu_cam0.copy(), t_cam0_imu=s.t_cam0_imu.copy(), td=float(be.td), P=be.P.copy(),
                       slam_ids=[int(i) for i in be.feature_states], n_imu_left=len(imu))
            if compiled:
                rec.update(n_win=be.n_window)
            else:
                rec.update(n_win=len(be.aug), win_ids=sorted(int(i) for i in be.aug), nui_ids=[int(i) for i in getattr(be, "nui_ids", [])],
                           stable={int(k_): np.array(v_) for k_, v_ in be.get_stable_map_points().items()},   # both getters clear what
                           active={int(k_): np.array(v_) for k_, v_ in be.get_active_map_points().items()})   # they return (larvio.cpp:2719-2733)
        out.append(rec)
    return out


def run_reference_on_calls(cfg_raw, calls, init=None, static_init=False):
    """The compiled reference over recorded calls; same record layout as run_oracle_on_calls (+ window, map points)."""
    if not os.path.exists(REF_BIN):
        raise FileNotFoundError(REF_BIN + " (build it with `make ref`; needs /root/reference)")
    with tempfile.TemporaryDirectory() as td:
        ypath = os.path.join(td, "cfg.yaml"); ipath = os.path.join(td, "in.bin"); opath = os.path.join(td, "out.bin")
        write_reference_yaml(cfg_raw, ypath, td + "/")
        buf = [0.0 if not static_init else 1.0]
        if init is not None:
            t, q, p, v, bg, ba = init
            buf += [float(t)] + list(map(float, q)) + list(map(float, p)) + list(map(float, v)) + list(map(float, bg)) + list(map(float, ba))
        else:
            buf += [0.0] * 17
        buf.append(float(len(calls)))
        for c in calls:
            buf += [c["t"], float(len(c["imu"]))] + c["imu"].reshape(-1).tolist() + [float(len(c["ids"]))]
            for i, d in zip(c["ids"], c["data"]):
                buf += [float(i)] + d.tolist()
        np.asarray(buf, np.float64).tofile(ipath)
        r = subprocess.run([REF_BIN, ypath, ipath, opath], capture_output=True, text=True, timeout=1800)
        if r.returncode != 0:
            raise RuntimeError("larvio_ref failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        o = np.fromfile(opath, np.float64)
    k = 0
    out = []

    def take(n):
        nonlocal k
        v = o[k:k + n].copy(); k += n
        return v
    for _ in calls:
        ok = bool(take(1)[0])
        rec = dict(ok=ok)
        if ok:
            rec["t"] = float(take(1)[0]); rec["q"] = take(4); rec["p"] = take(3); rec["v"] = take(3); rec["bg"] = take(3); rec["ba"] = take(3)
            rec["R_imu_cam0"] = take(9).reshape(3, 3); rec["t_cam0_imu"] = take(3); rec["td"] = float(take(1)[0])
            rec["Tg"] = take(9).reshape(3, 3); rec["As"] = take(9).reshape(3, 3); rec["Ma"] = take(9).reshape(3, 3)
            n_win, n_slam, n_nui, dim = (int(x) for x in take(4))
            rec["P"] = take(dim * dim).reshape(dim, dim)
            w = take(n_win * 8).reshape(n_win, 8)
            rec["n_win"] = n_win; rec["win_ids"] = [int(x) for x in w[:, 0]]; rec["win_q"] = w[:, 1:5]; rec["win_p"] = w[:, 5:8]
            rec["slam_ids"] = [int(x) for x in take(n_slam)]; rec["slam_pos"] = take(3 * n_slam).reshape(n_slam, 3)
            rec["nui_ids"] = [int(x) for x in take(n_nui)]
            ns = int(take(1)[0]); st = take(4 * ns).reshape(ns, 4); rec["stable"] = {int(r_[0]): r_[1:4] for r_ in st}
            na = int(take(1)[0]); ac = take(4 * na).reshape(na, 4); rec["active"] = {int(r_[0]): r_[1:4] for r_ in ac}
            rec["n_imu_left"] = int(take(1)[0])
        out.append(rec)
    assert k == len(o), "trailing output of larvio_ref not consumed"
    return out


def compare_runs(a, b):
    """Largest deviations between two runs (lists of per-call records)."""
    worst = dict(q=0.0, p=0.0, v=0.0, bg=0.0, ba=0.0, ext=0.0, td=0.0, P=0.0, n=0)
    for x, y in zip(a, b):
        assert x["ok"] == y["ok"], "processFeatures returned differently"
        if not x["ok"]:
            continue
        assert x["P"].shape == y["P"].shape, "state dimension differs: %s vs %s" % (x["P"].shape, y["P"].shape)
        assert x["n_win"] == y["n_win"] and list(x["slam_ids"]) == list(y["slam_ids"]), "window / SLAM feature bookkeeping differs"
        assert x["n_imu_left"] == y["n_imu_left"], "IMU buffer consumed differently"
        assert list(x.get("nui_ids", [])) == list(y.get("nui_ids", [])), "nuisance states differ"
        dq = min(np.abs(x["q"] - y["q"]).max(), np.abs(x["q"] + y["q"]).max())
        worst["q"] = max(worst["q"], float(dq))
        for k in ("p", "v", "bg", "ba"):
            worst[k] = max(worst[k], float(np.abs(x[k] - y[k]).max()))
        worst["ext"] = max(worst["ext"], float(np.abs(x["R_imu_cam0"] - y["R_imu_cam0"]).max()), float(np.abs(x["t_cam0_imu"] - y["t_cam0_imu"]).max()))
        worst["td"] = max(worst["td"], abs(x["td"] - y["td"]))
        worst["P"] = max(worst["P"], float(np.linalg.norm(x["P"] - y["P"]) / np.linalg.norm(y["P"])))
        worst["n"] += 1
    return worst


# ---- committed fixtures (tests/golden/ref_<case>.npz): the recorded calls + what the compiled reference answered ----------
def fingerprint_vector(d):
    """Fixed probe vector: the fixtures keep P z and diag(P) of every call instead of the d x d matrix."""
    i = np.arange(d, dtype=np.float64)
    return np.cos(1.7 * i + 0.3) + 0.25 * np.sin(0.37 * i * i)


def pack_fixture(overrides, init, static_init, calls, ref):
    import json
    out = dict(cfg_json=np.array(json.dumps(overrides)), static_init=np.array(int(static_init)),
               init=np.zeros(0) if init is None else np.concatenate([[init[0]], init[1], init[2], init[3], init[4], init[5]]).astype(np.float64))
    out["call_t"] = np.array([c["t"] for c in calls]); out["call_frame"] = np.array([c["frame"] for c in calls], np.int32)
    out["imu_ofs"] = np.cumsum([0] + [len(c["imu"]) for c in calls]).astype(np.int64)
    out["imu"] = np.concatenate([c["imu"].reshape(-1, 7) for c in calls])
    out["feat_ofs"] = np.cumsum([0] + [len(c["ids"]) for c in calls]).astype(np.int64)
    out["ids"] = np.concatenate([c["ids"] for c in calls]).astype(np.uint64)
    out["data"] = np.concatenate([c["data"].reshape(-1, 8) for c in calls])
    n = len(calls)
    ok = np.array([r["ok"] for r in ref], np.uint8)
    state = np.zeros((n, 16)); ext = np.zeros((n, 12)); td = np.zeros(n); calib = np.zeros((n, 27))
    meta = np.zeros((n, 5), np.int64)       # dim, n_win, n_slam, n_nui, n_imu_left
    Pz = []; Pd = []; slam = []; Pfro = np.zeros(n)
    pts = []; pts_n = np.zeros((n, 2), np.int64)     # map points the two getters returned after the call: rows [id x y z], stable then active
    last = None
    for i, r in enumerate(ref):
        if not r["ok"]:
            continue
        state[i] = np.concatenate([r["q"], r["p"], r["v"], r["bg"], r["ba"]])
        ext[i] = np.concatenate([r["R_imu_cam0"].reshape(-1), r["t_cam0_imu"]]); td[i] = r["td"]
        calib[i] = np.concatenate([r["Tg"].reshape(-1), r["As"].reshape(-1), r["Ma"].reshape(-1)])
        d = r["P"].shape[0]
        meta[i] = (d, r["n_win"], len(r["slam_ids"]), len(r["nui_ids"]), r["n_imu_left"])
        Pz.append(r["P"] @ fingerprint_vector(d)); Pd.append(np.diag(r["P"]).copy()); slam.append(np.array(r["slam_ids"], np.int64))
        Pfro[i] = np.linalg.norm(r["P"])
        for w_, key in enumerate(("stable", "active")):
            pts_n[i, w_] = len(r[key])
            pts += [np.concatenate([[float(k_)], r[key][k_]]) for k_ in sorted(r[key])]
        last = i
    out.update(ok=ok, state=state, ext=ext, td=td, calib=calib, meta=meta, Pfro=Pfro,
               Pz=np.concatenate(Pz) if Pz else np.zeros(0), Pdiag=np.concatenate(Pd) if Pd else np.zeros(0),
               slam_ids=np.concatenate(slam) if slam else np.zeros(0, np.int64),
               pts=np.array(pts).reshape(-1, 4), pts_n=pts_n)
    if last is not None:
        out["P_last"] = ref[last]["P"]; out["P_last_call"] = np.array(last)
    return out


def load_fixture(path):
    """-> (overrides, init tuple or None, static_init, calls, per-call reference records)."""
    import json
    z = np.load(path)
    overrides = json.loads(str(z["cfg_json"]))
    iv = z["init"]
    init = None if len(iv) == 0 else (float(iv[0]), iv[1:5], iv[5:8], iv[8:11], iv[11:14], iv[14:17])
    calls = []
    for i in range(len(z["call_t"])):
        a, b = z["imu_ofs"][i], z["imu_ofs"][i + 1]; f0, f1 = z["feat_ofs"][i], z["feat_ofs"][i + 1]
        calls.append(dict(frame=int(z["call_frame"][i]), t=float(z["call_t"][i]), imu=z["imu"][a:b], ids=z["ids"][f0:f1], data=z["data"][f0:f1]))
    ref = []
    kd = 0; ks = 0; kp = 0
    for i in range(len(calls)):
        r = dict(ok=bool(z["ok"][i]))
        if r["ok"]:
            st = z["state"][i]; d, n_win, n_slam, n_nui, n_left = (int(x) for x in z["meta"][i])
            r.update(q=st[0:4], p=st[4:7], v=st[7:10], bg=st[10:13], ba=st[13:16], R_imu_cam0=z["ext"][i][:9].reshape(3, 3),
                     t_cam0_imu=z["ext"][i][9:12], td=float(z["td"][i]), Tg=z["calib"][i][0:9].reshape(3, 3), As=z["calib"][i][9:18].reshape(3, 3),
                     Ma=z["calib"][i][18:27].reshape(3, 3), dim=d, n_win=n_win, n_nui=n_nui, n_imu_left=n_left,
                     Pz=z["Pz"][kd:kd + d], Pdiag=z["Pdiag"][kd:kd + d], Pfro=float(z["Pfro"][i]), slam_ids=[int(x) for x in z["slam_ids"][ks:ks + n_slam]])
            kd += d; ks += n_slam
            for w_, key in enumerate(("stable", "active")):
                m_ = int(z["pts_n"][i, w_]); rows = z["pts"][kp:kp + m_]; kp += m_
                r[key] = {int(q_[0]): q_[1:4] for q_ in rows}
            if "P_last_call" in z.files and int(z["P_last_call"]) == i:
                r["P"] = z["P_last"]
        ref.append(r)
    return overrides, init, bool(int(z["static_init"])), calls, ref


def compare_with_fixture(run, ref):
    """run: records with q p v bg ba R_imu_cam0 t_cam0_imu td P n_win slam_ids n_imu_left (oracle, compiled oracle or GPU);
    ref: load_fixture records.  Bookkeeping must be identical; returns the largest numeric deviations."""
    worst = dict(q=0.0, p=0.0, v=0.0, bg=0.0, ba=0.0, ext=0.0, td=0.0, Pz=0.0, Pdiag=0.0, P=0.0, n=0, pts=0.0, n_pts=0)
    for i, (x, y) in enumerate(zip(run, ref)):
        assert bool(x["ok"]) == y["ok"], "call %d: processFeatures returned %s, the reference %s" % (i, x["ok"], y["ok"])
        if not y["ok"]:
            continue
        assert x["P"].shape[0] == y["dim"], "call %d: state dimension %d, the reference has %d" % (i, x["P"].shape[0], y["dim"])
        assert x["n_win"] == y["n_win"], "call %d: window size differs" % i
        if "slam_ids" in x:
            assert list(x["slam_ids"]) == list(y["slam_ids"]), "call %d: SLAM features in the state differ" % i
        if "n_imu_left" in x:
            assert x["n_imu_left"] == y["n_imu_left"], "call %d: IMU buffer consumed differently" % i
        if "nui_ids" in x:
            assert len(x["nui_ids"]) == y["n_nui"], "call %d: %d nuisance states, the reference has %d" % (i, len(x["nui_ids"]), y["n_nui"])
        worst["q"] = max(worst["q"], float(min(np.abs(x["q"] - y["q"]).max(), np.abs(x["q"] + y["q"]).max())))
        for k in ("p", "v", "bg", "ba"):
            worst[k] = max(worst[k], float(np.abs(x[k] - y[k]).max()))
        if "R_imu_cam0" in x:
            worst["ext"] = max(worst["ext"], float(np.abs(x["R_imu_cam0"] - y["R_imu_cam0"]).max()), float(np.abs(x["t_cam0_imu"] - y["t_cam0_imu"]).max()))
            worst["td"] = max(worst["td"], abs(x["td"] - y["td"]))
        d = y["dim"]
        worst["Pz"] = max(worst["Pz"], float(np.linalg.norm(x["P"] @ fingerprint_vector(d) - y["Pz"]) / np.linalg.norm(y["Pz"])))
        worst["Pdiag"] = max(worst["Pdiag"], float(np.linalg.norm(np.diag(x["P"]) - y["Pdiag"]) / np.linalg.norm(y["Pdiag"])))
        if "P" in y:
            worst["P"] = max(worst["P"], float(np.linalg.norm(x["P"] - y["P"]) / np.linalg.norm(y["P"])))
        for key in ("stable", "active"):                 # getStableMapPointPositions / getActiveeMapPointPositions (larvio.h:86-87)
            if key in x:
                assert sorted(x[key]) == sorted(y[key]), "call %d: %s map points differ: %s vs %s" % (i, key, sorted(x[key]), sorted(y[key]))
                for k_ in y[key]:
                    worst["pts"] = max(worst["pts"], float(np.abs(np.asarray(x[key][k_]) - y[key][k_]).max())); worst["n_pts"] += 1
        worst["n"] += 1
    return worst


# ---- front-end fixtures (tests/golden/ref_fe_<case>.npz): what the reference's own ImageProcessor published -----------------
FE_CASES = {
    # name: (config overrides, sequence id, frames, synth kwargs, image edit)
    "fe_plain": (dict(max_features_in_one_grid=0), 0, 60, {}, None),
    "fe_blackout": (dict(max_features_in_one_grid=0), 0, 44, {}, "blackout"),            # every track lost for four frames, then repopulated
    "fe_failed_second": (dict(max_features_in_one_grid=0), 1, 30, {}, "grey_second"),     # initializeFirstFeatures fails -> back to FIRST_IMAGE
    "fe_400_tracks": (dict(max_features_in_one_grid=0, max_features_num=400, min_distance=14), 2, 24, {}, None),   # configs[4]'s front end
    "fe_static_start": (dict(max_features_in_one_grid=0), 3, 40, dict(static_until=1.0), None),
}


def fe_case_sequence(name):
    """(Config, Sequence, n_frames) of a front-end fixture, rebuilt deterministically from the synthetic generator."""
    import copy
    from larvio_b200.config import Config
    ov, sid, nf, kw, edit = FE_CASES[name]
    cfg = Config.load(os.path.join(ROOT, "configs", "euroc_mono.yaml"), **ov)
    seq = copy.copy(synth.make_sequence(cfg.raw, sid, nf, **kw))
    seq.images = seq.images.copy()
    if edit == "blackout":
        seq.images[20:24] = 117
    elif edit == "grey_second":
        seq.images[1] = 117
    return cfg, seq, nf


def pack_fe_fixture(name, seq, msgs):
    import hashlib
    has = np.array([m is not None for m in msgs], np.uint8)
    ofs = np.cumsum([0] + [0 if m is None else len(m["ids"]) for m in msgs]).astype(np.int64)
    ids = np.concatenate([m["ids"] for m in msgs if m is not None] or [np.zeros(0, np.uint64)]).astype(np.uint64)
    data = np.concatenate([m["data"].reshape(-1, 8) for m in msgs if m is not None] or [np.zeros((0, 8))])
    return dict(name=np.array(name), has=has, ofs=ofs, ids=ids, data=data, t=np.array([np.nan if m is None else m["t"] for m in msgs]),
                img_sha=np.frombuffer(hashlib.sha256(seq.images.tobytes()).digest(), np.uint8))


def load_fe_fixture(path):
    z = np.load(path)
    msgs = []
    for j in range(len(z["has"])):
        if not z["has"][j]:
            msgs.append(None); continue
        a, b = z["ofs"][j], z["ofs"][j + 1]
        msgs.append(dict(t=float(z["t"][j]), ids=z["ids"][a:b], data=z["data"][a:b]))
    return msgs, z["img_sha"]


def compare_fe(msgs, ref):
    """msgs: per frame None or dict(ids, data) of the implementation under test; ref: fixture.  Returns (published frames,
    frames whose ids or order differ, largest |difference| of the eight message columns over frames with equal ids)."""
    n_pub = 0; bad_ids = 0; worst = 0.0
    for j, (m, r) in enumerate(zip(msgs, ref)):
        assert (m is None) == (r is None), "frame %d: published %s, the reference %s" % (j, m is not None, r is not None)
        if r is None:
            continue
        n_pub += 1
        if len(m["ids"]) != len(r["ids"]) or not np.array_equal(np.asarray(m["ids"], np.uint64), r["ids"]):
            bad_ids += 1; continue
        if len(r["ids"]):
            worst = max(worst, float(np.abs(np.asarray(m["data"], np.float64) - r["data"]).max()))
    return n_pub, bad_ids, worst


# ---- the whole reference pipeline (oracle/_ref/larvio_ref_main): front end + static initialiser + filter, larvioMain's loop ------
REF_MAIN_BIN = os.path.join(ROOT, "oracle", "_ref", "larvio_ref_main")


def write_mav(dirpath, seq):
    """A synthetic sequence as an EuRoC ASL directory (PNG + csv, ns stamps) - what the replay tools read."""
    import cv2
    mav = os.path.join(str(dirpath), "mav0")
    os.makedirs(os.path.join(mav, "cam0", "data")); os.makedirs(os.path.join(mav, "imu0"))
    with open(os.path.join(mav, "cam0", "data.csv"), "w") as f:
        f.write("#timestamp [ns],filename\n")
        for t, im in zip(seq.img_t, seq.images):
            ns = int(round(t * 1e9)); cv2.imwrite(os.path.join(mav, "cam0", "data", "%d.png" % ns), im); f.write("%d,%d.png\r\n" % (ns, ns))
    with open(os.path.join(mav, "imu0", "data.csv"), "w") as f:
        f.write("#timestamp [ns],w_x,w_y,w_z,a_x,a_y,a_z\n")
        for r in seq.imu:
            f.write("%d,%.17g,%.17g,%.17g,%.17g,%.17g,%.17g\r\n" % (int(round(r[0] * 1e9)), *r[1:]))
    return mav


def parse_odometry_lines(text):
    """ODO / PTS lines of larvio_ref_main (rotation matrix) or larvio_shim_demo (quaternion x y z w) -> (odo rows, map-point lists)."""
    odo = [np.array(l.split()[1:], float) for l in text.splitlines() if l.startswith("ODO ")]
    pts = []
    for l in text.splitlines():
        if not l.startswith("PTS "):
            continue
        w = l.split()
        vals = np.array(w[3:], float).reshape(-1, 4)
        pts.append((w[1], {int(r[0]): r[1:4] for r in vals}))
    return odo, pts


def run_reference_pipeline(cfg_raw, mav_dir, with_logs=False):
    """The reference's main loop on the files of an EuRoC ASL directory (read like the replay tools read them).  with_logs: also
    return the two files LarVio itself writes into output_dir (larvio.cpp:388, 420-453): msckf_2_state.txt, msckf_2_takeoff.txt."""
    import sys
    from larvio_b200 import euroc
    if not os.path.exists(REF_MAIN_BIN):
        raise FileNotFoundError(REF_MAIN_BIN + " (build it with `make ref_main`; needs /root/reference)")
    ts = []; imgs = []; rows = []
    for t, img, r in euroc.Replay(str(mav_dir)):
        ts.append(t); imgs.append(img); rows.append(r)
    imgs = np.ascontiguousarray(np.stack(imgs), np.uint8); imu = np.concatenate(rows).reshape(-1, 7)
    with tempfile.TemporaryDirectory() as td:
        ypath = os.path.join(td, "cfg.yaml"); ipath = os.path.join(td, "in.bin")
        write_reference_yaml(cfg_raw, ypath, td + "/")
        with open(ipath, "wb") as f:
            np.array([len(ts), imgs.shape[1], imgs.shape[2], len(imu)], np.float64).tofile(f)
            np.asarray(ts, np.float64).tofile(f); np.ascontiguousarray(imu, np.float64).tofile(f); imgs.tofile(f)
        env = dict(os.environ, LVB_CV_SERVER=os.path.join(ROOT, "oracle", "cv_server.py"), LVB_CV_SERVER_PYTHON=sys.executable)
        r = subprocess.run([REF_MAIN_BIN, ypath, ipath], capture_output=True, text=True, timeout=3600, env=env)
        if r.returncode != 0:
            raise RuntimeError("larvio_ref_main failed (%d): %s" % (r.returncode, (r.stderr or r.stdout)[-3000:]))
        if os.environ.get("LVB_REF_TRACE"):
            print("\n".join(l for l in r.stderr.splitlines() if l.startswith("FRAME ")))
        logs = (open(os.path.join(td, "msckf_2_state.txt")).read(), open(os.path.join(td, "msckf_2_takeoff.txt")).read()) if with_logs else None
    txt = "\n".join(l for l in r.stdout.splitlines() if l.startswith(("ODO ", "PTS ")))
    return (txt, logs[0], logs[1]) if with_logs else txt


def quat_xyzw_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def compare_odometry(demo_text, ref_text):
    """larvio_shim_demo's lines (t q p v) against larvio_ref_main's (t R p v): same publications, same map-point lists."""
    odo, pts = parse_odometry_lines(demo_text); rodo, rpts = parse_odometry_lines(ref_text)
    assert len(odo) == len(rodo), "%d odometry messages, the reference published %d" % (len(odo), len(rodo))
    w = dict(n=len(odo), t=0.0, R=0.0, p=0.0, v=0.0, pts=0.0, n_lists=len(rpts))
    for a, b in zip(odo, rodo):
        w["t"] = max(w["t"], abs(a[0] - b[0]))
        w["R"] = max(w["R"], float(np.abs(quat_xyzw_to_rot(a[1:5]) - b[1:10].reshape(3, 3)).max()))
        w["p"] = max(w["p"], float(np.abs(a[5:8] - b[10:13]).max())); w["v"] = max(w["v"], float(np.abs(a[8:11] - b[13:16]).max()))
    assert len(pts) == len(rpts), "%d map-point lists, the reference returned %d" % (len(pts), len(rpts))
    for (ta, ma), (tb, mb) in zip(pts, rpts):
        assert ta == tb and sorted(ma) == sorted(mb), "map-point list %s differs: %s vs %s" % (tb, sorted(ma), sorted(mb))
        for k_ in mb:
            w["pts"] = max(w["pts"], float(np.abs(ma[k_] - mb[k_]).max()))
    return w


def run_oracle_pipeline(cfg_raw, mav_dir, with_logs=False):
    """oracle/frontend.py + oracle/initializer.py + oracle/backend.py behind the same loop, printing larvio_shim_demo's lines.
    with_logs: also the text of msckf_2_state.txt / msckf_2_takeoff.txt as the PRODUCT's writer (larvio_b200.euroc.state_line, the
    formatter of TrajectoryLog and of the replay tool's C++ twin) renders the same states."""
    from larvio_b200 import euroc
    from oracle.frontend import ImageProcessorOracle
    from oracle.backend import LarVioOracle
    from oracle.initializer import StaticInitializerOracle
    fe = ImageProcessorOracle(cfg_raw); be = LarVioOracle(cfg_raw); init = StaticInitializerOracle(cfg_raw)
    imu = []; lines = []; pubs = 0; first = False
    log_lines = []; take_off = None
    for t, img, rows in euroc.Replay(str(mav_dir)):
        imu.extend(rows.tolist())
        msg = fe.process_image(img, t, np.array(imu).reshape(-1, 7))
        if msg is None:
            continue
        if not be.is_gravity_set:
            if not first:                                       # larvio.cpp:365-372
                if len(imu) > 0 and imu[0][0] - msg.t - be.td <= 0.0:
                    first = True
                else:
                    continue
            o = init.try_inc_init(msg.ids, msg.data[:, :2], msg.t, np.array(imu).reshape(-1, 7))
            if o is None:
                continue
            be.set_initial_state(o["t"], o["q"], o["p"], o["v"], o["bg"], o["ba"])
            be.m_gyro_old = o["gyro_old"]; be.m_acc_old = o["acc_old"]
            del imu[:o["n_consumed"]]
            take_off = o["t"]
        if not be.process_features(msg, imu):
            continue
        s = be.imu_state
        log_lines.append(euroc.state_line(s.time - take_off, s.q, s.v, s.p, s.bg, s.ba, s.R_imu_cam0, s.t_cam0_imu))
        lines.append("ODO %.9f " % t + " ".join("%.17g" % x for x in list(s.q) + list(s.p) + list(s.v)))
        pubs += 1
        if pubs % 10 == 0:
            for tag, m in (("S", be.get_stable_map_points()), ("A", be.get_active_map_points())):
                if m:
                    lines.append("PTS %s %d " % (tag, len(m)) + " ".join("%d %.17g %.17g %.17g" % (k_, *m[k_]) for k_ in sorted(m)))
    if with_logs:
        return "\n".join(lines), "\n".join(log_lines) + "\n", "%.9f\n" % take_off
    return "\n".join(lines)
