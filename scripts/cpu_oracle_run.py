"""Run the full CPU oracle (front end + back end) on one synthetic sequence and report drift."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from larvio_b200.config import Config
from larvio_b200 import synth
from oracle.frontend import ImageProcessorOracle
from oracle.backend import LarVioOracle

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = Config.load('configs/euroc_mono.yaml', max_features_in_one_grid=int(os.environ.get('GRID', '0')), sw_size=int(os.environ.get('SW', '30')),
                  calib_imu_instrinsic=int(os.environ.get('CALIB', '0'))).raw
seq = synth.make_sequence(cfg, int(os.environ.get('SEQ', '0')), NF)
fe = ImageProcessorOracle(cfg); be = LarVioOracle(cfg)
imu = []; k = 0
t_fe = t_be = 0.0
err = []
for j in range(NF):
    k2 = synth.imu_window(seq, k, seq.img_t[j])
    imu.extend(seq.imu[k:k2].tolist()); k = k2
    t0 = time.time()
    msg = fe.process_image(seq.images[j], seq.img_t[j], np.array(imu).reshape(-1, 7))
    t_fe += time.time() - t0
    if msg is None:
        continue
    if not be.is_gravity_set:
        be.set_initial_state(seq.img_t[j], seq.gt_q[j], seq.gt_p[j], seq.gt_v[j], np.zeros(3), np.zeros(3))
    t0 = time.time()
    ok = be.process_features(msg, imu)
    t_be += time.time() - t0
    if ok:
        e = np.linalg.norm(be.imu_state.p - seq.gt_p[j])
        err.append(e)
        if j % 10 == 0:
            print(j, 'n_feat', len(msg.ids), 'win', len(be.aug), 'd', be.P.shape[0], 'pos_err %.4f' % e,
                  'stats', {a: b for a, b in be.stats.items() if a != 'm_hist'}, flush=True)
print(json.dumps(dict(frames=NF, rmse=float(np.sqrt(np.mean(np.square(err)))), final_err=float(err[-1]), t_fe=t_fe, t_be=t_be,
                      bg_true=seq.gyro_bias.tolist(), bg_est=be.imu_state.bg.tolist(), ba_true=seq.acc_bias.tolist(), ba_est=be.imu_state.ba.tolist())))
