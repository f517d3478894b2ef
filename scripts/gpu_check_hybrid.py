"""GPU check: hybrid MSCKF + EKF-SLAM (euroc defaults) fused step against the oracle, with per-frame diagnostics."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from larvio_b200.config import Config
from larvio_b200 import synth, api, harness
from oracle.frontend import ImageProcessorOracle
from oracle.backend import LarVioOracle

NF = int(os.environ.get('NF', '130')); IDS = [int(x) for x in os.environ.get('SEQS', '0,1').split(',')]; S = len(IDS)
cfg = Config.load('configs/euroc_mono.yaml', sw_size=int(os.environ.get('SW', '16')), calib_imu_instrinsic=int(os.environ.get('CALIB', '0')),
                  max_features_in_one_grid=int(os.environ.get('GRID', '1')))
seqs = [synth.make_sequence(cfg.raw, s, NF) for s in IDS]
STOP = int(os.environ.get('STOP', str(NF)))
b = api.Batch(cfg, n_seq=S)
fes = [ImageProcessorOracle(cfg.raw) for _ in range(S)]; bes = [LarVioOracle(cfg.raw) for _ in range(S)]
feed = harness.ImuFeeder(seqs, stride=128)
imu_o = [[] for _ in range(S)]; k = [0] * S; inited = [False] * S
worst = 0.0
for j in range(min(NF, STOP)):
    feed.push_until(j)
    msgs = []
    for s in range(S):
        k2 = synth.imu_window(seqs[s], k[s], seqs[s].img_t[j]); imu_o[s].extend(seqs[s].imu[k[s]:k2].tolist()); k[s] = k2
        msgs.append(fes[s].process_image(seqs[s].images[j], seqs[s].img_t[j], np.array(imu_o[s]).reshape(-1, 7)))
        if msgs[s] is not None and not inited[s]:
            a = (seqs[s].img_t[j], seqs[s].gt_q[j], seqs[s].gt_p[j], seqs[s].gt_v[j], np.zeros(3), np.zeros(3))
            bes[s].set_initial_state(*a); b.set_initial_state(s, *a); inited[s] = True
    imgs = np.stack([seqs[s].images[j] for s in range(S)]); t_img = np.array([seqs[s].img_t[j] for s in range(S)])
    if j >= int(os.environ.get('DBG_FROM', '10000')):
        os.environ['LVB_DEBUG_NAN'] = '1'; os.environ['LVB_DEBUG_FEAT'] = '1'
        print('--- frame', j, flush=True)
    try:
        ok = b.step(imgs, t_img, feed.buf, feed.n)
    except Exception as e:
        print('frame', j, 'GPU error', e); break
    for s in range(S):
        oko = bes[s].process_features(msgs[s], imu_o[s]) if msgs[s] is not None else False
        if not oko: continue
        st = b.get_state(s); P = b.get_covariance(s)
        dp = float(np.abs(st['p'] - bes[s].imu_state.p).max())
        same = P.shape == bes[s].P.shape
        rel = float(np.linalg.norm(P - bes[s].P) / np.linalg.norm(bes[s].P)) if same else -1
        worst = max(worst, dp)
        if os.environ.get('VERBOSE') and same:
            D = np.abs(P - bes[s].P); L = bes[s].LEG
            cal = b.get_calibration(s)
            print('  frame', j, 'seq', s, 'dp %.2e' % dp, 'Prel %.2e' % rel, 'maxdP LL %.2e LA %.2e AA %.2e' % (D[:L, :L].max(), D[:L, L:].max() if P.shape[0] > L else 0, D[L:, L:].max() if P.shape[0] > L else 0),
                  'dTg %.2e dbg %.2e dba %.2e' % (np.abs(cal['Tg'] - bes[s].Tg).max(), np.abs(st['bg'] - bes[s].imu_state.bg).max(), np.abs(st['ba'] - bes[s].imu_state.ba).max()),
                  'upd', bes[s].stats.get('updates'), flush=True)
            if j >= int(os.environ.get('DBG_FROM', '10000')):
                for g in bes[s].stats.get('gates', []): print('  [oracle gate] dof %d gamma %.9e chi2 %.6e %s' % (g[0], g[1], g[2], 'pass' if g[1] < g[2] else 'REJECT'))
        if j > 98 and (j % 2 == 0) or not same or not (dp <= 1e-7):
            print('frame', j, 'seq', s, 'dp %.2e' % dp, 'Prel %.2e' % rel, 'dims', P.shape[0], bes[s].P.shape[0], 'nslam', len(bes[s].feature_states),
                  {a: c for a, c in bes[s].stats.items() if a in ('n_ekf_new', 'n_ekf', 'n_msckf_features', 'n_ekf_lost', 'anchor_changes', 'prune_used')},
                  'icore', b.debug_icore(s)[:28], 'nanP', int(np.isnan(P).sum()), flush=True)
        if not same or not (dp <= 1e-3):
            print('DIVERGED'); sys.exit(0)
print(json.dumps(dict(worst_dp=worst)))
