"""GPU bring-up check 3: back end (lvb_process_features) against the numpy oracle, fed with the oracle
front end's feature messages; then the fused lvb_step against oracle FE+BE."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from larvio_b200.config import Config
from larvio_b200 import synth, api
from oracle.frontend import ImageProcessorOracle
from oracle.backend import LarVioOracle

NF = int(os.environ.get('NF', '60'))
S = int(os.environ.get('S', '2'))
SW = int(os.environ.get('SW', '12'))
cfg = Config.load('configs/euroc_mono.yaml', max_features_in_one_grid=0, sw_size=SW)
seqs = [synth.make_sequence(cfg.raw, s, NF) for s in range(S)]

def run(mode):
    b = api.Batch(cfg, n_seq=S)
    fes = [ImageProcessorOracle(cfg.raw) for _ in range(S)]
    bes = [LarVioOracle(cfg.raw) for _ in range(S)]
    imu_o = [[] for _ in range(S)]          # oracle buffers
    imu_g = [np.zeros((0, 7)) for _ in range(S)]   # gpu-side caller buffers
    k = [0] * S
    inited = [False] * S
    worst = dict(p=0.0, q=0.0, v=0.0, bg=0.0, ba=0.0, Prel=0.0, dim_mismatch=0, ok_mismatch=0, nimu_mismatch=0, frames=0)
    for j in range(NF):
        msgs = []
        for s in range(S):
            k2 = synth.imu_window(seqs[s], k[s], seqs[s].img_t[j])
            new = seqs[s].imu[k[s]:k2]
            imu_o[s].extend(new.tolist()); imu_g[s] = np.concatenate([imu_g[s], new]); k[s] = k2
            msgs.append(fes[s].process_image(seqs[s].images[j], seqs[s].img_t[j], np.array(imu_o[s]).reshape(-1, 7)))
        imu, n_imu = api.Batch.pack_imu(imu_g, stride=96)
        if mode == 'be':
            valid = np.array([m is not None for m in msgs], np.uint8)
            if not valid.any():
                continue
            feat = np.zeros((S, b.cap), api.FEATURE_DTYPE); n_feat = np.zeros(S, np.int32); t_msg = np.zeros(S)
            for s, m in enumerate(msgs):
                if m is None: continue
                n = len(m.ids); n_feat[s] = n; t_msg[s] = m.t
                feat['id'][s, :n] = m.ids
                for c, name in enumerate(['u', 'v', 'u_init', 'v_init', 'u_vel', 'v_vel', 'u_init_vel', 'v_init_vel']):
                    feat[name][s, :n] = m.data[:, c]
        for s in range(S):
            if msgs[s] is not None and not inited[s]:
                a = (seqs[s].img_t[j], seqs[s].gt_q[j], seqs[s].gt_p[j], seqs[s].gt_v[j], np.zeros(3), np.zeros(3))
                bes[s].set_initial_state(*a); b.set_initial_state(s, *a); inited[s] = True
        if mode == 'be':
            ok = b.process_features(valid, t_msg, feat, n_feat, imu, n_imu)
        else:
            imgs = np.stack([seqs[s].images[j] for s in range(S)])
            ok = b.step(imgs, np.array([seqs[s].img_t[j] for s in range(S)]), imu, n_imu)
        for s in range(S):
            imu_g[s] = np.stack([imu['t'][s, :n_imu[s]]] + [imu['gyro'][s, :n_imu[s], c] for c in range(3)] + [imu['acc'][s, :n_imu[s], c] for c in range(3)], 1) if n_imu[s] else np.zeros((0, 7))
            oko = bes[s].process_features(msgs[s], imu_o[s]) if msgs[s] is not None else False
            if bool(ok[s]) != bool(oko): worst['ok_mismatch'] += 1
            if len(imu_o[s]) != n_imu[s]: worst['nimu_mismatch'] += 1
            if not oko: continue
            worst['frames'] += 1
            st = b.get_state(s); o = bes[s].imu_state
            worst['p'] = max(worst['p'], float(np.abs(st['p'] - o.p).max())); worst['v'] = max(worst['v'], float(np.abs(st['v'] - o.v).max()))
            dq = min(np.abs(st['q'] - o.q).max(), np.abs(st['q'] + o.q).max()); worst['q'] = max(worst['q'], float(dq))
            worst['bg'] = max(worst['bg'], float(np.abs(st['bg'] - o.bg).max())); worst['ba'] = max(worst['ba'], float(np.abs(st['ba'] - o.ba).max()))
            P = b.get_covariance(s)
            if P.shape != bes[s].P.shape: worst['dim_mismatch'] += 1
            else: worst['Prel'] = max(worst['Prel'], float(np.linalg.norm(P - bes[s].P) / np.linalg.norm(bes[s].P)))
        if j % 10 == 0:
            print(mode, 'frame', j, {a: (('%.2e' % c) if isinstance(c, float) else c) for a, c in worst.items()}, 'err', float(np.linalg.norm(bes[0].imu_state.p - seqs[0].gt_p[j])), flush=True)
    worst['mode'] = mode
    print(json.dumps(worst))
    b.close()

t0 = time.time()
run('be')
print('be sec', time.time() - t0)
t0 = time.time()
run('step')
print('step sec', time.time() - t0)
