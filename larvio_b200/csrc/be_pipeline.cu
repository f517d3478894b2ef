// Batched LarVio::processFeatures (larvio.cpp:363-461), FP64: pure MSCKF and hybrid MSCKF + EKF-SLAM features (1-D or 3-D
// inverse depth), ZUPT, online extrinsics / td / IMU-intrinsic calibration; Schmidt nuisance states are refused.
//   be_propagate_kernel   batchImuProcessing / processModel / predictNewState / calPhi (:464-649, :3475-3530)
//   be_add_obs_kernel     addFeatureObservations (:804-856) + the checkZUPT detector (:2751-2766)
//   be_augment_kernel     stateAugmentation (:720-801)
//   be_classify_kernel    removeLostFeatures' classifier + Feature::checkMotion / initializePosition
//                         (:1926-2005, feature.hpp:334-552) and pruneImuStateBuffer's selection (:2331-2489)
//   be_feature_kernel     measurementJacobian_msckf / featureJacobian_msckf / gatingTest (:859-981, :1865-1880)
//   be_stack_kernel       stacking of the gated blocks (:2185-2204, :2494-2536)
//   be_qr_kernel          thin-QR measurement compression (:1430-1449, :2209-2229)
//   be_update_*           measurementUpdate_hybrid / _msckf (:1420-1602, :1605-1862) as
//                         T=HP, S=TH^T+s^2 I, S=LL^T, Y=L^-1 T, dx=Y^T L^-1 r, P-=Y^T Y
//   be_prune_*            findRedundantImuStates / pruneImuStateBuffer (:2259-2307, :2310-2641)
//   be_slam_*, be_anchor_kernel, be_remap_*   the hybrid filter: promotion rule (:1968-2002), featureJacobian_ekf[_new] with
//                         measurementJacobian_ekf_1didp / _3didp (:984-1417), new-feature states and grown covariance
//                         (:1661-1676, :1821-1854), anchor hand-over with updateFeatureCov_1didp / _3didp (:2345-2461,
//                         :2965-3293), rmLostFeaturesCov (:3296-3348), map-point lists (:455-458, :3342)
// One CTA per sequence for the sequential parts, (tiles x sequences) grids for the dense algebra.
#include <math.h>
#include <string.h>
#include "be_state.h"
#include <cstring>
#include "be_math.cuh"

#define RC(x) do { int rc_ = (x); if (rc_ != LVB_OK) return rc_; } while (0)

namespace {

__constant__ double c_chi2[100] = {
#include "be_chi2.inc"
};

struct BeCfg {
  double th_imu;              // imu_img_timeTh = 1/(2*imu_rate)
  double sg2, sa2, sbg2, sba2, sfeat2;
  double rot_thr, trans_thr, track_thr, feat_trans_thr, zupt_dis, zupt_nv, zupt_np, zupt_nq;
  int max_track_len, sw_size, least_obs, if_FEJ_config, estimate_td, if_ZUPT_valid;
  int hybrid;                 // EKF-SLAM features enabled (1-D or 3-D inverse depth: v.be.IDP)
  double x_min, y_min, grid_w, grid_h;
};

struct BeView {
  LvbBackEnd be;     // by value: device pointers + dims
  BeCfg cfg;
  const LvbFeature* msg; const int* msg_n; const double* msg_t; const uint8_t* msg_valid; int msg_stride;
  double* P;         // current covariance buffer
  double* Pn;        // the other buffer
};

#define LEGD (v.be.LEG)
__device__ __forceinline__ double* core_of(const BeView& v, int s) { return v.be.core + (size_t)s * BE_CORE; }
__device__ __forceinline__ int* icore_of(const BeView& v, int s) { return v.be.icore + (size_t)s * BE_ICORE; }
__device__ __forceinline__ double* win_of(const BeView& v, int s, int slot) { return v.be.win + ((size_t)s * v.be.Wcap + slot) * BE_WIN; }
__device__ __forceinline__ double* P_of(const BeView& v, int s) { return v.P + (size_t)s * v.be.LD * v.be.LD; }
// anchor pose of a SLAM feature: a window slot, or (use_schmidt) a frozen nuisance state (larvio.cpp:1000-1010, 1543-1549)
__device__ __forceinline__ const double* anchor_rec(const BeView& v, int s, int as) {
  return as >= BE_NUI_BASE ? v.be.nui_win + ((size_t)s * v.be.NUI + (as - BE_NUI_BASE)) * BE_WIN : v.be.win + ((size_t)s * v.be.Wcap + as) * BE_WIN;
}
// first covariance column of that pose: the nuisance block follows the nf feature blocks that are in the covariance (:1351-1366)
__device__ __forceinline__ int anchor_col(const BeView& v, int as, int n_win, int nf) {
  return as >= BE_NUI_BASE ? v.be.LEG + 6 * n_win + v.be.IDP * nf + 6 * (as - BE_NUI_BASE) : v.be.LEG + 6 * as;
}

// ====================================================================== propagate
// selector matrices of calPhi's IMU-intrinsic blocks (larvio.cpp:3534-3629): lower / diagonal / upper placement of a vector
__device__ __forceinline__ M3 imu_selector(int kind, V3 x) {
  M3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.0;
  if (kind == 0) { r.m[3] = x.x; r.m[7] = x.x; r.m[8] = x.y; }
  else if (kind == 1) { r.m[0] = x.x; r.m[4] = x.y; r.m[8] = x.z; }
  else { r.m[0] = x.y; r.m[1] = x.z; r.m[5] = x.z; }
  return r;
}
// per-sample record left by the sequential nominal-state pass for the parallel construction of Phi_k
enum { R_DT = 0, R_Q = 1, R_VEL = 5, R_POS = 8, R_VNEW = 11, R_PNEW = 14, R_FOV = 17, R_FOP = 20, R_FNV = 23, R_FNP = 26,
       R_GYRO = 29, R_ACC = 32, R_GYRO_OLD = 35, R_ACC_OLD = 38, PROP_REC = 42 };
template <int L> struct PropChunk { static constexpr int N = (L <= 22) ? 8 : 4; };   // samples whose Phi is built concurrently

// calPhi (larvio.cpp:3475-3800) for one IMU sample from its record: fills Phi (L x L, row pitch L + 1).
// role 0 writes the 15-column core blocks, roles 1..8 (L = 46 only) one group of three IMU-intrinsic columns each; the caller has
// set Phi to the identity.
template <int L>
__device__ void prop_build_phi(const double* rec, const double* core, bool fej, double* Phi_, int role) {
  constexpr int LP = L + 1;
#define Phi(r, c) Phi_[(r) * LP + (c)]
  const double dtime = rec[R_DT];
  const V3 bg = ld3(core + C_BG), ba = ld3(core + C_BA);
  const M3 Tg = m3_load(core + C_TG), As = m3_load(core + C_AS), Ma = m3_load(core + C_MA);
  const V3 m_gyro = ld3(rec + R_GYRO), m_acc = ld3(rec + R_ACC);
  const V3 f = m_acc - ba, acc = m3_vec(Ma, f);
  const V3 w = m_gyro - m3_vec(As, acc) - bg, gyro = m3_vec(Tg, w);
  const V3 f_old = ld3(rec + R_ACC_OLD) - ba, acc_old = m3_vec(Ma, f_old);
  const V3 w_old = ld3(rec + R_GYRO_OLD) - m3_vec(As, acc_old) - bg, gyro_old = m3_vec(Tg, w_old);
  const double q[4] = {rec[R_Q], rec[R_Q + 1], rec[R_Q + 2], rec[R_Q + 3]};
  const M3 Rq = quat_to_rot(q);
  const V3 vel = ld3(rec + R_VEL), pos = ld3(rec + R_POS), vnew = ld3(rec + R_VNEW), pnew = ld3(rec + R_PNEW);
  const V3 g = v3(0, 0, -9.81);
  const V3 axis = (gyro_old + gyro) * (dtime * 0.5) + cross(gyro_old, gyro) * (dtime * dtime / 12);
  const M3 Ah = skew(axis);
  const M3 C = Rq;       // C_bk2w from imu_state_old.orientation
  V3 vk, pk, vk1, pk1;
  if (fej) { vk = ld3(rec + R_FOV); pk = ld3(rec + R_FOP); vk1 = ld3(rec + R_FNV); pk1 = ld3(rec + R_FNP); }
  else { vk = vel; pk = pos; vk1 = vnew; pk1 = pnew; }
  const M3 I3 = m3_identity();
  const M3 TA = m3_mul(Tg, As), TAM = m3_mul(TA, Ma);
  if (role == 0) {
  const M3 twoIAh = m3_add(m3_scale(I3, 2.0), Ah);
  const M3 CtA = m3_scale(m3_mul(C, twoIAh), 0.5 * dtime);       // 0.5*C*(2I+Ah)*dtime
  const M3 Pqbg = m3_scale(m3_mul(CtA, Tg), -1.0);
  const M3 Pqba = m3_mul(CtA, TAM);
  const M3 Pvq = m3_scale(skew(vk1 - vk - g * dtime), -1.0);
  const M3 Pvbg = m3_add(m3_mul(skew(pk - pk1 + vk1 * dtime - g * (0.5 * dtime * dtime)), C),
                         m3_mul(m3_mul(skew(pk * 0.5 - pk1 * 0.5 + vk1 * (0.5 * dtime) - g * (dtime * dtime / 6)), C), Ah));
  const M3 Pvba = m3_sub(m3_scale(m3_mul(CtA, Ma), -1.0), m3_mul(Pvbg, TAM));
  const M3 Ppq = m3_scale(skew(pk1 - pk - vk * dtime - g * (0.5 * dtime * dtime)), -1.0);
  const M3 Ppbg = m3_add(m3_scale(m3_mul(skew(g), C), -dtime * dtime * dtime / 6),
                         m3_scale(m3_mul(m3_mul(skew(pk1 - pk - g * (dtime * dtime / 6)), C), Ah), dtime / 4));
  const M3 Ppba = m3_sub(m3_mul(m3_scale(m3_mul(C, m3_add(m3_scale(I3, 3.0), Ah)), -dtime * dtime / 6), Ma), m3_mul(Ppbg, TAM));
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Phi(r, 9 + c) = Pqbg.m[r * 3 + c];
      Phi(r, 12 + c) = Pqba.m[r * 3 + c];
      Phi(3 + r, c) = Pvq.m[r * 3 + c];
      Phi(3 + r, 9 + c) = Pvbg.m[r * 3 + c];
      Phi(3 + r, 12 + c) = Pvba.m[r * 3 + c];
      Phi(6 + r, c) = Ppq.m[r * 3 + c];
      Phi(6 + r, 3 + c) = (r == c) ? dtime : 0.0;
      Phi(6 + r, 9 + c) = Ppbg.m[r * 3 + c];
      Phi(6 + r, 12 + c) = Ppba.m[r * 3 + c];
    }
  }
  if constexpr (L > 22) {
    if (role == 0) return;
    // ---- IMU-intrinsic columns (:3532-3797): 8 groups of 3 columns; selectors Lo/Di/Up place the components of
    // (w | acc | f) sampled at k, k+1/2, k+1; Simpson weights for q, the reference's RK4 weights for v and p
    const V3 f_mid = (f + f_old) * 0.5, acc_mid = (acc + acc_old) * 0.5;
    const V3 w_mid = (w_old + w) * 0.5 + cross(w_old, w) * (dtime / 12);
    const M3 R_mid = m3_add(I3, m3_scale(Ah, 0.5)), R_kp1 = m3_add(I3, Ah);
    const M3 S_mid = m3_scale(skew(m3_vec(R_mid, acc_mid)), dtime * 0.5), S_kp1 = skew(m3_vec(R_kp1, acc));
    {
      const int gi = role - 1;
      const int kind = (gi < 6) ? gi % 3 : gi - 6;                 // 0 Lo, 1 Di, 2 Up
      const V3 xk = gi < 3 ? w_old : (gi < 6 ? acc_old : f_old);
      const V3 xh = gi < 3 ? w_mid : (gi < 6 ? acc_mid : f_mid);
      const V3 xp = gi < 3 ? w : (gi < 6 ? acc : f);
      const M3 Lf = gi < 3 ? I3 : (gi < 6 ? Tg : TA);
      const double sgn = gi < 3 ? 1.0 : -1.0;
      const bool direct = gi >= 6;
      const M3 sk = imu_selector(kind, xk), sh = imu_selector(kind, xh), sp = imu_selector(kind, xp);
      const M3 kq1 = m3_mul(Lf, sk), kq2 = m3_mul(R_mid, m3_mul(Lf, sh)), kq4 = m3_mul(R_kp1, m3_mul(Lf, sp));
      const M3 Rq2 = m3_scale(m3_add(m3_add(kq1, m3_scale(kq2, 4.0)), kq4), dtime / 6);
      M3 kv1, kv2, kv3, kv4;
      if (!direct) {
        kv1 = m3_scale(I3, 0.0); kv2 = m3_mul(S_mid, kq1); kv3 = m3_mul(S_mid, kq2); kv4 = m3_mul(S_kp1, Rq2);
      } else {
        const M3 Rh = m3_mul(R_mid, sh);
        kv1 = sk; kv2 = m3_add(Rh, m3_mul(S_mid, kq1)); kv3 = m3_add(Rh, m3_mul(S_mid, kq2));
        kv4 = m3_add(m3_mul(R_kp1, sp), m3_mul(S_kp1, Rq2));
      }
      const M3 fR = m3_scale(m3_add(m3_add(kv1, m3_scale(m3_add(kv2, kv3), 2.0)), kv4), dtime / 6);
      const double vs = direct ? 1.0 : -sgn;
      const M3 kp = m3_scale(m3_add(m3_scale(m3_add(kv1, kv2), dtime), fR), dtime / 6);   // kp1=0, kp2=dt kv1/2, kp3=dt kv2/2, kp4=fR
      const M3 Bq = m3_scale(m3_mul(C, Rq2), sgn), Bv = m3_scale(m3_mul(C, fR), vs), Bp = m3_scale(m3_mul(C, kp), vs);
      const int col = 22 + 3 * gi;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { Phi(r, col + c) = Bq.m[r * 3 + c]; Phi(3 + r, col + c) = Bv.m[r * 3 + c]; Phi(6 + r, col + c) = Bp.m[r * 3 + c]; }
    }
  }
#undef Phi
}

// batchImuProcessing / processModel / predictNewState / calPhi for one sequence per CTA.  Three phases per chunk of samples:
// (1) one thread integrates the nominal state sample by sample (sequential by nature: quaternion + the reference's RK4) and
// leaves a record per sample; (2) one thread per sample builds its transition matrix Phi_k from the record - the expensive
// part (calPhi, with the 24 IMU-intrinsic columns when L = 46), now concurrent; (3) all threads run the covariance recursion
// P_LL <- Phi (P_LL + D dt) Phi^T, PhiTot <- Phi PhiTot sample by sample.  The composed PhiTot is applied ONCE to the
// cross-covariance at the end.
template <int L>
__global__ void __launch_bounds__(256) be_propagate_kernel(BeView v) {
  extern __shared__ double psm[];                      // PhiC[CHK], PLL, Tmp, PhiTot: [L][L+1] each; Ddiag [L]; rec[CHK][PROP_REC]
  constexpr int CHK = PropChunk<L>::N;
  __shared__ int s_ok, s_used, s_stop, s_nact;
  __shared__ int s_act[CHK];
  __shared__ double s_dt;
  const int s = blockIdx.x, tid = threadIdx.x;
  double* core = core_of(v, s);
  int* ic = icore_of(v, s);
  const LvbImu* imu = v.be.imu + (size_t)s * v.be.imu_cap;
  const int n_imu = v.be.n_imu[s];
  constexpr int LP = L + 1;
  const int LD = v.be.LD;
  double* const PhiC_ = psm; double* const PLL_ = psm + CHK * L * LP; double* const Tmp_ = PLL_ + L * LP; double* const PhiTot_ = Tmp_ + L * LP;
  double* const Ddiag = PhiTot_ + L * LP;
  double* const rec_ = Ddiag + L;
#define PLL(r, c) PLL_[(r) * LP + (c)]
#define Tmp(r, c) Tmp_[(r) * LP + (c)]
#define PhiTot(r, c) PhiTot_[(r) * LP + (c)]
  double* P = P_of(v, s);
  if (tid == 0) {
    int ok = v.msg_valid[s] != 0;
    if (ok && !ic[I_FIRST]) {                                   // larvio.cpp:366-372
      if (n_imu > 0 && imu[0].t - v.msg_t[s] - core[C_TD] <= 0.0) ic[I_FIRST] = 1;
      else ok = 0;
    }
    if (ok && !ic[I_GRAVITY]) ok = 0;                            // :375-391 (initialiser out of scope)
    s_ok = ok; ic[I_OK] = ok; ic[I_CONSUMED] = 0; s_used = 0; s_dt = 0.0; s_stop = 0;
    ic[I_UPDATES] = 0;
  }
  if (tid < L) {
    double dv = 0.0;
    if (tid < 3) dv = v.cfg.sg2; else if (tid < 6) dv = v.cfg.sa2;
    else if (tid >= 9 && tid < 12) dv = v.cfg.sbg2; else if (tid >= 12 && tid < 15) dv = v.cfg.sba2;
    Ddiag[tid] = dv;
  }
  __syncthreads();
  if (!s_ok) return;
  for (int i = tid; i < L * L; i += blockDim.x) {
    const int r = i / L, c = i - r * L;
    PLL(r, c) = P[(size_t)r * LD + c];
    PhiTot(r, c) = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const double time_bound = v.msg_t[s] + core[C_TD];
  const bool fej = ic[I_FEJ] != 0;
  for (int k0 = 0; k0 < n_imu; k0 += CHK) {
    // ---- phase 1: nominal state of up to CHK samples (processModel :520-578 without Phi, predictNewState :581-649)
    if (tid == 0) {
      int nact = 0;
      for (int kk = 0; kk < CHK && k0 + kk < n_imu; ++kk) {
        const int k = k0 + kk;
        const double t = imu[k].t;
        if (t <= core[C_TIME]) { s_used++; continue; }           // already covered by the state
        if (t - time_bound > v.cfg.th_imu) { s_stop = 1; break; }
        s_used++;
        s_dt = t - time_bound;
        double* rec = rec_ + nact * PROP_REC;
        const V3 m_gyro = v3(imu[k].gyro[0], imu[k].gyro[1], imu[k].gyro[2]);
        const V3 m_acc = v3(imu[k].acc[0], imu[k].acc[1], imu[k].acc[2]);
        if (!ic[I_HAVE_OLD]) { st3(core + C_GYRO_OLD, m_gyro); st3(core + C_ACC_OLD, m_acc); ic[I_HAVE_OLD] = 1; }
        // acc = Ma f, w = m_gyro - As acc - bg, gyro = Tg w
        const V3 bg = ld3(core + C_BG), ba = ld3(core + C_BA);
        const M3 Tg = m3_load(core + C_TG), As = m3_load(core + C_AS), Ma = m3_load(core + C_MA);
        const V3 f = m_acc - ba, acc = m3_vec(Ma, f);
        const V3 w = m_gyro - m3_vec(As, acc) - bg, gyro = m3_vec(Tg, w);
        st3(rec + R_GYRO, m_gyro); st3(rec + R_ACC, m_acc);
        st3(rec + R_GYRO_OLD, ld3(core + C_GYRO_OLD)); st3(rec + R_ACC_OLD, ld3(core + C_ACC_OLD));
        const double dtime = t - core[C_TIME];
        rec[R_DT] = dtime;
        const double gn = norm(gyro);
        double q[4] = {core[C_Q], core[C_Q + 1], core[C_Q + 2], core[C_Q + 3]};
        const V3 vel = ld3(core + C_V), pos = ld3(core + C_P);
        for (int i = 0; i < 4; ++i) { core[C_OLD_Q + i] = q[i]; rec[R_Q + i] = q[i]; }
        st3(core + C_OLD_P, pos); st3(core + C_OLD_V, vel);
        st3(rec + R_VEL, vel); st3(rec + R_POS, pos);
        const V3 qv = v3(q[0], q[1], q[2]);
        const V3 ov = cross(qv, gyro) + gyro * q[3];     // (Omega q).head<3>()
        const double ow = -dot(gyro, qv);                // (Omega q)(3)
        double dq[4], dq2[4];
        if (gn > 1e-5) {
          const double c1 = cos(gn * dtime * 0.5), s1 = 1 / gn * sin(gn * dtime * 0.5);
          const double c2 = cos(gn * dtime * 0.25), s2 = 1 / gn * sin(gn * dtime * 0.25);
          dq[0] = c1 * q[0] + s1 * ov.x; dq[1] = c1 * q[1] + s1 * ov.y; dq[2] = c1 * q[2] + s1 * ov.z; dq[3] = c1 * q[3] + s1 * ow;
          dq2[0] = c2 * q[0] + s2 * ov.x; dq2[1] = c2 * q[1] + s2 * ov.y; dq2[2] = c2 * q[2] + s2 * ov.z; dq2[3] = c2 * q[3] + s2 * ow;
        } else {
          const double c1 = cos(gn * dtime * 0.5), c2 = cos(gn * dtime * 0.25);
          dq[0] = (q[0] + 0.5 * dtime * ov.x) * c1; dq[1] = (q[1] + 0.5 * dtime * ov.y) * c1;
          dq[2] = (q[2] + 0.5 * dtime * ov.z) * c1; dq[3] = (q[3] + 0.5 * dtime * ow) * c1;
          dq2[0] = (q[0] + 0.25 * dtime * ov.x) * c2; dq2[1] = (q[1] + 0.25 * dtime * ov.y) * c2;
          dq2[2] = (q[2] + 0.25 * dtime * ov.z) * c2; dq2[3] = (q[3] + 0.25 * dtime * ow) * c2;
        }
        const M3 dR = quat_to_rot(dq), dR2 = quat_to_rot(dq2), Rq = quat_to_rot(q);
        const V3 g = v3(0, 0, -9.81);
        const V3 k1v = m3_vec(Rq, acc) + g, k1p = vel;
        const V3 k1_v = vel + k1v * dtime * 0.5;
        const V3 k2v = m3_vec(dR2, acc) + g, k2p = k1_v;
        const V3 k2_v = vel + k2v * dtime * 0.5;
        const V3 k3v = m3_vec(dR2, acc) + g, k3p = k2_v;
        const V3 k3_v = vel + k3v * dtime;
        const V3 k4v = m3_vec(dR, acc) + g, k4p = k3_v;
        const double qn = 1.0 / sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
        for (int i = 0; i < 4; ++i) core[C_Q + i] = dq[i] * qn;
        const V3 vnew = vel + (k1v + 2.0 * k2v + 2.0 * k3v + k4v) * (dtime / 6);
        const V3 pnew = pos + (k1p + 2.0 * k2p + 2.0 * k3p + k4p) * (dtime / 6);
        st3(core + C_V, vnew); st3(core + C_P, pnew);
        st3(rec + R_VNEW, vnew); st3(rec + R_PNEW, pnew);
        // FEJ bookkeeping (:644-646)
        st3(core + C_FOLD_P, ld3(core + C_FNOW_P)); st3(core + C_FOLD_V, ld3(core + C_FNOW_V));
        st3(core + C_FNOW_P, pnew); st3(core + C_FNOW_V, vnew);
        st3(rec + R_FOP, ld3(core + C_FOLD_P)); st3(rec + R_FOV, ld3(core + C_FOLD_V));
        st3(rec + R_FNP, pnew); st3(rec + R_FNV, vnew);
        core[C_TIME] = t;
        st3(core + C_GYRO_OLD, m_gyro); st3(core + C_ACC_OLD, m_acc);
        ++nact;
      }
      s_nact = nact;
    }
    __syncthreads();
    const int nact = s_nact;
    // ---- phase 2: Phi_k of every sample of the chunk: identity by all threads, then one thread per (sample, column group)
    for (int i = tid; i < nact * L * L; i += blockDim.x) {
      const int a = i / (L * L), e = i - a * (L * L), r = e / L, c = e - r * L;
      PhiC_[(size_t)a * L * LP + r * LP + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    {
      constexpr int NROLE = (L > 22) ? 9 : 1;
      if (tid < nact * NROLE) {
        const int a = tid / NROLE, role = tid - a * NROLE;
        prop_build_phi<L>(rec_ + a * PROP_REC, core, fej, PhiC_ + (size_t)a * L * LP, role);
      }
    }
    __syncthreads();
    // ---- phase 3: covariance recursion, sample by sample
    for (int a = 0; a < nact; ++a) {
      const double* Phi_ = PhiC_ + (size_t)a * L * LP;
#define Phi(r, c) Phi_[(r) * LP + (c)]
      const double dtime = rec_[a * PROP_REC + R_DT];
      // Tmp = Phi * (PLL + D*dtime)
      for (int i = tid; i < L * L; i += blockDim.x) {
        const int r = i / L, c = i - r * L;
        double acc = 0.0;
        for (int k2 = 0; k2 < L; ++k2) acc += Phi(r, k2) * (PLL(k2, c) + (k2 == c ? Ddiag[c] * dtime : 0.0));
        Tmp(r, c) = acc;
      }
      __syncthreads();
      for (int i = tid; i < L * L; i += blockDim.x) {
        const int r = i / L, c = i - r * L;
        double acc = 0.0;
        for (int k2 = 0; k2 < L; ++k2) acc += Tmp(r, k2) * Phi(c, k2);
        PLL(r, c) = acc;
      }
      __syncthreads();
      for (int i = tid; i < L * L; i += blockDim.x) {
        const int r = i / L, c = i - r * L;
        double acc = 0.0;
        for (int k2 = 0; k2 < L; ++k2) acc += Phi(r, k2) * PhiTot(k2, c);
        Tmp(r, c) = acc;
        if (r < c) { const double m = 0.5 * (PLL(r, c) + PLL(c, r)); PLL(r, c) = m; PLL(c, r) = m; }
      }
      __syncthreads();
      for (int i = tid; i < L * L; i += blockDim.x) { const int r = i / L, c = i - r * L; PhiTot(r, c) = Tmp(r, c); }
      __syncthreads();
#undef Phi
    }
    if (s_stop) break;
  }
  __syncthreads();
  // write back P_LL, apply the composed transition to the cross terms
  const int d = ic[I_DIM];
  for (int i = tid; i < L * L; i += blockDim.x) { const int r = i / L, c = i - r * L; P[(size_t)r * LD + c] = PLL(r, c); }
  for (int c = L + tid; c < d; c += blockDim.x) {
    double col[BE_LEG_MAX];
    for (int r = 0; r < L; ++r) col[r] = P[(size_t)r * LD + c];
    for (int r = 0; r < L; ++r) {
      double acc = 0.0;
      for (int k2 = 0; k2 < L; ++k2) acc += PhiTot(r, k2) * col[k2];
      P[(size_t)r * LD + c] = acc;
      P[(size_t)c * LD + r] = acc;
    }
  }
#undef PLL
#undef Tmp
#undef PhiTot
  if (tid == 0) {
    ic[I_ID] = ic[I_NEXT_ID]; ic[I_NEXT_ID] += 1;     // :505
    core[C_DT] = s_dt;                                 // :508
    ic[I_CONSUMED] = s_used;                           // :510-512 (the host erases)
  }
}

static size_t be_propagate_smem(int L) {       // PhiC[CHK] + PLL + Tmp + PhiTot ([L][L+1] each) + Ddiag[L] + rec[CHK][PROP_REC]
  const int chk = (L <= 22) ? 8 : 4;
  return sizeof(double) * ((size_t)(chk + 3) * L * (L + 1) + L + (size_t)chk * PROP_REC);
}

// ====================================================================== addFeatureObservations
// One CTA per sequence, one thread per message feature.
__global__ void __launch_bounds__(512) be_add_obs_kernel(BeView v) {
  extern __shared__ unsigned long long sm_ids[];      // [T] table ids (0xfff.. = free)
  __shared__ unsigned long long hkey[2048];           // id -> slot hash (open addressing, load <= 0.5: T <= 1024)
  __shared__ short hval[2048];
  __shared__ short free_list[1024];                   // k-th free table slot in ascending slot order
  __shared__ int s_warp_tot[16], s_nfree;
  __shared__ int s_tracked, s_nbefore, s_ndis;
  __shared__ double s_dis[512];           // double like the reference's norms (larvio.cpp:842-847, 2751-2766)
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  double* core = core_of(v, s);
  const int T = v.be.T, Wcap = v.be.Wcap;
  unsigned long long* ft_id = v.be.ft_id + (size_t)s * T;
  int* ft_flags = v.be.ft_flags + (size_t)s * T;
  unsigned long long* ft_mask = v.be.ft_mask + (size_t)s * T;
  double* ft_obs = v.be.ft_obs + (size_t)s * T * Wcap * 4;
  if (tid == 0) { s_tracked = 0; s_nbefore = 0; s_ndis = 0; s_nfree = 0; }
  for (int i = tid; i < 2048; i += blockDim.x) hkey[i] = ~0ull;
  __syncthreads();
  int used_local = 0;
  for (int i = tid; i < T; i += blockDim.x) {
    const bool used = ft_flags[i] & 1;
    const unsigned long long id = used ? ft_id[i] : ~0ull;
    sm_ids[i] = id;
    used_local += used;
    if (used) {                                        // insert (ids are unique)
      unsigned h = (unsigned)((id * 0x9E3779B97F4A7C15ull) >> 53);
      while (atomicCAS(&hkey[h], ~0ull, id) != ~0ull) h = (h + 1) & 2047;
      hval[h] = (short)i;
    }
  }
  atomicAdd(&s_nbefore, used_local);
  __syncthreads();
  // free slots in ascending order (new features take them in message order, like std::map insertion into a fresh slot)
  for (int c0 = 0; c0 < T; c0 += blockDim.x) {
    const int i = c0 + tid;
    const bool fr = i < T && sm_ids[i] == ~0ull;
    const unsigned bal = __ballot_sync(0xffffffffu, fr);
    if ((tid & 31) == 0) s_warp_tot[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = s_nfree + __popc(bal & ((1u << (tid & 31)) - 1));
    for (int w2 = 0; w2 < (tid >> 5); ++w2) off += s_warp_tot[w2];
    if (fr) free_list[off] = (short)i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += s_warp_tot[w2]; s_nfree += t; }
    __syncthreads();
  }
  const int n_msg = min(v.msg_n[s], v.be.N);
  const int n_win = ic[I_NWIN];                       // slot the new state will occupy
  if (n_win >= Wcap) { if (tid == 0) ic[I_ERR] = 2; return; }   // window full: report before anything is written past its slots
  const long long sid = ic[I_ID];
  // slot of state id-1, if it is still in the window
  int prev_slot = -1;
  for (int w = n_win - 1; w >= 0; --w) if (v.be.win_id[(size_t)s * Wcap + w] == sid - 1) { prev_slot = w; break; }
  const double dt = core[C_DT];
  LvbFeature f; int slot = -1; bool is_new = false;
  if (tid < n_msg) {
    f = v.msg[(size_t)s * v.msg_stride + tid];
    unsigned h = (unsigned)((f.id * 0x9E3779B97F4A7C15ull) >> 53);
    while (hkey[h] != ~0ull) { if (hkey[h] == f.id) { slot = hval[h]; break; } h = (h + 1) & 2047; }
    is_new = slot < 0;
  }
  __syncthreads();
  // allocate free slots for new features in message order (block-level exclusive scan by ballot)
  __shared__ int warp_new[16];
  const unsigned bal = __ballot_sync(0xffffffffu, is_new);
  if ((tid & 31) == 0) warp_new[tid >> 5] = __popc(bal);
  __syncthreads();
  int rank = __popc(bal & ((1u << (tid & 31)) - 1));
  for (int w = 0; w < (tid >> 5); ++w) rank += warp_new[w];
  if (is_new) {
    slot = rank < s_nfree ? (int)free_list[rank] : -1;   // rank-th free slot
    if (slot < 0) atomicExch(&ic[I_ERR], 1);           // feature table overflow
  }
  if (tid < n_msg && slot >= 0) {
    double* o = ft_obs + ((size_t)slot * Wcap + n_win) * 4;
    if (is_new) {
      ft_id[slot] = f.id; ft_flags[slot] = 1;            // used, not initialised
      v.be.ft_anchor[(size_t)s * T + slot] = -1;
      unsigned long long m = 1ull << n_win;
      o[0] = f.u + f.u_vel * dt; o[1] = f.v + f.v_vel * dt; o[2] = f.u_vel; o[3] = f.v_vel;
      if (!(f.u_init == -1 && f.v_init == -1) && prev_slot >= 0) {      // :824-832
        const double dt_ = win_of(v, s, prev_slot)[W_DT];
        double* o2 = ft_obs + ((size_t)slot * Wcap + prev_slot) * 4;
        o2[0] = f.u_init + f.u_init_vel * dt_; o2[1] = f.v_init + f.v_init_vel * dt_; o2[2] = f.u_init_vel; o2[3] = f.v_init_vel;
        m |= 1ull << prev_slot;
      }
      ft_mask[slot] = m;
      v.be.ft_pos[((size_t)s * T + slot) * 3 + 0] = 0; v.be.ft_pos[((size_t)s * T + slot) * 3 + 1] = 0; v.be.ft_pos[((size_t)s * T + slot) * 3 + 2] = 0;
    } else {
      const unsigned long long m = ft_mask[slot];
      o[0] = f.u + f.u_vel * dt; o[1] = f.v + f.v_vel * dt; o[2] = f.u_vel; o[3] = f.v_vel;
      ft_mask[slot] = m | (1ull << n_win);
      atomicAdd(&s_tracked, 1);
      if (v.cfg.if_ZUPT_valid && prev_slot >= 0 && (m >> prev_slot & 1)) {     // :842-847
        const double* op = ft_obs + ((size_t)slot * Wcap + prev_slot) * 4;
        const double dx = f.u - op[0], dy = f.v - op[1];
        const int k = atomicAdd(&s_ndis, 1);
        if (k < 512) s_dis[k] = sqrt(dx * dx + dy * dy);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    core[C_TRACK_RATE] = (double)s_tracked / (double)s_nbefore;      // :851-853 (NaN when the map was empty)
    // checkZUPT (:2751-2766): the 9th largest displacement
    ic[I_ZUPT] = 0;
    const int nd = min(s_ndis, 512);
    if (v.cfg.if_ZUPT_valid && nd >= 20) {
      // selection of the 9th largest by repeated max (nd <= 512, once per frame)
      double cur = 1.0e300; int taken = 0; double val = 0.0;
      while (taken < 9) {
        double best = -1.0; int cnt = 0;
        for (int i = 0; i < nd; ++i) { const double d = s_dis[i]; if (d < cur) { if (d > best) { best = d; cnt = 1; } else if (d == best) ++cnt; } }
        if (best < 0.0) break;
        taken += cnt; val = best; cur = best;
      }
      if (val < v.cfg.zupt_dis) { ic[I_ZUPT] = 1; ic[I_ZUPT_EVENTS] += 1; }
    }
  }
}

// ====================================================================== stateAugmentation
__global__ void __launch_bounds__(256) be_augment_kernel(BeView v) {
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  double* core = core_of(v, s);
  const int LD = v.be.LD;
  double* P = P_of(v, s);
  const int d = ic[I_DIM];
  const int n_win = ic[I_NWIN];
  if (n_win >= v.be.Wcap || d + 6 > v.be.Dmax) { if (tid == 0) atomicExch(&ic[I_ERR], 2); return; }
  if (tid == 0) {
    double* w = win_of(v, s, n_win);
    v.be.win_id[(size_t)s * v.be.Wcap + n_win] = ic[I_ID];
    w[W_TIME] = core[C_TIME]; w[W_DT] = core[C_DT];
    for (int i = 0; i < 4; ++i) w[W_Q + i] = core[C_Q + i];
    for (int i = 0; i < 3; ++i) { w[W_P + i] = core[C_P + i]; w[W_PFEJ + i] = core[C_FNOW_P + i]; w[W_TCI + i] = core[C_TCI + i]; }
    for (int i = 0; i < 9; ++i) w[W_RIC + i] = core[C_RIC + i];
    const M3 R_b2w = quat_to_rot(core + C_Q);
    const M3 R_b2c = m3_load(core + C_RIC);
    const M3 R_c2w = m3_mul(R_b2w, m3_t(R_b2c));       // (R_b2c * R_w2b)^T
    rot_to_quat(R_c2w, w + W_QCAM);
    st3(w + W_PCAM, ld3(core + C_P) + m3_vec(R_b2w, ld3(core + C_TCI)));
  }
  const int sel[6] = {0, 1, 2, 6, 7, 8};
  const int nf = ic[I_NF], nnui = ic[I_NNUI];
  if (nf > 0 || nnui > 0) {
    // SLAM-feature (and nuisance) columns follow the pose block: the new pose is INSERTED before them (larvio.cpp:768-793).
    // P_aug[i][j] = P[m(i)][m(j)] with m = identity / selection rows / shifted feature indices -> generic re-map.
    int* cm = v.be.cmap + (size_t)s * LD;
    const int pe = d - v.be.IDP * nf - 6 * nnui;
    for (int i = tid; i < d + 6; i += blockDim.x) cm[i] = (i < pe) ? i : (i < pe + 6 ? sel[i - pe] : i - 6);
    __syncthreads();
    if (tid == 0) { ic[I_REMAP] = 1; ic[I_NEWDIM] = d + 6; ic[I_NWIN] = n_win + 1; }
    return;
  }
  for (int j = tid; j < d; j += blockDim.x) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double val = P[(size_t)sel[i] * LD + j];
      P[(size_t)(d + i) * LD + j] = val;
      P[(size_t)j * LD + d + i] = val;
    }
  }
  if (tid < 36) {
    const int i = tid / 6, k = tid - i * 6;
    P[(size_t)(d + i) * LD + d + k] = P[(size_t)sel[i] * LD + sel[k]];
  }
  __syncthreads();
  if (tid == 0) { ic[I_DIM] = d + 6; ic[I_NWIN] = n_win + 1; }
}

}  // namespace

namespace {

// ====================================================================== warp helpers
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ int nth_set_bit(unsigned long long m, int n) {   // index of the n-th (0-based) set bit
  for (int k = 0; k < n; ++k) m &= m - 1;
  return __ffsll((long long)m) - 1;
}

// ---------------------------------------------------------------- Feature::checkMotion (feature.hpp:334-381)
__device__ bool check_motion(const BeView& v, int s, unsigned long long mask, const double* obs, int Wcap, bool if_tracked) {
  const int first = __ffsll((long long)mask) - 1;
  unsigned long long m2 = mask;
  int last = 63 - __clzll((long long)m2);
  if (if_tracked) { m2 &= ~(1ull << last); last = 63 - __clzll((long long)m2); }
  const double* wf = win_of(v, s, first);
  const double* wl = win_of(v, s, last);
  const M3 Rf = quat_to_rot(wf + W_QCAM);
  V3 dir = v3(obs[(size_t)first * 4 + 0], obs[(size_t)first * 4 + 1], 1.0);
  dir = dir * (1.0 / norm(dir));
  dir = m3_vec(Rf, dir);
  const V3 tr = ld3(wl + W_PCAM) - ld3(wf + W_PCAM);
  const double par = dot(tr, dir);
  const V3 orth = tr - dir * par;
  return norm(orth) > v.cfg.feat_trans_thr;
}

// ---------------------------------------------------------------- Feature::initializePosition[_AssignAnchor]
// (feature.hpp:383-552 / :554-721).  One warp; lanes own observations (two per lane at most).
// tri_mask: observations that take part.  Returns validity; on success writes the world position.
__device__ bool triangulate_warp(const BeView& v, int s, unsigned long long tri_mask, const double* obs,
                                 bool is_init, double* pos /*[3] in/out*/, double* spec = nullptr /*[8]: ok, pos_w, p_lastcam*/) {
  const int lane = threadIdx.x & 31;
  const int n = __popcll(tri_mask);
  const int last = 63 - __clzll((long long)tri_mask);
  const double* wl = win_of(v, s, last);
  const M3 Rl = quat_to_rot(wl + W_QCAM);
  const V3 tl = ld3(wl + W_PCAM);
  // per-lane observations
  double zx[2], zy[2]; M3 R[2]; V3 t[2]; bool have[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = lane + 32 * e;
    have[e] = k < n;
    zx[e] = zy[e] = 0; R[e] = m3_identity(); t[e] = v3(0, 0, 0);
    if (have[e]) {
      const int slot = nth_set_bit(tri_mask, k);
      const double* w = win_of(v, s, slot);
      const M3 Ri = quat_to_rot(w + W_QCAM);
      R[e] = m3_mul(m3_t(Ri), Rl);                       // pose.inverse() * T_c_w_last
      t[e] = m3_tvec(Ri, tl - ld3(w + W_PCAM));
      zx[e] = obs[(size_t)slot * 4 + 0]; zy[e] = obs[(size_t)slot * 4 + 1];
    }
  }
  // ---- initial guess
  V3 init;
  if (!is_init) {
    // generateInitialGuess(cam_poses[0], measurements[last], measurements[0]) (feature.hpp:312-332)
    M3 R0; V3 t0;
    for (int i = 0; i < 9; ++i) R0.m[i] = shfl_d(R[0].m[i], 0);
    t0 = v3(shfl_d(t[0].x, 0), shfl_d(t[0].y, 0), shfl_d(t[0].z, 0));
    const double z2x = shfl_d(zx[0], 0), z2y = shfl_d(zy[0], 0);
    const int lk = n - 1;
    const double z1x = (lk < 32) ? shfl_d(zx[0], lk & 31) : shfl_d(zx[1], lk & 31);
    const double z1y = (lk < 32) ? shfl_d(zy[0], lk & 31) : shfl_d(zy[1], lk & 31);
    const V3 m = m3_vec(R0, v3(z1x, z1y, 1.0));
    const double A0 = m.x - z2x * m.z, A1 = m.y - z2y * m.z;
    const double b0 = z2x * t0.z - t0.x, b1 = z2y * t0.z - t0.y;
    const double depth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1);
    init = v3(z1x * depth, z1y * depth, depth);
  } else {
    init = m3_tvec(Rl, ld3(pos) - tl);
  }
  double sol[3] = {init.x / init.z, init.y / init.z, 1.0 / init.z};
  auto cost_of = [&](const double* x) {
    double c = 0.0;
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (have[e]) {
        const V3 h = m3_vec(R[e], v3(x[0], x[1], 1.0)) + t[e] * x[2];
        const double dx = h.x / h.z - zx[e], dy = h.y / h.z - zy[e];
        c += dx * dx + dy * dy;
      }
    return warp_sum_d(c);
  };
  double lambda = 1e-3;
  int inner = 0, outer = 0;
  bool reduced = false;
  double delta_norm = 0.0;
  double total = cost_of(sol);
  const double huber = 0.01;
  while (true) {
    double A[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};     // A: xx xy xz yy yz zz
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (have[e]) {
        const V3 h = m3_vec(R[e], v3(sol[0], sol[1], 1.0)) + t[e] * sol[2];
        // W = [R(:,0) R(:,1) t]
        const double W0[3] = {R[e].m[0], R[e].m[1], t[e].x}, W1[3] = {R[e].m[3], R[e].m[4], t[e].y}, W2[3] = {R[e].m[6], R[e].m[7], t[e].z};
        double J0[3], J1[3];
        for (int c = 0; c < 3; ++c) {
          J0[c] = 1 / h.z * W0[c] - h.x / (h.z * h.z) * W2[c];
          J1[c] = 1 / h.z * W1[c] - h.y / (h.z * h.z) * W2[c];
        }
        const double r0 = h.x / h.z - zx[e], r1 = h.y / h.z - zy[e];
        const double en = sqrt(r0 * r0 + r1 * r1);
        const double w = en <= huber ? 1.0 : sqrt(2.0 * huber / en);
        const double w2 = (w == 1.0) ? 1.0 : w * w;
        A[0] += w2 * (J0[0] * J0[0] + J1[0] * J1[0]); A[1] += w2 * (J0[0] * J0[1] + J1[0] * J1[1]);
        A[2] += w2 * (J0[0] * J0[2] + J1[0] * J1[2]); A[3] += w2 * (J0[1] * J0[1] + J1[1] * J1[1]);
        A[4] += w2 * (J0[1] * J0[2] + J1[1] * J1[2]); A[5] += w2 * (J0[2] * J0[2] + J1[2] * J1[2]);
        b[0] += w2 * (J0[0] * r0 + J1[0] * r1); b[1] += w2 * (J0[1] * r0 + J1[1] * r1); b[2] += w2 * (J0[2] * r0 + J1[2] * r1);
      }
    for (int i = 0; i < 6; ++i) A[i] = warp_sum_d(A[i]);
    for (int i = 0; i < 3; ++i) b[i] = warp_sum_d(b[i]);
    while (true) {
      // delta = (A + lambda I)^-1 b  (3x3 symmetric, LDL^T)
      const double a00 = A[0] + lambda, a01 = A[1], a02 = A[2], a11 = A[3] + lambda, a12 = A[4], a22 = A[5] + lambda;
      const double l10 = a01 / a00, l20 = a02 / a00;
      const double d1 = a11 - l10 * a01;
      const double l21 = (a12 - l20 * a01) / d1;
      const double d2 = a22 - l20 * a02 - l21 * l21 * d1;
      const double y0 = b[0], y1 = b[1] - l10 * y0, y2 = b[2] - l20 * y0 - l21 * y1;
      const double x2 = y2 / d2;
      const double x1 = y1 / d1 - l21 * x2;
      const double x0 = y0 / a00 - l10 * x1 - l20 * x2;
      const double ns[3] = {sol[0] - x0, sol[1] - x1, sol[2] - x2};
      delta_norm = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
      const double nc = cost_of(ns);
      if (nc < total) {
        reduced = true; sol[0] = ns[0]; sol[1] = ns[1]; sol[2] = ns[2]; total = nc;
        lambda = lambda / 10 > 1e-10 ? lambda / 10 : 1e-10;
      } else {
        reduced = false;
        lambda = lambda * 10 < 1e12 ? lambda * 10 : 1e12;
      }
      const bool cont = (inner < 10) && !reduced;
      ++inner;
      if (!cont) break;
    }
    inner = 0;
    const bool cont = (outer < 10) && (delta_norm > 5e-7);
    ++outer;
    if (!cont) break;
  }
  const V3 fin = v3(sol[0] / sol[2], sol[1] / sol[2], 1.0 / sol[2]);
  int bad = 0;
#pragma unroll
  for (int e = 0; e < 2; ++e)
    if (have[e]) { const V3 pp = m3_vec(R[e], fin) + t[e]; if (pp.z <= 0) bad = 1; }
  bad = __any_sync(0xffffffffu, bad);
  bool valid = !bad;
  if (total / (2.0 * n * n) > 4.7673e-04) valid = false;
  if (!(total == total)) valid = false;   // NaN guard
  if (lane == 0) {
    if (spec) {
      spec[0] = valid ? 1.0 : 0.0;
      if (valid) { st3(spec + 1, m3_vec(Rl, fin) + tl); st3(spec + 4, fin); }
    } else if (valid) st3(pos, m3_vec(Rl, fin) + tl);
  }
  return valid;
}

// ====================================================================== classification
// mode 0: removeLostFeatures (:1926-2005); mode 1: pruneImuStateBuffer selection (:2331-2489).
// one warp per table slot.  ft_action: 0 keep, 1 erase, 2 use in the update (mode 0: then erase).
__global__ void __launch_bounds__(128) be_classify_kernel(BeView v, int mode) {
  const int s = blockIdx.y;
  const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  int* ic = icore_of(v, s);
  const int T = v.be.T, Wcap = v.be.Wcap;
  if (slot >= T) return;
  const size_t fi = (size_t)s * T + slot;
  if (lane == 0) { v.be.ft_action[fi] = 0; v.be.ft_nrows[fi] = 0; v.be.ft_accept[fi] = 0; }
  if (!ic[I_OK]) return;
  if (mode == 1 && !ic[I_DO_PRUNE]) return;
  const int flags = v.be.ft_flags[fi];
  if (!(flags & 1)) return;
  const unsigned long long mask = v.be.ft_mask[fi];
  const double* obs = v.be.ft_obs + fi * Wcap * 4;
  double* pos = v.be.ft_pos + fi * 3;
  const int cur = ic[I_NWIN] - 1;
  const bool tracked = (mask >> cur) & 1;
  const int nobs = __popcll(mask);
  bool is_init = (flags & 2) != 0;
  const bool zupt = ic[I_ZUPT] != 0;
  int action = 0;
  unsigned long long usemask = 0;
  int nrows_override = -1;
  if (flags & 4) {                                            // EKF-SLAM feature in the state
    if (mode == 0 && tracked) { action = 4; usemask = 1ull << cur; nrows_override = 2; }
  } else if (mode == 0 && v.cfg.hybrid && tracked && nobs >= v.cfg.max_track_len) {
    // promotion candidate: the sequential rule (be_slam_decide_kernel) needs a two-view-initialised triangulation
    // without the current camera (initializeInvParamPosition / initializePosition), prepared here speculatively
    double* sp = v.be.ft_spec + fi * 8;
    if (lane == 0) sp[0] = 0.0;
    __syncwarp();
    if (!(flags & 8) && check_motion(v, s, mask, obs, Wcap, tracked))
      triangulate_warp(v, s, mask & ~(1ull << cur), obs, false, pos, sp);
    action = 5;
  } else if (mode == 0) {
    const unsigned long long tri = mask & ~(1ull << cur);     // initializePosition skips the current camera
    if (!tracked) {
      if (nobs < v.cfg.least_obs) action = 1;
      else {
        if (!is_init) {
          if (!check_motion(v, s, mask, obs, Wcap, tracked)) action = 1;
          else if (!triangulate_warp(v, s, tri, obs, false, pos)) action = 1;
          else is_init = true;
        }
        if (action == 0) { action = 2; usemask = mask; }
      }
    } else if (nobs >= v.cfg.max_track_len) {
      if (!is_init) {
        if (check_motion(v, s, mask, obs, Wcap, tracked))
          if (triangulate_warp(v, s, tri, obs, false, pos)) is_init = true;
      }
      if (is_init) { action = 2; usemask = mask; }
    }
    if (zupt && action == 2) { is_init = false; usemask = 0; }   // :2241-2246: no update, features dropped
  } else {
    unsigned long long rm = 0;
    for (int k = 0; k < ic[I_NRM]; ++k) rm |= 1ull << ic[I_RM0 + k];
    const unsigned long long involved = mask & rm;
    if (involved && !zupt && !(flags & 8) && __popcll(involved) > 1) {     // potential EKF features are not used (:2465)
      bool ok = true;
      if (!is_init) {
        if (!check_motion(v, s, mask, obs, Wcap, tracked)) ok = false;
        else if (!triangulate_warp(v, s, mask, obs, false, pos)) ok = false;     // _AssignAnchor: all observations
        else is_init = true;
      }
      if (ok) { action = 2; usemask = involved; }
    }
  }
  if (lane == 0) {
    v.be.ft_flags[fi] = (flags & ~2) | (is_init ? 2 : 0);
    v.be.ft_action[fi] = action;
    v.be.ft_usemask[fi] = usemask;
    v.be.ft_nrows[fi] = nrows_override >= 0 ? nrows_override : ((action == 2 && usemask) ? 2 * __popcll(usemask) : 0);
  }
}

// exclusive scan of raw row counts over the table (one CTA per sequence, 512 threads, T <= 1024)
__global__ void __launch_bounds__(512) be_scan_rows_kernel(BeView v, int which /*0: raw rows -> ft_rowofs, I_RAWROWS*/) {
  __shared__ int part[512];
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  const int T = v.be.T;
  const int per = (T + 511) / 512;
  int loc = 0;
  for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < T) loc += v.be.ft_nrows[(size_t)s * T + i]; }
  part[tid] = loc;
  __syncthreads();
  for (int o = 1; o < 512; o <<= 1) { int x = tid >= o ? part[tid - o] : 0; __syncthreads(); part[tid] += x; __syncthreads(); }
  int run = part[tid] - loc;
  for (int k = 0; k < per; ++k) {
    const int i = tid * per + k;
    if (i < T) {
      const int nr = v.be.ft_nrows[(size_t)s * T + i];
      v.be.ft_rowofs[(size_t)s * T + i] = run;
      if (nr && run + nr > v.be.RAWMAX) {        // capacity: the feature is not used this frame (reported)
        v.be.ft_nrows[(size_t)s * T + i] = 0; v.be.ft_usemask[(size_t)s * T + i] = 0; atomicExch(&ic[I_ERR], 3);
      }
      run += nr;
    }
  }
  if (tid == 511) ic[I_RAWROWS] = part[511];
}

}  // namespace

namespace {

// ---------------------------------------------------------------- measurementJacobian_ekf_1didp (:1117-1244) / _3didp (:984-1114)
// Observation of an EKF-SLAM feature (inverse depth, or (x/z, y/z, 1/z), in its anchor camera) from window slot ws.
// hf[a][0..IDP-1] are the feature's own columns.
struct Jac1d { double hf[2][3], ha[2][6], hx[2][6], he[2][6], r[2]; };
__device__ void meas_jac_idp(const BeView& v, int s, size_t fi, int ws, bool fej, Jac1d& J) {
  const int Wcap = v.be.Wcap;
  const double* wk = win_of(v, s, ws);
  const int as = v.be.ft_anchor[fi];
  const double* wa = anchor_rec(v, s, as);
  const bool nui = as >= BE_NUI_BASE;              // a nuisance anchor is used as frozen: its own camera pose, no first estimates (:1032-1046)
  const bool fej_a = fej && !nui;
  const M3 R_b2c = m3_load(wk + W_RIC);
  const V3 t_c_b = ld3(wk + W_TCI);
  const V3 f_an = v3(v.be.ft_oa[fi * 2], v.be.ft_oa[fi * 2 + 1], 1.0);
  const double inv = v.be.ft_inv[fi];
  const M3 R_bk2w = quat_to_rot(wk + W_Q), R_w2bk = m3_t(R_bk2w);
  const M3 R_w2ck = m3_mul(R_b2c, R_w2bk);
  const V3 t_bk_w = ld3(wk + W_P);
  const V3 t_ck_w = t_bk_w + m3_vec(R_bk2w, t_c_b);
  const M3 R_ba2w = quat_to_rot(wa + W_Q), R_w2ba = m3_t(R_ba2w);
  const V3 t_ba_w = ld3(wa + W_P);
  const M3 R_w2ca = nui ? m3_t(quat_to_rot(wa + W_QCAM)) : m3_mul(R_b2c, R_w2ba);
  const V3 p_w = ld3(v.be.ft_pos + fi * 3), p_fej = ld3(v.be.ft_pfej + fi * 3);
  V3 p_ca;
  if (fej_a) p_ca = m3_vec(R_b2c, m3_vec(R_w2ba, p_fej - ld3(wa + W_PFEJ)) - t_c_b);
  else p_ca = v3(f_an.x / inv, f_an.y / inv, 1.0 / inv);
  const double* z = v.be.ft_obs + (fi * Wcap + ws) * 4;
  const V3 p_ck = m3_vec(R_w2ck, p_w - t_ck_w);
  J.r[0] = z[0] - p_ck.x / p_ck.z; J.r[1] = z[1] - p_ck.y / p_ck.z;
  const double Jk[2][3] = {{1 / p_ck.z, 0, -p_ck.x / (p_ck.z * p_ck.z)}, {0, 1 / p_ck.z, -p_ck.y / (p_ck.z * p_ck.z)}};
  const V3 J_d = m3_vec(R_w2ck, m3_tvec(R_w2ca, f_an));              // R_w2ck * R_w2ca^T * f_an
  const V3 p_baf_w = fej_a ? (p_fej - ld3(wa + W_PFEJ)) : (p_w - t_ba_w);
  const V3 p_bkf_w = fej ? (p_fej - ld3(wk + W_PFEJ)) : (p_w - t_bk_w);
  const M3 Jxa_l = m3_scale(m3_mul(R_w2ck, skew(p_baf_w)), -1.0);    // J_xa = [-R skew(p_baf) | R]
  const M3 Jxk_l = m3_mul(R_w2ck, skew(p_bkf_w));                   // J_xk = [R skew(p_bkf) | -R]
  const M3 Rka = m3_mul(R_w2bk, R_ba2w);                            // R_w2bk * R_w2ba^T
  const M3 Sk = skew(m3_vec(R_w2bk, p_bkf_w) - t_c_b);
  const M3 Mx = m3_mul(Rka, skew(m3_tvec(R_b2c, p_ca)));
  const M3 Je_l = m3_mul(R_b2c, m3_sub(Sk, Mx));
  const M3 Je_r = m3_mul(R_b2c, m3_sub(Rka, m3_identity()));
  if (v.be.IDP == 3) {
    if (ws == as) {                                                   // the anchor's own observation (:1065-1073)
      for (int a = 0; a < 2; ++a) {
        for (int c = 0; c < 3; ++c) J.hf[a][c] = (a == c) ? 1.0 : 0.0;
        for (int c = 0; c < 6; ++c) { J.ha[a][c] = 0.0; J.hx[a][c] = 0.0; J.he[a][c] = 0.0; }
      }
      return;
    }
    // H_f = J_k * (R_w2ck R_w2ca^T) * J_f,  J_f = d(p_ca)/d(invParam) = [I | -f0/f2, -f1/f2, -1/f2] / f2
    const M3 Jp = m3_mul(R_w2ck, m3_t(R_w2ca));
    const double f0 = f_an.x, f1 = f_an.y, f2 = inv;
    double Jf[3][3] = {{1.0, 0.0, -f0 / f2}, {0.0, 1.0, -f1 / f2}, {0.0, 0.0, -1.0 / f2}};
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) Jf[i][c] = Jf[i][c] / f2;
    for (int a = 0; a < 2; ++a) {
      double kp[3];
      for (int c = 0; c < 3; ++c) kp[c] = Jk[a][0] * Jp.m[c] + Jk[a][1] * Jp.m[3 + c] + Jk[a][2] * Jp.m[6 + c];
      for (int c = 0; c < 3; ++c) J.hf[a][c] = kp[0] * Jf[0][c] + kp[1] * Jf[1][c] + kp[2] * Jf[2][c];
    }
  }
  const double J_rho = -1.0 / (inv * inv);
  const double jd[3] = {J_d.x, J_d.y, J_d.z};
  for (int a = 0; a < 2; ++a) {
    if (v.be.IDP == 1) { J.hf[a][0] = (Jk[a][0] * jd[0] + Jk[a][1] * jd[1] + Jk[a][2] * jd[2]) * J_rho; J.hf[a][1] = 0.0; J.hf[a][2] = 0.0; }
    for (int c = 0; c < 3; ++c) {
      double xa = 0, xr = 0, kl = 0, kr = 0, el = 0, er = 0;
      for (int q = 0; q < 3; ++q) {
        xa += Jk[a][q] * Jxa_l.m[q * 3 + c]; xr += Jk[a][q] * R_w2ck.m[q * 3 + c];
        kl += Jk[a][q] * Jxk_l.m[q * 3 + c]; kr += Jk[a][q] * -R_w2ck.m[q * 3 + c];
        el += Jk[a][q] * Je_l.m[q * 3 + c]; er += Jk[a][q] * Je_r.m[q * 3 + c];
      }
      J.ha[a][c] = xa; J.ha[a][3 + c] = xr; J.hx[a][c] = kl; J.hx[a][3 + c] = kr; J.he[a][c] = el; J.he[a][3 + c] = er;
    }
  }
}

// ====================================================================== per-feature Jacobian + gating
// One warp (one CTA) per table slot.  action 2: MSCKF feature (featureJacobian_msckf + gatingTest); action 3: new EKF-SLAM
// feature (same gate, then featureJacobian_ekf_new and the split of its block into null-space rows and the row that
// defines the new state, larvio.cpp:2033-2125); action 4: EKF-SLAM feature already in the state (featureJacobian_ekf +
// gate with dof 2).  Raw rows live in Hraw[s][rowofs ..) (dense, width LD).  For MSCKF blocks three Householder
// reflections annihilate H_f; rows 3.. are A^T H_x, A^T r (any orthonormal basis of the left null space gives the same
// gate value and the same update).  dynamic smem: see be_feature_smem_bytes().
__device__ __forceinline__ int be_feature_smem_doubles(int Wcap) { return 6 * Wcap + 8 + 4 * Wcap * Wcap + 2 * Wcap + (8 + 6 * Wcap) / 2 + 4; }

// gamma = r^T (H P H^T + sigma^2 I)^-1 r over the listed nonzero columns; returns pass/fail against chi2[dof = R]
__device__ bool gate_block(const BeView& v, const double* P, int LD, const double* Hj, double* Tp, const double* rj, int R,
                           const int* nzl, int nz, double* Ssm, double* vv, int lane, double* gamma_out = nullptr) {
  if (R <= 0 || R >= 100) return false;
  for (int j = lane; j < nz; j += 32) {
    const int c2 = nzl[j];
    for (int a0 = 0; a0 < R; a0 += 16) {
      double acc[16];
#pragma unroll
      for (int a = 0; a < 16; ++a) acc[a] = 0.0;
      for (int q = 0; q < nz; ++q) {
        const int c1 = nzl[q];
        const double pv = P[(size_t)c1 * LD + c2];
#pragma unroll
        for (int a = 0; a < 16; ++a)
          if (a0 + a < R) acc[a] += Hj[(size_t)(a0 + a) * LD + c1] * pv;
      }
#pragma unroll
      for (int a = 0; a < 16; ++a)
        if (a0 + a < R) Tp[(size_t)(a0 + a) * LD + c2] = acc[a];
    }
  }
  __syncwarp();
  for (int e = lane; e < R * R; e += 32) {
    const int a = e / R, b = e - a * R;
    double acc = (a == b) ? v.cfg.sfeat2 : 0.0;
    for (int q = 0; q < nz; ++q) { const int c1 = nzl[q]; acc += Tp[(size_t)a * LD + c1] * Hj[(size_t)b * LD + c1]; }
    Ssm[e] = acc;
  }
  __syncwarp();
  for (int j = 0; j < R; ++j) {
    double djj = Ssm[j * R + j];
    if (!(djj > 0.0)) return false;
    djj = sqrt(djj);
    __syncwarp();
    if (lane == 0) Ssm[j * R + j] = djj;
    for (int i = j + 1 + lane; i < R; i += 32) Ssm[i * R + j] /= djj;
    __syncwarp();
    for (int e = lane; e < (R - j - 1) * (R - j - 1); e += 32) {
      const int a = j + 1 + e / (R - j - 1), b = j + 1 + e % (R - j - 1);
      if (b <= a) Ssm[a * R + b] -= Ssm[a * R + j] * Ssm[b * R + j];
    }
    __syncwarp();
  }
  if (lane == 0) {
    double gamma = 0.0;
    for (int i = 0; i < R; ++i) {
      double x = rj[i];
      for (int q = 0; q < i; ++q) x -= Ssm[i * R + q] * vv[q];
      x /= Ssm[i * R + i];
      vv[i] = x;
      gamma += x * x;
    }
    vv[0] = gamma;
  }
  __syncwarp();
  const double gamma = vv[0];
  __syncwarp();
  if (gamma_out && lane == 0) *gamma_out = gamma;
  return gamma < c_chi2[R];
}

// ncf Householder reflections that triangularise Hf [nrows][ncf] (a feature's own columns), applied to the listed nonzero
// columns of the block H and to its residual: afterwards rows ncf.. of (H, rr) are the projection onto the left null space of
// Hf and rows 0..ncf-1 carry its column space, with the triangular factor (LAPACK sign convention) left in Hf.
__device__ void reflect_block(double* Hf, int ncf, double* H, double* rr, int nrows, int LD, const int* nzl, int nz, double* vv, int lane) {
  for (int k = 0; k < ncf; ++k) {
    double nrm = 0.0;
    for (int i = k + lane; i < nrows; i += 32) { const double x = Hf[i * ncf + k]; nrm += x * x; }
    nrm = sqrt(warp_sum_d(nrm));
    const double x0 = Hf[k * ncf + k];
    const double alpha = x0 >= 0 ? -nrm : nrm;
    for (int i = k + lane; i < nrows; i += 32) vv[i] = Hf[i * ncf + k] - (i == k ? alpha : 0.0);
    __syncwarp();
    double vtv = 0.0;
    for (int i = k + lane; i < nrows; i += 32) vtv += vv[i] * vv[i];
    vtv = warp_sum_d(vtv);
    if (vtv > 0.0) {
      const double beta = 2.0 / vtv;
      for (int c = k; c < ncf; ++c) {
        double dt_ = 0.0;
        for (int i = k + lane; i < nrows; i += 32) dt_ += vv[i] * Hf[i * ncf + c];
        dt_ = warp_sum_d(dt_) * beta;
        for (int i = k + lane; i < nrows; i += 32) Hf[i * ncf + c] -= dt_ * vv[i];
      }
      for (int j = lane; j < nz + 1; j += 32) {
        if (j < nz) {
          const int c = nzl[j];
          double dt_ = 0.0;
          for (int i = k; i < nrows; ++i) dt_ += vv[i] * H[(size_t)i * LD + c];
          dt_ *= beta;
          for (int i = k; i < nrows; ++i) H[(size_t)i * LD + c] -= dt_ * vv[i];
        } else {
          double dt_ = 0.0;
          for (int i = k; i < nrows; ++i) dt_ += vv[i] * rr[i];
          dt_ *= beta;
          for (int i = k; i < nrows; ++i) rr[i] -= dt_ * vv[i];
        }
      }
      __syncwarp();
      if (lane == 0) Hf[k * ncf + k] = alpha;           // the exact diagonal of the triangular factor
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) be_feature_kernel(BeView v, double* Traw) {
  extern __shared__ double fsm[];
  const int s = blockIdx.y, slot = blockIdx.x, lane = threadIdx.x;
  const int T = v.be.T, Wcap = v.be.Wcap, LD = v.be.LD;
  const size_t fi = (size_t)s * T + slot;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int nrows_all = v.be.ft_nrows[fi];
  const int action = v.be.ft_action[fi];
  if (nrows_all == 0 || action < 2 || action > 4) return;
  const unsigned long long um = v.be.ft_usemask[fi];
  const int ofs = v.be.ft_rowofs[fi];
  double* H = v.be.Hraw + ((size_t)s * v.be.RAWMAX + ofs) * LD;
  double* rr = v.be.rraw + (size_t)s * v.be.RAWMAX + ofs;
  double* Tj = Traw + ((size_t)s * v.be.RAWMAX + ofs) * LD;
  const double* P = P_of(v, s);
  const double* obs = v.be.ft_obs + fi * Wcap * 4;
  const V3 p_w = ld3(v.be.ft_pos + fi * 3);
  const bool fej = ic[I_FEJ] != 0;
  const int n_win = ic[I_NWIN];
  double* Hf = fsm;                                        // [2m][3]
  double* Ssm = fsm + 6 * Wcap + 8;                        // [(2m-3)^2]
  double* vv = Ssm + 4 * Wcap * Wcap;                      // [2m] reflector / rhs scratch
  int* nzl = reinterpret_cast<int*>(vv + 2 * Wcap);        // nonzero column list
  for (int i = lane; i < nrows_all * LD; i += 32) H[i] = 0.0;
  __syncwarp();

  if (action == 4) {
    // ---- featureJacobian_ekf (:1341-1417): 2 rows, gate with dof 2 (:2133-2143)
    const int cur = n_win - 1;
    const int as = v.be.ft_anchor[fi];
    const int idp = v.be.IDP;
    int fidx = -1;
    const int* fs = v.be.fs_slot + (size_t)s * 64;
    for (int i = 0; i < ic[I_NF]; ++i) if (fs[i] == slot) fidx = LEGD + 6 * n_win + idp * i;
    int nz4 = 0;
    if (lane == 0 && fidx >= 0) {
      Jac1d J;
      meas_jac_idp(v, s, fi, cur, fej, J);
      const int acol = anchor_col(v, as, n_win, ic[I_NF]);
      for (int a = 0; a < 2; ++a) {
        double* row = H + (size_t)a * LD;
        for (int c = 0; c < 6; ++c) { row[acol + c] = J.ha[a][c]; row[15 + c] = J.he[a][c]; }
        for (int c = 0; c < 6; ++c) row[LEGD + 6 * cur + c] = J.hx[a][c];
        for (int q = 0; q < idp; ++q) row[fidx + q] = J.hf[a][q];
        if (v.cfg.estimate_td) row[21] = obs[(size_t)cur * 4 + 2 + a];
        rr[a] = J.r[a];
      }
      for (int j = 0; j < 7; ++j) nzl[nz4++] = 15 + j;
      for (int c = 0; c < 6; ++c) nzl[nz4++] = acol + c;
      if (cur != as) for (int c = 0; c < 6; ++c) nzl[nz4++] = LEGD + 6 * cur + c;
      for (int q = 0; q < idp; ++q) nzl[nz4++] = fidx + q;
    }
    nz4 = __shfl_sync(0xffffffffu, nz4, 0);
    __syncwarp();
    const bool pass = fidx >= 0 && gate_block(v, P, LD, H, Tj, rr, 2, nzl, nz4, Ssm, vv, lane);
    if (lane == 0) v.be.ft_accept[fi] = pass ? 2 : 0;
    return;
  }

  // ---- measurementJacobian_msckf per observation (:859-921) into rows [0, 2m)
  const int m = __popcll(um);
  const int nrows = 2 * m;
  for (int k = lane; k < m; k += 32) {
    const int ws = nth_set_bit(um, k);
    const double* w = win_of(v, s, ws);
    const M3 R_b2c = m3_load(w + W_RIC);
    const V3 t_c_b = ld3(w + W_TCI);
    const M3 R_b2w = quat_to_rot(w + W_Q);
    const M3 R_w2c = m3_mul(R_b2c, m3_t(R_b2w));
    const V3 t_b_w = ld3(w + W_P);
    const V3 t_c_w = t_b_w + m3_vec(R_b2w, t_c_b);
    const V3 p_c = m3_vec(R_w2c, p_w - t_c_w);
    const V3 p_bf_w = fej ? (p_w - ld3(w + W_PFEJ)) : (p_w - t_b_w);
    double dz[2][3] = {{1 / p_c.z, 0, -p_c.x / (p_c.z * p_c.z)}, {0, 1 / p_c.z, -p_c.y / (p_c.z * p_c.z)}};
    const M3 A1 = m3_mul(R_w2c, skew(p_bf_w));                       // dpc_dxb.leftCols(3)
    const M3 E1 = m3_sub(m3_mul(A1, R_b2w), m3_mul(R_b2c, skew(t_c_b)));   // dpc_dxe.leftCols(3)
    double* r0 = H + (size_t)(2 * k) * LD;
    double* r1 = r0 + LD;
    const int cp = LEGD + 6 * ws;
    for (int c = 0; c < 3; ++c) {
      double hx0 = 0, hx1 = 0, hp0 = 0, hp1 = 0, he0 = 0, he1 = 0, ht0 = 0, ht1 = 0, hf0 = 0, hf1 = 0;
      for (int q = 0; q < 3; ++q) {
        hx0 += dz[0][q] * A1.m[q * 3 + c]; hx1 += dz[1][q] * A1.m[q * 3 + c];
        hp0 += dz[0][q] * -R_w2c.m[q * 3 + c]; hp1 += dz[1][q] * -R_w2c.m[q * 3 + c];
        he0 += dz[0][q] * E1.m[q * 3 + c]; he1 += dz[1][q] * E1.m[q * 3 + c];
        ht0 += dz[0][q] * -R_b2c.m[q * 3 + c]; ht1 += dz[1][q] * -R_b2c.m[q * 3 + c];
        hf0 += dz[0][q] * R_w2c.m[q * 3 + c]; hf1 += dz[1][q] * R_w2c.m[q * 3 + c];
      }
      r0[cp + c] = hx0; r1[cp + c] = hx1; r0[cp + 3 + c] = hp0; r1[cp + 3 + c] = hp1;
      r0[15 + c] = he0; r1[15 + c] = he1; r0[18 + c] = ht0; r1[18 + c] = ht1;
      Hf[(2 * k) * 3 + c] = hf0; Hf[(2 * k + 1) * 3 + c] = hf1;
    }
    if (v.cfg.estimate_td) { r0[21] = obs[(size_t)ws * 4 + 2]; r1[21] = obs[(size_t)ws * 4 + 3]; }
    rr[2 * k] = obs[(size_t)ws * 4 + 0] - p_c.x / p_c.z;
    rr[2 * k + 1] = obs[(size_t)ws * 4 + 1] - p_c.y / p_c.z;
  }
  // nonzero column list: 15..21 and the pose blocks of the used window slots
  const int nz = 7 + 6 * m;
  for (int j = lane; j < nz; j += 32) nzl[j] = (j < 7) ? 15 + j : LEGD + 6 * nth_set_bit(um, (j - 7) / 6) + ((j - 7) % 6);
  __syncwarp();
  // ---- three Householder reflections on H_f, applied to the nonzero columns of H and to r
  reflect_block(Hf, 3, H, rr, nrows, LD, nzl, nz, vv, lane);
  // ---- gating test (:1865-1880) on rows 3..nrows-1
  const int R = nrows - 3;
  const bool pass = gate_block(v, P, LD, H + (size_t)3 * LD, Tj + (size_t)3 * LD, rr + 3, R, nzl, nz, Ssm, vv, lane, v.be.ft_gamma + fi);
  if (action == 2) { if (lane == 0) v.be.ft_accept[fi] = pass ? R : 0; return; }

  // ---- action 3: new EKF-SLAM feature
  if (!pass) {                                   // :2053-2059, :2084-2092: stays in the map, not in the state
    if (lane == 0) { v.be.ft_accept[fi] = 0; v.be.ft_action[fi] = 0; }
    return;
  }
  // featureJacobian_ekf_new (:1247-1338): every observation (1-D inverse depth: except the anchor's, :1260-1262), rows from 2m
  const int idp = v.be.IDP;
  const int as = v.be.ft_anchor[fi];
  const unsigned long long um2 = (idp == 1) ? (um & ~(1ull << as)) : um;
  const int m2 = __popcll(um2), nr2 = 2 * m2;
  double* H2 = H + (size_t)nrows * LD;
  double* r2 = rr + nrows;
  double* hf = Hf;                               // own columns [2 m2][idp]
  for (int k = lane; k < m2; k += 32) {
    const int ws = nth_set_bit(um2, k);
    Jac1d J;
    meas_jac_idp(v, s, fi, ws, fej, J);
    for (int a = 0; a < 2; ++a) {
      double* row = H2 + (size_t)(2 * k + a) * LD;
      for (int c = 0; c < 6; ++c) { row[LEGD + 6 * as + c] = J.ha[a][c]; row[15 + c] = J.he[a][c]; }
      for (int c = 0; c < 6; ++c) row[LEGD + 6 * ws + c] = J.hx[a][c];
      if (v.cfg.estimate_td) row[21] = obs[(size_t)ws * 4 + 2 + a];
      for (int q = 0; q < idp; ++q) hf[(2 * k + a) * idp + q] = J.hf[a][q];
      r2[2 * k + a] = J.r[a];
    }
  }
  __syncwarp();
  if (nr2 <= idp) {                              // nothing left to define the state with: treated like a failed gate
    if (lane == 0) { v.be.ft_accept[fi] = 0; v.be.ft_action[fi] = 0; }
    return;
  }
  // Householder reflections on the own columns: rows 0..idp-1 <- column space (H_1, H_2 triangular, r_1), the rest <- left
  // null space (:2094-2125: any orthonormal basis of either space gives the same update up to row signs, which cancel)
  reflect_block(hf, idp, H2, r2, nr2, LD, nzl, nz, vv, lane);
  // (H_1 | H_2 | r_1) rows of this feature -> Hnew[s][idp * rank + j], rank = position among this frame's candidates
  const int* cand = v.be.cand + (size_t)s * 128;
  int rank = -1;
  for (int k = 0; k < ic[I_NCAND]; ++k) if (cand[k] == slot) rank = k;
  if (rank >= 0) {
    for (int j = 0; j < idp; ++j) {
      double* hn = v.be.Hnew + ((size_t)s * 64 * idp + (size_t)rank * idp + j) * (LD + 4);
      for (int c = lane; c < LD; c += 32) hn[c] = H2[(size_t)j * LD + c];
      if (lane == 0) {
        for (int q = 0; q < 3; ++q) hn[LD + q] = (q < idp && q >= j) ? hf[j * idp + q] : 0.0;
        hn[LD + 3] = r2[j];
      }
    }
  }
  if (lane == 0) v.be.ft_accept[fi] = (rank >= 0) ? nr2 - idp : 0;
}

// ====================================================================== stacking (column-major Hs)
// phase 0: gated MSCKF blocks (action 2) from row 0 -> I_ROWS = I_R = total (then be_qr_kernel compresses them);
// phase 1: gated in-state SLAM rows (action 4) and the null-space rows of new SLAM features (action 3) appended after
//          the (possibly compressed) MSCKF block, exactly like H_o in measurementUpdate_hybrid (:1616-1628).
__global__ void __launch_bounds__(512) be_stack_kernel(BeView v, int phase) {
  __shared__ int part[512];
  __shared__ int s_total;
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int T = v.be.T, LD = v.be.LD, RMAX = v.be.RMAX;
  const int per = (T + 511) / 512;
  auto rows_of = [&](int i) {
    const size_t fi = (size_t)s * T + i;
    const int act = v.be.ft_action[fi];
    const bool mine = (phase == 0) ? (act == 2) : (act == 3 || act == 4);
    return mine ? v.be.ft_accept[fi] : 0;
  };
  int loc = 0;
  for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < T) loc += rows_of(i); }
  part[tid] = loc;
  __syncthreads();
  for (int o = 1; o < 512; o <<= 1) { int x = tid >= o ? part[tid - o] : 0; __syncthreads(); part[tid] += x; __syncthreads(); }
  if (tid == 511) { s_total = part[511]; }
  __syncthreads();
  const int base = (phase == 0) ? 0 : ic[I_R];
  int total = s_total;
  const int cap = (phase == 0) ? RMAX : min(RMAX, v.be.LDS);   // the MSCKF block is compressed to <= cols rows before the update
  if (base + total > cap) { if (tid == 0) atomicExch(&ic[I_ERR], 4); total = max(cap - base, 0); }
  int run = base + part[tid] - loc;
  for (int k = 0; k < per; ++k) {
    const int i = tid * per + k;
    if (i < T) { const int a = rows_of(i); if (a) v.be.ft_nrows[(size_t)s * T + i] = run; run += a; }
  }
  __syncthreads();
  const int d = ic[I_DIM];
  double* Hs = v.be.Hs + (size_t)s * LD * RMAX;
  double* rs = v.be.rs + (size_t)s * RMAX;
  const int warp = tid >> 5, lane = tid & 31;
  if (phase == 0) {
    // MSCKF blocks are structurally sparse (columns 15..21 + the pose blocks of the observing window slots): clear the rows
    // (coalesced along the column-major stack), scatter only each block's own columns, and derive the list of nonzero
    // columns (kmap, I_NC) from the union of the slot masks - a superset of the numerically nonzero columns, which is all
    // the compression and the update need (be_colscan_kernel does the same from the data where rows of other kinds are stacked).
    __shared__ unsigned long long s_um;
    if (tid == 0) s_um = 0ull;
    const int nclr = min(total, cap);
    for (int c = warp; c < d; c += 16) { double* cj = Hs + (size_t)c * RMAX; for (int r = lane; r < nclr; r += 32) cj[r] = 0.0; }
    __syncthreads();
    unsigned long long um_loc = 0ull;
    for (int i = warp; i < T; i += 16) {
      const int a = rows_of(i);
      if (!a) continue;
      const size_t fi = (size_t)s * T + i;
      const int dst = v.be.ft_nrows[fi];
      if (dst + a > cap) continue;
      const unsigned long long um = v.be.ft_usemask[fi];
      um_loc |= um;
      const int nz = 7 + 6 * __popcll(um);
      const int first = v.be.ft_rowofs[fi] + 3;                 // skip the 3 rows that carry H_f
      const double* H = v.be.Hraw + ((size_t)s * v.be.RAWMAX + first) * LD;
      const double* rr = v.be.rraw + (size_t)s * v.be.RAWMAX + first;
      for (int e = lane; e < a * nz; e += 32) {
        const int row = e / nz, j = e - row * nz;
        const int col = (j < 7) ? 15 + j : LEGD + 6 * nth_set_bit(um, (j - 7) / 6) + ((j - 7) % 6);
        Hs[(size_t)col * RMAX + dst + row] = H[(size_t)row * LD + col];
      }
      for (int e = lane; e < a; e += 32) rs[dst + e] = rr[e];
    }
    if (lane == 0 && um_loc) atomicOr(&s_um, um_loc);
    __syncthreads();
    if (tid == 0) {
      int* km = v.be.kmap + (size_t)s * LD;
      int nc = 0;
      const unsigned long long um = s_um;
      if (total > 0) {
        for (int j = 0; j < 7; ++j) km[nc++] = 15 + j;
        for (int w = 0; w < 64; ++w) if ((um >> w) & 1ull) for (int c = 0; c < 6; ++c) km[nc++] = LEGD + 6 * w + c;
      }
      ic[I_NC] = nc;
    }
  } else {
  for (int i = warp; i < T; i += 16) {
    const int a = rows_of(i);
    if (!a) continue;
    const size_t fi = (size_t)s * T + i;
    const int dst = v.be.ft_nrows[fi];
    if (dst + a > cap) continue;
    const int act = v.be.ft_action[fi];
    // first raw row of the block that goes into H_o: MSCKF skips the 3 rows that carry H_f; a new SLAM feature keeps
    // the rows after its 2m gating rows and after the IDP rows that define the new state
    int first = v.be.ft_rowofs[fi];
    if (act == 2) first += 3;
    else if (act == 3) first += 2 * __popcll(v.be.ft_usemask[fi]) + v.be.IDP;
    const double* H = v.be.Hraw + ((size_t)s * v.be.RAWMAX + first) * LD;
    const double* rr = v.be.rraw + (size_t)s * v.be.RAWMAX + first;
    for (int e = lane; e < a * d; e += 32) {
      const int row = e / d, col = e - row * d;
      Hs[(size_t)col * RMAX + dst + row] = H[(size_t)row * LD + col];
    }
    for (int e = lane; e < a; e += 32) rs[dst + e] = rr[e];
  }
  }
  if (tid == 0) {
    if (phase == 0) { ic[I_ROWS] = total; ic[I_R] = total; }
    else ic[I_R] = base + total;
  }
}

// new SLAM features that passed the gate, in candidate (= id) order: state indices, fs_slot, compact Hnew (:2019-2093)
__global__ void be_slam_accept_kernel(BeView v) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.be.S) return;
  int* ic = icore_of(v, s);
  ic[I_NNEW] = 0;
  if (!ic[I_OK]) return;
  const int T = v.be.T, LD = v.be.LD;
  int* cand = v.be.cand + (size_t)s * 128;
  int* fs = v.be.fs_slot + (size_t)s * 64;
  const int nf = ic[I_NF];
  int k2 = 0;
  for (int k = 0; k < ic[I_NCAND]; ++k) {
    const int slot = cand[k];
    const size_t fi = (size_t)s * T + slot;
    if (v.be.ft_action[fi] != 3) continue;                 // rejected by the gate (action reset to 0)
    if (k2 != k) {
      const int idp = v.be.IDP;
      const double* src = v.be.Hnew + ((size_t)s * 64 + k) * idp * (LD + 4);
      double* dst = v.be.Hnew + ((size_t)s * 64 + k2) * idp * (LD + 4);
      for (int c = 0; c < idp * (LD + 4); ++c) dst[c] = src[c];
    }
    fs[nf + k2] = slot;
    v.be.ft_flags[fi] |= 4;                                // in_state
    ++k2;
  }
  ic[I_NNEW] = k2;
}

// second half of measurementUpdate_hybrid (:1661-1676, :1821-1854): the states of the new SLAM features and the grown
// covariance.  The reference solves with H_2.ldlt(), which reads the lower triangle of the upper-triangular factor, i.e. its
// diagonal (App. C-13): HH = H_1 / diag(H_2) row by row; only P22's noise term uses the full factor (H_2^T H_2)^-1 (:1823-1825).
// Scratch: HH [nn][LD] in Tm (free after P -= Y^T Y), nHHP [nn][LD] in Sm.
__global__ void __launch_bounds__(256) be_slam_grow_kernel(BeView v) {
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int n_new = ic[I_NNEW];
  if (n_new == 0) return;
  const int idp = v.be.IDP, nn = idp * n_new;
  const int d = ic[I_DIM], LD = v.be.LD, T = v.be.T;
  double* P = P_of(v, s);
  const double* dx = v.be.dx + (size_t)s * v.be.LDS;
  const bool have_dx = ic[I_R] > 0;
  const int* fs = v.be.fs_slot + (size_t)s * 64;
  const int nf = ic[I_NF];
  const double* Hn = v.be.Hnew + (size_t)s * 64 * idp * (LD + 4);
  double* HH = v.be.Tm + (size_t)s * v.be.RAWMAX * LD;
  for (int e = tid; e < nn * d; e += blockDim.x) {
    const int k = e / d, c = e - k * d;
    const double* hn = Hn + (size_t)k * (LD + 4);
    HH[(size_t)k * LD + c] = hn[c] / hn[LD + k % idp];
  }
  __syncthreads();
  // new rows/cols of P: nHHP = -HH P (P already updated), P22 = -nHHP HH^T + sigma^2 (H_2^T H_2)^-1
  double* Sd = v.be.Sm + (size_t)s * v.be.LDS * v.be.LDS;          // scratch for nHHP [nn][LD]
  for (int e = tid; e < nn * d; e += blockDim.x) {
    const int k = e / d, c = e - k * d;
    double acc = 0.0;
    for (int q = 0; q < d; ++q) acc += HH[(size_t)k * LD + q] * P[(size_t)q * LD + c];
    Sd[(size_t)k * LD + c] = -acc;
  }
  __syncthreads();
  for (int e = tid; e < nn * d; e += blockDim.x) {
    const int k = e / d, c = e - k * d;
    const double val = Sd[(size_t)k * LD + c];
    P[(size_t)(d + k) * LD + c] = val;
    P[(size_t)c * LD + d + k] = val;
  }
  for (int e = tid; e < nn * nn; e += blockDim.x) {
    const int a = e / nn, b = e - a * nn;
    double acc = 0.0;
    for (int q = 0; q < d; ++q) acc += Sd[(size_t)a * LD + q] * HH[(size_t)b * LD + q];
    acc = -acc;
    if (idp == 1) {
      if (a == b) { const double h2 = Hn[(size_t)a * (LD + 4) + LD]; acc += v.cfg.sfeat2 / (h2 * h2); }
    } else if (a / 3 == b / 3) {
      // (R^T R)^-1 = R^-1 R^-T of this feature's upper-triangular 3x3 factor
      const double* h0 = Hn + (size_t)(a / 3 * 3) * (LD + 4) + LD;
      const double r00 = h0[0], r01 = h0[1], r02 = h0[2], r11 = h0[(LD + 4) + 1], r12 = h0[(LD + 4) + 2], r22 = h0[2 * (LD + 4) + 2];
      double Ri[3][3] = {{1.0 / r00, -r01 / (r00 * r11), (r01 * r12 - r02 * r11) / (r00 * r11 * r22)}, {0.0, 1.0 / r11, -r12 / (r11 * r22)}, {0.0, 0.0, 1.0 / r22}};
      const int ia = a % 3, ib = b % 3;
      acc += v.cfg.sfeat2 * (Ri[ia][0] * Ri[ib][0] + Ri[ia][1] * Ri[ib][1] + Ri[ia][2] * Ri[ib][2]);
    }
    P[(size_t)(d + a) * LD + d + b] = acc;
  }
  __syncthreads();
  for (int e = tid; e < nn * nn; e += blockDim.x) {             // symmetrise the new block (:1852-1853)
    const int a = e / nn, b = e - a * nn;
    if (a < b) { const double mval = 0.5 * (P[(size_t)(d + a) * LD + d + b] + P[(size_t)(d + b) * LD + d + a]); P[(size_t)(d + a) * LD + d + b] = mval; P[(size_t)(d + b) * LD + d + a] = mval; }
  }
  // dx_new = -HH dx_leg + r_1 / diag(H_2); inverse-depth parameters and world position of the new features
  if (tid < n_new) {
    const int k = tid;
    double dxn[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < idp; ++j) {
      const double* hn = Hn + (size_t)(k * idp + j) * (LD + 4);
      double acc = 0.0;
      if (have_dx) for (int q = 0; q < d; ++q) acc += HH[(size_t)(k * idp + j) * LD + q] * dx[q];
      dxn[j] = -acc + hn[LD + 3] / hn[LD + j];
    }
    const size_t fi = (size_t)s * T + fs[nf + k];
    if (idp == 3) { v.be.ft_oa[fi * 2] += dxn[0]; v.be.ft_oa[fi * 2 + 1] += dxn[1]; }
    const double inv = v.be.ft_inv[fi] + dxn[idp - 1];
    v.be.ft_inv[fi] = inv;
    const double* wa = win_of(v, s, v.be.ft_anchor[fi]);
    const V3 p_c = v3(v.be.ft_oa[fi * 2] / inv, v.be.ft_oa[fi * 2 + 1] / inv, 1.0 / inv);
    st3(v.be.ft_pos + fi * 3, m3_vec(quat_to_rot(wa + W_QCAM), p_c) + ld3(wa + W_PCAM));
  }
  __syncthreads();
  const int nnui = ic[I_NNUI];
  if (nnui > 0) {
    // the new columns were appended behind the nuisance block (where featureJacobian_ekf_new puts them, :1291-1300); the
    // reference then moves the nuisance block to the end (:1832-1845): [.. old features | new features | nuisance]
    int* cm = v.be.cmap + (size_t)s * LD;
    const int n0 = d - 6 * nnui;
    for (int i = tid; i < d + nn; i += blockDim.x) cm[i] = (i < n0) ? i : (i < n0 + nn ? d + (i - n0) : i - nn);
    __syncthreads();
    if (tid == 0) { ic[I_REMAP] = 1; ic[I_NEWDIM] = d + nn; }
  }
  if (tid == 0) { ic[I_DIM] = d + nn; ic[I_NF] = nf + n_new; ic[I_NNEW] = 0; }
}

// Schmidt update (:1579-1589, :1805-1814, :2940-2950): the nuisance block of P keeps its prior through every update - saved
// before the update kernels, written back after P -= Y^T Y (before new feature columns are attached)
__global__ void __launch_bounds__(256) be_nui_block_kernel(BeView v, int restore) {
  const int s = blockIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int nn = 6 * ic[I_NNUI];
  if (nn == 0) return;
  const int LD = v.be.LD, W = 6 * v.be.NUI;
  const int n0 = LEGD + 6 * ic[I_NWIN] + v.be.IDP * ic[I_NF];
  double* P = P_of(v, s);
  double* B = v.be.nui_P + (size_t)s * W * W;
  for (int e = threadIdx.x; e < nn * nn; e += blockDim.x) {
    const int a = e / nn, b = e - a * nn;
    if (restore) P[(size_t)(n0 + a) * LD + n0 + b] = B[a * W + b];
    else B[a * W + b] = P[(size_t)(n0 + a) * LD + n0 + b];
  }
}

// ====================================================================== QR compression (SPQR thin QR at larvio.cpp:1430-1449, 2151-2171)
// Householder, column by column over the structurally nonzero columns (kmap), on a shared-memory panel [nc + 1 columns (H | r)]
// [rows], column-major.  Rows arrive in blocks: the panel holds as many rows as fit (<= QR_SMEM_DOUBLES), is reduced to its
// nc x nc triangle, and the next block of rows is stacked under that triangle and reduced again (a thin QR is invariant to this
// row blocking) - so the bursts of 300..650 rows that a frame with many tracks ending together produces never leave shared memory.
// Per column ONE block barrier: the warp that updates column j+1 also derives its reflector (norm below the diagonal -> alpha, v0
// written in place of the diagonal entry, beta into a parity-buffered slot), so the next iteration reads the column as the
// reflector directly; groups of 8/16/32 lanes own a trailing column each and walk it two rows at a time (LDS.128).  Systems whose
// column count leaves no room for a useful row block (hybrid configurations with hundreds of nonzero columns) take the same
// algorithm on the global column-major stack with two barriers per column.
// (Round 2, per 16-sequence launch: 5-barrier global version 66 us, 2-barrier global version 53 us.)
constexpr int QR_SMEM_DOUBLES = 26000;               // 203 KB of the 227 KB a CTA can have

__global__ void __launch_bounds__(512) be_qr_kernel(BeView v) {
  extern __shared__ double qsm[];      // panel [(c + 1)][Rp] (shared-memory path) or reflector [RMAX] (global path)
  __shared__ double red[16];
  __shared__ double s_n2, s_x0;
  __shared__ double s_beta[2];             // shared-memory path: beta of the next column's reflector, double-buffered by column parity
  __shared__ double s_diag[BE_DMAX_PAD];
  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int R = ic[I_ROWS], c = ic[I_NC];        // compression over the structurally nonzero columns only (be_stack_kernel / be_colscan_kernel)
  if (R <= c || R == 0) return;
  const int RMAX = v.be.RMAX, LD = v.be.LD;
  double* Hs = v.be.Hs + (size_t)s * LD * RMAX;
  double* rs = v.be.rs + (size_t)s * RMAX;
  const int* km = v.be.kmap + (size_t)s * LD;
  // rows the panel can hold: stride Rp = 8 (mod 16) doubles (even: 16-byte aligned pairs; lane groups of neighbouring columns
  // land in different bank halves)
  int cap_rows = QR_SMEM_DOUBLES / (c + 1);
  cap_rows = ((cap_rows - 8) / 16) * 16 + 8;
  const bool in_smem = cap_rows >= c + 32 && (size_t)RMAX <= (size_t)QR_SMEM_DOUBLES;
  if (in_smem) {
    const int Rp = min(cap_rows, ((R + 7) / 16) * 16 + 8);
    int done = 0, top = 0;
    while (done < R) {
      const int take = min(R - done, Rp - top);
      const int m = top + take;
      // ---- rows [done, done + take) of (H | r) under the triangle of the rows reduced so far; zero padding up to Rp
      for (int k = warp; k <= c; k += 16) {
        const double* src = ((k < c) ? Hs + (size_t)km[k] * RMAX : rs) + done;
        double* dst = qsm + (size_t)k * Rp;
        for (int i = lane; i < Rp - top; i += 32) dst[top + i] = (i < take) ? src[i] : 0.0;
      }
      __syncthreads();
      {                                                  // reflector of column 0: the only block-wide reduction of the pass
        double part = 0.0;
        for (int i = tid; i < m; i += 512) { const double x = qsm[i]; part += x * x; }
        part = warp_sum_d(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
          double n2 = 0.0; for (int w = 0; w < 16; ++w) n2 += red[w];
          const double x0 = qsm[0], nrm = sqrt(n2), alpha = x0 >= 0 ? -nrm : nrm, v0 = x0 - alpha, vtv = n2 - x0 * x0 + v0 * v0;
          qsm[0] = v0; s_diag[0] = alpha; s_beta[0] = vtv > 0.0 ? 2.0 / vtv : 0.0;
        }
        __syncthreads();
      }
      const int G = (m > 192) ? 32 : (m > 64) ? 16 : 8;  // lanes per trailing column; each lane walks pairs of rows
      const int gpw = 32 / G, grp = lane / G, gl = lane - grp * G;
      const int m_even = (m + 1) & ~1;                   // rows m..Rp-1 are zero
      for (int j = 0; j < c; ++j) {
        const double* cj = qsm + (size_t)j * Rp;         // rows >= j: the reflector (v0 already in place of the diagonal entry)
        const double beta = s_beta[j & 1];
        const int e0 = j & ~1;                           // pairs start at an even row; row j-1 of an odd j is masked out
        for (int kb = j + 1 + warp * gpw; kb <= c; kb += 16 * gpw) {      // warp-uniform trip count: full-mask shuffles below
          const int k = kb + grp;
          const bool act = k <= c;
          double* ck = qsm + (size_t)(act ? k : j + 1) * Rp;
          double dt_ = 0.0;
          if (act && beta != 0.0)
            for (int i = e0 + 2 * gl; i < m_even; i += 2 * G) {
              double2 r2 = *reinterpret_cast<const double2*>(cj + i);
              const double2 a2 = *reinterpret_cast<const double2*>(ck + i);
              if (i < j) r2.x = 0.0;
              dt_ += r2.x * a2.x + r2.y * a2.y;
            }
          for (int o = G >> 1; o; o >>= 1) dt_ += __shfl_xor_sync(0xffffffffu, dt_, o);
          dt_ *= beta;
          if (act && beta != 0.0)
            for (int i = e0 + 2 * gl; i < m_even; i += 2 * G) {
              double2 r2 = *reinterpret_cast<const double2*>(cj + i);
              double2 a2 = *reinterpret_cast<double2*>(ck + i);
              if (i < j) r2.x = 0.0;
              a2.x -= dt_ * r2.x; a2.y -= dt_ * r2.y;
              *reinterpret_cast<double2*>(ck + i) = a2;
            }
          if (kb == j + 1) {                             // (warp 0) the reflector of column j+1 for the next iteration
            __syncwarp();
            double part = 0.0;
            const int j1 = j + 1, f0 = j1 & ~1;
            if (grp == 0 && k < c)
              for (int i = f0 + 2 * gl; i < m_even; i += 2 * G) {
                const double2 a2 = *reinterpret_cast<const double2*>(ck + i);
                part += ((i < j1) ? 0.0 : a2.x * a2.x) + a2.y * a2.y;
              }
            for (int o = G >> 1; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0 && k < c) {
              const double x0 = ck[j1], nrm = sqrt(part), alpha = x0 >= 0 ? -nrm : nrm, v0 = x0 - alpha, vtv = part - x0 * x0 + v0 * v0;
              ck[j1] = v0; s_diag[j1] = alpha; s_beta[j1 & 1] = vtv > 0.0 ? 2.0 / vtv : 0.0;      // the other parity slot: slower warps may still read this column's
            }
          }
        }
        __syncthreads();
      }
      // ---- the panel's rows < c become the triangle: diagonal from s_diag, zeros below (the reflectors are dropped)
      for (int k = warp; k < c; k += 16) {
        double* ck = qsm + (size_t)k * Rp;
        for (int i = k + lane; i < c; i += 32) ck[i] = (i == k) ? s_diag[k] : 0.0;
      }
      __syncthreads();
      done += take; top = c;
    }
    // the triangular factor (rows < c) and Q^T r back to the global stack; rows >= c are dead after I_R = c
    for (int k = warp; k <= c; k += 16) {
      const double* src = qsm + (size_t)k * Rp;
      double* dst = (k < c) ? Hs + (size_t)km[k] * RMAX : rs;
      for (int i = lane; i < c; i += 32) dst[i] = src[i];
    }
  } else {
  {                                                    // column 0: the only block-wide reduction
    double part = 0.0;
    for (int i = tid; i < R; i += 512) { const double x = Hs[(size_t)km[0] * RMAX + i]; part += x * x; }
    part = warp_sum_d(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) { double n2 = 0.0; for (int w = 0; w < 16; ++w) n2 += red[w]; s_n2 = n2; s_x0 = Hs[(size_t)km[0] * RMAX]; }
    __syncthreads();
  }
  for (int j = 0; j < c; ++j) {
    double* cj = Hs + (size_t)km[j] * RMAX;
    const double n2 = s_n2, x0 = s_x0;                 // squared norm of rows >= j of column j, and its diagonal entry
    const double nrm = sqrt(n2);
    const double alpha = x0 >= 0 ? -nrm : nrm;
    const double v0 = x0 - alpha;
    const double vtv = n2 - x0 * x0 + v0 * v0;
    const double beta = vtv > 0.0 ? 2.0 / vtv : 0.0;
    for (int i = j + tid; i < R; i += 512) {
      const double x = cj[i];
      qsm[i] = (i == j) ? v0 : x;
      cj[i] = (i == j) ? alpha : 0.0;
    }
    __syncthreads();
    for (int k = j + 1 + warp; k <= c; k += 16) {
      double* ck = (k < c) ? Hs + (size_t)km[k] * RMAX : rs;
      if (beta != 0.0) {
        double dt_ = 0.0;
        for (int i = j + lane; i < R; i += 32) dt_ += qsm[i] * ck[i];
        dt_ = warp_sum_d(dt_) * beta;
        for (int i = j + lane; i < R; i += 32) ck[i] -= dt_ * qsm[i];
      }
      if (k == j + 1 && k < c) {                       // hand the next column's norm and diagonal to the next iteration
        __syncwarp();
        double part = 0.0;
        for (int i = j + 1 + lane; i < R; i += 32) { const double x = ck[i]; part += x * x; }
        part = warp_sum_d(part);
        if (lane == 0) { s_n2 = part; s_x0 = ck[j + 1]; }
      }
    }
    __syncthreads();
  }
  }
  if (tid == 0) {
    ic[I_R] = c;
    if (v.be.stats) { atomicAdd(&v.be.stats[8], 1ull); atomicAdd(&v.be.stats[9], (unsigned long long)R * c * c); }
  }
}

}  // namespace

namespace {

// ====================================================================== nonzero columns of the stacked Jacobian
// The stacked H_x of an update is structurally sparse: tracks are consumed at max_track_len observations, so every block touches
// the extrinsic/td columns 15..21 and the pose blocks of a handful of recent window slots only - at a full
// 30-pose window ~50 of 202 columns.  Zero columns contribute nothing to H P H^T, to the gain or to P - K H P, and a thin QR of
// H restricted to its nonzero columns spans the same row space; so compression and update run over the listed columns only
// (kmap[0 .. I_NC)), which turns the 2Rc^2 / 2rd^2 terms of the reference's dense algebra into 2R nc^2 / 2 r nc d.  The zero
// test is exact: structural zeros are never computed, they are the memset / stack-kernel zeros.
__global__ void __launch_bounds__(256) be_colscan_kernel(BeView v, int rows_idx) {
  __shared__ int flag[BE_DMAX_PAD];
  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int R = ic[rows_idx], d = ic[I_DIM];
  const int RMAX = v.be.RMAX, LD = v.be.LD;
  const double* Hs = v.be.Hs + (size_t)s * LD * RMAX;
  int* km = v.be.kmap + (size_t)s * LD;
  for (int j = warp; j < d; j += 8) {
    const double* cj = Hs + (size_t)j * RMAX;
    bool nz = false;
    for (int i = lane; i < R; i += 32) nz |= (cj[i] != 0.0);
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) flag[j] = nz ? 1 : 0;
  }
  __syncthreads();
  if (tid == 0) {
    int nc = 0;
    for (int j = 0; j < d; ++j) if (flag[j]) km[nc++] = j;
    ic[I_NC] = nc;
  }
}

// ====================================================================== batched FP64 GEMM (generic strides)
// C[M x N] = alpha * A[M x K] * B[K x N] + beta * C (+ diag on the diagonal), dims from icore.
struct GemmArgs {
  const double* A; size_t sA; int rsA, csA;
  const double* B; size_t sB; int rsB, csB;
  double* C; size_t sC; int rsC, csC;
  const int* icore; int m_idx, n_idx, k_idx;
  double alpha, beta, diag;
  const double* diag_vec; size_t sD;     // optional per-row diagonal term (ZUPT's block-diagonal R), else `diag`
  const int* kmap; size_t sK;            // optional: the contraction runs over the listed indices kmap[0 .. K) of A's and B's k axis
};

constexpr int GT = 64, GK = 16;
// The dense contractions of the update (T = H P, S = T H^T, P -= Y^T Y) on the FP64 tensor path.  tcgen05 has no f64 kind, so
// this is warp-level `mma.sync.m8n8k4.f64` (SASS: DMMA): 8 warps tile a 64x64 block as 4 (M) x 2 (N), each warp owns 2 x 4
// fragments of 8x8; per 4-deep k step a thread issues 6 shared-memory loads for 8 DMMAs (64 FMAs).  (A scalar-FMA version with
// 4x4 register tiles needed 8 loads per 16 FMAs and measured 10.4 us vs 9.1 us per launch, round 2.)
constexpr int GTP = GT + 4;      // 68: (k % 4) * 68 + (m % 8) hits 16 distinct 8-byte banks per half warp
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) be_gemm_kernel(GemmArgs g) {
  __shared__ double As[GK][GTP], Bs[GK][GTP];
  const int s = blockIdx.z;
  const int* ic = g.icore + (size_t)s * BE_ICORE;
  if (!ic[I_OK]) return;
  const int M = ic[g.m_idx], N = ic[g.n_idx], K = ic[g.k_idx];
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  if (m0 >= M || n0 >= N || K <= 0) return;
  const double* A = g.A + (size_t)s * g.sA;
  const double* B = g.B + (size_t)s * g.sB;
  double* C = g.C + (size_t)s * g.sC;
  const int* km = g.kmap ? g.kmap + (size_t)s * g.sK : nullptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;                  // fragment coordinates: row/col group, index inside the group
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;      // this warp's 16 x 32 patch of the 64 x 64 tile
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int i = tid; i < GT * GK; i += 256) {
      int mm, kk;
      if (g.rsA == 1) { mm = i % GT; kk = i / GT; } else { kk = i % GK; mm = i / GK; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[(size_t)gm * g.rsA + (size_t)(km ? km[gk] : gk) * g.csA] : 0.0;
    }
    for (int i = tid; i < GT * GK; i += 256) {
      int nn, kk;
      if (g.csB == 1) { nn = i % GT; kk = i / GT; } else { kk = i % GK; nn = i / GK; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? B[(size_t)(km ? km[gk] : gk) * g.rsB + (size_t)gn * g.csB] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double a[2], b[4];
      // A fragment (8 x 4, "row"): element (row = grp, k = tig); B fragment (4 x 8, "col"): element (k = tig, col = grp)
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = As[kk + tig][wm + 8 * i + grp];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk + tig][wn + 8 * j + grp];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // C fragment (8 x 8): elements (row = grp, col = 2 * tig + {0, 1})
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gm = m0 + wm + 8 * i + grp, gn = n0 + wn + 8 * j + 2 * tig + e;
        if (gm < M && gn < N) {
          double* c = C + (size_t)gm * g.rsC + (size_t)gn * g.csC;
          double val = g.alpha * acc[i][j][e];
          if (g.beta != 0.0) val += g.beta * *c;
          if (gm == gn) val += g.diag_vec ? g.diag_vec[(size_t)s * g.sD + gm] : g.diag;
          *c = val;
        }
      }
}

// ====================================================================== Cholesky S = L L^T (lower, in place) + z = L^-1 r
// The packed lower triangle (r <= 208 -> <= 174 KB) lives in shared memory for the whole factorisation; the
// right-looking update runs out of smem, L is written back to Sm (TRSM reads it) at the end.
__global__ void __launch_bounds__(1024) be_chol_kernel(BeView v) {
  extern __shared__ double csm[];   // packed lower triangle [r(r+1)/2] + column cache [r] + z [r]
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int r = ic[I_R];
  const int RC = v.be.chol_cap;
  if (r <= 0 || r > RC) return;                  // larger systems go through be_chol_gmem_kernel
  const int LD = v.be.LDS;
  double* S = v.be.Sm + (size_t)s * LD * LD;
  double* Lp = csm;
  double* col = csm + (size_t)RC * (RC + 1) / 2;
  double* z = col + RC;
#define LP(i, j) Lp[(size_t)(i) * ((i) + 1) / 2 + (j)]
  for (int e = tid; e < r * r; e += 1024) { const int i = e / r, j = e - i * r; if (j <= i) LP(i, j) = S[(size_t)i * LD + j]; }
  __syncthreads();
  for (int j = 0; j < r; ++j) {
    const double djj = sqrt(LP(j, j));
    const double inv = 1.0 / djj;
    __syncthreads();
    for (int i = j + tid; i < r; i += 1024) {
      const double x = (i == j) ? djj : LP(i, j) * inv;
      LP(i, j) = x;
      col[i] = x;
    }
    __syncthreads();
    // trailing update of the lower triangle: rows a in (j, r) over the warps, columns b in (j, a] over the lanes
    for (int a = j + 1 + (tid >> 5); a < r; a += 32) {
      const double ca = col[a];
      double* row = &LP(a, 0);
      for (int b2 = j + 1 + (tid & 31); b2 <= a; b2 += 32) row[b2] -= ca * col[b2];
    }
    __syncthreads();
  }
  for (int e = tid; e < r * r; e += 1024) { const int i = e / r, j = e - i * r; if (j <= i) S[(size_t)i * LD + j] = LP(i, j); }
  // z = L^-1 r  (warp 0, row oriented, out of smem)
  if (tid < 32) {
    const double* rs = v.be.rs + (size_t)s * v.be.RMAX;
    for (int i = 0; i < r; ++i) {
      double part = 0.0;
      for (int q = tid; q < i; q += 32) part += LP(i, q) * z[q];
      part = warp_sum_d(part);
      if (tid == 0) z[i] = (rs[i] - part) / LP(i, i);
      __syncwarp();
    }
    double* zg = v.be.zvec + (size_t)s * LD;
    for (int i = tid; i < r; i += 32) zg[i] = z[i];
  }
#undef LP
}


// global-memory variant for windows whose packed S does not fit in shared memory (sw_size > ~33)
__global__ void __launch_bounds__(512) be_chol_gmem_kernel(BeView v, int min_r) {
  extern __shared__ double csm[];   // column cache [LDS]
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int r = ic[I_R];
  if (r <= min_r) return;
  const int LD = v.be.LDS;
  double* S = v.be.Sm + (size_t)s * LD * LD;
  for (int j = 0; j < r; ++j) {
    const double djj = sqrt(S[(size_t)j * LD + j]);
    __syncthreads();
    for (int i = j + tid; i < r; i += 512) {
      const double x = (i == j) ? djj : S[(size_t)i * LD + j] / djj;
      S[(size_t)i * LD + j] = x;
      csm[i] = x;
    }
    __syncthreads();
    for (int a = j + 1 + (tid >> 5); a < r; a += 16) {
      const double ca = csm[a];
      double* row = S + (size_t)a * LD;
      for (int b = j + 1 + (tid & 31); b <= a; b += 32) row[b] -= ca * csm[b];
    }
    __syncthreads();
  }
  // z = L^-1 r  (warp 0, row oriented)
  if (tid < 32) {
    double* z = v.be.zvec + (size_t)s * LD;
    const double* rs = v.be.rs + (size_t)s * v.be.RMAX;
    for (int i = 0; i < r; ++i) {
      double part = 0.0;
      for (int q = tid; q < i; q += 32) part += S[(size_t)i * LD + q] * z[q];
      part = warp_sum_d(part);
      if (tid == 0) z[i] = (rs[i] - part) / S[(size_t)i * LD + i];
      __syncwarp();
    }
  }
}

// ====================================================================== Y = L^-1 T (in place on Tm), 64 columns per CTA
__global__ void __launch_bounds__(64) be_trsm_kernel(BeView v) {
  __shared__ double Lt[32][33];
  const int s = blockIdx.y, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int r = ic[I_R], d = ic[I_DIM];
  if (r <= 0) return;
  const int c = blockIdx.x * 64 + tid;
  if (blockIdx.x * 64 >= d) return;
  const bool act = c < d;
  const int LD = v.be.LD, LS = v.be.LDS;
  const double* L = v.be.Sm + (size_t)s * LS * LS;
  double* Tm = v.be.Tm + (size_t)s * v.be.RAWMAX * LD;
  const int nb = (r + 31) / 32;
  for (int ib = 0; ib < nb; ++ib) {
    const int i0 = ib * 32;
    for (int e = tid; e < 32 * 32; e += 64) {
      const int a = e / 32, b = e % 32;
      Lt[a][b] = (i0 + a < r && i0 + b < r) ? L[(size_t)(i0 + a) * LS + i0 + b] : ((a == b) ? 1.0 : 0.0);
    }
    __syncthreads();
    double y[32];
    if (act) {
#pragma unroll
      for (int a = 0; a < 32; ++a) y[a] = (i0 + a < r) ? Tm[(size_t)(i0 + a) * LD + c] : 0.0;
#pragma unroll
      for (int a = 0; a < 32; ++a) {
        double x = y[a];
#pragma unroll
        for (int q = 0; q < 32; ++q) if (q < a) x -= Lt[a][q] * y[q];
        y[a] = x / Lt[a][a];
      }
#pragma unroll
      for (int a = 0; a < 32; ++a) if (i0 + a < r) Tm[(size_t)(i0 + a) * LD + c] = y[a];
    }
    for (int jb = ib + 1; jb < nb; ++jb) {
      const int j0 = jb * 32;
      __syncthreads();
      for (int e = tid; e < 32 * 32; e += 64) {
        const int a = e / 32, b = e % 32;
        Lt[a][b] = (j0 + a < r && i0 + b < r) ? L[(size_t)(j0 + a) * LS + i0 + b] : 0.0;
      }
      __syncthreads();
      if (act) {
#pragma unroll 4
        for (int a = 0; a < 32; ++a) {
          if (j0 + a >= r) break;
          double x = 0.0;
#pragma unroll
          for (int q = 0; q < 32; ++q) x += Lt[a][q] * y[q];
          Tm[(size_t)(j0 + a) * LD + c] -= x;
        }
      }
    }
    __syncthreads();
  }
}


// ====================================================================== dx = Y^T z, state correction (:1476-1534, :1692-1750)
__global__ void __launch_bounds__(256) be_correct_kernel(BeView v) {
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int r = ic[I_R], d = ic[I_DIM];
  if (r <= 0) return;
  const int LD = v.be.LD;
  const double* Y = v.be.Tm + (size_t)s * v.be.RAWMAX * LD;
  const double* z = v.be.zvec + (size_t)s * v.be.LDS;
  double* dx = v.be.dx + (size_t)s * v.be.LDS;
  for (int c = tid; c < d; c += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < r; ++i) acc += Y[(size_t)i * LD + c] * z[i];
    dx[c] = acc;
  }
  __syncthreads();
  double* core = core_of(v, s);
  if (tid == 0) {
    double dq[4], qn[4];
    small_angle_quat(v3(dx[0], dx[1], dx[2]), dq);
    quat_mul(dq, core + C_Q, qn);
    for (int i = 0; i < 4; ++i) core[C_Q + i] = qn[i];
    for (int i = 0; i < 3; ++i) { core[C_V + i] += dx[3 + i]; core[C_P + i] += dx[6 + i]; core[C_BG + i] += dx[9 + i]; core[C_BA + i] += dx[12 + i]; }
    small_angle_quat(v3(dx[15], dx[16], dx[17]), dq);
    const M3 Rn = m3_mul(m3_load(core + C_RIC), m3_t(quat_to_rot(dq)));
    m3_store(core + C_RIC, Rn);
    for (int i = 0; i < 3; ++i) core[C_TCI + i] += dx[18 + i];
    core[C_TD] += dx[21];
    if (LEGD > 22) {            // T1 T2 T3 A1 A2 A3 M1 M2 += dx(22:46), then updateImuMx (:1497-1507, 3803-3847)
      const int lo[3] = {3, 6, 7}, di[3] = {0, 4, 8}, up[3] = {1, 2, 5};      // (1,0)(2,0)(2,1) / diagonal / (0,1)(0,2)(1,2)
      for (int i = 0; i < 3; ++i) {
        core[C_TG + lo[i]] += dx[22 + i]; core[C_TG + di[i]] += dx[25 + i]; core[C_TG + up[i]] += dx[28 + i];
        core[C_AS + lo[i]] += dx[31 + i]; core[C_AS + di[i]] += dx[34 + i]; core[C_AS + up[i]] += dx[37 + i];
        core[C_MA + lo[i]] += dx[40 + i]; core[C_MA + di[i]] += dx[43 + i];
      }
    }
    ic[I_UPDATES] += 1;
    if (v.be.stats) {
      atomicAdd(&v.be.stats[4], 1ull); atomicAdd(&v.be.stats[5], (unsigned long long)r);
      atomicAdd(&v.be.stats[6], (unsigned long long)r * d * d); atomicAdd(&v.be.stats[7], (unsigned long long)ic[I_ROWS]);
      atomicAdd(&v.be.stats[15], (unsigned long long)r * ic[I_NC] * d);
    }
  }
  __syncthreads();
  const int n_win = ic[I_NWIN];
  const M3 R_c2b = m3_t(m3_load(core + C_RIC));
  const V3 t_c_b = ld3(core + C_TCI);
  for (int i = tid; i < n_win; i += blockDim.x) {
    double* w = win_of(v, s, i);
    const double* da = dx + LEGD + 6 * i;
    double dq[4], qn[4];
    small_angle_quat(v3(da[0], da[1], da[2]), dq);
    quat_mul(dq, w + W_Q, qn);
    for (int k = 0; k < 4; ++k) w[W_Q + k] = qn[k];
    for (int k = 0; k < 3; ++k) w[W_P + k] += da[3 + k];
    const M3 R_b2w = quat_to_rot(w + W_Q);
    rot_to_quat(m3_mul(R_b2w, R_c2b), w + W_QCAM);
    st3(w + W_PCAM, ld3(w + W_P) + m3_vec(R_b2w, t_c_b));
  }
  __syncthreads();
  // inverse depth of the SLAM features already in the state and their world positions (:1536-1575, :1752-1801)
  const int nf = ic[I_NF], idp = v.be.IDP;
  const int base = LEGD + 6 * n_win;
  for (int i = tid; i < nf; i += blockDim.x) {
    const size_t fi = (size_t)s * v.be.T + v.be.fs_slot[(size_t)s * 64 + i];
    if (idp == 3) { v.be.ft_oa[fi * 2] += dx[base + 3 * i]; v.be.ft_oa[fi * 2 + 1] += dx[base + 3 * i + 1]; }
    const double inv = v.be.ft_inv[fi] + dx[base + idp * i + idp - 1];
    v.be.ft_inv[fi] = inv;
    const double* wa = anchor_rec(v, s, v.be.ft_anchor[fi]);
    const V3 p_c = v3(v.be.ft_oa[fi * 2] / inv, v.be.ft_oa[fi * 2 + 1] / inv, 1.0 / inv);
    st3(v.be.ft_pos + fi * 3, m3_vec(quat_to_rot(wa + W_QCAM), p_c) + ld3(wa + W_PCAM));
  }
}

// measurementUpdate_ZUPT_vpq (:2791-2830): the 9-row system [v; dp; dq] of a detected standstill, written into the
// stacked-Jacobian buffers so that the common update kernels apply it; sequences without ZUPT get an empty system.
__global__ void __launch_bounds__(256) be_zupt_build_kernel(BeView v) {
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  if (!ic[I_ZUPT]) { if (tid == 0) { ic[I_ROWS] = 0; ic[I_R] = 0; } return; }
  const int RMAX = v.be.RMAX, LD = v.be.LD;
  const int nf0 = ic[I_NF];
  // with nuisance states behind the feature block the reference's conservativeResize (:2770-2773) cuts the TAIL of the nuisance
  // block off and leaves nui_ids behind: its own state is inconsistent from there on - reported, not imitated
  if (tid == 0 && ic[I_NNUI] > 0) atomicExch(&ic[I_ERR], 6);
  const int d = ic[I_DIM] - v.be.IDP * nf0, N = ic[I_NWIN];        // read before thread 0 rewrites I_DIM below
  if (nf0 > 0) {                                        // :2770-2782: every SLAM feature leaves the state
    for (int i = tid; i < nf0; i += blockDim.x) {
      const size_t fi = (size_t)s * v.be.T + v.be.fs_slot[(size_t)s * 64 + i];
      v.be.ft_flags[fi] &= ~(2 | 4 | 8);
    }
  }
  __syncthreads();
  if (tid == 0) { ic[I_DIM] = d; ic[I_NF] = 0; core_of(v, s)[C_LAST_ZUPT] = core_of(v, s)[C_TIME]; }
  double* Hs = v.be.Hs + (size_t)s * LD * RMAX;
  double* rs = v.be.rs + (size_t)s * RMAX;
  double* dg = v.be.dx + (size_t)s * v.be.LDS;        // per-row measurement variances (dx is free until be_correct)
  for (int e = tid; e < 9 * d; e += blockDim.x) { const int col = e / 9, row = e - col * 9; Hs[(size_t)col * RMAX + row] = 0.0; }
  __syncthreads();
  if (tid < 3) {
    const int L = LEGD;
    Hs[(size_t)(3 + tid) * RMAX + tid] = 1.0;                                   // zupt_v
    Hs[(size_t)(L + 6 * N - 3 + tid) * RMAX + 3 + tid] = 1.0;                   // zupt_p current
    Hs[(size_t)(L + 6 * N - 9 + tid) * RMAX + 3 + tid] = -1.0;                  // zupt_p previous
    Hs[(size_t)(L + 6 * N - 6 + tid) * RMAX + 6 + tid] = -0.5;                  // zupt_q current
    Hs[(size_t)(L + 6 * N - 12 + tid) * RMAX + 6 + tid] = 0.5;                  // zupt_q previous
    dg[tid] = v.cfg.zupt_nv; dg[3 + tid] = v.cfg.zupt_np; dg[6 + tid] = v.cfg.zupt_nq;
  }
  if (tid == 0) {
    const double* core = core_of(v, s);
    const double* wc = win_of(v, s, N - 1);
    const double* wp = win_of(v, s, N - 2);
    for (int i = 0; i < 3; ++i) { rs[i] = -core[C_V + i]; rs[3 + i] = -(wc[W_P + i] - wp[W_P + i]); }
    const double qpc[4] = {-wp[W_Q], -wp[W_Q + 1], -wp[W_Q + 2], wp[W_Q + 3]};
    double dq[4];
    quat_mul(wc + W_Q, qpc, dq);
    rs[6] = dq[0]; rs[7] = dq[1]; rs[8] = dq[2];
    ic[I_ROWS] = 9; ic[I_R] = 9;
  }
}

// erase processed / invalid features (:2007-2009, :2248-2253)
__global__ void be_apply_actions_kernel(BeView v) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || i >= v.be.T) return;
  const size_t fi = (size_t)s * v.be.T + i;
  const int act = v.be.ft_action[fi];
  if (act == 1 || act == 2) { v.be.ft_flags[fi] = 0; v.be.ft_mask[fi] = 0; }   // SLAM features (3, 4) stay in the map
  v.be.ft_action[fi] = 0;
}

// ====================================================================== pruning
// findRedundantImuStates (:2259-2307) -> window slots to drop
__global__ void be_prune_select_kernel(BeView v) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.be.S) return;
  int* ic = icore_of(v, s);
  ic[I_DO_PRUNE] = 0; ic[I_NRM] = 0; ic[I_NEWNUI] = 0;
  if (!ic[I_OK]) return;
  const int n = ic[I_NWIN];
  if (ic[I_ZUPT]) {                                       // :2321-2325: the previous state goes, whatever the window size
    ic[I_RM0] = n - 2; ic[I_RM1] = -1; ic[I_NRM] = 1; ic[I_DO_PRUNE] = 1;
    return;
  }
  if (n < v.cfg.sw_size) return;
  const double* core = core_of(v, s);
  int key = n - 4, st = key + 1, first = 0;
  const double* wk = win_of(v, s, key);
  const M3 key_R = quat_to_rot(wk + W_QCAM);
  const V3 key_p = ld3(wk + W_PCAM);
  int rm[2];
  for (int k = 0; k < 2; ++k) {
    const double* w = win_of(v, s, st);
    const M3 Rr = m3_t(quat_to_rot(w + W_QCAM));
    const double dist = norm(ld3(w + W_PCAM) - key_p);
    double q[4];
    rot_to_quat(m3_mul(Rr, key_R), q);
    const double nv = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double angle = 2.0 * atan2(nv, fabs(q[3]));
    if (angle < v.cfg.rot_thr && dist < v.cfg.trans_thr && core[C_TRACK_RATE] > v.cfg.track_thr) { rm[k] = st; ++st; }
    else { rm[k] = first; ++first; st -= 2; }
  }
  if (rm[0] > rm[1]) { const int t = rm[0]; rm[0] = rm[1]; rm[1] = t; }
  ic[I_RM0] = rm[0]; ic[I_RM1] = rm[1]; ic[I_NRM] = 2; ic[I_DO_PRUNE] = 1;
}

// getNewAnchorId (:3412-3472): among the first n-2 window states that observed the feature and stay, the one with the
// smallest reprojection distance; otherwise the newest state
__device__ int new_anchor_slot(const BeView& v, int s, size_t fi, int n_win, int r0, int r1) {
  if (n_win <= 2) return n_win - 1;
  const unsigned long long mask = v.be.ft_mask[fi];
  const V3 p_w = ld3(v.be.ft_pos + fi * 3);
  const double* obs = v.be.ft_obs + fi * v.be.Wcap * 4;
  int best = -1; double min_dis = 99999;
  for (int i = 0; i < n_win - 2; ++i) {
    if (!((mask >> i) & 1) || i == r0 || i == r1) continue;
    const double* w = win_of(v, s, i);
    const V3 pn = m3_tvec(quat_to_rot(w + W_QCAM), p_w - ld3(w + W_PCAM));
    const double dx = pn.x / pn.z - obs[(size_t)i * 4], dy = pn.y / pn.z - obs[(size_t)i * 4 + 1];
    const double dis = sqrt(dx * dx + dy * dy);
    if (min_dis > dis) { min_dis = dis; best = i; }
  }
  return best >= 0 ? best : n_win - 1;
}

// anchor changes of pruneImuStateBuffer (:2345-2461) for features whose anchor pose is about to leave the window:
// in-state SLAM features get updateFeatureCov_1didp (:3125-3293) / _3didp (:2965-3122), one after the other in feature-id
// order (each rewrites its rows/columns of P that the next one reads); potential SLAM features outside the state are only
// re-anchored.  With 3-D inverse depth the new anchor is always the newest state (:2361-2378, :2420-2437).
__global__ void __launch_bounds__(256) be_anchor_kernel(BeView v) {
  __shared__ double Jv[3][20]; __shared__ int Jc[20];
  __shared__ int s_list[64]; __shared__ int s_n, s_nj;
  extern __shared__ double pfl[];          // [IDP][LD]
  const int idp = v.be.IDP;
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_DO_PRUNE]) return;
  const int T = v.be.T, LD = v.be.LD, Wcap = v.be.Wcap;
  const int n_win = ic[I_NWIN], r0 = ic[I_RM0], r1 = (ic[I_NRM] > 1) ? ic[I_RM1] : -1;
  const bool fej = ic[I_FEJ] != 0;
  double* P = P_of(v, s);
  const double* core = core_of(v, s);
  const int d = ic[I_DIM];
  if (tid == 0) {
    // potential SLAM features outside the state (ekf flag, initialised): plain re-anchoring (:2408-2461)
    for (int i = 0; i < T; ++i) {
      const size_t fi = (size_t)s * T + i;
      const int fl = v.be.ft_flags[fi];
      if (!(fl & 1) || (fl & 4) || !(fl & 8) || !(fl & 2)) continue;
      const int as = v.be.ft_anchor[fi];
      if (as != r0 && as != r1) continue;
      if (!((v.be.ft_mask[fi] >> as) & 1)) continue;
      const int ns = (idp == 3) ? n_win - 1 : new_anchor_slot(v, s, fi, n_win, r0, r1);
      const double* w = win_of(v, s, ns);
      const V3 pn = m3_tvec(quat_to_rot(w + W_QCAM), ld3(v.be.ft_pos + fi * 3) - ld3(w + W_PCAM));
      v.be.ft_inv[fi] = 1.0 / pn.z;
      if (idp == 3) { v.be.ft_oa[fi * 2] = pn.x / pn.z; v.be.ft_oa[fi * 2 + 1] = pn.y / pn.z; }      // invParam (:2429-2432)
      else { v.be.ft_oa[fi * 2] = v.be.ft_obs[(fi * Wcap + ns) * 4]; v.be.ft_oa[fi * 2 + 1] = v.be.ft_obs[(fi * Wcap + ns) * 4 + 1]; }
      v.be.ft_anchor[fi] = ns;
    }
    // in-state features to re-anchor, sorted by feature id
    const int* fs = v.be.fs_slot + (size_t)s * 64;
    int n = 0;
    // use_schmidt (:2351-2358): a MATURE anchor pose (more than 2 states old - never one of the two newest poses the pruning
    // rule may pick) is not handed over: it stays behind the feature block as a nuisance state, with every feature it anchors
    int newnui = 0;
    if (v.be.NUI > 0) {
      for (int k = 0; k < 2; ++k) {
        const int r = k ? r1 : r0;
        if (r < 0 || (long long)ic[I_ID] - v.be.win_id[(size_t)s * Wcap + r] <= 2) continue;
        for (int i = 0; i < ic[I_NF]; ++i) if (v.be.ft_anchor[(size_t)s * T + fs[i]] == r) { newnui |= 1 << k; break; }
      }
      const int cnt = (newnui & 1) + ((newnui >> 1) & 1);
      if (ic[I_NNUI] + cnt > v.be.NUI) { atomicExch(&ic[I_ERR], 6); newnui = 0; }     // capacity: reported, handled as without Schmidt
      ic[I_NEWNUI] = newnui;
    }
    for (int i = 0; i < ic[I_NF]; ++i) {
      const size_t fi = (size_t)s * T + fs[i];
      const int as = v.be.ft_anchor[fi];
      if (as != r0 && as != r1) continue;
      if ((as == r0 && (newnui & 1)) || (as == r1 && (newnui & 2))) continue;
      int k = n++;
      while (k > 0 && v.be.ft_id[(size_t)s * T + fs[s_list[k - 1]]] > v.be.ft_id[fi]) { s_list[k] = s_list[k - 1]; --k; }
      s_list[k] = i;
    }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  const int* fs = v.be.fs_slot + (size_t)s * 64;
  for (int it = 0; it < n; ++it) {
    const int fcnt = s_list[it];
    const size_t fi = (size_t)s * T + fs[fcnt];
    const int fidx = LEGD + 6 * n_win + idp * fcnt;
    if (tid == 0 && idp == 3) {
      const int os = v.be.ft_anchor[fi];
      const int ns = n_win - 1;
      const double* wo = win_of(v, s, os);
      const double* wn = win_of(v, s, ns);
      const V3 p_w = ld3(v.be.ft_pos + fi * 3), p_fej = ld3(v.be.ft_pfej + fi * 3);
      const V3 pnew = m3_tvec(quat_to_rot(wn + W_QCAM), p_w - ld3(wn + W_PCAM));
      const double iv[3] = {pnew.x / pnew.z, pnew.y / pnew.z, 1.0 / pnew.z};      // invParam in the new anchor (:2366-2370)
      v.be.ft_oa[fi * 2] = iv[0]; v.be.ft_oa[fi * 2 + 1] = iv[1]; v.be.ft_inv[fi] = iv[2];
      // ---- updateFeatureCov_3didp, literally: the reference looks the "new" pose and its column block up with
      // old_state_id (:3000, :3066), so every "new" quantity below is the OLD anchor's and H_x_new lands in the old block
      const M3 R_b2c = m3_load(core + C_RIC); const V3 t_c_b = ld3(core + C_TCI);
      const M3 R_b2w_old = quat_to_rot(wo + W_Q), R_c2w_old = quat_to_rot(wo + W_QCAM);
      V3 p_old;
      if (fej) p_old = m3_vec(R_b2c, m3_tvec(R_b2w_old, p_fej - ld3(wo + W_PFEJ)) - t_c_b);
      else p_old = m3_tvec(R_c2w_old, p_w - ld3(wo + W_PCAM));
      const M3 R_w2b_new = m3_t(R_b2w_old), R_w2c_new = m3_t(R_c2w_old);
      const V3 pbn = fej ? (p_fej - ld3(wo + W_PFEJ)) : (p_w - ld3(wo + W_P));
      double Jfp[3][3] = {{1.0, 0.0, -iv[0]}, {0.0, 1.0, -iv[1]}, {0.0, 0.0, -iv[2]}};
      for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) Jfp[i][c] = iv[2] * Jfp[i][c];
      const M3 Jp = m3_mul(R_w2c_new, R_c2w_old);
      const M3 Jxn_l = m3_mul(R_w2c_new, skew(pbn));                    // J_x_new = [R skew(p_bf_new) | -R]
      const M3 Sk = skew(m3_vec(R_w2b_new, pbn) - t_c_b);
      const M3 Rno = m3_mul(R_w2b_new, R_b2w_old);
      const M3 Mx = m3_mul(Rno, skew(m3_tvec(R_b2c, p_old)));
      const M3 Jet = m3_mul(R_b2c, m3_sub(Sk, Mx));
      const M3 Jep = m3_mul(R_b2c, m3_sub(Rno, m3_identity()));
      double Jpf[3][3] = {{1.0, 0.0, -p_old.x}, {0.0, 1.0, -p_old.y}, {0.0, 0.0, -p_old.z}};
      for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) Jpf[i][c] = p_old.z * Jpf[i][c];
      for (int c = 0; c < 3; ++c) { Jc[c] = fidx + c; Jc[3 + c] = LEGD + 6 * os + c; Jc[6 + c] = LEGD + 6 * os + 3 + c; Jc[9 + c] = 15 + c; Jc[12 + c] = 18 + c; }
      for (int a = 0; a < 3; ++a) {
        double fp[3];                                                   // row a of J_fp_new * J_p
        for (int c = 0; c < 3; ++c) fp[c] = Jfp[a][0] * Jp.m[c] + Jfp[a][1] * Jp.m[3 + c] + Jfp[a][2] * Jp.m[6 + c];
        for (int c = 0; c < 3; ++c) {
          Jv[a][c] = fp[0] * Jpf[0][c] + fp[1] * Jpf[1][c] + fp[2] * Jpf[2][c];
          Jv[a][3 + c] = Jfp[a][0] * Jxn_l.m[c] + Jfp[a][1] * Jxn_l.m[3 + c] + Jfp[a][2] * Jxn_l.m[6 + c];
          Jv[a][6 + c] = -(Jfp[a][0] * R_w2c_new.m[c] + Jfp[a][1] * R_w2c_new.m[3 + c] + Jfp[a][2] * R_w2c_new.m[6 + c]);
          Jv[a][9 + c] = Jfp[a][0] * Jet.m[c] + Jfp[a][1] * Jet.m[3 + c] + Jfp[a][2] * Jet.m[6 + c];
          Jv[a][12 + c] = Jfp[a][0] * Jep.m[c] + Jfp[a][1] * Jep.m[3 + c] + Jfp[a][2] * Jep.m[6 + c];
        }
      }
      s_nj = 15;
      v.be.ft_anchor[fi] = ns;
    }
    if (tid == 0 && idp == 1) {
      const int os = v.be.ft_anchor[fi];
      const int ns = new_anchor_slot(v, s, fi, n_win, r0, r1);
      const double* wo = win_of(v, s, os);
      const double* wn = win_of(v, s, ns);
      const V3 p_w = ld3(v.be.ft_pos + fi * 3), p_fej = ld3(v.be.ft_pfej + fi * 3);
      const M3 R_c2w_new = quat_to_rot(wn + W_QCAM);
      const V3 pnew = m3_tvec(R_c2w_new, p_w - ld3(wn + W_PCAM));
      v.be.ft_inv[fi] = 1.0 / pnew.z;                                 // :2393-2397
      v.be.ft_oa[fi * 2] = pnew.x / pnew.z; v.be.ft_oa[fi * 2 + 1] = pnew.y / pnew.z;
      // ---- updateFeatureCov_1didp
      const M3 R_b2c = m3_load(core + C_RIC); const V3 t_c_b = ld3(core + C_TCI);
      const M3 R_b2w_old = quat_to_rot(wo + W_Q), R_c2w_old = quat_to_rot(wo + W_QCAM);
      V3 p_old;
      if (fej) p_old = m3_vec(R_b2c, m3_tvec(R_b2w_old, p_fej - ld3(wo + W_PFEJ)) - t_c_b);
      else p_old = m3_tvec(R_c2w_old, p_w - ld3(wo + W_PCAM));
      const V3 p_old_ = m3_tvec(R_c2w_old, p_w - ld3(wo + W_PCAM));
      const double inv_old = 1.0 / p_old_.z;
      const V3 f_old = v3(p_old_.x / p_old_.z, p_old_.y / p_old_.z, 1.0);
      const M3 R_b2w_new = quat_to_rot(wn + W_Q), R_w2b_new = m3_t(R_b2w_new);
      const M3 R_w2c_new = m3_t(R_c2w_new);
      const double inv_new = v.be.ft_inv[fi];
      V3 pbo, pbn;
      if (fej) { pbo = p_fej - ld3(wo + W_PFEJ); pbn = p_fej - ld3(wn + W_PFEJ); }
      else { pbo = p_w - ld3(wo + W_P); pbn = p_w - ld3(wn + W_P); }
      const double Jr = -inv_new * inv_new;
      const double J_d = m3_vec(R_w2c_new, m3_vec(R_c2w_old, f_old)).z;
      const M3 Jto = m3_scale(m3_mul(R_w2c_new, skew(pbo)), -1.0);
      const M3 Jtn = m3_mul(R_w2c_new, skew(pbn));
      const M3 Sk = skew(m3_vec(R_w2b_new, pbn) - t_c_b);
      const M3 Rno = m3_mul(R_w2b_new, R_b2w_old);
      const M3 Mx = m3_mul(Rno, skew(m3_tvec(R_b2c, p_old)));
      const M3 Jet = m3_mul(R_b2c, m3_sub(Sk, Mx));
      const M3 Jep = m3_mul(R_b2c, m3_sub(Rno, m3_identity()));
      int k = 0;
      Jc[k] = fidx; Jv[0][k++] = Jr * J_d * (-1.0 / (inv_old * inv_old));
      for (int c = 0; c < 3; ++c) { Jc[k] = LEGD + 6 * os + c; Jv[0][k++] = Jr * Jto.m[6 + c]; }
      for (int c = 0; c < 3; ++c) { Jc[k] = LEGD + 6 * os + 3 + c; Jv[0][k++] = Jr * R_w2c_new.m[6 + c]; }
      for (int c = 0; c < 3; ++c) { Jc[k] = LEGD + 6 * ns + c; Jv[0][k++] = Jr * Jtn.m[6 + c]; }
      for (int c = 0; c < 3; ++c) { Jc[k] = LEGD + 6 * ns + 3 + c; Jv[0][k++] = Jr * -R_w2c_new.m[6 + c]; }
      for (int c = 0; c < 3; ++c) { Jc[k] = 15 + c; Jv[0][k++] = Jr * Jet.m[6 + c]; }
      for (int c = 0; c < 3; ++c) { Jc[k] = 18 + c; Jv[0][k++] = Jr * Jep.m[6 + c]; }
      s_nj = k;
      v.be.ft_anchor[fi] = ns;
    }
    __syncthreads();
    const int nj = s_nj;
    // P_fl = J P (IDP rows), P_ff = P_fl J^T; the feature's rows and columns of P are replaced (:3100-3121, :3272-3292)
    for (int e = tid; e < idp * d; e += blockDim.x) {
      const int a = e / d, c = e - a * d;
      double acc = 0.0;
      for (int k = 0; k < nj; ++k) acc += Jv[a][k] * P[(size_t)Jc[k] * LD + c];
      pfl[a * LD + c] = acc;
    }
    __syncthreads();
    double pff = 0.0;
    if (tid < idp * idp) { const int a = tid / idp, b = tid - a * idp; for (int k = 0; k < nj; ++k) pff += pfl[a * LD + Jc[k]] * Jv[b][k]; }
    __syncthreads();
    for (int e = tid; e < idp * d; e += blockDim.x) {
      const int a = e / d, c = e - a * d;
      if (c < fidx || c >= fidx + idp) { P[(size_t)(fidx + a) * LD + c] = pfl[a * LD + c]; P[(size_t)c * LD + fidx + a] = pfl[a * LD + c]; }
    }
    if (tid < idp * idp) P[(size_t)(fidx + tid / idp) * LD + fidx + tid % idp] = pff;
    __syncthreads();
    if (idp == 3 && tid < 3) {                    // (P + P^T) / 2 (:3122): only the 3x3 block can be unsymmetric
      const int a = tid == 2 ? 1 : 0, b = tid == 0 ? 1 : 2;
      const double mval = 0.5 * (P[(size_t)(fidx + a) * LD + fidx + b] + P[(size_t)(fidx + b) * LD + fidx + a]);
      P[(size_t)(fidx + a) * LD + fidx + b] = mval; P[(size_t)(fidx + b) * LD + fidx + a] = mval;
    }
    __syncthreads();
  }
}

// drop the observations of the removed states from every feature and re-pack slots (:2530-2532, :2556-2558),
// then re-pack the window arrays (:2636-2637)
__global__ void __launch_bounds__(256) be_prune_tables_kernel(BeView v) {
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_DO_PRUNE]) return;
  const int T = v.be.T, Wcap = v.be.Wcap;
  const int n = ic[I_NWIN];
  const int r0 = ic[I_RM0], r1 = (ic[I_NRM] > 1) ? ic[I_RM1] : -1;
  // new nuisance states (:2611-2613): the pose record as the update left it, the number of features it anchors; in the order
  // of the removed ids (r0 < r1), behind the existing ones
  const int newnui = ic[I_NEWNUI];
  const int idx0 = ic[I_NNUI], idx1 = idx0 + (newnui & 1);
  if (newnui) {
    if (tid < BE_WIN) {
      if (newnui & 1) v.be.nui_win[((size_t)s * v.be.NUI + idx0) * BE_WIN + tid] = win_of(v, s, r0)[tid];
      if (newnui & 2) v.be.nui_win[((size_t)s * v.be.NUI + idx1) * BE_WIN + tid] = win_of(v, s, r1)[tid];
    }
    if (tid == 0) {
      int c0 = 0, c1 = 0;
      const int* fs = v.be.fs_slot + (size_t)s * 64;
      for (int i = 0; i < ic[I_NF]; ++i) { const int as = v.be.ft_anchor[(size_t)s * T + fs[i]]; c0 += (as == r0); c1 += (as == r1); }
      if (newnui & 1) v.be.nui_cnt[(size_t)s * v.be.NUI + idx0] = c0;
      if (newnui & 2) v.be.nui_cnt[(size_t)s * v.be.NUI + idx1] = c1;
    }
  }
  __syncthreads();
  for (int i = tid; i < T; i += blockDim.x) {
    const size_t fi = (size_t)s * T + i;
    if (!(v.be.ft_flags[fi] & 1)) continue;
    const unsigned long long m = v.be.ft_mask[fi];
    unsigned long long nm = 0;
    double* obs = v.be.ft_obs + fi * Wcap * 4;
    int dst = 0;
    for (int w = 0; w < n; ++w) {
      if (w == r0 || w == r1) continue;
      if ((m >> w) & 1) {
        nm |= 1ull << dst;
        if (dst != w) for (int k = 0; k < 4; ++k) obs[(size_t)dst * 4 + k] = obs[(size_t)w * 4 + k];
      }
      ++dst;
    }
    v.be.ft_mask[fi] = nm;
    const int as = v.be.ft_anchor[fi];                 // anchors are window slots: follow the re-pack (nuisance anchors stay)
    if ((v.be.ft_flags[fi] & 4) && as == r0 && (newnui & 1)) v.be.ft_anchor[fi] = BE_NUI_BASE + idx0;
    else if ((v.be.ft_flags[fi] & 4) && r1 >= 0 && as == r1 && (newnui & 2)) v.be.ft_anchor[fi] = BE_NUI_BASE + idx1;
    else if (as >= 0 && as < BE_NUI_BASE) v.be.ft_anchor[fi] = as - (as > r0 ? 1 : 0) - ((r1 >= 0 && as > r1) ? 1 : 0);
  }
  __syncthreads();
  if (tid < BE_WIN + 1) {
    int dst = 0;
    for (int w = 0; w < n; ++w) {
      if (w == r0 || w == r1) continue;
      if (dst != w) {
        if (tid < BE_WIN) win_of(v, s, dst)[tid] = win_of(v, s, w)[tid];
        else v.be.win_id[(size_t)s * Wcap + dst] = v.be.win_id[(size_t)s * Wcap + w];
      }
      ++dst;
    }
  }
}

// covariance re-pack (:2563-2634): gather the kept rows/cols into Sm, then copy back
__global__ void __launch_bounds__(256) be_prune_cov_gather_kernel(BeView v) {
  const int s = blockIdx.y, row = blockIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_DO_PRUNE]) return;
  const int d = ic[I_DIM], LD = v.be.LD;
  const int nrm = ic[I_NRM], newnui = ic[I_NEWNUI];
  const int ndk = d - 6 * nrm;                                     // columns that stay where they are (relative order)
  const int nd = ndk + 6 * ((newnui & 1) + ((newnui >> 1) & 1));   // + the pose blocks that become nuisance states, moved to the end (:2569-2610)
  if (row >= nd) return;
  const int a0 = LEGD + 6 * ic[I_RM0], a1 = (nrm > 1) ? LEGD + 6 * ic[I_RM1] : (1 << 30);
  auto src = [&](int i) {
    if (i >= ndk) { const int j = i - ndk; const bool first = (j < 6) && (newnui & 1); return (first ? a0 : a1) + j % 6; }
    int x = i; if (x >= a0) x += 6; if (x >= a1) x += 6; return x;
  };
  const double* P = P_of(v, s);
  double* Sd = v.be.Sm + (size_t)s * v.be.LDS * v.be.LDS;
  const int sr = src(row);
  for (int c = threadIdx.x; c < nd; c += blockDim.x) Sd[(size_t)row * LD + c] = P[(size_t)sr * LD + src(c)];
}
__global__ void __launch_bounds__(256) be_prune_cov_scatter_kernel(BeView v) {
  const int s = blockIdx.y, row = blockIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_DO_PRUNE]) return;
  const int d = ic[I_DIM], LD = v.be.LD;
  const int newnui = ic[I_NEWNUI];
  const int nd = d - 6 * ic[I_NRM] + 6 * ((newnui & 1) + ((newnui >> 1) & 1));
  if (row >= nd) return;
  double* P = P_of(v, s);
  const double* Sd = v.be.Sm + (size_t)s * v.be.LDS * v.be.LDS;
  for (int c = threadIdx.x; c < nd; c += blockDim.x) P[(size_t)row * LD + c] = Sd[(size_t)row * LD + c];
}

// one map point into list `which` of sequence s (upsert by id: a std::map assignment in the reference)
__device__ inline void pts_put(const BeView& v, int which, int s, unsigned long long id, const double* xyz) {
  const int cap = v.be.PCAP;
  const size_t b = ((size_t)which * v.be.S + s);
  unsigned long long* ids = v.be.pts_id + b * cap;
  double* pos = v.be.pts_xyz + b * cap * 3;
  int n = v.be.pts_n[b];
  int k = 0;
  while (k < n && ids[k] != id) ++k;
  if (k == n) {
    if (n >= cap) { v.be.pts_drop[b] += 1; return; }
    ids[k] = id; v.be.pts_n[b] = n + 1;
  }
  pos[k * 3] = xyz[0]; pos[k * 3 + 1] = xyz[1]; pos[k * 3 + 2] = xyz[2];
}

// end of processFeatures: shrink dims after pruning, FEJ switch (:414-419), active map points (:455-458).  One warp per
// sequence: lane 0 does the scalar bookkeeping, the id search of the map-point upsert runs across the lanes.
__global__ void __launch_bounds__(32) be_frame_end_kernel(BeView v) {
  const int s = blockIdx.x, lane = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int nf = ic[I_NF];
  __syncwarp();
  if (lane == 0) {
    if (ic[I_DO_PRUNE]) {
      const int nk = (ic[I_NEWNUI] & 1) + ((ic[I_NEWNUI] >> 1) & 1);       // removed poses that stay in the covariance as nuisance states
      ic[I_DIM] -= 6 * (ic[I_NRM] - nk); ic[I_NWIN] -= ic[I_NRM]; ic[I_NNUI] += nk; ic[I_NEWNUI] = 0;
    }
    const double* core = core_of(v, s);
    if (v.cfg.if_FEJ_config && !ic[I_FEJ] && core[C_TIME] - core[C_TAKEOFF] >= 0) ic[I_FEJ] = 1;
  }
  if (v.be.PCAP > 0 && nf > 0) {
    const int cap = v.be.PCAP;
    const size_t b = (size_t)1 * v.be.S + s;                      // list 1: active map points
    unsigned long long* ids = v.be.pts_id + b * cap;
    double* pos = v.be.pts_xyz + b * cap * 3;
    const int* fs = v.be.fs_slot + (size_t)s * 64;
    int n = v.be.pts_n[b];
    for (int i = 0; i < nf; ++i) {
      const size_t fi = (size_t)s * v.be.T + fs[i];
      const unsigned long long id = v.be.ft_id[fi];
      int k = -1;
      for (int k0 = 0; k0 < n && k < 0; k0 += 32) {
        const unsigned hit = __ballot_sync(0xffffffffu, k0 + lane < n && ids[k0 + lane] == id);
        if (hit) k = k0 + __ffs(hit) - 1;
      }
      if (k < 0) {
        if (n >= cap) { if (lane == 0) v.be.pts_drop[b] += 1; continue; }
        k = n++;
        if (lane == 0) ids[k] = id;
      }
      if (lane < 3) pos[k * 3 + lane] = v.be.ft_pos[fi * 3 + lane];
      __syncwarp();
    }
    if (lane == 0) v.be.pts_n[b] = n;
  }
}

}  // namespace

namespace {

// ---------------------------------------------------------------- generic covariance re-map P'[i][j] = P[m(i)][m(j)]
// used by stateAugmentation with SLAM features in the state (:768-793), rmLostFeaturesCov (:3296-3348) and the drop of all
// SLAM features on ZUPT.  Gather into the Sm scratch, copy back, commit the new dimension.
__global__ void __launch_bounds__(256) be_remap_gather_kernel(BeView v) {
  const int s = blockIdx.y, row = blockIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_REMAP]) return;
  const int nd = ic[I_NEWDIM], LD = v.be.LD;
  if (row >= nd) return;
  const int* cm = v.be.cmap + (size_t)s * LD;
  const double* P = P_of(v, s);
  double* Sd = v.be.Sm + (size_t)s * v.be.LDS * v.be.LDS;
  const int sr = cm[row];
  for (int c = threadIdx.x; c < nd; c += blockDim.x) Sd[(size_t)row * LD + c] = P[(size_t)sr * LD + cm[c]];
}
__global__ void __launch_bounds__(256) be_remap_scatter_kernel(BeView v) {
  const int s = blockIdx.y, row = blockIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_REMAP]) return;
  const int nd = ic[I_NEWDIM], LD = v.be.LD;
  if (row >= nd) return;
  double* P = P_of(v, s);
  const double* Sd = v.be.Sm + (size_t)s * v.be.LDS * v.be.LDS;
  for (int c = threadIdx.x; c < nd; c += blockDim.x) P[(size_t)row * LD + c] = Sd[(size_t)row * LD + c];
}
__global__ void be_remap_commit_kernel(BeView v) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.be.S) return;
  int* ic = icore_of(v, s);
  if (!ic[I_OK] || !ic[I_REMAP]) return;
  ic[I_DIM] = ic[I_NEWDIM]; ic[I_REMAP] = 0;
}

__device__ __forceinline__ int grid_code(const BeView& v, double x, double y) {
  const int row = (int)((y - v.cfg.y_min) / v.cfg.grid_h), col = (int)((x - v.cfg.x_min) / v.cfg.grid_w);
  return row * v.be.grid_cols + col;
}

// grid_map[code] of the reference (a std::map<int, vector>): the rows*cols cells are emptied by updateGridMap every step
// (larvio.cpp:3355-3357) and live in cand[s][64..127]; any other code is created on first touch (:3366, :1972, :1988) and
// never emptied - those counts persist in grid_oor.  nullptr: a code outside the tracked range [-64, 192).
__device__ __forceinline__ int* grid_cell(const BeView& v, int s, int* grid, int code) {
  if (code >= 0 && code < v.be.grid_rows * v.be.grid_cols) return grid + code;
  if (code >= -BE_GRID_OOR_NEG && code < BE_GRID_OOR - BE_GRID_OOR_NEG) return v.be.grid_oor + (size_t)s * BE_GRID_OOR + code + BE_GRID_OOR_NEG;
  return nullptr;
}

// ---------------------------------------------------------------- in-state SLAM features at the start of removeLostFeatures
// (:1896-1924): lost ones leave the state (covariance re-map, feature erased); grid occupancy of the rest (updateGridMap).
// grid counts live in cand[s][64..127].
__global__ void __launch_bounds__(64) be_slam_pre_kernel(BeView v) {
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (!ic[I_OK]) return;
  const int T = v.be.T, Wcap = v.be.Wcap, LD = v.be.LD;
  int* grid = v.be.cand + (size_t)s * 128 + 64;
  if (tid < 64) grid[tid] = 0;
  __syncthreads();
  if (tid != 0) return;
  const int nf = ic[I_NF], nnui = ic[I_NNUI];
  if (nf == 0 && nnui == 0) return;
  int* fs = v.be.fs_slot + (size_t)s * 64;
  const int cur = ic[I_NWIN] - 1, d = ic[I_DIM];
  const int idp = v.be.IDP;
  const int base = d - idp * nf - 6 * nnui;
  int* cm = v.be.cmap + (size_t)s * LD;
  int* ncnt = v.be.nui_cnt + (size_t)s * (v.be.NUI > 0 ? v.be.NUI : 1);
  int keep = 0;
  for (int i = 0; i < base; ++i) cm[i] = i;
  for (int i = 0; i < nf; ++i) {
    const int slot = fs[i];
    const size_t fi = (size_t)s * T + slot;
    if ((v.be.ft_mask[fi] >> cur) & 1) {
      for (int q = 0; q < idp; ++q) cm[base + idp * keep + q] = base + idp * i + q;
      fs[keep++] = slot;
    } else {
      const int as = v.be.ft_anchor[fi];
      if (as >= BE_NUI_BASE) ncnt[as - BE_NUI_BASE] -= 1;       // nui_features[id_anchor].erase(feature) (:3329-3339)
      pts_put(v, 0, s, v.be.ft_id[fi], v.be.ft_pos + fi * 3);   // lost_slam_features[id] = map_server[id] (:3342)
      v.be.ft_flags[fi] = 0; v.be.ft_mask[fi] = 0;              // rmLostFeaturesCov erases the feature
    }
  }
  // rmUselessNuisanceState (:3850-3895): nuisance states that anchor no feature any more leave the covariance; the others
  // close ranks (their order is the order of nui_ids), and the anchors of the remaining features follow
  int nkeep = 0;
  if (nnui > 0) {
    int newidx[BE_NUI_MAX];
    const int c0 = base + idp * keep, o0 = base + idp * nf;
    for (int k = 0; k < nnui; ++k) {
      if (ncnt[k] > 0) {
        for (int q = 0; q < 6; ++q) cm[c0 + 6 * nkeep + q] = o0 + 6 * k + q;
        if (nkeep != k) {
          double* dst = v.be.nui_win + ((size_t)s * v.be.NUI + nkeep) * BE_WIN;
          const double* src = v.be.nui_win + ((size_t)s * v.be.NUI + k) * BE_WIN;
          for (int q = 0; q < BE_WIN; ++q) dst[q] = src[q];
          ncnt[nkeep] = ncnt[k];
        }
        newidx[k] = nkeep++;
      } else newidx[k] = -1;
    }
    if (nkeep != nnui)
      for (int i = 0; i < keep; ++i) {
        const size_t fi = (size_t)s * T + fs[i];
        const int as = v.be.ft_anchor[fi];
        if (as >= BE_NUI_BASE) v.be.ft_anchor[fi] = BE_NUI_BASE + newidx[as - BE_NUI_BASE];
      }
  }
  for (int i = 0; i < keep; ++i) {                               // updateGridMap (:3351-3370) over the features that stay
    const size_t fi = (size_t)s * T + fs[i];
    const double* o = v.be.ft_obs + (fi * Wcap + cur) * 4;
    int* cell = grid_cell(v, s, grid, grid_code(v, o[0], o[1]));
    if (cell) (*cell)++;
  }
  if (keep != nf || nkeep != nnui) { ic[I_REMAP] = 1; ic[I_NEWDIM] = base + idp * keep + 6 * nkeep; ic[I_NF] = keep; ic[I_NNUI] = nkeep; }
}

// ---------------------------------------------------------------- sequential part of the promotion rule (:1968-2002)
// Candidates = tracked, not in state, observed >= max_track_len times; processed in ascending feature id like the
// reference's std::map; the classify kernel left a speculative two-view-initialised triangulation in ft_spec.
__global__ void __launch_bounds__(256) be_slam_decide_kernel(BeView v) {
  __shared__ int unsorted[1024], sorted[1024];
  __shared__ int s_n;
  const int s = blockIdx.x, tid = threadIdx.x;
  int* ic = icore_of(v, s);
  if (tid == 0) { ic[I_NCAND] = 0; ic[I_NNEW] = 0; s_n = 0; }
  if (!ic[I_OK]) return;
  __syncthreads();
  const int T = v.be.T, Wcap = v.be.Wcap;
  const double* core = core_of(v, s);
  int* cand = v.be.cand + (size_t)s * 128;
  int* grid = cand + 64;
  // candidates (action 5) sorted by feature id: unordered gather, then rank by counting smaller ids
  for (int i = tid; i < T; i += blockDim.x)
    if (v.be.ft_action[(size_t)s * T + i] == 5) { const int k = atomicAdd(&s_n, 1); if (k < 1024) unsorted[k] = i; }
  __syncthreads();
  const int n = min(s_n, 1024);
  for (int k = tid; k < n; k += blockDim.x) {
    const unsigned long long id = v.be.ft_id[(size_t)s * T + unsorted[k]];
    int rank = 0;
    for (int q = 0; q < n; ++q) rank += v.be.ft_id[(size_t)s * T + unsorted[q]] < id;
    sorted[rank] = unsorted[k];
  }
  __syncthreads();
  if (tid != 0) return;
  const int* cand_sorted = sorted;
  const int cur = ic[I_NWIN] - 1;
  const bool zupt = ic[I_ZUPT] != 0;
  int n_new = 0;
  const int nf = ic[I_NF];
  int newlist[64];
  for (int k = 0; k < n; ++k) {
    const int slot = cand_sorted[k];
    const size_t fi = (size_t)s * T + slot;
    int flags = v.be.ft_flags[fi];
    const unsigned long long mask = v.be.ft_mask[fi];
    const double* o = v.be.ft_obs + (fi * Wcap + cur) * 4;
    const int code = grid_code(v, o[0], o[1]);
    const double* sp = v.be.ft_spec + fi * 8;
    const bool spec_ok = sp[0] != 0.0;
    int* cell = v.cfg.hybrid ? grid_cell(v, s, grid, code) : nullptr;
    if (v.cfg.hybrid && !cell) atomicExch(&ic[I_ERR], 5);
    const bool slam = cell && *cell < v.be.max_per_cell &&
                      core[C_TIME] - core[C_LAST_ZUPT] > 5 && (nf + n_new) < v.be.NFmax;
    auto commit_spec = [&](bool ekf) {
      double* pos = v.be.ft_pos + fi * 3;
      if (!(flags & 2)) { v.be.ft_pfej[fi * 3] = pos[0]; v.be.ft_pfej[fi * 3 + 1] = pos[1]; v.be.ft_pfej[fi * 3 + 2] = pos[2]; }
      pos[0] = sp[1]; pos[1] = sp[2]; pos[2] = sp[3];
      const double inv = 1.0 / sp[6];
      v.be.ft_inv[fi] = inv; v.be.ft_oa[fi * 2] = sp[4] * inv; v.be.ft_oa[fi * 2 + 1] = sp[5] * inv;
      v.be.ft_anchor[fi] = 63 - __clzll((long long)(mask & ~(1ull << cur)));
      flags |= 2;
      if (ekf) flags |= 8;
    };
    int action = 0; unsigned long long usemask = 0; int nrows = 0;
    const int m = __popcll(mask);
    if (slam) {
      if (!(flags & 8)) {               // not yet a potential EKF feature: re-initialise from scratch (:1977-1981)
        flags &= ~2;
        if (spec_ok) commit_spec(true);
      }
      if (flags & 2) {
        action = 3; usemask = mask; nrows = 2 * m + (v.be.IDP == 3 ? 2 * m : 2 * (m - 1));      // gate rows + featureJacobian_ekf_new rows (the anchor's own observation only counts with 3-D inverse depth, :1260-1262)
        (*cell)++; newlist[n_new++] = slot;
      }
    } else {
      if (!(flags & 2) && spec_ok) commit_spec(false);
      if (flags & 2) { action = 2; usemask = mask; nrows = 2 * m; }
    }
    if (zupt && action == 2) { flags &= ~2; usemask = 0; nrows = 0; }      // :2241-2246
    v.be.ft_flags[fi] = flags; v.be.ft_action[fi] = action; v.be.ft_usemask[fi] = usemask; v.be.ft_nrows[fi] = nrows;
  }
  for (int k = 0; k < n_new; ++k) cand[k] = newlist[k];     // new SLAM features in id order = their future state order
  ic[I_NCAND] = n_new;
}

}  // namespace

// ====================================================================== host side
template <typename T>
static int bdalloc(LvbHandle* h, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  LVB_CUDA(cudaMalloc(&q, bytes));
  LVB_CUDA(cudaMemsetAsync(q, 0, bytes, h->stream));
  h->allocs.push_back(q);
  *p = (T*)q;
  return LVB_OK;
}
#define BDA(ptr, n) do { int rc_ = bdalloc(h, &(ptr), (size_t)(n)); if (rc_ != LVB_OK) return rc_; } while (0)
#define BPIN(ptr, T, n) do { void* q_ = nullptr; LVB_CUDA(cudaHostAlloc(&q_, sizeof(T) * (size_t)(n), cudaHostAllocDefault)); memset(q_, 0, sizeof(T) * (size_t)(n)); ptr = (T*)q_; } while (0)

static const char* be_unsupported_reason(const LvbConfig& c) {
  if (c.max_features_in_one_grid > 0 && c.aug_grid_rows * c.aug_grid_cols != 0) {
    if (c.max_features_in_one_grid * c.aug_grid_rows * c.aug_grid_cols > 64) return "more than 64 EKF-SLAM features";
  }
  if (c.sw_size + 1 > 64) return "sw_size > 63";
  if (c.sw_size < 5) return "sw_size < 5 (findRedundantImuStates needs four older poses, larvio.cpp:2259-2307)";
  return nullptr;
}

__global__ void be_init_kernel(BeView v, double c_ori, double c_vel, double c_pos, double c_bg, double c_ba, double c_er,
                               double c_et, int est_ext, int est_td, double td, const double* T_cam_imu) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.be.S) return;
  double* core = core_of(v, s);
  int* ic = icore_of(v, s);
  for (int i = 0; i < BE_CORE; ++i) core[i] = 0.0;
  for (int i = 0; i < BE_ICORE; ++i) ic[i] = 0;
  core[C_Q + 3] = 1.0;
  core[C_TD] = td;
  for (int i = 0; i < 3; ++i) { core[C_TG + 4 * i] = 1.0; core[C_MA + 4 * i] = 1.0; }      // Ma = Tg = I, As = 0 (:129-131)
  // R_imu_cam0 = R_file, t_cam0_imu = -R_file^T t_file (larvio.cpp:189-203)
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) core[C_RIC + r * 3 + c] = T_cam_imu[r * 4 + c];
  for (int r = 0; r < 3; ++r) {
    double a = 0; for (int k = 0; k < 3; ++k) a += T_cam_imu[k * 4 + r] * T_cam_imu[k * 4 + 3];
    core[C_TCI + r] = -a;
  }
  ic[I_DIM] = LEGD;
  double* P = P_of(v, s);
  const int LD = v.be.LD;
  for (int i = 0; i < LEGD; ++i) for (int j = 0; j < LEGD; ++j) P[(size_t)i * LD + j] = 0.0;
  for (int i = 0; i < 3; ++i) {
    P[(size_t)i * LD + i] = c_ori; P[(size_t)(3 + i) * LD + 3 + i] = c_vel; P[(size_t)(6 + i) * LD + 6 + i] = c_pos;
    P[(size_t)(9 + i) * LD + 9 + i] = c_bg; P[(size_t)(12 + i) * LD + 12 + i] = c_ba;
    if (est_ext) { P[(size_t)(15 + i) * LD + 15 + i] = c_er; P[(size_t)(18 + i) * LD + 18 + i] = c_et; }
  }
  if (est_td) P[(size_t)21 * LD + 21] = 4e-6;
  for (int i = 22; i < LEGD; ++i) P[(size_t)i * LD + i] = 1e-4;                              // :183-186
}

static BeView make_beview(LvbHandle* h) {
  BeView v;
  v.be = *h->be;
  const LvbConfig& c = h->cfg;
  v.cfg.th_imu = 1.0 / (2.0 * c.imu_rate);
  v.cfg.sg2 = c.noise_gyro * c.noise_gyro; v.cfg.sa2 = c.noise_acc * c.noise_acc;
  v.cfg.sbg2 = c.noise_gyro_bias * c.noise_gyro_bias; v.cfg.sba2 = c.noise_acc_bias * c.noise_acc_bias;
  v.cfg.sfeat2 = c.noise_feature * c.noise_feature;
  v.cfg.rot_thr = c.rotation_threshold; v.cfg.trans_thr = c.translation_threshold; v.cfg.track_thr = c.tracking_rate_threshold;
  v.cfg.feat_trans_thr = c.feature_translation_threshold; v.cfg.zupt_dis = c.zupt_max_feature_dis;
  v.cfg.zupt_nv = c.zupt_noise_v * c.zupt_noise_v; v.cfg.zupt_np = c.zupt_noise_p * c.zupt_noise_p; v.cfg.zupt_nq = c.zupt_noise_q * c.zupt_noise_q;
  v.cfg.max_track_len = c.max_track_len; v.cfg.sw_size = c.sw_size; v.cfg.least_obs = c.least_observation_number;
  v.cfg.if_FEJ_config = c.if_FEJ; v.cfg.estimate_td = c.estimate_td; v.cfg.if_ZUPT_valid = c.if_ZUPT_valid;
  v.cfg.hybrid = h->be->NFmax > 0;
  {
    const double x_max = (c.width - c.cx) / c.fx, y_max = (c.height - c.cy) / c.fy;
    v.cfg.x_min = -c.cx / c.fx; v.cfg.y_min = -c.cy / c.fy;
    const int cells = c.aug_grid_rows * c.aug_grid_cols;
    v.cfg.grid_w = cells ? (x_max - v.cfg.x_min) / c.aug_grid_cols : (x_max - v.cfg.x_min);
    v.cfg.grid_h = cells ? (y_max - v.cfg.y_min) / c.aug_grid_rows : (y_max - v.cfg.y_min);
  }
  v.msg = nullptr; v.msg_n = nullptr; v.msg_t = nullptr; v.msg_valid = nullptr; v.msg_stride = 0;
  v.P = h->be->P[0]; v.Pn = h->be->P[1];
  return v;
}

int be_alloc(LvbHandle* h) {
  LvbBackEnd* be = new LvbBackEnd();
  memset(be, 0, sizeof(*be));
  h->be = be;
  const LvbConfig& c = h->cfg;
  be->S = h->S; be->N = h->fe.N;
  int sw = c.sw_size; if (sw < 5) sw = 5; if (sw > 63) sw = 63;
  be->Wcap = sw + 1;
  be->T = 2 * be->N;
  be->grid_rows = c.aug_grid_rows; be->grid_cols = c.aug_grid_cols; be->max_per_cell = c.max_features_in_one_grid > 0 ? c.max_features_in_one_grid : 0;
  be->NFmax = be->max_per_cell * be->grid_rows * be->grid_cols;
  if (be->NFmax > 64) be->NFmax = 0;                     // such configs are refused at the first back-end call
  be->NUI = (c.use_schmidt && be->NFmax > 0) ? BE_NUI_MAX : 0;
  be->IDP = (c.feature_idp_dim == 1) ? 1 : 3;           // anything but 1 means 3 (larvio.cpp:270-274)
  be->LEG = h->cfg.calib_imu_instrinsic ? 46 : 22;
  be->Dmax = be->LEG + 6 * be->Wcap + be->IDP * be->NFmax + 6 * be->NUI;
  be->LD = ((be->Dmax + 7) / 8) * 8;
  be->LDS = be->NFmax ? ((be->Dmax + 2 * be->NFmax + 16 * be->NFmax + 7) / 8) * 8 : be->LD;
  // row capacities of one measurement pass: every track that reaches max_track_len in the same frame contributes
  // 2m raw / 2m-3 stacked rows (all N tracks born in the first frame do so together); overflow is reported (LVB_E_CAPACITY)
  {
    const int mlen = c.max_track_len > 2 ? c.max_track_len : 2;
    const int raw = 2 * (mlen + 1) * be->N, stk = (2 * mlen - 3) * be->N + 32;
    be->RAWMAX = raw > 4096 ? ((raw + 255) / 256) * 256 : 4096;
    be->RMAX = stk > 2048 ? ((stk + 255) / 256) * 256 : 2048;
  }
  be->imu_cap = 64;
  be->PCAP = be->NFmax > 0 ? (8 * be->NFmax > 64 ? 8 * be->NFmax : 64) : 0;
  be->stats = h->fe.stats;
  const size_t S = be->S, T = be->T, LD = be->LD;
  BDA(be->core, S * BE_CORE); BDA(be->icore, S * BE_ICORE);
  BDA(be->kmap, S * LD);
  BDA(be->pts_id, 2 * S * be->PCAP); BDA(be->pts_xyz, 2 * S * be->PCAP * 3); BDA(be->pts_n, 2 * S); BDA(be->pts_drop, 2 * S);
  BDA(be->win_id, S * be->Wcap); BDA(be->win, S * be->Wcap * BE_WIN);
  BDA(be->P[0], S * LD * LD); BDA(be->P[1], 8);
  BDA(be->ft_id, S * T); BDA(be->ft_flags, S * T); BDA(be->ft_pos, S * T * 3); BDA(be->ft_mask, S * T);
  BDA(be->ft_obs, S * T * be->Wcap * 4);
  BDA(be->ft_action, S * T); BDA(be->ft_rowofs, S * T); BDA(be->ft_nrows, S * T); BDA(be->ft_accept, S * T); BDA(be->ft_usemask, S * T);
  BDA(be->Hraw, S * (size_t)be->RAWMAX * LD); BDA(be->rraw, S * (size_t)be->RAWMAX);
  BDA(be->Hs, S * LD * (size_t)be->RMAX); BDA(be->rs, S * (size_t)be->RMAX);
  BDA(be->Tm, S * (size_t)be->RAWMAX * LD);      // doubles as the per-feature H*P scratch of the gate
  BDA(be->Sm, S * (size_t)be->LDS * be->LDS); BDA(be->zvec, S * (size_t)be->LDS); BDA(be->dx, S * (size_t)be->LDS);
  BDA(be->ft_inv, S * T); BDA(be->ft_oa, S * T * 2); BDA(be->ft_anchor, S * T); BDA(be->ft_pfej, S * T * 3); BDA(be->ft_spec, S * T * 8);
  BDA(be->ft_gamma, S * T); BDA(be->fs_slot, S * 64); BDA(be->cmap, S * LD); BDA(be->cand, S * 128); BDA(be->grid_oor, S * BE_GRID_OOR);
  { const size_t nu = be->NUI > 0 ? be->NUI : 1; BDA(be->nui_win, S * nu * BE_WIN); BDA(be->nui_cnt, S * nu); BDA(be->nui_P, S * 36 * nu * nu); } BDA(be->Hnew, S * 64 * be->IDP * (LD + 4));
  BDA(be->imu, S * be->imu_cap); BDA(be->n_imu, S);
  BDA(be->msg_in, S * be->N); BDA(be->msg_in_n, S); BDA(be->msg_in_t, S); BDA(be->msg_in_valid, S);
  BPIN(be->pin_imu, LvbImu, S * be->imu_cap); BPIN(be->pin_n_imu, int, S); BPIN(be->pin_icore, int, S * BE_ICORE);
  BPIN(be->pin_feat, LvbFeature, S * be->N); BPIN(be->pin_feat_n, int, S); BPIN(be->pin_feat_t, double, S); BPIN(be->pin_valid, uint8_t, S);
  double* d_T = nullptr;
  BDA(d_T, 16);
  LVB_CUDA(cudaMemcpyAsync(d_T, c.T_cam_imu, sizeof(double) * 16, cudaMemcpyHostToDevice, h->stream));
  BeView v = make_beview(h);
  LVB_PROF(h, "be_init_kernel");
  be_init_kernel<<<(be->S + 63) / 64, 64, 0, h->stream>>>(v, c.cov_orientation, c.cov_velocity, c.cov_position, c.cov_gyro_bias,
                                                          c.cov_acc_bias, c.cov_extrin_rot, c.cov_extrin_trans, c.estimate_extrin,
                                                          c.estimate_td, c.td, d_T);
  LVB_LAUNCH_CHECK(h);
  // dynamic shared memory opt-ins
  const int fsm_bytes = (int)sizeof(double) * (6 * be->Wcap + 8 + 4 * be->Wcap * be->Wcap + 2 * be->Wcap + (8 + 6 * be->Wcap) / 2 + 4);
  LVB_CUDA(cudaFuncSetAttribute(be_feature_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm_bytes));
  LVB_CUDA(cudaFuncSetAttribute(be_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * QR_SMEM_DOUBLES)));
  LVB_CUDA(cudaFuncSetAttribute(be_propagate_kernel<22>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)be_propagate_smem(22)));
  LVB_CUDA(cudaFuncSetAttribute(be_propagate_kernel<46>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)be_propagate_smem(46)));
  // with the nonzero-column compression r stays near 50..110 rows whatever the window; the pure-MSCKF BASELINE window (LDS 208)
  // fits entirely, so be_chol_gmem_kernel is not even launched there
  be->chol_cap = be->LDS < 224 ? be->LDS : 224;          // 224 rows of packed triangle + 2 vectors = 205 KB of the 227 KB a CTA can have
  const size_t chol_bytes = sizeof(double) * ((size_t)be->chol_cap * (be->chol_cap + 1) / 2 + 2 * be->chol_cap);
  LVB_CUDA(cudaFuncSetAttribute(be_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_bytes));
  h->use_graph = getenv("LVB_NO_GRAPH") ? 0 : 1;          // one CUDA graph launch per step (lvb_step_graph); the env switch is for debugging
  return LVB_OK;
}

void be_free(LvbHandle* h) {
  if (!h->be) return;
  LvbBackEnd* be = h->be;
  void* pins[] = {be->pin_imu, be->pin_n_imu, be->pin_icore, be->pin_feat, be->pin_feat_n, be->pin_feat_t, be->pin_valid};
  for (void* p : pins) if (p) cudaFreeHost(p);
  delete be;
  h->be = nullptr;
}

static int launch_gemm(LvbHandle* h, const GemmArgs& g, int max_m, int max_n) {
  LvbBackEnd* be = h->be;
  LVB_PROF(h, "be_gemm_kernel");
  be_gemm_kernel<<<dim3((max_n + GT - 1) / GT, (max_m + GT - 1) / GT, be->S), 256, 0, h->stream>>>(g);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

// compression + EKF update on the stacked system currently in Hs/rs
static int be_debug_check(LvbHandle* h, const char* stage);
#define DBG(name) RC(be_debug_check(h, name))
static int be_colscan(LvbHandle* h, BeView& v, int rows_idx) {
  LVB_PROF(h, "be_colscan_kernel");
  be_colscan_kernel<<<h->be->S, 256, 0, h->stream>>>(v, rows_idx);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

static int be_qr(LvbHandle* h, BeView& v) {       // kmap / I_NC of the MSCKF block come from be_stack_kernel(phase 0)
  LVB_PROF(h, "be_qr_kernel");
  be_qr_kernel<<<h->be->S, 512, sizeof(double) * QR_SMEM_DOUBLES, h->stream>>>(v);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

// rescan: the stacked system changed since the last column scan (rows appended after the compression, or built directly)
static int be_update(LvbHandle* h, BeView& v, bool zupt_rows = false, bool rescan = false) {
  LvbBackEnd* be = h->be;
  if (zupt_rows || rescan) RC(be_colscan(h, v, I_R));
  cudaStream_t st = h->stream;
  if (be->NUI > 0) {                               // use_schmidt: the nuisance block keeps its prior (saved here, restored below)
    LVB_PROF(h, "be_nui_block_kernel");
    be_nui_block_kernel<<<be->S, 256, 0, st>>>(v, 0);
    LVB_LAUNCH_CHECK(h);
  }
  const size_t LD = be->LD;
  GemmArgs g;
  g.icore = be->icore; g.diag_vec = nullptr; g.sD = be->LDS;
  g.kmap = be->kmap; g.sK = LD;                     // T and S contract over the nonzero columns of H only (be_colscan_kernel)
  // T = H P
  g.A = be->Hs; g.sA = LD * be->RMAX; g.rsA = 1; g.csA = be->RMAX;
  g.B = v.P; g.sB = LD * LD; g.rsB = (int)LD; g.csB = 1;
  g.C = be->Tm; g.sC = (size_t)be->RAWMAX * LD; g.rsC = (int)LD; g.csC = 1;
  g.m_idx = I_R; g.n_idx = I_DIM; g.k_idx = I_NC; g.alpha = 1.0; g.beta = 0.0; g.diag = 0.0;
  RC(launch_gemm(h, g, be->LDS, be->Dmax));
  DBG("gemm T=HP");
  // S = T H^T + sigma^2 I
  g.A = be->Tm; g.sA = (size_t)be->RAWMAX * LD; g.rsA = (int)LD; g.csA = 1;
  g.B = be->Hs; g.sB = LD * be->RMAX; g.rsB = be->RMAX; g.csB = 1;
  g.C = be->Sm; g.sC = (size_t)be->LDS * be->LDS; g.rsC = be->LDS; g.csC = 1;
  g.m_idx = I_R; g.n_idx = I_R; g.k_idx = I_NC; g.alpha = 1.0; g.beta = 0.0; g.diag = v.cfg.sfeat2;
  if (zupt_rows) g.diag_vec = be->dx;
  RC(launch_gemm(h, g, be->LDS, be->LDS));
  g.diag_vec = nullptr;
  const size_t chol_bytes = sizeof(double) * ((size_t)be->chol_cap * (be->chol_cap + 1) / 2 + 2 * be->chol_cap);
  LVB_PROF(h, "be_chol_kernel");
  be_chol_kernel<<<be->S, 1024, chol_bytes, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  DBG("be_chol_kernel");
  if (be->LDS > be->chol_cap) {                  // innovation systems beyond the shared-memory capacity (rare: > 224 rows)
    LVB_PROF(h, "be_chol_gmem_kernel");
    be_chol_gmem_kernel<<<be->S, 512, sizeof(double) * be->LDS, st>>>(v, be->chol_cap);
    LVB_LAUNCH_CHECK(h);
    DBG("be_chol_gmem_kernel");
  }
  LVB_PROF(h, "be_trsm_kernel");
  be_trsm_kernel<<<dim3((be->Dmax + 63) / 64, be->S), 64, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  DBG("be_trsm_kernel");
  LVB_PROF(h, "be_correct_kernel");
  be_correct_kernel<<<be->S, 256, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  DBG("be_correct_kernel");
  // P -= Y^T Y
  g.kmap = nullptr;
  g.A = be->Tm; g.sA = (size_t)be->RAWMAX * LD; g.rsA = 1; g.csA = (int)LD;
  g.B = be->Tm; g.sB = (size_t)be->RAWMAX * LD; g.rsB = (int)LD; g.csB = 1;
  g.C = v.P; g.sC = LD * LD; g.rsC = (int)LD; g.csC = 1;
  g.m_idx = I_DIM; g.n_idx = I_DIM; g.k_idx = I_R; g.alpha = -1.0; g.beta = 1.0; g.diag = 0.0;
  RC(launch_gemm(h, g, be->Dmax, be->Dmax));
  DBG("gemm P-=YtY");
  if (be->NUI > 0) {
    LVB_PROF(h, "be_nui_block_kernel");
    be_nui_block_kernel<<<be->S, 256, 0, st>>>(v, 1);
    LVB_LAUNCH_CHECK(h);
  }
  return LVB_OK;
}

// LVB_DEBUG_NAN=1: after every back-end launch, pull P / H / T / S of every sequence and report the first non-finite value
static int be_debug_check(LvbHandle* h, const char* stage) {
  if (!getenv("LVB_DEBUG_NAN")) return LVB_OK;
  LvbBackEnd* be = h->be;
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  std::vector<int> ic((size_t)be->S * BE_ICORE);
  LVB_CUDA(cudaMemcpy(ic.data(), be->icore, sizeof(int) * ic.size(), cudaMemcpyDeviceToHost));
  std::vector<double> P((size_t)be->LD * be->LD);
  for (int s = 0; s < be->S; ++s) {
    const int* c = ic.data() + (size_t)s * BE_ICORE;
    if (!c[I_OK]) continue;
    LVB_CUDA(cudaMemcpy(P.data(), be->P[0] + (size_t)s * be->LD * be->LD, sizeof(double) * P.size(), cudaMemcpyDeviceToHost));
    const int d = c[I_DIM];
    int bad = 0;
    for (int i = 0; i < d && !bad; ++i) for (int j = 0; j < d; ++j) if (!std::isfinite(P[(size_t)i * be->LD + j])) { bad = 1; fprintf(stderr, "[lvb debug] %s: seq %d P[%d][%d] not finite (d=%d nwin=%d nf=%d R=%d rows=%d nnew=%d)\n", stage, s, i, j, d, c[I_NWIN], c[I_NF], c[I_R], c[I_ROWS], c[I_NNEW]); break; }
    if (!bad) fprintf(stderr, "[lvb debug] %s: seq %d ok (d=%d nwin=%d nf=%d R=%d rows=%d nnew=%d ncand=%d)\n", stage, s, d, c[I_NWIN], c[I_NF], c[I_R], c[I_ROWS], c[I_NNEW], c[I_NCAND]);
    if (getenv("LVB_DEBUG_FEAT") && !strcmp(stage, "be_feature_kernel")) {
      const int T = be->T;
      std::vector<unsigned long long> ids(T); std::vector<int> act(T), acc(T), nr(T); std::vector<double> gm(T);
      cudaMemcpy(ids.data(), be->ft_id + (size_t)s * T, sizeof(unsigned long long) * T, cudaMemcpyDeviceToHost);
      cudaMemcpy(act.data(), be->ft_action + (size_t)s * T, sizeof(int) * T, cudaMemcpyDeviceToHost);
      cudaMemcpy(acc.data(), be->ft_accept + (size_t)s * T, sizeof(int) * T, cudaMemcpyDeviceToHost);
      cudaMemcpy(nr.data(), be->ft_nrows + (size_t)s * T, sizeof(int) * T, cudaMemcpyDeviceToHost);
      cudaMemcpy(gm.data(), be->ft_gamma + (size_t)s * T, sizeof(double) * T, cudaMemcpyDeviceToHost);
      for (int i = 0; i < T; ++i) if (act[i] >= 2) fprintf(stderr, "[lvb feat] seq %d id %llu action %d nrows %d accept %d gamma %.9e\n", s, ids[i], act[i], nr[i], acc[i], gm[i]);
    }
    // stacked system, T, S, z, dx, Hnew
    const int r = c[I_R];
    if (r > 0) {
      std::vector<double> Hs((size_t)be->LD * be->RMAX), Tm((size_t)r * be->LD), Sm((size_t)be->LDS * be->LDS), zz(be->LDS), dxx(be->LDS), rs(be->RMAX);
      cudaMemcpy(Hs.data(), be->Hs + (size_t)s * be->LD * be->RMAX, sizeof(double) * Hs.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(rs.data(), be->rs + (size_t)s * be->RMAX, sizeof(double) * rs.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(Tm.data(), be->Tm + (size_t)s * be->RAWMAX * be->LD, sizeof(double) * Tm.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(Sm.data(), be->Sm + (size_t)s * be->LDS * be->LDS, sizeof(double) * Sm.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(zz.data(), be->zvec + (size_t)s * be->LDS, sizeof(double) * zz.size(), cudaMemcpyDeviceToHost);
      cudaMemcpy(dxx.data(), be->dx + (size_t)s * be->LDS, sizeof(double) * dxx.size(), cudaMemcpyDeviceToHost);
      int bh = -1, bt = -1, bs = -1, bz = -1, bd = -1, br = -1; double mindiag = 1e300;
      for (int i = 0; i < r; ++i) {
        for (int j = 0; j < d; ++j) { if (bh < 0 && !std::isfinite(Hs[(size_t)j * be->RMAX + i])) bh = i; if (bt < 0 && !std::isfinite(Tm[(size_t)i * be->LD + j])) bt = i; }
        for (int j = 0; j <= i; ++j) if (bs < 0 && !std::isfinite(Sm[(size_t)i * be->LDS + j])) bs = i;
        if (std::isfinite(Sm[(size_t)i * be->LDS + i])) mindiag = std::min(mindiag, Sm[(size_t)i * be->LDS + i]);
        if (bz < 0 && !std::isfinite(zz[i])) bz = i;
        if (br < 0 && !std::isfinite(rs[i])) br = i;
      }
      for (int j = 0; j < d; ++j) if (bd < 0 && !std::isfinite(dxx[j])) bd = j;
      fprintf(stderr, "[lvb debug]     seq %d first non-finite row: H %d r %d T %d S %d z %d dx %d ; min diag(S) %.3e\n", s, bh, br, bt, bs, bz, bd, mindiag);
    }
  }
  return LVB_OK;
}

static int be_remap(LvbHandle* h, BeView& v) {
  LvbBackEnd* be = h->be;
  LVB_PROF(h, "be_remap_gather_kernel");
  be_remap_gather_kernel<<<dim3(be->Dmax, be->S), 256, 0, h->stream>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_remap_scatter_kernel");
  be_remap_scatter_kernel<<<dim3(be->Dmax, be->S), 256, 0, h->stream>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_remap_commit_kernel");
  be_remap_commit_kernel<<<(be->S + 63) / 64, 64, 0, h->stream>>>(v);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

static int be_measurement_pass(LvbHandle* h, BeView& v, int mode) {
  LvbBackEnd* be = h->be;
  cudaStream_t st = h->stream;
  const bool hybrid = be->NFmax > 0;
  if (hybrid && mode == 0) {
    LVB_PROF(h, "be_slam_pre_kernel");
    be_slam_pre_kernel<<<be->S, 64, 0, st>>>(v);
    LVB_LAUNCH_CHECK(h);
    DBG("be_slam_pre_kernel");
    RC(be_remap(h, v));
  }
  if (hybrid && mode == 1) {
    LVB_PROF(h, "be_anchor_kernel");
    be_anchor_kernel<<<be->S, 256, sizeof(double) * be->IDP * be->LD, st>>>(v);
    LVB_LAUNCH_CHECK(h);
    DBG("be_anchor_kernel");
  }
  LVB_PROF(h, "be_classify_kernel");
  be_classify_kernel<<<dim3((be->T + 3) / 4, be->S), 128, 0, st>>>(v, mode);
  LVB_LAUNCH_CHECK(h);
  DBG("be_classify_kernel");
  if (hybrid && mode == 0) {
    LVB_PROF(h, "be_slam_decide_kernel");
    be_slam_decide_kernel<<<be->S, 256, 0, st>>>(v);
    LVB_LAUNCH_CHECK(h);
    DBG("be_slam_decide_kernel");
  }
  LVB_PROF(h, "be_scan_rows_kernel");
  be_scan_rows_kernel<<<be->S, 512, 0, st>>>(v, 0);
  LVB_LAUNCH_CHECK(h);
  DBG("be_scan_rows_kernel");
  const int fsm_bytes = (int)sizeof(double) * (6 * be->Wcap + 8 + 4 * be->Wcap * be->Wcap + 2 * be->Wcap + (8 + 6 * be->Wcap) / 2 + 4);
  LVB_PROF(h, "be_feature_kernel");
  be_feature_kernel<<<dim3(be->T, be->S), 32, fsm_bytes, st>>>(v, be->Tm);
  LVB_LAUNCH_CHECK(h);
  DBG("be_feature_kernel");
  if (hybrid && mode == 0) {
    LVB_PROF(h, "be_slam_accept_kernel");
    be_slam_accept_kernel<<<(be->S + 31) / 32, 32, 0, st>>>(v);
    LVB_LAUNCH_CHECK(h);
    DBG("be_slam_accept_kernel");
  }
  LVB_PROF(h, "be_stack_kernel");
  be_stack_kernel<<<be->S, 512, 0, st>>>(v, 0);
  LVB_LAUNCH_CHECK(h);
  DBG("be_stack_kernel");
  RC(be_qr(h, v));
  if (hybrid && mode == 0) {
    LVB_PROF(h, "be_stack_kernel");
    be_stack_kernel<<<be->S, 512, 0, st>>>(v, 1);
    LVB_LAUNCH_CHECK(h);
    DBG("be_stack_kernel");
  }
  RC(be_update(h, v, false, hybrid && mode == 0));
  if (hybrid && mode == 0) {
    LVB_PROF(h, "be_slam_grow_kernel");
    be_slam_grow_kernel<<<be->S, 256, 0, st>>>(v);
    LVB_LAUNCH_CHECK(h);
    DBG("be_slam_grow_kernel");
    if (be->NUI > 0) RC(be_remap(h, v));             // the nuisance block moves behind the new feature columns (:1832-1845)
  }
  return LVB_OK;
}

// processFeatures for the whole batch in three pieces, shared by the plain path (be_process) and the one-graph-per-step
// path (lvb_step_graph): host preparation (pinned staging of the IMU buffers), a capturable enqueue (fixed launch sequence:
// grids follow capacities, every per-sequence decision is a device-side flag) and the read-back.
//
// The WHOLE caller buffer of every sequence is staged, as batchImuProcessing sees the whole buffer (larvio.cpp:464-512): the
// staging capacity grows with the largest pending count (rounded up to 64) instead of truncating.
static int be_grow_imu(LvbHandle* h, int need) {
  LvbBackEnd* be = h->be;
  const int cap = ((need + 63) / 64) * 64;
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  LvbImu* d = nullptr; void* pin = nullptr;
  LVB_CUDA(cudaMalloc((void**)&d, sizeof(LvbImu) * (size_t)be->S * cap));
  LVB_CUDA(cudaHostAlloc(&pin, sizeof(LvbImu) * (size_t)be->S * cap, cudaHostAllocDefault));
  h->allocs.push_back(d);                       // the old device buffer stays on the handle's free list until lvb_destroy
  if (be->pin_imu) cudaFreeHost(be->pin_imu);
  be->imu = d; be->pin_imu = (LvbImu*)pin; be->imu_cap = cap;
  for (int k = 0; k < 2; ++k) if (h->gexec[k]) { cudaGraphExecDestroy(h->gexec[k]); h->gexec[k] = nullptr; }   // captured pointers are stale
  return LVB_OK;
}

static int be_host_prep(LvbHandle* h, const LvbImu* imu, const int* n_imu, int imu_stride) {
  LvbBackEnd* be = h->be;
  const int S = be->S;
  int need = 0;
  for (int s = 0; s < S; ++s) {
    if (n_imu[s] < 0 || n_imu[s] > imu_stride) return lvb_set_err(LVB_E_ARG, "sequence %d: n_imu %d outside [0, imu_stride %d]", s, n_imu[s], imu_stride);
    if (n_imu[s] > need) need = n_imu[s];
  }
  if (need > be->imu_cap) RC(be_grow_imu(h, need));
  for (int s = 0; s < S; ++s) {
    be->pin_n_imu[s] = n_imu[s];
    memcpy(be->pin_imu + (size_t)s * be->imu_cap, imu + (size_t)s * imu_stride, sizeof(LvbImu) * n_imu[s]);
  }
  return LVB_OK;
}

// msg arrays are device pointers
static int be_enqueue(LvbHandle* h, const LvbFeature* d_msg, const int* d_msg_n, const double* d_msg_t, const uint8_t* d_valid, int msg_stride) {
  LvbBackEnd* be = h->be;
  cudaStream_t st = h->stream;
  const int S = be->S;
  LVB_CUDA(cudaMemcpyAsync(be->imu, be->pin_imu, sizeof(LvbImu) * (size_t)S * be->imu_cap, cudaMemcpyHostToDevice, st));
  LVB_CUDA(cudaMemcpyAsync(be->n_imu, be->pin_n_imu, sizeof(int) * S, cudaMemcpyHostToDevice, st));
  BeView v = make_beview(h);
  v.msg = d_msg; v.msg_n = d_msg_n; v.msg_t = d_msg_t; v.msg_valid = d_valid; v.msg_stride = msg_stride;
  LVB_PROF(h, "be_propagate_kernel");
  if (be->LEG == 22) be_propagate_kernel<22><<<S, 256, be_propagate_smem(22), st>>>(v);
  else be_propagate_kernel<46><<<S, 256, be_propagate_smem(46), st>>>(v);
  LVB_LAUNCH_CHECK(h);
  const int nthr = be->N <= 256 ? 256 : 512;
  LVB_PROF(h, "be_add_obs_kernel");
  be_add_obs_kernel<<<S, nthr, sizeof(unsigned long long) * be->T, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_augment_kernel");
  be_augment_kernel<<<S, 256, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  if (be->NFmax > 0) RC(be_remap(h, v));
  if (h->cfg.if_ZUPT_valid) {                        // checkZUPT -> measurementUpdate_ZUPT_vpq (:405-406)
    LVB_PROF(h, "be_zupt_build_kernel");
    be_zupt_build_kernel<<<S, 256, 0, st>>>(v);
    LVB_LAUNCH_CHECK(h);
    RC(be_update(h, v, true));
  }
  RC(be_measurement_pass(h, v, 0));                  // removeLostFeatures
  LVB_PROF(h, "be_apply_actions_kernel");
  be_apply_actions_kernel<<<dim3((be->T + 127) / 128, S), 128, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_prune_select_kernel");
  be_prune_select_kernel<<<(S + 63) / 64, 64, 0, st>>>(v);   // pruneImuStateBuffer
  LVB_LAUNCH_CHECK(h);
  RC(be_measurement_pass(h, v, 1));
  LVB_PROF(h, "be_prune_tables_kernel");
  be_prune_tables_kernel<<<S, 256, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_prune_cov_gather_kernel");
  be_prune_cov_gather_kernel<<<dim3(be->Dmax, S), 256, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_prune_cov_scatter_kernel");
  be_prune_cov_scatter_kernel<<<dim3(be->Dmax, S), 256, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_PROF(h, "be_frame_end_kernel");
  be_frame_end_kernel<<<S, 32, 0, st>>>(v);
  LVB_LAUNCH_CHECK(h);
  LVB_CUDA(cudaMemcpyAsync(be->pin_icore, be->icore, sizeof(int) * (size_t)S * BE_ICORE, cudaMemcpyDeviceToHost, st));
  return LVB_OK;
}

// imu/n_imu: the caller's host buffers (mutated like the reference's)
static int be_finish(LvbHandle* h, LvbImu* imu, int* n_imu, int imu_stride, uint8_t* ok_out) {
  LvbBackEnd* be = h->be;
  const int S = be->S;
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  int err = 0;
  for (int s = 0; s < S; ++s) {
    const int* ic = be->pin_icore + (size_t)s * BE_ICORE;
    if (ok_out) ok_out[s] = (uint8_t)ic[I_OK];
    if (ic[I_OK]) {
      const int used = ic[I_CONSUMED];
      if (used > 0) {                                  // larvio.cpp:510-512: erase consumed samples in place
        LvbImu* b = imu + (size_t)s * imu_stride;
        memmove(b, b + used, sizeof(LvbImu) * (n_imu[s] - used));
        n_imu[s] -= used;
      }
      if (ic[I_ERR]) err = ic[I_ERR];
    }
  }
  if (err) return lvb_set_err(LVB_E_CAPACITY, "back-end capacity exceeded (code %d: 1 feature table, 2 window, 3 raw rows, 4 stacked rows, 5 grid cell code outside [-64, 192), 6 nuisance states: more than 16, or a ZUPT while some exist)", err);
  return LVB_OK;
}

static int be_check_config(LvbHandle* h) {
  if (const char* why = be_unsupported_reason(h->cfg)) return lvb_set_err(LVB_E_UNSUPPORTED, "%s", why);
  if (h->be->N > 512) return lvb_set_err(LVB_E_UNSUPPORTED, "max_features_num > 512");
  return LVB_OK;
}

int be_process(LvbHandle* h, const LvbFeature* d_msg, const int* d_msg_n, const double* d_msg_t, const uint8_t* d_valid,
               int msg_stride, LvbImu* imu, int* n_imu, int imu_stride, uint8_t* ok_out) {
  RC(be_check_config(h));
  RC(be_host_prep(h, imu, n_imu, imu_stride));
  RC(be_enqueue(h, d_msg, d_msg_n, d_msg_t, d_valid, msg_stride));
  return be_finish(h, imu, n_imu, imu_stride, ok_out);
}

// ---- ONE graph launch per step (default; LVB_NO_GRAPH=1 falls back to stream launches).  The enqueue halves of processImage / processFeatures are captured
// once per pyramid parity and replayed; host preparation (pinned staging of IMU, homographies, stamps; the image copy into
// the fixed staging batch) and the read-back stay outside.  Not used while the profiler or LVB_DEBUG_NAN is on (both
// synchronise between launches).
static int lvb_step_graph(LvbHandle* h, const uint8_t* images, int images_on_device, const double* t_img, LvbImu* imu,
                          int* n_imu, int imu_stride, uint8_t* published) {
  RC(be_check_config(h));
  cudaStream_t st = h->stream;
  const uint8_t* d_images = nullptr;
  RC(fe_host_prep(h, images, images_on_device, t_img, imu, n_imu, imu_stride, true, &d_images));
  RC(be_host_prep(h, imu, n_imu, imu_stride));
  const int par = h->fe.cur & 1;
  if (!h->gexec[par]) {
    cudaGraph_t graph = nullptr;
    const long long l0 = h->launches;
    LVB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = fe_enqueue(h, d_images);
    if (rc == LVB_OK) rc = be_enqueue(h, h->fe.msg, h->fe.msg_n, h->fe.msg_t, h->fe.has_msg, h->fe.N);
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != LVB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return lvb_set_err(LVB_E_CUDA, "cudaStreamEndCapture -> %s", cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&h->gexec[par], graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) return lvb_set_err(LVB_E_CUDA, "cudaGraphInstantiate -> %s", cudaGetErrorString(e2));
    h->glaunches[par] = h->launches - l0;
  } else {
    h->launches += h->glaunches[par];
  }
  LVB_CUDA(cudaGraphLaunch(h->gexec[par], st));
  h->fe.cur ^= 1;
  return be_finish(h, imu, n_imu, imu_stride, published);
}

extern "C" int lvb_process_features(LvbHandle* h, const uint8_t* valid, const double* t_msg, const LvbFeature* feat,
                                    const int* n_feat, int feat_stride, LvbImu* imu, int* n_imu, int imu_stride,
                                    uint8_t* ok) {
  if (!h || !valid || !t_msg || !feat || !n_feat || !imu || !n_imu) return lvb_set_err(LVB_E_ARG, "lvb_process_features: null argument");
  LVB_CUDA(cudaSetDevice(h->device));
  LvbBackEnd* be = h->be;
  const int S = be->S, N = be->N;
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  for (int s = 0; s < S; ++s) {
    int n = valid[s] ? n_feat[s] : 0;
    if (n > N) return lvb_set_err(LVB_E_CAPACITY, "feature message of %d entries exceeds capacity %d", n, N);
    be->pin_feat_n[s] = n; be->pin_feat_t[s] = t_msg[s]; be->pin_valid[s] = valid[s];
    if (n) memcpy(be->pin_feat + (size_t)s * N, feat + (size_t)s * feat_stride, sizeof(LvbFeature) * n);
  }
  LVB_CUDA(cudaMemcpyAsync(be->msg_in, be->pin_feat, sizeof(LvbFeature) * (size_t)S * N, cudaMemcpyHostToDevice, h->stream));
  LVB_CUDA(cudaMemcpyAsync(be->msg_in_n, be->pin_feat_n, sizeof(int) * S, cudaMemcpyHostToDevice, h->stream));
  LVB_CUDA(cudaMemcpyAsync(be->msg_in_t, be->pin_feat_t, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
  LVB_CUDA(cudaMemcpyAsync(be->msg_in_valid, be->pin_valid, S, cudaMemcpyHostToDevice, h->stream));
  return be_process(h, be->msg_in, be->msg_in_n, be->msg_in_t, be->msg_in_valid, N, imu, n_imu, imu_stride, ok);
}

extern "C" int lvb_step(LvbHandle* h, const uint8_t* images, int images_on_device, const double* t_img, LvbImu* imu,
                        int* n_imu, int imu_stride, uint8_t* published) {
  if (!h || !images || !t_img || !imu || !n_imu) return lvb_set_err(LVB_E_ARG, "lvb_step: null argument");
  LVB_CUDA(cudaSetDevice(h->device));
  if (h->use_graph && !h->prof.on && !getenv("LVB_DEBUG_NAN"))
    return lvb_step_graph(h, images, images_on_device, t_img, imu, n_imu, imu_stride, published);
  RC(fe_process(h, images, images_on_device, t_img, imu, n_imu, imu_stride));
  LvbFrontEnd& fe = h->fe;
  // the message stays in HBM: fe.msg / msg_n / msg_t / has_msg feed processFeatures directly
  return be_process(h, fe.msg, fe.msg_n, fe.msg_t, fe.has_msg, fe.N, imu, n_imu, imu_stride, published);
}

__global__ void be_set_state_kernel(BeView v, int s, double t, const double* vals /*q4 p3 v3 bg3 ba3*/) {
  double* core = core_of(v, s);
  int* ic = icore_of(v, s);
  core[C_TIME] = t;
  for (int i = 0; i < 4; ++i) core[C_Q + i] = vals[i];
  for (int i = 0; i < 3; ++i) {
    core[C_P + i] = vals[4 + i]; core[C_V + i] = vals[7 + i]; core[C_BG + i] = vals[10 + i]; core[C_BA + i] = vals[13 + i];
    core[C_FNOW_P + i] = vals[4 + i]; core[C_FNOW_V + i] = vals[7 + i];
  }
  core[C_TAKEOFF] = t;
  core[C_LAST_ZUPT] = t;
  ic[I_GRAVITY] = 1;
  // the initialiser only ever runs behind the bFirstFeatures gate (larvio.cpp:366-376), and the call that initialises goes on
  // to batchImuProcessing with the samples the initialiser left (StaticInitializer.cpp:149-150): a state handed in means the
  // gate has been passed, it must not be re-evaluated on the shortened buffer
  ic[I_FIRST] = 1;
}

extern "C" int lvb_set_initial_state(LvbHandle* h, int seq, double t, const double* q_xyzw, const double* p,
                                     const double* vel, const double* bg, const double* ba) {
  if (!h || seq < 0 || seq >= h->S || !q_xyzw || !p || !vel || !bg || !ba) return lvb_set_err(LVB_E_ARG, "lvb_set_initial_state: bad argument");
  LVB_CUDA(cudaSetDevice(h->device));
  double vals[16];
  for (int i = 0; i < 4; ++i) vals[i] = q_xyzw[i];
  for (int i = 0; i < 3; ++i) { vals[4 + i] = p[i]; vals[7 + i] = vel[i]; vals[10 + i] = bg[i]; vals[13 + i] = ba[i]; }
  double* d = h->be->dx;   // scratch
  LVB_CUDA(cudaMemcpyAsync(d, vals, sizeof(vals), cudaMemcpyHostToDevice, h->stream));
  BeView v = make_beview(h);
  LVB_PROF(h, "be_set_state_kernel");
  be_set_state_kernel<<<1, 1, 0, h->stream>>>(v, seq, t, d);
  LVB_LAUNCH_CHECK(h);
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  return LVB_OK;
}

extern "C" int lvb_get_states(LvbHandle* h, double* out) {
  if (!h || !out) return lvb_set_err(LVB_E_ARG, "lvb_get_states: null argument");
  LVB_CUDA(cudaSetDevice(h->device));
  std::vector<double> core((size_t)h->S * BE_CORE);
  LVB_CUDA(cudaMemcpyAsync(core.data(), h->be->core, sizeof(double) * core.size(), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  for (int s = 0; s < h->S; ++s) {
    const double* c = core.data() + (size_t)s * BE_CORE;
    double* o = out + (size_t)s * 17;
    o[0] = c[C_TIME];
    for (int i = 0; i < 4; ++i) o[1 + i] = c[C_Q + i];
    for (int i = 0; i < 3; ++i) { o[5 + i] = c[C_P + i]; o[8 + i] = c[C_V + i]; o[11 + i] = c[C_BG + i]; o[14 + i] = c[C_BA + i]; }
  }
  return LVB_OK;
}

extern "C" int lvb_get_calibration(LvbHandle* h, int seq, double* R_imu_cam9, double* t_cam_imu3, double* td,
                                   double* Tg9, double* As9, double* Ma9) {
  if (!h || seq < 0 || seq >= h->S) return lvb_set_err(LVB_E_ARG, "lvb_get_calibration: bad argument");
  LVB_CUDA(cudaSetDevice(h->device));
  double c[BE_CORE];
  LVB_CUDA(cudaMemcpyAsync(c, h->be->core + (size_t)seq * BE_CORE, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 9; ++i) {
    if (R_imu_cam9) R_imu_cam9[i] = c[C_RIC + i];
    if (Tg9) Tg9[i] = c[C_TG + i];
    if (As9) As9[i] = c[C_AS + i];
    if (Ma9) Ma9[i] = c[C_MA + i];
  }
  for (int i = 0; i < 3; ++i) if (t_cam_imu3) t_cam_imu3[i] = c[C_TCI + i];
  if (td) *td = c[C_TD];
  return LVB_OK;
}

extern "C" int lvb_get_covariance(LvbHandle* h, int seq, double* P, int cap_dim, int* dim) {
  if (!h || seq < 0 || seq >= h->S || !P || !dim) return lvb_set_err(LVB_E_ARG, "lvb_get_covariance: bad argument");
  LVB_CUDA(cudaSetDevice(h->device));
  LvbBackEnd* be = h->be;
  int ic[BE_ICORE];
  LVB_CUDA(cudaMemcpyAsync(ic, be->icore + (size_t)seq * BE_ICORE, sizeof(ic), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  const int d = ic[I_DIM];
  *dim = d;
  if (d > cap_dim) return lvb_set_err(LVB_E_CAPACITY, "covariance is %d x %d, caller capacity %d", d, d, cap_dim);
  std::vector<double> tmp((size_t)be->LD * be->LD);
  LVB_CUDA(cudaMemcpyAsync(tmp.data(), be->P[0] + (size_t)seq * be->LD * be->LD, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) P[(size_t)i * d + j] = tmp[(size_t)i * be->LD + j];
  return LVB_OK;
}

extern "C" int lvb_get_state(LvbHandle* h, int seq, double* t, double* q, double* p, double* vel, double* bg, double* ba,
                             double* P_pose36, double* P_vel9) {
  if (!h || seq < 0 || seq >= h->S) return lvb_set_err(LVB_E_ARG, "lvb_get_state: bad argument");
  LVB_CUDA(cudaSetDevice(h->device));
  LvbBackEnd* be = h->be;
  double c[BE_CORE];
  LVB_CUDA(cudaMemcpyAsync(c, be->core + (size_t)seq * BE_CORE, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  std::vector<double> rows((size_t)9 * be->LD);
  LVB_CUDA(cudaMemcpyAsync(rows.data(), be->P[0] + (size_t)seq * be->LD * be->LD, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  if (t) *t = c[C_TIME];
  for (int i = 0; i < 4; ++i) if (q) q[i] = c[C_Q + i];
  for (int i = 0; i < 3; ++i) { if (p) p[i] = c[C_P + i]; if (vel) vel[i] = c[C_V + i]; if (bg) bg[i] = c[C_BG + i]; if (ba) ba[i] = c[C_BA + i]; }
  // getPpose (larvio.cpp: rows/cols {0-2, 6-8}), getPvel (3-5)
  const int sel[6] = {0, 1, 2, 6, 7, 8};
  if (P_pose36) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) P_pose36[i * 6 + j] = rows[(size_t)sel[i] * be->LD + sel[j]];
  if (P_vel9) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) P_vel9[i * 3 + j] = rows[(size_t)(3 + i) * be->LD + 3 + j];
  return LVB_OK;
}

extern "C" int lvb_get_points(LvbHandle* h, int seq, int which, unsigned long long* ids, double* xyz, int cap, int* n) {
  if (!h || !n || seq < 0 || seq >= h->S || (which != 0 && which != 1) || cap < 0) return lvb_set_err(LVB_E_ARG, "lvb_get_points: bad argument");
  LVB_CUDA(cudaSetDevice(h->device));
  LvbBackEnd* be = h->be;
  *n = 0;
  if (be->PCAP == 0) return LVB_OK;                       // pure MSCKF: there are no EKF-SLAM features
  const size_t b = (size_t)which * be->S + seq;
  int cnt[2] = {0, 0};
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  LVB_CUDA(cudaMemcpy(&cnt[0], be->pts_n + b, sizeof(int), cudaMemcpyDeviceToHost));
  LVB_CUDA(cudaMemcpy(&cnt[1], be->pts_drop + b, sizeof(int), cudaMemcpyDeviceToHost));
  if (cnt[0] > cap) return lvb_set_err(LVB_E_CAPACITY, "lvb_get_points: %d points pending, caller capacity %d (nothing was cleared)", cnt[0], cap);
  if (cnt[0] > 0 && (!ids || !xyz)) return lvb_set_err(LVB_E_ARG, "lvb_get_points: null output");
  if (cnt[0] > 0) {
    LVB_CUDA(cudaMemcpy(ids, be->pts_id + b * be->PCAP, sizeof(unsigned long long) * cnt[0], cudaMemcpyDeviceToHost));
    LVB_CUDA(cudaMemcpy(xyz, be->pts_xyz + b * be->PCAP * 3, sizeof(double) * 3 * cnt[0], cudaMemcpyDeviceToHost));
  }
  LVB_CUDA(cudaMemset(be->pts_n + b, 0, sizeof(int)));   // the reference's getters clear their map (larvio.cpp:2722, 2729)
  LVB_CUDA(cudaMemset(be->pts_drop + b, 0, sizeof(int)));
  *n = cnt[0];
  if (cnt[1] > 0) return lvb_set_err(LVB_E_CAPACITY, "lvb_get_points: %d map points were dropped since the last read (list capacity %d); read more often", cnt[1], be->PCAP);
  return LVB_OK;
}

extern "C" int lvb_get_window(LvbHandle* h, int seq, double* qp, int cap, int* n) {
  if (!h || seq < 0 || seq >= h->S || !qp || !n) return lvb_set_err(LVB_E_ARG, "lvb_get_window: bad argument");
  LVB_CUDA(cudaSetDevice(h->device));
  LvbBackEnd* be = h->be;
  int ic[BE_ICORE];
  std::vector<double> w((size_t)be->Wcap * BE_WIN);
  LVB_CUDA(cudaMemcpyAsync(ic, be->icore + (size_t)seq * BE_ICORE, sizeof(ic), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaMemcpyAsync(w.data(), be->win + (size_t)seq * be->Wcap * BE_WIN, sizeof(double) * w.size(), cudaMemcpyDeviceToHost, h->stream));
  LVB_CUDA(cudaStreamSynchronize(h->stream));
  const int k = ic[I_NWIN] < cap ? ic[I_NWIN] : cap;
  for (int i = 0; i < k; ++i) {
    for (int j = 0; j < 4; ++j) qp[i * 7 + j] = w[(size_t)i * BE_WIN + W_Q + j];
    for (int j = 0; j < 3; ++j) qp[i * 7 + 4 + j] = w[(size_t)i * BE_WIN + W_P + j];
  }
  *n = k;
  return LVB_OK;
}
