// CPU oracle of LarVio::processFeatures, compiled (TEST INFRASTRUCTURE - see oracle/__init__.py: only tests/, smoke() and
// bench.py's CPU legs may load this; the product never does).
//
// C++17, dense loops, no Eigen: the same restatement of /root/reference/src/larvio.cpp as oracle/backend.py, for every configuration of the
// filter: LEG_DIM 22 or 46 (IMU-intrinsic calibration), pure MSCKF (max_features_in_one_grid: 0, the configuration BASELINE.json's metric is
// quoted on) and the hybrid filter with 1-D or 3-D inverse-depth EKF-SLAM features (promotion rule with the grid map, featureJacobian_ekf /
// _ekf_new, measurementUpdate_hybrid, anchor hand-over with updateFeatureCov_1didp / _3didp, the standstill that drops them, Schmidt
// nuisance states), with FEJ, online extrinsics / td and ZUPT.  It exists so that the CPU arm of bench.py times compiled code, as the
// reference is compiled code.  Pinned to golden vectors produced by the reference's OWN larvio.cpp
// (tests/golden/ref_*.npz, tests/test_cpu.py::test_compiled_oracle_matches_the_compiled_reference: all 13 fixtures, <= 1e-9) and to
// oracle/backend.py (test_compiled_backend_matches_the_numpy_oracle).
//
// processFeatures :363-461, batchImuProcessing :464-517, processModel :520-578, predictNewState :581-649, calPhi :3475-3530,
// stateAugmentation :720-801, addFeatureObservations :804-856, measurementJacobian_msckf :859-921, featureJacobian_msckf
// :924-981, measurementUpdate_msckf :1420-1602, gatingTest :1865-1880, removeLostFeatures :1883-2256, findRedundantImuStates
// :2259-2307, pruneImuStateBuffer :2310-2641, checkZUPT :2751-2788, measurementUpdate_ZUPT_vpq :2791-2962;
// Feature::{cost, jacobian, generateInitialGuess, checkMotion, initializePosition, _AssignAnchor} feature.hpp:252-721.
// Third-party pieces as in backend.py: SPQR thin QR -> Householder QR over the nonzero columns, JacobiSVD left null space ->
// three Householder reflections, LDLT -> Cholesky; the update is invariant to these basis choices up to rounding.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct V3 { double x, y, z; };
struct M3 { double m[9]; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 ld(const double* p) { return {p[0], p[1], p[2]}; }
inline void st(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline M3 eye3() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 skew(V3 w) { return {{0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0}}; }
inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
inline M3 tr(const M3& a) { return {{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
inline V3 mv(const M3& a, V3 v) { return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z}; }
inline V3 mtv(const M3& a, V3 v) { return {a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z}; }
inline M3 add(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
inline M3 sub(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
inline M3 scl(const M3& a, double s) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s; return r; }

// math_utils.hpp:26-231 (Eigen conventions, q = [x y z w])
inline M3 quat_to_rot(const double* q) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  return {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
           2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
}
inline void rot_to_quat(const M3& R, double* q) {
  double t = R.m[0] + R.m[4] + R.m[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (R.m[7] - R.m[5]) * t; q[1] = (R.m[2] - R.m[6]) * t; q[2] = (R.m[3] - R.m[1]) * t;
  } else {
    int i = 0;
    if (R.m[4] > R.m[0]) i = 1;
    if (R.m[8] > R.m[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R.m[i * 4] - R.m[j * 4] - R.m[k * 4] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (R.m[k * 3 + j] - R.m[j * 3 + k]) * t;
    q[j] = (R.m[j * 3 + i] + R.m[i * 3 + j]) * t;
    q[k] = (R.m[k * 3 + i] + R.m[i * 3 + k]) * t;
  }
}
inline void quat_mul(const double* a, const double* b, double* o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by; o[1] = aw * by - ax * bz + ay * bw + az * bx;
  o[2] = aw * bz + ax * by - ay * bx + az * bw; o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
inline void small_angle_quat(V3 dth, double* q) {
  const V3 dq = dth * 0.5;
  const double n2 = dot(dq, dq);
  if (n2 <= 1) { q[0] = dq.x; q[1] = dq.y; q[2] = dq.z; q[3] = std::sqrt(1 - n2); }
  else { const double s = 1.0 / std::sqrt(1 + n2); q[0] = dq.x * s; q[1] = dq.y * s; q[2] = dq.z * s; q[3] = s; }
}

struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
  double* row(int i) { return a.data() + (size_t)i * c; }
  const double* row(int i) const { return a.data() + (size_t)i * c; }
};

struct Obs { double z[2], vel[2]; };
struct ImuS {
  long long id = 0; double time = 0, dt = 0;
  double q[4] = {0, 0, 0, 1}, p[3] = {0, 0, 0}, v[3] = {0, 0, 0}, bg[3] = {0, 0, 0}, ba[3] = {0, 0, 0};
  M3 R_ic = eye3(); double t_ci[3] = {0, 0, 0};
};
struct Aug {
  long long id = 0; double time = 0, dt = 0;
  double q[4], p[3], p_fej[3]; M3 R_ic; double t_ci[3]; double q_cam[4], p_cam[3];
};
struct Feature {
  long long id = 0;
  std::map<long long, Obs> obs;
  double pos[3] = {0, 0, 0}, pos_fej[3] = {0, 0, 0};
  bool init = false;
  long long anchor = -1;
  // hybrid filter (1-D inverse-depth EKF-SLAM features, feature.hpp:231-246): in the state / potential EKF feature, inverse depth and
  // corrected bearing in the anchor camera
  bool in_state = false, ekf = false;
  double inv_depth = 0.0, obs_anchor[3] = {0, 0, 1};
  double inv_param[3] = {0, 0, 0};           // 3-D inverse depth: (x/z, y/z, 1/z) in the anchor camera (feature.hpp:231)
};

const V3 GRAV = {0.0, 0.0, -9.81};

struct Cfg {               // order = the vector oracle/backend_c.py passes
  double imu_rate, rotation_threshold, translation_threshold, tracking_rate_threshold, feature_translation_threshold, td;
  double noise_gyro, noise_acc, noise_gyro_bias, noise_acc_bias, noise_feature;
  double cov_ori, cov_vel, cov_pos, cov_bg, cov_ba, cov_er, cov_et;
  double zupt_max_feature_dis, zupt_noise_v, zupt_noise_p, zupt_noise_q;
  double max_track_len, sw_size, least_obs, if_FEJ, estimate_td, estimate_extrin, if_ZUPT_valid;
  double T_cam_imu[16];
  double chi2[100];
  double max_features, grid_rows, grid_cols, x_min, y_min, grid_w, grid_h;     // larvio.cpp:226-268 (0 features per cell = pure MSCKF)
  double calib_imu;                                                            // calib_imu_instrinsic: LEG_DIM 46 (:158-161)
  double use_schmidt;                                                          // :277
  double feature_idp_dim;                                                      // 1, anything else means 3 (:270-274)
};

struct Filter {
  Cfg c;
  int idp = 1;                               // state columns per SLAM feature
  int LEG;                                   // legacy error-state size: 22, or 46 with the 24 IMU-intrinsic states T1 T2 T3 A1 A2 A3 M1 M2
  M3 Tg = eye3(), As = {{0, 0, 0, 0, 0, 0, 0, 0, 0}}, Ma = eye3();   // larvio.cpp:129-131
  double intr[24] = {0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1};
  double th, sg2, sa2, sbg2, sba2, sfeat2, td;
  ImuS s, imu_old, fej_now, fej_old;
  Mat P;
  std::map<long long, Aug> aug;
  std::map<long long, Feature> map;
  bool if_FEJ = false, if_ZUPT = false, first = false, gravity_set = false, have_old = false;
  double g_old[3], a_old[3];
  long long next_id = 0;
  double tracking_rate = 0, take_off = 0, last_zupt = 0;
  std::vector<double> coarse;
  long long zupt_events = 0, updates = 0;
  std::vector<long long> fstates;          // state_server.feature_states: ids of the SLAM features in the state, in covariance order
  std::map<int, int> grid;                 // grid_map: features per cell code (cells outside rows*cols are never emptied, :3355-3357)
  bool hybrid() const { return c.max_features * c.grid_rows * c.grid_cols != 0; }
  // use_schmidt (:277): poses that left the window while they anchored SLAM features, kept frozen behind the feature block
  std::vector<long long> nui_ids; std::map<long long, Aug> nui_states; std::map<long long, std::vector<long long>> nui_features;
  bool schmidt() const { return c.use_schmidt != 0; }
  bool anchor_is_nui(const Feature& ft) const { return aug.find(ft.anchor) == aug.end(); }
  const Aug& anchor_state(const Feature& ft) const { auto it = aug.find(ft.anchor); return it != aug.end() ? it->second : nui_states.at(ft.anchor); }

  explicit Filter(const Cfg& cc) : c(cc) {
    th = 1.0 / (2.0 * c.imu_rate);
    sg2 = c.noise_gyro * c.noise_gyro; sa2 = c.noise_acc * c.noise_acc; sbg2 = c.noise_gyro_bias * c.noise_gyro_bias;
    sba2 = c.noise_acc_bias * c.noise_acc_bias; sfeat2 = c.noise_feature * c.noise_feature; td = c.td;
    LEG = c.calib_imu != 0 ? 46 : 22;
    idp = (c.feature_idp_dim == 1) ? 1 : 3;
    P = Mat(LEG, LEG);
    for (int i = 22; i < LEG; ++i) P(i, i) = 1e-4;                               // :183-186
    for (int i = 0; i < 3; ++i) {
      P(i, i) = c.cov_ori; P(3 + i, 3 + i) = c.cov_vel; P(6 + i, 6 + i) = c.cov_pos; P(9 + i, 9 + i) = c.cov_bg; P(12 + i, 12 + i) = c.cov_ba;
      if (c.estimate_extrin != 0) { P(15 + i, 15 + i) = c.cov_er; P(18 + i, 18 + i) = c.cov_et; }
    }
    if (c.estimate_td != 0) P(21, 21) = 4e-6;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.R_ic.m[i * 3 + j] = c.T_cam_imu[i * 4 + j];     // larvio.cpp:189-202
    for (int i = 0; i < 3; ++i) { double a = 0; for (int k = 0; k < 3; ++k) a += c.T_cam_imu[k * 4 + i] * c.T_cam_imu[k * 4 + 3]; s.t_ci[i] = -a; }
    fej_now = s; fej_old = s; imu_old = s;
  }

  // ---------------------------------------------------------------- :581-649
  void predict_new_state(double dt, V3 gyro, V3 acc) {
    const double gn = norm(gyro);
    double Om[16] = {0};
    const M3 sk = skew(gyro);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Om[i * 4 + j] = -sk.m[i * 3 + j];
    Om[3] = gyro.x; Om[7] = gyro.y; Om[11] = gyro.z; Om[12] = -gyro.x; Om[13] = -gyro.y; Om[14] = -gyro.z;
    imu_old = s;
    double dq[4], dq2[4];
    auto apply = [&](double cdiag, double coff, double post, double* out) {     // (cdiag I + coff Om) * post @ q
      for (int i = 0; i < 4; ++i) {
        double acc_ = 0;
        for (int j = 0; j < 4; ++j) acc_ += (((i == j) ? cdiag : 0.0) + coff * Om[i * 4 + j]) * post * s.q[j];
        out[i] = acc_;
      }
    };
    if (gn > 1e-5) {
      apply(std::cos(gn * dt * 0.5), 1 / gn * std::sin(gn * dt * 0.5), 1.0, dq);
      apply(std::cos(gn * dt * 0.25), 1 / gn * std::sin(gn * dt * 0.25), 1.0, dq2);
    } else {
      apply(1.0, 0.5 * dt, std::cos(gn * dt * 0.5), dq);
      apply(1.0, 0.25 * dt, std::cos(gn * dt * 0.25), dq2);
    }
    const M3 dR = quat_to_rot(dq), dR2 = quat_to_rot(dq2);
    const V3 v = ld(s.v), p = ld(s.p);
    const V3 k1v = mv(quat_to_rot(s.q), acc) + GRAV, k1p = v;
    const V3 k1_v = v + k1v * (dt / 2);
    const V3 k2v = mv(dR2, acc) + GRAV, k2p = k1_v;
    const V3 k2_v = v + k2v * (dt / 2);
    const V3 k3v = mv(dR2, acc) + GRAV, k3p = k2_v;
    const V3 k3_v = v + k3v * dt;
    const V3 k4v = mv(dR, acc) + GRAV, k4p = k3_v;
    const double qn = std::sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
    for (int i = 0; i < 4; ++i) s.q[i] = dq[i] / qn;
    st(s.v, v + (k1v + k2v * 2 + k3v * 2 + k4v) * (dt / 6));
    st(s.p, p + (k1p + k2p * 2 + k3p * 2 + k4p) * (dt / 6));
    fej_old = fej_now;
    fej_now = s;
  }

  // ---------------------------------------------------------------- :520-578 with calPhi :3475-3530 and its IMU-intrinsic columns :3532-3797
  void process_model(double time, V3 m_gyro, V3 m_acc) {
    const V3 f = m_acc - ld(s.ba), acc = mv(Ma, f);
    const V3 w = m_gyro - mv(As, acc) - ld(s.bg), gyro = mv(Tg, w);
    const V3 f_old = ld(a_old) - ld(s.ba), acc_old = mv(Ma, f_old);
    const V3 w_old = ld(g_old) - mv(As, acc_old) - ld(s.bg), gyro_old = mv(Tg, w_old);
    const double dtime = time - s.time;
    predict_new_state(dtime, gyro, acc);
    const V3 axis = (gyro_old + gyro) * (dtime / 2) + cross(gyro_old, gyro) * (dtime * dtime / 12);
    const M3 Ah = skew(axis);
    const M3 C = quat_to_rot(imu_old.q);
    const M3 I3 = eye3();
    const M3 TA = mul(Tg, As);
    V3 vk, pk, vk1, pk1;
    if (if_FEJ) { vk = ld(fej_old.v); pk = ld(fej_old.p); vk1 = ld(fej_now.v); pk1 = ld(fej_now.p); }
    else { vk = ld(imu_old.v); pk = ld(imu_old.p); vk1 = ld(s.v); pk1 = ld(s.p); }
    const V3 g = GRAV;
    const int L = LEG;
    Mat Phi(L, L);
    for (int i = 0; i < L; ++i) Phi(i, i) = 1.0;
    auto put = [&](int r0, int c0, const M3& B) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Phi(r0 + i, c0 + j) = B.m[i * 3 + j]; };
    const M3 CA2 = mul(C, add(scl(I3, 2.0), Ah));
    put(0, 9, mul(scl(CA2, -0.5 * dtime), Tg));
    put(0, 12, mul(mul(scl(CA2, 0.5 * dtime), TA), Ma));
    put(3, 0, scl(skew(vk1 - vk - g * dtime), -1.0));
    const M3 P39 = add(mul(skew(pk * 1.0 - pk1 + vk1 * dtime - g * (0.5 * dtime * dtime)), C),
                       mul(mul(skew(pk * 0.5 - pk1 * 0.5 + vk1 * (0.5 * dtime) - g * (dtime * dtime / 6)), C), Ah));
    put(3, 9, P39);
    put(3, 12, sub(mul(scl(CA2, -0.5 * dtime), Ma), mul(mul(P39, TA), Ma)));
    put(6, 0, scl(skew(pk1 - pk - vk * dtime - g * (0.5 * dtime * dtime)), -1.0));
    put(6, 3, scl(I3, dtime));
    const M3 P69 = add(scl(mul(skew(g), C), -dtime * dtime * dtime / 6), scl(mul(mul(skew(pk1 - pk - g * (dtime * dtime / 6)), C), Ah), dtime / 4));
    put(6, 9, P69);
    put(6, 12, sub(mul(scl(mul(C, add(scl(I3, 3.0), Ah)), -dtime * dtime / 6), Ma), mul(mul(P69, TA), Ma)));
    if (L > 22) {
      // selectors of a 3-vector x: Lo (1,0)=x0 (2,1)=x0 (2,2)=x1 ; Di = diag(x) ; Up (0,0)=x1 (0,1)=x2 (1,2)=x2   (:3532-3797)
      auto Lo = [](V3 x) { M3 m = {{0, 0, 0, 0, 0, 0, 0, 0, 0}}; m.m[3] = x.x; m.m[7] = x.x; m.m[8] = x.y; return m; };
      auto Di = [](V3 x) { M3 m = {{0, 0, 0, 0, 0, 0, 0, 0, 0}}; m.m[0] = x.x; m.m[4] = x.y; m.m[8] = x.z; return m; };
      auto Up = [](V3 x) { M3 m = {{0, 0, 0, 0, 0, 0, 0, 0, 0}}; m.m[0] = x.y; m.m[1] = x.z; m.m[5] = x.z; return m; };
      const V3 f_mid = (f + f_old) * 0.5, acc_mid = (acc + acc_old) * 0.5;
      const V3 w_mid = (w_old + w) * 0.5 + cross(w_old, w) * (dtime / 12);
      const M3 R_mid = add(I3, scl(Ah, 0.5)), R_kp1 = add(I3, Ah);
      const M3 S_mid = skew(mv(R_mid, acc_mid)), S_kp1 = skew(mv(R_kp1, acc));
      const M3 Z3 = {{0, 0, 0, 0, 0, 0, 0, 0, 0}};
      for (int grp = 0; grp < 8; ++grp) {
        const int col = 22 + 3 * grp;
        const int which = grp % 3;                                   // 0 Lo, 1 Di, 2 Up (the last group pair is Lo, Di)
        const int fam = grp / 3;                                     // 0: T (w), 1: A (acc), 2: M (f)
        V3 xk, xh, xp; M3 Lf; double sgn; bool direct;
        if (fam == 0) { xk = w_old; xh = w_mid; xp = w; Lf = I3; sgn = 1; direct = false; }
        else if (fam == 1) { xk = acc_old; xh = acc_mid; xp = acc; Lf = Tg; sgn = -1; direct = false; }
        else { xk = f_old; xh = f_mid; xp = f; Lf = TA; sgn = -1; direct = true; }
        auto sel = [&](V3 x) { return which == 0 ? Lo(x) : which == 1 ? Di(x) : Up(x); };
        const M3 kq1 = mul(Lf, sel(xk)), kq2 = mul(R_mid, mul(Lf, sel(xh))), kq4 = mul(R_kp1, mul(Lf, sel(xp)));
        const M3 Rq = scl(add(add(kq1, scl(kq2, 4.0)), kq4), dtime / 6);
        put(0, col, scl(mul(C, Rq), sgn));
        M3 kv1, kv2, kv3, kv4;
        if (!direct) {
          kv1 = Z3; kv2 = scl(mul(scl(S_mid, dtime), kq1), 0.5); kv3 = scl(mul(scl(S_mid, dtime), kq2), 0.5); kv4 = mul(S_kp1, Rq);
        } else {
          kv1 = sel(xk);
          kv2 = add(mul(R_mid, sel(xh)), scl(mul(scl(S_mid, dtime), kq1), 0.5));
          kv3 = add(mul(R_mid, sel(xh)), scl(mul(scl(S_mid, dtime), kq2), 0.5));
          kv4 = add(mul(R_kp1, sel(xp)), mul(S_kp1, Rq));
        }
        const M3 fR = scl(add(add(kv1, scl(kv2, 2.0)), add(scl(kv3, 2.0), kv4)), dtime / 6);
        const double vs = direct ? 1.0 : -sgn;                       // Phi_v: -C f for T, +C f for A, +C v for M
        put(3, col, scl(mul(C, fR), vs));
        const M3 kp2 = scl(kv1, dtime / 2), kp3 = scl(kv2, dtime / 2);
        put(6, col, scl(mul(C, scl(add(add(scl(kp2, 2.0), scl(kp3, 2.0)), fR), dtime / 6)), vs));
      }
    }
    // Q = Phi G Qc G^T Phi^T dt, G Qc G^T = blkdiag(sg2 C C^T, sa2 C C^T, 0, sbg2 I, sba2 I, 0)
    const M3 CCt = mul(C, tr(C));
    Mat M(15, 15);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { M(i, j) = sg2 * CCt.m[i * 3 + j]; M(3 + i, 3 + j) = sa2 * CCt.m[i * 3 + j]; }
    for (int i = 0; i < 3; ++i) { M(9 + i, 9 + i) = sbg2; M(12 + i, 12 + i) = sba2; }
    Mat PM(L, 15), Q(L, L);
    for (int i = 0; i < L; ++i) for (int j = 0; j < 15; ++j) { double a_ = 0; for (int k = 0; k < 15; ++k) a_ += Phi(i, k) * M(k, j); PM(i, j) = a_; }
    for (int i = 0; i < L; ++i) for (int j = 0; j < L; ++j) { double a_ = 0; for (int k = 0; k < 15; ++k) a_ += PM(i, k) * Phi(j, k); Q(i, j) = a_ * dtime; }
    // P[:L,:L] = Phi P Phi^T + Q ; P[:L,L:] = Phi P[:L,L:] ; P[L:,:L] = transpose ; P = (P + P^T)/2
    const int d = P.r;
    std::vector<double> top((size_t)L * d);
    for (int i = 0; i < L; ++i) {
      double* o = top.data() + (size_t)i * d;
      for (int j = 0; j < d; ++j) o[j] = 0.0;
      for (int k = 0; k < L; ++k) { const double ph = Phi(i, k); if (ph == 0.0) continue; const double* pr = P.row(k); for (int j = 0; j < d; ++j) o[j] += ph * pr[j]; }
    }
    Mat LL(L, L);
    for (int i = 0; i < L; ++i) for (int j = 0; j < L; ++j) { double a_ = 0; for (int k = 0; k < L; ++k) a_ += top[(size_t)i * d + k] * Phi(j, k); LL(i, j) = a_ + Q(i, j); }
    for (int i = 0; i < L; ++i) {
      for (int j = 0; j < L; ++j) P(i, j) = LL(i, j);
      for (int j = L; j < d; ++j) { P(i, j) = top[(size_t)i * d + j]; P(j, i) = top[(size_t)i * d + j]; }
    }
    for (int i = 0; i < L; ++i) for (int j = i + 1; j < L; ++j) { const double m_ = 0.5 * (P(i, j) + P(j, i)); P(i, j) = m_; P(j, i) = m_; }
    s.time = time; fej_now.time = time;
  }

  // ---------------------------------------------------------------- :464-517
  int batch_imu(double time_bound, const double* imu, int m) {
    int used = 0; double dt = 0.0;
    for (int k = 0; k < m; ++k) {
      const double* row = imu + (size_t)k * 7;
      const double t = row[0];
      if (t <= s.time) { ++used; continue; }
      if (t - time_bound > th) break;
      dt = t - time_bound;
      if (!have_old) { for (int i = 0; i < 3; ++i) { g_old[i] = row[1 + i]; a_old[i] = row[4 + i]; } have_old = true; }
      process_model(t, ld(row + 1), ld(row + 4));
      ++used;
      for (int i = 0; i < 3; ++i) { g_old[i] = row[1 + i]; a_old[i] = row[4 + i]; }
    }
    s.id = next_id++;
    s.dt = dt;
    return used;
  }

  // ---------------------------------------------------------------- :720-801
  void augment() {
    Aug a;
    a.id = s.id; a.time = s.time; a.dt = s.dt;
    memcpy(a.q, s.q, sizeof a.q); memcpy(a.p, s.p, sizeof a.p); memcpy(a.p_fej, fej_now.p, sizeof a.p_fej);
    a.R_ic = s.R_ic; memcpy(a.t_ci, s.t_ci, sizeof a.t_ci);
    const M3 R_b2w = quat_to_rot(s.q);
    const M3 R_w2c = mul(s.R_ic, tr(R_b2w));
    rot_to_quat(tr(R_w2c), a.q_cam);
    st(a.p_cam, ld(s.p) + mv(R_b2w, ld(s.t_ci)));
    aug[s.id] = a;
    const int d = P.r;
    const int sel[6] = {0, 1, 2, 6, 7, 8};
    Mat Pn(d + 6, d + 6);
    for (int i = 0; i < d; ++i) memcpy(Pn.row(i), P.row(i), sizeof(double) * d);
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < d; ++j) { const double v_ = P(sel[i], j); Pn(d + i, j) = v_; Pn(j, d + i) = v_; }
      for (int j = 0; j < 6; ++j) Pn(d + i, d + j) = P(sel[i], sel[j]);
    }
    const int nf = idp * (int)fstates.size() + 6 * (int)nui_ids.size();
    if (nf > 0) {                                       // the new pose goes in FRONT of the SLAM-feature (and nuisance) block (:768-793)
      std::vector<int> order; const int pe = d - nf;
      for (int i = 0; i < pe; ++i) order.push_back(i);
      for (int i = 0; i < 6; ++i) order.push_back(d + i);
      for (int i = pe; i < d; ++i) order.push_back(i);
      P = permuted(Pn, order);
    } else P = std::move(Pn);
  }

  static Mat permuted(const Mat& A, const std::vector<int>& order) {
    const int n = (int)order.size();
    Mat o(n, n);
    for (int i = 0; i < n; ++i) { const double* src = A.row(order[i]); double* dst = o.row(i); for (int j = 0; j < n; ++j) dst[j] = src[order[j]]; }
    return o;
  }

  // ---------------------------------------------------------------- :804-856
  void add_observations(const long long* ids, const double* data, int n) {
    const long long sid = s.id;
    const double curr_num = (double)map.size();
    double tracked = 0;
    const double dt = s.dt;
    for (int k = 0; k < n; ++k) {
      const double* f = data + (size_t)k * 8;
      const double u = f[0], v = f[1], u_init = f[2], v_init = f[3], u_vel = f[4], v_vel = f[5], u_init_vel = f[6], v_init_vel = f[7];
      auto it = map.find(ids[k]);
      if (it == map.end()) {
        Feature ft; ft.id = ids[k];
        ft.obs[sid] = Obs{{u + u_vel * dt, v + v_vel * dt}, {u_vel, v_vel}};
        if (!(u_init == -1 && v_init == -1)) {
          auto pa = aug.find(sid - 1);
          if (pa != aug.end()) { const double dt_ = pa->second.dt; ft.obs[sid - 1] = Obs{{u_init + u_init_vel * dt_, v_init + v_init_vel * dt_}, {u_init_vel, v_init_vel}}; }
        }
        map[ids[k]] = std::move(ft);
      } else {
        Feature& ft = it->second;
        ft.obs[sid] = Obs{{u + u_vel * dt, v + v_vel * dt}, {u_vel, v_vel}};
        tracked += 1;
        if (c.if_ZUPT_valid != 0) {
          auto po = ft.obs.find(sid - 1);
          if (po != ft.obs.end()) { const double dx = u - po->second.z[0], dy = v - po->second.z[1]; coarse.push_back(std::sqrt(dx * dx + dy * dy)); }
        }
      }
    }
    tracking_rate = tracked / curr_num;              // 0/0 -> NaN like the reference's double division
  }

  // ---------------------------------------------------------------- state injection of an update (:1497-1533)
  // new_at / num_old: in measurementUpdate_hybrid the corrections of the features added by this update sit behind the old covariance
  void inject(const std::vector<double>& dx, int new_at = -1, int num_old = -1) {
    double dq[4], qn[4];
    small_angle_quat({dx[0], dx[1], dx[2]}, dq);
    quat_mul(dq, s.q, qn); memcpy(s.q, qn, sizeof qn);
    for (int i = 0; i < 3; ++i) { s.v[i] += dx[3 + i]; s.p[i] += dx[6 + i]; s.bg[i] += dx[9 + i]; s.ba[i] += dx[12 + i]; }
    small_angle_quat({dx[15], dx[16], dx[17]}, dq);
    s.R_ic = mul(s.R_ic, tr(quat_to_rot(dq)));
    for (int i = 0; i < 3; ++i) s.t_ci[i] += dx[18 + i];
    td += dx[21];
    if (LEG > 22) {                                   // T1 T2 T3 A1 A2 A3 M1 M2 += dx(22:46), then updateImuMx (:1497-1507, 3803-3847)
      for (int k = 0; k < 24; ++k) intr[k] += dx[22 + k];
      const double* T1 = intr; const double* T2 = intr + 3; const double* T3 = intr + 6; const double* A1 = intr + 9; const double* A2 = intr + 12;
      const double* A3 = intr + 15; const double* M1 = intr + 18; const double* M2 = intr + 21;
      Tg = {{T2[0], T3[0], T3[1], T1[0], T2[1], T3[2], T1[1], T1[2], T2[2]}};
      As = {{A2[0], A3[0], A3[1], A1[0], A2[1], A3[2], A1[1], A1[2], A2[2]}};
      Ma.m[0] = M2[0]; Ma.m[3] = M1[0]; Ma.m[4] = M2[1]; Ma.m[6] = M1[1]; Ma.m[7] = M1[2]; Ma.m[8] = M2[2];   // the upper triangle is never written (:3839-3844)
    }
    int i = 0;
    for (auto& kv : aug) {
      Aug& a = kv.second;
      const double* da = dx.data() + LEG + 6 * i;
      small_angle_quat({da[0], da[1], da[2]}, dq);
      quat_mul(dq, a.q, qn); memcpy(a.q, qn, sizeof qn);
      for (int k = 0; k < 3; ++k) a.p[k] += da[3 + k];
      const M3 R_b2w = quat_to_rot(a.q);
      rot_to_quat(mul(R_b2w, tr(s.R_ic)), a.q_cam);
      st(a.p_cam, ld(a.p) + mv(R_b2w, ld(s.t_ci)));
      ++i;
    }
    const int base = LEG + 6 * (int)aug.size();       // inverse depth of the in-state features, world position from the anchor (:1536-1575)
    for (int k = 0; k < (int)fstates.size(); ++k) {
      Feature& ft = map.at(fstates[k]);
      const int at = (num_old < 0 || k < num_old) ? base + idp * k : new_at + idp * (k - num_old);
      const Aug& an = anchor_state(ft);
      V3 p_c;
      if (idp == 3) {
        for (int q = 0; q < 3; ++q) ft.inv_param[q] += dx[at + q];
        p_c = {ft.inv_param[0] / ft.inv_param[2], ft.inv_param[1] / ft.inv_param[2], 1.0 / ft.inv_param[2]};
      } else {
        ft.inv_depth += dx[at];
        p_c = {ft.obs_anchor[0] / ft.inv_depth, ft.obs_anchor[1] / ft.inv_depth, 1.0 / ft.inv_depth};
      }
      st(ft.pos, mv(quat_to_rot(an.q_cam), p_c) + ld(an.p_cam));
    }
  }

  // nonzero columns of a stacked Jacobian (exact zero test)
  static std::vector<int> nonzero_cols(const Mat& H) {
    std::vector<char> nz(H.c, 0);
    for (int i = 0; i < H.r; ++i) { const double* r = H.row(i); for (int j = 0; j < H.c; ++j) if (r[j] != 0.0) nz[j] = 1; }
    std::vector<int> out;
    for (int j = 0; j < H.c; ++j) if (nz[j]) out.push_back(j);
    return out;
  }

  // T = H P over the nonzero columns of H
  Mat times_P(const Mat& H, const std::vector<int>& nz) const {
    const int d = P.r;
    Mat T(H.r, d);
    for (int i = 0; i < H.r; ++i) {
      double* o = T.row(i);
      for (int c_ : nz) { const double h = H(i, c_); if (h == 0.0) continue; const double* pr = P.row(c_); for (int j = 0; j < d; ++j) o[j] += h * pr[j]; }
    }
    return T;
  }

  // in-place Cholesky (lower) of the m x m matrix S; false if not positive definite
  static bool cholesky(Mat& S) {
    const int m = S.r;
    for (int j = 0; j < m; ++j) {
      double djj = S(j, j);
      for (int k = 0; k < j; ++k) djj -= S(j, k) * S(j, k);
      if (!(djj > 0.0)) return false;
      djj = std::sqrt(djj); S(j, j) = djj;
      for (int i = j + 1; i < m; ++i) {
        double a_ = S(i, j);
        const double* ri = S.row(i); const double* rj = S.row(j);
        for (int k = 0; k < j; ++k) a_ -= ri[k] * rj[k];
        S(i, j) = a_ / djj;
      }
    }
    return true;
  }

  // ---------------------------------------------------------------- gatingTest :1865-1880
  bool gating(const Mat& H, const std::vector<double>& r, int dof) const {
    const int m = H.r;
    if (dof < 1 || dof >= 100) return false;
    const std::vector<int> nz = nonzero_cols(H);
    Mat S(m, m);
    std::vector<double> t(nz.size());
    for (int i = 0; i < m; ++i) {
      for (size_t a_ = 0; a_ < nz.size(); ++a_) { double acc = 0; for (int b : nz) acc += H(i, b) * P(b, nz[a_]); t[a_] = acc; }
      for (int j = 0; j <= i; ++j) { double acc = 0; for (size_t a_ = 0; a_ < nz.size(); ++a_) acc += t[a_] * H(j, nz[a_]); S(i, j) = acc + (i == j ? sfeat2 : 0.0); }
    }
    if (!cholesky(S)) return false;
    double gamma = 0;
    std::vector<double> y(m);
    for (int i = 0; i < m; ++i) { double x = r[i]; for (int k = 0; k < i; ++k) x -= S(i, k) * y[k]; x /= S(i, i); y[i] = x; gamma += x * x; }
    return gamma < c.chi2[dof];
  }

  // ---------------------------------------------------------------- thin QR of H (rows > cols) over its nonzero columns (:1430-1449)
  static void compress(Mat& H, std::vector<double>& r) {
    const int m = H.r, d = H.c;
    const std::vector<int> nz = nonzero_cols(H);
    const int nc = (int)nz.size();
    if (m <= nc) return;
    std::vector<double> A((size_t)nc * m);                 // column-major panel
    for (int j = 0; j < nc; ++j) for (int i = 0; i < m; ++i) A[(size_t)j * m + i] = H(i, nz[j]);
    std::vector<double> v(m);
    for (int k = 0; k < nc; ++k) {
      double* ck = A.data() + (size_t)k * m;
      double n2 = 0; for (int i = k; i < m; ++i) n2 += ck[i] * ck[i];
      const double nrm = std::sqrt(n2);
      if (nrm == 0.0) continue;
      const double alpha = ck[k] >= 0 ? -nrm : nrm;
      double vtv = 0;
      for (int i = k; i < m; ++i) { v[i] = ck[i] - (i == k ? alpha : 0.0); vtv += v[i] * v[i]; }
      if (vtv == 0.0) continue;
      const double beta = 2.0 / vtv;
      for (int j = k + 1; j < nc; ++j) {
        double* cj = A.data() + (size_t)j * m;
        double dt_ = 0; for (int i = k; i < m; ++i) dt_ += v[i] * cj[i];
        dt_ *= beta;
        for (int i = k; i < m; ++i) cj[i] -= dt_ * v[i];
      }
      double dt_ = 0; for (int i = k; i < m; ++i) dt_ += v[i] * r[i];
      dt_ *= beta;
      for (int i = k; i < m; ++i) r[i] -= dt_ * v[i];
      ck[k] = alpha; for (int i = k + 1; i < m; ++i) ck[i] = 0.0;
    }
    Mat R(nc, d);
    for (int j = 0; j < nc; ++j) for (int i = 0; i <= j; ++i) R(i, nz[j]) = A[(size_t)j * m + i];
    H = std::move(R);
    r.resize(nc);
  }

  // ---------------------------------------------------------------- measurementUpdate_msckf :1420-1602 (Rdiag: ZUPT's per-row variances)
  void update(const Mat& H, const std::vector<double>& r, const double* Rdiag = nullptr) {
    std::vector<double> dx; Mat Y;
    if (!solve_update(H, r, Rdiag, dx, Y)) return;
    inject(dx);
    if (schmidt() && !nui_ids.empty()) {
      const int n0 = LEG + 6 * (int)aug.size() + idp * (int)fstates.size();
      const Mat B = nuisance_block(n0);
      apply_cov(Y);
      restore_nuisance_block(n0, B);
    } else apply_cov(Y);
  }
  // gain in square-root form: Y = L^-1 H P (S = L L^T), dx = Y^T L^-1 r; false: nothing to do
  bool solve_update(const Mat& H, const std::vector<double>& r, const double* Rdiag, std::vector<double>& dx, Mat& T) {
    const int m = H.r, d = P.r;
    if (m == 0) return false;
    const std::vector<int> nz = nonzero_cols(H);
    T = times_P(H, nz);                                      // H P
    Mat S(m, m);
    for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) { double acc = 0; for (int c_ : nz) acc += T(i, c_) * H(j, c_); S(i, j) = acc + (i == j ? (Rdiag ? Rdiag[i] : sfeat2) : 0.0); }
    if (!cholesky(S)) return false;
    // Y = L^-1 T, z = L^-1 r ; dx = Y^T z ; P -= Y^T Y
    std::vector<double> z(m);
    for (int i = 0; i < m; ++i) {
      double* ti = T.row(i);
      double x = r[i];
      for (int k = 0; k < i; ++k) { const double l = S(i, k); if (l == 0.0) continue; const double* tk = T.row(k); for (int j = 0; j < d; ++j) ti[j] -= l * tk[j]; x -= l * z[k]; }
      const double inv = 1.0 / S(i, i);
      for (int j = 0; j < d; ++j) ti[j] *= inv;
      z[i] = x * inv;
    }
    dx.assign(d, 0.0);
    for (int i = 0; i < m; ++i) { const double* ti = T.row(i); for (int j = 0; j < d; ++j) dx[j] += ti[j] * z[i]; }
    return true;
  }
  void apply_cov(const Mat& T) {                             // P -= Y^T Y, symmetrised
    const int m = T.r, d = P.r;
    for (int k = 0; k < m; ++k) {
      const double* tk = T.row(k);
      for (int i = 0; i < d; ++i) { const double a_ = tk[i]; if (a_ == 0.0) continue; double* pr = P.row(i); for (int j = 0; j < d; ++j) pr[j] -= a_ * tk[j]; }
    }
    for (int i = 0; i < d; ++i) for (int j = i + 1; j < d; ++j) { const double m_ = 0.5 * (P(i, j) + P(j, i)); P(i, j) = m_; P(j, i) = m_; }
    ++updates;
  }

  // ---------------------------------------------------------------- Feature::checkMotion feature.hpp:334-381
  bool check_motion(const Feature& ft, bool if_tracked) const {
    auto first = ft.obs.begin();
    auto last = ft.obs.end(); --last;
    if (if_tracked) --last;
    const Aug& af = aug.at(first->first);
    const Aug& al = aug.at(last->first);
    V3 dir = {first->second.z[0], first->second.z[1], 1.0};
    dir = dir * (1.0 / norm(dir));
    dir = mv(quat_to_rot(af.q_cam), dir);
    const V3 t = ld(al.p_cam) - ld(af.p_cam);
    const double par = dot(t, dir);
    return norm(t - dir * par) > c.feature_translation_threshold;
  }

  // ---------------------------------------------------------------- Feature::initializePosition[_AssignAnchor] feature.hpp:383-721
  struct RelPose { M3 R; V3 t; double z[2]; };
  static double cost(const RelPose& p, const double* x) {
    const V3 h = mv(p.R, {x[0], x[1], 1.0}) + p.t * x[2];
    const double dx = h.x / h.z - p.z[0], dy = h.y / h.z - p.z[1];
    return dx * dx + dy * dy;
  }
  bool initialize_position(Feature& ft, long long skip_id, bool use_skip, bool force_guess = false) {
    std::vector<RelPose> rel; std::vector<long long> cam_ids;
    std::vector<std::pair<M3, V3>> poses;
    for (auto& kv : ft.obs) {
      auto ia = aug.find(kv.first);
      if (ia == aug.end()) continue;
      if (use_skip && kv.first == skip_id) continue;
      RelPose rp; rp.z[0] = kv.second.z[0]; rp.z[1] = kv.second.z[1];
      rel.push_back(rp); cam_ids.push_back(kv.first);
      poses.push_back({quat_to_rot(ia->second.q_cam), ld(ia->second.p_cam)});
    }
    const int n = (int)rel.size();
    if (n == 0) return false;
    const M3 Rl = poses.back().first; const V3 tl = poses.back().second;
    for (int i = 0; i < n; ++i) { rel[i].R = mul(tr(poses[i].first), Rl); rel[i].t = mtv(poses[i].first, tl - poses[i].second); }
    V3 init;
    if (!ft.init || force_guess) {                                       // generateInitialGuess feature.hpp:312-332
      const V3 m = mv(rel[0].R, {rel[n - 1].z[0], rel[n - 1].z[1], 1.0});
      const double* z2 = rel[0].z;
      const double A0 = m.x - z2[0] * m.z, A1 = m.y - z2[1] * m.z;
      const double b0 = z2[0] * rel[0].t.z - rel[0].t.x, b1 = z2[1] * rel[0].t.z - rel[0].t.y;
      const double depth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1);
      init = {rel[n - 1].z[0] * depth, rel[n - 1].z[1] * depth, depth};
    } else init = mtv(Rl, ld(ft.pos) - tl);
    double sol[3] = {init.x / init.z, init.y / init.z, 1.0 / init.z};
    double lam = 1e-3;
    int inner = 0, outer = 0; bool reduced = false; double delta_norm = 0;
    double total = 0; for (auto& p : rel) total += cost(p, sol);
    while (true) {
      double A[9] = {0}, b[3] = {0};
      for (auto& p : rel) {
        const V3 h = mv(p.R, {sol[0], sol[1], 1.0}) + p.t * sol[2];
        const double W[9] = {p.R.m[0], p.R.m[1], p.t.x, p.R.m[3], p.R.m[4], p.t.y, p.R.m[6], p.R.m[7], p.t.z};
        double J[6];
        for (int k = 0; k < 3; ++k) { J[k] = 1 / h.z * W[k] - h.x / (h.z * h.z) * W[6 + k]; J[3 + k] = 1 / h.z * W[3 + k] - h.y / (h.z * h.z) * W[6 + k]; }
        const double r0 = h.x / h.z - p.z[0], r1 = h.y / h.z - p.z[1];
        const double e = std::sqrt(r0 * r0 + r1 * r1);
        const double w = e <= 0.01 ? 1.0 : std::sqrt(2.0 * 0.01 / e);
        const double ww = (w == 1.0) ? 1.0 : w * w;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) A[i * 3 + j] += ww * (J[i] * J[j] + J[3 + i] * J[3 + j]); b[i] += ww * (J[i] * r0 + J[3 + i] * r1); }
      }
      while (true) {
        double Mx[9]; for (int i = 0; i < 9; ++i) Mx[i] = A[i]; Mx[0] += lam; Mx[4] += lam; Mx[8] += lam;
        // 3x3 solve, Gaussian elimination with partial pivoting (what np.linalg.solve / Eigen's ldlt see is well conditioned here)
        double Ab[3][4] = {{Mx[0], Mx[1], Mx[2], b[0]}, {Mx[3], Mx[4], Mx[5], b[1]}, {Mx[6], Mx[7], Mx[8], b[2]}};
        for (int col = 0; col < 3; ++col) {
          int piv = col;
          for (int i = col + 1; i < 3; ++i) if (std::fabs(Ab[i][col]) > std::fabs(Ab[piv][col])) piv = i;
          if (piv != col) for (int j = 0; j < 4; ++j) std::swap(Ab[col][j], Ab[piv][j]);
          for (int i = col + 1; i < 3; ++i) { const double f_ = Ab[i][col] / Ab[col][col]; for (int j = col; j < 4; ++j) Ab[i][j] -= f_ * Ab[col][j]; }
        }
        double delta[3];
        for (int i = 2; i >= 0; --i) { double x = Ab[i][3]; for (int j = i + 1; j < 3; ++j) x -= Ab[i][j] * delta[j]; delta[i] = x / Ab[i][i]; }
        const double ns[3] = {sol[0] - delta[0], sol[1] - delta[1], sol[2] - delta[2]};
        delta_norm = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
        double nc_ = 0; for (auto& p : rel) nc_ += cost(p, ns);
        if (nc_ < total) { reduced = true; sol[0] = ns[0]; sol[1] = ns[1]; sol[2] = ns[2]; total = nc_; lam = lam / 10 > 1e-10 ? lam / 10 : 1e-10; }
        else { reduced = false; lam = lam * 10 < 1e12 ? lam * 10 : 1e12; }
        const bool cont = (inner < 10) && !reduced;
        ++inner;
        if (!cont) break;
      }
      inner = 0;
      const bool cont = (outer < 10) && (delta_norm > 5e-7);
      ++outer;
      if (!cont) break;
    }
    const V3 fin = {sol[0] / sol[2], sol[1] / sol[2], 1.0 / sol[2]};
    bool valid = true;
    for (auto& p : rel) { const V3 pp = mv(p.R, fin) + p.t; if (pp.z <= 0) { valid = false; break; } }
    if (total / (2.0 * n * n) > 4.7673e-04) valid = false;
    if (!(total == total)) valid = false;
    if (valid) {
      if (!ft.init) memcpy(ft.pos_fej, ft.pos, sizeof ft.pos);      // feature.hpp:538-539 (the previous estimate, literal)
      ft.init = true;
      st(ft.pos, mv(Rl, fin) + tl);
      ft.anchor = cam_ids.back();
      ft.inv_depth = 1.0 / fin.z;                                    // feature.hpp:541-546
      ft.obs_anchor[0] = fin.x * ft.inv_depth; ft.obs_anchor[1] = fin.y * ft.inv_depth; ft.obs_anchor[2] = 1.0;
      ft.inv_param[0] = fin.x / fin.z; ft.inv_param[1] = fin.y / fin.z; ft.inv_param[2] = 1.0 / fin.z;      // feature.hpp:711-713
    }
    return valid;
  }
  // initializeInvParamPosition (feature.hpp:723-890): always from the two-view guess; marks a potential EKF-SLAM feature
  bool initialize_inv_param(Feature& ft, long long curr_id) {
    const bool ok = initialize_position(ft, curr_id, true, true);
    if (ok) ft.ekf = true;
    return ok;
  }

  // ---------------------------------------------------------------- featureJacobian_msckf :924-981 (+ measurementJacobian_msckf :859-921)
  // returns the block projected onto the left null space of H_f (three reflections; rows 3.. of the reflected block)
  void feature_jacobian(const Feature& ft, const std::vector<long long>& state_ids, Mat& Ho, std::vector<double>& ro) const {
    std::vector<long long> valid;
    for (long long sid : state_ids) if (ft.obs.count(sid)) valid.push_back(sid);
    const int rows = 2 * (int)valid.size(), d = P.r;
    Mat Hx(rows, d); std::vector<double> Hf((size_t)rows * 3), r(rows);
    std::vector<int> cols; for (int j = 15; j < 22; ++j) cols.push_back(j);
    const V3 p_w = ld(ft.pos);
    int k = 0;
    for (long long sid : valid) {
      const auto ia = aug.find(sid);
      const Aug& a = ia->second;
      const int cntr = (int)std::distance(aug.begin(), ia);
      const M3 R_b2w = quat_to_rot(a.q);
      const M3 R_w2c = mul(a.R_ic, tr(R_b2w));
      const V3 t_c_w = ld(a.p) + mv(R_b2w, ld(a.t_ci));
      const V3 p_c = mv(R_w2c, p_w - t_c_w);
      const V3 p_bf_w = if_FEJ ? (p_w - ld(a.p_fej)) : (p_w - ld(a.p));
      const double dz[2][3] = {{1 / p_c.z, 0, -p_c.x / (p_c.z * p_c.z)}, {0, 1 / p_c.z, -p_c.y / (p_c.z * p_c.z)}};
      const M3 A1 = mul(R_w2c, skew(p_bf_w));
      const M3 E1 = sub(mul(A1, R_b2w), mul(a.R_ic, skew(ld(a.t_ci))));
      const Obs& o = ft.obs.at(sid);
      const int cp = LEG + 6 * cntr;
      for (int q = 0; q < 6; ++q) cols.push_back(cp + q);
      for (int rr = 0; rr < 2; ++rr) {
        double* row = Hx.row(k + rr);
        for (int cc = 0; cc < 3; ++cc) {
          double hx = 0, hp = 0, he = 0, ht = 0, hf = 0;
          for (int q = 0; q < 3; ++q) {
            hx += dz[rr][q] * A1.m[q * 3 + cc]; hp += dz[rr][q] * -R_w2c.m[q * 3 + cc]; he += dz[rr][q] * E1.m[q * 3 + cc];
            ht += dz[rr][q] * -a.R_ic.m[q * 3 + cc]; hf += dz[rr][q] * R_w2c.m[q * 3 + cc];
          }
          row[cp + cc] = hx; row[cp + 3 + cc] = hp; row[15 + cc] = he; row[18 + cc] = ht; Hf[(size_t)(k + rr) * 3 + cc] = hf;
        }
        if (c.estimate_td != 0) row[21] = o.vel[rr];
      }
      r[k] = o.z[0] - p_c.x / p_c.z; r[k + 1] = o.z[1] - p_c.y / p_c.z;
      k += 2;
    }
    std::vector<double> v(rows);
    for (int kk = 0; kk < 3 && kk < rows; ++kk) {
      double n2 = 0; for (int i = kk; i < rows; ++i) n2 += Hf[(size_t)i * 3 + kk] * Hf[(size_t)i * 3 + kk];
      const double nrm = std::sqrt(n2);
      const double alpha = Hf[(size_t)kk * 3 + kk] >= 0 ? -nrm : nrm;
      double vtv = 0;
      for (int i = kk; i < rows; ++i) { v[i] = Hf[(size_t)i * 3 + kk] - (i == kk ? alpha : 0.0); vtv += v[i] * v[i]; }
      if (!(vtv > 0.0)) continue;
      const double beta = 2.0 / vtv;
      for (int cc = kk; cc < 3; ++cc) { double dt_ = 0; for (int i = kk; i < rows; ++i) dt_ += v[i] * Hf[(size_t)i * 3 + cc]; dt_ *= beta; for (int i = kk; i < rows; ++i) Hf[(size_t)i * 3 + cc] -= dt_ * v[i]; }
      for (int cc : cols) { double dt_ = 0; for (int i = kk; i < rows; ++i) dt_ += v[i] * Hx(i, cc); dt_ *= beta; for (int i = kk; i < rows; ++i) Hx(i, cc) -= dt_ * v[i]; }
      double dt_ = 0; for (int i = kk; i < rows; ++i) dt_ += v[i] * r[i]; dt_ *= beta; for (int i = kk; i < rows; ++i) r[i] -= dt_ * v[i];
    }
    const int R = rows > 3 ? rows - 3 : 0;
    Ho = Mat(R, d); ro.assign(R, 0.0);
    for (int i = 0; i < R; ++i) { memcpy(Ho.row(i), Hx.row(i + 3), sizeof(double) * d); ro[i] = r[i + 3]; }
  }

  static void append(Mat& H, std::vector<double>& r, const Mat& Hj, const std::vector<double>& rj) {
    if (H.c == 0) H.c = Hj.c;
    H.a.insert(H.a.end(), Hj.a.begin(), Hj.a.end()); H.r += Hj.r;
    r.insert(r.end(), rj.begin(), rj.end());
  }

  // ================================================================ hybrid filter: 1-D inverse-depth EKF-SLAM features
  // ---------------------------------------------------------------- measurementJacobian_ekf_1didp :1117-1244
  struct Jac1 { double hf[2][3], ha[2][6], hx[2][6], he[2][6], r[2]; };
  void meas_jacobian_idp(long long sid, const Feature& ft, Jac1& J) const {
    const Aug& k = aug.at(sid); const Aug& a = anchor_state(ft);
    const bool nui = anchor_is_nui(ft);                  // a nuisance anchor is used as frozen: own camera pose, no first estimates (:1168-1183)
    const bool fej_a = if_FEJ && !nui;
    const M3 R_b2c = k.R_ic; const V3 t_c_b = ld(k.t_ci);
    const V3 f_an = idp == 3 ? V3{ft.inv_param[0], ft.inv_param[1], 1.0} : V3{ft.obs_anchor[0], ft.obs_anchor[1], ft.obs_anchor[2]};
    const double inv_d = idp == 3 ? ft.inv_param[2] : ft.inv_depth;
    const M3 R_bk2w = quat_to_rot(k.q), R_w2bk = tr(R_bk2w);
    const M3 R_w2ck = mul(R_b2c, R_w2bk); const V3 t_ck_w = ld(k.p) + mv(R_bk2w, t_c_b);
    const M3 R_ba2w = quat_to_rot(a.q), R_w2ba = tr(R_ba2w);
    const M3 R_w2ca = nui ? tr(quat_to_rot(a.q_cam)) : mul(R_b2c, R_w2ba);
    const V3 p_w = ld(ft.pos), p_fej = ld(ft.pos_fej);
    V3 p_ca;
    if (fej_a) p_ca = mv(R_b2c, mv(R_w2ba, p_fej - ld(a.p_fej)) - t_c_b);
    else p_ca = {f_an.x / inv_d, f_an.y / inv_d, 1.0 / inv_d};
    const Obs& o = ft.obs.at(sid);
    const V3 p_ck = mv(R_w2ck, p_w - t_ck_w);
    J.r[0] = o.z[0] - p_ck.x / p_ck.z; J.r[1] = o.z[1] - p_ck.y / p_ck.z;
    for (int a_ = 0; a_ < 2; ++a_) { for (int q = 0; q < 3; ++q) J.hf[a_][q] = 0.0; for (int q = 0; q < 6; ++q) { J.ha[a_][q] = 0.0; J.hx[a_][q] = 0.0; J.he[a_][q] = 0.0; } }
    if (idp == 3 && sid == ft.anchor) { J.hf[0][0] = 1.0; J.hf[1][1] = 1.0; return; }        // the anchor's own observation (:1065-1073)
    const double Jk[2][3] = {{1 / p_ck.z, 0, -p_ck.x / (p_ck.z * p_ck.z)}, {0, 1 / p_ck.z, -p_ck.y / (p_ck.z * p_ck.z)}};
    const V3 J_d = mv(R_w2ck, mtv(R_w2ca, f_an));
    const V3 p_baf_w = fej_a ? (p_fej - ld(a.p_fej)) : (p_w - ld(a.p));
    const V3 p_bkf_w = if_FEJ ? (p_fej - ld(k.p_fej)) : (p_w - ld(k.p));
    const M3 Jxa_l = scl(mul(R_w2ck, skew(p_baf_w)), -1.0);
    const M3 Jxk_l = mul(R_w2ck, skew(p_bkf_w));
    const M3 Rka = mul(R_w2bk, R_ba2w);
    const M3 Sk = skew(mv(R_w2bk, p_bkf_w) - t_c_b);
    const M3 Mx = mul(Rka, skew(mtv(R_b2c, p_ca)));
    const M3 Je_l = mul(R_b2c, sub(Sk, Mx));
    const M3 Je_r = mul(R_b2c, sub(Rka, eye3()));
    const double J_rho = -1.0 / (inv_d * inv_d);
    const double jd[3] = {J_d.x, J_d.y, J_d.z};
    if (idp == 3) {            // H_f = J_k (R_w2ck R_w2ca^T) J_f, J_f = d(p_ca)/d(invParam) = [I | -f0/f2, -f1/f2, -1/f2] / f2   (:1075-1114)
      const M3 Jp = mul(R_w2ck, tr(R_w2ca));
      const double f0 = ft.inv_param[0], f1 = ft.inv_param[1], f2 = ft.inv_param[2];
      double Jf[3][3] = {{1.0, 0.0, -f0 / f2}, {0.0, 1.0, -f1 / f2}, {0.0, 0.0, -1.0 / f2}};
      for (int i = 0; i < 3; ++i) for (int q = 0; q < 3; ++q) Jf[i][q] = Jf[i][q] / f2;
      for (int a_ = 0; a_ < 2; ++a_) {
        double kp[3];
        for (int q = 0; q < 3; ++q) kp[q] = Jk[a_][0] * Jp.m[q] + Jk[a_][1] * Jp.m[3 + q] + Jk[a_][2] * Jp.m[6 + q];
        for (int q = 0; q < 3; ++q) J.hf[a_][q] = kp[0] * Jf[0][q] + kp[1] * Jf[1][q] + kp[2] * Jf[2][q];
      }
    }
    for (int a_ = 0; a_ < 2; ++a_) {
      if (idp == 1) J.hf[a_][0] = (Jk[a_][0] * jd[0] + Jk[a_][1] * jd[1] + Jk[a_][2] * jd[2]) * J_rho;
      for (int cc = 0; cc < 3; ++cc) {
        double xa = 0, xr = 0, kl = 0, kr = 0, el = 0, er = 0;
        for (int q = 0; q < 3; ++q) {
          xa += Jk[a_][q] * Jxa_l.m[q * 3 + cc]; xr += Jk[a_][q] * R_w2ck.m[q * 3 + cc];
          kl += Jk[a_][q] * Jxk_l.m[q * 3 + cc]; kr += Jk[a_][q] * -R_w2ck.m[q * 3 + cc];
          el += Jk[a_][q] * Je_l.m[q * 3 + cc]; er += Jk[a_][q] * Je_r.m[q * 3 + cc];
        }
        J.ha[a_][cc] = xa; J.ha[a_][3 + cc] = xr; J.hx[a_][cc] = kl; J.hx[a_][3 + cc] = kr; J.he[a_][cc] = el; J.he[a_][3 + cc] = er;
      }
    }
  }
  int window_index(long long sid) const { return (int)std::distance(aug.begin(), aug.find(sid)); }
  int feature_index(long long fid) const { return (int)(std::find(fstates.begin(), fstates.end(), fid) - fstates.begin()); }

  // ---------------------------------------------------------------- featureJacobian_ekf_new :1247-1338 (columns: state + one per feature in fstates)
  void feature_jacobian_ekf_new(const Feature& ft, Mat& H, std::vector<double>& r) const {
    std::vector<long long> valid;
    for (auto& o : ft.obs) if (idp == 3 || o.first != ft.anchor) valid.push_back(o.first);      // 1-D: the anchor's own observation is not used (:1260-1262)
    const int ncol = LEG + 6 * (int)aug.size() + idp * (int)fstates.size() + 6 * (int)nui_ids.size();
    H = Mat(2 * (int)valid.size(), ncol); r.assign(2 * valid.size(), 0.0);
    // the features this call adds are not in the covariance yet: their columns follow it, i.e. the nuisance block (:1291-1300)
    const int num_old = (int)fstates.size() - (ncol - P.r) / idp;
    const int a_idx = LEG + 6 * window_index(ft.anchor), f_idx = P.r + idp * (feature_index(ft.id) - num_old);
    int k = 0;
    for (long long sid : valid) {
      Jac1 J; meas_jacobian_idp(sid, ft, J);
      const int cidx = LEG + 6 * window_index(sid);
      for (int a_ = 0; a_ < 2; ++a_) {
        double* row = H.row(k + a_);
        for (int q = 0; q < idp; ++q) row[f_idx + q] = J.hf[a_][q];
        for (int q = 0; q < 6; ++q) row[a_idx + q] = J.ha[a_][q];
        for (int q = 0; q < 6; ++q) row[cidx + q] = J.hx[a_][q];
        for (int q = 0; q < 6; ++q) row[15 + q] = J.he[a_][q];
        if (c.estimate_td != 0) row[21] = ft.obs.at(sid).vel[a_];
        r[k + a_] = J.r[a_];
      }
      k += 2;
    }
  }
  // ---------------------------------------------------------------- featureJacobian_ekf :1341-1417
  void feature_jacobian_ekf(const Feature& ft, Mat& H, std::vector<double>& r) const {
    const long long sid = s.id;
    H = Mat(2, P.r); r.assign(2, 0.0);
    Jac1 J; meas_jacobian_idp(sid, ft, J);
    const int f_idx = LEG + 6 * (int)aug.size() + idp * feature_index(ft.id), cidx = LEG + 6 * window_index(sid);
    int a_idx = LEG + 6 * window_index(ft.anchor);
    if (anchor_is_nui(ft)) {                             // :1351-1366: the anchor's columns are in the nuisance block
      const int num_new = (LEG + 6 * (int)aug.size() + idp * (int)fstates.size() + 6 * (int)nui_ids.size() - P.r) / idp;
      a_idx = LEG + 6 * (int)aug.size() + idp * ((int)fstates.size() - num_new) + 6 * (int)(std::find(nui_ids.begin(), nui_ids.end(), ft.anchor) - nui_ids.begin());
    }
    for (int a_ = 0; a_ < 2; ++a_) {
      double* row = H.row(a_);
      for (int q = 0; q < idp; ++q) row[f_idx + q] = J.hf[a_][q];
      for (int q = 0; q < 6; ++q) row[a_idx + q] = J.ha[a_][q];
      for (int q = 0; q < 6; ++q) row[cidx + q] = J.hx[a_][q];
      for (int q = 0; q < 6; ++q) row[15 + q] = J.he[a_][q];
      if (c.estimate_td != 0) row[21] = ft.obs.at(sid).vel[a_];
      r[a_] = J.r[a_];
    }
  }
  // ---------------------------------------------------------------- rmLostFeaturesCov :3296-3348
  void rm_lost_features_cov(const std::vector<long long>& lost) {
    for (long long fid : lost) {
      const int seq = feature_index(fid), i0 = LEG + 6 * (int)aug.size() + idp * seq;
      std::vector<int> keep; for (int i = 0; i < P.r; ++i) if (i < i0 || i >= i0 + idp) keep.push_back(i);
      P = permuted(P, keep);
      fstates.erase(fstates.begin() + seq);
      if (schmidt()) {                                                // :3329-3339
        const long long an = map.at(fid).anchor;
        auto it = nui_features.find(an);
        if (it != nui_features.end() && std::find(nui_ids.begin(), nui_ids.end(), an) != nui_ids.end())
          it->second.erase(std::find(it->second.begin(), it->second.end(), fid));
      }
      map.erase(fid);
    }
  }
  // ---------------------------------------------------------------- rmUselessNuisanceState :3850-3895
  void rm_useless_nuisance() {
    std::vector<long long> rm; for (long long id : nui_ids) if (nui_features[id].empty()) rm.push_back(id);
    for (long long id : rm) {
      const int seq = (int)(std::find(nui_ids.begin(), nui_ids.end(), id) - nui_ids.begin());
      const int n0 = LEG + 6 * (int)aug.size() + idp * (int)fstates.size() + 6 * seq;
      std::vector<int> keep; for (int i = 0; i < P.r; ++i) if (i < n0 || i >= n0 + 6) keep.push_back(i);
      P = permuted(P, keep);
      nui_ids.erase(nui_ids.begin() + seq); nui_states.erase(id); nui_features.erase(id);
    }
  }
  // the nuisance block of P keeps its prior through an update (:1579-1589, :1805-1814, :2940-2950)
  Mat nuisance_block(int n0) const { const int nn = 6 * (int)nui_ids.size(); Mat B(nn, nn); for (int i = 0; i < nn; ++i) for (int j = 0; j < nn; ++j) B(i, j) = P(n0 + i, n0 + j); return B; }
  void restore_nuisance_block(int n0, const Mat& B) { for (int i = 0; i < B.r; ++i) for (int j = 0; j < B.c; ++j) P(n0 + i, n0 + j) = B(i, j); }
  // ---------------------------------------------------------------- updateGridMap :3351-3370
  int grid_code(const double* xy) const { const int row = (int)((xy[1] - c.y_min) / c.grid_h), col = (int)((xy[0] - c.x_min) / c.grid_w); return row * (int)c.grid_cols + col; }
  void update_grid_map() {
    const int cells = (int)(c.grid_rows * c.grid_cols);
    if (cells == 0) return;
    for (int i = 0; i < cells; ++i) grid[i] = 0;                       // only these: cells outside the range keep their count for ever
    for (long long fid : fstates) grid[grid_code(map.at(fid).obs.at(s.id).z)] += 1;
  }
  // ---------------------------------------------------------------- getNewAnchorId :3412-3472
  long long new_anchor_id(const Feature& ft, const std::vector<long long>& involved) const {
    std::vector<long long> order; for (auto& kv : aug) order.push_back(kv.first);
    const int size = (int)order.size();
    if (size <= 2) return order.back();
    long long best = -1; double min_dis = 99999.0;
    for (int i = 0; i < size - 2; ++i) {
      const long long sid = order[i];
      if (!ft.obs.count(sid) || std::find(involved.begin(), involved.end(), sid) != involved.end()) continue;
      const Aug& a = aug.at(sid);
      const V3 pn = mtv(quat_to_rot(a.q_cam), ld(ft.pos) - ld(a.p_cam));
      const Obs& o = ft.obs.at(sid);
      const double dx = pn.x / pn.z - o.z[0], dy = pn.y / pn.z - o.z[1];
      const double dis = std::sqrt(dx * dx + dy * dy);
      if (min_dis > dis) { min_dis = dis; best = sid; }
    }
    return best >= 0 ? best : order.back();
  }
  // ---------------------------------------------------------------- updateFeatureCov_1didp :3125-3293
  void update_feature_cov_1didp(const Feature& ft, long long old_id, long long new_id) {
    const int N = (int)aug.size();
    const V3 p_w = ld(ft.pos), p_fej = ld(ft.pos_fej);
    const M3 R_b2c = s.R_ic; const V3 t_c_b = ld(s.t_ci);
    const Aug& o = aug.at(old_id); const Aug& n = aug.at(new_id);
    const M3 R_b2w_old = quat_to_rot(o.q), R_c2w_old = quat_to_rot(o.q_cam);
    V3 p_old;
    if (if_FEJ) p_old = mv(R_b2c, mtv(R_b2w_old, p_fej - ld(o.p_fej)) - t_c_b);
    else p_old = mtv(R_c2w_old, p_w - ld(o.p_cam));
    const V3 p_old_ = mtv(R_c2w_old, p_w - ld(o.p_cam));
    const double inv_old = 1.0 / p_old_.z;
    const V3 f_old = {p_old_.x / p_old_.z, p_old_.y / p_old_.z, 1.0};
    const M3 R_b2w_new = quat_to_rot(n.q), R_w2b_new = tr(R_b2w_new);
    const M3 R_w2c_new = tr(quat_to_rot(n.q_cam));
    const double inv_new = ft.inv_depth;
    V3 pbo, pbn;
    if (if_FEJ) { pbo = p_fej - ld(o.p_fej); pbn = p_fej - ld(n.p_fej); }
    else { pbo = p_w - ld(o.p); pbn = p_w - ld(n.p); }
    const double Jr = -inv_new * inv_new;
    const double J_d = mv(R_w2c_new, mv(R_c2w_old, f_old)).z;
    const M3 Jto = scl(mul(R_w2c_new, skew(pbo)), -1.0);
    const M3 Jtn = mul(R_w2c_new, skew(pbn));
    const M3 Sk = skew(mv(R_w2b_new, pbn) - t_c_b);
    const M3 Rno = mul(R_w2b_new, R_b2w_old);
    const M3 Mx = mul(Rno, skew(mtv(R_b2c, p_old)));
    const M3 Jet = mul(R_b2c, sub(Sk, Mx));
    const M3 Jep = mul(R_b2c, sub(Rno, eye3()));
    const int d = P.r;
    std::vector<double> J(d, 0.0);
    const int oc = window_index(old_id), nc = window_index(new_id), fi = LEG + 6 * N + feature_index(ft.id);
    J[fi] = Jr * J_d * (-1.0 / (inv_old * inv_old));
    for (int q = 0; q < 3; ++q) {
      J[LEG + 6 * oc + q] = Jr * Jto.m[6 + q]; J[LEG + 6 * oc + 3 + q] = Jr * R_w2c_new.m[6 + q];
    }
    for (int q = 0; q < 3; ++q) {                      // += not =: with old == new the reference's second assignment overwrites; they never coincide here
      J[LEG + 6 * nc + q] = Jr * Jtn.m[6 + q]; J[LEG + 6 * nc + 3 + q] = Jr * -R_w2c_new.m[6 + q];
    }
    for (int q = 0; q < 3; ++q) { J[15 + q] = Jr * Jet.m[6 + q]; J[18 + q] = Jr * Jep.m[6 + q]; }
    std::vector<double> Pfl(d, 0.0);
    for (int i = 0; i < d; ++i) { if (J[i] == 0.0) continue; const double* pr = P.row(i); for (int j = 0; j < d; ++j) Pfl[j] += J[i] * pr[j]; }
    double Pff = 0; for (int j = 0; j < d; ++j) Pff += Pfl[j] * J[j];
    for (int j = 0; j < d; ++j) { P(fi, j) = Pfl[j]; P(j, fi) = Pfl[j]; }
    P(fi, fi) = Pff;
  }
  // ---------------------------------------------------------------- updateFeatureCov_3didp :2965-3122, literally: the reference looks the "new" pose and
  // its column block up with old_state_id (:3000, :3066), so H_x_new overwrites H_x_old in the old block
  void update_feature_cov_3didp(const Feature& ft, long long old_id) {
    const int N = (int)aug.size();
    const V3 p_w = ld(ft.pos), p_fej = ld(ft.pos_fej);
    const M3 R_b2c = s.R_ic; const V3 t_c_b = ld(s.t_ci);
    const Aug& o = aug.at(old_id); const Aug& n = o;                       // sic
    const M3 R_b2w_old = quat_to_rot(o.q), R_c2w_old = quat_to_rot(o.q_cam);
    V3 p_old;
    if (if_FEJ) p_old = mv(R_b2c, mtv(R_b2w_old, p_fej - ld(o.p_fej)) - t_c_b);
    else p_old = mtv(R_c2w_old, p_w - ld(o.p_cam));
    const M3 R_b2w_new = quat_to_rot(n.q), R_w2b_new = tr(R_b2w_new);
    const M3 R_w2c_new = tr(quat_to_rot(n.q_cam));
    const double* iv = ft.inv_param;
    V3 pbo, pbn;
    if (if_FEJ) { pbo = p_fej - ld(o.p_fej); pbn = p_fej - ld(n.p_fej); }
    else { pbo = p_w - ld(o.p); pbn = p_w - ld(n.p); }
    double Jfp[3][3] = {{1.0, 0.0, -iv[0]}, {0.0, 1.0, -iv[1]}, {0.0, 0.0, -iv[2]}};
    for (int i = 0; i < 3; ++i) for (int q = 0; q < 3; ++q) Jfp[i][q] = iv[2] * Jfp[i][q];
    const M3 Jp = mul(R_w2c_new, R_c2w_old);
    const M3 Jxo_l = scl(mul(R_w2c_new, skew(pbo)), -1.0);               // J_x_old = [-R skew(p_bf_old) | R]
    const M3 Jxn_l = mul(R_w2c_new, skew(pbn));                         // J_x_new = [R skew(p_bf_new) | -R]
    const M3 Sk = skew(mv(R_w2b_new, pbn) - t_c_b);
    const M3 Rno = mul(R_w2b_new, R_b2w_old);
    const M3 Mx = mul(Rno, skew(mtv(R_b2c, p_old)));
    const M3 Jet = mul(R_b2c, sub(Sk, Mx));
    const M3 Jep = mul(R_b2c, sub(Rno, eye3()));
    double Jpf[3][3] = {{1.0, 0.0, -p_old.x}, {0.0, 1.0, -p_old.y}, {0.0, 0.0, -p_old.z}};
    for (int i = 0; i < 3; ++i) for (int q = 0; q < 3; ++q) Jpf[i][q] = p_old.z * Jpf[i][q];
    const int d = P.r;
    Mat J(3, d);
    const int oc = window_index(old_id), fi = LEG + 6 * N + 3 * feature_index(ft.id);
    auto m3rows = [](const double A[3][3], const M3& B, int a_, int q) { return A[a_][0] * B.m[q] + A[a_][1] * B.m[3 + q] + A[a_][2] * B.m[6 + q]; };
    for (int a_ = 0; a_ < 3; ++a_) {
      double fp[3]; for (int q = 0; q < 3; ++q) fp[q] = m3rows(Jfp, Jp, a_, q);
      for (int q = 0; q < 3; ++q) J(a_, fi + q) = fp[0] * Jpf[0][q] + fp[1] * Jpf[1][q] + fp[2] * Jpf[2][q];
      (void)Jxo_l;                                                      // H_x_old is written first and then overwritten by H_x_new (same block)
      for (int q = 0; q < 3; ++q) { J(a_, LEG + 6 * oc + q) = m3rows(Jfp, Jxn_l, a_, q); J(a_, LEG + 6 * oc + 3 + q) = -m3rows(Jfp, R_w2c_new, a_, q); }
      for (int q = 0; q < 3; ++q) { J(a_, 15 + q) = m3rows(Jfp, Jet, a_, q); J(a_, 18 + q) = m3rows(Jfp, Jep, a_, q); }
    }
    Mat Pfl(3, d);
    for (int a_ = 0; a_ < 3; ++a_) for (int i = 0; i < d; ++i) { const double jv = J(a_, i); if (jv == 0.0) continue; const double* pr = P.row(i); double* o_ = Pfl.row(a_); for (int j = 0; j < d; ++j) o_[j] += jv * pr[j]; }
    double Pff[3][3];
    for (int a_ = 0; a_ < 3; ++a_) for (int b = 0; b < 3; ++b) { double acc = 0; for (int j = 0; j < d; ++j) acc += Pfl(a_, j) * J(b, j); Pff[a_][b] = acc; }
    for (int a_ = 0; a_ < 3; ++a_) for (int j = 0; j < d; ++j) if (j < fi || j >= fi + 3) { P(fi + a_, j) = Pfl(a_, j); P(j, fi + a_) = Pfl(a_, j); }
    for (int a_ = 0; a_ < 3; ++a_) for (int b = 0; b < 3; ++b) P(fi + a_, fi + b) = Pff[a_][b];
    for (int a_ = 0; a_ < 3; ++a_) for (int b = a_ + 1; b < 3; ++b) { const double m_ = 0.5 * (P(fi + a_, fi + b) + P(fi + b, fi + a_)); P(fi + a_, fi + b) = m_; P(fi + b, fi + a_) = m_; }
  }
  // ---------------------------------------------------------------- measurementUpdate_hybrid :1605-1862
  // H_new: rows [0, nn) = the rows that define the nn new states (H_1 | H_2 upper triangular | r_1), rows [nn, ..) = their null-space rows
  void update_hybrid(const Mat& H_new, const std::vector<double>& r_new, int nn, const Mat& H_ekf, const std::vector<double>& r_ekf,
                     const Mat& H_msckf, const std::vector<double>& r_msckf) {
    const int d = P.r;
    const int k = H_new.r - nn;
    if (r_new.size() + r_ekf.size() + r_msckf.size() == 0) return;
    Mat H_o; std::vector<double> r_o; H_o.c = d;
    auto add_rows = [&](const Mat& H, const std::vector<double>& r, int first) {
      for (int i = first; i < H.r; ++i) { H_o.a.insert(H_o.a.end(), H.row(i), H.row(i) + d); H_o.r += 1; r_o.push_back(r[i]); }
    };
    { Mat Hm(H_msckf.r, d); for (int i = 0; i < H_msckf.r; ++i) memcpy(Hm.row(i), H_msckf.row(i), sizeof(double) * H_msckf.c); add_rows(Hm, r_msckf, 0); }
    add_rows(H_ekf, r_ekf, 0);
    if (k > 0) { Mat Hn(H_new.r, d); for (int i = 0; i < H_new.r; ++i) memcpy(Hn.row(i), H_new.row(i), sizeof(double) * d); add_rows(Hn, r_new, nn); }
    std::vector<double> dx_leg(d, 0.0); Mat Y;
    const bool have = solve_update(H_o, r_o, nullptr, dx_leg, Y);
    if (!have) dx_leg.assign(d, 0.0);
    // HH = H_1 / diag(H_2): Eigen's LDLT reads the lower triangle of the triangular factor (App. C-13, :1665-1666)
    Mat HH(nn, d); std::vector<double> dx = dx_leg;
    for (int i = 0; i < nn; ++i) {
      const double h2 = H_new(i, d + i);
      double acc = 0;
      for (int j = 0; j < d; ++j) { HH(i, j) = H_new(i, j) / h2; acc += HH(i, j) * dx_leg[j]; }
      dx.push_back(-acc + r_new[i] / h2);
    }
    inject(dx, d, (int)fstates.size() - nn / idp);
    const int nui_cols = schmidt() ? 6 * (int)nui_ids.size() : 0;
    if (have) {
      if (nui_cols) {
        const int n0 = LEG + 6 * (int)aug.size() + idp * ((int)fstates.size() - nn / idp);
        const Mat B = nuisance_block(n0);
        apply_cov(Y);
        restore_nuisance_block(n0, B);
      } else apply_cov(Y);
    }
    if (nn) {
      Mat nHHP(nn, d);
      for (int i = 0; i < nn; ++i) { double* o_ = nHHP.row(i); for (int q = 0; q < d; ++q) { const double h = HH(i, q); if (h == 0.0) continue; const double* pr = P.row(q); for (int j = 0; j < d; ++j) o_[j] -= h * pr[j]; } }
      // (H_2^T H_2)^-1 through the inverse of the upper-triangular factor (the FULL factor here, :1823-1825)
      Mat Ri(nn, nn);
      for (int j = 0; j < nn; ++j) {
        Ri(j, j) = 1.0 / H_new(j, d + j);
        for (int i = j - 1; i >= 0; --i) { double acc = 0; for (int q = i + 1; q <= j; ++q) acc += H_new(i, d + q) * Ri(q, j); Ri(i, j) = -acc / H_new(i, d + i); }
      }
      Mat Pn(d + nn, d + nn);
      for (int i = 0; i < d; ++i) memcpy(Pn.row(i), P.row(i), sizeof(double) * d);
      for (int i = 0; i < nn; ++i) for (int j = 0; j < d; ++j) { Pn(d + i, j) = nHHP(i, j); Pn(j, d + i) = nHHP(i, j); }
      for (int a_ = 0; a_ < nn; ++a_) for (int b = 0; b < nn; ++b) {
        double acc = 0; for (int q = 0; q < d; ++q) acc += nHHP(a_, q) * HH(b, q);
        double inv = 0; for (int q = 0; q < nn; ++q) inv += Ri(a_, q) * Ri(b, q);
        Pn(d + a_, d + b) = -acc + sfeat2 * inv;
      }
      for (int i = 0; i < d + nn; ++i) for (int j = i + 1; j < d + nn; ++j) { const double m_ = 0.5 * (Pn(i, j) + Pn(j, i)); Pn(i, j) = m_; Pn(j, i) = m_; }
      if (nui_cols) {                                                // :1832-1845: the new columns go in FRONT of the nuisance block
        std::vector<int> order;
        for (int i = 0; i < d - nui_cols; ++i) order.push_back(i);
        for (int i = d; i < d + nn; ++i) order.push_back(i);
        for (int i = d - nui_cols; i < d; ++i) order.push_back(i);
        P = permuted(Pn, order);
      } else P = std::move(Pn);
    }
  }

  // ---------------------------------------------------------------- checkZUPT :2751-2788 + measurementUpdate_ZUPT_vpq :2791-2962
  bool check_zupt() {
    std::vector<double> dd; dd.swap(coarse);
    if (dd.size() < 20) return false;
    std::sort(dd.begin(), dd.end());
    if (dd[dd.size() - 9] < c.zupt_max_feature_dis) {
      ++zupt_events;
      if (!fstates.empty()) {                                        // :2770-2782: every SLAM feature leaves the state
        const int nd = P.r - idp * (int)fstates.size();
        std::vector<int> keep; for (int i = 0; i < nd; ++i) keep.push_back(i);
        P = permuted(P, keep);
        for (long long fid : fstates) { Feature& ft = map.at(fid); ft.init = false; ft.ekf = false; ft.in_state = false; }
        fstates.clear();
      }
      const int N = (int)aug.size(), d = P.r;
      Mat H(9, d); std::vector<double> r(9);
      for (int i = 0; i < 3; ++i) {
        H(i, 3 + i) = 1.0;
        H(3 + i, LEG + 6 * N - 3 + i) = 1.0; H(3 + i, LEG + 6 * N - 9 + i) = -1.0;
        H(6 + i, LEG + 6 * N - 6 + i) = -0.5; H(6 + i, LEG + 6 * N - 12 + i) = 0.5;
      }
      const Aug& cur = aug.at(s.id); const Aug& prv = aug.at(s.id - 1);
      for (int i = 0; i < 3; ++i) { r[i] = -s.v[i]; r[3 + i] = -(cur.p[i] - prv.p[i]); }
      const double qc[4] = {-prv.q[0], -prv.q[1], -prv.q[2], prv.q[3]};
      double qe[4]; quat_mul(cur.q, qc, qe);
      for (int i = 0; i < 3; ++i) r[6 + i] = qe[i];
      const double nv = c.zupt_noise_v * c.zupt_noise_v, np_ = c.zupt_noise_p * c.zupt_noise_p, nq = c.zupt_noise_q * c.zupt_noise_q;
      const double Rd[9] = {nv, nv, nv, np_, np_, np_, nq, nq, nq};
      update(H, r, Rd);
      last_zupt = s.time;
      return true;
    }
    return false;
  }

  // ---------------------------------------------------------------- removeLostFeatures :1883-2256
  void remove_lost_features() {
    const long long sid_now = s.id;
    std::vector<long long> invalid, msckf_ids, lost_ids, ekf_new, ekf_lost, ekf_ids;
    for (auto& kv : map) if (kv.second.in_state) (kv.second.obs.count(sid_now) ? ekf_ids : ekf_lost).push_back(kv.first);
    rm_lost_features_cov(ekf_lost);
    if (schmidt()) rm_useless_nuisance();                             // :1920-1921
    update_grid_map();
    const bool hyb = hybrid();
    for (auto& kv : map) {
      Feature& ft = kv.second;
      if (ft.in_state) continue;
      const bool tracked_now = ft.obs.count(sid_now) != 0;
      if (!tracked_now) {
        if ((int)ft.obs.size() < (int)c.least_obs) { invalid.push_back(kv.first); continue; }
        if (!ft.init) {
          if (!check_motion(ft, tracked_now)) { invalid.push_back(kv.first); continue; }
          if (!initialize_position(ft, sid_now, true)) { invalid.push_back(kv.first); continue; }
        }
        msckf_ids.push_back(kv.first); lost_ids.push_back(kv.first);
      } else {
        if (!((int)ft.obs.size() >= (int)c.max_track_len)) continue;
        const int code = hyb ? grid_code(ft.obs.at(sid_now).z) : 0;
        if (hyb && grid[code] < (int)c.max_features && s.time - last_zupt > 5 &&
            (int)(fstates.size() + ekf_new.size()) < (int)(c.max_features * c.grid_rows * c.grid_cols)) {      // :1968-1990
          if (!ft.ekf) { ft.init = false; if (check_motion(ft, tracked_now)) initialize_inv_param(ft, sid_now); }
          if (!ft.init) continue;
          ekf_new.push_back(kv.first);
          grid[code] += 1;
        } else {
          if (!ft.init) { if (check_motion(ft, tracked_now)) initialize_position(ft, sid_now, true); }
          if (!ft.init) continue;
          msckf_ids.push_back(kv.first); lost_ids.push_back(kv.first);
        }
      }
    }
    for (long long fid : invalid) map.erase(fid);
    if (msckf_ids.empty() && ekf_new.empty() && ekf_ids.empty()) return;
    if (!if_ZUPT) {
      const int d = P.r;
      // ---- new EKF-SLAM features (:2019-2125)
      for (long long fid : ekf_new) { map.at(fid).in_state = true; fstates.push_back(fid); }
      const int n_all = (int)ekf_new.size();
      std::vector<long long> kept; std::vector<Mat> Hn; std::vector<std::vector<double>> rn;
      for (long long fid : ekf_new) {
        Feature& ft = map.at(fid);
        std::vector<long long> sids; for (auto& o : ft.obs) sids.push_back(o.first);
        Mat Hj; std::vector<double> rj; feature_jacobian_ekf_new(ft, Hj, rj);
        Mat Hm; std::vector<double> rm; feature_jacobian(ft, sids, Hm, rm);
        if (Hm.r > 0 && gating(Hm, rm, 2 * (int)sids.size() - 3)) { kept.push_back(fid); Hn.push_back(std::move(Hj)); rn.push_back(std::move(rj)); }
        else ft.in_state = false;
      }
      fstates.resize(fstates.size() - n_all);
      for (long long fid : kept) fstates.push_back(fid);
      const int nn = idp * (int)kept.size();
      Mat H_new; std::vector<double> r_new;
      if (nn) {
        // the columns of the features that failed the gate disappear; then Householder reflections on the nn own columns:
        // rows [0, nn) <- column space (H_1 | H_2 triangular | r_1), the rest <- left null space (any orthonormal bases give the same update)
        int rows = 0; for (auto& h : Hn) rows += h.r;
        H_new = Mat(rows, d + nn); r_new.assign(rows, 0.0);
        int r0 = 0;
        for (int f = 0; f < (int)kept.size(); ++f) {
          const int src_col = d + idp * (int)(std::find(ekf_new.begin(), ekf_new.end(), kept[f]) - ekf_new.begin());
          for (int i = 0; i < Hn[f].r; ++i) {
            memcpy(H_new.row(r0 + i), Hn[f].row(i), sizeof(double) * d);
            for (int q = 0; q < idp; ++q) H_new(r0 + i, d + idp * f + q) = Hn[f](i, src_col + q);
            r_new[r0 + i] = rn[f][i];
          }
          r0 += Hn[f].r;
        }
        std::vector<double> v(rows);
        for (int k = 0; k < nn; ++k) {
          double n2 = 0; for (int i = k; i < rows; ++i) n2 += H_new(i, d + k) * H_new(i, d + k);
          const double nrm = std::sqrt(n2);
          if (nrm == 0.0) continue;
          const double alpha = H_new(k, d + k) >= 0 ? -nrm : nrm;
          double vtv = 0;
          for (int i = k; i < rows; ++i) { v[i] = H_new(i, d + k) - (i == k ? alpha : 0.0); vtv += v[i] * v[i]; }
          if (vtv == 0.0) continue;
          const double beta = 2.0 / vtv;
          for (int cc = 0; cc < d + nn; ++cc) {
            double dt_ = 0; for (int i = k; i < rows; ++i) dt_ += v[i] * H_new(i, cc);
            if (dt_ == 0.0) continue;
            dt_ *= beta;
            for (int i = k; i < rows; ++i) H_new(i, cc) -= dt_ * v[i];
          }
          double dt_ = 0; for (int i = k; i < rows; ++i) dt_ += v[i] * r_new[i];
          dt_ *= beta;
          for (int i = k; i < rows; ++i) r_new[i] -= dt_ * v[i];
        }
      } else H_new = Mat(0, d);
      // ---- in-state EKF-SLAM features (:2127-2177)
      Mat H_ekf; std::vector<double> r_ekf; H_ekf.c = d;
      for (long long fid : ekf_ids) {
        Mat Hj; std::vector<double> rj; feature_jacobian_ekf(map.at(fid), Hj, rj);
        if (gating(Hj, rj, 2)) append(H_ekf, r_ekf, Hj, rj);
      }
      if (H_ekf.r > H_ekf.c) compress(H_ekf, r_ekf);
      // ---- MSCKF features (:2179-2233)
      Mat H; std::vector<double> r; H.c = d;
      for (long long fid : msckf_ids) {
        const Feature& ft = map.at(fid);
        std::vector<long long> sids; for (auto& o : ft.obs) sids.push_back(o.first);
        Mat Hj; std::vector<double> rj;
        feature_jacobian(ft, sids, Hj, rj);
        if (Hj.r > 0 && gating(Hj, rj, 2 * (int)sids.size() - 3)) append(H, r, Hj, rj);
      }
      if (H.r > LEG + 6 * (int)aug.size()) compress(H, r);
      if (!hyb) update(H, r);
      else update_hybrid(H_new, r_new, nn, H_ekf, r_ekf, H, r);
    } else {
      for (long long fid : msckf_ids) map.at(fid).init = false;
    }
    for (long long fid : lost_ids) map.erase(fid);
  }

  // ---------------------------------------------------------------- findRedundantImuStates :2259-2307
  std::vector<long long> find_redundant() const {
    std::vector<long long> ids; for (auto& kv : aug) ids.push_back(kv.first);
    int key_i = (int)ids.size() - 4, st_i = key_i + 1, first_i = 0;
    const Aug& key = aug.at(ids[key_i]);
    const M3 key_R = quat_to_rot(key.q_cam);
    std::vector<long long> rm;
    for (int k = 0; k < 2; ++k) {
      const Aug& a = aug.at(ids[st_i]);
      const M3 Rr = tr(quat_to_rot(a.q_cam));
      const double dist = norm(ld(a.p_cam) - ld(key.p_cam));
      double qq[4]; rot_to_quat(mul(Rr, key_R), qq);
      const double n = std::sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2]);
      const double angle = 2 * std::atan2(n, std::fabs(qq[3]));
      if (angle < c.rotation_threshold && dist < c.translation_threshold && tracking_rate > c.tracking_rate_threshold) { rm.push_back(ids[st_i]); ++st_i; }
      else { rm.push_back(ids[first_i]); ++first_i; st_i -= 2; }
    }
    std::sort(rm.begin(), rm.end());
    return rm;
  }

  // ---------------------------------------------------------------- pruneImuStateBuffer :2310-2641 (pure MSCKF)
  void prune() {
    std::vector<long long> rm_ids;
    if (!if_ZUPT) {
      if ((int)aug.size() < (int)c.sw_size) return;
      rm_ids = find_redundant();
    } else rm_ids.push_back(s.id - 1);
    const long long sid_now = s.id;
    std::vector<long long> used, new_nui;
    auto involved_of = [&](const Feature& ft) { std::vector<long long> v; for (long long sid : rm_ids) if (ft.obs.count(sid)) v.push_back(sid); return v; };
    for (auto& kv : map) {
      Feature& ft = kv.second;
      const std::vector<long long> inv = involved_of(ft);
      if (inv.empty()) continue;
      const bool anchor_goes = std::find(inv.begin(), inv.end(), ft.anchor) != inv.end();
      if (ft.in_state) {                                                      // :2345-2405: hand the anchor over
        if (anchor_goes && schmidt() && s.id - ft.anchor > 2) {               // :2351-2358: a mature anchor becomes a nuisance state
          nui_features[ft.anchor].push_back(kv.first);
          if (std::find(new_nui.begin(), new_nui.end(), ft.anchor) == new_nui.end()) new_nui.push_back(ft.anchor);
          continue;
        }
        if (anchor_goes) {
          const long long new_id = idp == 3 ? s.id : new_anchor_id(ft, inv);          // 3-D: always the newest state (:2361-2378)
          const Aug& a = aug.at(new_id);
          const V3 pn = mtv(quat_to_rot(a.q_cam), ld(ft.pos) - ld(a.p_cam));
          if (idp == 3) {
            ft.inv_param[0] = pn.x / pn.z; ft.inv_param[1] = pn.y / pn.z; ft.inv_param[2] = 1.0 / pn.z;
            update_feature_cov_3didp(ft, ft.anchor);
          } else {
            ft.inv_depth = 1.0 / pn.z;
            ft.obs_anchor[0] = pn.x / pn.z; ft.obs_anchor[1] = pn.y / pn.z;
            update_feature_cov_1didp(ft, ft.anchor, new_id);
          }
          ft.anchor = new_id;
        }
        continue;
      }
      if (hybrid() && ft.init && anchor_goes) {                               // :2407-2460: potential features are only re-anchored
        const long long new_id = idp == 3 ? s.id : new_anchor_id(ft, inv);
        const Aug& a = aug.at(new_id);
        const V3 pn = mtv(quat_to_rot(a.q_cam), ld(ft.pos) - ld(a.p_cam));
        if (idp == 3) { ft.inv_param[0] = pn.x / pn.z; ft.inv_param[1] = pn.y / pn.z; ft.inv_param[2] = 1.0 / pn.z; }
        else {
          ft.inv_depth = 1.0 / pn.z;
          ft.obs_anchor[0] = ft.obs.at(new_id).z[0]; ft.obs_anchor[1] = ft.obs.at(new_id).z[1];
        }
        ft.anchor = new_id;
      }
      if (!if_ZUPT && !ft.ekf && inv.size() > 1) {
        const bool tracked = ft.obs.count(sid_now) != 0;
        if (!ft.init) {
          if (!check_motion(ft, tracked)) continue;
          if (!initialize_position(ft, 0, false)) continue;
        }
        used.push_back(kv.first);
      }
    }
    if (!if_ZUPT && !used.empty()) {
      Mat H; std::vector<double> r;
      H.c = P.r;
      size_t ui = 0;
      for (auto& kv : map) {
        Feature& ft = kv.second;
        const std::vector<long long> inv = involved_of(ft);
        if (ui < used.size() && used[ui] == kv.first) {
          ++ui;
          Mat Hj; std::vector<double> rj;
          feature_jacobian(ft, inv, Hj, rj);
          if (Hj.r > 0 && gating(Hj, rj, 2 * (int)inv.size() - 3)) append(H, r, Hj, rj);
        }
        for (long long sid : inv) ft.obs.erase(sid);
      }
      if (H.r > 0) {
        if (H.r > H.c) compress(H, r);
        update(H, r);
      }
    } else {
      for (auto& kv : map) for (long long sid : rm_ids) kv.second.obs.erase(sid);
    }
    for (long long sid : rm_ids) {
      const auto ia = aug.find(sid);
      const int seq = (int)std::distance(aug.begin(), ia);
      const int a0 = LEG + 6 * seq, d = P.r;
      if (schmidt() && std::find(new_nui.begin(), new_nui.end(), sid) != new_nui.end()) {     // :2569-2613: the pose block moves behind everything else
        std::vector<int> order;
        for (int i = 0; i < d; ++i) if (i < a0 || i >= a0 + 6) order.push_back(i);
        for (int i = a0; i < a0 + 6; ++i) order.push_back(i);
        P = permuted(P, order);
        nui_ids.push_back(sid); nui_states[sid] = ia->second;
        aug.erase(ia);
        continue;
      }
      Mat Pn(d - 6, d - 6);
      for (int i = 0, ii = 0; i < d; ++i) {
        if (i >= a0 && i < a0 + 6) continue;
        const double* src = P.row(i); double* dst = Pn.row(ii++);
        memcpy(dst, src, sizeof(double) * a0);
        memcpy(dst + a0, src + a0 + 6, sizeof(double) * (d - a0 - 6));
      }
      P = std::move(Pn);
      aug.erase(ia);
    }
  }

  // ---------------------------------------------------------------- processFeatures :363-461
  int process_features(double t_msg, const long long* ids, const double* data, int n, const double* imu, int m, int* consumed) {
    *consumed = 0;
    if (!first) {
      if (m > 0 && imu[0] - t_msg - td <= 0.0) first = true;
      else return 0;
    }
    if (!gravity_set) return 0;
    *consumed = batch_imu(t_msg + td, imu, m);
    add_observations(ids, data, n);
    augment();
    if (c.if_ZUPT_valid != 0) if_ZUPT = check_zupt();
    remove_lost_features();
    prune();
    if (c.if_FEJ != 0 && !if_FEJ && s.time - take_off >= 0) if_FEJ = true;
    return 1;
  }
};

}  // namespace

extern "C" {

void* lvo_create(const double* cfg, int n) {
  if (n != (int)(sizeof(Cfg) / sizeof(double))) return nullptr;
  Cfg c; memcpy(&c, cfg, sizeof c);
  return new Filter(c);
}
void lvo_destroy(void* h) { delete (Filter*)h; }
int lvo_cfg_doubles() { return (int)(sizeof(Cfg) / sizeof(double)); }

void lvo_set_initial_state(void* h, double t, const double* q, const double* p, const double* v, const double* bg, const double* ba) {
  Filter* f = (Filter*)h;
  f->s.time = t;
  memcpy(f->s.q, q, sizeof(double) * 4); memcpy(f->s.p, p, sizeof(double) * 3); memcpy(f->s.v, v, sizeof(double) * 3);
  memcpy(f->s.bg, bg, sizeof(double) * 3); memcpy(f->s.ba, ba, sizeof(double) * 3);
  f->gravity_set = true; f->first = true /* the initialiser runs behind the gate of larvio.cpp:366-372 */; f->take_off = t; f->last_zupt = t; f->fej_now = f->s;
}

int lvo_process_features(void* h, double t_msg, const long long* ids, const double* data, int n, const double* imu, int m, int* consumed) {
  return ((Filter*)h)->process_features(t_msg, ids, data, n, imu, m, consumed);
}

// out[31] = t, q(4), p(3), v(3), bg(3), ba(3), R_imu_cam0(9), t_cam0_imu(3), td, n_window
void lvo_get_state(void* h, double* out) {
  const Filter* f = (const Filter*)h;
  out[0] = f->s.time;
  memcpy(out + 1, f->s.q, sizeof(double) * 4); memcpy(out + 5, f->s.p, sizeof(double) * 3); memcpy(out + 8, f->s.v, sizeof(double) * 3);
  memcpy(out + 11, f->s.bg, sizeof(double) * 3); memcpy(out + 14, f->s.ba, sizeof(double) * 3);
  memcpy(out + 17, f->s.R_ic.m, sizeof(double) * 9); memcpy(out + 26, f->s.t_ci, sizeof(double) * 3);
  out[29] = f->td; out[30] = (double)f->aug.size();
}
int lvo_dim(void* h) { return ((Filter*)h)->P.r; }
void lvo_get_cov(void* h, double* out) { const Filter* f = (const Filter*)h; memcpy(out, f->P.a.data(), sizeof(double) * f->P.a.size()); }
// ids of the EKF-SLAM features in the state, in covariance order; returns their number
int lvo_get_slam(void* h, long long* ids, int cap) {
  const Filter* f = (const Filter*)h;
  for (int i = 0; i < (int)f->fstates.size() && i < cap; ++i) ids[i] = f->fstates[i];
  return (int)f->fstates.size();
}
long long lvo_counter(void* h, int which) { const Filter* f = (const Filter*)h; return which == 0 ? f->updates : which == 1 ? f->zupt_events : (long long)f->map.size(); }

}  // extern "C"
