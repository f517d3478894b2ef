// Inclinometer (static-scene) initialiser, host side: StaticInitializer::tryIncInit / initializeGravityAndBias /
// assignInitialState (src/StaticInitializer.cpp:13-161) behind FlexibleInitializer::tryIncInit (src/FlexibleInitializer.cpp:
// 10-26), which LarVio::processFeatures calls until gravity is set (src/larvio.cpp:375-391).  Per sequence it is a few
// hundred flops once per run, on data the caller already holds on the host (the feature message returned by
// lvb_process_images and its own IMU buffer), so it stays host C++; its result goes to the device through
// lvb_set_initial_state.  The dynamic (SfM + Ceres) initialiser is out of scope.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <new>
#include <vector>
#include "../../include/larvio_b200.h"

int lvb_set_err(int code, const char* fmt, ...);

struct LvbStaticInit {
  double max_feature_dis;      // zupt_max_feature_dis (larvio.cpp:343-344)
  int static_num;              // (int)(static_duration * pub_frequency) (larvio.cpp:223-224)
  double td;
  int counter;                 // staticImgCounter
  double lower_time_bound;
  std::map<uint64_t, std::pair<double, double>> init_features;
  bool done;
};

extern "C" LvbStaticInit* lvb_static_init_create(const LvbConfig* cfg) {
  if (!cfg) { lvb_set_err(LVB_E_ARG, "lvb_static_init_create: null config"); return nullptr; }
  LvbStaticInit* s = new (std::nothrow) LvbStaticInit();
  if (!s) return nullptr;
  s->max_feature_dis = cfg->zupt_max_feature_dis;
  s->static_num = (int)(cfg->static_duration * cfg->pub_frequency);
  s->td = cfg->td;
  s->counter = 0; s->lower_time_bound = 0.0; s->done = false;
  return s;
}

extern "C" void lvb_static_init_destroy(LvbStaticInit* s) { delete s; }

// Eigen::Quaterniond::FromTwoVectors(a, b) (coeffs x y z w); the antiparallel branch of Eigen (an SVD of [a b]) is
// replaced by an explicit orthogonal axis - it needs an IMU lying exactly upside down.
static void from_two_vectors(const double a[3], const double b[3], double q[4]) {
  const double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  const double v0[3] = {a[0] / na, a[1] / na, a[2] / na}, v1[3] = {b[0] / nb, b[1] / nb, b[2] / nb};
  const double c = v1[0] * v0[0] + v1[1] * v0[1] + v1[2] * v0[2];
  if (c < -1.0 + 1e-12) {
    double ax[3] = {1.0, 0.0, 0.0};
    if (std::fabs(v0[0]) > 0.9) { ax[0] = 0.0; ax[1] = 1.0; }
    double o[3] = {v0[1] * ax[2] - v0[2] * ax[1], v0[2] * ax[0] - v0[0] * ax[2], v0[0] * ax[1] - v0[1] * ax[0]};
    const double no = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    q[0] = o[0] / no; q[1] = o[1] / no; q[2] = o[2] / no; q[3] = 0.0;
    return;
  }
  const double axis[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
  const double s = std::sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
  q[0] = axis[0] * invs; q[1] = axis[1] * invs; q[2] = axis[2] * invs; q[3] = s * 0.5;
}

extern "C" int lvb_static_init_try(LvbStaticInit* s, const LvbFeature* feat, int n_feat, double t_msg, const LvbImu* imu,
                                   int n_imu, double* state17, double* gyro_old3, double* acc_old3, int* n_consumed) {
  if (!s || (n_feat > 0 && !feat) || (n_imu > 0 && !imu) || !state17 || !gyro_old3 || !acc_old3 || !n_consumed)
    return lvb_set_err(LVB_E_ARG, "lvb_static_init_try: bad argument");
  *n_consumed = 0;
  if (s->counter == 0) {                                             // first image of a candidate static stretch (:16-25)
    s->counter++;
    s->init_features.clear();
    for (int i = 0; i < n_feat; ++i) s->init_features[feat[i].id] = std::make_pair(feat[i].u, feat[i].v);
    s->lower_time_bound = t_msg + s->td;
    return 0;
  }
  std::map<uint64_t, std::pair<double, double>> curr;
  std::vector<double> dis;
  for (int i = 0; i < n_feat; ++i) {                                 // (:28-38)
    curr[feat[i].id] = std::make_pair(feat[i].u, feat[i].v);
    auto it = s->init_features.find(feat[i].id);
    if (it != s->init_features.end()) {
      const double dx = feat[i].u - it->second.first, dy = feat[i].v - it->second.second;
      dis.push_back(std::sqrt(dx * dx + dy * dy));
    }
  }
  if (dis.size() < 20) { s->counter = 0; return 0; }                 // (:40-44)
  std::sort(dis.begin(), dis.end());
  const double max_dis = dis[dis.size() - 19];                       // 19th largest (:46-50)
  if (max_dis < s->max_feature_dis) {
    s->counter++;
    s->init_features.swap(curr);
    if (s->counter < s->static_num) return 0;
  } else {
    s->counter = 0;
    return 0;
  }
  // ---- initializeGravityAndBias (:78-124) with Ma = Tg = I, As = 0 (their initial values, larvio.cpp:129-131)
  const double time_bound = t_msg + s->td;
  double sw[3] = {0, 0, 0}, sa[3] = {0, 0, 0};
  int used = 0; double last_t = 0.0;
  for (int i = 0; i < n_imu; ++i) {
    const double t = imu[i].t;
    if (t < s->lower_time_bound) continue;
    if (t > time_bound) break;
    for (int k = 0; k < 3; ++k) { sw[k] += imu[i].gyro[k]; sa[k] += imu[i].acc[k]; }
    ++used; last_t = t;
  }
  if (used == 0) return lvb_set_err(LVB_E_ARG, "lvb_static_init_try: no IMU sample inside the static stretch");
  double bg[3], g_imu[3];
  for (int k = 0; k < 3; ++k) { bg[k] = sw[k] / used; g_imu[k] = sa[k] / used; }
  const double gn = std::sqrt(g_imu[0] * g_imu[0] + g_imu[1] * g_imu[1] + g_imu[2] * g_imu[2]);
  const double minus_g_world[3] = {-0.0, -0.0, gn};
  double q[4];
  from_two_vectors(g_imu, minus_g_world, q);
  state17[0] = last_t;
  for (int k = 0; k < 4; ++k) state17[1 + k] = q[k];
  for (int k = 0; k < 3; ++k) { state17[5 + k] = 0.0; state17[8 + k] = 0.0; state17[11 + k] = bg[k]; state17[14 + k] = 0.0; }
  // ---- assignInitialState (:127-158): samples up to state_time are consumed, the next one seeds m_gyro_old / m_acc_old
  int useful = 0;
  for (int i = 0; i < n_imu; ++i) { if (imu[i].t > last_t) break; ++useful; }
  if (useful >= n_imu) useful--;
  for (int k = 0; k < 3; ++k) { gyro_old3[k] = imu[useful].gyro[k]; acc_old3[k] = imu[useful].acc[k]; }
  *n_consumed = useful;
  s->done = true;
  return 1;
}
