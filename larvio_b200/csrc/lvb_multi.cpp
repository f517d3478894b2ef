// Sharding as a C-ABI feature: S independent sequences split into contiguous blocks over n_gpu devices,
// one LvbHandle and one host thread per block, no exchange step between them (sequences never interact; the reference's only
// cross-instance coupling, its static id counters, larvio.cpp:33-39, is per sequence here).  A multi-process launcher
// (bench.py under torchrun: one rank per GPU, NCCL for the input pool and the trajectory gather) is the other way to use
// several GPUs; this is the single-process one, for a replay tool or a ROS node that owns the whole node.
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include "../../include/larvio_b200.h"

int lvb_set_err(int code, const char* fmt, ...);      // lvb_capi.cu (thread-local message behind lvb_last_error)

struct LvbMulti {
  std::vector<LvbHandle*> h;      // one per shard
  std::vector<int> first;         // first sequence of shard k; first[n] = S
  int S = 0, W = 0, H = 0;
  std::vector<int> rc;            // per-shard status of the last call
  std::vector<std::string> err;
};

extern "C" int lvbm_create(const LvbConfig* cfg, int n_seq, const int* gpu_ids, int n_gpu, LvbMulti** out) {
  if (!cfg || !gpu_ids || !out || n_gpu <= 0 || n_seq < n_gpu) return lvb_set_err(LVB_E_ARG, "lvbm_create: need at least one sequence per shard");
  LvbMulti* m = new LvbMulti();
  m->S = n_seq; m->W = cfg->width; m->H = cfg->height;
  m->first.resize(n_gpu + 1);
  for (int k = 0; k <= n_gpu; ++k) m->first[k] = (int)(((long long)n_seq * k) / n_gpu);
  m->h.assign(n_gpu, nullptr); m->rc.assign(n_gpu, LVB_OK); m->err.assign(n_gpu, std::string());
  for (int k = 0; k < n_gpu; ++k) {
    const int rc = lvb_create(cfg, m->first[k + 1] - m->first[k], gpu_ids[k], &m->h[k]);
    if (rc != LVB_OK) {
      for (int q = 0; q < k; ++q) lvb_destroy(m->h[q]);
      delete m;
      return rc;                    // lvb_last_error() holds lvb_create's message
    }
  }
  *out = m;
  return LVB_OK;
}

extern "C" void lvbm_destroy(LvbMulti* m) {
  if (!m) return;
  for (LvbHandle* h : m->h) lvb_destroy(h);
  delete m;
}

extern "C" int lvbm_n_shards(const LvbMulti* m) { return m ? (int)m->h.size() : 0; }

// shard and local index of a global sequence number
static bool locate(const LvbMulti* m, int seq, int* shard, int* local) {
  if (!m || seq < 0 || seq >= m->S) return false;
  int k = 0;
  while (seq >= m->first[k + 1]) ++k;
  *shard = k; *local = seq - m->first[k];
  return true;
}

extern "C" int lvbm_set_initial_state(LvbMulti* m, int seq, double t, const double* q_xyzw, const double* p, const double* v,
                                      const double* bg, const double* ba) {
  int k, l;
  if (!locate(m, seq, &k, &l)) return lvb_set_err(LVB_E_ARG, "lvbm_set_initial_state: sequence %d", seq);
  return lvb_set_initial_state(m->h[k], l, t, q_xyzw, p, v, bg, ba);
}

// lvb_step for all S sequences: every shard runs on its own host thread (the calls block on their own stream only).
// Arrays are the [S]-sized arrays of lvb_step; images are HOST memory (each shard copies its block to its device).
extern "C" int lvbm_step(LvbMulti* m, const uint8_t* images, const double* t_img, LvbImu* imu, int* n_imu, int imu_stride,
                         uint8_t* published) {
  if (!m || !images || !t_img || !imu || !n_imu) return lvb_set_err(LVB_E_ARG, "lvbm_step: null argument");
  const size_t npx = (size_t)m->W * m->H;
  const int n = (int)m->h.size();
  auto work = [&](int k) {
    const int lo = m->first[k];
    m->rc[k] = lvb_step(m->h[k], images + npx * lo, 0, t_img + lo, imu + (size_t)lo * imu_stride, n_imu + lo, imu_stride,
                        published ? published + lo : nullptr);
    if (m->rc[k] != LVB_OK) m->err[k] = lvb_last_error();      // the error string is thread-local: carry it over
  };
  std::vector<std::thread> th;
  for (int k = 1; k < n; ++k) th.emplace_back(work, k);
  work(0);
  for (auto& t : th) t.join();
  for (int k = 0; k < n; ++k)
    if (m->rc[k] != LVB_OK) return lvb_set_err(m->rc[k], "shard %d (sequences %d..%d): %s", k, m->first[k], m->first[k + 1] - 1, m->err[k].c_str());
  return LVB_OK;
}

// out[S][17] = t, q(4), p(3), v(3), bg(3), ba(3): the trajectory gather
extern "C" int lvbm_get_states(LvbMulti* m, double* out) {
  if (!m || !out) return lvb_set_err(LVB_E_ARG, "lvbm_get_states: null argument");
  for (size_t k = 0; k < m->h.size(); ++k) {
    const int rc = lvb_get_states(m->h[k], out + (size_t)m->first[k] * 17);
    if (rc != LVB_OK) return rc;
  }
  return LVB_OK;
}

extern "C" long long lvbm_launch_count(const LvbMulti* m) {
  long long n = 0;
  if (m) for (LvbHandle* h : m->h) n += lvb_launch_count(h);
  return n;
}
