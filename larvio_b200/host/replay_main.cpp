// Batched replay driver: app/larvioMain.cpp:30-117 for S EuRoC ASL sequences in lock-step, host C++ over the C ABI.
//
//   larvio_replay <config.yaml> <output_dir> [--device D] [--max-frames K] <mav0_dir> [<mav0_dir> ...]
//
// Per sequence: cam0/data.csv + imu0/data.csv are read and aligned like DataReader.hpp does (loadImageList,
// loadImuFile, findFirstAlign), images are decoded on the host (8-bit grey PNG), IMU samples up to 50 ms past each
// image are appended to that sequence's buffer (larvioMain.cpp:98-102), processImage / processFeatures run for the whole
// batch through lvb_process_images / lvb_process_features, the filter self-starts from a standstill through the
// inclinometer initialiser (larvio.cpp:375-391), and every published frame appends one line to
// <output_dir>/seq<k>/msckf_2_state.txt in the reference's format (larvio.cpp:420-453).  No GUI, no ROS.
#include <sys/stat.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/larvio_b200.h"

extern "C" {
const char* lvbio_last_error(void);
int lvbio_png_read_gray8(const char* path, uint8_t* out, int cap_bytes, int* w, int* h);
int lvbio_euroc_read_imu(const char* csv_path, LvbImu* out, int cap, int* n);
int lvbio_euroc_read_image_list(const char* csv_path, double* t, char* names, int name_len, int cap, int* n);
int lvbio_first_align(const double* t_img, int n_img, const LvbImu* imu, int n_imu, int* img0, int* imu0);
}

namespace {
constexpr int NAME_LEN = 64;
struct Seq {
  std::string dir;
  std::vector<double> t_img;
  std::vector<char> names;
  std::vector<LvbImu> imu;
  size_t k = 0;                       // next IMU sample to hand over
  std::vector<LvbImu> buf;            // the caller-owned imu_msg_buffer of this sequence
  LvbStaticInit* init = nullptr;
  bool started = false;
  double take_off = 0.0;
  FILE* log = nullptr;
};

// Quaterniond(R) coefficients (w, x, y, z) of a rotation matrix (row-major)
void rot_to_quat_wxyz(const double* R, double* q) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s <config.yaml> <output_dir> [--device D] [--max-frames K] <mav0_dir> [<mav0_dir> ...]\n", argv[0]);
    return 2;
  }
  const std::string cfg_path = argv[1], out_dir = argv[2];
  int device = 0; long max_frames = -1;
  std::vector<Seq> seqs;
  for (int i = 3; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--max-frames") && i + 1 < argc) max_frames = std::atol(argv[++i]);
    else { Seq q; q.dir = argv[i]; seqs.push_back(q); }
  }
  if (seqs.empty()) { std::fprintf(stderr, "no sequence directory given\n"); return 2; }
  LvbConfig cfg;
  if (lvb_parse_config(cfg_path.c_str(), &cfg) != LVB_OK) { std::fprintf(stderr, "%s\n", lvb_last_error()); return 1; }
  const int S = (int)seqs.size();
  // ---- load and align every sequence
  size_t n_frames = (size_t)-1;
  for (int s = 0; s < S; ++s) {
    Seq& q = seqs[s];
    int n = 0;
    const std::string icsv = q.dir + "/cam0/data.csv", mcsv = q.dir + "/imu0/data.csv";
    if (lvbio_euroc_read_image_list(icsv.c_str(), nullptr, nullptr, NAME_LEN, 0, &n) != 0 || n == 0) { std::fprintf(stderr, "%s: %s\n", icsv.c_str(), lvbio_last_error()); return 1; }
    q.t_img.resize(n); q.names.resize((size_t)n * NAME_LEN);
    lvbio_euroc_read_image_list(icsv.c_str(), q.t_img.data(), q.names.data(), NAME_LEN, n, &n);
    if (lvbio_euroc_read_imu(mcsv.c_str(), nullptr, 0, &n) != 0 || n == 0) { std::fprintf(stderr, "%s: %s\n", mcsv.c_str(), lvbio_last_error()); return 1; }
    q.imu.resize(n);
    lvbio_euroc_read_imu(mcsv.c_str(), q.imu.data(), n, &n);
    int i0 = 0, m0 = 0;
    if (lvbio_first_align(q.t_img.data(), (int)q.t_img.size(), q.imu.data(), (int)q.imu.size(), &i0, &m0) != 0) { std::fprintf(stderr, "%s: %s\n", q.dir.c_str(), lvbio_last_error()); return 1; }
    q.t_img.erase(q.t_img.begin(), q.t_img.begin() + i0);
    q.names.erase(q.names.begin(), q.names.begin() + (size_t)i0 * NAME_LEN);
    q.imu.erase(q.imu.begin(), q.imu.begin() + m0);
    if (q.t_img.size() < n_frames) n_frames = q.t_img.size();
    q.init = lvb_static_init_create(&cfg);
    const std::string sd = out_dir + "/seq" + std::to_string(s);
    mkdir(out_dir.c_str(), 0755); mkdir(sd.c_str(), 0755);
    q.log = std::fopen((sd + "/msckf_2_state.txt").c_str(), "w");
    if (!q.init || !q.log) { std::fprintf(stderr, "cannot set up sequence %d (%s)\n", s, sd.c_str()); return 1; }
  }
  if (max_frames >= 0 && (size_t)max_frames < n_frames) n_frames = (size_t)max_frames;
  // ---- the batch
  LvbHandle* h = nullptr;
  if (lvb_create(&cfg, S, device, &h) != LVB_OK) { std::fprintf(stderr, "lvb_create: %s\n", lvb_last_error()); return 1; }
  const int cap = lvb_feature_capacity(h);
  const size_t npx = (size_t)cfg.width * cfg.height;
  const int IMU_STRIDE = 4096;
  std::vector<uint8_t> images(npx * S), has(S), valid(S), ok(S);
  std::vector<double> t_img(S), t_msg(S);
  std::vector<LvbImu> imu((size_t)S * IMU_STRIDE);
  std::vector<int> n_imu(S), out_n(S);
  std::vector<LvbFeature> feat((size_t)S * cap);
  long published = 0;
  for (size_t j = 0; j < n_frames; ++j) {
    for (int s = 0; s < S; ++s) {
      Seq& q = seqs[s];
      const std::string path = q.dir + "/cam0/data/" + std::string(&q.names[j * NAME_LEN]);
      int w = 0, hh = 0;
      if (lvbio_png_read_gray8(path.c_str(), &images[npx * s], (int)npx, &w, &hh) != 0 || w != cfg.width || hh != cfg.height) {
        std::fprintf(stderr, "%s: %s (expected %dx%d)\n", path.c_str(), lvbio_last_error(), cfg.width, cfg.height);
        return 1;
      }
      t_img[s] = q.t_img[j];
      while (q.k < q.imu.size() && q.imu[q.k].t - t_img[s] < 0.05) q.buf.push_back(q.imu[q.k++]);     // larvioMain.cpp:98-102
      if ((int)q.buf.size() > IMU_STRIDE) { std::fprintf(stderr, "sequence %d: IMU buffer overflow\n", s); return 1; }
      std::memcpy(&imu[(size_t)s * IMU_STRIDE], q.buf.data(), sizeof(LvbImu) * q.buf.size());
      n_imu[s] = (int)q.buf.size();
    }
    if (lvb_process_images(h, images.data(), t_img.data(), imu.data(), n_imu.data(), IMU_STRIDE, feat.data(), out_n.data(), has.data()) != LVB_OK) {
      std::fprintf(stderr, "lvb_process_images: %s\n", lvb_last_error()); return 1;
    }
    bool any = false;
    for (int s = 0; s < S; ++s) {
      Seq& q = seqs[s];
      valid[s] = 0; t_msg[s] = t_img[s];
      if (!has[s]) continue;
      if (!q.started) {                                                  // larvio.cpp:375-391
        double st[17], g0[3], a0[3]; int used = 0;
        const int rc = lvb_static_init_try(q.init, &feat[(size_t)s * cap], out_n[s], t_img[s], q.buf.data(), (int)q.buf.size(), st, g0, a0, &used);
        if (rc < 0) { std::fprintf(stderr, "lvb_static_init_try: %s\n", lvb_last_error()); return 1; }
        if (rc == 0) continue;
        if (lvb_set_initial_state(h, s, st[0], st + 1, st + 5, st + 8, st + 11, st + 14) != LVB_OK) { std::fprintf(stderr, "%s\n", lvb_last_error()); return 1; }
        q.buf.erase(q.buf.begin(), q.buf.begin() + used);
        std::memcpy(&imu[(size_t)s * IMU_STRIDE], q.buf.data(), sizeof(LvbImu) * q.buf.size());
        n_imu[s] = (int)q.buf.size();
        q.started = true; q.take_off = st[0];
        FILE* ft = std::fopen((out_dir + "/seq" + std::to_string(s) + "/msckf_2_takeoff.txt").c_str(), "w");
        if (ft) { std::fprintf(ft, "%.9f\n", q.take_off); std::fclose(ft); }
      }
      valid[s] = 1; any = true;
    }
    if (!any) continue;
    if (lvb_process_features(h, valid.data(), t_msg.data(), feat.data(), out_n.data(), cap, imu.data(), n_imu.data(), IMU_STRIDE, ok.data()) != LVB_OK) {
      std::fprintf(stderr, "lvb_process_features: %s\n", lvb_last_error()); return 1;
    }
    for (int s = 0; s < S; ++s) {
      Seq& q = seqs[s];
      if (!valid[s]) continue;
      q.buf.assign(&imu[(size_t)s * IMU_STRIDE], &imu[(size_t)s * IMU_STRIDE] + n_imu[s]);   // consumed samples were erased (larvio.cpp:510-512)
      if (!ok[s]) continue;
      double t, qx[4], p[3], v[3], bg[3], ba[3], R[9], tc[3], td;
      lvb_get_state(h, s, &t, qx, p, v, bg, ba, nullptr, nullptr);
      lvb_get_calibration(h, s, R, tc, &td, nullptr, nullptr, nullptr);
      double qbc[4];
      rot_to_quat_wxyz(R, qbc);
      std::fprintf(q.log, "%g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g\n", t - q.take_off, qx[3], qx[0], qx[1], qx[2],
                   v[0], v[1], v[2], p[0], p[1], p[2], bg[0], bg[1], bg[2], ba[0], ba[1], ba[2], qbc[0], qbc[1], qbc[2], qbc[3], tc[0], tc[1], tc[2]);
      ++published;
    }
  }
  for (Seq& q : seqs) { std::fclose(q.log); lvb_static_init_destroy(q.init); }
  lvb_destroy(h);
  std::printf("replayed %zu frames of %d sequences, %ld odometry lines\n", n_frames, S, published);
  return 0;
}
