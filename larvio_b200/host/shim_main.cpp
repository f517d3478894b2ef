// app/larvioMain.cpp:87-117 written against the drop-in façade (larvio_shim.hpp): the reference's two classes, its
// loop, its ownership rules (caller-owned IMU vector that processFeatures erases from, raw MonoCameraMeasurement pointer),
// one EuRoC ASL sequence from disk.  Exists so that the façade is LINKED and RUN (tests/test_gpu.py compares its output
// with the CPU oracle), and as the shortest example of switching a LARVIO application over.
//
//   larvio_shim_demo <config.yaml> <mav0_dir> [max_frames]
//
// stdout, one line per published odometry:   ODO t qx qy qz qw px py pz vx vy vz
// followed, when map points are pending:     PTS <S|A> n  id x y z  id x y z ...   (S stable, A active; read every 10th publish)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "larvio_shim.hpp"

extern "C" {
const char* lvbio_last_error(void);
int lvbio_png_read_gray8(const char* path, uint8_t* out, int cap_bytes, int* w, int* h);
int lvbio_euroc_read_imu(const char* csv_path, LvbImu* out, int cap, int* n);
int lvbio_euroc_read_image_list(const char* csv_path, double* t, char* names, int name_len, int cap, int* n);
int lvbio_first_align(const double* t_img, int n_img, const LvbImu* imu, int n_imu, int* img0, int* imu0);
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s <config.yaml> <mav0_dir> [max_frames]\n", argv[0]); return 2; }
  std::string cfg = argv[1];
  const std::string dir = argv[2];
  const long max_frames = argc > 3 ? std::atol(argv[3]) : -1;
  constexpr int NAME_LEN = 64;
  int n = 0;
  const std::string icsv = dir + "/cam0/data.csv", mcsv = dir + "/imu0/data.csv";
  if (lvbio_euroc_read_image_list(icsv.c_str(), nullptr, nullptr, NAME_LEN, 0, &n) != 0 || n == 0) { std::fprintf(stderr, "%s: %s\n", icsv.c_str(), lvbio_last_error()); return 1; }
  std::vector<double> t_img(n); std::vector<char> names((size_t)n * NAME_LEN);
  lvbio_euroc_read_image_list(icsv.c_str(), t_img.data(), names.data(), NAME_LEN, n, &n);
  int m = 0;
  if (lvbio_euroc_read_imu(mcsv.c_str(), nullptr, 0, &m) != 0 || m == 0) { std::fprintf(stderr, "%s: %s\n", mcsv.c_str(), lvbio_last_error()); return 1; }
  std::vector<LvbImu> imu_all(m);
  lvbio_euroc_read_imu(mcsv.c_str(), imu_all.data(), m, &m);
  int i0 = 0, m0 = 0;                                                // findFirstAlign (larvioMain.cpp:74-85)
  if (lvbio_first_align(t_img.data(), n, imu_all.data(), m, &i0, &m0) != 0) { std::fprintf(stderr, "%s: %s\n", dir.c_str(), lvbio_last_error()); return 1; }
  t_img.erase(t_img.begin(), t_img.begin() + i0);
  names.erase(names.begin(), names.begin() + (size_t)i0 * NAME_LEN);
  imu_all.erase(imu_all.begin(), imu_all.begin() + m0);

  // one batch slot shared by both objects, like the reference shares one config file and one imu buffer between them
  auto session = std::make_shared<larvio::Session>(cfg);
  larvio::ImageProcessor ip(cfg, session);
  larvio::LarVio est(cfg, session);
  if (!ip.initialize() || !est.initialize()) { std::fprintf(stderr, "initialize: %s\n", lvb_last_error()); return 1; }
  LvbConfig c;
  lvb_parse_config(cfg.c_str(), &c);

  std::vector<larvio::ImuData> imu_msg_buffer;                      // larvioMain.cpp:88
  std::vector<uint8_t> pixels((size_t)c.width * c.height);
  size_t k = 0; long pubs = 0;
  const size_t nf = (max_frames >= 0 && (size_t)max_frames < t_img.size()) ? (size_t)max_frames : t_img.size();
  for (size_t j = 0; j < nf; ++j) {
    int w = 0, h = 0;
    const std::string path = dir + "/cam0/data/" + std::string(&names[j * NAME_LEN]);
    if (lvbio_png_read_gray8(path.c_str(), pixels.data(), (int)pixels.size(), &w, &h) != 0) { std::fprintf(stderr, "%s: %s\n", path.c_str(), lvbio_last_error()); return 1; }
    larvio::ImgData img{t_img[j], pixels.data(), w, h, w};
    while (k < imu_all.size() && imu_all[k].t - t_img[j] < 0.05) {   // larvioMain.cpp:98-102
      larvio::ImuData d;
      d.timeStampToSec = imu_all[k].t;
      for (int a = 0; a < 3; ++a) { d.angular_velocity[a] = imu_all[k].gyro[a]; d.linear_acceleration[a] = imu_all[k].acc[a]; }
      imu_msg_buffer.push_back(d);
      ++k;
    }
    larvio::MonoCameraMeasurementPtr features = new larvio::MonoCameraMeasurement;   // :105
    const bool bProcess = ip.processImage(&img, imu_msg_buffer, features);           // :107
    bool bPubOdo = false;
    if (bProcess) bPubOdo = est.processFeatures(features, imu_msg_buffer);          // :114
    delete features;
    if (!bPubOdo) continue;
    const larvio::Pose T = est.getTbw();
    double v[3];
    est.getVel(v);
    std::printf("ODO %.9f %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", t_img[j], T.q_xyzw[0], T.q_xyzw[1], T.q_xyzw[2], T.q_xyzw[3],
                T.p[0], T.p[1], T.p[2], v[0], v[1], v[2]);
    if (++pubs % 10 == 0) {
      for (int which = 0; which < 2; ++which) {
        std::map<unsigned long long, larvio::LarVio::Point3> pts;
        if (which == 0) est.getStableMapPointPositions(pts); else est.getActiveeMapPointPositions(pts);
        if (pts.empty()) continue;
        std::printf("PTS %c %zu", which == 0 ? 'S' : 'A', pts.size());
        for (const auto& kv : pts) std::printf(" %llu %.17g %.17g %.17g", kv.first, kv.second.x, kv.second.y, kv.second.z);
        std::printf("\n");
      }
    }
  }
  std::fprintf(stderr, "shim demo: %zu frames, %ld odometry messages\n", nf, pubs);
  return 0;
}
