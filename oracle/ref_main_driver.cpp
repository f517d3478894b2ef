// TEST INFRASTRUCTURE (oracle/): the reference's WHOLE per-frame pipeline - src/image_processor.cpp, src/ORBDescriptor.cpp,
// src/larvio.cpp, src/StaticInitializer.cpp, src/FlexibleInitializer.cpp, all compiled unmodified (Makefile target `ref_main`,
// output oracle/_ref/larvio_ref_main) against the stand-in headers of oracle/ref_shim/ - driven by the loop of
// app/larvioMain.cpp:87-117 (one IMU buffer shared by processImage and processFeatures, which erases what it consumed), through
// the PUBLIC interface only.  Prints the odometry it would visualise: per published frame
//   ODO t R(9, row-major) p(3) v(3)        <- getTbw(), getVel()
// and after every 10th publication the two map-point getters (larvio.h:86-87)
//   PTS S|A n  id x y z ...
// tests/golden/make_ref_main_golden.py stores these lines; the drop-in facade (larvio_b200/bin/larvio_shim_demo, the same loop on
// the shim classes over the CUDA library) is compared with them on the GPU box.
//
// usage: larvio_ref_main <config.yaml> <in.bin>        (in.bin as oracle/ref_fe_driver.cpp; LVB_CV_SERVER as there)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <larvio/image_processor.h>
#include <larvio/larvio.h>

using namespace larvio;

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s config.yaml in.bin\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[2], "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
  double hdr[4];
  if (std::fread(hdr, 8, 4, f) != 4) return 2;
  const int nf = (int)hdr[0], H = (int)hdr[1], W = (int)hdr[2], ni = (int)hdr[3];
  std::vector<double> img_t(nf), imu((size_t)ni * 7);
  if (std::fread(img_t.data(), 8, nf, f) != (size_t)nf || std::fread(imu.data(), 8, imu.size(), f) != imu.size()) return 2;
  std::vector<unsigned char> pix((size_t)nf * H * W);
  if (std::fread(pix.data(), 1, pix.size(), f) != pix.size()) return 2;
  std::fclose(f);

  std::string cfg = argv[1];
  ImageProcessor ip(cfg);
  LarVio est(cfg);
  if (!ip.initialize() || !est.initialize()) return 3;
  std::vector<ImuData> imu_msg_buffer;                                   // larvioMain.cpp:88
  size_t k = 0; long pubs = 0;
  FILE* out = stdout;
  for (int j = 0; j < nf; ++j) {
    ImageDataPtr img(new ImgData);
    img->timeStampToSec = img_t[j];
    img->image = cv::Mat(H, W, CV_8UC1);
    std::memcpy(img->image.data, &pix[(size_t)j * H * W], (size_t)H * W);
    while (k < (size_t)ni && imu[k * 7] - img_t[j] < 0.05) {              // :98-102
      const double* r = &imu[k * 7];
      imu_msg_buffer.push_back(ImuData(r[0], r[1], r[2], r[3], r[4], r[5], r[6])); ++k;
    }
    MonoCameraMeasurementPtr features = new MonoCameraMeasurement;       // :105
    const bool bProcess = ip.processImage(img, imu_msg_buffer, features);   // :107
    bool bPubOdo = false;
    const size_t n_feat = features->features.size();
    if (bProcess) bPubOdo = est.processFeatures(features, imu_msg_buffer);  // :114
    if (std::getenv("LVB_REF_TRACE"))      // debugging aid: what each frame did
      std::fprintf(stderr, "FRAME %d t %.4f processImage %d features %zu processFeatures %d imu left %zu\n", j, img_t[j], (int)bProcess, n_feat, (int)bPubOdo, imu_msg_buffer.size());
    delete features;
    if (!bPubOdo) continue;
    const Eigen::Isometry3d T = est.getTbw();
    const Eigen::Vector3d v = est.getVel();
    std::fprintf(out, "\nODO %.9f", img_t[j]);      // leading newline: the reference prints diagnostics without one
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) std::fprintf(out, " %.17g", T.linear()(a, b));
    for (int a = 0; a < 3; ++a) std::fprintf(out, " %.17g", T.translation()(a));
    for (int a = 0; a < 3; ++a) std::fprintf(out, " %.17g", v(a));
    std::fprintf(out, "\n");
    if (++pubs % 10 == 0) {
      for (int which = 0; which < 2; ++which) {
        std::map<FeatureIDType, Eigen::Vector3d> pts;
        if (which == 0) est.getStableMapPointPositions(pts); else est.getActiveeMapPointPositions(pts);
        if (pts.empty()) continue;
        std::fprintf(out, "\nPTS %c %zu", which == 0 ? 'S' : 'A', pts.size());
        for (const auto& kv : pts) std::fprintf(out, " %lld %.17g %.17g %.17g", (long long)kv.first, kv.second(0), kv.second(1), kv.second(2));
        std::fprintf(out, "\n");
      }
    }
  }
  std::fflush(out);
  return 0;
}
