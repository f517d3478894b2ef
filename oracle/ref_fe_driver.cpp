// TEST INFRASTRUCTURE (oracle/): drives the REFERENCE's own front end - /root/reference/src/image_processor.cpp and
// src/ORBDescriptor.cpp compiled unmodified (Makefile target `ref_fe`, output oracle/_ref/larvio_ref_fe) against the stand-in
// headers of oracle/ref_shim/, whose OpenCV functions are executed by the cv2 module through oracle/cv_server.py - over a
// recorded image + IMU stream, the way app/larvioMain.cpp:87-117 does, and dumps every MonoCameraMeasurement it publishes.
// tests/golden/make_ref_fe_golden.py turns the dump into the fixture that pins oracle/frontend.py and the CUDA front end to the
// reference's bookkeeping (feature ids, their order, lifetimes, the FIRST/SECOND/OTHER image state machine).
//
// usage: larvio_ref_fe <config.yaml> <in.bin> <out.bin>       (environment: LVB_CV_SERVER = path of oracle/cv_server.py)
// in.bin : f64 header [n_frames, height, width, n_imu], f64 img_t[n_frames], f64 imu[n_imu][7] (t w a), u8 images[n_frames][h][w]
// out.bin: f64 stream, per frame: has_msg, and when set: t, n, n x [id u v u_init v_init u_vel v_vel u_init_vel v_init_vel]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <larvio/image_processor.h>

using namespace larvio;

int main(int argc, char** argv) {
  if (argc != 4) { std::fprintf(stderr, "usage: %s config.yaml in.bin out.bin\n", argv[0]); return 2; }
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
  if (!ip.initialize()) return 3;
  std::vector<ImuData> imu_buf;
  std::vector<double> out;
  int k = 0;
  for (int j = 0; j < nf; ++j) {
    while (k < ni && imu[(size_t)k * 7] - img_t[j] < 0.05) {            // app/larvioMain.cpp:98
      const double* r = &imu[(size_t)k * 7];
      imu_buf.push_back(ImuData(r[0], r[1], r[2], r[3], r[4], r[5], r[6])); ++k;
    }
    ImageDataPtr msg(new ImgData);
    msg->timeStampToSec = img_t[j];
    msg->image = cv::Mat(H, W, CV_8UC1);
    std::memcpy(msg->image.data, &pix[(size_t)j * H * W], (size_t)H * W);
    MonoCameraMeasurement feat;
    const bool has = ip.processImage(msg, imu_buf, &feat);
    out.push_back(has ? 1.0 : 0.0);
    if (!has) continue;
    out.push_back(feat.timeStampToSec); out.push_back((double)feat.features.size());
    for (const auto& m : feat.features) {
      out.push_back((double)m.id); out.push_back(m.u); out.push_back(m.v); out.push_back(m.u_init); out.push_back(m.v_init);
      out.push_back(m.u_vel); out.push_back(m.v_vel); out.push_back(m.u_init_vel); out.push_back(m.v_init_vel);
    }
  }
  f = std::fopen(argv[3], "wb");
  if (!f) return 2;
  std::fwrite(out.data(), 8, out.size(), f);
  std::fclose(f);
  return 0;
}
