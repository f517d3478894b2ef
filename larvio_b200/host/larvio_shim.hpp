// Drop-in C++ façade: the reference's two classes for ONE sequence on top of the C ABI.
//
//   larvio::ImageProcessor  <->  include/larvio/image_processor.h:39-65
//   larvio::LarVio          <->  include/larvio/larvio.h:42-87
//
// POD restatements replace the third-party types of the reference signatures (none of those libraries
// is needed): cv::Mat -> ImgData{w,h,stride,ptr}; Eigen::Isometry3d -> Pose{q_xyzw,p};
// boost::shared_ptr<ImgData> -> const ImgData*.  Both objects of one sequence share one LvbHandle
// (the front end and the filter of a sequence live in the same batch slot), exactly like the driver
// shares one imu buffer between them (app/larvioMain.cpp:107,114).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/larvio_b200.h"

namespace larvio {

struct ImuData {                    // include/sensors/ImuData.hpp:16-38
  double timeStampToSec;
  double angular_velocity[3];
  double linear_acceleration[3];
};
struct ImgData {                    // include/sensors/ImageData.hpp:16-19 (cv::Mat -> raw view)
  double timeStampToSec;
  const uint8_t* data; int width, height, stride;
};
typedef const ImgData* ImageDataPtr;
struct MonoFeatureMeasurement {     // include/larvio/feature_msg.h:15-47
  unsigned long long id;
  double u, v, u_init, v_init, u_vel, v_vel, u_init_vel, v_init_vel;
};
struct MonoCameraMeasurement { double timeStampToSec; std::vector<MonoFeatureMeasurement> features; };
typedef MonoCameraMeasurement* MonoCameraMeasurementPtr;    // raw pointer, like feature_msg.h:56
struct Pose { double q_xyzw[4]; double p[3]; };

class Session {                     // one batch slot shared by the two façades
 public:
  explicit Session(const std::string& cfg, int device = 0) : cfg_(cfg), device_(device), h_(nullptr) {}
  ~Session() { if (h_) lvb_destroy(h_); }
  bool open() { return h_ || lvb_create_from_file(cfg_.c_str(), 1, device_, &h_) == LVB_OK; }
  LvbHandle* h() { return h_; }
 private:
  std::string cfg_; int device_; LvbHandle* h_;
};

class ImageProcessor {
 public:
  ImageProcessor(std::string& config_file, std::shared_ptr<Session> s = nullptr)
      : s_(s ? s : std::make_shared<Session>(config_file)) {}
  bool initialize() { return s_->open(); }                       // false on unreadable config, like :116-126
  // processImage(msg, imu_buffer, features) -> "a feature message was emitted"
  bool processImage(const ImageDataPtr& msg, const std::vector<ImuData>& imu, MonoCameraMeasurementPtr features) {
    std::vector<uint8_t> img((size_t)msg->width * msg->height);
    for (int y = 0; y < msg->height; ++y) std::memcpy(&img[(size_t)y * msg->width], msg->data + (size_t)y * msg->stride, msg->width);
    std::vector<LvbImu> b(imu.size() ? imu.size() : 1);
    for (size_t i = 0; i < imu.size(); ++i) {
      b[i].t = imu[i].timeStampToSec;
      for (int k = 0; k < 3; ++k) { b[i].gyro[k] = imu[i].angular_velocity[k]; b[i].acc[k] = imu[i].linear_acceleration[k]; }
    }
    const int cap = lvb_feature_capacity(s_->h());
    std::vector<LvbFeature> out(cap);
    int n_imu = (int)imu.size(), n_out = 0; uint8_t has = 0; double t = msg->timeStampToSec;
    if (lvb_process_images(s_->h(), img.data(), &t, b.data(), &n_imu, (int)b.size(), out.data(), &n_out, &has) != LVB_OK) return false;
    if (!has) return false;
    features->timeStampToSec = t;
    features->features.resize(n_out);
    for (int i = 0; i < n_out; ++i) std::memcpy(&features->features[i], &out[i], sizeof(LvbFeature));
    return true;
  }
  std::shared_ptr<Session> session() { return s_; }
 private:
  std::shared_ptr<Session> s_;
};

class LarVio {
 public:
  LarVio(std::string& config_file, std::shared_ptr<Session> s = nullptr)
      : s_(s ? s : std::make_shared<Session>(config_file)), cfg_(config_file) {}
  ~LarVio() { if (init_) lvb_static_init_destroy(init_); }
  bool initialize() {
    if (!s_->open()) return false;
    LvbConfig c;
    if (lvb_parse_config(cfg_.c_str(), &c) != LVB_OK) return false;
    init_ = lvb_static_init_create(&c);
    td0_ = c.td;
    return init_ != nullptr;
  }
  // what FlexibleInitializer::tryIncInit leaves behind (larvio.cpp:376-386), for callers that start the filter themselves
  bool setInitialState(double t, const Pose& T_b_w, const double v[3], const double bg[3], const double ba[3]) {
    gravity_set_ = lvb_set_initial_state(s_->h(), 0, t, T_b_w.q_xyzw, T_b_w.p, v, bg, ba) == LVB_OK;
    return gravity_set_;
  }
  // processFeatures(msg, imu_buffer): consumed samples are erased from the caller's vector (larvio.cpp:510-512)
  bool processFeatures(MonoCameraMeasurementPtr msg, std::vector<ImuData>& imu) {
    std::vector<LvbImu> b(imu.size() ? imu.size() : 1);
    for (size_t i = 0; i < imu.size(); ++i) {
      b[i].t = imu[i].timeStampToSec;
      for (int k = 0; k < 3; ++k) { b[i].gyro[k] = imu[i].angular_velocity[k]; b[i].acc[k] = imu[i].linear_acceleration[k]; }
    }
    std::vector<LvbFeature> f(msg->features.size() ? msg->features.size() : 1);
    for (size_t i = 0; i < msg->features.size(); ++i) std::memcpy(&f[i], &msg->features[i], sizeof(LvbFeature));
    int n_imu = (int)imu.size(), n_feat = (int)msg->features.size(); uint8_t valid = 1, ok = 0; double t = msg->timeStampToSec;
    if (!gravity_set_) {                                            // larvio.cpp:375-391 (static initialiser only)
      if (!first_features_) {                                       // :365-372: no initialiser before an IMU sample precedes a message
        if (n_imu > 0 && b[0].t - t - td0_ <= 0.0) first_features_ = true;
        else return false;
      }
      double st[17], g0[3], a0[3]; int used = 0;
      if (lvb_static_init_try(init_, f.data(), n_feat, t, b.data(), n_imu, st, g0, a0, &used) != 1) return false;
      if (lvb_set_initial_state(s_->h(), 0, st[0], st + 1, st + 5, st + 8, st + 11, st + 14) != LVB_OK) return false;
      gravity_set_ = true;
      imu.erase(imu.begin(), imu.begin() + used);                   // StaticInitializer.cpp:149-150
      b.erase(b.begin(), b.begin() + used); n_imu -= used;
      if (b.empty()) b.resize(1);
    }
    if (lvb_process_features(s_->h(), &valid, &t, f.data(), &n_feat, (int)f.size(), b.data(), &n_imu, (int)b.size(), &ok) != LVB_OK) return false;
    imu.erase(imu.begin(), imu.begin() + ((int)imu.size() - n_imu));
    return ok != 0;
  }
  Pose getTbw() { Pose T{}; double t; lvb_get_state(s_->h(), 0, &t, T.q_xyzw, T.p, nullptr, nullptr, nullptr, nullptr, nullptr); return T; }
  void getVel(double v[3]) { double t; lvb_get_state(s_->h(), 0, &t, nullptr, nullptr, v, nullptr, nullptr, nullptr, nullptr); }
  void getPpose(double P[36]) { double t; lvb_get_state(s_->h(), 0, &t, nullptr, nullptr, nullptr, nullptr, nullptr, P, nullptr); }
  void getPvel(double P[9]) { double t; lvb_get_state(s_->h(), 0, &t, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, P); }
  // larvio.h:86-87: map<FeatureIDType, Vector3d> -> map<id, Point3>; both clear what they return, like larvio.cpp:2719-2733
  struct Point3 { double x, y, z; };
  void getStableMapPointPositions(std::map<unsigned long long, Point3>& mMapPoints) { get_points(0, mMapPoints); }
  void getActiveeMapPointPositions(std::map<unsigned long long, Point3>& mMapPoints) { get_points(1, mMapPoints); }
  void getSwPoses(std::vector<Pose>& out) {
    double qp[64 * 7]; int n = 0;
    lvb_get_window(s_->h(), 0, qp, 64, &n);
    out.resize(n);
    for (int i = 0; i < n; ++i) { std::memcpy(out[i].q_xyzw, qp + i * 7, 4 * sizeof(double)); std::memcpy(out[i].p, qp + i * 7 + 4, 3 * sizeof(double)); }
  }
 private:
  void get_points(int which, std::map<unsigned long long, Point3>& m) {
    unsigned long long ids[512]; double xyz[512 * 3]; int n = 0;
    lvb_get_points(s_->h(), 0, which, ids, xyz, 512, &n);
    for (int i = 0; i < n; ++i) m[ids[i]] = Point3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  }
  std::shared_ptr<Session> s_;
  std::string cfg_;
  LvbStaticInit* init_ = nullptr;
  bool gravity_set_ = false;
  bool first_features_ = false;
  double td0_ = 0.0;
};

}  // namespace larvio
