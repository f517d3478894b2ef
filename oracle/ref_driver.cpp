// TEST INFRASTRUCTURE (oracle/): drives the REFERENCE's own filter - /root/reference/src/larvio.cpp, StaticInitializer.cpp and
// FlexibleInitializer.cpp compiled unmodified from where they lie (Makefile target `ref`, outputs into oracle/_ref/) against
// the stand-in headers of oracle/ref_shim/ - over a recorded stream of feature messages + IMU samples, the way
// app/larvioMain.cpp:87-117 does, and dumps the filter state after every LarVio::processFeatures call.  Used by
// tests/golden/make_ref_golden.py (run in the build container, where /root/reference exists) to produce the golden vectors
// that pin oracle/backend.py and the CUDA back end to the reference's logic.  Never linked into the product.
//
// usage: larvio_ref <config.yaml> <in.bin> <out.bin>
// in.bin  (little-endian f64 stream): mode (0 = forced initial state as oracle/backend.py::set_initial_state, 1 = the
//         reference's own static initialiser), 17 init values [t q(4) p(3) v(3) bg(3) ba(3)], n_calls, then per call:
//         t_msg, n_imu, n_imu x [t w(3) a(3)], n_feat, n_feat x [id u v u_init v_init u_vel v_vel u_init_vel v_init_vel]
// out.bin (f64 stream) per call: ok, and when ok: t q(4) p(3) v(3) bg(3) ba(3) R_imu_cam0(9, row-major) t_cam0_imu(3) td
//         Tg(9) As(9) Ma(9) n_win n_slam n_nui dim P(dim*dim, row-major) window[n_win x (id q(4) p(3))] slam ids[n_slam]
//         slam positions [n_slam x 3] nuisance ids[n_nui] n_stable n_stable x [id xyz] n_active n_active x [id xyz]
//         n_imu_left
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>
#include <Eigen/Dense>
#include <boost/shared_ptr.hpp>
#define private public   // the harness needs the state server; the class itself is compiled from the unmodified sources
#include <larvio/larvio.h>
#undef private

using namespace larvio;

static std::vector<double> read_all(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::fprintf(stderr, "larvio_ref: cannot open %s\n", path); std::exit(2); }
  std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  std::vector<double> v(n / 8);
  if (std::fread(v.data(), 8, v.size(), f) != v.size()) { std::fprintf(stderr, "larvio_ref: short read\n"); std::exit(2); }
  std::fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc != 4) { std::fprintf(stderr, "usage: %s config.yaml in.bin out.bin\n", argv[0]); return 2; }
  std::string cfg = argv[1];
  std::vector<double> in = read_all(argv[2]);
  size_t k = 0;
  auto next = [&]() { if (k >= in.size()) { std::fprintf(stderr, "larvio_ref: input exhausted\n"); std::exit(2); } return in[k++]; };

  LarVio vio(cfg);
  if (!vio.initialize()) return 3;

  int mode = (int)next();
  double init[17]; for (double& d : init) d = next();
  int n_calls = (int)next();
  std::vector<double> out;
  std::vector<ImuData> imu_buf;
  bool forced = false;

  for (int c = 0; c < n_calls; ++c) {
    MonoCameraMeasurement msg;
    msg.timeStampToSec = next();
    int n_imu = (int)next();
    for (int i = 0; i < n_imu; ++i) {
      double r[7]; for (double& d : r) d = next();
      imu_buf.push_back(ImuData(r[0], r[1], r[2], r[3], r[4], r[5], r[6]));
    }
    int n_feat = (int)next();
    msg.features.resize(n_feat);
    for (int i = 0; i < n_feat; ++i) {
      MonoFeatureMeasurement& m = msg.features[i];
      m.id = (unsigned long long)next();
      m.u = next(); m.v = next(); m.u_init = next(); m.v_init = next();
      m.u_vel = next(); m.v_vel = next(); m.u_init_vel = next(); m.v_init_vel = next();
    }
    if (mode == 0 && !forced) {
      // what FlexibleInitializer::tryIncInit + larvio.cpp:376-386 leave behind, with the state handed in (the same
      // assignment oracle/backend.py::set_initial_state and lvb_set_initial_state make)
      IMUState& s = vio.state_server.imu_state;
      s.time = init[0];
      s.orientation = Eigen::Vector4d(init[1], init[2], init[3], init[4]);
      s.position = Eigen::Vector3d(init[5], init[6], init[7]);
      s.velocity = Eigen::Vector3d(init[8], init[9], init[10]);
      s.gyro_bias = Eigen::Vector3d(init[11], init[12], init[13]);
      s.acc_bias = Eigen::Vector3d(init[14], init[15], init[16]);
      vio.is_gravity_set = true;
      vio.take_off_stamp = s.time;
      vio.last_ZUPT_time = s.time;
      vio.last_update_time = s.time;
      vio.state_server.imu_state_FEJ_now = s;
      // previous sample of the trapezoidal terms = the first sample that will be integrated
      for (const ImuData& d : imu_buf)
        if (d.timeStampToSec > s.time) { vio.m_gyro_old = d.angular_velocity; vio.m_acc_old = d.linear_acceleration; break; }
      forced = true;
    }
    bool ok;
    const char* stg = std::getenv("LVB_REF_STAGES");   // debugging aid: run ONE call stage by stage (the body of larvio.cpp:391-417) and dump P after each
    if (stg && std::atoi(stg) == c && vio.is_gravity_set && vio.bFirstFeatures) {
      auto dump = [&](const char* what) {
        const Eigen::MatrixXd& P = vio.state_server.state_cov;
        std::string path = std::string(std::getenv("LVB_REF_STAGES_DIR") ? std::getenv("LVB_REF_STAGES_DIR") : "/tmp") + "/stage_" + what + ".bin";
        FILE* f = std::fopen(path.c_str(), "wb"); double d = (double)P.rows(); std::fwrite(&d, 8, 1, f);
        for (int i = 0; i < P.rows(); ++i) for (int j = 0; j < P.cols(); ++j) { double v = P(i, j); std::fwrite(&v, 8, 1, f); }
        std::fclose(f);
      };
      dump("0_start");
      vio.batchImuProcessing(msg.timeStampToSec + vio.state_server.td, imu_buf); dump("1_propagated");
      vio.addFeatureObservations(&msg);
      vio.stateAugmentation(); dump("2_augmented");
      if (vio.if_ZUPT_valid) vio.if_ZUPT = vio.checkZUPT();
      vio.removeLostFeatures(); dump("3_updated");
      vio.pruneImuStateBuffer(); dump("4_pruned");
      if (vio.if_FEJ_config && !vio.if_FEJ && vio.state_server.imu_state.time - vio.take_off_stamp >= 0) vio.if_FEJ = true;
      for (auto fid : vio.state_server.feature_states) vio.active_slam_features[fid] = vio.map_server[fid];
      ok = true;
    } else {
      ok = vio.processFeatures(&msg, imu_buf);
    }
    if (const char* tr = std::getenv("LVB_REF_TRACE_FEATURE")) {   // debugging aid: one feature's bookkeeping after every call
      long long fid = std::atoll(tr);
      auto it = vio.map_server.find(fid);
      if (it == vio.map_server.end()) std::fprintf(stderr, "call %d: feature %lld not in map_server\n", c, fid);
      else {
        const Feature& f = it->second;
        std::fprintf(stderr, "call %d: feature %lld obs %zu init %d in_state %d ekf %d anchor %lld invDepth %.17g pos %.17g %.17g %.17g\n", c, fid,
                     f.observations.size(), (int)f.is_initialized, (int)f.in_state, (int)f.ekf_feature, (long long)f.id_anchor, f.invDepth,
                     f.position(0), f.position(1), f.position(2));
      }
    }
    out.push_back(ok ? 1.0 : 0.0);
    if (!ok) continue;
    const IMUState& s = vio.state_server.imu_state;
    out.push_back(s.time);
    for (int i = 0; i < 4; ++i) out.push_back(s.orientation(i));
    for (int i = 0; i < 3; ++i) out.push_back(s.position(i));
    for (int i = 0; i < 3; ++i) out.push_back(s.velocity(i));
    for (int i = 0; i < 3; ++i) out.push_back(s.gyro_bias(i));
    for (int i = 0; i < 3; ++i) out.push_back(s.acc_bias(i));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out.push_back(s.R_imu_cam0(i, j));
    for (int i = 0; i < 3; ++i) out.push_back(s.t_cam0_imu(i));
    out.push_back(vio.state_server.td);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out.push_back(vio.state_server.Tg(i, j));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out.push_back(vio.state_server.As(i, j));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out.push_back(vio.state_server.Ma(i, j));
    const Eigen::MatrixXd& P = vio.state_server.state_cov;
    out.push_back((double)vio.state_server.imu_states_augment.size());
    out.push_back((double)vio.state_server.feature_states.size());
    out.push_back((double)vio.state_server.nui_ids.size());
    out.push_back((double)P.rows());
    for (int i = 0; i < P.rows(); ++i) for (int j = 0; j < P.cols(); ++j) out.push_back(P(i, j));
    for (const auto& kv : vio.state_server.imu_states_augment) {
      out.push_back((double)kv.first);
      for (int i = 0; i < 4; ++i) out.push_back(kv.second.orientation(i));
      for (int i = 0; i < 3; ++i) out.push_back(kv.second.position(i));
    }
    for (auto fid : vio.state_server.feature_states) out.push_back((double)fid);
    for (auto fid : vio.state_server.feature_states) {
      const Eigen::Vector3d& p = vio.map_server[fid].position;
      for (int i = 0; i < 3; ++i) out.push_back(p(i));
    }
    for (auto id : vio.state_server.nui_ids) out.push_back((double)id);
    std::map<FeatureIDType, Eigen::Vector3d> stable, active;
    vio.getStableMapPointPositions(stable);
    vio.getActiveeMapPointPositions(active);
    out.push_back((double)stable.size());
    for (const auto& kv : stable) { out.push_back((double)kv.first); for (int i = 0; i < 3; ++i) out.push_back(kv.second(i)); }
    out.push_back((double)active.size());
    for (const auto& kv : active) { out.push_back((double)kv.first); for (int i = 0; i < 3; ++i) out.push_back(kv.second(i)); }
    out.push_back((double)imu_buf.size());
  }
  FILE* f = std::fopen(argv[3], "wb");
  if (!f) { std::fprintf(stderr, "larvio_ref: cannot write %s\n", argv[3]); return 2; }
  std::fwrite(out.data(), 8, out.size(), f);
  std::fclose(f);
  return 0;
}
