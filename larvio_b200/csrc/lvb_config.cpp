// Reader for the reference's config dialect (OpenCV FileStorage "%YAML:1.0") without OpenCV.
// Keys and their meaning: front end image_processor.cpp:51-92, back end larvio.cpp:65-277.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/larvio_b200.h"

int lvb_set_err(int code, const char* fmt, ...);

namespace {

std::string strip_comment(const std::string& raw) {
  bool inq = false;
  for (size_t j = 0; j < raw.size(); ++j) {
    if (raw[j] == '"') inq = !inq;
    else if (raw[j] == '#' && !inq) return raw.substr(0, j);
  }
  return raw;
}
std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  return s.substr(a, b - a + 1);
}

struct Parsed {
  std::map<std::string, std::string> scalars;        // "key" or "map.key"
  std::map<std::string, std::vector<double>> mats;
};

bool parse_file(const char* path, Parsed& P) {
  FILE* f = fopen(path, "r");
  if (!f) return false;
  std::vector<std::string> lines;
  char buf[4096];
  while (fgets(buf, sizeof(buf), f)) lines.push_back(buf);
  fclose(f);
  std::string cur_map;
  for (size_t i = 0; i < lines.size();) {
    std::string line = strip_comment(lines[i++]);
    std::string t = trim(line);
    if (t.empty() || t.compare(0, 5, "%YAML") == 0 || t == "---") continue;
    size_t indent = line.find_first_not_of(" \t");
    size_t colon = t.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(t.substr(0, colon)), val = trim(t.substr(colon + 1));
    if (indent == 0) {
      cur_map.clear();
      if (val.compare(0, 15, "!!opencv-matrix") == 0) {
        std::string data;
        bool in_data = false, done = false;
        while (i < lines.size() && !done) {
          std::string l2 = strip_comment(lines[i++]);
          std::string t2 = trim(l2);
          if (!in_data) {
            if (t2.compare(0, 5, "data:") == 0) { in_data = true; t2 = t2.substr(5); }
            else continue;
          }
          data += " " + t2;
          if (t2.find(']') != std::string::npos) done = true;
        }
        std::vector<double> v;
        const char* p = data.c_str();
        while (*p) {
          if ((*p >= '0' && *p <= '9') || *p == '-' || *p == '+' || *p == '.') {
            char* e = nullptr;
            double d = strtod(p, &e);
            if (e == p) { ++p; continue; }
            v.push_back(d);
            p = e;
          } else ++p;
        }
        P.mats[key] = v;
      } else if (val.empty()) {
        cur_map = key;
      } else {
        if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
        P.scalars[key] = val;
      }
    } else if (!cur_map.empty()) {
      P.scalars[cur_map + "." + key] = val;
    }
  }
  return true;
}

bool getd(const Parsed& P, const char* k, double* out) {
  auto it = P.scalars.find(k);
  if (it == P.scalars.end()) return false;
  *out = atof(it->second.c_str());
  return true;
}
}  // namespace

extern "C" int lvb_parse_config(const char* yaml_path, LvbConfig* c) {
  if (!yaml_path || !c) return lvb_set_err(LVB_E_ARG, "lvb_parse_config: null argument");
  Parsed P;
  if (!parse_file(yaml_path, P))   // image_processor.cpp:46-49 / larvio.cpp:60-63: unreadable config
    return lvb_set_err(LVB_E_CONFIG, "config_file error: cannot open %s", yaml_path);
  memset(c, 0, sizeof(*c));
  double d;
#define REQ_D(field, key) if (!getd(P, key, &d)) return lvb_set_err(LVB_E_CONFIG, "missing key %s", key); c->field = d;
#define REQ_I(field, key) if (!getd(P, key, &d)) return lvb_set_err(LVB_E_CONFIG, "missing key %s", key); c->field = (int)d;
  REQ_I(width, "resolution_width") REQ_I(height, "resolution_height")
  auto dm = P.scalars.find("distortion_model");
  c->distortion_model = (dm != P.scalars.end() && dm->second == "equidistant") ? 1 : 0;
  REQ_D(fx, "intrinsics.fx") REQ_D(fy, "intrinsics.fy") REQ_D(cx, "intrinsics.cx") REQ_D(cy, "intrinsics.cy")
  REQ_D(dist[0], "distortion_coeffs.k1") REQ_D(dist[1], "distortion_coeffs.k2")
  REQ_D(dist[2], "distortion_coeffs.p1") REQ_D(dist[3], "distortion_coeffs.p2")
  auto m = P.mats.find("T_cam_imu");
  if (m == P.mats.end() || m->second.size() != 16) return lvb_set_err(LVB_E_CONFIG, "T_cam_imu must be a 4x4 opencv-matrix");
  for (int i = 0; i < 16; ++i) c->T_cam_imu[i] = m->second[i];
  REQ_I(pyramid_levels, "pyramid_levels") REQ_I(patch_size, "patch_size") REQ_I(max_iteration, "max_iteration")
  REQ_I(max_features_num, "max_features_num") REQ_I(min_distance, "min_distance") REQ_I(flag_equalize, "flag_equalize")
  REQ_D(track_precision, "track_precision") REQ_D(pub_frequency, "pub_frequency") REQ_D(img_rate, "img_rate")
  REQ_D(imu_rate, "imu_rate") REQ_D(rotation_threshold, "rotation_threshold")
  REQ_D(translation_threshold, "translation_threshold") REQ_D(tracking_rate_threshold, "tracking_rate_threshold")
  REQ_D(feature_translation_threshold, "feature_translation_threshold") REQ_D(td, "td")
  REQ_D(noise_gyro, "noise_gyro") REQ_D(noise_acc, "noise_acc") REQ_D(noise_gyro_bias, "noise_gyro_bias")
  REQ_D(noise_acc_bias, "noise_acc_bias") REQ_D(noise_feature, "noise_feature")
  REQ_D(cov_orientation, "initial_covariance_orientation") REQ_D(cov_velocity, "initial_covariance_velocity")
  REQ_D(cov_position, "initial_covariance_position") REQ_D(cov_gyro_bias, "initial_covariance_gyro_bias")
  REQ_D(cov_acc_bias, "initial_covariance_acc_bias") REQ_D(cov_extrin_rot, "initial_covariance_extrin_rot")
  REQ_D(cov_extrin_trans, "initial_covariance_extrin_trans")
  REQ_D(zupt_max_feature_dis, "zupt_max_feature_dis") REQ_D(zupt_noise_v, "zupt_noise_v")
  REQ_D(zupt_noise_p, "zupt_noise_p") REQ_D(zupt_noise_q, "zupt_noise_q") REQ_D(static_duration, "static_duration")
  REQ_I(max_track_len, "max_track_len") REQ_I(sw_size, "sw_size") REQ_I(least_observation_number, "least_observation_number")
  REQ_I(if_FEJ, "if_FEJ") REQ_I(estimate_extrin, "estimate_extrin") REQ_I(estimate_td, "estimate_td")
  REQ_I(calib_imu_instrinsic, "calib_imu_instrinsic") REQ_I(if_ZUPT_valid, "if_ZUPT_valid")
  REQ_I(max_features_in_one_grid, "max_features_in_one_grid") REQ_I(aug_grid_rows, "aug_grid_rows")
  REQ_I(aug_grid_cols, "aug_grid_cols") REQ_I(feature_idp_dim, "feature_idp_dim") REQ_I(use_schmidt, "use_schmidt")
#undef REQ_D
#undef REQ_I
  return LVB_OK;
}
