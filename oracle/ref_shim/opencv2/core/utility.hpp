// TEST INFRASTRUCTURE (oracle/ref_shim): the handful of OpenCV core names LarVio::loadParameters (larvio.cpp:58-246) needs to
// read its %YAML:1.0 settings file: cv::FileStorage / FileNode (scalars, one nested map, !!opencv-matrix), cv::Mat with a
// cv::Rect sub-view, cv::Matx / cv::Vec and cv2eigen.  No OpenCV code; a line-oriented reader for the subset of YAML the
// reference's config files use.
#ifndef LVB_REF_SHIM_OPENCV_CORE_UTILITY
#define LVB_REF_SHIM_OPENCV_CORE_UTILITY
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <Eigen/Dense>
namespace cv {
struct Rect { int x, y, width, height; Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
class Mat {
 public:
  int rows, cols; std::vector<double> data;
  Mat() : rows(0), cols(0) {}
  Mat(int r, int c) : rows(r), cols(c), data((size_t)r * c, 0.0) {}
  double at(int i, int j) const { return data[(size_t)i * cols + j]; }
  Mat operator()(const Rect& r) const {
    Mat o(r.height, r.width);
    for (int i = 0; i < r.height; ++i) for (int j = 0; j < r.width; ++j) o.data[(size_t)i * r.width + j] = at(r.y + i, r.x + j);
    return o;
  }
};
template <class T, int M, int N> class Matx {
 public:
  T val[M * N];
  Matx() { for (auto& v : val) v = T(0); }
  Matx(const Mat& m) {
    if (m.rows != M || m.cols != N) { std::fprintf(stderr, "ref_shim/cv: Matx from a Mat of another size\n"); std::abort(); }
    for (int i = 0; i < M * N; ++i) val[i] = (T)m.data[i];
  }
  T operator()(int i, int j) const { return val[i * N + j]; }
  T operator()(int i) const { return val[i]; }
};
template <class T, int M> using Vec = Matx<T, M, 1>;
typedef Matx<double, 3, 3> Matx33d;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;
template <class T, int M, int N, class E> void cv2eigen(const Matx<T, M, N>& src, E& dst) {
  dst = Eigen::Matrix<double, M, N>::Zero();
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) dst(i, j) = src(i, j);
}

class FileNode {
 public:
  enum Kind { NONE, SCALAR, MAP, MATRIX };
  Kind kind; std::string text; std::map<std::string, std::shared_ptr<FileNode>> kids; Mat mat;
  FileNode() : kind(NONE) {}
  bool empty() const { return kind == NONE; }
  FileNode operator[](const std::string& k) const { auto it = kids.find(k); return it == kids.end() ? FileNode() : *it->second; }
  FileNode operator[](const char* k) const { return (*this)[std::string(k)]; }
  double real() const { return kind == SCALAR ? std::strtod(text.c_str(), nullptr) : 0.0; }
  operator double() const { return real(); }
  operator float() const { return (float)real(); }
  operator int() const { return (int)std::lrint(real()); }
  operator std::string() const { return kind == SCALAR ? text : std::string(); }
};
inline void operator>>(const FileNode& n, std::string& s) { s = (std::string)n; }
inline void operator>>(const FileNode& n, Mat& m) { m = n.mat; }
inline void operator>>(const FileNode& n, double& v) { v = (double)n; }
inline void operator>>(const FileNode& n, int& v) { v = (int)n; }

class FileStorage {
 public:
  enum Mode { READ = 0 };
  FileStorage(const std::string& path, int) : ok_(false) {
    std::ifstream f(path);
    if (!f) return;
    std::vector<std::pair<int, std::string>> lines;  // (indent, content without comment)
    std::string ln;
    while (std::getline(f, ln)) {
      if (!ln.empty() && ln[0] == '%') continue;
      bool inq = false; size_t cut = ln.size();
      for (size_t i = 0; i < ln.size(); ++i) { if (ln[i] == '"') inq = !inq; if (ln[i] == '#' && !inq) { cut = i; break; } }
      ln = ln.substr(0, cut);
      size_t a = ln.find_first_not_of(" \t\r");
      if (a == std::string::npos) continue;
      size_t b = ln.find_last_not_of(" \t\r");
      if (ln.substr(a, b - a + 1) == "---") continue;
      lines.push_back({(int)a, ln.substr(a, b - a + 1)});
    }
    size_t i = 0;
    parse_map(lines, i, lines.empty() ? 0 : lines[0].first, root_);
    root_.kind = FileNode::MAP;
    ok_ = true;
  }
  bool isOpened() const { return ok_; }
  void release() {}
  FileNode operator[](const std::string& k) const { return root_[k]; }
  FileNode operator[](const char* k) const { return root_[std::string(k)]; }
 private:
  bool ok_; FileNode root_;
  static std::string strip(std::string s) {
    size_t a = s.find_first_not_of(" \t"); if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t"); s = s.substr(a, b - a + 1);
    if (s.size() >= 2 && s.front() == '"' && s.back() == '"') s = s.substr(1, s.size() - 2);
    return s;
  }
  static void parse_map(const std::vector<std::pair<int, std::string>>& L, size_t& i, int indent, FileNode& into) {
    while (i < L.size() && L[i].first >= indent) {
      const std::string& s = L[i].second;
      size_t c = s.find(':');
      if (c == std::string::npos) { ++i; continue; }
      std::string key = strip(s.substr(0, c)), val = strip(s.substr(c + 1));
      auto node = std::make_shared<FileNode>();
      int my_indent = L[i].first;
      ++i;
      if (val == "!!opencv-matrix") {
        FileNode sub; size_t j = i; std::string data;
        // rows / cols / dt / data: [ ... ] possibly over several lines
        int rows = 0, cols = 0; bool in_data = false;
        while (j < L.size() && L[j].first > my_indent) {
          const std::string& t = L[j].second;
          if (!in_data) {
            size_t cc = t.find(':'); std::string k2 = strip(t.substr(0, cc)), v2 = cc == std::string::npos ? "" : strip(t.substr(cc + 1));
            if (k2 == "rows") rows = std::atoi(v2.c_str());
            else if (k2 == "cols") cols = std::atoi(v2.c_str());
            else if (k2 == "data") { in_data = true; data += v2; }
          } else data += " " + t;
          ++j;
          if (in_data && data.find(']') != std::string::npos) break;
        }
        i = j;
        for (auto& ch : data) if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
        std::istringstream is(data); double v; node->mat = Mat(rows, cols); size_t n = 0;
        while (is >> v && n < node->mat.data.size()) node->mat.data[n++] = v;
        node->kind = FileNode::MATRIX;
      } else if (val.empty()) {
        node->kind = FileNode::MAP;
        if (i < L.size() && L[i].first > my_indent) parse_map(L, i, L[i].first, *node);
      } else {
        node->kind = FileNode::SCALAR; node->text = val;
      }
      into.kids[key] = node;
    }
  }
};
}  // namespace cv
#endif
