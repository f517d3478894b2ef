// TEST INFRASTRUCTURE (oracle/ref_shim): the subset of the OpenCV C++ API that the reference's sources use
// (src/image_processor.cpp, src/ORBDescriptor.cpp, src/larvio.cpp::loadParameters), so that they compile UNMODIFIED from where
// they lie under /root/reference (Makefile targets `ref`, `ref_fe`).  OpenCV's C++ headers are not in this image; its Python
// module is.  What is here, written from scratch:
//  * value types whose arithmetic lives in OpenCV's HEADERS (Point_, Size_, Rect, Scalar, Matx / Vec with the same
//    accumulation order and rounding as matx.hpp, cvRound/cvFloor/cvCeil), cv::Mat as a typed, reference-counted buffer with
//    ROI views that remember their parent (filters may read real pixels outside an ROI, as OpenCV's do), cv::FileStorage for
//    the %YAML:1.0 subset of config/euroc.yaml;
//  * NO image processing: every OpenCV FUNCTION (CLAHE, buildOpticalFlowPyramid, calcOpticalFlowPyrLK, goodFeaturesToTrack,
//    undistortPoints, findFundamentalMat, GaussianBlur, copyMakeBorder, resize, Rodrigues, fastAtan2) is forwarded over a pipe to
//    oracle/cv_server.py, which runs it with the cv2 module of this image (OpenCV 4.13).  Drawing calls are no-ops.
// Never included by the product.
#ifndef LVB_REF_SHIM_CV_HPP
#define LVB_REF_SHIM_CV_HPP
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

inline int cvRound(double v) { return (int)std::lrint(v); }          // round half to even, like OpenCV's SSE2 cvtsd2si path
inline int cvRound(float v) { return (int)std::lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

inline void shim_die(const char* what) { std::fprintf(stderr, "ref_shim/cv: %s\n", what); std::abort(); }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };
enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8, RANSAC = 8, LMEDS = 4 };
enum { COLOR_GRAY2RGB = 8, COLOR_GRAY2BGR = 8 };

// ---------------------------------------------------------------------------------------------------------------------
template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  T dot(const Point_& o) const { return (T)(x * o.x + y * o.y); }
  Point_& operator*=(T s) { x = (T)(x * s); y = (T)(y * s); return *this; }
  Point_& operator+=(const Point_& o) { x = (T)(x + o.x); y = (T)(y + o.y); return *this; }
};
template <class T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>((T)(a.x - b.x), (T)(a.y - b.y)); }
template <class T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>((T)(a.x + b.x), (T)(a.y + b.y)); }
// types.hpp: Point_<_Tp> * float/double/int -> saturate_cast<_Tp>(a.x * b)
inline Point_<float> operator*(const Point_<float>& a, float b) { return Point_<float>(a.x * b, a.y * b); }
inline Point_<float> operator*(const Point_<float>& a, double b) { return Point_<float>((float)(a.x * b), (float)(a.y * b)); }
inline Point_<float> operator*(const Point_<float>& a, int b) { return Point_<float>(a.x * (float)b, a.y * (float)b); }
template <class T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <class T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }   // types.hpp

template <class T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct Scalar {
  double val[4];
  Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

struct TermCriteria {
  enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
  int type, maxCount; double epsilon;
  TermCriteria() : type(0), maxCount(0), epsilon(0) {}
  TermCriteria(int t, int n, double e) : type(t), maxCount(n), epsilon(e) {}
};

struct KeyPoint {
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

// ---------------------------------------------------------------------------------------------------------------------
// cv::Mat: a typed buffer shared between headers; an ROI view keeps the whole allocation, so a consumer can see the pixels
// around it (OpenCV's locateROI).
struct MatBuf { std::vector<uchar> bytes; int rows, cols, type; size_t step; };

inline int cv_elem_size(int type) {
  const int depth = type & 7, cn = (type >> 3) + 1;
  const int sz = depth == CV_8U ? 1 : depth == CV_32S ? 4 : depth == CV_32F ? 4 : depth == CV_64F ? 8 : 0;
  if (!sz) shim_die("unsupported Mat depth");
  return sz * cn;
}

class Mat {
 public:
  int rows, cols;
  size_t step;                 // bytes per row of the underlying allocation
  uchar* data;
  std::shared_ptr<MatBuf> buf;
  int ox, oy;                  // position of this header inside buf (elements)
  int type_;
  int pyr_handle, pyr_level;   // set on the outputs of buildOpticalFlowPyramid: names the source image kept by the server

  Mat() : rows(0), cols(0), step(0), data(nullptr), ox(0), oy(0), type_(0), pyr_handle(-1), pyr_level(-1) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& v) : Mat() { create(r, c, type); setTo(v); }
  Mat(const Mat& m, const Rect& r) : Mat() { *this = m(r); }

  void create(int r, int c, int type) {
    if (buf && rows == r && cols == c && type_ == type) return;     // like OpenCV: a matching header is kept (ROI outputs are written in place)
    auto b = std::make_shared<MatBuf>();
    b->rows = r; b->cols = c; b->type = type; b->step = (size_t)c * cv_elem_size(type);
    b->bytes.assign(b->step * (size_t)r + 16, 0);
    buf = b; rows = r; cols = c; type_ = type; step = b->step; ox = oy = 0; data = b->bytes.data(); pyr_handle = pyr_level = -1;
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  void release() { buf.reset(); rows = cols = 0; data = nullptr; step = 0; }
  bool empty() const { return !buf || rows == 0 || cols == 0; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return (size_t)cv_elem_size(type_); }
  size_t elemSize1() const { return elemSize() / channels(); }
  size_t step1() const { return step / elemSize1(); }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }

  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }

  Mat operator()(const Rect& r) const {
    if (r.x < 0 || r.y < 0 || r.x + r.width > cols || r.y + r.height > rows) shim_die("Mat ROI outside the matrix");
    Mat m; m.buf = buf; m.rows = r.height; m.cols = r.width; m.step = step; m.type_ = type_;
    m.ox = ox + r.x; m.oy = oy + r.y; m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    return m;
  }
  Mat row(int r) const { return (*this)(Rect(0, r, cols, 1)); }
  Mat clone() const {
    Mat m; if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat& dst) const { Mat c = clone(); if (dst.buf && dst.rows == rows && dst.cols == cols && dst.type_ == type_) dst.assign_pixels(c); else dst = c; }
  void assign_pixels(const Mat& src) {
    if (src.rows != rows || src.cols != cols || src.type_ != type_) shim_die("pixel assignment between matrices of different shape");
    for (int r = 0; r < rows; ++r) std::memcpy(ptr(r), src.ptr(r), (size_t)cols * elemSize());
  }
  Mat& setTo(const Scalar& v) {
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols * channels(); ++c) {
        const double s = v.val[c % channels()];
        switch (depth()) {
          case CV_8U: ptr<uchar>(r)[c] = (uchar)s; break;
          case CV_32S: ptr<int>(r)[c] = (int)s; break;
          case CV_32F: ptr<float>(r)[c] = (float)s; break;
          default: ptr<double>(r)[c] = s;
        }
      }
    return *this;
  }
  Mat& operator=(const Scalar& v) { return setTo(v); }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
  // real pixels available around this header inside its allocation (OpenCV: locateROI)
  void margins(int& top, int& bottom, int& left, int& right) const {
    top = oy; left = ox; bottom = buf ? buf->rows - oy - rows : 0; right = buf ? buf->cols - ox - cols : 0;
  }
  // this header grown by (t, b, l, r) elements inside its allocation
  Mat grown(int t, int b, int l, int r) const {
    Mat m = *this; m.rows = rows + t + b; m.cols = cols + l + r; m.ox = ox - l; m.oy = oy - t;
    m.data = data - (size_t)t * step - (size_t)l * elemSize();
    return m;
  }
};

template <class T> struct MatDepth;
template <> struct MatDepth<uchar> { enum { value = CV_8U }; };
template <> struct MatDepth<int> { enum { value = CV_32S }; };
template <> struct MatDepth<float> { enum { value = CV_32F }; };
template <> struct MatDepth<double> { enum { value = CV_64F }; };

template <class T> class Mat_;
template <class T> struct MatCommaInitializer_ {
  Mat m; int k;
  explicit MatCommaInitializer_(const Mat& m_) : m(m_), k(0) {}
  template <class U> MatCommaInitializer_& operator,(U v) { m.at<T>(k / m.cols, k % m.cols) = (T)v; ++k; return *this; }
  operator Mat() const { return m; }
};
template <class T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, MatDepth<T>::value) {}
  T& operator()(int r, int c) { return at<T>(r, c); }
};
template <class T, class U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) { MatCommaInitializer_<T> ci(m); return (ci, v); }

// ---------------------------------------------------------------------------------------------------------------------
// Matx / Vec: arithmetic as in OpenCV's matx.hpp (sums start at 0 and accumulate left to right in _Tp).
template <class T, int M, int N> class Matx {
 public:
  T val[M * N];
  Matx() { for (int i = 0; i < M * N; ++i) val[i] = T(0); }
  Matx(T v0, T v1) : Matx() { static_assert(M * N >= 2, ""); val[0] = v0; val[1] = v1; }
  Matx(T v0, T v1, T v2) : Matx() { static_assert(M * N >= 3, ""); val[0] = v0; val[1] = v1; val[2] = v2; }
  Matx(T v0, T v1, T v2, T v3) : Matx() { static_assert(M * N >= 4, ""); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  Matx(T v0, T v1, T v2, T v3, T v4, T v5, T v6, T v7, T v8) : Matx() {
    static_assert(M * N >= 9, ""); const T v[9] = {v0, v1, v2, v3, v4, v5, v6, v7, v8}; for (int i = 0; i < 9; ++i) val[i] = v[i];
  }
  Matx(const Mat& m) {      // Mat -> Matx (mat.inl.hpp): element-wise conversion
    if (m.rows != M || m.cols != N) shim_die("Matx from a Mat of another size");
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j)
      val[i * N + j] = m.depth() == CV_64F ? (T)m.at<double>(i, j) : m.depth() == CV_32F ? (T)m.at<float>(i, j) : (T)m.at<uchar>(i, j);
  }
  template <class U> explicit Matx(const Matx<U, M, N>& o) { for (int i = 0; i < M * N; ++i) val[i] = (T)o.val[i]; }
  static Matx eye() { Matx m; for (int i = 0; i < (M < N ? M : N); ++i) m.val[i * N + i] = T(1); return m; }
  static Matx zeros() { return Matx(); }
  T& operator()(int i, int j) { return val[i * N + j]; }
  const T& operator()(int i, int j) const { return val[i * N + j]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  Matx<T, N, M> t() const { Matx<T, N, M> r; for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) r.val[j * M + i] = val[i * N + j]; return r; }
  // 3x3: Matx_FastInvOp (operations.hpp): adjugate times 1/det
  Matx<T, N, M> inv() const {
    static_assert(M == 3 && N == 3, "only the 3x3 inverse is provided");
    const Matx& a = *this; Matx<T, 3, 3> b;
    T d = a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) + a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    if (d == 0) return Matx<T, 3, 3>();
    d = 1 / d;
    b(0, 0) = (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * d; b(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * d; b(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * d;
    b(1, 0) = (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * d; b(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * d; b(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * d;
    b(2, 0) = (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * d; b(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * d; b(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * d;
    return b;
  }
};
template <class T, int M> class Vec : public Matx<T, M, 1> {
 public:
  Vec() {}
  Vec(T a, T b) : Matx<T, M, 1>(a, b) {}
  Vec(T a, T b, T c) : Matx<T, M, 1>(a, b, c) {}
  Vec(T a, T b, T c, T d) : Matx<T, M, 1>(a, b, c, d) {}
  Vec(const Matx<T, M, 1>& o) : Matx<T, M, 1>(o) {}
  Vec(const Mat& m) : Matx<T, M, 1>(m) {}
  template <class U> Vec(const Vec<U, M>& o) { for (int i = 0; i < M; ++i) this->val[i] = (T)o.val[i]; }    // Vec<T2> conversion operator of matx.hpp
  T& operator[](int i) { return this->val[i]; }
  const T& operator[](int i) const { return this->val[i]; }
  Vec& operator+=(const Vec& o) { for (int i = 0; i < M; ++i) this->val[i] = (T)(this->val[i] + o.val[i]); return *this; }
  Vec& operator*=(float s) { for (int i = 0; i < M; ++i) this->val[i] = (T)(this->val[i] * s); return *this; }
  Vec& operator*=(double s) { for (int i = 0; i < M; ++i) this->val[i] = (T)(this->val[i] * s); return *this; }
};
template <class T, int M, int N, int L> inline Matx<T, M, N> operator*(const Matx<T, M, L>& a, const Matx<T, L, N>& b) {
  Matx<T, M, N> c;
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { T s = 0; for (int k = 0; k < L; ++k) s += a(i, k) * b(k, j); c(i, j) = s; }
  return c;
}
template <class T, int M, int N> inline Vec<T, M> operator*(const Matx<T, M, N>& a, const Vec<T, N>& b) {
  Vec<T, M> c;
  for (int i = 0; i < M; ++i) { T s = 0; for (int k = 0; k < N; ++k) s += a(i, k) * b.val[k]; c.val[i] = s; }
  return c;
}
// Matx33d * Vec3f (image_processor.cpp:254): the product is formed in double (the oracle's reading, oracle/frontend.py:_integrate_imu)
template <int M, int N> inline Vec<double, M> operator*(const Matx<double, M, N>& a, const Vec<float, N>& b) {
  Vec<double, M> c;
  for (int i = 0; i < M; ++i) { double s = 0; for (int k = 0; k < N; ++k) s += a(i, k) * (double)b.val[k]; c.val[i] = s; }
  return c;
}
template <class T, int M, int N> inline Matx<T, M, N> operator-(const Matx<T, M, N>& a) { Matx<T, M, N> c; for (int i = 0; i < M * N; ++i) c.val[i] = -a.val[i]; return c; }
template <class T, int M> inline Vec<T, M> operator*(const Vec<T, M>& a, double s) { Vec<T, M> c; for (int i = 0; i < M; ++i) c.val[i] = (T)(a.val[i] * s); return c; }   // saturate_cast<_Tp>(a * alpha)
template <class T, int M> inline Vec<T, M> operator*(const Vec<T, M>& a, float s) { Vec<T, M> c; for (int i = 0; i < M; ++i) c.val[i] = (T)(a.val[i] * s); return c; }
template <class T, int M> inline Vec<T, M> operator*(const Vec<T, M>& a, int s) { Vec<T, M> c; for (int i = 0; i < M; ++i) c.val[i] = (T)(a.val[i] * s); return c; }
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<float, 3, 3> Matx33f;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;
typedef Vec<float, 3> Vec3f;
typedef Vec<int, 2> Vec2i;
typedef Vec<double, 2> Vec2d;

template <class T, int M, int N, class E> inline void cv2eigen(const Matx<T, M, N>& src, E& dst) {
  E tmp = E::Zero();
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) tmp(i, j) = src(i, j);
  dst = tmp;
}

// ---------------------------------------------------------------------------------------------------------------------
// the pipe to oracle/cv_server.py
struct NdArray { int dtype; std::vector<int> dims; std::vector<uchar> bytes; };   // dtype: 0 u8, 1 i32, 2 f32, 3 f64
class Server {
 public:
  static Server& get() { static Server s; return s; }
  std::vector<NdArray> call(int op, const std::vector<NdArray>& args) {
    start();
    wr32(op); wr32((int)args.size());
    for (const NdArray& a : args) { wr32(a.dtype); wr32((int)a.dims.size()); for (int d : a.dims) wr32(d); wr(a.bytes.data(), a.bytes.size()); }
    std::fflush(to_);
    const int n = rd32();
    std::vector<NdArray> out(n);
    for (NdArray& a : out) {
      a.dtype = rd32(); const int nd = rd32(); a.dims.resize(nd); size_t cnt = 1; for (int& d : a.dims) { d = rd32(); cnt *= (size_t)d; }
      const int es = a.dtype == 0 ? 1 : a.dtype == 3 ? 8 : 4;
      a.bytes.resize(cnt * es); rd(a.bytes.data(), a.bytes.size());
    }
    return out;
  }
 private:
  FILE* to_ = nullptr; FILE* from_ = nullptr;
  void start() {
    if (to_) return;
    const char* py = std::getenv("LVB_CV_SERVER_PYTHON"); const char* script = std::getenv("LVB_CV_SERVER");
    if (!script) shim_die("LVB_CV_SERVER (path of oracle/cv_server.py) is not set");
    int a[2], b[2];
    if (pipe(a) || pipe(b)) shim_die("pipe failed");
    const pid_t pid = fork();
    if (pid < 0) shim_die("fork failed");
    if (pid == 0) {
      dup2(a[0], 0); dup2(b[1], 1); close(a[0]); close(a[1]); close(b[0]); close(b[1]);
      execlp(py ? py : "python", py ? py : "python", script, (char*)nullptr);
      _exit(127);
    }
    close(a[0]); close(b[1]);
    to_ = fdopen(a[1], "wb"); from_ = fdopen(b[0], "rb");
  }
  void wr(const void* p, size_t n) { if (n && std::fwrite(p, 1, n, to_) != n) shim_die("write to cv_server failed"); }
  void rd(void* p, size_t n) { if (n && std::fread(p, 1, n, from_) != n) shim_die("cv_server closed the pipe (see its stderr)"); }
  void wr32(int v) { wr(&v, 4); }
  int rd32() { int v; rd(&v, 4); return v; }
};

inline NdArray nd_from_mat(const Mat& m) {
  NdArray a; a.dtype = m.depth() == CV_8U ? 0 : m.depth() == CV_32S ? 1 : m.depth() == CV_32F ? 2 : 3;
  a.dims = {m.rows, m.cols}; if (m.channels() > 1) a.dims.push_back(m.channels());
  const size_t rowb = (size_t)m.cols * m.elemSize();
  a.bytes.resize(rowb * m.rows);
  for (int r = 0; r < m.rows; ++r) std::memcpy(a.bytes.data() + rowb * r, m.ptr(r), rowb);
  return a;
}
inline Mat mat_from_nd(const NdArray& a) {
  const int depth = a.dtype == 0 ? CV_8U : a.dtype == 1 ? CV_32S : a.dtype == 2 ? CV_32F : CV_64F;
  const int cn = a.dims.size() > 2 ? a.dims[2] : 1;
  Mat m(a.dims.size() > 0 ? a.dims[0] : 0, a.dims.size() > 1 ? a.dims[1] : 1, CV_MAKETYPE(depth, cn));
  if (!a.bytes.empty()) std::memcpy(m.data, a.bytes.data(), a.bytes.size());
  return m;
}
template <class T> inline NdArray nd_scalars(int dtype, std::initializer_list<T> v) {
  NdArray a; a.dtype = dtype; a.dims = {(int)v.size()}; a.bytes.resize(v.size() * sizeof(T)); std::memcpy(a.bytes.data(), v.begin(), a.bytes.size()); return a;
}
inline NdArray nd_ints(std::initializer_list<int> v) { return nd_scalars<int>(1, v); }
inline NdArray nd_doubles(std::initializer_list<double> v) { return nd_scalars<double>(3, v); }
inline NdArray nd_points(const std::vector<Point2f>& p) {
  NdArray a; a.dtype = 2; a.dims = {(int)p.size(), 2}; a.bytes.resize(p.size() * 8);
  if (!p.empty()) std::memcpy(a.bytes.data(), p.data(), a.bytes.size());
  return a;
}
inline void points_from_nd(const NdArray& a, std::vector<Point2f>& p) {
  p.resize(a.dims.empty() ? 0 : a.dims[0]);
  if (!p.empty()) std::memcpy(p.data(), a.bytes.data(), p.size() * 8);
}
template <class T, int M, int N> inline NdArray nd_matx(const Matx<T, M, N>& m) {
  NdArray a; a.dtype = 3; a.dims = {M, N}; a.bytes.resize(sizeof(double) * M * N);
  double* d = reinterpret_cast<double*>(a.bytes.data()); for (int i = 0; i < M * N; ++i) d[i] = (double)m.val[i];
  return a;
}

// ---------------------------------------------------------------------------------------------------------------------
// OpenCV functions, forwarded
template <class T> using Ptr = std::shared_ptr<T>;

class CLAHE {
 public:
  double clip; Size tiles;
  void apply(const Mat& src, Mat& dst) {
    auto r = Server::get().call(1, {nd_from_mat(src), nd_doubles({clip}), nd_ints({tiles.width, tiles.height})});
    dst = mat_from_nd(r[0]);
  }
};
inline Ptr<CLAHE> createCLAHE(double clipLimit = 40.0, Size tileGridSize = Size(8, 8)) { auto p = std::make_shared<CLAHE>(); p->clip = clipLimit; p->tiles = tileGridSize; return p; }

inline int next_pyramid_handle() { static int h = 0; return ++h; }

// copyMakeBorder; an ROI source that is not BORDER_ISOLATED first takes the real pixels around it (OpenCV: locateROI + adjustROI)
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar& = Scalar()) {
  Mat s = src; int t = top, b = bottom, l = left, r = right;
  if (!(borderType & BORDER_ISOLATED)) {
    int mt, mb, ml, mr; src.margins(mt, mb, ml, mr);
    const int dt = std::min(mt, top), db = std::min(mb, bottom), dl = std::min(ml, left), dr = std::min(mr, right);
    s = src.grown(dt, db, dl, dr); t -= dt; b -= db; l -= dl; r -= dr;
  }
  auto out = Server::get().call(8, {nd_from_mat(s), nd_ints({t, b, l, r}), nd_ints({borderType & ~BORDER_ISOLATED})});
  Mat res = mat_from_nd(out[0]);
  if (dst.buf && dst.rows == res.rows && dst.cols == res.cols && dst.type() == res.type()) dst.assign_pixels(res); else dst = res;
}

// in-place capable GaussianBlur; real neighbours of an ROI are read like OpenCV's filter engine does (no BORDER_ISOLATED here)
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
  int mt, mb, ml, mr; src.margins(mt, mb, ml, mr);
  const int ry = ksize.height / 2, rx = ksize.width / 2;
  int dt = std::min(mt, ry), db = std::min(mb, ry), dl = std::min(ml, rx), dr = std::min(mr, rx);
  if (borderType & BORDER_ISOLATED) dt = db = dl = dr = 0;
  Mat s = src.grown(dt, db, dl, dr);
  auto out = Server::get().call(7, {nd_from_mat(s), nd_ints({ksize.width, ksize.height}), nd_doubles({sigmaX, sigmaY}), nd_ints({borderType & ~BORDER_ISOLATED})});
  Mat res = mat_from_nd(out[0])(Rect(dl, dt, src.cols, src.rows)).clone();
  if (dst.buf && dst.rows == res.rows && dst.cols == res.cols && dst.type() == res.type()) dst.assign_pixels(res); else dst = res;
}

inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR) {
  auto out = Server::get().call(9, {nd_from_mat(src), nd_ints({dsize.width, dsize.height}), nd_ints({interpolation})});
  Mat res = mat_from_nd(out[0]);
  if (dst.buf && dst.rows == res.rows && dst.cols == res.cols && dst.type() == res.type()) dst.assign_pixels(res); else dst = res;
}

// buildOpticalFlowPyramid(img, pyr, winSize, maxLevel, withDerivatives = true, ...): the level images come back from cv2; each one is
// placed inside a buffer padded by winSize with pyrBorder, as OpenCV lays them out (so that copyMakeBorder / filters on pyr[0] find
// the same neighbours), the derivative slots stay empty headers (only calcOpticalFlowPyrLK reads them, and cv2 rebuilds them)
inline int buildOpticalFlowPyramid(const Mat& img, std::vector<Mat>& pyramid, Size winSize, int maxLevel, bool withDerivatives = true,
                                   int pyrBorder = BORDER_REFLECT_101, int = BORDER_CONSTANT, bool = true) {
  const int h = next_pyramid_handle();
  auto out = Server::get().call(2, {nd_from_mat(img), nd_ints({winSize.width, winSize.height}), nd_ints({maxLevel}), nd_ints({h})});
  const int n = reinterpret_cast<const int*>(out[0].bytes.data())[0];
  pyramid.assign((size_t)(n + 1) * (withDerivatives ? 2 : 1), Mat());
  for (int l = 0; l <= n; ++l) {
    Mat lvl = mat_from_nd(out[1 + l]);
    Mat padded; copyMakeBorder(lvl, padded, winSize.height, winSize.height, winSize.width, winSize.width, pyrBorder | BORDER_ISOLATED);
    Mat view = padded(Rect(winSize.width, winSize.height, lvl.cols, lvl.rows));
    view.pyr_handle = h; view.pyr_level = l;
    pyramid[(size_t)l * (withDerivatives ? 2 : 1)] = view;
  }
  return n;
}

struct NoArrayT {};
inline NoArrayT noArray() { return NoArrayT(); }

inline void calcOpticalFlowPyrLK(const std::vector<Mat>& prevPyr, const std::vector<Mat>& nextPyr, const std::vector<Point2f>& prevPts,
                                 std::vector<Point2f>& nextPts, std::vector<uchar>& status, NoArrayT, Size winSize = Size(21, 21), int maxLevel = 3,
                                 TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0, double = 1e-4) {
  if (prevPyr.empty() || nextPyr.empty() || prevPyr[0].pyr_handle < 0 || nextPyr[0].pyr_handle < 0) shim_die("calcOpticalFlowPyrLK: not a pyramid of buildOpticalFlowPyramid");
  if (prevPts.empty()) { nextPts.clear(); status.clear(); return; }
  if (nextPts.size() != prevPts.size()) shim_die("calcOpticalFlowPyrLK: initial flow of another length");
  auto out = Server::get().call(3, {nd_ints({prevPyr[0].pyr_handle}), nd_ints({nextPyr[0].pyr_handle}), nd_points(prevPts), nd_points(nextPts),
                                    nd_ints({winSize.width, winSize.height}), nd_ints({maxLevel}), nd_doubles({(double)criteria.type, (double)criteria.maxCount, criteria.epsilon}), nd_ints({flags})});
  points_from_nd(out[0], nextPts);
  status.assign(out[1].bytes.begin(), out[1].bytes.end());
}

inline void goodFeaturesToTrack(const Mat& image, std::vector<Point2f>& corners, int maxCorners, double qualityLevel, double minDistance,
                                const Mat& mask = Mat(), int = 3, bool = false, double = 0.04) {
  NdArray m; m.dtype = 0; m.dims = {0, 0};
  auto out = Server::get().call(4, {nd_from_mat(image), nd_doubles({(double)maxCorners, qualityLevel, minDistance}), mask.empty() ? m : nd_from_mat(mask)});
  points_from_nd(out[0], corners);
}

inline void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Matx33d& K, const Vec4d& dist,
                            const Matx33d& R = Matx33d::eye(), const Matx33d& P = Matx33d::eye()) {
  auto out = Server::get().call(5, {nd_points(src), nd_matx(K), nd_matx(dist), nd_matx(R), nd_matx(P), nd_ints({0})});
  points_from_nd(out[0], dst);
}
namespace fisheye {
inline void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Matx33d& K, const Vec4d& dist,
                            const Matx33d& R = Matx33d::eye(), const Matx33d& P = Matx33d::eye()) {
  auto out = Server::get().call(5, {nd_points(src), nd_matx(K), nd_matx(dist), nd_matx(R), nd_matx(P), nd_ints({1})});
  points_from_nd(out[0], dst);
}
}  // namespace fisheye

inline Mat findFundamentalMat(const std::vector<Point2f>& p1, const std::vector<Point2f>& p2, int method, double param1, double param2,
                              std::vector<uchar>& mask) {
  auto out = Server::get().call(6, {nd_points(p1), nd_points(p2), nd_doubles({(double)method, param1, param2})});
  mask.assign(out[0].bytes.begin(), out[0].bytes.end());
  return Mat();
}

inline void Rodrigues(const Vec3f& src, Matx33f& dst) {
  NdArray a; a.dtype = 2; a.dims = {3}; a.bytes.resize(12); std::memcpy(a.bytes.data(), src.val, 12);
  auto out = Server::get().call(10, {a});
  if (out[0].dtype != 2) shim_die("Rodrigues: float32 in, float32 out expected");
  std::memcpy(dst.val, out[0].bytes.data(), 36);
}

inline float fastAtan2(float y, float x) {
  NdArray a; a.dtype = 2; a.dims = {2}; a.bytes.resize(8); const float v[2] = {y, x}; std::memcpy(a.bytes.data(), v, 8);
  auto out = Server::get().call(11, {a});
  float r; std::memcpy(&r, out[0].bytes.data(), 4); return r;
}

// drawing / colour conversion for the debug image (image_processor.cpp:1131-1167): no-ops, the image is never read back
inline void cvtColor(const Mat&, Mat&, int) {}
inline void destroyAllWindows() {}
template <class P> inline void circle(Mat&, P, int, const Scalar&, int = 1) {}
template <class P> inline void line(Mat&, P, P, const Scalar&, int = 1) {}

class RNG {   // cv::RNG (multiply-with-carry), only reached through ORBdescriptor::makeRandomPattern for patch sizes other than 31
 public:
  uint64_t state;
  explicit RNG(uint64_t s = 0xffffffff) : state(s ? s : 0xffffffff) {}
  unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// ---------------------------------------------------------------------------------------------------------------------
// cv::FileStorage for the %YAML:1.0 subset of the reference's settings files
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
    std::vector<std::pair<int, std::string>> lines;
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
        size_t j = i; std::string data; int rows = 0, cols = 0; bool in_data = false;
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
        std::istringstream is(data); double v; node->mat = Mat(rows, cols, CV_64FC1); int n = 0;
        while (is >> v && n < rows * cols) { node->mat.at<double>(n / cols, n % cols) = v; ++n; }
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
