// Small FP64 helpers of the back end (math_utils.hpp:26-231 and the Eigen calls the reference makes).
#pragma once
#include <cuda_runtime.h>

struct V3 { double x, y, z; };
struct M3 { double m[9]; };   // row-major

__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double norm(V3 a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

__device__ __forceinline__ M3 m3_identity() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0; return r; }
__device__ __forceinline__ M3 m3_load(const double* p) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = p[i]; return r; }
__device__ __forceinline__ void m3_store(double* p, const M3& a) { for (int i = 0; i < 9; ++i) p[i] = a.m[i]; }
__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
__device__ __forceinline__ M3 m3_t(const M3& a) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[j * 3 + i]; return r; }
__device__ __forceinline__ V3 m3_vec(const M3& a, V3 v) {
  return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ V3 m3_tvec(const M3& a, V3 v) {   // a^T v
  return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ M3 m3_scale(const M3& a, double s) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s; return r; }
__device__ __forceinline__ M3 m3_add(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
__device__ __forceinline__ M3 m3_sub(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
__device__ __forceinline__ M3 skew(V3 w) {   // math_utils.hpp:26-45
  M3 r; r.m[0] = 0; r.m[1] = -w.z; r.m[2] = w.y; r.m[3] = w.z; r.m[4] = 0; r.m[5] = -w.x; r.m[6] = -w.y; r.m[7] = w.x; r.m[8] = 0; return r;
}
// Eigen Quaterniond(w,x,y,z).toRotationMatrix() for q = [x y z w]  (== quaternionToRotation, math_utils.hpp:149-163)
__device__ __forceinline__ M3 quat_to_rot(const double* q) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  M3 r;
  r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y - w * z); r.m[2] = 2 * (x * z + w * y);
  r.m[3] = 2 * (x * y + w * z); r.m[4] = 1 - 2 * (x * x + z * z); r.m[5] = 2 * (y * z - w * x);
  r.m[6] = 2 * (x * z - w * y); r.m[7] = 2 * (y * z + w * x); r.m[8] = 1 - 2 * (x * x + y * y);
  return r;
}
// Eigen Quaterniond(Matrix3d).coeffs() -> [x y z w]
__device__ __forceinline__ void rot_to_quat(const M3& R, double* q) {
  double t = R.m[0] + R.m[4] + R.m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R.m[7] - R.m[5]) * t; q[1] = (R.m[2] - R.m[6]) * t; q[2] = (R.m[3] - R.m[1]) * t;
  } else {
    int i = 0;
    if (R.m[4] > R.m[0]) i = 1;
    if (R.m[8] > R.m[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R.m[i * 4] - R.m[j * 4] - R.m[k * 4] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R.m[k * 3 + j] - R.m[j * 3 + k]) * t;
    q[j] = (R.m[j * 3 + i] + R.m[i * 3 + j]) * t;
    q[k] = (R.m[k * 3 + i] + R.m[i * 3 + k]) * t;
  }
}
// Hamilton product a*b of Eigen quaternions stored [x y z w]
__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* r) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by - ax * bz + ay * bw + az * bx;
  r[2] = aw * bz + ax * by - ay * bx + az * bw;
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
}
// smallAngleQuaternion, math_utils.hpp:92-110
__device__ __forceinline__ void small_angle_quat(V3 dtheta, double* q) {
  const V3 dq = dtheta * 0.5;
  const double n2 = dot(dq, dq);
  if (n2 <= 1) { q[0] = dq.x; q[1] = dq.y; q[2] = dq.z; q[3] = sqrt(1 - n2); }
  else { const double s = 1.0 / sqrt(1 + n2); q[0] = dq.x * s; q[1] = dq.y * s; q[2] = dq.z * s; q[3] = s; }
}
