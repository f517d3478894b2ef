// Steered-BRIEF (ORB) descriptors at arbitrary sub-pixel points, one warp per keypoint.
// Replaces ORBdescriptor::IC_Angle (ORBDescriptor.cpp:486-513), computeOrbDescriptor (:334-382),
// computeDescriptors (:385-416) and computeDescriptorDistance (ORBDescriptor.h:44-60) plus the
// "Hamming <= 58" gate of image_processor.cpp:450-462, 689-699, 917-930.
// Image sources: the angle is taken on the UNBLURRED level-0 image, the 256 pair tests on the 7x7
// blurred image whose 32-px frame is unblurred REFLECT_101: taps that land
// outside the image read the padded raw level instead of the blur plane.
#include "lvb_internal.h"
#include "fe_device.cuh"

namespace {

__constant__ signed char c_pattern[1024] = {
#include "orb_pattern.inc"
};
// umax of a radius-15 disc (ORBDescriptor.cpp:316-329)
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2 (degrees), SURVEY.md App. A.3 — no FMA contraction.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 57.283626556396484f, p3 = -18.66744613647461f, p5 = 8.914000511169434f, p7 = -2.539724588394165f;
  const float eps = 2.220446049250313e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

struct OrbArgs {
  const uint8_t* pyr; const uint8_t* blur; LvbPyramidLayout L;
  int stride;
  const float2* pts;      // [S][stride], indexed through perm when given
  const int* perm; const int* n_pts;
  float* angles;          // optional [S][stride]
  uint8_t* desc_out;      // optional [S][stride][32] (indexed by i, or by slot when out_by_slot)
  int out_by_slot;
  const uint8_t* desc_ref;  // optional [S][stride][32] (indexed by slot): Hamming gate against it
  uint8_t* status;        // optional [S][stride]: 1 if distance <= max_dist
  int* dist_out;          // optional [S][stride]
  int max_dist;
};

// angle + this lane's descriptor byte (8 pair tests) of one keypoint; all 32 lanes of the warp cooperate
__device__ __forceinline__ unsigned orb_describe(const uint8_t* raw, const uint8_t* blr, const LvbLevel& lv, float2 pt, int lane, float* angle_out) {
  const int cx = __float2int_rn(pt.x), cy = __float2int_rn(pt.y);
  // ---- intensity centroid on the raw image
  int m10 = 0, m01 = 0;
  {
    const int u = lane - 15;          // lanes 0..30 -> u = -15..15
    if (lane < 31) {
      for (int v = -15; v <= 15; ++v) {
        const int d = c_umax[v < 0 ? -v : v];
        if (u >= -d && u <= d) {
          const int val = raw[(ptrdiff_t)(cy + v) * lv.pitch + (cx + u)];
          m10 += u * val;
          m01 += v * val;
        }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  const float ang_rad = __fmul_rn(angle, 0.017453292519943295f);   // factorPI = (float)(CV_PI/180.f)
  const float ca = (float)cos((double)ang_rad), sa = (float)sin((double)ang_rad);
  // ---- 8 pair tests per lane -> one descriptor byte
  unsigned val = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const signed char* p = &c_pattern[(lane * 8 + k) * 4];
    int t[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float px = (float)p[2 * e], py = (float)p[2 * e + 1];
      const float x = __fsub_rn(__fmul_rn(px, ca), __fmul_rn(py, sa));
      const float y = __fadd_rn(__fmul_rn(px, sa), __fmul_rn(py, ca));
      const int gx = cx + __float2int_rn(x), gy = cy + __float2int_rn(y);
      if (gx >= 0 && gx < lv.w && gy >= 0 && gy < lv.h) t[e] = blr[(size_t)gy * lv.w + gx];
      else t[e] = raw[(ptrdiff_t)gy * lv.pitch + gx];
    }
    val |= (unsigned)(t[0] < t[1]) << k;
  }
  if (angle_out) *angle_out = angle;
  return val;
}

__global__ void __launch_bounds__(128) orb_kernel(OrbArgs a) {
  const int s = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= a.n_pts[s]) return;
  const int slot = a.perm ? a.perm[(size_t)s * a.stride + i] : i;
  const float2 pt = a.pts[(size_t)s * a.stride + slot];
  const LvbLevel lv = a.L.lv[0];
  float angle;
  const unsigned val = orb_describe(lvb_level_origin(a.pyr, a.L, s, 0), a.blur + (size_t)s * lv.w * lv.h, lv, pt, lane, &angle);
  const size_t oi = (size_t)s * a.stride + i;
  if (a.desc_out) a.desc_out[(a.out_by_slot ? ((size_t)s * a.stride + slot) : oi) * 32 + lane] = (uint8_t)val;
  if (a.angles && lane == 0) a.angles[oi] = angle;
  if (a.desc_ref) {
    const size_t ri = (size_t)s * a.stride + slot;
    int d = __popc(val ^ (unsigned)a.desc_ref[ri * 32 + lane]);
#pragma unroll
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) {
      if (a.status) a.status[oi] = d <= a.max_dist ? 1 : 0;
      if (a.dist_out) a.dist_out[oi] = d;
    }
  }
}

// The descriptor gate of one frame in ONE launch (image_processor.cpp:436-462, 673-699, 904-930).  blockIdx.z = 0: tracked
// features, descriptor at the current position against the one stored at birth (:807, kept forever).  blockIdx.z = 1: features
// detected in the previous frame: descriptor in the previous image (kept as the birth descriptor) and in the current image,
// compared with each other.
struct OrbGateArgs {
  const uint8_t* pyr_cur; const uint8_t* blur_cur; const uint8_t* pyr_prev; const uint8_t* blur_prev; LvbPyramidLayout L;
  int stride; int max_dist;
  const float2* cur_pts[2];     // [S][stride] by slot: tracked position in the current image, chain 0 / 1
  const float2* new_prev_pts;   // [S][stride] by slot: where the new features were detected (previous image)
  const int* perm[2]; const int* n_pts[2];
  const uint8_t* birth_desc;    // [S][stride][32] by slot: descriptors of the tracked features
  uint8_t* new_desc;            // [S][stride][32] by slot: birth descriptors of the new features (output)
  uint8_t* status[2];           // [S][stride] by rank
};
__global__ void __launch_bounds__(128) orb_gate_kernel(const __grid_constant__ OrbGateArgs a) {
  const int s = blockIdx.y, c = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= a.n_pts[c][s]) return;
  const size_t base = (size_t)s * a.stride;
  const int slot = a.perm[c][base + i];
  const LvbLevel lv = a.L.lv[0];
  const unsigned vcur = orb_describe(lvb_level_origin(a.pyr_cur, a.L, s, 0), a.blur_cur + (size_t)s * lv.w * lv.h, lv, a.cur_pts[c][base + slot], lane, nullptr);
  unsigned vref;
  if (c == 0) vref = a.birth_desc[(base + slot) * 32 + lane];
  else {
    vref = orb_describe(lvb_level_origin(a.pyr_prev, a.L, s, 0), a.blur_prev + (size_t)s * lv.w * lv.h, lv, a.new_prev_pts[base + slot], lane, nullptr);
    a.new_desc[(base + slot) * 32 + lane] = (uint8_t)vref;
  }
  int d = __popc(vcur ^ vref);
#pragma unroll
  for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) a.status[c][base + i] = d <= a.max_dist ? 1 : 0;
}

// ---------------------------------------------------------------- undistortPoints (App. A.7)
// radtan: 5 fixed-point iterations in double; equidistant: cv::fisheye (Newton on theta, 10 its).
struct UndArgs {
  LvbCamera cam;
  int stride; const float2* pts; const int* perm; const int* n_pts; float2* out; int to_pixels;
};

__global__ void undistort_kernel(UndArgs a) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_pts[s]) return;
  const int slot = a.perm ? a.perm[(size_t)s * a.stride + i] : i;
  a.out[(size_t)s * a.stride + i] = lvb_undistort_point(a.cam, a.pts[(size_t)s * a.stride + slot], a.to_pixels);
}

}  // namespace

LvbCamera lvb_camera(const LvbConfig& c) {
  LvbCamera cam;
  cam.fx = c.fx; cam.fy = c.fy; cam.cx = c.cx; cam.cy = c.cy;
  for (int i = 0; i < 4; ++i) cam.dist[i] = c.dist[i];
  cam.model = c.distortion_model;
  return cam;
}

int fe_orb_launch(LvbHandle* h, const uint8_t* pyr, const uint8_t* blur, int n_seq, int stride,
                  const float2* pts, const int* perm, const int* n_pts, float* angles, uint8_t* desc_out,
                  int out_by_slot, const uint8_t* desc_ref, uint8_t* status, int* dist_out) {
  OrbArgs a;
  a.pyr = pyr; a.blur = blur; a.L = h->fe.L; a.stride = stride; a.pts = pts; a.perm = perm; a.n_pts = n_pts;
  a.angles = angles; a.desc_out = desc_out; a.out_by_slot = out_by_slot; a.desc_ref = desc_ref;
  a.status = status; a.dist_out = dist_out; a.max_dist = 58;
  dim3 grd((stride + 3) / 4, n_seq);
  LVB_PROF(h, "orb_kernel");
  orb_kernel<<<grd, 128, 0, h->stream>>>(a);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

int fe_orb_gate_launch(LvbHandle* h, const uint8_t* pyr_cur, const uint8_t* blur_cur, const uint8_t* pyr_prev, const uint8_t* blur_prev,
                       int n_seq, int stride, const float2* const cur_pts[2], const float2* new_prev_pts, int* const perm[2],
                       int* const n_pts[2], const uint8_t* birth_desc, uint8_t* new_desc, uint8_t* const status[2]) {
  OrbGateArgs a;
  a.pyr_cur = pyr_cur; a.blur_cur = blur_cur; a.pyr_prev = pyr_prev; a.blur_prev = blur_prev; a.L = h->fe.L; a.stride = stride; a.max_dist = 58;
  for (int c = 0; c < 2; ++c) { a.cur_pts[c] = cur_pts[c]; a.perm[c] = perm[c]; a.n_pts[c] = n_pts[c]; a.status[c] = status[c]; }
  a.new_prev_pts = new_prev_pts; a.birth_desc = birth_desc; a.new_desc = new_desc;
  LVB_PROF(h, "orb_gate_kernel");
  orb_gate_kernel<<<dim3((stride + 3) / 4, n_seq, 2), 128, 0, h->stream>>>(a);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

int fe_undistort_launch(LvbHandle* h, int n_seq, int stride, const float2* pts, const int* perm,
                        const int* n_pts, float2* out, int to_pixels) {
  UndArgs a;
  a.cam = lvb_camera(h->cfg); a.stride = stride; a.pts = pts; a.perm = perm; a.n_pts = n_pts; a.out = out;
  a.to_pixels = to_pixels;
  dim3 grd((stride + 127) / 128, n_seq);
  LVB_PROF(h, "undistort_kernel");
  undistort_kernel<<<grd, 128, 0, h->stream>>>(a);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}
