// Fundamental-matrix RANSAC inlier masks == cv::findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask)
// at image_processor.cpp:498-500, 755-757, 968-970 (only the mask is consumed there).
// OpenCV (un-vendored third-party) algorithm restated in oracle/ransac.py and pinned against cv2:
//   n < 7            -> no mask (the reference then keeps every element, image_processor.h:219-223)
//   n == 7           -> all ones
//   8 <= n < 15      -> LMedS registrator (fundam.cpp switches below 15 points): 300 hypotheses,
//                       least median, sigma-scaled threshold.  For n <= 13 the median index n/2 falls
//                       inside the 7 exactly-fitted sample points, so OpenCV's winner is decided by
//                       rounding noise and cannot be reproduced (known non-parity edge)
//   n >= 15          -> RANSAC: cv::RNG(-1) MWC subsets of 7 (duplicates redrawn, last-point
//                       collinearity test), 7-point cubic, goodCount > max(best, 6) update rule with
//                       adaptive iteration count, Sampson-style max line distance^2 <= 1 as float.
// One CTA per point set.  Thread 0 replays the RNG stream (it does not depend on model quality), the
// 32 lanes of warp 0 solve 32 minimal problems at once in FP64, all warps score models, thread 0
// replays the sequential acceptance rule.  The 2-D null space comes from Gauss-Jordan with full
// pivoting instead of an SVD: the pencil of F matrices and therefore every candidate model is the same.
#include <float.h>
#include <string.h>
#include "lvb_internal.h"
#include "fe_device.cuh"

namespace {

constexpr int BATCH = 32;      // hypotheses per round
constexpr int MAXPTS = 512;

struct RansacSet {             // one chain of point sets ([S] sets, blockIdx.y selects the chain)
  const float2* p1; const float2* p2; const int* n;
  const int* perm;             // with `undistort`: p1/p2 are raw pixel coordinates indexed through perm (image_processor.cpp:479-486)
  uint8_t* mask; int* fail;    // fail[s]=1 -> chain already aborted, skip
};
struct RansacArgs {
  RansacSet set[2];
  int stride; const int* enable;
  int undistort; LvbCamera cam;   // undistortPoints(P = K) fused into the load of the point sets
  double threshold, confidence; int max_iters;
  unsigned long long* stats;   // [14] += 1 per point set in the 8..13 regime (OpenCV's LMedS winner there is rounding noise)
};

struct Rng {
  unsigned long long state;
  __device__ unsigned next() {
    state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

__device__ bool collinear_last(const float2* p, const int* idx, int count) {
  const int i = count - 1;
  const float2 pi = p[idx[i]];
  for (int j = 0; j < i; ++j) {
    const double dx1 = (double)p[idx[j]].x - (double)pi.x, dy1 = (double)p[idx[j]].y - (double)pi.y;
    for (int k = 0; k < j; ++k) {
      const double dx2 = (double)p[idx[k]].x - (double)pi.x, dy2 = (double)p[idx[k]].y - (double)pi.y;
      if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
    }
  }
  return false;
}

__device__ int solve_cubic(const double* c, double* x) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) return 0;
      x[0] = -a3 / a2; return 1;
    }
    double d = a2 * a2 - 4 * a1 * a3;
    if (d >= 0) {
      d = sqrt(d);
      const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
      if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
      else { x[0] = q2 / a1; x[1] = a3 / q2; }
      return d > 0 ? 2 : 1;
    }
    return 0;
  }
  a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
  const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
  const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
  const double Qc = Q * Q * Q;
  double d = Qc - R * R;
  if (d > 0) {
    const double theta = acos(R / sqrt(Qc)), sq = sqrt(Q);
    const double t0 = -2 * sq, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
    x[0] = t0 * cos(t1) - t2;
    x[1] = t0 * cos(t1 + (2. * 3.14159265358979323846 / 3)) - t2;
    x[2] = t0 * cos(t1 + (4. * 3.14159265358979323846 / 3)) - t2;
    return 3;
  }
  if (d == 0) {
    if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
    else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
    return x[0] == x[1] ? 1 : 2;
  }
  d = sqrt(-d);
  double e = pow(d + fabs(R), 1. / 3);
  if (R > 0) e = -e;
  x[0] = (e + Q / e) - a1 * (1. / 3);
  return 1;
}

// 7-point solver of one lane. A lives in shared memory, element-major: A[e*32 + lane].
__device__ int seven_point(double* A, const float2* p1, const float2* p2, const int* idx, double* F /*[3][9]*/) {
#define AA(r, c) A[((r) * 9 + (c)) * 32]
  // Hartley normalisation of the seven pairs, as run7Point does in the OpenCV (4.13) the oracle runs: centroid to the
  // origin, mean distance sqrt(2).  The candidate set is the same as for the raw-pixel system; the ORDER of the
  // candidates (which decides ties between equally good models of one sample) follows the normalised system.
  double m1cx = 0, m1cy = 0, m2cx = 0, m2cy = 0;
  for (int i = 0; i < 7; ++i) { m1cx += (double)p1[idx[i]].x; m1cy += (double)p1[idx[i]].y; m2cx += (double)p2[idx[i]].x; m2cy += (double)p2[idx[i]].y; }
  const double tcnt = 1. / 7;
  m1cx *= tcnt; m1cy *= tcnt; m2cx *= tcnt; m2cy *= tcnt;
  double scale1 = 0, scale2 = 0;
  for (int i = 0; i < 7; ++i) {
    const double ax = (double)p1[idx[i]].x - m1cx, ay = (double)p1[idx[i]].y - m1cy, bx = (double)p2[idx[i]].x - m2cx, by = (double)p2[idx[i]].y - m2cy;
    scale1 += sqrt(ax * ax + ay * ay); scale2 += sqrt(bx * bx + by * by);
  }
  scale1 *= tcnt; scale2 *= tcnt;
  if (scale1 < (double)FLT_EPSILON || scale2 < (double)FLT_EPSILON) return 0;
  scale1 = sqrt(2.) / scale1; scale2 = sqrt(2.) / scale2;
  for (int i = 0; i < 7; ++i) {
    const double x0 = ((double)p1[idx[i]].x - m1cx) * scale1, y0 = ((double)p1[idx[i]].y - m1cy) * scale1;
    const double x1 = ((double)p2[idx[i]].x - m2cx) * scale2, y1 = ((double)p2[idx[i]].y - m2cy) * scale2;
    AA(i, 0) = x1 * x0; AA(i, 1) = x1 * y0; AA(i, 2) = x1;
    AA(i, 3) = y1 * x0; AA(i, 4) = y1 * y0; AA(i, 5) = y1;
    AA(i, 6) = x0; AA(i, 7) = y0; AA(i, 8) = 1.0;
  }
  int perm[9];
  for (int c = 0; c < 9; ++c) perm[c] = c;
  for (int k = 0; k < 7; ++k) {
    // full pivot search in rows k.., columns k..
    int pr = k, pc = k; double best = -1;
    for (int r = k; r < 7; ++r)
      for (int c = k; c < 9; ++c) { const double v = fabs(AA(r, perm[c])); if (v > best) { best = v; pr = r; pc = c; } }
    if (pr != k) for (int c = 0; c < 9; ++c) { const double t = AA(k, c); AA(k, c) = AA(pr, c); AA(pr, c) = t; }
    { const int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t; }
    const double piv = AA(k, perm[k]);
    const double ip = 1.0 / piv;
    for (int c = 0; c < 9; ++c) AA(k, c) *= ip;
    for (int r = 0; r < 7; ++r) {
      if (r == k) continue;
      const double f = AA(r, perm[k]);
      if (f != 0) for (int c = 0; c < 9; ++c) AA(r, c) -= f * AA(k, c);
    }
  }
  double n1[9], n2[9];
  for (int c = 0; c < 9; ++c) { n1[c] = 0; n2[c] = 0; }
  n1[perm[7]] = 1.0; n2[perm[8]] = 1.0;
  for (int k = 0; k < 7; ++k) { n1[perm[k]] = -AA(k, perm[7]); n2[perm[k]] = -AA(k, perm[8]); }
#undef AA
  // Re-express the null space in the basis cv::SVDecomp(FULL_UV) returns (oracle/ransac.py
  // opencv_null_basis): Vt[7] = unit projection of R1 onto null(A), Vt[8] = unit projection of R2
  // made orthogonal to Vt[7], with R1/R2 the +-1/9 sign vectors OpenCV draws from RNG(0x12345678).
  // The basis fixes the ORDER of the candidate F matrices, which decides ties between equally good
  // models of one sample.
  const double R1[9] = {-1, -1, 1, -1, -1, -1, -1, 1, 1}, R2[9] = {1, -1, 1, 1, 1, 1, 1, -1, 1};
  double e1[9], e2[9];
  {
    double s1 = 0, dot = 0, s2 = 0;
    for (int c = 0; c < 9; ++c) s1 += n1[c] * n1[c];
    s1 = 1.0 / sqrt(s1);
    for (int c = 0; c < 9; ++c) { e1[c] = n1[c] * s1; dot += e1[c] * n2[c]; }
    for (int c = 0; c < 9; ++c) { e2[c] = n2[c] - dot * e1[c]; s2 += e2[c] * e2[c]; }
    s2 = 1.0 / sqrt(s2);
    for (int c = 0; c < 9; ++c) e2[c] *= s2;
  }
  double f1[9], f2[9];
  {
    double a1 = 0, b1 = 0, a2 = 0, b2 = 0;
    for (int c = 0; c < 9; ++c) { a1 += e1[c] * R1[c]; b1 += e2[c] * R1[c]; a2 += e1[c] * R2[c]; b2 += e2[c] * R2[c]; }
    double nn = 0, dot = 0, n2n = 0;
    for (int c = 0; c < 9; ++c) { f1[c] = a1 * e1[c] + b1 * e2[c]; nn += f1[c] * f1[c]; }
    nn = 1.0 / sqrt(nn);
    for (int c = 0; c < 9; ++c) { f1[c] *= nn; f2[c] = a2 * e1[c] + b2 * e2[c]; dot += f1[c] * f2[c]; }
    for (int c = 0; c < 9; ++c) { f2[c] -= dot * f1[c]; n2n += f2[c] * f2[c]; }
    n2n = 1.0 / sqrt(n2n);
    for (int c = 0; c < 9; ++c) { f2[c] *= n2n; f1[c] -= f2[c]; }
  }
  double c4[4], r[3] = {0, 0, 0};
  double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c4[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c4[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
          f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
          f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
          f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c4[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  c4[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
          f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
          f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
          f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  const int n = solve_cubic(c4, r);
  if (n < 1 || n > 3) return 0;
  for (int k = 0; k < n; ++k) {
    double lambda = r[k], mu = 1.;
    const double s = f1[8] * r[k] + f2[8];
    double* Fk = F + k * 9;
    if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; Fk[8] = 1.; }
    else Fk[8] = 0.;
    for (int i = 0; i < 8; ++i) Fk[i] = f1[i] * lambda + f2[i] * mu;
    // de-normalise: T2^T F T1, T = [[s, 0, -s cx], [0, s, -s cy], [0, 0, 1]]; then F(3,3) = 1
    double M[9], G[9];
    for (int j = 0; j < 3; ++j) {
      M[j] = scale2 * Fk[j]; M[3 + j] = scale2 * Fk[3 + j];
      M[6 + j] = (-scale2 * m2cx) * Fk[j] + (-scale2 * m2cy) * Fk[3 + j] + Fk[6 + j];
    }
    for (int i = 0; i < 3; ++i) {
      G[3 * i] = M[3 * i] * scale1; G[3 * i + 1] = M[3 * i + 1] * scale1;
      G[3 * i + 2] = M[3 * i] * (-scale1 * m1cx) + M[3 * i + 1] * (-scale1 * m1cy) + M[3 * i + 2];
    }
    if (fabs(G[8]) > (double)FLT_EPSILON) { const double ig = 1. / G[8]; for (int i = 0; i < 9; ++i) G[i] *= ig; }
    for (int i = 0; i < 9; ++i) Fk[i] = G[i];
  }
  return n;
}

__device__ __forceinline__ float sampson_err(const double* F, float2 a, float2 b) {
  const double x1 = a.x, y1 = a.y, x2 = b.x, y2 = b.y;
  double A = F[0] * x1 + F[1] * y1 + F[2], B = F[3] * x1 + F[4] * y1 + F[5], C = F[6] * x1 + F[7] * y1 + F[8];
  const double s2 = 1. / (A * A + B * B), d2 = x2 * A + y2 * B + C;
  A = F[0] * x2 + F[3] * y2 + F[6]; B = F[1] * x2 + F[4] * y2 + F[7]; C = F[2] * x2 + F[5] * y2 + F[8];
  const double s1 = 1. / (A * A + B * B), d1 = x1 * A + y1 * B + C;
  return (float)fmax(d1 * d1 * s1, d2 * d2 * s2);
}

__device__ int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.), 1.); ep = fmin(fmax(ep, 0.), 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num); denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

__global__ void __launch_bounds__(256) ransac_kernel(const __grid_constant__ RansacArgs aa) {
  const RansacSet a = aa.set[blockIdx.y];
  __shared__ float2 sp1[MAXPTS], sp2[MAXPTS];
  __shared__ double sA[63 * 32];
  __shared__ double sF[BATCH][3][9];
  __shared__ int s_nmodels[BATCH];
  __shared__ int s_idx[BATCH][7];
  __shared__ int s_valid[BATCH];       // subset found
  __shared__ int s_good[BATCH][3];
  __shared__ float s_med[BATCH][3];
  __shared__ float s_err[8][16];       // LMedS scratch (n < 15), one row per warp
  __shared__ int s_ctl[8];             // 0 done, 1 iter, 2 niters, 3 max_good, 4 best hyp, 5 best model, 6 have_best, 7 abort
  __shared__ double s_minmed;
  __shared__ unsigned long long s_rng;
  const int s = blockIdx.x;
  if (aa.enable && !aa.enable[s]) return;
  if (a.fail && a.fail[s]) return;
  const int n = min(a.n[s], MAXPTS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* mask = a.mask + (size_t)s * aa.stride;
  if (n < 7) { for (int i = tid; i < n; i += 256) mask[i] = 1; return; }   // "no mask" => keep all
  for (int i = tid; i < n; i += 256) {
    if (aa.undistort) {
      const int slot = a.perm[(size_t)s * aa.stride + i];
      sp1[i] = lvb_undistort_point(aa.cam, a.p1[(size_t)s * aa.stride + slot], 1);
      sp2[i] = lvb_undistort_point(aa.cam, a.p2[(size_t)s * aa.stride + slot], 1);
    } else {
      sp1[i] = a.p1[(size_t)s * aa.stride + i]; sp2[i] = a.p2[(size_t)s * aa.stride + i];
    }
  }
  __syncthreads();
  if (n == 7) { for (int i = tid; i < n; i += 256) mask[i] = 1; return; }
  const bool lmeds = n < 15;
  if (tid == 0 && n <= 13 && aa.stats) atomicAdd(&aa.stats[14], 1ull);
  if (tid == 0) {
    s_ctl[0] = 0; s_ctl[1] = 0; s_ctl[3] = 0; s_ctl[6] = 0; s_ctl[7] = 0;
    int niters = max(aa.max_iters, 1);
    if (lmeds) { niters = update_num_iters(aa.confidence, 0.45, 7, aa.max_iters); niters = max(niters, 3); }
    s_ctl[2] = niters;
    s_minmed = DBL_MAX;
    s_rng = 0xffffffffffffffffull;
  }
  __syncthreads();
  const float thr2 = (float)(aa.threshold * aa.threshold);
  double bestF[9];
  for (int q = 0; q < 9; ++q) bestF[q] = 0;

  // The adaptive iteration count collapses to a handful after the first good model (clean tracks: w ~ 0.95 -> ~4 iterations),
  // so the first round draws and scores 8 hypotheses, later rounds 32 (over-drawing is harmless: a call owns its RNG).
  for (int round = 0;; ++round) {
    const int bsz = round == 0 ? 8 : BATCH;
    // ---- thread 0: next bsz subsets from the RNG stream
    if (tid == 0) {
      Rng rng; rng.state = s_rng;
      const int max_attempts = lmeds ? 1000 : 10000;
      for (int hb = bsz; hb < BATCH; ++hb) s_valid[hb] = 0;
      for (int hb = 0; hb < bsz; ++hb) {
        bool found = false;
        int idx[7];
        for (int at = 0; at < max_attempts && !found; ++at) {
          for (int i = 0; i < 7; ++i) {
            int v;
            bool dup;
            do {
              v = rng.uniform(0, n);
              dup = false;
              for (int j = 0; j < i; ++j) dup |= (idx[j] == v);
            } while (dup);
            idx[i] = v;
          }
          found = !collinear_last(sp1, idx, 7) && !collinear_last(sp2, idx, 7);
        }
        s_valid[hb] = found;
        for (int i = 0; i < 7; ++i) s_idx[hb][i] = idx[i];
        if (!found) { for (int h2 = hb + 1; h2 < bsz; ++h2) s_valid[h2] = 0; break; }
      }
      s_rng = rng.state;
    }
    __syncthreads();
    // ---- warp 0: solve the 32 minimal problems
    if (warp == 0) {
      int nm = 0;
      if (s_valid[lane]) nm = seven_point(&sA[lane], sp1, sp2, s_idx[lane], &sF[lane][0][0]);
      s_nmodels[lane] = nm;
    }
    __syncthreads();
    // ---- all warps: score every (hypothesis, model)
    for (int hm = warp; hm < bsz * 3; hm += 8) {
      const int hb = hm / 3, m = hm - hb * 3;
      if (m >= s_nmodels[hb]) { if (lane == 0) { s_good[hb][m] = -1; s_med[hb][m] = 0.f; } continue; }
      const double* F = &sF[hb][m][0];
      if (!lmeds) {
        int good = 0;
        for (int i = lane; i < n; i += 32) good += (sampson_err(F, sp1[i], sp2[i]) <= thr2);
#pragma unroll
        for (int o = 16; o; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
        if (lane == 0) s_good[hb][m] = good;
      } else {
        if (lane < n) s_err[warp][lane] = sampson_err(F, sp1[lane], sp2[lane]);
        __syncwarp();
        if (lane == 0) {
          // element of rank n/2 (std::nth_element on the raw bits; errors are >= 0)
          float e[16];
          for (int i = 0; i < n; ++i) e[i] = s_err[warp][i];
          for (int i = 1; i < n; ++i) { float v = e[i]; int j = i - 1; while (j >= 0 && e[j] > v) { e[j + 1] = e[j]; --j; } e[j + 1] = v; }
          s_med[hb][m] = e[n / 2];
          s_good[hb][m] = 0;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- thread 0: replay the sequential acceptance rule
    if (tid == 0) {
      int iter = s_ctl[1], niters = s_ctl[2], max_good = s_ctl[3];
      int hb = 0;
      for (; hb < bsz && iter < niters; ++hb) {
        if (!s_valid[hb]) { if (iter == 0) s_ctl[7] = 1; s_ctl[0] = 1; break; }
        for (int m = 0; m < s_nmodels[hb]; ++m) {
          if (!lmeds) {
            const int good = s_good[hb][m];
            if (good > max(max_good, 6)) {
              max_good = good; s_ctl[4] = hb; s_ctl[5] = m; s_ctl[6] = 1;
              niters = update_num_iters(aa.confidence, (double)(n - good) / n, 7, niters);
            }
          } else {
            const double med = (double)s_med[hb][m];
            if (med < s_minmed) { s_minmed = med; s_ctl[4] = hb; s_ctl[5] = m; s_ctl[6] = 1; }
          }
        }
        ++iter;
      }
      s_ctl[1] = iter; s_ctl[2] = niters; s_ctl[3] = max_good;
      if (iter >= niters) s_ctl[0] = 1;
      // remember whether the best model lives in THIS batch
    }
    __syncthreads();
    // every thread keeps a private copy of the best model when it was (re)assigned in this batch
    if (s_ctl[6] == 1) {
      const double* F = &sF[s_ctl[4]][s_ctl[5]][0];
      for (int q = 0; q < 9; ++q) bestF[q] = F[q];
    }
    __syncthreads();
    if (tid == 0 && s_ctl[6] == 1) s_ctl[6] = 2;   // 2 = have a best model from an earlier batch
    __syncthreads();
    if (s_ctl[0]) break;
  }
  // ---- final mask
  if (s_ctl[6] == 0) {
    // no model at all: cv leaves the mask unspecified; we report all-zero
    for (int i = tid; i < n; i += 256) mask[i] = 0;
    return;
  }
  float t = thr2;
  if (lmeds) {
    double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(s_minmed);
    sigma = fmax(sigma, 0.001);
    t = (float)(sigma * sigma);
  }
  for (int i = tid; i < n; i += 256) mask[i] = sampson_err(bestF, sp1[i], sp2[i]) <= t ? 1 : 0;
}

}  // namespace

int fe_ransac_launch(LvbHandle* h, int n_seq, int stride, const float2* p1, const float2* p2, const int* n,
                     uint8_t* mask, const int* enable, int* fail) {
  RansacArgs a;
  memset(&a, 0, sizeof(a));
  a.set[0].p1 = p1; a.set[0].p2 = p2; a.set[0].n = n; a.set[0].perm = nullptr; a.set[0].mask = mask; a.set[0].fail = fail;
  a.stride = stride; a.enable = enable; a.undistort = 0;
  a.threshold = 1.0; a.confidence = 0.99; a.max_iters = 1000; a.stats = h->fe.stats;
  LVB_PROF(h, "ransac_kernel");
  ransac_kernel<<<dim3(n_seq, 1), 256, 0, h->stream>>>(a);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

// Both chains of a frame (tracked + new features) in ONE launch, undistortPoints(P = K) of the raw pixel pairs fused
// into the load (image_processor.cpp:479-500, 736-757, 949-970): prev[c] / curr[c] are indexed by slot through perm[c].
int fe_ransac_launch2(LvbHandle* h, int n_seq, int stride, const float2* const prev[2], const float2* const curr[2],
                      int* const perm[2], int* const n[2], uint8_t* const mask[2], int* const fail[2]) {
  RansacArgs a;
  memset(&a, 0, sizeof(a));
  for (int c = 0; c < 2; ++c) {
    a.set[c].p1 = prev[c]; a.set[c].p2 = curr[c]; a.set[c].n = n[c]; a.set[c].perm = perm[c]; a.set[c].mask = mask[c]; a.set[c].fail = fail[c];
  }
  a.stride = stride; a.enable = nullptr; a.undistort = 1; a.cam = lvb_camera(h->cfg);
  a.threshold = 1.0; a.confidence = 0.99; a.max_iters = 1000; a.stats = h->fe.stats;
  LVB_PROF(h, "ransac_kernel");
  ransac_kernel<<<dim3(n_seq, 2), 256, 0, h->stream>>>(a);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}
