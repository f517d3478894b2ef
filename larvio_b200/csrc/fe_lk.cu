// Pyramidal iterative Lucas-Kanade with initial flow, one warp per point, all levels in one launch.
// Replaces the six cv::calcOpticalFlowPyrLK call sites of image_processor.cpp (:368-377, :405-414,
// :558-567, :618-627, :830-839, :870-879; win 21x21, 3 levels, <=30 iterations, eps 0.01,
// minEigThreshold 1e-4, OPTFLOW_USE_INITIAL_FLOW) together with the gates that follow each call:
// predictFeatureTracking (:266-293) as prologue, the in-image test (:571-578) and the
// forward/backward consistency test (:630-642) as epilogue.  Arithmetic follows SURVEY.md App. A.5:
// Q14 bilinear weights, int16 patches with 5 fractional bits, float32 normal equations.
// Scharr derivatives are computed on the fly from the padded level image (zero outside the image,
// REFLECT_101 at its edge == the pad content), so no derivative pyramid exists in HBM.
#include "lvb_internal.h"

namespace {

constexpr int WIN = 21;
constexpr int NPIX = WIN * WIN;            // 441
constexpr int TILE = WIN + 3;              // 24: window + bilinear +1 + Scharr apron 1 on both sides
constexpr int DT = WIN + 1;                // 22: derivative grid
constexpr int WARPS = 4;
constexpr int W_BITS = 14;

// OpenCV's SSE path accumulates the normal equations in 4 float lanes + a scalar tail (pixels 16..20 of each
// window row); to be BIT-identical we replay exactly that order (oracle/lk_exact.py): per sum, accumulator p < 4
// receives, row by row, the terms of pixels p, 4+p, 8+p, 12+p (A) resp. the int-pair terms of both 8-pixel
// chunks (b); accumulator 4 receives pixels 16..20.  Terms are produced by all lanes, the 15 (A) / 10 (b)
// sequential chains run on one lane each.
constexpr int NTERM_A = WIN * 16 + WIN * 5 + 3;   // per sum: 336 lane terms + 105 tail terms (+3 pad -> 16-B aligned sums)
constexpr int NTERM_B = WIN * 8 + WIN * 5 + 3;    // per sum: 168 pair terms + 105 tail terms (+3 pad)
struct alignas(16) WarpSmem {
  short Iw[NPIX + 7];                      // 896
  short2 dIw[NPIX];                        // 1764
  short dd[NPIX + 7];                      // 896  I_t per iteration
  union {
    struct { uint8_t tile[TILE * TILE]; short2 dtile[DT * DT]; } st;   // 576 + 1936 (window set-up only)
    alignas(16) float termA[3 * NTERM_A];  // 5328
    alignas(16) float termB[2 * NTERM_B];  // 2208
  } u;
};

// Term layout per sum: accumulator p (0..3) owns WIN*K consecutive floats (row-major over (row, k)), followed by the
// WIN*5 tail terms; K = 4 for the A sums, 2 for the b sums.  Each chain adds its terms strictly in order, reading
// them as float4 (all segment lengths are multiples of 4 except the 105-term tail, padded by one zero... 105 = 26*4+1).
template <int K>
__device__ __forceinline__ float run_chain(const float* T, int acc) {
  float a = 0.f;
  if (acc < 4) {
    // WIN*K floats per accumulator: 84 (16-byte multiples) for K = 4, 42 (8-byte multiples) for K = 2 -> float2 loads
    const float2* q = reinterpret_cast<const float2*>(T + acc * (WIN * K));
#pragma unroll
    for (int i = 0; i < WIN * K / 2; ++i) {
      const float2 v = q[i];
      a = __fadd_rn(a, v.x); a = __fadd_rn(a, v.y);
    }
  } else {
    const float* qt = T + 4 * (WIN * K);          // 336 / 168 floats: 16-byte aligned
    const float4* q = reinterpret_cast<const float4*>(qt);
#pragma unroll
    for (int i = 0; i < (WIN * 5) / 4; ++i) {
      const float4 v = q[i];
      a = __fadd_rn(a, v.x); a = __fadd_rn(a, v.y); a = __fadd_rn(a, v.z); a = __fadd_rn(a, v.w);
    }
    a = __fadd_rn(a, qt[WIN * 5 - 1]);
  }
  return a;
}
// total = tail + ((l0 + l2) + (l1 + l3)); chains of this sum live on lanes base..base+4
__device__ __forceinline__ float combine_chains(float mine, int base) {
  const float l0 = __shfl_sync(0xffffffffu, mine, base), l1 = __shfl_sync(0xffffffffu, mine, base + 1);
  const float l2 = __shfl_sync(0xffffffffu, mine, base + 2), l3 = __shfl_sync(0xffffffffu, mine, base + 3);
  const float tl = __shfl_sync(0xffffffffu, mine, base + 4);
  return __fadd_rn(tl, __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct LkArgs {
  const uint8_t* pyrA; const uint8_t* pyrB;
  LvbPyramidLayout L;
  int stride;                 // per-sequence stride of point arrays
  const float2* ptsA;         // [S][stride] source points (indexed through perm if given)
  const int* perm;            // [S][stride] or null
  const int* n_pts;           // [S]
  const float2* init;         // [S][stride] initial flow (ignored when Hmat != null)
  int init_by_slot;
  const float* Hmat;          // [S][9] or null: initial flow = K R K^-1 * ptA
  float2* out;                // [S][stride] tracked positions (indexed by i, not by perm)
  uint8_t* status;            // [S][stride]
  int gate_mode;              // 0 none, 1 in-image, 2 in-image + |out - ref| <= 1
  const float2* ref;          // [S][stride] reference for gate 2 (indexed by perm like ptsA)
  int max_iter; double eps2; double min_eig;
  int max_level;
};

struct LkArgs2 { LkArgs a[2]; };
__global__ void __launch_bounds__(WARPS * 32) lk_kernel(const __grid_constant__ LkArgs2 aa) {
  __shared__ WarpSmem sm[WARPS];
  const LkArgs& a = aa.a[blockIdx.z];
  const int s = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * WARPS + warp;
  if (i >= a.n_pts[s]) return;
  WarpSmem& w = sm[warp];
  const int slot = a.perm ? a.perm[(size_t)s * a.stride + i] : i;
  const float2 pA = a.ptsA[(size_t)s * a.stride + slot];
  float2 nxt;
  if (a.Hmat) {
    const float* H = a.Hmat + (size_t)s * 9;
    // cv::Matx33f * Vec3f : s = 0; s += H(r,k)*v(k)   (image_processor.cpp:285-290)
    float q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = __fmul_rn(H[r * 3 + 0], pA.x);
      acc = __fadd_rn(acc, __fmul_rn(H[r * 3 + 1], pA.y));
      acc = __fadd_rn(acc, H[r * 3 + 2]);
      q[r] = acc;
    }
    nxt.x = __fdiv_rn(q[0], q[2]);
    nxt.y = __fdiv_rn(q[1], q[2]);
  } else {
    nxt = a.init[(size_t)s * a.stride + (a.init_by_slot ? slot : i)];
  }
  bool status = true;
  const float halfWin = (WIN - 1) * 0.5f;

  for (int level = a.max_level; level >= 0; --level) {
    const LvbLevel lv = a.L.lv[level];
    const float scale = 1.0f / (float)(1 << level);
    float2 prevPt = make_float2(__fmul_rn(pA.x, scale), __fmul_rn(pA.y, scale));
    if (level == a.max_level) nxt = make_float2(__fmul_rn(nxt.x, scale), __fmul_rn(nxt.y, scale));
    else nxt = make_float2(__fmul_rn(nxt.x, 2.f), __fmul_rn(nxt.y, 2.f));
    prevPt.x = __fsub_rn(prevPt.x, halfWin);
    prevPt.y = __fsub_rn(prevPt.y, halfWin);
    const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
    if (ipx < -WIN || ipx >= lv.w || ipy < -WIN || ipy >= lv.h) {
      if (level == 0) status = false;
      continue;
    }
    const uint8_t* orgA = lvb_level_origin(a.pyrA, a.L, s, level);
    const uint8_t* orgB = lvb_level_origin(a.pyrB, a.L, s, level);
    __syncwarp();
    // ---- stage the 24x24 source tile (origin ipx-1, ipy-1)
    for (int t = lane; t < TILE * TILE; t += 32) {
      const int ty = t / TILE, tx = t - ty * TILE;
      w.u.st.tile[t] = __ldg(orgA + (ptrdiff_t)(ipy - 1 + ty) * lv.pitch + (ipx - 1 + tx));
    }
    __syncwarp();
    // ---- Scharr derivatives on the 22x22 grid (zero outside the image)
    for (int t = lane; t < DT * DT; t += 32) {
      const int dy = t / DT, dx = t - dy * DT;
      const int gx = ipx + dx, gy = ipy + dy;
      short2 d = make_short2(0, 0);
      if (gx >= 0 && gx < lv.w && gy >= 0 && gy < lv.h) {
        const uint8_t* c = &w.u.st.tile[(dy + 1) * TILE + (dx + 1)];
        const int tl = c[-TILE - 1], tc = c[-TILE], tr = c[-TILE + 1];
        const int ml = c[-1], mr = c[1];
        const int bl = c[TILE - 1], bc = c[TILE], br = c[TILE + 1];
        d.x = (short)(3 * (tr + br - tl - bl) + 10 * (mr - ml));
        d.y = (short)(3 * (bl + br - tl - tr) + 10 * (bc - tc));
      }
      w.u.st.dtile[t] = d;
    }
    __syncwarp();
    // ---- window of the previous image
    float fa = __fsub_rn(prevPt.x, (float)ipx), fb = __fsub_rn(prevPt.y, (float)ipy);
    int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    for (int p = lane; p < NPIX; p += 32) {
      const int y = p / WIN, x = p - y * WIN;
      const uint8_t* c = &w.u.st.tile[(y + 1) * TILE + (x + 1)];
      const int ival = (c[0] * iw00 + c[1] * iw01 + c[TILE] * iw10 + c[TILE + 1] * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
      const short2 d00 = w.u.st.dtile[y * DT + x], d01 = w.u.st.dtile[y * DT + x + 1];
      const short2 d10 = w.u.st.dtile[(y + 1) * DT + x], d11 = w.u.st.dtile[(y + 1) * DT + x + 1];
      const int ix = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
      const int iy = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
      w.Iw[p] = (short)ival;
      w.dIw[p] = make_short2((short)ix, (short)iy);
    }
    __syncwarp();
    // terms of A11, A12, A22 in OpenCV's lane order (the tile/dtile staging area is dead from here on)
    for (int p = lane; p < NPIX; p += 32) {
      const int y = p / WIN, x = p - y * WIN;
      const short2 d = w.dIw[p];
      const float fx = (float)d.x, fy = (float)d.y;
      const int slot = (x < 16) ? ((x & 3) * (WIN * 4) + y * 4 + (x >> 2)) : (WIN * 16 + y * 5 + (x - 16));
      w.u.termA[slot] = __fmul_rn(fx, fx);
      w.u.termA[NTERM_A + slot] = __fmul_rn(fx, fy);
      w.u.termA[2 * NTERM_A + slot] = __fmul_rn(fy, fy);
    }
    __syncwarp();
    float chainv = 0.f;
    if (lane < 15) chainv = run_chain<4>(w.u.termA + (lane / 5) * NTERM_A, lane % 5);
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    float A11 = __fmul_rn(combine_chains(chainv, 0), FLT_SCALE);
    float A12 = __fmul_rn(combine_chains(chainv, 5), FLT_SCALE);
    float A22 = __fmul_rn(combine_chains(chainv, 10), FLT_SCALE);
    __syncwarp();
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dif = __fsub_rn(A11, A22);
    const float rad = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
    if ((double)minEig < a.min_eig || D < 1.1920929e-07f) {
      if (level == 0) status = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);
    float2 np = make_float2(__fsub_rn(nxt.x, halfWin), __fsub_rn(nxt.y, halfWin));
    int poff[14];            // byte offset of this lane's window pixels inside level B (pixel p = lane + 32k)
#pragma unroll
    for (int k2 = 0; k2 < 14; ++k2) { const int p = lane + 32 * k2; const int y = p / WIN; poff[k2] = y * lv.pitch + (p - y * WIN); }
    float2 prevDelta = make_float2(0.f, 0.f);
    for (int j = 0; j < a.max_iter; ++j) {
      const int inx = (int)floorf(np.x), iny = (int)floorf(np.y);
      if (inx < -WIN || inx >= lv.w || iny < -WIN || iny >= lv.h) {
        if (level == 0) status = false;
        break;
      }
      fa = __fsub_rn(np.x, (float)inx);
      fb = __fsub_rn(np.y, (float)iny);
      iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
      iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
      iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
      iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      const uint8_t* jbase = orgB + (ptrdiff_t)iny * lv.pitch + inx;
#pragma unroll
      for (int k2 = 0; k2 < 14; ++k2) {
        const int p = lane + 32 * k2;
        if (p < NPIX) {
          const uint8_t* c = jbase + poff[k2];
          const int j00 = __ldg(c), j01 = __ldg(c + 1), j10 = __ldg(c + lv.pitch), j11 = __ldg(c + lv.pitch + 1);
          w.dd[p] = (short)(((j00 * iw00 + j01 * iw01 + j10 * iw10 + j11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - (int)w.Iw[p]);
        }
      }
      __syncwarp();
      // pair terms (both 8-pixel chunks of every row) and tail terms, then the 10 chains
      for (int t = lane; t < WIN * 13; t += 32) {
        float t1, t2;
        int slot;
        if (t < WIN * 8) {
          const int y = t >> 3, ch = (t >> 2) & 1, p = t & 3;
          const int i0 = y * WIN + 8 * ch + p, i1 = i0 + 4;
          const int d0 = w.dd[i0], d1 = w.dd[i1];
          const short2 g0 = w.dIw[i0], g1 = w.dIw[i1];
          t1 = (float)(d0 * g0.x + d1 * g1.x);
          t2 = (float)(d0 * g0.y + d1 * g1.y);
          slot = p * (WIN * 2) + y * 2 + ch;
        } else {
          const int q = t - WIN * 8, y = q / 5, x = 16 + q - y * 5;
          const int i0 = y * WIN + x;
          const int d0 = w.dd[i0];
          const short2 g0 = w.dIw[i0];
          t1 = (float)(d0 * g0.x);
          t2 = (float)(d0 * g0.y);
          slot = t;
        }
        w.u.termB[slot] = t1;
        w.u.termB[NTERM_B + slot] = t2;
      }
      __syncwarp();
      float cv = 0.f;
      if (lane < 10) cv = run_chain<2>(w.u.termB + (lane / 5) * NTERM_B, lane % 5);
      const float b1 = __fmul_rn(combine_chains(cv, 0), FLT_SCALE);
      const float b2 = __fmul_rn(combine_chains(cv, 5), FLT_SCALE);
      __syncwarp();
      float2 delta;
      delta.x = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      delta.y = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      np.x = __fadd_rn(np.x, delta.x);
      np.y = __fadd_rn(np.y, delta.y);
      nxt = make_float2(__fadd_rn(np.x, halfWin), __fadd_rn(np.y, halfWin));
      const double dd = (double)delta.x * (double)delta.x + (double)delta.y * (double)delta.y;
      if (dd <= a.eps2) break;
      if (j > 0 && (double)fabsf(__fadd_rn(delta.x, prevDelta.x)) < 0.01 && (double)fabsf(__fadd_rn(delta.y, prevDelta.y)) < 0.01) {
        nxt.x = __fsub_rn(nxt.x, __fmul_rn(delta.x, 0.5f));
        nxt.y = __fsub_rn(nxt.y, __fmul_rn(delta.y, 0.5f));
        break;
      }
      prevDelta = delta;
    }
  }

  if (lane == 0) {
    const LvbLevel l0 = a.L.lv[0];
    if (status && a.gate_mode >= 1) {
      if (nxt.y < 0.f || nxt.y > (float)(l0.h - 1) || nxt.x < 0.f || nxt.x > (float)(l0.w - 1)) status = false;
    }
    if (status && a.gate_mode == 2) {
      const float2 r = a.ref[(size_t)s * a.stride + slot];
      const float dx = __fsub_rn(nxt.x, r.x), dy = __fsub_rn(nxt.y, r.y);
      const float dis = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
      if (dis > 1.f) status = false;
    }
    a.out[(size_t)s * a.stride + i] = nxt;
    a.status[(size_t)s * a.stride + i] = status ? 1 : 0;
  }
}

// ---- staged variant (LVB_EXPERIMENT=lk_fused): identical arithmetic and summation order; the I_t pass and
// the term pass of every iteration are one pass (each term's lane gathers its own 2 x 4 or 4 source bytes), which removes
// the dd round trip through shared memory, one warp barrier and the per-pixel index arithmetic (~150 of ~640 issue slots).
__global__ void __launch_bounds__(WARPS * 32) lk_fused_kernel(const __grid_constant__ LkArgs2 aa) {
  __shared__ WarpSmem sm[WARPS];
  const LkArgs& a = aa.a[blockIdx.z];
  const int s = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * WARPS + warp;
  if (i >= a.n_pts[s]) return;
  WarpSmem& w = sm[warp];
  const int slot = a.perm ? a.perm[(size_t)s * a.stride + i] : i;
  const float2 pA = a.ptsA[(size_t)s * a.stride + slot];
  float2 nxt;
  if (a.Hmat) {
    const float* H = a.Hmat + (size_t)s * 9;
    // cv::Matx33f * Vec3f : s = 0; s += H(r,k)*v(k)   (image_processor.cpp:285-290)
    float q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = __fmul_rn(H[r * 3 + 0], pA.x);
      acc = __fadd_rn(acc, __fmul_rn(H[r * 3 + 1], pA.y));
      acc = __fadd_rn(acc, H[r * 3 + 2]);
      q[r] = acc;
    }
    nxt.x = __fdiv_rn(q[0], q[2]);
    nxt.y = __fdiv_rn(q[1], q[2]);
  } else {
    nxt = a.init[(size_t)s * a.stride + (a.init_by_slot ? slot : i)];
  }
  bool status = true;
  const float halfWin = (WIN - 1) * 0.5f;

  for (int level = a.max_level; level >= 0; --level) {
    const LvbLevel lv = a.L.lv[level];
    const float scale = 1.0f / (float)(1 << level);
    float2 prevPt = make_float2(__fmul_rn(pA.x, scale), __fmul_rn(pA.y, scale));
    if (level == a.max_level) nxt = make_float2(__fmul_rn(nxt.x, scale), __fmul_rn(nxt.y, scale));
    else nxt = make_float2(__fmul_rn(nxt.x, 2.f), __fmul_rn(nxt.y, 2.f));
    prevPt.x = __fsub_rn(prevPt.x, halfWin);
    prevPt.y = __fsub_rn(prevPt.y, halfWin);
    const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
    if (ipx < -WIN || ipx >= lv.w || ipy < -WIN || ipy >= lv.h) {
      if (level == 0) status = false;
      continue;
    }
    const uint8_t* orgA = lvb_level_origin(a.pyrA, a.L, s, level);
    const uint8_t* orgB = lvb_level_origin(a.pyrB, a.L, s, level);
    __syncwarp();
    // ---- stage the 24x24 source tile (origin ipx-1, ipy-1)
    for (int t = lane; t < TILE * TILE; t += 32) {
      const int ty = t / TILE, tx = t - ty * TILE;
      w.u.st.tile[t] = __ldg(orgA + (ptrdiff_t)(ipy - 1 + ty) * lv.pitch + (ipx - 1 + tx));
    }
    __syncwarp();
    // ---- Scharr derivatives on the 22x22 grid (zero outside the image)
    for (int t = lane; t < DT * DT; t += 32) {
      const int dy = t / DT, dx = t - dy * DT;
      const int gx = ipx + dx, gy = ipy + dy;
      short2 d = make_short2(0, 0);
      if (gx >= 0 && gx < lv.w && gy >= 0 && gy < lv.h) {
        const uint8_t* c = &w.u.st.tile[(dy + 1) * TILE + (dx + 1)];
        const int tl = c[-TILE - 1], tc = c[-TILE], tr = c[-TILE + 1];
        const int ml = c[-1], mr = c[1];
        const int bl = c[TILE - 1], bc = c[TILE], br = c[TILE + 1];
        d.x = (short)(3 * (tr + br - tl - bl) + 10 * (mr - ml));
        d.y = (short)(3 * (bl + br - tl - tr) + 10 * (bc - tc));
      }
      w.u.st.dtile[t] = d;
    }
    __syncwarp();
    // ---- window of the previous image
    float fa = __fsub_rn(prevPt.x, (float)ipx), fb = __fsub_rn(prevPt.y, (float)ipy);
    int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    for (int p = lane; p < NPIX; p += 32) {
      const int y = p / WIN, x = p - y * WIN;
      const uint8_t* c = &w.u.st.tile[(y + 1) * TILE + (x + 1)];
      const int ival = (c[0] * iw00 + c[1] * iw01 + c[TILE] * iw10 + c[TILE + 1] * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
      const short2 d00 = w.u.st.dtile[y * DT + x], d01 = w.u.st.dtile[y * DT + x + 1];
      const short2 d10 = w.u.st.dtile[(y + 1) * DT + x], d11 = w.u.st.dtile[(y + 1) * DT + x + 1];
      const int ix = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
      const int iy = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
      w.Iw[p] = (short)ival;
      w.dIw[p] = make_short2((short)ix, (short)iy);
    }
    __syncwarp();
    // terms of A11, A12, A22 in OpenCV's lane order (the tile/dtile staging area is dead from here on)
    for (int p = lane; p < NPIX; p += 32) {
      const int y = p / WIN, x = p - y * WIN;
      const short2 d = w.dIw[p];
      const float fx = (float)d.x, fy = (float)d.y;
      const int slot = (x < 16) ? ((x & 3) * (WIN * 4) + y * 4 + (x >> 2)) : (WIN * 16 + y * 5 + (x - 16));
      w.u.termA[slot] = __fmul_rn(fx, fx);
      w.u.termA[NTERM_A + slot] = __fmul_rn(fx, fy);
      w.u.termA[2 * NTERM_A + slot] = __fmul_rn(fy, fy);
    }
    __syncwarp();
    float chainv = 0.f;
    if (lane < 15) chainv = run_chain<4>(w.u.termA + (lane / 5) * NTERM_A, lane % 5);
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    float A11 = __fmul_rn(combine_chains(chainv, 0), FLT_SCALE);
    float A12 = __fmul_rn(combine_chains(chainv, 5), FLT_SCALE);
    float A22 = __fmul_rn(combine_chains(chainv, 10), FLT_SCALE);
    __syncwarp();
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dif = __fsub_rn(A11, A22);
    const float rad = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
    if ((double)minEig < a.min_eig || D < 1.1920929e-07f) {
      if (level == 0) status = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);
    float2 np = make_float2(__fsub_rn(nxt.x, halfWin), __fsub_rn(nxt.y, halfWin));
    float2 prevDelta = make_float2(0.f, 0.f);
    for (int j = 0; j < a.max_iter; ++j) {
      const int inx = (int)floorf(np.x), iny = (int)floorf(np.y);
      if (inx < -WIN || inx >= lv.w || iny < -WIN || iny >= lv.h) {
        if (level == 0) status = false;
        break;
      }
      fa = __fsub_rn(np.x, (float)inx);
      fb = __fsub_rn(np.y, (float)iny);
      iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
      iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
      iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
      iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      const uint8_t* jbase = orgB + (ptrdiff_t)iny * lv.pitch + inx;
      // fused I_t + term pass: work item t < 168 is a pair term of OpenCV's b sums (pixels i0 and i0 + 4 of one 8-pixel chunk),
      // t >= 168 a scalar-tail term; the lane that owns the term gathers its own pixels, so I_t never goes through shared memory
      const int pitch = lv.pitch;
      for (int t = lane; t < WIN * 13; t += 32) {
        float t1, t2;
        if (t < WIN * 8) {
          const int y = t >> 3, ch = (t >> 2) & 1, p = t & 3;
          const int i0 = y * WIN + 8 * ch + p, i1 = i0 + 4;
          const uint8_t* c0 = jbase + y * pitch + 8 * ch + p;
          const int a00 = __ldg(c0), a01 = __ldg(c0 + 1), a10 = __ldg(c0 + pitch), a11 = __ldg(c0 + pitch + 1);
          const int b00 = __ldg(c0 + 4), b01 = __ldg(c0 + 5), b10 = __ldg(c0 + pitch + 4), b11 = __ldg(c0 + pitch + 5);
          const int d0 = (int)(short)(((a00 * iw00 + a01 * iw01 + a10 * iw10 + a11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - (int)w.Iw[i0]);
          const int d1 = (int)(short)(((b00 * iw00 + b01 * iw01 + b10 * iw10 + b11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - (int)w.Iw[i1]);
          const short2 g0 = w.dIw[i0], g1 = w.dIw[i1];
          t1 = (float)(d0 * g0.x + d1 * g1.x);
          t2 = (float)(d0 * g0.y + d1 * g1.y);
          const int slot = p * (WIN * 2) + y * 2 + ch;
          w.u.termB[slot] = t1;
          w.u.termB[NTERM_B + slot] = t2;
        } else {
          const int q = t - WIN * 8, y = q / 5, x = 16 + q - y * 5;
          const int i0 = y * WIN + x;
          const uint8_t* c0 = jbase + y * pitch + x;
          const int a00 = __ldg(c0), a01 = __ldg(c0 + 1), a10 = __ldg(c0 + pitch), a11 = __ldg(c0 + pitch + 1);
          const int d0 = (int)(short)(((a00 * iw00 + a01 * iw01 + a10 * iw10 + a11 * iw11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - (int)w.Iw[i0]);
          const short2 g0 = w.dIw[i0];
          t1 = (float)(d0 * g0.x);
          t2 = (float)(d0 * g0.y);
          w.u.termB[t] = t1;
          w.u.termB[NTERM_B + t] = t2;
        }
      }
      __syncwarp();
      float cv = 0.f;
      if (lane < 10) cv = run_chain<2>(w.u.termB + (lane / 5) * NTERM_B, lane % 5);
      const float b1 = __fmul_rn(combine_chains(cv, 0), FLT_SCALE);
      const float b2 = __fmul_rn(combine_chains(cv, 5), FLT_SCALE);
      __syncwarp();
      float2 delta;
      delta.x = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      delta.y = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      np.x = __fadd_rn(np.x, delta.x);
      np.y = __fadd_rn(np.y, delta.y);
      nxt = make_float2(__fadd_rn(np.x, halfWin), __fadd_rn(np.y, halfWin));
      const double dd = (double)delta.x * (double)delta.x + (double)delta.y * (double)delta.y;
      if (dd <= a.eps2) break;
      if (j > 0 && (double)fabsf(__fadd_rn(delta.x, prevDelta.x)) < 0.01 && (double)fabsf(__fadd_rn(delta.y, prevDelta.y)) < 0.01) {
        nxt.x = __fsub_rn(nxt.x, __fmul_rn(delta.x, 0.5f));
        nxt.y = __fsub_rn(nxt.y, __fmul_rn(delta.y, 0.5f));
        break;
      }
      prevDelta = delta;
    }
  }

  if (lane == 0) {
    const LvbLevel l0 = a.L.lv[0];
    if (status && a.gate_mode >= 1) {
      if (nxt.y < 0.f || nxt.y > (float)(l0.h - 1) || nxt.x < 0.f || nxt.x > (float)(l0.w - 1)) status = false;
    }
    if (status && a.gate_mode == 2) {
      const float2 r = a.ref[(size_t)s * a.stride + slot];
      const float dx = __fsub_rn(nxt.x, r.x), dy = __fsub_rn(nxt.y, r.y);
      const float dis = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
      if (dis > 1.f) status = false;
    }
    a.out[(size_t)s * a.stride + i] = nxt;
    a.status[(size_t)s * a.stride + i] = status ? 1 : 0;
  }
}

}  // namespace

static void fill_lk_args(LvbHandle* h, LkArgs& a, const uint8_t* pyrA, const uint8_t* pyrB, int stride, const float2* ptsA,
                         const int* perm, const int* n_pts, const float2* init, int init_by_slot, const float* Hmat, float2* out,
                         uint8_t* status, int gate_mode, const float2* ref) {
  a.pyrA = pyrA; a.pyrB = pyrB; a.L = h->fe.L; a.stride = stride; a.ptsA = ptsA; a.perm = perm;
  a.n_pts = n_pts; a.init = init; a.init_by_slot = init_by_slot; a.Hmat = Hmat; a.out = out; a.status = status;
  a.gate_mode = gate_mode; a.ref = ref;
  int mi = h->cfg.max_iteration; if (mi < 0) mi = 0; if (mi > 100) mi = 100;      // cv clamps maxCount to [0,100]
  double eps = h->cfg.track_precision; if (eps < 0) eps = 0; if (eps > 10) eps = 10;
  a.max_iter = mi; a.eps2 = eps * eps; a.min_eig = 1e-4;
  a.max_level = h->cfg.pyramid_levels;
}

int fe_lk_launch(LvbHandle* h, const uint8_t* pyrA, const uint8_t* pyrB, int n_seq, int stride,
                 const float2* ptsA, const int* perm, const int* n_pts, const float2* init, int init_by_slot,
                 const float* Hmat, float2* out, uint8_t* status, int gate_mode, const float2* ref) {
  if (h->cfg.patch_size != WIN) return lvb_set_err(LVB_E_UNSUPPORTED, "patch_size %d (kernel is built for 21)", h->cfg.patch_size);
  LkArgs2 aa;
  fill_lk_args(h, aa.a[0], pyrA, pyrB, stride, ptsA, perm, n_pts, init, init_by_slot, Hmat, out, status, gate_mode, ref);
  aa.a[1] = aa.a[0];
  dim3 grd((stride + WARPS - 1) / WARPS, n_seq, 1);
  if (h->experiments & LVB_EXP_LK_FUSED) {
    LVB_PROF(h, "lk_fused_kernel");
    lk_fused_kernel<<<grd, WARPS * 32, 0, h->stream>>>(aa);
    LVB_LAUNCH_CHECK(h);
    return LVB_OK;
  }
  LVB_PROF(h, "lk_kernel");
  lk_kernel<<<grd, WARPS * 32, 0, h->stream>>>(aa);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}

// both chains (tracked + new) of one direction in a single launch (gridDim.z = 2)
int fe_lk_launch2(LvbHandle* h, const uint8_t* pyrA, const uint8_t* pyrB, int n_seq, int stride,
                  const float2* const ptsA[2], int* const perm[2], int* const n_pts[2], const float2* const init[2],
                  int init_by_slot, const float* Hmat, float2* const out[2], uint8_t* const status[2], int gate_mode,
                  const float2* const ref[2]) {
  LkArgs2 aa;
  for (int c = 0; c < 2; ++c)
    fill_lk_args(h, aa.a[c], pyrA, pyrB, stride, ptsA[c], perm[c], n_pts[c], init ? init[c] : nullptr, init_by_slot, Hmat, out[c],
                 status[c], gate_mode, ref ? ref[c] : nullptr);
  dim3 grd((stride + WARPS - 1) / WARPS, n_seq, 2);
  if (h->experiments & LVB_EXP_LK_FUSED) {
    LVB_PROF(h, "lk_fused_kernel");
    lk_fused_kernel<<<grd, WARPS * 32, 0, h->stream>>>(aa);
    LVB_LAUNCH_CHECK(h);
    return LVB_OK;
  }
  LVB_PROF(h, "lk_kernel");
  lk_kernel<<<grd, WARPS * 32, 0, h->stream>>>(aa);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}
