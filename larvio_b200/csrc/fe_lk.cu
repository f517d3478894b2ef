// Pyramidal iterative Lucas-Kanade with initial flow, one warp per point, all levels in one launch.
// Replaces the six cv::calcOpticalFlowPyrLK call sites of image_processor.cpp (:368-377, :405-414,
// :558-567, :618-627, :830-839, :870-879; win 21x21, 3 levels, <=30 iterations, eps 0.01,
// minEigThreshold 1e-4, OPTFLOW_USE_INITIAL_FLOW) together with the gates that follow each call:
// predictFeatureTracking (:266-293) as prologue, the in-image test (:571-578) and the
// forward/backward consistency test (:630-642) as epilogue.  Arithmetic follows SURVEY.md App. A.5:
// Q14 bilinear weights, int16 patches with 5 fractional bits, float32 normal equations.
// Scharr derivatives are computed on the fly from the padded level image (zero outside the image,
// REFLECT_101 at its edge == the pad content), so no derivative pyramid exists in HBM.
//
// Data movement (sm_100a): per pyramid level a warp stages two 32x32-byte tiles with TMA
// (cp.async.bulk.tensor.3d + a per-warp mbarrier): the source window of the previous image and a
// search tile of the next image centred on the predicted position.  All <=30 iterations of the level run out of
// that shared-memory tile (re-staged only if the window walks more than 5 px away), so the dependent
// iteration chain never waits on global memory.  Each lane keeps its 16 window pixels (patch value,
// both derivatives) in registers for the whole level.
//
// Bit-exactness with OpenCV's SSE accumulation order (oracle/lk_exact.py): the sums of the normal
// equations are sums of integer-valued floats, accumulated by OpenCV in 4 SSE lanes + a scalar tail,
// i.e. five sequential float chains per sum.  A float chain whose partial sums all stay below 2^24 is exact,
// hence equal to the integer sum in ANY order; we prove that per chain with a Cauchy-Schwarz bound
// (sum|d g| <= sqrt(sum d^2 * sum g^2) < 2^24) from integer warp reductions and then take the
// integer sums (fast path).  When the bound fails the terms are written out and the five chains are replayed
// in OpenCV's order, one lane per chain (slow path) - the result is bit-identical either way.
#include <cuda.h>
#include <string.h>
#include "lvb_internal.h"

namespace {

constexpr int WIN = 21;
constexpr int DT = WIN + 1;                // 22: derivative grid
constexpr int WARPS = 4;
constexpr int W_BITS = 14;
constexpr int TB = 32;                     // staged tiles are TB x TB bytes (TMA box), row pitch TB
constexpr int MARGIN = 5;                  // the search window may drift +-MARGIN px inside the staged tile

// OpenCV's SSE path accumulates the normal equations in 4 float lanes + a scalar tail (pixels 16..20 of each
// window row); per sum, accumulator p < 4 receives, row by row, the terms of pixels p, 4+p, 8+p, 12+p (A) resp.
// the int-pair terms of both 8-pixel chunks (b); accumulator 4 receives pixels 16..20.
constexpr int NTERM_A = WIN * 16 + WIN * 5 + 3;   // per sum: 336 lane terms + 105 tail terms (+3 pad -> 16-B aligned sums)
constexpr int NTERM_B = WIN * 8 + WIN * 5 + 3;    // per sum: 168 pair terms + 105 tail terms (+3 pad)
struct alignas(128) WarpSmem {
  uint8_t tileB[TB * TB + 128];            // search tile of the next image (+ slack: masked lanes read a few bytes past the end)
  union {
    struct { uint8_t tileA[TB * TB]; int2 dgrid[DT * DT + 4]; } st;   // window set-up: source tile, Scharr (dx, dy) on the 22x22 grid
    float termA[3 * NTERM_A];              // slow path of the A sums
    float termB[2 * NTERM_B];              // slow path of the b sums
  } u;
  unsigned long long bar[2];               // [0] source tile, [1] search tile
};

// Term layout per sum: accumulator p (0..3) owns WIN*K consecutive floats (row-major over (row, k)), followed by the
// WIN*5 tail terms; K = 4 for the A sums, 2 for the b sums.  Each chain adds its terms strictly in order.
template <int K>
__device__ __forceinline__ float run_chain(const float* T, int acc) {
  float a = 0.f;
  if (acc < 4) {
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
__device__ __forceinline__ float combine5(float l0, float l1, float l2, float l3, float tl) {
  return __fadd_rn(tl, __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3)));
}
__device__ __forceinline__ float combine_chains(float mine, int base) {
  const float l0 = __shfl_sync(0xffffffffu, mine, base), l1 = __shfl_sync(0xffffffffu, mine, base + 1);
  const float l2 = __shfl_sync(0xffffffffu, mine, base + 2), l3 = __shfl_sync(0xffffffffu, mine, base + 3);
  const float tl = __shfl_sync(0xffffffffu, mine, base + 4);
  return combine5(l0, l1, l2, l3, tl);
}

// ---- sm_100a primitives
__device__ __forceinline__ int dp2a_lo(int w, unsigned b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi(int w, unsigned b, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tma_load_tile(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(TB * TB) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// Eight horizontally adjacent pixels of one tile row pair: bytes [b, b+9) of row `rowp` (top) and of the row below,
// bilinear weights wt = iw00 | iw01 << 16, wb = iw10 | iw11 << 16, per-pixel accumulator start c[j].
// out[j] = c[j] + iw00*T[j] + iw01*T[j+1] + iw10*B[j] + iw11*B[j+1]            (two IDP.2A per pixel)
__device__ __forceinline__ void bilinear8(const uint8_t* rowp, int b, int wt, int wb, const int* c, int* out) {
  const uint8_t* q = rowp + (b & ~7);
  const uint2 t0 = *reinterpret_cast<const uint2*>(q), t1 = *reinterpret_cast<const uint2*>(q + 8);
  const uint2 u0 = *reinterpret_cast<const uint2*>(q + TB), u1 = *reinterpret_cast<const uint2*>(q + TB + 8);
  const bool hi = (b & 4) != 0;
  const unsigned tw0 = hi ? t0.y : t0.x, tw1 = hi ? t1.x : t0.y, tw2 = hi ? t1.y : t1.x;
  const unsigned bw0 = hi ? u0.y : u0.x, bw1 = hi ? u1.x : u0.y, bw2 = hi ? u1.y : u1.x;
  const int k = (b & 3) * 8;
  const unsigned ta0 = __funnelshift_rc(tw0, tw1, k), ta1 = __funnelshift_rc(tw1, tw2, k);          // bytes 0-3, 4-7
  const unsigned ts0 = __funnelshift_rc(tw0, tw1, k + 8), ts1 = __funnelshift_rc(tw1, tw2, k + 8);  // bytes 1-4, 5-8
  const unsigned ba0 = __funnelshift_rc(bw0, bw1, k), ba1 = __funnelshift_rc(bw1, bw2, k);
  const unsigned bs0 = __funnelshift_rc(bw0, bw1, k + 8), bs1 = __funnelshift_rc(bw1, bw2, k + 8);
  out[0] = dp2a_lo(wt, ta0, dp2a_lo(wb, ba0, c[0]));
  out[1] = dp2a_lo(wt, ts0, dp2a_lo(wb, bs0, c[1]));
  out[2] = dp2a_hi(wt, ta0, dp2a_hi(wb, ba0, c[2]));
  out[3] = dp2a_hi(wt, ts0, dp2a_hi(wb, bs0, c[3]));
  out[4] = dp2a_lo(wt, ta1, dp2a_lo(wb, ba1, c[4]));
  out[5] = dp2a_lo(wt, ts1, dp2a_lo(wb, bs1, c[5]));
  out[6] = dp2a_hi(wt, ta1, dp2a_hi(wb, ba1, c[6]));
  out[7] = dp2a_hi(wt, ts1, dp2a_hi(wb, bs1, c[7]));
}

struct LkArgs {
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
};

struct LkArgs2 {
  CUtensorMap mapA[LVB_MAX_LEVELS];   // padded level images of the previous / next pyramid: (x, y, sequence), box TB x TB x 1
  CUtensorMap mapB[LVB_MAX_LEVELS];
  LkArgs a[2];
  int lw[LVB_MAX_LEVELS], lh[LVB_MAX_LEVELS];
  int max_iter; double eps2; double min_eig;
  int max_level;
  unsigned long long* stats;  // [10] iterations, [11] slow-path iterations, [12] slow-path window set-ups, [13] tile re-stages
};

__global__ void __launch_bounds__(WARPS * 32, 4) lk_kernel(const __grid_constant__ LkArgs2 aa) {
  __shared__ WarpSmem sm[WARPS];
  const LkArgs& a = aa.a[blockIdx.z];
  const int s = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * WARPS + warp;
  if (i >= a.n_pts[s]) return;
  WarpSmem& w = sm[warp];
  const unsigned bar0 = smem_u32(&w.bar[0]), bar1 = smem_u32(&w.bar[1]);
  if (lane == 0) {
    mbar_init(bar0); mbar_init(bar1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  unsigned ph0 = 0, ph1 = 0;

  // ---- lane roles: 21 "row" lanes own the 16 SSE pixels of one window row (two 8-pixel items), 11 "tail" lanes own
  // the scalar-tail pixels 16..20 of two rows each.  Row lanes are interleaved with tail lanes and alternate the item
  // order every 4 rows so that the 8-byte shared-memory reads of a half-warp spread over the banks.
  int row_l; bool tail;
  if (lane < 8) { row_l = lane; tail = false; }
  else if (lane < 12) { row_l = lane - 8; tail = true; }
  else if (lane < 25) { row_l = lane - 4; tail = false; }
  else { row_l = lane - 21; tail = true; }
  const int cls = (row_l >> 2) & 1;
  const int rowA = row_l, rowB = tail ? min(row_l + 11, WIN - 1) : row_l;
  const int xA = tail ? 16 : 8 * cls, xB = tail ? 16 : 8 * (1 - cls);
  const int nvA = tail ? 5 : 8, nvB = tail ? ((row_l + 11 < WIN) ? 5 : 0) : 8;   // valid pixels per item
  const unsigned rm = tail ? 0u : 0xffffffffu;

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
  const float FLT_SCALE = 1.f / (float)(1 << 20);
  unsigned n_iter = 0, n_slow = 0, n_slowA = 0, n_restage = 0;

  for (int level = aa.max_level; level >= 0; --level) {
    const int lw = aa.lw[level], lh = aa.lh[level];
    const float scale = 1.0f / (float)(1 << level);
    float2 prevPt = make_float2(__fmul_rn(pA.x, scale), __fmul_rn(pA.y, scale));
    if (level == aa.max_level) nxt = make_float2(__fmul_rn(nxt.x, scale), __fmul_rn(nxt.y, scale));
    else nxt = make_float2(__fmul_rn(nxt.x, 2.f), __fmul_rn(nxt.y, 2.f));
    prevPt.x = __fsub_rn(prevPt.x, halfWin);
    prevPt.y = __fsub_rn(prevPt.y, halfWin);
    const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
    if (ipx < -WIN || ipx >= lw || ipy < -WIN || ipy >= lh) {
      if (level == 0) status = false;
      continue;
    }
    float2 np = make_float2(__fsub_rn(nxt.x, halfWin), __fsub_rn(nxt.y, halfWin));
    // ---- stage the source tile (origin ipx-1, ipy-1: window + bilinear +1 + Scharr apron) and the search tile
    int tx, ty;
    {
      // float -> int with saturation (np may be far outside for a lost point); the iteration re-checks the range anyway
      const float cx = fminf(fmaxf(floorf(np.x), -4096.f), 8192.f), cy = fminf(fmaxf(floorf(np.y), -4096.f), 8192.f);
      tx = (int)cx - MARGIN; ty = (int)cy - MARGIN;
    }
    __syncwarp();
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic writes to the union (term scratch)
      tma_load_tile(smem_u32(w.u.st.tileA), &aa.mapA[level], ipx - 1 + LVB_PAD, ipy - 1 + LVB_PAD, s, bar0);
      tma_load_tile(smem_u32(w.tileB), &aa.mapB[level], tx + LVB_PAD, ty + LVB_PAD, s, bar1);
    }
    mbar_wait(bar0, ph0); ph0 ^= 1;
    // ---- Scharr derivatives on the 22x22 grid (zero outside the image)
    {
      int dy = 0, dx = lane;                                             // grid point t = lane + 32 k  ->  (dy, dx)
      if (dx >= DT) { dx -= DT; dy = 1; }
      for (int t = lane; t < DT * DT; t += 32) {
        const int gx = ipx + dx, gy = ipy + dy;
        int2 d = make_int2(0, 0);
        if (gx >= 0 && gx < lw && gy >= 0 && gy < lh) {
          const uint8_t* c = &w.u.st.tileA[(dy + 1) * TB + (dx + 1)];
          const int tl = c[-TB - 1], tc = c[-TB], tr = c[-TB + 1];
          const int ml = c[-1], mr = c[1];
          const int bl = c[TB - 1], bc = c[TB], br = c[TB + 1];
          d.x = 3 * (tr + br - tl - bl) + 10 * (mr - ml);
          d.y = 3 * (bl + br - tl - tr) + 10 * (bc - tc);
        }
        w.u.st.dgrid[t] = d;
        dx += 32 - DT; dy += 1;                                          // 32 = 22 + 10
        if (dx >= DT) { dx -= DT; dy += 1; }
      }
    }
    __syncwarp();
    // ---- this lane's 16 pixels of the previous-image window: patch value folded into the accumulator start
    //      cI = 256 - (Iw << 9)  (so that I_t = (cI + sum w*J) >> 9 in the iterations), derivatives gX, gY
    float fa = __fsub_rn(prevPt.x, (float)ipx), fb = __fsub_rn(prevPt.y, (float)ipy);
    int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    int cI[16], gX[16], gY[16];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int y = it ? rowB : rowA, x0 = it ? xB : xA, nv = it ? nvB : nvA;
      int c256[8], iv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c256[j] = 1 << (W_BITS - 5 - 1);
      bilinear8(&w.u.st.tileA[(y + 1) * TB], x0 + 1, (iw00 & 0xffff) | (iw01 << 16), (iw10 & 0xffff) | (iw11 << 16), c256, iv);
      const int2* g0 = &w.u.st.dgrid[y * DT + x0];
      const int2* g1 = g0 + DT;
      int2 p0 = g0[0], p1 = g1[0];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int2 q0 = g0[j + 1], q1 = g1[j + 1];
        const int ix = (p0.x * iw00 + q0.x * iw01 + p1.x * iw10 + q1.x * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
        const int iy = (p0.y * iw00 + q0.y * iw01 + p1.y * iw10 + q1.y * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
        const bool ok = j < nv;
        cI[it * 8 + j] = ok ? (1 << (W_BITS - 5 - 1)) - ((iv[j] >> (W_BITS - 5)) << (W_BITS - 5)) : 0;
        gX[it * 8 + j] = ok ? ix : 0;
        gY[it * 8 + j] = ok ? iy : 0;
        p0 = q0; p1 = q1;
      }
    }
    // ---- A sums.  Per chain c (4 SSE accumulators + tail): G11 = sum gx^2, G22 = sum gy^2, G12 = sum gx gy, Gab = sum |gx gy|
    unsigned G11[5], G22[5], Gab[5]; int G12[5];
    {
      unsigned g11[4] = {0, 0, 0, 0}, g22[4] = {0, 0, 0, 0}, gab[4] = {0, 0, 0, 0};
      int g12[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int xy = gX[j] * gY[j];
        g11[j & 3] += (unsigned)(gX[j] * gX[j]); g22[j & 3] += (unsigned)(gY[j] * gY[j]);
        g12[j & 3] += xy; gab[j & 3] += (unsigned)abs(xy);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        G11[p] = __reduce_add_sync(0xffffffffu, g11[p] & rm); G22[p] = __reduce_add_sync(0xffffffffu, g22[p] & rm);
        G12[p] = __reduce_add_sync(0xffffffffu, (int)((unsigned)g12[p] & rm)); Gab[p] = __reduce_add_sync(0xffffffffu, gab[p] & rm);
      }
      G11[4] = __reduce_add_sync(0xffffffffu, (g11[0] + g11[1] + g11[2] + g11[3]) & ~rm);
      G22[4] = __reduce_add_sync(0xffffffffu, (g22[0] + g22[1] + g22[2] + g22[3]) & ~rm);
      G12[4] = __reduce_add_sync(0xffffffffu, (int)((unsigned)(g12[0] + g12[1] + g12[2] + g12[3]) & ~rm));
      Gab[4] = __reduce_add_sync(0xffffffffu, (gab[0] + gab[1] + gab[2] + gab[3]) & ~rm);
    }
    unsigned gmaxA = 0;
#pragma unroll
    for (int c = 0; c < 5; ++c) gmaxA = max(gmaxA, max(max(G11[c], G22[c]), Gab[c]));
    float A11, A12, A22;
    if (gmaxA < (1u << 24)) {             // every chain exact: the float chains equal the integer sums
      A11 = __fmul_rn(combine5((float)G11[0], (float)G11[1], (float)G11[2], (float)G11[3], (float)G11[4]), FLT_SCALE);
      A12 = __fmul_rn(combine5((float)G12[0], (float)G12[1], (float)G12[2], (float)G12[3], (float)G12[4]), FLT_SCALE);
      A22 = __fmul_rn(combine5((float)G22[0], (float)G22[1], (float)G22[2], (float)G22[3], (float)G22[4]), FLT_SCALE);
    } else {                               // replay OpenCV's chains (the staging area is dead from here on)
      ++n_slowA;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int it = j >> 3, jj = j & 7;
        const int y = it ? rowB : rowA, x = (it ? xB : xA) + jj, nv = it ? nvB : nvA;
        if (jj < nv) {
          const float fx = (float)gX[j], fy = (float)gY[j];
          const int sl = (x < 16) ? ((x & 3) * (WIN * 4) + y * 4 + (x >> 2)) : (WIN * 16 + y * 5 + (x - 16));
          w.u.termA[sl] = __fmul_rn(fx, fx);
          w.u.termA[NTERM_A + sl] = __fmul_rn(fx, fy);
          w.u.termA[2 * NTERM_A + sl] = __fmul_rn(fy, fy);
        }
      }
      __syncwarp();
      float chainv = 0.f;
      if (lane < 15) chainv = run_chain<4>(w.u.termA + (lane / 5) * NTERM_A, lane % 5);
      A11 = __fmul_rn(combine_chains(chainv, 0), FLT_SCALE);
      A12 = __fmul_rn(combine_chains(chainv, 5), FLT_SCALE);
      A22 = __fmul_rn(combine_chains(chainv, 10), FLT_SCALE);
      __syncwarp();
    }
    // per-chain bound for the b sums: exact while sum d^2 < thr[c] <= 2^48 / max(G11, G22)   (Cauchy-Schwarz)
    unsigned thr[5];
#pragma unroll
    for (int c = 0; c < 5; ++c)
      thr[c] = __float2uint_rd(__fdiv_rd(281474976710656.f, __uint2float_ru(max(G11[c], G22[c]))));

    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dif = __fsub_rn(A11, A22);
    const float rad = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
    mbar_wait(bar1, ph1); ph1 ^= 1;                                      // search tile has landed (needed before any `continue`)
    if ((double)minEig < aa.min_eig || D < 1.1920929e-07f) {
      if (level == 0) status = false;
      continue;
    }
    D = __fdiv_rn(1.f, D);
    float2 prevDelta = make_float2(0.f, 0.f);
    for (int j = 0; j < aa.max_iter; ++j) {
      const int inx = (int)floorf(np.x), iny = (int)floorf(np.y);
      if (inx < -WIN || inx >= lw || iny < -WIN || iny >= lh) {
        if (level == 0) status = false;
        break;
      }
      int ox = inx - tx, oy = iny - ty;
      if (ox < 0 || ox > 2 * MARGIN || oy < 0 || oy > 2 * MARGIN) {    // the window left the staged tile: re-centre it
        tx = inx - MARGIN; ty = iny - MARGIN; ox = MARGIN; oy = MARGIN;
        ++n_restage;
        __syncwarp();
        if (lane == 0) tma_load_tile(smem_u32(w.tileB), &aa.mapB[level], tx + LVB_PAD, ty + LVB_PAD, s, bar1);
        mbar_wait(bar1, ph1); ph1 ^= 1;
      }
      fa = __fsub_rn(np.x, (float)inx);
      fb = __fsub_rn(np.y, (float)iny);
      iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
      iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
      iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
      iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
      const int wt = (iw00 & 0xffff) | (iw01 << 16), wb = (iw10 & 0xffff) | (iw11 << 16);
      // I_t of this lane's 16 pixels: d = (cI + bilinear) >> 9
      int d[16];
      bilinear8(&w.tileB[(oy + rowA) * TB], ox + xA, wt, wb, &cI[0], &d[0]);
      bilinear8(&w.tileB[(oy + rowB) * TB], ox + xB, wt, wb, &cI[8], &d[8]);
      int s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
      unsigned dq[4] = {0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        d[q] >>= (W_BITS - 5);
        s1[q & 3] += d[q] * gX[q]; s2[q & 3] += d[q] * gY[q];
        dq[q & 3] += (unsigned)(d[q] * d[q]);
      }
      ++n_iter;
      // integer warp reductions per chain
      const unsigned dmax = __reduce_max_sync(0xffffffffu, max(max(dq[0], dq[1]), max(dq[2], dq[3])));
      int S1[5], S2[5]; unsigned DQ[5];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        S1[p] = __reduce_add_sync(0xffffffffu, (int)((unsigned)s1[p] & rm));
        S2[p] = __reduce_add_sync(0xffffffffu, (int)((unsigned)s2[p] & rm));
        DQ[p] = __reduce_add_sync(0xffffffffu, dq[p] & rm);
      }
      S1[4] = __reduce_add_sync(0xffffffffu, (int)((unsigned)(s1[0] + s1[1] + s1[2] + s1[3]) & ~rm));
      S2[4] = __reduce_add_sync(0xffffffffu, (int)((unsigned)(s2[0] + s2[1] + s2[2] + s2[3]) & ~rm));
      DQ[4] = __reduce_add_sync(0xffffffffu, (dq[0] + dq[1] + dq[2] + dq[3]) & ~rm);
      bool exact = dmax < (1u << 26);                                    // no 32-bit wrap in the DQ reductions
#pragma unroll
      for (int c = 0; c < 5; ++c) exact = exact && (DQ[c] < thr[c]);
      float b1, b2;
      if (exact) {
        b1 = __fmul_rn(combine5((float)S1[0], (float)S1[1], (float)S1[2], (float)S1[3], (float)S1[4]), FLT_SCALE);
        b2 = __fmul_rn(combine5((float)S2[0], (float)S2[1], (float)S2[2], (float)S2[3], (float)S2[4]), FLT_SCALE);
      } else {
        ++n_slow;
        __syncwarp();
        if (!tail) {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int ch = (it ? xB : xA) >> 3;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const int q0 = it * 8 + p, q1 = q0 + 4;
              const int sl = p * (WIN * 2) + rowA * 2 + ch;
              w.u.termB[sl] = (float)(d[q0] * gX[q0] + d[q1] * gX[q1]);
              w.u.termB[NTERM_B + sl] = (float)(d[q0] * gY[q0] + d[q1] * gY[q1]);
            }
          }
        } else {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int y = it ? rowB : rowA, nv = it ? nvB : nvA;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              if (q < nv) {
                const int sl = WIN * 8 + y * 5 + q;
                w.u.termB[sl] = (float)(d[it * 8 + q] * gX[it * 8 + q]);
                w.u.termB[NTERM_B + sl] = (float)(d[it * 8 + q] * gY[it * 8 + q]);
              }
            }
          }
        }
        __syncwarp();
        float cv = 0.f;
        if (lane < 10) cv = run_chain<2>(w.u.termB + (lane / 5) * NTERM_B, lane % 5);
        b1 = __fmul_rn(combine_chains(cv, 0), FLT_SCALE);
        b2 = __fmul_rn(combine_chains(cv, 5), FLT_SCALE);
        __syncwarp();
      }
      float2 delta;
      delta.x = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      delta.y = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      np.x = __fadd_rn(np.x, delta.x);
      np.y = __fadd_rn(np.y, delta.y);
      nxt = make_float2(__fadd_rn(np.x, halfWin), __fadd_rn(np.y, halfWin));
      const double dd = (double)delta.x * (double)delta.x + (double)delta.y * (double)delta.y;
      if (dd <= aa.eps2) break;
      if (j > 0 && (double)fabsf(__fadd_rn(delta.x, prevDelta.x)) < 0.01 && (double)fabsf(__fadd_rn(delta.y, prevDelta.y)) < 0.01) {
        nxt.x = __fsub_rn(nxt.x, __fmul_rn(delta.x, 0.5f));
        nxt.y = __fsub_rn(nxt.y, __fmul_rn(delta.y, 0.5f));
        break;
      }
      prevDelta = delta;
    }
  }

  if (lane == 0) {
    if (status && a.gate_mode >= 1) {
      if (nxt.y < 0.f || nxt.y > (float)(aa.lh[0] - 1) || nxt.x < 0.f || nxt.x > (float)(aa.lw[0] - 1)) status = false;
    }
    if (status && a.gate_mode == 2) {
      const float2 r = a.ref[(size_t)s * a.stride + slot];
      const float dx = __fsub_rn(nxt.x, r.x), dy = __fsub_rn(nxt.y, r.y);
      const float dis = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
      if (dis > 1.f) status = false;
    }
    a.out[(size_t)s * a.stride + i] = nxt;
    a.status[(size_t)s * a.stride + i] = status ? 1 : 0;
    if (aa.stats) {
      atomicAdd(&aa.stats[10], (unsigned long long)n_iter);
      if (n_slow) atomicAdd(&aa.stats[11], (unsigned long long)n_slow);
      if (n_slowA) atomicAdd(&aa.stats[12], (unsigned long long)n_slowA);
      if (n_restage) atomicAdd(&aa.stats[13], (unsigned long long)n_restage);
    }
  }
}

}  // namespace

// ---- TMA descriptors of a padded pyramid block: one 3-D u8 tensor (x, y, sequence) per level, box TB x TB x 1
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

static int make_level_maps(LvbHandle* h, const uint8_t* pyr, int n_seq, CUtensorMap* out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return lvb_set_err(LVB_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  const LvbPyramidLayout& L = h->fe.L;
  for (int l = 0; l < L.n_levels; ++l) {
    const LvbLevel& lv = L.lv[l];
    const cuuint64_t dims[3] = {(cuuint64_t)lv.pitch, (cuuint64_t)lv.rows, (cuuint64_t)n_seq};
    const cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)L.bytes_per_seq};
    const cuuint32_t box[3] = {(cuuint32_t)TB, (cuuint32_t)TB, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(&out[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)(pyr + lv.offset), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return lvb_set_err(LVB_E_CUDA, "cuTensorMapEncodeTiled(level %d) -> %d", l, (int)r);
  }
  return LVB_OK;
}

// the handle's own ping-pong pyramids keep their descriptors; temporary pyramids (stage-level entry points) get fresh ones
static int level_maps(LvbHandle* h, const uint8_t* pyr, int n_seq, CUtensorMap* out) {
  for (int k = 0; k < 2; ++k) {
    if (pyr == h->fe.pyr[k] && n_seq == h->fe.S) {
      if (!h->lk_maps_ok[k]) {
        int rc = make_level_maps(h, pyr, n_seq, h->lk_maps[k]);
        if (rc != LVB_OK) return rc;
        h->lk_maps_ok[k] = true;
      }
      memcpy(out, h->lk_maps[k], sizeof(CUtensorMap) * LVB_MAX_LEVELS);
      return LVB_OK;
    }
  }
  return make_level_maps(h, pyr, n_seq, out);
}

static void fill_lk_args(LkArgs& a, int stride, const float2* ptsA, const int* perm, const int* n_pts, const float2* init,
                         int init_by_slot, const float* Hmat, float2* out, uint8_t* status, int gate_mode, const float2* ref) {
  a.stride = stride; a.ptsA = ptsA; a.perm = perm;
  a.n_pts = n_pts; a.init = init; a.init_by_slot = init_by_slot; a.Hmat = Hmat; a.out = out; a.status = status;
  a.gate_mode = gate_mode; a.ref = ref;
}

static int fill_lk_common(LvbHandle* h, LkArgs2& aa, const uint8_t* pyrA, const uint8_t* pyrB, int n_seq) {
  if (h->cfg.patch_size != WIN) return lvb_set_err(LVB_E_UNSUPPORTED, "patch_size %d (kernel is built for 21)", h->cfg.patch_size);
  memset(&aa, 0, sizeof(aa));
  int rc = level_maps(h, pyrA, n_seq, aa.mapA);
  if (rc != LVB_OK) return rc;
  rc = level_maps(h, pyrB, n_seq, aa.mapB);
  if (rc != LVB_OK) return rc;
  for (int l = 0; l < h->fe.L.n_levels; ++l) { aa.lw[l] = h->fe.L.lv[l].w; aa.lh[l] = h->fe.L.lv[l].h; }
  int mi = h->cfg.max_iteration; if (mi < 0) mi = 0; if (mi > 100) mi = 100;      // cv clamps maxCount to [0,100]
  double eps = h->cfg.track_precision; if (eps < 0) eps = 0; if (eps > 10) eps = 10;
  aa.max_iter = mi; aa.eps2 = eps * eps; aa.min_eig = 1e-4;
  aa.max_level = h->cfg.pyramid_levels;
  aa.stats = h->fe.stats;
  return LVB_OK;
}

int fe_lk_launch(LvbHandle* h, const uint8_t* pyrA, const uint8_t* pyrB, int n_seq, int stride,
                 const float2* ptsA, const int* perm, const int* n_pts, const float2* init, int init_by_slot,
                 const float* Hmat, float2* out, uint8_t* status, int gate_mode, const float2* ref) {
  LkArgs2 aa;
  int rc = fill_lk_common(h, aa, pyrA, pyrB, n_seq);
  if (rc != LVB_OK) return rc;
  fill_lk_args(aa.a[0], stride, ptsA, perm, n_pts, init, init_by_slot, Hmat, out, status, gate_mode, ref);
  aa.a[1] = aa.a[0];
  dim3 grd((stride + WARPS - 1) / WARPS, n_seq, 1);
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
  int rc = fill_lk_common(h, aa, pyrA, pyrB, n_seq);
  if (rc != LVB_OK) return rc;
  for (int c = 0; c < 2; ++c)
    fill_lk_args(aa.a[c], stride, ptsA[c], perm[c], n_pts[c], init ? init[c] : nullptr, init_by_slot, Hmat, out[c],
                 status[c], gate_mode, ref ? ref[c] : nullptr);
  dim3 grd((stride + WARPS - 1) / WARPS, n_seq, 2);
  LVB_PROF(h, "lk_kernel");
  lk_kernel<<<grd, WARPS * 32, 0, h->stream>>>(aa);
  LVB_LAUNCH_CHECK(h);
  return LVB_OK;
}
