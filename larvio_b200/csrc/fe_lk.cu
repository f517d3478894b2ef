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
// Data movement (sm_100a): per pyramid level a warp stages two tiles with TMA (cp.async.bulk.tensor.3d + a per-warp
// mbarrier, SASS UTMALDG): the 48x24-byte source window of the previous image and a 48x32-byte search tile of the next
// image around the predicted position.  TMA wants the innermost coordinate 16-byte aligned, so a tile starts at the
// 16-aligned column left of what is needed and the window sits at a 0..15 byte offset inside it.  All <=30 iterations
// of the level run out of the shared-memory tile (re-staged only if the window walks more than 5 px away), so the
// dependent iteration chain never waits on global memory.  Each lane keeps its 16 window pixels (patch value, both
// derivatives) in registers for the whole level; bilinear blends are IDP.2A / IDP.4A dot products on packed bytes.
//
// Bit-exactness with OpenCV's SSE accumulation order (oracle/lk_exact.py): OpenCV accumulates every sum of the normal
// equations in 4 SSE lanes + a scalar tail, i.e. five sequential float chains per sum, and on CLAHE-contrast images the
// partial sums exceed 2^24, so the order is part of the result.  All lanes produce the terms into per-chain arrays
// (zero-padded to a common length: x + 0 is exact), then one lane per chain adds its terms strictly in order - 15
// chains for (A11, A12, A22) once per level, 10 for (b1, b2) per iteration, all in one uniform 27 x float4 loop.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "lvb_internal.h"

namespace {

constexpr int WIN = 21;
constexpr int DT = WIN + 1;                // 22: derivative grid
constexpr int WARPS = 4;
constexpr int W_BITS = 14;
constexpr int TP = 48;                     // tile row pitch = TMA box width in bytes (a multiple of 16)
constexpr int TA_ROWS = 24, TB_ROWS = 32;  // source tile rows (window + bilinear + Scharr apron), search tile rows
constexpr int MARGIN = 5;                  // the search window may drift +-MARGIN px inside the staged tile
constexpr int CH = 108;                    // floats per chain: 105 tail terms / 84 (A) or 42 (b) lane terms, zero-padded

struct alignas(128) WarpSmem {
  uint8_t tileB[TP * TB_ROWS + 128];       // search tile of the next image (+ slack: masked lanes read a few bytes past the end)
  uint8_t tileA[TP * TA_ROWS];             // source tile of the previous image
  float term[15 * CH];                     // chain c = sum * 5 + accumulator (accumulator 4 = scalar tail); b sums use chains 0..9
  unsigned long long bar[2];               // [0] source tile, [1] search tile
};

// One chain: strictly sequential float adds of its CH terms (pads are +0.0f: exact no-ops).
__device__ __forceinline__ float run_chain(const float* T) {
  const float4* q = reinterpret_cast<const float4*>(T);
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < CH / 4; ++i) {
    const float4 v = q[i];
    a = __fadd_rn(a, v.x); a = __fadd_rn(a, v.y); a = __fadd_rn(a, v.z); a = __fadd_rn(a, v.w);
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

// ---- sm_100a primitives
__device__ __forceinline__ int dp2a_lo(int w, unsigned b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi(int w, unsigned b, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory"); }
// x must be a multiple of 16 (bytes): TMA faults on a misaligned innermost coordinate
__device__ __forceinline__ void tma_load_tile(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ int floor16(int v) { return v & ~15; }        // two's complement: also floors negative values

// 12 bytes [b, b+12) of a tile row as three words, b arbitrary: three aligned 8-byte reads, word select, funnel shifts
__device__ __forceinline__ void row_bytes12(const uint8_t* rowp, int b, unsigned& a0, unsigned& a1, unsigned& a2) {
  const uint8_t* q = rowp + (b & ~7);
  const uint2 r0 = *reinterpret_cast<const uint2*>(q), r1 = *reinterpret_cast<const uint2*>(q + 8), r2 = *reinterpret_cast<const uint2*>(q + 16);
  const bool hi = (b & 4) != 0;
  const unsigned w0 = hi ? r0.y : r0.x, w1 = hi ? r1.x : r0.y, w2 = hi ? r1.y : r1.x, w3 = hi ? r2.x : r1.y;
  const int k = (b & 3) * 8;
  a0 = __funnelshift_r(w0, w1, k); a1 = __funnelshift_r(w1, w2, k); a2 = __funnelshift_r(w2, w3, k);
}

// Eight horizontally adjacent pixels of one tile row pair: bytes [b, b+9) of row `rowp` (top) and of the row below,
// bilinear weights wt = iw00 | iw01 << 16, wb = iw10 | iw11 << 16, per-pixel accumulator start c[j].
// out[j] = c[j] + iw00*T[j] + iw01*T[j+1] + iw10*B[j] + iw11*B[j+1]            (two IDP.2A per pixel)
__device__ __forceinline__ void bilinear8(const uint8_t* rowp, int b, int wt, int wb, const int* c, int* out) {
  const uint8_t* q = rowp + (b & ~7);
  const uint2 t0 = *reinterpret_cast<const uint2*>(q), t1 = *reinterpret_cast<const uint2*>(q + 8);
  const uint2 u0 = *reinterpret_cast<const uint2*>(q + TP), u1 = *reinterpret_cast<const uint2*>(q + TP + 8);
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

// One 8-pixel item (window row y, columns x0..x0+7, nv of them valid) of the previous-image window from the staged source
// tile; tile byte (r, oxa + c) = level pixel (ipy - 1 + r, ipx - 1 + c).  Scharr at grid point (g, c) [level pixel
// (ipy + g, ipx + c)] needs tile rows g..g+2 and tile columns c..c+2:
//   dx = 3 (Hd[g] + Hd[g+2]) + 10 Hd[g+1],  Hd[r][c] = tile[r][c+2] - tile[r][c]              (IDP.4A with weights -1 0 1)
//   dy = Hs[g+2] - Hs[g],                   Hs[r][c] = 3 tile[r][c] + 10 tile[r][c+1] + 3 tile[r][c+2]   (weights 3 10 3)
// then OpenCV's Q14 bilinear blend with rounding of the four neighbouring grid values, and the Q14 blend >> 9 of the pixels.
__device__ __forceinline__ void window_item(const uint8_t* tileA, int oxa, int y, int x0, int nv, int iw00, int iw01, int iw10, int iw11,
                                            bool grid_inside, int ipx, int ipy, int lw, int lh, int* cI, int* gX, int* gY) {
  int dxg[2][9], dyg[2][9];
  int hs1[9], hs2[9], hd1[9], hd2[9];       // the two previous tile rows
  unsigned p1[9], p2[9];                    // byte-shifted words of tile rows y+1 and y+2 (patch value)
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    unsigned a0, a1, a2;
    row_bytes12(tileA + (y + t) * TP, oxa + x0, a0, a1, a2);
    unsigned sc[9];
    sc[0] = a0; sc[1] = __funnelshift_r(a0, a1, 8); sc[2] = __funnelshift_r(a0, a1, 16); sc[3] = __funnelshift_r(a0, a1, 24);
    sc[4] = a1; sc[5] = __funnelshift_r(a1, a2, 8); sc[6] = __funnelshift_r(a1, a2, 16); sc[7] = __funnelshift_r(a1, a2, 24);
    sc[8] = a2;
    int hs[9], hd[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) { hs[c] = dp4a_us(sc[c], 0x00030A03, 0); hd[c] = dp4a_us(sc[c], 0x000100FF, 0); }
    if (t >= 2) {
#pragma unroll
      for (int c = 0; c < 9; ++c) { dxg[t - 2][c] = 3 * (hd2[c] + hd[c]) + 10 * hd1[c]; dyg[t - 2][c] = hs[c] - hs2[c]; }
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      hs2[c] = hs1[c]; hd2[c] = hd1[c]; hs1[c] = hs[c]; hd1[c] = hd[c];
      if (t == 1) p1[c] = sc[c];
      if (t == 2) p2[c] = sc[c];
    }
  }
  if (!grid_inside) {                       // derivative image is zero outside the level (warp-uniform branch)
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const int gx = ipx + x0 + c, gy = ipy + y + g;
        if (gx < 0 || gx >= lw || gy < 0 || gy >= lh) { dxg[g][c] = 0; dyg[g][c] = 0; }
      }
  }
  const int wt = (iw00 & 0xffff) | (iw01 << 16), wb = (iw10 & 0xffff) | (iw11 << 16);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ix = (dxg[0][j] * iw00 + dxg[0][j + 1] * iw01 + dxg[1][j] * iw10 + dxg[1][j + 1] * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
    const int iy = (dyg[0][j] * iw00 + dyg[0][j + 1] * iw01 + dyg[1][j] * iw10 + dyg[1][j + 1] * iw11 + (1 << (W_BITS - 1))) >> W_BITS;
    const int iv = dp2a_lo(wt, p1[j + 1], dp2a_lo(wb, p2[j + 1], 1 << (W_BITS - 5 - 1)));     // tile bytes (j+1, j+2) of rows y+1, y+2
    const bool ok = j < nv;
    cI[j] = ok ? (1 << (W_BITS - 5 - 1)) - ((iv >> (W_BITS - 5)) << (W_BITS - 5)) : 0;
    gX[j] = ok ? ix : 0;
    gY[j] = ok ? iy : 0;
  }
}

// debug twin of tma_load_tile: same box, same zero fill outside the padded level
__device__ __forceinline__ void debug_load_tile(uint8_t* dst, int rows, const uint8_t* pyr, const LvbPyramidLayout& L, int level, int s, int x, int y, int lane) {
  const LvbLevel lv = L.lv[level];
  const uint8_t* base = pyr + (size_t)s * L.bytes_per_seq + lv.offset;
  for (int t = lane; t < TP * rows; t += 32) {
    const int r = t / TP, c = t - r * TP;
    const int gx = x + c, gy = y + r;
    dst[t] = (gx >= 0 && gx < lv.pitch && gy >= 0 && gy < lv.rows) ? base[(size_t)gy * lv.pitch + gx] : (uint8_t)0;
  }
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
  CUtensorMap mapA[LVB_MAX_LEVELS];   // previous pyramid, box TP x TA_ROWS x 1 per level: (x, y, sequence) over the padded level images
  CUtensorMap mapB[LVB_MAX_LEVELS];   // next pyramid, box TP x TB_ROWS x 1
  LkArgs a[2];
  int lw[LVB_MAX_LEVELS], lh[LVB_MAX_LEVELS];
  int max_iter; double eps2; double min_eig;
  int max_level;
  unsigned long long* stats;  // [10] iterations, [13] search-tile re-stages
  // debug only (LVB_DEBUG_LK_NOTMA=1): stage the tiles with plain loads instead of TMA, to tell a staging fault from an arithmetic one
  int no_tma; const uint8_t* pyrA; const uint8_t* pyrB; LvbPyramidLayout L;
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
  for (int t = lane; t < 15 * CH / 4; t += 32) reinterpret_cast<float4*>(w.term)[t] = make_float4(0.f, 0.f, 0.f, 0.f);   // chain pads stay zero
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
  // where this lane's terms go.  A sums: pixel (y, x): x < 16 -> chain x & 3, position y * 4 + (x >> 2); else chain 4, position
  // y * 5 + (x - 16).  b sums: pair term p of chunk ch -> chain p, position y * 2 + ch; tail pixel j -> chain 4, position y * 5 + j.
  // Per item the k-th value lands at base + k * step (row lanes: the next chain; tail lanes: the next position).
  const int stepB = tail ? 1 : CH;
  const int baseB_A = tail ? 4 * CH + rowA * 5 : rowA * 2 + (xA >> 3);
  const int baseB_B = tail ? 4 * CH + rowB * 5 : rowB * 2 + (xB >> 3);
  const int pair_mask = tail ? 0 : -1;                                   // row lanes add pixel k + 4 to pixel k (OpenCV's madd pairs)

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
  unsigned n_iter = 0, n_restage = 0;

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
    // ---- stage the source tile (needs columns ipx-1 .. ipx+22, rows ipy-1 .. ipy+22) and the search tile
    const int txa = floor16(ipx - 1 + LVB_PAD), oxa = ipx - 1 + LVB_PAD - txa;                     // padded-level coordinates
    int txb, ty;                                                                                   // search tile origin (padded x, level y)
    {
      // float -> int with saturation (np may be far outside for a lost point); the iteration re-checks the range anyway
      const float cx = fminf(fmaxf(floorf(np.x), -4096.f), 8192.f), cy = fminf(fmaxf(floorf(np.y), -4096.f), 8192.f);
      txb = max(floor16((int)cx - MARGIN + LVB_PAD), 0); ty = max((int)cy - MARGIN, -LVB_PAD);   // tile origins stay inside the padded level
    }
    __syncwarp();
    if (aa.no_tma) {
      debug_load_tile(w.tileA, TA_ROWS, aa.pyrA, aa.L, level, s, txa, ipy - 1 + LVB_PAD, lane);
      debug_load_tile(w.tileB, TB_ROWS, aa.pyrB, aa.L, level, s, txb, ty + LVB_PAD, lane);
      __syncwarp();
    } else {
      if (lane == 0) {
        tma_load_tile(smem_u32(w.tileA), &aa.mapA[level], txa, ipy - 1 + LVB_PAD, s, bar0, TP * TA_ROWS);
        tma_load_tile(smem_u32(w.tileB), &aa.mapB[level], txb, ty + LVB_PAD, s, bar1, TP * TB_ROWS);
      }
      mbar_wait(bar0, ph0); ph0 ^= 1;
    }
    // ---- this lane's 16 pixels of the previous-image window, straight from the staged source tile: Scharr derivatives
    //      (zero outside the image) at the 2 x 9 grid points each 8-pixel item touches, their bilinear blend gX, gY, and
    //      the patch value folded into the accumulator start cI = 256 - (Iw << 9)  (I_t = (cI + sum w*J) >> 9 later)
    float fa = __fsub_rn(prevPt.x, (float)ipx), fb = __fsub_rn(prevPt.y, (float)ipy);
    int iw00 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw01 = __float2int_rn(__fmul_rn(__fmul_rn(fa, __fsub_rn(1.f, fb)), (float)(1 << W_BITS)));
    int iw10 = __float2int_rn(__fmul_rn(__fmul_rn(__fsub_rn(1.f, fa), fb), (float)(1 << W_BITS)));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    const bool grid_inside = ipx >= 0 && ipx + DT <= lw && ipy >= 0 && ipy + DT <= lh;       // warp-uniform: no zero padding needed
    int cI[16], gX[16], gY[16];
    window_item(w.tileA, oxa, rowA, xA, nvA, iw00, iw01, iw10, iw11, grid_inside, ipx, ipy, lw, lh, &cI[0], &gX[0], &gY[0]);
    window_item(w.tileA, oxa, rowB, xB, nvB, iw00, iw01, iw10, iw11, grid_inside, ipx, ipy, lw, lh, &cI[8], &gX[8], &gY[8]);
    // ---- A sums: terms fx*fx, fx*fy, fy*fy into chains 0..4 / 5..9 / 10..14, then 15 lanes replay OpenCV's order
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int it = j >> 3, jj = j & 7;
      const int y = it ? rowB : rowA, x = (it ? xB : xA) + jj, nv = it ? nvB : nvA;
      if (jj < nv) {
        const float fx = (float)gX[j], fy = (float)gY[j];
        const int sl = (x < 16) ? ((x & 3) * CH + y * 4 + (x >> 2)) : (4 * CH + y * 5 + (x - 16));
        w.term[sl] = __fmul_rn(fx, fx);
        w.term[5 * CH + sl] = __fmul_rn(fx, fy);
        w.term[10 * CH + sl] = __fmul_rn(fy, fy);
      }
    }
    __syncwarp();
    float chainv = 0.f;
    if (lane < 15) chainv = run_chain(w.term + lane * CH);
    const float A11 = __fmul_rn(combine_chains(chainv, 0), FLT_SCALE);
    const float A12 = __fmul_rn(combine_chains(chainv, 5), FLT_SCALE);
    const float A22 = __fmul_rn(combine_chains(chainv, 10), FLT_SCALE);
    __syncwarp();
    // the b sums reuse chains 0..9; their lane chains hold 42 terms where the A chains held 84: clear the difference once
    for (int t = lane; t < 8 * 42; t += 32) { const int c = t / 42, q = t - c * 42; w.term[(c + (c >> 2)) * CH + 42 + q] = 0.f; }

    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dif = __fsub_rn(A11, A22);
    const float rad = __fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WIN * WIN));
    if (!aa.no_tma) { mbar_wait(bar1, ph1); ph1 ^= 1; }                  // search tile has landed (needed before any `continue`)
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
      int ox = inx + LVB_PAD - txb, oy = iny - ty;
      if (ox < 0 || ox > TP - DT || oy < 0 || oy > 2 * MARGIN) {         // the window left the staged tile: re-centre it
        txb = max(floor16(inx - MARGIN + LVB_PAD), 0); ty = max(iny - MARGIN, -LVB_PAD); ox = inx + LVB_PAD - txb; oy = iny - ty;
        ++n_restage;
        __syncwarp();
        if (aa.no_tma) { debug_load_tile(w.tileB, TB_ROWS, aa.pyrB, aa.L, level, s, txb, ty + LVB_PAD, lane); __syncwarp(); }
        else {
          if (lane == 0) tma_load_tile(smem_u32(w.tileB), &aa.mapB[level], txb, ty + LVB_PAD, s, bar1, TP * TB_ROWS);
          mbar_wait(bar1, ph1); ph1 ^= 1;
        }
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
      bilinear8(&w.tileB[(oy + rowA) * TP], ox + xA, wt, wb, &cI[0], &d[0]);
      bilinear8(&w.tileB[(oy + rowB) * TP], ox + xB, wt, wb, &cI[8], &d[8]);
      ++n_iter;
      // terms of b1 (chains 0..4) and b2 (chains 5..9): row lanes 4 pair terms per item, tail lanes 5 single terms
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int nv = it ? nvB : nvA;
        float* t1 = w.term + (it ? baseB_B : baseB_A);
        float* t2 = t1 + 5 * CH;
        int p1[8], p2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int dq = d[it * 8 + q] >> (W_BITS - 5); p1[q] = dq * gX[it * 8 + q]; p2[q] = dq * gY[it * 8 + q]; }
        if (nv > 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            t1[k * stepB] = (float)(p1[k] + (p1[k + 4] & pair_mask));
            t2[k * stepB] = (float)(p2[k] + (p2[k + 4] & pair_mask));
          }
          if (tail) { t1[4] = (float)p1[4]; t2[4] = (float)p2[4]; }
        }
      }
      __syncwarp();
      float cv = 0.f;
      if (lane < 10) cv = run_chain(w.term + lane * CH);
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
      if (n_restage) atomicAdd(&aa.stats[13], (unsigned long long)n_restage);
    }
  }
}

}  // namespace

// ---- TMA descriptors of a padded pyramid block: one 3-D u8 tensor (x, y, sequence) per level, box TP x rows x 1
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

static int make_level_maps(LvbHandle* h, const uint8_t* pyr, int n_seq, int box_rows, CUtensorMap* out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return lvb_set_err(LVB_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  const LvbPyramidLayout& L = h->fe.L;
  for (int l = 0; l < L.n_levels; ++l) {
    const LvbLevel& lv = L.lv[l];
    const cuuint64_t dims[3] = {(cuuint64_t)lv.pitch, (cuuint64_t)lv.rows, (cuuint64_t)n_seq};
    const cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)L.bytes_per_seq};
    const cuuint32_t box[3] = {(cuuint32_t)TP, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(&out[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)(pyr + lv.offset), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return lvb_set_err(LVB_E_CUDA, "cuTensorMapEncodeTiled(level %d) -> %d", l, (int)r);
  }
  return LVB_OK;
}

// the handle's own ping-pong pyramids keep their descriptors (two box shapes each: source role, search role); temporary
// pyramids (stage-level entry points) get fresh ones
static int level_maps(LvbHandle* h, const uint8_t* pyr, int n_seq, int role /*0 source tile, 1 search tile*/, CUtensorMap* out) {
  const int box_rows = role == 0 ? TA_ROWS : TB_ROWS;
  for (int k = 0; k < 2; ++k) {
    if (pyr == h->fe.pyr[k] && n_seq == h->fe.S) {
      if (!h->lk_maps_ok[k][role]) {
        int rc = make_level_maps(h, pyr, n_seq, box_rows, h->lk_maps[k][role]);
        if (rc != LVB_OK) return rc;
        h->lk_maps_ok[k][role] = true;
      }
      memcpy(out, h->lk_maps[k][role], sizeof(CUtensorMap) * LVB_MAX_LEVELS);
      return LVB_OK;
    }
  }
  return make_level_maps(h, pyr, n_seq, box_rows, out);
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
  int rc = level_maps(h, pyrA, n_seq, 0, aa.mapA);
  if (rc != LVB_OK) return rc;
  rc = level_maps(h, pyrB, n_seq, 1, aa.mapB);
  if (rc != LVB_OK) return rc;
  for (int l = 0; l < h->fe.L.n_levels; ++l) { aa.lw[l] = h->fe.L.lv[l].w; aa.lh[l] = h->fe.L.lv[l].h; }
  int mi = h->cfg.max_iteration; if (mi < 0) mi = 0; if (mi > 100) mi = 100;      // cv clamps maxCount to [0,100]
  double eps = h->cfg.track_precision; if (eps < 0) eps = 0; if (eps > 10) eps = 10;
  aa.max_iter = mi; aa.eps2 = eps * eps; aa.min_eig = 1e-4;
  aa.max_level = h->cfg.pyramid_levels;
  aa.stats = h->fe.stats;
  aa.no_tma = getenv("LVB_DEBUG_LK_NOTMA") ? 1 : 0; aa.pyrA = pyrA; aa.pyrB = pyrB; aa.L = h->fe.L;
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
