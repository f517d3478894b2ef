// Front-end image stage: CLAHE(3.0, 8x8) -> padded 3-level LK pyramid -> ORB 7x7 blur.
// Replaces ImageProcessor::createImagePyramids (image_processor.cpp:318-334: cv::CLAHE::apply +
// cv::buildOpticalFlowPyramid) and the per-frame work of ORBdescriptor::initializeLayerAndPyramid
// (ORBDescriptor.cpp:418-483, level 0 only).  Integer stages are bit-exact w.r.t. OpenCV 4.13
// (SURVEY.md App. A.1, A.2, A.4); Scharr derivative images are NOT materialised (the LK kernel
// recomputes them from the padded level), which removes 4x the pixel bytes from HBM.
#include "lvb_internal.h"

namespace {

constexpr int kTiles = 8;   // CLAHE tilesX = tilesY = 8

// ---------------------------------------------------------------- CLAHE histogram -> LUT
// grid (64, n): one CTA per tile.  App. A.1.
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ img, int W, int H,
                                                         uint8_t* __restrict__ lut, int clip_limit) {
  __shared__ int hist[256];
  __shared__ int warp_sums[8];
  __shared__ int s_clipped;
  const int tile = blockIdx.x, s = blockIdx.y;
  const int tw = W / kTiles, th = H / kTiles;
  const int tx = tile % kTiles, ty = tile / kTiles;
  const uint8_t* src = img + (size_t)s * W * H + (size_t)(ty * th) * W + tx * tw;
  const int tid = threadIdx.x;
  hist[tid] = 0;
  if (tid == 0) s_clipped = 0;
  __syncthreads();
  const int total = tw * th;
  for (int i = tid; i < total; i += 256) {
    int y = i / tw, x = i - y * tw;
    atomicAdd(&hist[src[(size_t)y * W + x]], 1);
  }
  __syncthreads();
  int hv = hist[tid];
  int excess = hv > clip_limit ? hv - clip_limit : 0;
  if (hv > clip_limit) hv = clip_limit;
  // block reduce excess
  int v = excess;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) warp_sums[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += warp_sums[i];
    s_clipped = t;
  }
  __syncthreads();
  const int clipped = s_clipped;
  const int batch = clipped / 256;
  int residual = clipped - batch * 256;
  hv += batch;
  if (residual != 0) {
    int step = 256 / residual;
    if (step < 1) step = 1;
    // bins 0, step, 2*step, ... get +1 while residual lasts (and index < 256)
    if (tid % step == 0 && tid / step < residual) hv += 1;
  }
  // inclusive scan of hv over 256 bins
  int x = hv;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((tid & 31) >= o) x += y;
  }
  __syncthreads();
  if ((tid & 31) == 31) warp_sums[tid >> 5] = x;
  __syncthreads();
  int base = 0;
  for (int i = 0; i < (tid >> 5); ++i) base += warp_sums[i];
  const int cum = x + base;
  const float lut_scale = 255.0f / (float)total;
  float f = __fmul_rn((float)cum, lut_scale);
  int r = __float2int_rn(f);
  r = r < 0 ? 0 : (r > 255 ? 255 : r);
  lut[((size_t)s * 64 + tile) * 256 + tid] = (uint8_t)r;
}

// REFLECT_101 pad written by the producer of a level: pixel (x, y) of a w x h level is also the value of its mirror images
// -x, 2(w-1)-x (and the same in y) wherever those fall inside the LVB_PAD-wide pad (w, h > LVB_PAD + 1 at every level we build).
__device__ __forceinline__ void store_pad_copies(uint8_t* org, int pitch, int w, int h, int x, int y, uint8_t v) {
  int xs[2], ys[2], nx = 0, ny = 0;
  if (x >= 1 && x <= LVB_PAD) xs[nx++] = -x;
  if (x <= w - 2 && x >= w - 1 - LVB_PAD) xs[nx++] = 2 * (w - 1) - x;
  if (y >= 1 && y <= LVB_PAD) ys[ny++] = -y;
  if (y <= h - 2 && y >= h - 1 - LVB_PAD) ys[ny++] = 2 * (h - 1) - y;
  for (int a = 0; a < nx; ++a) org[(ptrdiff_t)y * pitch + xs[a]] = v;
  for (int b = 0; b < ny; ++b) {
    org[(ptrdiff_t)ys[b] * pitch + x] = v;
    for (int a = 0; a < nx; ++a) org[(ptrdiff_t)ys[b] * pitch + xs[a]] = v;
  }
}

// ---------------------------------------------------------------- CLAHE apply -> L0 interior + its REFLECT_101 pad
// each thread: 4 horizontally adjacent pixels.
__global__ void __launch_bounds__(256) clahe_apply_kernel(const uint8_t* __restrict__ img, int W, int H,
                                                           const uint8_t* __restrict__ lut,
                                                           uint8_t* __restrict__ pyr, LvbPyramidLayout L,
                                                           int equalize) {
  const int s = blockIdx.z;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= W || y >= H) return;
  const uint8_t* src = img + (size_t)s * W * H + (size_t)y * W + x0;
  uint8_t* dst = lvb_level_origin(pyr, L, s, 0) + (size_t)y * L.lv[0].pitch + x0;
  uchar4 in = *reinterpret_cast<const uchar4*>(src);
  uint8_t px[4] = {in.x, in.y, in.z, in.w};
  uint8_t out[4];
  if (!equalize) {
    for (int k = 0; k < 4; ++k) out[k] = px[k];
  } else {
    const int tw = W / kTiles, th = H / kTiles;
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    int ty2 = ty1 + 1;
    const float ya = __fsub_rn(tyf, (float)ty1);
    const float ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0);
    ty2 = min(ty2, kTiles - 1);
    const uint8_t* lutS = lut + (size_t)s * 64 * 256;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
      int tx1 = (int)floorf(txf);
      int tx2 = tx1 + 1;
      const float xa = __fsub_rn(txf, (float)tx1);
      const float xa1 = __fsub_rn(1.0f, xa);
      tx1 = max(tx1, 0);
      tx2 = min(tx2, kTiles - 1);
      const int v = px[k];
      const float l11 = (float)__ldg(lutS + (ty1 * kTiles + tx1) * 256 + v);
      const float l12 = (float)__ldg(lutS + (ty1 * kTiles + tx2) * 256 + v);
      const float l21 = (float)__ldg(lutS + (ty2 * kTiles + tx1) * 256 + v);
      const float l22 = (float)__ldg(lutS + (ty2 * kTiles + tx2) * 256 + v);
      const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
      const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
      const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
      int r = __float2int_rn(res);
      out[k] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
  }
  *reinterpret_cast<uchar4*>(dst) = make_uchar4(out[0], out[1], out[2], out[3]);
  if (x0 <= LVB_PAD || x0 + 3 >= W - 1 - LVB_PAD || y <= LVB_PAD || y >= H - 1 - LVB_PAD) {     // border band: also fill the pad
    uint8_t* org = lvb_level_origin(pyr, L, s, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) store_pad_copies(org, L.lv[0].pitch, W, H, x0 + k, y, out[k]);
  }
}

// ---------------------------------------------------------------- pyrDown (5x5 binomial, (x+128)>>8)
// source level must already have a valid REFLECT_101 pad (>= 2 px). App. A.2.
// each thread produces 4 adjacent destination pixels from 5 source rows x 16 bytes (four aligned 32-bit loads per row)
__global__ void __launch_bounds__(256) pyrdown_kernel(uint8_t* __restrict__ pyr, LvbPyramidLayout L, int src_level) {
  const int s = blockIdx.z;
  const LvbLevel ls = L.lv[src_level], ld = L.lv[src_level + 1];
  const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;              // first of 4 destination columns (block = 32 x 8)
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x4 >= ld.w || y >= ld.h) return;
  const uint8_t* so = lvb_level_origin((const uint8_t*)pyr, L, s, src_level);
  uint8_t* dorg = lvb_level_origin(pyr, L, s, src_level + 1);
  int acc[4] = {0, 0, 0, 0};
  const int wv[5] = {1, 4, 6, 4, 1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    // bytes [2*x4 - 4, 2*x4 + 12) of source row 2y-2+j; the level origin is 8-byte aligned and 2*x4 is a multiple of 8
    const uint32_t* row = reinterpret_cast<const uint32_t*>(so + (ptrdiff_t)(2 * y - 2 + j) * ls.pitch + (2 * x4 - 4));
    const uint32_t w0 = row[0], w1 = row[1], w2 = row[2], w3 = row[3];
    uint8_t b[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) { b[q] = (w0 >> (8 * q)) & 255; b[4 + q] = (w1 >> (8 * q)) & 255; b[8 + q] = (w2 >> (8 * q)) & 255; b[12 + q] = (w3 >> (8 * q)) & 255; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 4 + 2 * k;      // byte index of source column 2*(x4+k)
      acc[k] += wv[j] * (b[c - 2] + 4 * b[c - 1] + 6 * b[c] + 4 * b[c + 1] + b[c + 2]);
    }
  }
  uint8_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = (uint8_t)((acc[k] + 128) >> 8);
  uint8_t* dst = dorg + (ptrdiff_t)y * ld.pitch + x4;
  if (x4 + 3 < ld.w) *reinterpret_cast<uchar4*>(dst) = make_uchar4(o[0], o[1], o[2], o[3]);
  else for (int k = 0; k < 4 && x4 + k < ld.w; ++k) dst[k] = o[k];
  if (x4 <= LVB_PAD || x4 + 3 >= ld.w - 1 - LVB_PAD || y <= LVB_PAD || y >= ld.h - 1 - LVB_PAD) {   // the destination level's pad
    for (int k = 0; k < 4 && x4 + k < ld.w; ++k) store_pad_copies(dorg, ld.pitch, ld.w, ld.h, x4 + k, y, o[k]);
  }
}

// ---------------------------------------------------------------- 7x7 sigma=2 fixed-point blur of L0
// App. A.4: kernel [18 34 48 56 48 34 18]/256 in both directions, out = (acc + 32768) >> 16.
// CTA tile 64x16 output, staged with a 3-px apron in shared memory.
// CTA tile 128x32 output; input staged as aligned 32-bit words with a 4-byte / 3-row apron; each thread filters
// 4 adjacent pixels horizontally (u16 partial sums in shared memory) and a 4x4 block vertically.
constexpr int BT_W = 128, BT_H = 32;
__global__ void __launch_bounds__(256) blur7_kernel(const uint8_t* __restrict__ pyr, LvbPyramidLayout L,
                                                     uint8_t* __restrict__ blur) {
  __shared__ uint32_t tin[BT_H + 6][(BT_W + 8) / 4 + 1];
  __shared__ unsigned short hs[BT_H + 6][BT_W + 4];
  const int s = blockIdx.z;
  const LvbLevel lv = L.lv[0];
  const int bx = blockIdx.x * BT_W, by = blockIdx.y * BT_H;
  const uint8_t* org = lvb_level_origin(pyr, L, s, 0);
  const int tid = threadIdx.x;
  constexpr int WPR = (BT_W + 8) / 4;      // 34 words per staged row: bytes [bx-4, bx+BT_W+4)
  for (int i = tid; i < (BT_H + 6) * WPR; i += 256) {
    const int r = i / WPR, wq = i - r * WPR;
    int gy = by - 3 + r;
    gy = min(gy, lv.h + LVB_PAD - 1);
    int gx = bx - 4 + 4 * wq;
    gx = min(gx, lv.pitch - LVB_PAD - 4);   // stay inside the padded row
    tin[r][wq] = *reinterpret_cast<const uint32_t*>(org + (ptrdiff_t)gy * lv.pitch + gx);
  }
  __syncthreads();
  const int k7[7] = {18, 34, 48, 56, 48, 34, 18};
  for (int it = tid; it < (BT_H + 6) * (BT_W / 4); it += 256) {
    const int r = it / (BT_W / 4), g = it - r * (BT_W / 4);
    const uint32_t w0 = tin[r][g], w1 = tin[r][g + 1], w2 = tin[r][g + 2];
    int b[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) { b[q] = (w0 >> (8 * q)) & 255; b[4 + q] = (w1 >> (8 * q)) & 255; b[8 + q] = (w2 >> (8 * q)) & 255; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int a = 0;
#pragma unroll
      for (int t = 0; t < 7; ++t) a += k7[t] * b[k + 1 + t];     // pixel 4g+k sits at byte 4+k; taps at 1+k .. 7+k
      hs[r][4 * g + k] = (unsigned short)a;
    }
  }
  __syncthreads();
  const int g = tid & 31, rg = tid >> 5;         // 4 columns x 4 rows per thread
  const int gx = bx + 4 * g;
  if (gx >= lv.w) return;
  int hv[10][4];
#pragma unroll
  for (int r = 0; r < 10; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) hv[r][k] = hs[rg * 4 + r][4 * g + k];
#pragma unroll
  for (int ry = 0; ry < 4; ++ry) {
    const int gy = by + rg * 4 + ry;
    if (gy >= lv.h) break;
    uint8_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int a = 0;
#pragma unroll
      for (int t = 0; t < 7; ++t) a += k7[t] * hv[ry + t][k];
      o[k] = (uint8_t)((a + 32768) >> 16);
    }
    *reinterpret_cast<uchar4*>(blur + (size_t)s * lv.w * lv.h + (size_t)gy * lv.w + gx) = make_uchar4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace

// Build CLAHE'd L0 + pyramid + blur for n images already in device memory.
int fe_build_pyramid(LvbHandle* h, const uint8_t* d_images, int n, uint8_t* pyr, uint8_t* blur) {
  LvbFrontEnd& fe = h->fe;
  const int W = fe.W, H = fe.H;
  cudaStream_t st = h->stream;
  if (h->cfg.flag_equalize) {
    const int tile_total = (W / kTiles) * (H / kTiles);
    int clip = (int)(3.0 * tile_total / 256);   // CLAHE clipLimit 3.0, histSize 256
    if (clip < 1) clip = 1;
    LVB_PROF(h, "clahe_lut_kernel");
    clahe_lut_kernel<<<dim3(64, n), 256, 0, st>>>(d_images, W, H, fe.lut, clip);
    LVB_LAUNCH_CHECK(h);
  }
  {
    dim3 blk(32, 8);
    dim3 grd((W / 4 + 31) / 32, (H + 7) / 8, n);
    LVB_PROF(h, "clahe_apply_kernel");
    clahe_apply_kernel<<<grd, blk, 0, st>>>(d_images, W, H, fe.lut, pyr, fe.L, h->cfg.flag_equalize);
    LVB_LAUNCH_CHECK(h);
  }
  for (int l = 0; l < fe.L.n_levels; ++l) {       // every producer writes the REFLECT_101 pad of the level it produces
    dim3 blk(32, 8);
    if (l + 1 < fe.L.n_levels) {
      const LvbLevel& ld = fe.L.lv[l + 1];
      dim3 g2(((ld.w + 3) / 4 + 31) / 32, (ld.h + 7) / 8, n);
      LVB_PROF(h, "pyrdown_kernel");
      pyrdown_kernel<<<g2, blk, 0, st>>>(pyr, fe.L, l);
      LVB_LAUNCH_CHECK(h);
    }
  }
  if (blur) {
    dim3 grd((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H, n);
    LVB_PROF(h, "blur7_kernel");
    blur7_kernel<<<grd, 256, 0, st>>>(pyr, fe.L, blur);
    LVB_LAUNCH_CHECK(h);
  }
  return LVB_OK;
}
