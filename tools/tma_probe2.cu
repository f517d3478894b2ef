// Second probe: which async-copy forms execute on this box.  One variant per process (TMA_PROBE=a|b|c|d|e).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ unsigned finish(const uint8_t* tile, unsigned b, int n) {
  unsigned ok = 0;
  for (int it = 0; it < 4000000 && !ok; ++it)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(0) : "memory");
  unsigned s = 0;
  for (int i = threadIdx.x; i < n; i += 32) s += tile[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return ok ? s : 0xdeadbeefu;
}
// a: 2-D tensor map, shared::cluster form
__global__ void k_a(const __grid_constant__ CUtensorMap m, int x, int y, unsigned* out) {
  __shared__ alignas(128) uint8_t tile[1024]; __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar), d = smem_u32(tile);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b)); asm volatile("fence.proxy.async.shared::cta;"); }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" :: "r"(d), "l"(&m), "r"(x), "r"(y), "r"(b) : "memory");
  }
  const unsigned s = finish(tile, b, 1024);
  if (threadIdx.x == 0) out[0] = s;
}
// b: 3-D tensor map, shared::cta form (PTX 8.6)
__global__ void k_b(const __grid_constant__ CUtensorMap m, int x, int y, int z, unsigned* out) {
  __shared__ alignas(128) uint8_t tile[1024]; __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar), d = smem_u32(tile);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b)); asm volatile("fence.proxy.async.shared::cta;"); }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" :: "r"(d), "l"(&m), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
  }
  const unsigned s = finish(tile, b, 1024);
  if (threadIdx.x == 0) out[0] = s;
}
// c: libcu++ (the programming guide's own example), 2-D
__global__ void k_c(const __grid_constant__ CUtensorMap m, int x, int y, unsigned* out) {
  __shared__ alignas(128) uint8_t tile[1024];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) { cde::cp_async_bulk_tensor_2d_global_to_shared(tile, &m, x, y, bar); token = cuda::device::barrier_arrive_tx(bar, 1, 1024); }
  else token = bar.arrive();
  bar.wait(std::move(token));
  unsigned s = 0;
  for (int i = threadIdx.x; i < 1024; i += 32) s += tile[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) out[0] = s;
}
// d: non-tensor bulk copies, one 48-byte row segment per lane (16-byte aligned source), one mbarrier
__global__ void k_d(const uint8_t* src, int pitch, unsigned* out) {
  __shared__ alignas(128) uint8_t tile[32 * 48]; __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b)); asm volatile("fence.proxy.async.shared::cta;"); }
  __syncwarp();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(32 * 48) : "memory");
  __syncwarp();
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(tile + threadIdx.x * 48)), "l"(src + (size_t)threadIdx.x * pitch), "r"(48), "r"(b) : "memory");
  const unsigned s = finish(tile, b, 32 * 48);
  if (threadIdx.x == 0) out[0] = s;
}
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int W = 800, H = 528, S = 3;
  const size_t per = (size_t)W * H + 256;
  std::vector<uint8_t> h(per * S);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 7 + (i >> 9)) & 0xff);
  uint8_t* d; CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncFn enc = (EncFn)fp;
  CUtensorMap m2, m3;
  { const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; const cuuint64_t st[1] = {(cuuint64_t)W}; const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    printf("encode2d -> %d\n", (int)enc(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  { const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S}; const cuuint64_t st[2] = {(cuuint64_t)W, (cuuint64_t)per}; const cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
    printf("encode3d -> %d\n", (int)enc(&m3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)); }
  unsigned* out; CK(cudaMalloc(&out, 64)); unsigned ho = 0;
  const int x = 48, y = 11;
  unsigned ref = 0; for (int r = 0; r < 32; ++r) for (int c = 0; c < 32; ++c) ref += h[(size_t)(y + r) * W + x + c];
  unsigned ref48 = 0; for (int r = 0; r < 32; ++r) for (int c = 0; c < 48; ++c) ref48 += h[(size_t)(y + r) * W + x + c];
  const char* w = getenv("TMA_PROBE"); const char v = w ? w[0] : 'a';
  if (v == 'a') k_a<<<1, 32>>>(m2, x, y, out);
  if (v == 'b') k_b<<<1, 32>>>(m3, x, y, 0, out);
  if (v == 'c') k_c<<<1, 32>>>(m2, x, y, out);
  if (v == 'd') k_d<<<1, 32>>>(d + (size_t)y * W + x, W, out);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&ho, out, 4, cudaMemcpyDeviceToHost);
  printf("variant %c: %s  sum %u (expected %u / 48-wide %u)\n", v, cudaGetErrorString(e), ho, ref, ref48);
  return e == cudaSuccess ? 0 : 2;
}
