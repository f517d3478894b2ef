// Stand-alone probe of the TMA tile load the LK kernel uses (3-D u8 tensor, box 32x32x1, per-warp mbarrier):
// which way of handing the descriptor to the kernel works on this driver.   nvcc -arch=sm_100a tools/tma_probe.cu -o /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
struct Nested { CUtensorMap m[4]; int pad[8]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ void load_and_sum(const CUtensorMap* map, int x, int y, int z, unsigned* out, int fence_tensormap) {
  __shared__ alignas(128) uint8_t tile[4][1024 + 128];
  __shared__ unsigned long long bar[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned b = smem_u32(&bar[warp]), d = smem_u32(tile[warp]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    if (fence_tensormap) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(map) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(d), "l"(map), "r"(x + warp), "r"(y), "r"(z), "r"(b) : "memory");
  }
  unsigned ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(0) : "memory");
  unsigned s = 0;
  for (int i = lane; i < 1024; i += 32) s += tile[warp][i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[warp] = ok ? s : 0xdeadbeefu;
}
__global__ void k_direct(const __grid_constant__ CUtensorMap m, int x, int y, int z, unsigned* out) { load_and_sum(&m, x, y, z, out, 0); }
__global__ void k_nested(const __grid_constant__ Nested n, int level, int x, int y, int z, unsigned* out) { load_and_sum(&n.m[level], x, y, z, out, 0); }
__global__ void k_global(const CUtensorMap* m, int x, int y, int z, unsigned* out) { load_and_sum(m, x, y, z, out, 1); }
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
  printf("entry point %p query %d\n", fp, (int)q);
  EncFn enc = (EncFn)fp;
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S};
  const cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)per};
  const cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)r);
  unsigned* out; CK(cudaMalloc(&out, 64)); unsigned ho[4];
  const int x = 37, y = 11, z = 2;
  unsigned ref[4];
  for (int w = 0; w < 4; ++w) { unsigned s = 0; for (int r2 = 0; r2 < 32; ++r2) for (int c = 0; c < 32; ++c) s += h[per * z + (size_t)(y + r2) * W + x + w + c]; ref[w] = s; }
  auto report = [&](const char* name) -> int {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-8s FAILED: %s\n", name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
    printf("%-8s sums %u %u %u %u  (expected %u %u %u %u) %s\n", name, ho[0], ho[1], ho[2], ho[3], ref[0], ref[1], ref[2], ref[3],
           (ho[0] == ref[0] && ho[3] == ref[3]) ? "OK" : "MISMATCH");
    return 0;
  };
  const char* which = getenv("TMA_PROBE");
  if (!which || which[0] == 'd') { k_direct<<<1, 128>>>(m, x, y, z, out); if (report("direct")) return 2; }
  if (!which || which[0] == 'n') { Nested n; memset(&n, 0, sizeof(n)); n.m[2] = m; k_nested<<<1, 128>>>(n, 2, x, y, z, out); if (report("nested")) return 3; }
  if (!which || which[0] == 'g') { CUtensorMap* dm; CK(cudaMalloc(&dm, sizeof(m))); CK(cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice)); k_global<<<1, 128>>>(dm, x, y, z, out); if (report("global")) return 4; }
  // negative / out-of-bounds coordinates (zero fill)
  if (!which) { k_direct<<<1, 128>>>(m, -5, -7, 0, out); cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost); printf("oob      %s sums %u %u\n", cudaGetErrorString(e), ho[0], ho[1]); }
  return 0;
}
