#include "lvb_internal.h"
template <typename T>
static int dalloc(LvbHandle* h, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  LVB_CUDA(cudaMalloc(&q, bytes));
  LVB_CUDA(cudaMemsetAsync(q, 0, bytes, h->stream));
  h->allocs.push_back(q);
  *p = (T*)q;
  return LVB_OK;
}
#define DA(ptr, n) do { int rc_ = dalloc(h, &(ptr), (size_t)(n)); if (rc_ != LVB_OK) return rc_; } while (0)
int fe_alloc(LvbHandle* h) {
  LvbFrontEnd& fe = h->fe;
  const size_t S = fe.S, N = fe.N, npx = (size_t)fe.W * fe.H;
  DA(fe.img_in, S * npx);
  DA(fe.lut, S * 64 * 256);
  for (int k = 0; k < 2; ++k) { DA(fe.pyr[k], S * fe.L.bytes_per_seq); DA(fe.blur[k], S * npx); }
  return LVB_OK;
}
