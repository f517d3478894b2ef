// Internal layout of a larvio_b200 handle: every per-sequence container of the reference
//  as fixed-capacity SoA arrays resident in HBM for the whole run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/larvio_b200.h"

#define LVB_PAD 24          // pyramid level padding (>= patch_size 21 + 1, see App. C-10)
#define LVB_MAX_LEVELS 4
#define LVB_DESC_BYTES 32

struct LvbLevel {
  int w, h;        // image size of this level
  int pitch;       // bytes per padded row
  int rows;        // padded rows
  size_t offset;   // byte offset of padded row 0 inside one sequence's pyramid block
};

struct LvbPyramidLayout {
  int n_levels;
  LvbLevel lv[LVB_MAX_LEVELS];
  size_t bytes_per_seq;
};

// pointer to pixel (0,0) of level l of sequence s in pyramid block `base`
__host__ __device__ inline const uint8_t* lvb_level_origin(const uint8_t* base, const LvbPyramidLayout& L,
                                                           int s, int l) {
  return base + (size_t)s * L.bytes_per_seq + L.lv[l].offset + (size_t)LVB_PAD * L.lv[l].pitch + LVB_PAD;
}
__host__ __device__ inline uint8_t* lvb_level_origin(uint8_t* base, const LvbPyramidLayout& L, int s, int l) {
  return base + (size_t)s * L.bytes_per_seq + L.lv[l].offset + (size_t)LVB_PAD * L.lv[l].pitch + LVB_PAD;
}

struct LvbCamera {
  double fx, fy, cx, cy;
  double dist[4];
  int model;
};

// ---- front-end state (image_processor.h:260-325), one slot set per sequence ----
struct LvbFrontEnd {
  int S, W, H, N;                 // sequences, image size, per-sequence track capacity
  LvbPyramidLayout L;
  uint8_t* img_in;                // [S][H][W]  staging of the raw input image
  uint8_t* lut;                   // [S][64][256] CLAHE LUTs
  uint8_t* pyr[2];                // ping-pong padded pyramids (prev/curr)
  uint8_t* blur[2];               // ping-pong 7x7-blurred L0 planes [S][H][W]
  int cur;                        // index of the CURRENT pyramid (host-side toggle)
  // persistent tracks
  float2* prev_pts; float2* curr_pts; float2* init_pts;       // [S][N]
  unsigned long long* ids; int* lifetime; uint8_t* desc;      // [S][N], [S][N], [S][N][32]
  int* n_tracks;                                              // [S]
  float2* new_pts; int* n_new;                                // [S][N], [S]
  // per-sequence scalars (device)
  int* image_state;               // 1 FIRST, 2 SECOND, 3 OTHER
  unsigned long long* next_id;
  double* last_pub_time; double* prev_img_time; double* curr_img_time;
  // per-frame host->device inputs
  float* Hmat;                    // [S][9] K R K^-1 (float, row-major)
  int* active;                    // [S] bFirstImg gate result for this frame
  double* t_img;                  // [S]
  // per-frame control flags produced by fe_frame_begin
  int* do_first; int* do_second; int* do_other; int* do_publish;
  // chain scratch (two chains: 0 = tracked, 1 = new)
  float2* ch_prev[2]; float2* ch_curr[2]; float2* ch_back[2];  // [S][N]
  float2* ch_uprev[2]; float2* ch_ucurr[2];
  uint8_t* ch_status[2];          // [S][N]
  int* ch_perm[2];                // [S][N] alive slot indices (order preserving)
  int* ch_n[2];                   // [S] alive count
  int* ch_fail[2];                // [S] chain aborted flag
  uint8_t* ch_desc[2];            // [S][N][32] descriptors computed this frame (prev for new chain)
  // detector scratch
  float* eig;                     // [S][H][W]
  uint8_t* mask;                  // [S][H][W]
  float* eig_max;                 // [S]
  unsigned long long* cand;       // [S][cand_cap] packed (value bits<<32 | ~index)
  int* n_cand; int cand_cap;
  int* want;                      // [S]
  // outputs
  LvbFeature* msg;                // [S][N]
  int* msg_n;                     // [S]
  uint8_t* has_msg;               // [S]
  double* msg_t;                  // [S]
};

struct LvbBackEnd;   // be_state.h

struct LvbHandle {
  LvbConfig cfg;
  int S;
  int device;
  cudaStream_t stream;
  LvbFrontEnd fe;
  LvbBackEnd* be;
  long long launches;
  // host-side per-sequence bookkeeping (mirrors ImageProcessor members that only the host needs)
  std::vector<uint8_t> h_first_img;      // bFirstImg
  std::vector<double> h_prev_img_time;
  std::vector<uint8_t> h_have_prev;
  // pinned staging
  uint8_t* pin_images; size_t pin_images_bytes;
  float* pin_H; int* pin_active; double* pin_t;
  LvbFeature* pin_msg; int* pin_msg_n; uint8_t* pin_has;
  std::vector<void*> allocs;
};

extern thread_local std::string g_lvb_err;
int lvb_set_err(int code, const char* fmt, ...);

#define LVB_CUDA(x)                                                                        \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess)                                                                 \
      return lvb_set_err(LVB_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
  } while (0)

#define LVB_LAUNCH_CHECK(h)                                                                \
  do {                                                                                     \
    (h)->launches++;                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                   \
    if (e_ != cudaSuccess)                                                                 \
      return lvb_set_err(LVB_E_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

// ---- stage launchers (each file owns its kernels) ----
int fe_build_pyramid(LvbHandle* h, const uint8_t* d_images /*[n][H][W]*/, int n, uint8_t* pyr, uint8_t* blur);
int fe_lk_launch(LvbHandle* h, const uint8_t* pyrA, const uint8_t* pyrB, int n_seq, int stride,
                 const float2* ptsA, const int* perm, const int* n_pts, const float2* init,
                 const float* Hmat, float2* out, uint8_t* status, int gate_mode, const float2* ref);
