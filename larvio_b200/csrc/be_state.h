// Back-end (LarVio) device state: StateServer + MapServer (larvio.h:99-142,199; imu_state.h:29-148;
// feature.hpp:34-250) as fixed-capacity SoA tables, one slice per sequence, FP64 throughout.
#pragma once
#include "lvb_internal.h"

#define BE_DMAX_PAD 640      // >= any state dimension: 46 + 6*64 + 3*64
#define BE_LEG_MAX 46        // legacy error-state size with IMU-intrinsic calibration (larvio.cpp:158-161); 22 without (LvbBackEnd::LEG)

// ---- core[s][BE_CORE] doubles
enum {
  C_TIME = 0, C_Q = 1, C_P = 5, C_V = 8, C_BG = 11, C_BA = 14, C_RIC = 17 /*R_imu_cam0 row-major*/, C_TCI = 26,
  C_TD = 29, C_DT = 30, C_FNOW_P = 31, C_FNOW_V = 34, C_FOLD_P = 37, C_FOLD_V = 40, C_OLD_Q = 43, C_OLD_P = 47,
  C_OLD_V = 50, C_GYRO_OLD = 53, C_ACC_OLD = 56, C_TRACK_RATE = 59, C_TAKEOFF = 60, C_LAST_ZUPT = 61,
  C_TG = 64, C_AS = 73, C_MA = 82 /*IMU intrinsics Tg, As, Ma row-major (larvio.cpp:129-155); T1..M2 are their entries*/, BE_CORE = 96
};
// ---- icore[s][BE_ICORE] ints
enum {
  I_ID = 0, I_NEXT_ID = 1, I_NWIN = 2, I_GRAVITY = 3, I_FIRST = 4, I_FEJ = 5, I_HAVE_OLD = 6, I_DIM = 7,
  I_ZUPT = 8, I_OK = 9 /*processFeatures return of this frame*/, I_CONSUMED = 10, I_ROWS = 11 /*stacked rows*/,
  I_R = 12 /*rows after compression*/, I_NUSED = 13, I_RAWROWS = 14, I_ERR = 15, I_RM0 = 16, I_RM1 = 17, I_NRM = 18,
  I_DO_PRUNE = 19, I_ZUPT_EVENTS = 20, I_UPDATES = 21, I_NF = 22 /*EKF-SLAM features in state*/, I_REMAP = 23, I_NEWDIM = 24,
  I_NNEW = 25 /*new SLAM features accepted this frame*/, I_NCAND = 26, I_RO = 27 /*rows of H_o before the new-feature rows*/,
  I_NC = 28 /*structurally nonzero columns of the stacked Jacobian (kmap)*/,
  I_NNUI = 29 /*Schmidt nuisance states in the covariance (use_schmidt, larvio.cpp:2351-2358)*/,
  I_NEWNUI = 30 /*bit k: window slot I_RM0+k becomes a nuisance state in this prune (:2569-2613)*/, BE_ICORE = 32,
  BE_NUI_MAX = 16 /*capacity of the nuisance block*/, BE_NUI_BASE = 64 /*ft_anchor >= BE_NUI_BASE: nuisance state ft_anchor - BE_NUI_BASE*/,
  BE_GRID_OOR = 256, BE_GRID_OOR_NEG = 64
};
// ---- win[s][slot][BE_WIN] doubles (IMUState_Aug)
enum { W_TIME = 0, W_DT = 1, W_Q = 2, W_P = 6, W_PFEJ = 9, W_RIC = 12, W_TCI = 21, W_QCAM = 24, W_PCAM = 28, BE_WIN = 32 };

struct LvbBackEnd {
  int S, N;            // sequences, message capacity per sequence
  int LEG;             // legacy error-state size: 22, or 46 with calib_imu_instrinsic
  int Wcap;            // window capacity (sw_size + 1)
  int T;               // feature-table capacity per sequence
  int LD;              // leading dimension of P / row length of stacked Jacobians (>= Dmax, multiple of 8)
  int Dmax;
  int RAWMAX;          // raw (unprojected) Jacobian rows per sequence per pass
  int RMAX;            // stacked (projected, gated) rows per sequence per pass
  int imu_cap;         // IMU samples per sequence per call
  int chol_cap;        // largest innovation system factorised out of shared memory (packed triangle); larger ones use the global-memory kernel
  int NFmax;           // EKF-SLAM feature capacity (max_features_in_one_grid * grid cells), 0 = pure MSCKF
  int IDP;             // state columns per EKF-SLAM feature: 1 (inverse depth) or 3 (x/z, y/z, 1/z in the anchor camera), larvio.cpp:270-274
  int LDS;             // leading dimension of Sm (rows of the stacked H_o can exceed the state dimension in hybrid mode)
  int grid_rows, grid_cols, max_per_cell;
  double* core; int* icore;
  long long* win_id; double* win;
  double* P[2]; int pcur;                     // ping-pong covariance [S][LD][LD], row-major
  unsigned long long* ft_id; int* ft_flags; double* ft_pos; unsigned long long* ft_mask; double* ft_obs;  // [S][T]...
  int* ft_action; int* ft_rowofs; int* ft_nrows; int* ft_accept; unsigned long long* ft_usemask;
  // EKF-SLAM bookkeeping per table slot: invDepth, corrected anchor observation (x,y) - with 3-D inverse depth the triple
  // (ft_oa[0], ft_oa[1], ft_inv) IS invParam = (x/z, y/z, 1/z) (feature.hpp:231) -, anchor window slot, first-estimate position;
  // speculative triangulation results of this frame; state order list
  double* ft_inv; double* ft_oa; int* ft_anchor; double* ft_pfej; double* ft_spec; int* fs_slot; int* cmap; int* cand;
  // occupancy of the reference's grid_map cells whose code lies OUTSIDE [0, rows*cols): [S][BE_GRID_OOR], index = code + BE_GRID_OOR_NEG.
  // updateGridMap (larvio.cpp:3355-3357) empties the rows*cols cells only; a cell created by an observation beyond the image
  // border (row == grid_rows, col < 0, ...) is never emptied, so these counts persist for the life of the handle
  int* grid_oor;
  // Schmidt nuisance states (use_schmidt: 1): poses that left the window while they anchored SLAM features stay BEHIND the feature
  // block of the covariance, frozen (larvio.cpp:2351-2358, 2569-2613).  NUI = capacity (0 without use_schmidt); nui_win[S][NUI][BE_WIN]
  // = the IMUState_Aug copies (nui_imu_states), nui_cnt[S][NUI] = features each one still anchors (nui_features[id].size()),
  // nui_P[S][(6 NUI)^2] = the block an update must leave untouched (:1579-1589)
  int NUI; double* nui_win; int* nui_cnt; double* nui_P;
  double* ft_gamma;                           // [S][T] last gating statistic of each slot (diagnostics)
  // map points for getStableMapPointPositions / getActiveeMapPointPositions (larvio.h:86-87): [which][S][PCAP], which 0 = SLAM
  // features that left the state (lost_slam_features, larvio.cpp:3342), 1 = features in the state at the end of a step
  // (active_slam_features, :455-458); both accumulate until lvb_get_points reads and clears them, like the reference's getters
  unsigned long long* pts_id; double* pts_xyz; int* pts_n; int* pts_drop; int PCAP;
  double* Hnew;                               // [S][64 * IDP][LD + 4]: per new state column (H_1 row | H_2 row [3] | r_1), features added this frame
  double* Hraw; double* rraw;                 // [S][RAWMAX][LD], [S][RAWMAX]
  double* Hs; double* rs;                     // stacked, COLUMN-major [S][LD cols][RMAX rows], [S][RMAX]
  double* Tm;                                 // H*P   [S][RMAX? -> Dmax rows][LD]   (rows <= Dmax after compression)
  double* Sm;                                 // [S][LD][LD]
  double* zvec; double* dx;                   // [S][LD]
  int* kmap;                                  // [S][LD] ascending list of the nonzero columns of the current stacked Jacobian (I_NC entries)
  LvbImu* imu; int* n_imu;                    // per-call IMU staging [S][imu_cap]
  LvbFeature* msg_in; int* msg_in_n; double* msg_in_t; uint8_t* msg_in_valid;   // host-provided messages
  double chi2[100];
  unsigned long long* stats;                  // shared with the front end's counter block (fe.stats)
  // host pinned
  LvbImu* pin_imu; int* pin_n_imu; int* pin_icore; LvbFeature* pin_feat; int* pin_feat_n; double* pin_feat_t; uint8_t* pin_valid;
};
