// csrc/ekf_handle.h — the opaque ekf_handle of include/ekf_b200.h (internal to libekf_b200).
#pragma once
#include <string>
#include <vector>

#include "../../include/ekf_b200.h"
#include "ekf_kernels.h"

struct DeletedPatch {  // Patch archived by removeFeature (vslamRansac.cpp:394-404)
  int real_index;
  double XYZ_pos[3];
  double cov_4_delete[9];
};

struct ekf_handle {
  ekf_config cfg;
  DevCfg dcfg;
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  int Ncap = 0, ncap = 0, ld = 0, N = 0, n = EKF_CAM;
  double *mu = nullptr, *muB = nullptr, *Sigma = nullptr, *SigmaB = nullptr;
  double *W = nullptr, *nu = nullptr, *Lb = nullptr, *Dinv = nullptr, *Dblk = nullptr, *yb = nullptr, *delta = nullptr, *mu_i = nullptr;
  int *xyz_flag = nullptr, *xyz_rmap = nullptr, *xyz_pos = nullptr, *xyz_coding = nullptr;
  double *xyz_y = nullptr, *xyz_J = nullptr;
  int *cand = nullptr, *map_dev = nullptr, *keep_dev = nullptr, *newpos_dev = nullptr, *gemm_counters = nullptr;
  DevCtl* ctl = nullptr;
  DevCtl* ctl_host = nullptr;   // pinned mirror for the two mid-step read-backs (n_li, n_hi)
  FeatTab ft{}, ftB{};
  uint8_t* frame = nullptr;
  size_t frame_cap = 0;
  uint8_t* raw = nullptr;   // full-resolution / colour staging for captureNewFrame's resize + BGR2GRAY
  size_t raw_cap = 0;
  FrameView fv{nullptr, 0, 0, 0};
  // host frames are uploaded on their own stream so that the copy runs beside predict (which does not read pixels); every
  // reader of the frame first makes the filter's stream wait for ev_frame (frame_ready, ekf_api.cu)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_frame = nullptr, ev_prev = nullptr;
  bool frame_pending = false;
  EkfTensorMap frame_map{};   // TMA tensor map of `frame` for the matcher's window staging (re-encoded when the frame view changes)
  uint32_t* picks_dev = nullptr;
  int picks_cap = 0;
  double* out_dev = nullptr;   // packed step record (device)
  double* out_host = nullptr;  // pinned mirror
  size_t out_bytes = 0;
  double dT = 1.0, old_ts = -1.0;
  int patchnumbre = 1, noise_cov_factor = 0;
  bool predicted = false, have_frame = false;
  bool cam_cache_ok = false;   // out_host holds the camera state and its 14 x 14 covariance block as of the last update (getState / getSigma)
  int lower_only = 0;
  // corner detector scratch (ekf_detect.cu), sized to the frame on first use
  uint8_t* det_mask = nullptr; float* det_eig = nullptr; unsigned long long* det_keys = nullptr; int* det_counters = nullptr;
  float* det_xy = nullptr; size_t det_cap = 0;
  // look-ahead pipeline of the stacked update (ekf_api.cu::stacked_update_lookahead)
  cudaStream_t gemm_stream = nullptr, corr_stream = nullptr;
  double* Wbuf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // Wbuf[0] == W; [4], [5]: V_b of the chain-short schedule; [6], [7]: hot rows of W'' (S look-ahead)
  double* Gbuf = nullptr;
  // chain-short schedule (ekf_api.cu::stacked_update_chain_short): second set of factor outputs (V_{b-1} reads one set while
  // factor_b writes the other), delta ping-pong, gy = G_b y_{b-1}, the stream of the n-row solves and per-block events
  double *Dinv2 = nullptr, *Dblk2 = nullptr, *yb2 = nullptr, *delta1 = nullptr, *delta2 = nullptr, *gy = nullptr;
  cudaStream_t v_stream = nullptr, gather_stream = nullptr;
  cudaEvent_t ev_F[2] = {nullptr, nullptr}, ev_corr2[2] = {nullptr, nullptr}, ev_dd[2] = {nullptr, nullptr}, ev_A = nullptr;
  double *bt_H = nullptr, *bt_zmh = nullptr, *Sgbuf = nullptr;   // BlkTab storage (k_blk_prep), -G G^T of the current block (k_blk_Sg)
  int *bt_pos = nullptr, *bt_nd = nullptr;
  ushort2* tile_order = nullptr;          // [blocks][T (T + 1) / 2]: see k_blk_tile_order
  int* tile_nhot = nullptr;               // [blocks]
  unsigned int* tile_counters = nullptr;  // [blocks]
  int tile_T_cap = 0, tile_blk_cap = 0;
  int split_dd = 0;       // chain-short only (EKF_SPLIT_DD): 1 = tile list, hot tiles first, the next gather gated on them; 2 = tile list only
  // resident-chain schedule (EKF_SCHED=2): flag words [3][blocks] + tickets, token sequence, stream of the resident factor CTA
  unsigned int* chain_flags = nullptr;
  unsigned int chain_seq = 0;
  cudaStream_t chain_stream = nullptr;
  cudaEvent_t ev_chain = nullptr;
  // S look-ahead of the chain-short schedule (EKF_S_LOOKAHEAD): G2 = H_b V_{b-2}, -G2 G2^T per block parity, events
  double *G2buf = nullptr, *Sg2buf[2] = {nullptr, nullptr};
  cudaEvent_t ev_mini[2] = {nullptr, nullptr}, ev_Sg2[2] = {nullptr, nullptr};
  int v_after_dd = 0;      // chain-short: V_b waits for the downdate of block b-1 (EKF_V_AFTER_DD)
  int gather_hp = 0;       // chain-short: the gather of W'_{b+1} on a high-priority stream (EKF_GATHER_HP)
  int s_lookahead = 0;
  int dd_release = 2;      // EKF_SCHED=3 only (EKF_DD_RELEASE): which kernel of block b releases the downdate of block b-1
  bool prelaunched = false;   // block tables + first two gathers were started before the n_li read-back (chain_short_prelaunch)
  int prelaunch_on = 1;       // EKF_PRELAUNCH=0 switches that off
  int sched = 1;          // 1: chain-short (default for pipe_small <= n < lookahead), 0: factor-beside-downdate (EKF_SCHED)
  cudaEvent_t ev_gather[3] = {nullptr, nullptr, nullptr}, ev_V[3] = {nullptr, nullptr, nullptr}, ev_fork = nullptr, ev_join = nullptr, ev_S = nullptr, ev_G = nullptr, ev_corr = nullptr;
  int pipe_small = 1000;  // minimum state dimension for the factor-beside-downdate schedule (0 = never)
  int lookahead = 6000;   // minimum state dimension for the look-ahead pipeline (0 = never)
  // row-block partitioned update across ranks (ekf_dist.cu): NCCL communicator of this handle, or null
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  long long dist_bytes = 0;   // bytes this rank contributed to all-gathers so far
  // Peer-memory exchange of the partitioned look-ahead update (ekf_dist.cu::p2p_setup): the panels go to the peers
  // by NVLink stores from inside the producing kernels instead of NCCL calls.
  struct P2P {
    bool on = false;
    void* mapped[8][4] = {};              // cudaIpcOpenMemHandle results per peer (to close on detach)
    double* peerW[3][8] = {};             // peers' Wbuf[0..2] (own entry = local pointer)
    double* peerSpart[8] = {};            // peers' partial-S slots [world][EKF_UB * EKF_UB]
    unsigned long long* peerFlags[8] = {};  // peers' flag words: [0, 8) S epochs by writer rank, [8, 16) V epochs
    double* xs = nullptr;                 // local allocation: partial-S slots, then the flag words
    unsigned long long epoch = 0;         // one per update block, the same sequence on every rank
  } p2p;
  ekf_step_stats stats{};
  long long launches = 0;
  std::string err;
  std::vector<DeletedPatch> deleted;
  // per-kernel-class CUDA-event timing (ekf_set_profiling)
  bool prof_on = false;
  struct ProfRec { int cls; int nl; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[EKF_PROF_CLASSES] = {0};
  long long prof_launches[EKF_PROF_CLASSES] = {0};
  // host cache of the feature table (valid when cache_ok)
  bool cache_ok = false;
  std::vector<int> c_pos, c_coding, c_innov, c_li, c_hi, c_removef, c_ntot, c_nfind, c_real, c_posz;
  std::vector<float> c_center, c_quality, c_ncc;
  std::vector<double> c_z, c_h, c_Hc, c_S2;
  // host mirror maintained by add/remove (always valid)
  std::vector<int> m_pos, m_coding;
};


// ekf_dist.cu: in-place all-gather of a row-partitioned device buffer.  Rank r owns rows
// [r * rows_per_rank, (r + 1) * rows_per_rank) of `buf` (row_elems doubles per row).
int ekf_dist_allgather_rows(ekf_handle* h, double* buf, int rows_per_rank, size_t row_elems);
int ekf_dist_allreduce_sum(ekf_handle* h, double* buf, size_t count);
#define EKF_DIST_PAD_ROWS 512   // extra rows allocated for W / delta / Sigma so that world * rows_per_rank fits
