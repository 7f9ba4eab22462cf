// csrc/ekf_common.cuh — device-side data layout shared by all kernels of libekf_b200.
//
// HBM layout of one filter (all fp64 unless noted):
//   mu      [n_cap]            state vector, reference order (vslamRansac.hpp:27-92)
//   Sigma   [n_cap x ld]       covariance, ROW-major, ld = n_cap rounded up to 8 doubles (64 B rows)
//   SigmaB  [n_cap x ld]       second buffer; removeFeature compacts into it and swaps
//   W       [(n_cap+1) x ldw]  W = Sigma H^T (n x k), then V = W L^-T in place; row n holds z-h
//   Sm      [k_cap x lds]      innovation covariance S, then its Cholesky factor L (lower)
//   feature table (SoA, capacity N_cap): pos, coding, flags, counters, h, z, compact H (2 x 13),
//           2x2 S blocks, u8 templates (patch / matching_patch)
//   frame   [h x stride] u8
// Sizes known only on the device (matched rows, inlier counts ...) live in DevCtl; every kernel
// downstream reads them there and exits early, so a whole step is enqueued without host syncs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

// Function attributes (the > 48 KB dynamic shared-memory opt-in) belong to a device, not to the process: a handle on a
// second GPU of the same process needs its own cudaFuncSetAttribute.  One bit per device ordinal, set once the attribute
// call SUCCEEDED on that device; safe against concurrent first launches (setting an attribute twice is harmless).
struct PerDeviceOnce {
  std::atomic<unsigned long long> bits{0};
  // returns cudaSuccess when the attribute is (now) set on the current device
  template <class F>
  cudaError_t ensure(F&& set_attr) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (bits.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = set_attr();
    if (e == cudaSuccess) bits.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

#define EKF_CAM 14  // STATE_DIM (vslamRansac.cpp:22)

struct CamParams {
  double fx, fy, u0, v0, k1, k2, k3, p1, p2;
};

// Scalars of ekf_config the kernels need.
struct DevCfg {
  CamParams cam;
  double Vmax[6];  // diag(sigma_v^2, sigma_w^2) (vslamRansac.cpp:194-200)
  double sigma_pixel_2;
  double th_low;       // li_threshold_factor * sigma_pixel (vslamRansac.cpp:968)
  double th_hi;        // vslamRansac.cpp:1066
  double ransac_p;
  double linearity_threshold;
  double rho_0, sigma_rho_0;
  double T_camera;      // fraction of dT used for the blur pose (ConfigVSLAM.cpp:39)
  float ncc_threshold, search_clamp, sigma_size_f, quality_ratio;
  int window, sigma_pixel, nhyp0, forsePlane, abs_int_quirk;
  int tstride;  // bytes per template record (window^2 rounded up to 16)
  int kernel_min_size;  // Patch::blur threshold (vslamRansac.cpp:148); >= 100000: motion blur off
};

// Device-resident control block.
struct DevCtl {
  double mu_cam_new[13];  // Predict_State output, committed by k_predict_features
  double cam_old[7];      // r,q of mu before the low-innovation update (vslamRansac.cpp:1069-1072)
  int m_innov;            // features in innovation after predict
  int n_matched;          // after matching (vslamRansac.cpp:984)
  int n_li, n_hi;
  int k_rows;             // rows of the stacked update being processed
  int ransac_hyps;
  int n_visible;          // vslamRansac.cpp:1301-1303
  int n_remove;           // features flagged by update_quality_index / rho <= 0
  unsigned int ticket;    // last-block-done counter (self-resetting)
  int chol_fail;          // non-positive pivot seen
  int blur_count;         // templates blurred by the last predict
  int blur_too_large;     // a blur kernel exceeded the supported 256 x 256
  int rs_sel, rs_count;   // cluster RANSAC (k_ransac_cluster): feature picked by CTA 0, inlier count summed over the CTAs
};

// Feature table (structure of arrays, device pointers).
struct FeatTab {
  int* pos;         // position_in_state
  int* coding;      // 0 inverse depth, 1 XYZ
  int* innov;       // isInInnovation
  int* li;          // isInLi
  int* hi;          // isInHi
  int* removef;     // removeFlag
  int* n_tot;
  int* n_find;
  int* real_index;
  int* pos_in_z;
  int* sel;         // compacted list of selected features (innovation / Li / Hi), in patch order
  float* center;    // 2 per feature
  float* quality;
  float* last_ncc;
  double* z;        // 2 per feature
  double* h;        // 2 per feature
  double* Hc;       // 26 per feature: 2 x 13 row-major, cols [0,7) camera, [7,13) feature
  double* S2;       // 4 per feature: 2x2 block of St
  uint8_t* patch;   // cfg.tstride bytes per feature (w*w used)
  uint8_t* mpatch;  // cfg.tstride bytes per feature (matching_patch)
};

// Feature table of filter b inside a batch of tables laid out [filter][Ncap] per field.
__host__ __device__ __forceinline__ FeatTab feattab_slice(const FeatTab& t, int b, int Ncap, int tstride) {
  FeatTab o;
  const size_t k = (size_t)b * Ncap;
  o.pos = t.pos + k; o.coding = t.coding + k; o.innov = t.innov + k; o.li = t.li + k; o.hi = t.hi + k;
  o.removef = t.removef + k; o.n_tot = t.n_tot + k; o.n_find = t.n_find + k; o.real_index = t.real_index + k;
  o.pos_in_z = t.pos_in_z + k; o.sel = t.sel + k; o.center = t.center + 2 * k; o.quality = t.quality + k;
  o.last_ncc = t.last_ncc + k; o.z = t.z + 2 * k; o.h = t.h + 2 * k; o.Hc = t.Hc + 26 * k; o.S2 = t.S2 + 4 * k;
  o.patch = t.patch + k * tstride; o.mpatch = t.mpatch + k * tstride;
  return o;
}

// Compact per-update tables of the selected features, in ft.sel order (k_blk_prep): what the block kernels of the stacked update
// read of the feature table, without the sel -> pos / coding -> H chain of dependent loads.  H == nullptr: not prepared (the kernels
// then go through the feature table).
struct BlkTab {
  const double* H;    // [cnt][26]
  const double* zmh;  // [cnt][2]  z - h
  const int* pos;     // [cnt]
  const int* nd;      // [cnt]  7 + 3 (XYZ) or 7 + 6 (inverse depth)
};

// Flag words of the resident-chain schedule (k_chain_factor): one word per update block and kind, written once per stacked update
// with that update's token (no reset needed), plus self-resetting tickets for "last CTA of the kernel" detection.
struct ChainFlags {
  unsigned int* gather;   // [blocks]  W'_b is complete (last CTA of its gather)
  unsigned int* sg;       // [blocks]  -G_b G_b^T, G_b and gy are complete (last CTA of k_blk_Sg)
  unsigned int* fact;     // [blocks]  L_b, D_b, y_b are complete (k_chain_factor)
  unsigned int* tickets;  // [0] gather, [1] Sg
  unsigned int token0;    // block b of this update publishes token0 + b
};
__device__ __forceinline__ void chain_publish(unsigned int* flag, unsigned int token) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(flag), "r"(token) : "memory");
}
// bounded wait of ONE thread for *flag == token; returns false after ~seconds (the caller reports instead of hanging)
__device__ __forceinline__ bool chain_wait(const unsigned int* flag, unsigned int token) {
  unsigned int seen = 0;
  for (long long spins = 0; spins < 2000000ll; ++spins) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(flag) : "memory");
    if (seen == token) return true;
    __nanosleep(40);
  }
  return false;
}
// end of a producer kernel: the last CTA to get here publishes `token` (every thread of the CTA must call)
__device__ __forceinline__ void chain_publish_last_cta(unsigned int* ticket, unsigned int* flag, unsigned int token) {
  __threadfence();   // EVERY thread: its own stores are performed before the CTA's ticket is taken (a fence by thread 0 alone does not cover the other warps' stores in flight)
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x * gridDim.y - 1) { *ticket = 0u; __threadfence(); chain_publish(flag, token); }
  }
}

// device-side view of the peer mappings, passed by value to the kernels that push panels to the peers
struct P2PView {
  double* w[8];                 // destination panel (Wbuf[b % 3]) on every rank
  double* spart[8];             // partial-S slots on every rank
  unsigned long long* flags[8]; // flag words on every rank
  int rank, world;
  unsigned long long epoch;
};

struct FrameView {
  const uint8_t* px;
  int w, h, stride;
};

#define EKF_CUDA_CHECK(expr)                                                     \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ekf_fail_cuda(h, _e, #expr, __FILE__, __LINE__); \
  } while (0)

// 13-entry index list of the state entries a feature's measurement depends on:
// camera r,q (0..6) and the feature block.  XYZ features use 10 entries.
__device__ __forceinline__ int ekf_idx13(int c, int pos) { return c < 7 ? c : pos + (c - 7); }

// Ordered compaction of flag[0..N) into list / pos_in_z by ONE block (blockDim multiple of 32).
// Returns the count to every thread.  list may be null.
static __device__ __noinline__ int block_compact(const int* flag, int N, int* list, int* pos_in_z) {
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int start = 0; start < N; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const int f = (i < N) ? (flag[i] != 0) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += warp_cnt[w];
    off += __popc(m & ((1u << lane) - 1u));
    if (f) {
      if (list) list[off] = i;
      if (pos_in_z) pos_in_z[i] = 2 * off;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < nw; ++w) tot += warp_cnt[w];
      base_s += tot;
    }
    __syncthreads();
  }
  return base_s;
}

