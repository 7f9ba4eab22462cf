// csrc/ekf_gemm.cu — K4d: covariance downdate  Sigma <- Sigma - V V^T  on the fp64 tensor pipe.
//
// the reference forms (I - K H) Sigma as a dense n x n x n product
// (vslamRansac.cpp:1060,1279).  With V = W L^-T (W = Sigma H^T, S = L L^T) the same update is the
// rank-k downdate Sigma - V V^T: 2 n^2 k flops, read + write of Sigma (16 n^2 bytes).
//
// sm_100a has no tcgen05 kind for fp64; the fp64 tensor path is mma.sync.m8n8k4.f64 (SASS
// DMMA.8x8x4).  Kernel shape: 128 x 64 output tile per CTA, 4 warps (2 x 2), each warp a 64 x 32
// sub-tile = 8 x 4 DMMA tiles (64 accumulator doubles per thread); K is consumed in 16-wide slabs
// staged global -> shared with cp.async (3 stages), rows padded to 20 doubles so the 8-row x 4-col
// fragment reads are bank-conflict free (row*20 mod 16 covers 0,4,8,12).
#include <cstdlib>

#include "ekf_kernels.h"

#define GT_N 64
#define GT_K 16
#define GT_LD 20
#define GT_THREADS 128

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// C[M x N] (ldc) -= A[M x K] (lda, K contiguous) * B[N x K]^T (ldb, K contiguous).
// K (kconst, or *kdev when kdev != nullptr) must be even; operands are zero-filled past K.
// The accumulators are initialised with the C tile (loaded while the first cp.async stages are in
// flight) and the B fragments are negated, so D = C + A (-B)^T needs no read in the epilogue.
// Two CTAs per SM (128 threads, 92 KB smem each) overlap one tile's C traffic with the other's DMMAs.
// lower_only: skip tiles entirely above the diagonal and mirror the strictly-lower elements into
// the upper triangle (C symmetric on input => symmetric on output).
template <int GT_M, int GT_STAGES, int MINB>
__global__ void __launch_bounds__(GT_THREADS, MINB)
k_gemm_nt_sub(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
              int M, int N, int kconst, const int* __restrict__ kdev, int lower_only, unsigned stagger_ns,
              const ushort2* __restrict__ tlist, const int* __restrict__ n_hot, unsigned int* hot_counter) {
  extern __shared__ __align__(16) double gsm[];
  const int K = kdev ? *kdev : kconst;
  if (K <= 0) return;
  // tlist != null (square tiles, lower_only): 1-D grid over a list of the lower-triangle tiles in which the tiles the NEXT gather
  // reads come first (k_blk_tile_order); a CTA that finishes one of those first *n_hot tiles bumps hot_counter, and the gather —
  // launched beside this kernel — starts as soon as the counter is full instead of after the whole downdate.
  int tm = blockIdx.y, tn = blockIdx.x;
  if (tlist) { const ushort2 t = tlist[blockIdx.x]; tm = t.x; tn = t.y; }
  const int m0 = tm * GT_M, n0 = tn * GT_N;
  if (lower_only && n0 > m0 + GT_M - 1) return;
  // De-synchronise the first wave.  All 4 x 148 CTAs of a launch start together, so they also RETIRE together (tile life ~ 20 us):
  // for most of the downdate no slot frees up, and a high-priority CTA of the gain chain launched meanwhile waits up to a tile
  // life for its registers.  With the tile list (1-D grid in dispatch order) the k-th group of 148 CTAs sleeps k * stagger_ns
  // first: retirements then come every stagger_ns, while the SM's DMMA pipe stays busy with the CTAs already running.
  if (stagger_ns > 0) {
    if (tlist) { if (blockIdx.x < 592u) { const unsigned k = (blockIdx.x / 148u) & 3u; if (k) __nanosleep(k * stagger_ns); } }
    else {
      const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x;
      if (lin < 296u && ((lin / 148u) & 1u)) __nanosleep(stagger_ns);
    }
  }
  double* As = gsm;                                   // [stages][GT_M][GT_LD]
  double* Bs = gsm + GT_STAGES * GT_M * GT_LD;        // [stages][GT_N][GT_LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;            // 2 x 2 warps, each (GT_M / 2) x 32
  constexpr int WR = GT_M / 2, WI = GT_M / 16;        // warp tile rows, 8-row DMMA tiles per warp
  const int g = lane >> 2, t4 = lane & 3;

  const int ktiles = (K + GT_K - 1) / GT_K;
  auto load_stage = [&](int kt, int stage) {
    const int k0 = kt * GT_K;
#pragma unroll
    for (int it = 0; it < (GT_M * 8) / GT_THREADS; ++it) {
      const int chunk = tid + it * GT_THREADS;
      const int row = chunk >> 3, cc = (chunk & 7) * 2;
      const int gr = m0 + row;
      const bool ok = (gr < M) && (k0 + cc < K);
      const double* src = A + (size_t)(ok ? gr : 0) * lda + (ok ? k0 + cc : 0);
      cp_async16(As + ((size_t)stage * GT_M + row) * GT_LD + cc, src, ok ? 16 : 0);
    }
#pragma unroll
    for (int it = 0; it < (GT_N * 8) / GT_THREADS; ++it) {
      const int chunk = tid + it * GT_THREADS;
      const int row = chunk >> 3, cc = (chunk & 7) * 2;
      const int gr = n0 + row;
      const bool ok = (gr < N) && (k0 + cc < K);
      const double* src = B + (size_t)(ok ? gr : 0) * ldb + (ok ? k0 + cc : 0);
      cp_async16(Bs + ((size_t)stage * GT_N + row) * GT_LD + cc, src, ok ? 16 : 0);
    }
  };
#pragma unroll
  for (int s = 0; s < GT_STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }
  // accumulators <- C tile
  double acc[WI][4][2];
#pragma unroll
  for (int i = 0; i < WI; ++i) {
    const int r = m0 + wm * WR + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn * 32 + j * 8 + 2 * t4;
      double2 v = make_double2(0.0, 0.0);
      if (r < M && c + 1 < N) v = *reinterpret_cast<const double2*>(C + (size_t)r * ldc + c);
      else if (r < M && c < N) v.x = C[(size_t)r * ldc + c];
      acc[i][j][0] = v.x; acc[i][j][1] = v.y;
    }
  }
  // warps whose 64 x 32 sub-tile lies outside the matrix, or (lower_only) entirely above the
  // diagonal, skip the tensor work (they still help with the loads): edge and diagonal tiles cost
  // half, which also removes most of the last partial wave.
  const bool warp_active = (m0 + wm * WR < M) && (n0 + wn * 32 < N) && !(lower_only && (n0 + wn * 32 > m0 + wm * WR + WR - 1));
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<GT_STAGES - 2>();
    __syncthreads();
    const int nk = kt + GT_STAGES - 1;
    if (nk < ktiles) load_stage(nk, nk % GT_STAGES);
    cp_async_commit();
    const double* as = As + (size_t)(kt % GT_STAGES) * GT_M * GT_LD + (size_t)(wm * WR + g) * GT_LD + t4;
    const double* bs = Bs + (size_t)(kt % GT_STAGES) * GT_N * GT_LD + (size_t)(wn * 32 + g) * GT_LD + t4;
    if (warp_active)
#pragma unroll
    for (int k4 = 0; k4 < GT_K / 4; ++k4) {
      double af[WI], bf[4];
#pragma unroll
      for (int i = 0; i < WI; ++i) af[i] = as[(size_t)i * 8 * GT_LD + k4 * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = -bs[(size_t)j * 8 * GT_LD + k4 * 4];
#pragma unroll
      for (int i = 0; i < WI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  // epilogue: store.  Thread holds rows g (+8 i), column pairs 2*t4 (+8 j).
#pragma unroll
  for (int i = 0; i < WI; ++i) {
    const int r = m0 + wm * WR + i * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn * 32 + j * 8 + 2 * t4;
      if (c >= N) continue;
      double* p = C + (size_t)r * ldc + c;
      if (!lower_only) {
        if (c + 1 < N) *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
        else *p = acc[i][j][0];
      } else {
        // only the lower triangle (col <= row) is authoritative; strictly-lower values are mirrored
        if (c + 1 <= r && c + 1 < N) {
          *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
          C[(size_t)c * ldc + r] = acc[i][j][0];
          if (c + 1 < r) C[(size_t)(c + 1) * ldc + r] = acc[i][j][1];
        } else if (c <= r) {
          *p = acc[i][j][0];
          if (c < r) C[(size_t)c * ldc + r] = acc[i][j][0];
        }
      }
    }
  }
  if (tlist && (int)blockIdx.x < *n_hot) {
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicAdd(hot_counter, 1u); }
  }
}

template <int GT_M, int GT_STAGES, int MINB>
static int launch_variant(cudaStream_t st, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int N,
                          int kconst, const int* kdev, int lower_only, unsigned stagger, const ushort2* tlist = nullptr, int n_tiles = 0, const int* n_hot = nullptr,
                          unsigned int* hot_counter = nullptr) {
  constexpr size_t smem = (size_t)GT_STAGES * (GT_M + GT_N) * GT_LD * sizeof(double);
  static PerDeviceOnce once;   // per device, not per process
  const cudaError_t ea = once.ensure([&] {
    return cudaFuncSetAttribute(k_gemm_nt_sub<GT_M, GT_STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  });
  if (ea != cudaSuccess) return (int)ea;
  dim3 grid((N + GT_N - 1) / GT_N, (M + GT_M - 1) / GT_M);
  if (tlist) grid = dim3(n_tiles, 1);
  k_gemm_nt_sub<GT_M, GT_STAGES, MINB><<<grid, GT_THREADS, smem, st>>>(C, ldc, A, lda, B, ldb, M, N, kconst, kdev, lower_only, stagger,
                                                                       tlist, n_hot, hot_counter);
  return 0;
}

// Tile shape: 64 x 64 at 4 CTAs / SM (32 accumulators per thread, 2-stage cp.async ring).  Against the
// 128 x 64 / 2 CTAs / SM variant (EKF_GEMM_TM=128) it packs the 148 x 4 CTA slots with almost no idle
// last wave (n = 3014, lower triangle: 0.47 -> 0.66 of the DGEMM peak) and keeps four independent CTAs
// per SM in flight (n = 11972: 0.69 -> 0.88).
int launch_gemm_nt_sub(cudaStream_t st, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int N,
                       int kconst, const int* kdev, int lower_only, int* counters, long long* launches,
                       const ushort2* tlist, int n_tiles, const int* n_hot, unsigned int* hot_counter) {
  (void)counters;
  if (M <= 0 || N <= 0) return 0;
  static int stagger = -1, force_tm = -1;
  if (stagger < 0) { const char* e = getenv("EKF_GEMM_STAGGER_NS"); stagger = e ? atoi(e) : 0; }
  if (force_tm < 0) { const char* e = getenv("EKF_GEMM_TM"); force_tm = e ? atoi(e) : 0; }
  const bool small_tile = force_tm ? (force_tm == 64) : true;
  if (tlist && !(small_tile && lower_only && M == N)) return (int)cudaErrorInvalidValue;   // the list enumerates 64 x 64 lower-triangle tiles
  const int rc = small_tile ? launch_variant<64, 2, 4>(st, C, ldc, A, lda, B, ldb, M, N, kconst, kdev, lower_only, (unsigned)stagger, tlist, n_tiles, n_hot, hot_counter)
                            : launch_variant<128, 3, 2>(st, C, ldc, A, lda, B, ldb, M, N, kconst, kdev, lower_only, (unsigned)stagger);
  if (rc) return rc;
  if (launches) *launches += 1;
  return 0;
}

bool gemm_uses_square_tiles() {
  const char* e = getenv("EKF_GEMM_TM");
  return !(e && atoi(e) == 128);
}

// Load both tile variants now (see update_kernels_init: a first launch beside the resident factor kernel must not trigger a lazy load).
int gemm_kernels_preload() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, k_gemm_nt_sub<64, 2, 4>);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncGetAttributes(&fa, k_gemm_nt_sub<128, 3, 2>);
  return (int)e;
}
