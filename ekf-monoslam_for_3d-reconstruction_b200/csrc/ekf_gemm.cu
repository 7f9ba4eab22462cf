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
#include "ekf_kernels.h"

#define GT_M 128
#define GT_N 64
#define GT_K 16
#define GT_LD 20
#define GT_STAGES 3
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
// Persistent CTAs (two per SM, 128 threads, 92 KB smem each) pull 128 x 64 tiles from an atomic
// counter: no wave quantisation, and the two co-resident CTAs drift out of phase so one tile's C
// traffic overlaps the other's DMMAs.
// lower_only: only tiles touching the lower triangle are enumerated; strictly-lower elements are
// mirrored into the upper triangle (C symmetric on input => symmetric on output).
__global__ void __launch_bounds__(GT_THREADS, 2)
k_gemm_nt_sub(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
              int M, int N, int kconst, const int* __restrict__ kdev, int lower_only, int* __restrict__ counters, int ntiles,
              int tiles_n) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ int s_tile;
  const int K = kdev ? *kdev : kconst;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;            // 2 x 2 warps, each 64 x 32
  const int g = lane >> 2, t4 = lane & 3;
  double* As = gsm;                                   // [stages][GT_M][GT_LD]
  double* Bs = gsm + GT_STAGES * GT_M * GT_LD;        // [stages][GT_N][GT_LD]
  while (K > 0) {
  if (tid == 0) s_tile = atomicAdd(&counters[0], 1);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= ntiles) break;
  int tm, tn;
  if (lower_only) {  // row tm holds tiles tn = 0 .. 2 tm + 1; rows before it hold tm^2 + tm tiles
    tm = (int)((sqrtf(4.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
    while (tm * tm + tm > tile) --tm;
    while ((tm + 1) * (tm + 1) + (tm + 1) <= tile) ++tm;
    tn = tile - (tm * tm + tm);
  } else {
    tm = tile / tiles_n; tn = tile - tm * tiles_n;
  }
  const int m0 = tm * GT_M, n0 = tn * GT_N;

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
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + wm * 64 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + wn * 32 + j * 8 + 2 * t4;
      double2 v = make_double2(0.0, 0.0);
      if (r < M && c + 1 < N) v = *reinterpret_cast<const double2*>(C + (size_t)r * ldc + c);
      else if (r < M && c < N) v.x = C[(size_t)r * ldc + c];
      acc[i][j][0] = v.x; acc[i][j][1] = v.y;
    }
  }
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<GT_STAGES - 2>();
    __syncthreads();
    const int nk = kt + GT_STAGES - 1;
    if (nk < ktiles) load_stage(nk, nk % GT_STAGES);
    cp_async_commit();
    const double* as = As + (size_t)(kt % GT_STAGES) * GT_M * GT_LD + (size_t)(wm * 64 + g) * GT_LD + t4;
    const double* bs = Bs + (size_t)(kt % GT_STAGES) * GT_N * GT_LD + (size_t)(wn * 32 + g) * GT_LD + t4;
#pragma unroll
    for (int k4 = 0; k4 < GT_K / 4; ++k4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = as[(size_t)i * 8 * GT_LD + k4 * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = -bs[(size_t)j * 8 * GT_LD + k4 * 4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  // epilogue: store.  Thread holds rows g (+8 i), column pairs 2*t4 (+8 j).
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + wm * 64 + i * 8 + g;
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
  __syncthreads();  // smem stages and s_tile are reused by the next tile
  }
  // the last CTA to leave resets the counters for the next launch
  if (tid == 0) {
    __threadfence();
    const int done = atomicAdd(&counters[1], 1);
    if (done == (int)gridDim.x - 1) { counters[0] = 0; counters[1] = 0; __threadfence(); }
  }
}

static const size_t kGemmSmem = (size_t)GT_STAGES * (GT_M + GT_N) * GT_LD * sizeof(double);

int launch_gemm_nt_sub(cudaStream_t st, double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int N,
                       int kconst, const int* kdev, int lower_only, int* counters, long long* launches) {
  static bool attr_done = false;
  static int num_sms = 0;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_nt_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_done = true;
  }
  if (M <= 0 || N <= 0) return 0;
  const int tiles_m = (M + GT_M - 1) / GT_M, tiles_n = (N + GT_N - 1) / GT_N;
  int ntiles = tiles_m * tiles_n;
  if (lower_only) {  // requires M == N: row tm has min(2 tm + 2, tiles_n) tiles
    if (M != N) return (int)cudaErrorInvalidValue;
    ntiles = 0;
    for (int tm = 0; tm < tiles_m; ++tm) ntiles += 2 * tm + 2;
  }
  const int grid = ntiles < 2 * num_sms ? ntiles : 2 * num_sms;
  k_gemm_nt_sub<<<grid, GT_THREADS, kGemmSmem, st>>>(C, ldc, A, lda, B, ldb, M, N, kconst, kdev, lower_only, counters, ntiles,
                                                    tiles_n);
  if (launches) *launches += 1;
  return 0;
}
