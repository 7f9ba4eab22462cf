// csrc/ekf_factor.cuh — K4b: Cholesky S = L L^T of one NB x NB innovation block by ONE CTA in shared
// memory, the inverses of its 32 x 32 diagonal blocks, and y = L^-1 nu.  Shared by the single-filter
// update (NB = 128, ekf_update.cu) and the fused batched-filter kernel (NB = 64, ekf_batch.cu).
//
// Right-looking Cholesky in steps of TWO columns (the measurement rows come in pairs, one pair per
// feature): every thread reads the 2 x 2 pivot block (a shared-memory broadcast) and factors it
// redundantly — two rsqrt on the dependency chain per step, no data exchange — then all threads scale
// the two panel columns and apply the rank-2 update to the trailing lower triangle.  Two block
// barriers per step, NB / 2 steps.  nu rides along as row NB of the matrix: its panel entries come
// out as y = L^-1 nu (forward substitution is the same recurrence).  The previous version (one warp
// factoring 32 x 32 blocks in registers + an explicit L^-1) spent most of its time with 15 warps
// waiting on one; V = W L^-T is now a blocked triangular solve in the callers, which only needs the
// 32 x 32 diagonal-block inverses computed here.
#pragma once
#include <cuda_runtime.h>

#define FACT_THREADS 512
#define FACT_WARPS (FACT_THREADS / 32)

__device__ __forceinline__ void dmma884f(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// shared-memory doubles needed by cta_chol22<NB>: A[(NB + 1)][NB + 1] + pivot reciprocals[NB]
template <int NB>
__host__ __device__ constexpr int cta_chol22_smem_doubles() { return (NB + 1) * (NB + 1) + NB; }

// fsm: workspace (see above).  Sb: NB x NB symmetric positive definite (leading dimension lds; only
// the lower triangle is read), nu: NB.  Outputs (shared or global memory):
//   Lout  NB x NB, leading dimension ldl: L on and below the diagonal, zero above
//   Dout  (NB / 32) blocks of 32 rows x ldd: inverse of the J-th 32 x 32 diagonal block of L
//   yout  NB: L^-1 nu
// Lout may alias Sb.  blockDim.x must be FACT_THREADS.  Ends with a barrier.
template <int NB>
__device__ void cta_chol22(double* fsm, const double* Sb, int lds, const double* nu, double* Lout, int ldl, double* Dout,
                           int ldd, double* yout, int* chol_fail) {
  constexpr int NP = NB / 32, FLD = NB + 1;
  double* A = fsm;                   // [(NB + 1)][FLD]
  double* rinvs = A + (NB + 1) * FLD;  // [NB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < NB * NB; e += FACT_THREADS) {
    const int r = e / NB, c = e - r * NB;
    if (c <= r) A[r * FLD + c] = Sb[(size_t)r * lds + c];
  }
  for (int e = tid; e < NB; e += FACT_THREADS) A[NB * FLD + e] = nu[e];
  __syncthreads();
  for (int p = 0; p < NB; p += 2) {
    // 2 x 2 pivot block, factored by every thread
    const double a00 = A[p * FLD + p], a10 = A[(p + 1) * FLD + p], a11 = A[(p + 1) * FLD + p + 1];
    const double r00 = rsqrt(a00);
    const double c10 = a10 * r00;
    const double t11 = a11 - c10 * c10;
    const double r11 = rsqrt(t11);
    // panel: rows p + 2 .. NB (row NB is nu)
    for (int r = p + 2 + tid; r <= NB; r += FACT_THREADS) {
      const double l0 = A[r * FLD + p] * r00;
      const double l1 = (A[r * FLD + p + 1] - l0 * c10) * r11;
      A[r * FLD + p] = l0;
      A[r * FLD + p + 1] = l1;
    }
    __syncthreads();
    if (tid == 0) {
      if (!(a00 > 0.0) || !(t11 > 0.0)) *chol_fail = 1;
      A[p * FLD + p] = a00 * r00; A[(p + 1) * FLD + p] = c10; A[(p + 1) * FLD + p + 1] = t11 * r11;
      rinvs[p] = r00; rinvs[p + 1] = r11;
    }
    // trailing update of the lower triangle (and of row NB): warp per row, lanes over columns
    for (int r = p + 2 + warp; r <= NB; r += FACT_WARPS) {
      const double l0 = A[r * FLD + p], l1 = A[r * FLD + p + 1];
      const int cmax = r < NB ? r : NB - 1;
      for (int c = p + 2 + lane; c <= cmax; c += 32) A[r * FLD + c] -= l0 * A[c * FLD + p] + l1 * A[c * FLD + p + 1];
    }
    __syncthreads();
  }
  // inverses of the diagonal blocks: warp J solves X L_JJ^T = I by substitution, lane = row r of
  // X = L_JJ^-T, i.e. column r of L_JJ^-1
  if (warp < NP) {
    const int o = 32 * warp;
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      double sx = (c == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int dd = 0; dd < c; ++dd) sx -= x[dd] * A[(o + c) * FLD + o + dd];
      x[c] = sx * rinvs[o + c];
    }
    double* X = Dout + (size_t)warp * 32 * ldd;
#pragma unroll
    for (int c = 0; c < 32; ++c) X[c * ldd + lane] = x[c];  // Dinv[c][r] = X[r][c]
  }
  for (int e = tid; e < NB * NB; e += FACT_THREADS) {
    const int r = e / NB, c = e - r * NB;
    Lout[(size_t)r * ldl + c] = (c <= r) ? A[r * FLD + c] : 0.0;
  }
  for (int e = tid; e < NB; e += FACT_THREADS) yout[e] = A[NB * FLD + e];
  __syncthreads();
}

// V = W L^-T for one 8-row tile owned by ONE warp, in place in shared memory (blocked triangular solve):
//   for J = 0 .. NB/32-1:  T_J = W_J - sum_{P<J} X_P L_JP^T ;  X_J = T_J Dinv_J^T
// Wt: the tile's rows (row stride ldw, NB columns); L: NB x NB lower (stride ldl); D: blocks of 32 x 32
// (row stride ldd inside a block, block stride 32 * ldd).  All operands in shared memory.
// Returns, per thread, the partial dot products of its row (lane / 4) with y over its columns.
template <int NB>
__device__ __forceinline__ double warp_trsm_tile(double* Wt, int ldw, const double* L, int ldl, const double* D, int ldd,
                                                 const double* y) {
  constexpr int NP = NB / 32;
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  double part = 0.0;
#pragma unroll
  for (int J = 0; J < NP; ++J) {
    double t[4][2];  // T_J: 4 column tiles of 8
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      const double2 v = *reinterpret_cast<const double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4);
      t[ct][0] = v.x; t[ct][1] = v.y;
    }
    for (int kq = 0; kq < 8 * J; ++kq) {   // K = 32 J: columns already solved
      const double a = -Wt[(size_t)g * ldw + 4 * kq + t4];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) dmma884f(t[ct][0], t[ct][1], a, L[(size_t)(32 * J + 8 * ct + g) * ldl + 4 * kq + t4]);
    }
    __syncwarp();
#pragma unroll
    for (int ct = 0; ct < 4; ++ct)
      *reinterpret_cast<double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(t[ct][0], t[ct][1]);
    __syncwarp();
    // X_J = T_J Dinv_J^T (Dinv lower triangular: column tile ct needs k < 8 ct + 8)
    double af[8];
#pragma unroll
    for (int kq = 0; kq < 8; ++kq) af[kq] = Wt[(size_t)g * ldw + 32 * J + 4 * kq + t4];
    __syncwarp();
    const double* Dj = D + (size_t)J * 32 * ldd;
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int kq = 0; kq < 2 * ct + 2; ++kq) dmma884f(d0, d1, af[kq], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + t4]);
      *reinterpret_cast<double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(d0, d1);
      part += d0 * y[32 * J + 8 * ct + 2 * t4] + d1 * y[32 * J + 8 * ct + 2 * t4 + 1];
    }
    __syncwarp();
  }
  return part;
}
