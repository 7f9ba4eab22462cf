// csrc/ekf_factor.cuh — K4b: Cholesky S = L L^T of one NB x NB innovation block by ONE CTA of 512 threads
// in shared memory, its inverse Linv = L^-1 and y = Linv nu.  Shared by the single-filter update
// (NB = 128, ekf_update.cu) and the fused batched-filter kernel (NB = 64, ekf_batch.cu).
//
// Blocked right-looking Cholesky, 32-column panels.  Per panel:
//  (a) warp 0 factors the 32x32 diagonal block entirely in registers (lane = row; pivots and column
//      entries travel by warp shuffle, so the per-column dependency chain has no block barrier; the
//      next pivot's rsqrt is started as soon as its element is final, in the shadow of the remaining
//      rank-1 updates of the current column);
//  (b) the rows below solve X L_JJ^T = A_panel by substitution, lane = row, L_JJ broadcast from
//      shared memory, reciprocals of the pivots reused;
//  (c) all warps apply the rank-32 update to the trailing block on the fp64 tensor pipe (DMMA).
// Then Linv by 32x32 blocks: diagonal blocks by substitution, off-diagonal blocks
// Linv[J][I] = -Dinv_J sum_P L[J][P] Linv[P][I] (DMMA), parked transposed in the unused upper triangle.
#pragma once
#include <cuda_runtime.h>

#define FACT_THREADS 512
#define FACT_WARPS (FACT_THREADS / 32)
#ifdef FACT_DEBUG
__device__ long long g_fact_stamp[32];
#define FSTAMP(i) do { if (threadIdx.x == 0) g_fact_stamp[i] = clock64(); } while (0)
#else
#define FSTAMP(i) do {} while (0)
#endif

__device__ __forceinline__ void dmma884f(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
// One warp: 8x8 tile D = sum_{k<32} A(r,k) B(k,n), A(r,k) = Ap[r*ar + k*ak], B(k,n) = Bp[k*bk + n*bn].
// All 16 fragments are loaded first and the 8 DMMAs run as 4 independent chains of 2 (the DMMA
// accumulate latency, not its issue rate, bounds these tiny products).  Returns the thread's two
// elements (row lane/4, columns 2*(lane%4), +1).
__device__ __forceinline__ void warp_tile_mma32(const double* Ap, int ar, int ak, const double* Bp, int bk, int bn,
                                                double& d0, double& d1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  double af[8], bf[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { af[q] = Ap[g * ar + (4 * q + t4) * ak]; bf[q] = Bp[(4 * q + t4) * bk + g * bn]; }
  double c[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) { c[q][0] = 0.0; c[q][1] = 0.0; }
#pragma unroll
  for (int q = 0; q < 8; ++q) dmma884f(c[q & 3][0], c[q & 3][1], af[q], bf[q]);
  d0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
  d1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
}


// fsm: cta_factor_smem_doubles<NB>() doubles of shared memory.  Sb: NB x NB (leading dimension lds),
// nu: NB.  Outputs Linv (leading dimension ldl, zero above the diagonal) and yout = Linv nu; Sb / nu /
// Linv / yout may live in shared or global memory (Linv may alias Sb).  blockDim.x must be FACT_THREADS.
template <int NB>
__host__ __device__ constexpr int cta_factor_smem_doubles() {
  return NB * (NB + 1) + (NB / 32) * 32 * 33 + 2 * NB + ((NB / 32) > 1 ? (NB / 32) - 1 : 1) * 32 * 33;
}
template <int NB>
__device__ void cta_factor(double* fsm, const double* Sb, int lds, const double* nu, double* Linv, int ldl, double* yout,
                           int* chol_fail) {
  constexpr int NP = NB / 32, FLD = NB + 1;
  double* A = fsm;                          // [NB][FLD]; lower: L, strictly upper: Linv^T blocks
  double* Di = A + NB * FLD;        // [4][32][33] inverses of the diagonal blocks
  double* col = Di + NP * 32 * 33;           // [2 * NB] scratch (pivot reciprocals, nu)
  double* Tb = col + 2 * NB;            // [4][32][33] block products
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  // Blocked right-looking Cholesky, 32-column panels.  Per panel:
  //  (a) warp 0 factors the 32x32 diagonal block entirely in registers (lane = row; pivots and column
  //      entries travel by warp shuffle, so the per-column dependency chain has no block barrier; the
  //      next pivot's rsqrt is started as soon as its element is final, in the shadow of the remaining
  //      rank-1 updates of the current column);
  //  (b) the rows below solve X L_JJ^T = A_panel by substitution, lane = row, L_JJ broadcast from
  //      shared memory, reciprocals of the pivots reused;
  //  (c) all threads apply the rank-32 update to the trailing block with register tiles.
  for (int e = tid; e < NB * NB; e += FACT_THREADS) A[(e / NB) * FLD + (e % NB)] = Sb[(size_t)(e / NB) * lds + (e % NB)];
  for (int e = tid; e < NB; e += FACT_THREADS) col[NB + e] = nu[e];
  double* rinvs = col;  // [NB] reciprocals of the pivots
  FSTAMP(0);
  __syncthreads();
  FSTAMP(1);
  for (int J = 0; J < NP; ++J) {
    const int o = 32 * J;
    if (ty == 0) {
      double Rr[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) Rr[c] = A[(o + tx) * FLD + o + c];
      double piv = __shfl_sync(0xffffffffu, Rr[0], 0);
      double rinv = rsqrt(piv);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (tx == 0 && !(piv > 0.0)) *chol_fail = 1;
        const double lij = (tx > j) ? Rr[j] * rinv : ((tx == j) ? piv * rinv : 0.0);
        Rr[j] = lij;
        if (tx == j) rinvs[o + j] = rinv;
        if (j + 1 < 32) {
          const double l1 = __shfl_sync(0xffffffffu, lij, j + 1);
          Rr[j + 1] -= lij * l1;
          piv = __shfl_sync(0xffffffffu, Rr[j + 1], j + 1);
          rinv = rsqrt(piv);
        }
#pragma unroll
        for (int c = j + 2; c < 32; ++c) {
          const double lc = __shfl_sync(0xffffffffu, lij, c);
          Rr[c] -= lij * lc;  // rows < c compute values that are never read
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c <= tx) A[(o + tx) * FLD + o + c] = Rr[c];
    }
    __syncthreads();
    FSTAMP(2 + 3 * J);
    const int m = NB - o - 32;  // rows below the panel
    if (m > 0) {
      if (ty < (m >> 5)) {
        const int r = o + 32 + ty * 32 + tx;
        double x[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = A[r * FLD + o + c];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          double sx = x[c];
#pragma unroll
          for (int dd = 0; dd < c; ++dd) sx -= x[dd] * A[(o + c) * FLD + o + dd];
          x[c] = sx * rinvs[o + c];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) A[r * FLD + o + c] = x[c];
      }
      __syncthreads();
      FSTAMP(3 + 3 * J);
      // trailing block on the tensor pipe: lower 8x8 tiles (ti >= tj) of A22 -= P P^T, K = 32,
      // round-robin over the 16 warps
      {
        const int base = o + 32, nt8 = m >> 3, g = tx >> 2, t4 = tx & 3;
        const int ntile = nt8 * (nt8 + 1) / 2;
        for (int t = ty; t < ntile; t += FACT_WARPS) {
          int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
          while (ti * (ti + 1) / 2 > t) --ti;
          while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
          const int tj = t - ti * (ti + 1) / 2;
          double d0, d1;
          const double* Pa = A + (size_t)(base + 8 * ti) * FLD + o;
          const double* Pb = A + (size_t)(base + 8 * tj) * FLD + o;
          warp_tile_mma32(Pa, FLD, 1, Pb, 1, FLD, d0, d1);
          double* dst = A + (size_t)(base + 8 * ti + g) * FLD + base + 8 * tj + 2 * t4;
          dst[0] -= d0;
          dst[1] -= d1;
        }
      }
      __syncthreads();
      FSTAMP(4 + 3 * J);
    }
  }
  // Inverses of the four diagonal blocks: warp J solves X L_JJ^T = I by the same substitution as
  // the panel solve (lane = row r of X = L_JJ^-T, i.e. column r of L_JJ^-1).
  if (ty < NP) {
    const int o = 32 * ty;
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      double sx = (c == tx) ? 1.0 : 0.0;
#pragma unroll
      for (int dd = 0; dd < c; ++dd) sx -= x[dd] * A[(o + c) * FLD + o + dd];
      x[c] = sx * rinvs[o + c];
    }
    double* X = Di + ty * 32 * 33;
#pragma unroll
    for (int c = 0; c < 32; ++c) X[c * 33 + tx] = x[c];  // Linv[c][r] = X[r][c]
  }
  __syncthreads();
  FSTAMP(14);
  // Off-diagonal blocks of Linv on the tensor pipe, by distance from the diagonal:
  //   Linv[J][I] = -Dinv_J sum_{P=I..J-1} L[J][P] X(P,I),  X(I,I) = Dinv_I, X(P,I) for P > I parked
  //   transposed in the unused upper triangle: A[32 I + c][32 P + r] = Linv[32 P + r][32 I + c].
  // 16 warps = the 16 8x8 tiles of a 32x32 block.
  {
    const int wti = ty >> 2, wtj = ty & 3, g = tx >> 2, t4 = tx & 3;
    for (int dist = 1; dist < NP; ++dist) {
      const int nblk = NP - dist;
      for (int b = 0; b < nblk; ++b) {
        const int I = b, J = b + dist;
        double a0 = 0.0, a1 = 0.0;
        for (int P = I; P < J; ++P) {
          double d0, d1;
          const double* Ap = A + (size_t)(J * 32 + wti * 8) * FLD + P * 32;
          if (P == I) warp_tile_mma32(Ap, FLD, 1, Di + I * 32 * 33 + wtj * 8, 33, 1, d0, d1);
          else warp_tile_mma32(Ap, FLD, 1, A + (size_t)(I * 32 + wtj * 8) * FLD + P * 32, 1, FLD, d0, d1);
          a0 += d0; a1 += d1;
        }
        double* T = Tb + b * 32 * 33;
        T[(wti * 8 + g) * 33 + wtj * 8 + 2 * t4] = a0;
        T[(wti * 8 + g) * 33 + wtj * 8 + 2 * t4 + 1] = a1;
      }
      __syncthreads();
      for (int b = 0; b < nblk; ++b) {
        const int I = b, J = b + dist;
        double d0, d1;
        warp_tile_mma32(Di + J * 32 * 33 + (wti * 8) * 33, 33, 1, Tb + b * 32 * 33 + wtj * 8, 33, 1, d0, d1);
        const int r = wti * 8 + g, c = wtj * 8 + 2 * t4;
        A[(size_t)(I * 32 + c) * FLD + J * 32 + r] = -d0;
        A[(size_t)(I * 32 + c + 1) * FLD + J * 32 + r] = -d1;
      }
      __syncthreads();
    }
  }
  FSTAMP(15);
  // write Linv (row r by warp r mod 16: coalesced stores, conflict-free transposed reads) and
  // y = Linv nu from the same values
  for (int r = ty; r < NB; r += FACT_WARPS) {
    const int Jr = r >> 5;
    double part = 0.0;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int c = tx + 32 * q;
      double v = 0.0;
      if (q < Jr) v = A[(size_t)c * FLD + r];
      else if (q == Jr) v = Di[Jr * 32 * 33 + (r & 31) * 33 + tx];  // zero above the diagonal
      Linv[(size_t)r * ldl + c] = v;
      part += v * col[NB + c];
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2);
    if (tx == 0) yout[r] = part;
  }
  FSTAMP(16);
}
