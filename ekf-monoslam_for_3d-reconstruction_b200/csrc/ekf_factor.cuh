// csrc/ekf_factor.cuh — K4b: Cholesky S = L L^T of one NB x NB innovation block by ONE CTA in shared
// memory, the inverses of its 32 x 32 diagonal blocks, and y = L^-1 nu.  Shared by the single-filter
// update (NB = 128, ekf_update.cu) and the fused batched-filter kernel (NB = 64, ekf_batch.cu).
//
// Right-looking Cholesky in steps of FOUR columns (two features): every thread reads the 4 x 4 pivot
// block (a shared-memory broadcast) and factors it redundantly — four rsqrt on the dependency chain
// per step, no data exchange — then one thread per row solves the four panel columns and all warps
// apply the rank-4 update to the trailing lower triangle with one DMMA.8x8x4 per 8 x 8 tile.  Two block
// barriers per step, NB / 4 steps.  nu rides along as row NB of the matrix: its panel entries come
// out as y = L^-1 nu (forward substitution is the same recurrence).  The previous version (one warp
// factoring 32 x 32 blocks in registers + an explicit L^-1) spent most of its time with 15 warps
// waiting on one; V = W L^-T is now a blocked triangular solve in the callers, which only needs the
// 32 x 32 diagonal-block inverses computed here.
#pragma once
#include <cuda_runtime.h>

#define FACT_THREADS 512
#define FACT_WARPS (FACT_THREADS / 32)
#ifdef FACT_DEBUG   // tools/factor_probe.cu: cycles per phase, accumulated by thread 0
__device__ long long g_fact_acc[16];
#define FACC(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); _facc[i] += _t - _fc; _fc = _t; } } while (0)
#define FACC_INIT __shared__ long long _facc[16]; if (threadIdx.x < 16) _facc[threadIdx.x] = 0; __syncthreads(); long long _fc = clock64()
#define FACC_DUMP do { if (threadIdx.x < 16) g_fact_acc[threadIdx.x] = _facc[threadIdx.x]; } while (0)
#else
#define FACC(i) do {} while (0)
#define FACC_INIT do {} while (0)
#define FACC_DUMP do {} while (0)
#endif

__device__ __forceinline__ void dmma884f(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// shared-memory doubles needed by cta_chol22<NB>
template <int NB>
__host__ __device__ constexpr int cta_chol22_smem_doubles() {
  // A[(NB + 8)][NB + 8] + pivot reciprocals[NB] + tile list (2 bytes per tile, T (T + 1) / 2 + T tiles)
  return (NB + 8) * (NB + 8) + NB + (((NB / 8) * (NB / 8 + 1) / 2 + NB / 8) * 2 + 7) / 8;
}

// fsm: workspace (see above).  Sb: NB x NB symmetric positive definite (leading dimension lds; only
// the lower triangle is read), nu: NB.  Outputs (shared or global memory):
//   Lout  NB x NB, leading dimension ldl: L on and below the diagonal, zero above
//   Dout  (NB / 32) blocks of 32 rows x ldd: inverse of the J-th 32 x 32 diagonal block of L
//   yout  NB: L^-1 nu
// Lout may alias Sb.  blockDim.x must be FACT_THREADS.  Ends with a barrier.
//
// Steps of FOUR columns (two features): the 4 x 4 pivot block is factored redundantly by every thread
// (four rsqrt on the chain), one thread per row solves the panel, and the rank-4 trailing update is one
// DMMA.8x8x4 per 8 x 8 tile of the lower triangle (tiles on a fixed 8-aligned grid; fragment rows that
// lie left of / above the trailing block are zeroed, so finished entries of L are never touched).
template <int NB>
__device__ void cta_chol22(double* fsm, const double* Sb, int lds, const double* nu, double* Lout, int ldl, double* Dout,
                           int ldd, double* yout, int* chol_fail) {
  // row stride = 8 (mod 16) doubles: the C-tile accesses of the rank-4 update are 16-byte vectors without bank
  // conflicts and the A/B fragment loads 2-way; an odd stride made every access 4-way conflicted (smem-bound)
  constexpr int NP = NB / 32, FLD = NB + 8, T = NB / 8;
  double* A = fsm;                      // [(NB + 8)][FLD]: rows 0..NB-1 S / L, row NB nu / y, rows NB+1.. padding
  double* rinvs = A + (NB + 8) * FLD;   // [NB]
  // tile list ordered by tile column J descending: the tiles of a trailing block that starts at tile
  // index I0 (J >= I0) are always a prefix of it
  unsigned char* tl = reinterpret_cast<unsigned char*>(rinvs + NB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  FACC_INIT;
  if (tid == 0) {
    int k = 0;
    for (int J = T - 1; J >= 0; --J)
      for (int I = J; I <= T; ++I) { tl[2 * k] = (unsigned char)I; tl[2 * k + 1] = (unsigned char)J; ++k; }
  }
  for (int e = tid; e < NB * NB; e += FACT_THREADS) {
    const int r = e / NB, c = e - r * NB;
    A[r * FLD + c] = (c <= r) ? Sb[(size_t)r * lds + c] : 0.0;
  }
  for (int e = tid; e < 8 * FLD; e += FACT_THREADS) A[NB * FLD + e] = (e < NB) ? nu[e] : 0.0;
  __syncthreads();
  FACC(0);
  for (int p = 0; p < NB; p += 4) {
    // 4 x 4 pivot block (shared-memory broadcast), factored by every thread
    const double* Pv = A + p * FLD + p;
    const double a00 = Pv[0], a10 = Pv[FLD], a11 = Pv[FLD + 1], a20 = Pv[2 * FLD], a21 = Pv[2 * FLD + 1], a22 = Pv[2 * FLD + 2],
                 a30 = Pv[3 * FLD], a31 = Pv[3 * FLD + 1], a32 = Pv[3 * FLD + 2], a33 = Pv[3 * FLD + 3];
    const double r0 = rsqrt(a00);
    const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0;
    const double t11 = a11 - l10 * l10;
    const double r1 = rsqrt(t11);
    const double l21 = (a21 - l20 * l10) * r1, l31 = (a31 - l30 * l10) * r1;
    const double t22 = a22 - l20 * l20 - l21 * l21;
    const double r2 = rsqrt(t22);
    const double l32 = (a32 - l30 * l20 - l31 * l21) * r2;
    const double t33 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
    const double r3 = rsqrt(t33);
    if (r3 == 123.456) *chol_fail = 2;   // keeps the chain from being sunk below the stamp in probe builds (never true)
    FACC(1);
    // panel: rows p + 4 .. NB (row NB is nu), one thread per row
    for (int r = p + 4 + tid; r <= NB; r += FACT_THREADS) {
      double* row = A + r * FLD + p;
      const double x0 = row[0] * r0;
      const double x1 = (row[1] - x0 * l10) * r1;
      const double x2 = (row[2] - x0 * l20 - x1 * l21) * r2;
      const double x3 = (row[3] - x0 * l30 - x1 * l31 - x2 * l32) * r3;
      row[0] = x0; row[1] = x1; row[2] = x2; row[3] = x3;
    }
    FACC(2);
    __syncthreads();
    FACC(3);
    if (tid == 0) {
      if (!(a00 > 0.0) || !(t11 > 0.0) || !(t22 > 0.0) || !(t33 > 0.0)) *chol_fail = 1;
      double* Pw = A + p * FLD + p;
      Pw[0] = a00 * r0;
      Pw[FLD] = l10; Pw[FLD + 1] = t11 * r1;
      Pw[2 * FLD] = l20; Pw[2 * FLD + 1] = l21; Pw[2 * FLD + 2] = t22 * r2;
      Pw[3 * FLD] = l30; Pw[3 * FLD + 1] = l31; Pw[3 * FLD + 2] = l32; Pw[3 * FLD + 3] = t33 * r3;
      rinvs[p] = r0; rinvs[p + 1] = r1; rinvs[p + 2] = r2; rinvs[p + 3] = r3;
    }
    // rank-4 trailing update on the tensor pipe: lower 8 x 8 tiles (I >= J >= I0) plus the nu tile row
    {
      const int base = p + 4, I0 = base >> 3, nt = T - I0;
      const int ntile = nt * (nt + 1) / 2 + nt;        // tiles (I, J) with J >= I0, I = J .. T (tile row T holds nu)
      // four tiles of a warp in flight at a time: the DMMA accumulate latency, not its issue rate, is the cost here
      for (int t0 = warp; t0 < ntile; t0 += 4 * FACT_WARPS) {
        double a[4], b[4], d0[4], d1[4];
        double* cp[4];
        bool wr0[4], wr1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = t0 + u * FACT_WARPS;
          const bool live = t < ntile;
          const int I = live ? tl[2 * t] : T, J = live ? tl[2 * t + 1] : 0;
          const int ra = 8 * I + g, rb = 8 * J + g;
          const bool arow = live && ra >= base && !(I == T && g > 0);   // rows above the trailing block / padding rows past nu: nothing
          a[u] = arow ? -A[ra * FLD + p + t4] : 0.0;
          b[u] = (live && rb >= base) ? A[rb * FLD + p + t4] : 0.0;
          cp[u] = A + ra * FLD + 8 * J + 2 * t4;
          const int col = 8 * J + 2 * t4;
          // finished entries (pivot rows / panel columns inside a boundary tile) are left alone: thread 0 is
          // writing the pivot block's final values concurrently
          wr0[u] = arow && col >= base;
          wr1[u] = arow && col + 1 >= base;
          const double2 cv = live ? *reinterpret_cast<const double2*>(cp[u]) : make_double2(0.0, 0.0);
          d0[u] = cv.x; d1[u] = cv.y;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) dmma884f(d0[u], d1[u], a[u], b[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (wr0[u] && wr1[u]) *reinterpret_cast<double2*>(cp[u]) = make_double2(d0[u], d1[u]);
          else if (wr0[u]) cp[u][0] = d0[u];
          else if (wr1[u]) cp[u][1] = d1[u];
        }
      }
    }
    FACC(4);
    __syncthreads();
    FACC(5);
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
  FACC(6);
  for (int e = tid; e < NB * NB; e += FACT_THREADS) {
    const int r = e / NB, c = e - r * NB;
    Lout[(size_t)r * ldl + c] = (c <= r) ? A[r * FLD + c] : 0.0;
  }
  for (int e = tid; e < NB; e += FACT_THREADS) yout[e] = A[NB * FLD + e];
  __syncthreads();
  FACC(7);
  FACC_DUMP;
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
