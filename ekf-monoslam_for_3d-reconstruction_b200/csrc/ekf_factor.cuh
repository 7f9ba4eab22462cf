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

// shared-memory doubles needed by cta_chol_panel<NB>
template <int NB>
__host__ __device__ constexpr int cta_chol_panel_smem_doubles() {
  // A[(NB + 8)][NB + 8] + pivot reciprocals[NB] + pivot parameters[16] + tile list (2 bytes per tile)
  return (NB + 8) * (NB + 8) + NB + 16 + (((NB / 8) * (NB / 8 + 1) / 2 + NB / 8) * 2 + 7) / 8;
}

// 4 x 4 pivot block at (p, p): factored by every lane of the calling warp (four rsqrt on the chain);
// lane 0 publishes the parameters the panel solve needs and writes the block's final values.
template <bool SKIP = false>
__device__ __forceinline__ void chol_pivot4(double* A, int FLD, int p, double* piv, double* rinvs, int* chol_fail, int lane) {
  double* Pv = A + p * FLD + p;
  if (SKIP) { if (lane < 16) { piv[lane] = 0.5; rinvs[p + (lane & 3)] = 0.5; } return; }
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
  __syncwarp();
  if (lane == 0) {
    if (!(a00 > 0.0) || !(t11 > 0.0) || !(t22 > 0.0) || !(t33 > 0.0)) *chol_fail = 1;
    Pv[0] = a00 * r0;
    Pv[FLD] = l10; Pv[FLD + 1] = t11 * r1;
    Pv[2 * FLD] = l20; Pv[2 * FLD + 1] = l21; Pv[2 * FLD + 2] = t22 * r2;
    Pv[3 * FLD] = l30; Pv[3 * FLD + 1] = l31; Pv[3 * FLD + 2] = l32; Pv[3 * FLD + 3] = t33 * r3;
    rinvs[p] = r0; rinvs[p + 1] = r1; rinvs[p + 2] = r2; rinvs[p + 3] = r3;
    piv[0] = r0; piv[1] = r1; piv[2] = r2; piv[3] = r3;
    piv[4] = l10; piv[5] = l20; piv[6] = l30; piv[7] = l21; piv[8] = l31; piv[9] = l32;
  }
}

// fsm: workspace (see above).  Sb: NB x NB symmetric positive definite (leading dimension lds; only
// the lower triangle is read), nu: NB.  Outputs (shared or global memory):
//   Lout  NB x NB, leading dimension ldl: L on and below the diagonal, zero above
//   Dout  (NB / 32) blocks of 32 rows x ldd: inverse of the J-th 32 x 32 diagonal block of L
//   yout  NB: L^-1 nu
// Lout may alias Sb.  blockDim.x must be FACT_THREADS.  Ends with a barrier.
//
// Panel-blocked right-looking Cholesky with a pivot look-ahead.  Columns are processed in panels of 32;
// inside a panel, steps of FOUR columns (two features):
//   * every thread reads the ten parameters of the step's factored 4 x 4 pivot block (broadcast) and one
//     thread per row solves the four panel columns;
//   * the rank-4 update is applied only to the remaining columns of the PANEL (at most 62 8 x 8 tiles, one
//     DMMA.8x8x4 each, one round of <= 5 tiles per warp), and while warps 1.. do that, warp 0 updates the tile
//     holding the NEXT pivot block first and factors it (four rsqrt on the chain) — the pivot chain is off
//     the other warps' critical path;
//   * after the panel's last step the columns right of the panel receive one rank-32 update (8 DMMAs per
//     tile, C tile read and written once), again with warp 0 running ahead to the next pivot.
// Tiles lie on a fixed 8-aligned grid; fragment rows left of / above the trailing block are zeroed, so
// finished entries of L are never touched.  nu rides along as row NB (tile row T): its panel entries come
// out as y = L^-1 nu.  The earlier version applied every rank-4 update to the whole trailing matrix and
// factored the pivot between two barriers: 57 us per 128 x 128 block against 40 us now (tools/factor_ko_probe.cu; the
// per-phase accounting is in DESIGN.md section 4: the updates are bound by shared-memory bandwidth).
// KO: timing knock-outs for tools/factor_probe.cu only (1 pivot chain, 2 panel solve, 4 rank-4 update, 8 rank-32 update,
// 16 diagonal-block inverses, 32 barriers); production code instantiates KO = 0.
template <int NB, int KO = 0>
__device__ void cta_chol_panel(double* fsm, const double* Sb, int lds, const double* nu, double* Lout, int ldl, double* Dout,
                               int ldd, double* yout, int* chol_fail) {
  // row stride = 8 (mod 16) doubles: the C-tile accesses of the updates are 16-byte vectors without bank
  // conflicts and the A/B fragment loads 2-way; an odd stride made every access 4-way conflicted (smem-bound)
  constexpr int NP = NB / 32, FLD = NB + 8, T = NB / 8, NTILES = T * (T + 1) / 2 + T;
  constexpr int UW = FACT_WARPS - 1;    // warps 1.. carry the updates, warp 0 the pivot look-ahead
  double* A = fsm;                      // [(NB + 8)][FLD]: rows 0..NB-1 S / L, row NB nu / y, rows NB+1.. zero padding
  double* rinvs = A + (NB + 8) * FLD;   // [NB]
  double* piv = rinvs + NB;             // [16] r0..r3, l10 l20 l30 l21 l31 l32 of the current step
  // tile list: panels ascending; inside a panel tile columns J descending, rows I = J .. T.  The tiles a
  // step needs (J >= I0 inside its panel) are a prefix of the panel's segment, the tiles right of a panel
  // are the rest of the list.
  unsigned char* tl = reinterpret_cast<unsigned char*>(piv + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  auto seg = [](int P) { const int j = 4 * P; return j * (T + 1) - j * (j - 1) / 2; };   // tiles with J < 4 P
  FACC_INIT;
  if (tid < NTILES) {
    int P = 0;
    while (P + 1 < NP && seg(P + 1) <= tid) ++P;
    int rem = tid - seg(P), J = 4 * P + 3;
    while (rem >= T - J + 1) { rem -= T - J + 1; --J; }
    tl[2 * tid] = (unsigned char)(J + rem); tl[2 * tid + 1] = (unsigned char)J;
  }
  {
    // lower triangle of S as 16-byte vectors, every load of a thread in flight before the first store (one latency,
    // not NB * NB / FACT_THREADS of them); Sb rows must be 16-byte aligned (lds even)
    constexpr int NV = NB * NB / 2 / FACT_THREADS;
    double2 v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int e = tid + q * FACT_THREADS, r = e / (NB / 2), c = (e % (NB / 2)) * 2;
      v[q] = (c <= r) ? *reinterpret_cast<const double2*>(Sb + (size_t)r * lds + c) : make_double2(0.0, 0.0);
    }
    const double nuv = (tid < NB) ? nu[tid] : 0.0;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int e = tid + q * FACT_THREADS, r = e / (NB / 2), c = (e % (NB / 2)) * 2;
      if (c + 1 > r) v[q].y = 0.0;
      *reinterpret_cast<double2*>(A + r * FLD + c) = v[q];
    }
    for (int e = tid; e < 8 * FLD; e += FACT_THREADS) A[NB * FLD + e] = (e == tid && e < NB) ? nuv : 0.0;
  }
  __syncthreads();
  if (warp == 0) chol_pivot4<(KO & 1) != 0>(A, FLD, 0, piv, rinvs, chol_fail, lane);
  __syncthreads();
  FACC(0);
  for (int p = 0; p < NB; p += 4) {
    const int P = p >> 5, pend = 32 * P + 32, base = p + 4;
    {
      const double r0 = piv[0], r1 = piv[1], r2 = piv[2], r3 = piv[3], l10 = piv[4], l20 = piv[5], l30 = piv[6], l21 = piv[7],
                   l31 = piv[8], l32 = piv[9];
      // panel: rows p + 4 .. NB (row NB is nu), one thread per row
      for (int r = base + tid; r <= NB && !(KO & 2); r += FACT_THREADS) {
        double* row = A + r * FLD + p;
        const double x0 = row[0] * r0;
        const double x1 = (row[1] - x0 * l10) * r1;
        const double x2 = (row[2] - x0 * l20 - x1 * l21) * r2;
        const double x3 = (row[3] - x0 * l30 - x1 * l31 - x2 * l32) * r3;
        row[0] = x0; row[1] = x1; row[2] = x2; row[3] = x3;
      }
    }
    FACC(1);
    if (!(KO & 32)) __syncthreads();
    FACC(2);
    if (base < pend) {
      // rank-4 update of the panel's remaining columns [base, pend): tiles (I, J), I0 <= J < 4 P + 4, I = J .. T
      const int I0 = base >> 3, nJ = 4 * P + 4 - I0;
      const int ntile = nJ * (T + 1) - (I0 + 4 * P + 3) * nJ / 2;
      const int pidx = ntile - (T - I0 + 1);          // tile (I0, I0): holds the next pivot block
      const unsigned char* tp = tl + 2 * seg(P);
      if (warp == 0) {
        const int ra = 8 * I0 + g;
        const bool arow = ra >= base;
        const double a = arow ? -A[ra * FLD + p + t4] : 0.0;
        const double b = arow ? A[ra * FLD + p + t4] : 0.0;
        double* cp = A + ra * FLD + 8 * I0 + 2 * t4;
        const int col = 8 * I0 + 2 * t4;
        double2 cv = *reinterpret_cast<const double2*>(cp);
        dmma884f(cv.x, cv.y, a, b);
        if (arow && col >= base) *reinterpret_cast<double2*>(cp) = cv;   // base is a multiple of 4: col, col + 1 on the same side
        __syncwarp();
        chol_pivot4<(KO & 1) != 0>(A, FLD, base, piv, rinvs, chol_fail, lane);
      } else if (!(KO & 4)) {
        for (int t0 = warp - 1; t0 < ntile; t0 += 5 * UW) {
          double a[5], b[5], d0[5], d1[5];
          double* cp[5];
          bool wr[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int t = t0 + u * UW;
            const bool live = t < ntile && t != pidx;
            const int I = live ? tp[2 * t] : T, J = live ? tp[2 * t + 1] : 0;
            const int ra = 8 * I + g, rb = 8 * J + g;
            const bool arow = live && ra >= base && !(I == T && g > 0);   // rows above the trailing block / padding rows past nu: nothing
            a[u] = arow ? -A[ra * FLD + p + t4] : 0.0;
            b[u] = (live && rb >= base) ? A[rb * FLD + p + t4] : 0.0;
            cp[u] = A + ra * FLD + 8 * J + 2 * t4;
            wr[u] = arow && 8 * J + 2 * t4 >= base;    // finished columns inside a boundary tile are left alone
            const double2 cv = live ? *reinterpret_cast<const double2*>(cp[u]) : make_double2(0.0, 0.0);
            d0[u] = cv.x; d1[u] = cv.y;
          }
#pragma unroll
          for (int u = 0; u < 5; ++u) dmma884f(d0[u], d1[u], a[u], b[u]);
#pragma unroll
          for (int u = 0; u < 5; ++u)
            if (wr[u]) *reinterpret_cast<double2*>(cp[u]) = make_double2(d0[u], d1[u]);
        }
      }
    } else if (pend < NB) {
      // rank-32 update of everything right of the finished panel: tiles seg(P + 1) .. NTILES, K = columns [32 P, pend)
      const int first = seg(P + 1), ntile = NTILES - first;
      const int I0 = 4 * P + 4, pidx = seg(P + 2) - first - (T - I0 + 1);   // tile (I0, I0) in the list
      const unsigned char* tp = tl + 2 * first;
      const int k0 = 32 * P;
      // Fragments come as 16-byte loads: lane (g, t4) takes columns 8 q + 2 t4 and 8 q + 2 t4 + 1 of its row, the first feeds a
      // DMMA over the even columns of the 8-chunk, the second one over the odd columns (a and b use the same column
      // assignment, so the two sums cover the chunk) — half the load instructions of 8-byte fragments and, at this row
      // stride, no bank conflicts (the updates are shared-memory-bandwidth bound); two accumulators per tile.
      if (warp == 0) {
        const int ra = 8 * I0 + g;
        double* cp = A + ra * FLD + 8 * I0 + 2 * t4;
        double2 cv = *reinterpret_cast<const double2*>(cp);
        double c1x = 0.0, c1y = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 f = *reinterpret_cast<const double2*>(A + ra * FLD + k0 + 8 * q + 2 * t4);
          dmma884f(cv.x, cv.y, -f.x, f.x);
          dmma884f(c1x, c1y, -f.y, f.y);
        }
        *reinterpret_cast<double2*>(cp) = make_double2(cv.x + c1x, cv.y + c1y);
        __syncwarp();
        chol_pivot4<(KO & 1) != 0>(A, FLD, base, piv, rinvs, chol_fail, lane);
      } else if (!(KO & 8)) {
        for (int t0 = warp - 1; t0 < ntile; t0 += 4 * UW) {
          double d0[4], d1[4], e0[4], e1[4];
          double* cp[4];
          const double *ap[4], *bp[4];
          bool live[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * UW;
            live[u] = t < ntile && t != pidx;
            const int I = live[u] ? tp[2 * t] : T, J = live[u] ? tp[2 * t + 1] : 0;
            ap[u] = A + (8 * I + g) * FLD + k0 + 2 * t4;
            bp[u] = A + (8 * J + g) * FLD + k0 + 2 * t4;
            cp[u] = A + (8 * I + g) * FLD + 8 * J + 2 * t4;
            const double2 cv = *reinterpret_cast<const double2*>(cp[u]);
            d0[u] = cv.x; d1[u] = cv.y; e0[u] = 0.0; e1[u] = 0.0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const double2 av = *reinterpret_cast<const double2*>(ap[u] + 8 * q), bv = *reinterpret_cast<const double2*>(bp[u] + 8 * q);
              dmma884f(d0[u], d1[u], -av.x, bv.x);
              dmma884f(e0[u], e1[u], -av.y, bv.y);
            }
          }
          // padding rows past nu (tile row T, g > 0) are zero in every column: their update is zero
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (live[u]) *reinterpret_cast<double2*>(cp[u]) = make_double2(d0[u] + e0[u], d1[u] + e1[u]);
        }
      }
    }
    FACC(3);
    if (!(KO & 32)) __syncthreads();
    FACC(4);
  }
  // inverses of the diagonal blocks: warp J solves X L_JJ^T = I by substitution, lane = row r of
  // X = L_JJ^-T, i.e. column r of L_JJ^-1
  if (warp < NP && !(KO & 16)) {
    const int o = 32 * warp;
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      double s0 = (c == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // four chains: the sum is latency-bound
#pragma unroll
      for (int dd = 0; dd < c; ++dd) {
        const double pr = x[dd] * A[(o + c) * FLD + o + dd];
        if ((dd & 3) == 0) s0 -= pr; else if ((dd & 3) == 1) s1 -= pr; else if ((dd & 3) == 2) s2 -= pr; else s3 -= pr;
      }
      x[c] = ((s0 + s1) + (s2 + s3)) * rinvs[o + c];
    }
    double* X = Dout + (size_t)warp * 32 * ldd;
#pragma unroll
    for (int c = 0; c < 32; ++c) X[c * ldd + lane] = x[c];  // Dinv[c][r] = X[r][c]
  }
  FACC(5);
  for (int e = tid; e < NB * NB / 2; e += FACT_THREADS) {   // diagonal tiles carry update residue above the diagonal: masked
    const int r = e / (NB / 2), c = (e % (NB / 2)) * 2;
    double2 v = *reinterpret_cast<const double2*>(A + r * FLD + c);
    if (c > r) v.x = 0.0;
    if (c + 1 > r) v.y = 0.0;
    *reinterpret_cast<double2*>(Lout + (size_t)r * ldl + c) = v;
  }
  for (int e = tid; e < NB; e += FACT_THREADS) yout[e] = A[NB * FLD + e];
  __syncthreads();
  FACC(6);
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
    // The result latency of DMMA.8x8x4 is ~150 cycles: two accumulator sets (even / odd k-steps) per column tile keep
    // eight independent chains in flight instead of four.
    double t2[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    for (int kq = 0; kq < 8 * J; kq += 2) {   // K = 32 J: columns already solved
      const double a0 = -Wt[(size_t)g * ldw + 4 * kq + t4], a1 = -Wt[(size_t)g * ldw + 4 * kq + 4 + t4];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        dmma884f(t[ct][0], t[ct][1], a0, L[(size_t)(32 * J + 8 * ct + g) * ldl + 4 * kq + t4]);
        dmma884f(t2[ct][0], t2[ct][1], a1, L[(size_t)(32 * J + 8 * ct + g) * ldl + 4 * kq + 4 + t4]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int ct = 0; ct < 4; ++ct)
      *reinterpret_cast<double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(t[ct][0] + t2[ct][0], t[ct][1] + t2[ct][1]);
    __syncwarp();
    // X_J = T_J Dinv_J^T (Dinv lower triangular: column tile ct needs k < 8 ct + 8)
    double af[8];
#pragma unroll
    for (int kq = 0; kq < 8; ++kq) af[kq] = Wt[(size_t)g * ldw + 32 * J + 4 * kq + t4];
    __syncwarp();
    const double* Dj = D + (size_t)J * 32 * ldd;
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;   // 2 ct + 2 k-steps: always even
#pragma unroll
      for (int kq = 0; kq < 2 * ct + 2; kq += 2) {
        dmma884f(d0, d1, af[kq], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + t4]);
        dmma884f(e0, e1, af[kq + 1], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + 4 + t4]);
      }
      d0 += e0; d1 += e1;
      *reinterpret_cast<double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(d0, d1);
      part += d0 * y[32 * J + 8 * ct + 2 * t4] + d1 * y[32 * J + 8 * ct + 2 * t4 + 1];
    }
    __syncwarp();
  }
  return part;
}

// The same solve with L stored PACKED: block row J (rows 32 J .. 32 J + 31, columns 0 .. 32 J - 1: all the solve reads of it) at
// Lj[J] with row stride ldj[J] — 49 KB instead of 135 KB at NB = 128.
// V = W L^-T for one 8-row tile owned by ONE warp, in place in shared memory (blocked triangular solve):
//   for J = 0 .. NB/32-1:  T_J = W_J - sum_{P<J} X_P L_JP^T ;  X_J = T_J Dinv_J^T
// Wt: the tile's rows (row stride ldw, NB columns); L: NB x NB lower (stride ldl); D: blocks of 32 x 32
// (row stride ldd inside a block, block stride 32 * ldd).  All operands in shared memory.
// Returns, per thread, the partial dot products of its row (lane / 4) with y over its columns.
template <int NB>
__device__ __forceinline__ double warp_trsm_tile_packed(double* Wt, int ldw, const double* const (&Lj)[NB / 32], const int (&ldj)[NB / 32],
                                                        const double* D, int ldd, const double* y) {
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
    // The result latency of DMMA.8x8x4 is ~150 cycles: two accumulator sets (even / odd k-steps) per column tile keep
    // eight independent chains in flight instead of four.
    double t2[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    for (int kq = 0; kq < 8 * J; kq += 2) {   // K = 32 J: columns already solved
      const double a0 = -Wt[(size_t)g * ldw + 4 * kq + t4], a1 = -Wt[(size_t)g * ldw + 4 * kq + 4 + t4];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        dmma884f(t[ct][0], t[ct][1], a0, Lj[J][(size_t)(8 * ct + g) * ldj[J] + 4 * kq + t4]);
        dmma884f(t2[ct][0], t2[ct][1], a1, Lj[J][(size_t)(8 * ct + g) * ldj[J] + 4 * kq + 4 + t4]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int ct = 0; ct < 4; ++ct)
      *reinterpret_cast<double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(t[ct][0] + t2[ct][0], t[ct][1] + t2[ct][1]);
    __syncwarp();
    // X_J = T_J Dinv_J^T (Dinv lower triangular: column tile ct needs k < 8 ct + 8)
    double af[8];
#pragma unroll
    for (int kq = 0; kq < 8; ++kq) af[kq] = Wt[(size_t)g * ldw + 32 * J + 4 * kq + t4];
    __syncwarp();
    const double* Dj = D + (size_t)J * 32 * ldd;
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;   // 2 ct + 2 k-steps: always even
#pragma unroll
      for (int kq = 0; kq < 2 * ct + 2; kq += 2) {
        dmma884f(d0, d1, af[kq], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + t4]);
        dmma884f(e0, e1, af[kq + 1], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + 4 + t4]);
      }
      d0 += e0; d1 += e1;
      *reinterpret_cast<double2*>(Wt + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(d0, d1);
      part += d0 * y[32 * J + 8 * ct + 2 * t4] + d1 * y[32 * J + 8 * ct + 2 * t4 + 1];
    }
    __syncwarp();
  }
  return part;
}

// The same solve for ONE 8-row tile by NW cooperating warps, right-looking: warp w owns the 8-column tiles ct = w, w + NW, .. of every
// 32-column block in registers (DMMA accumulator layout).  Per block J: the owners publish T_J, every warp forms its tiles of
// X_J = T_J Dinv_J^T (<= 8 k-steps in two accumulator chains), X_J goes to the output tile, and every warp subtracts X_J L_{J'J}^T from
// its own tiles of the later blocks J' (8 k-steps each, independent tiles interleaved).  The longest chain of dependent DMMAs is
// 8 per block instead of 4 J + 4 (with the whole tile in one warp: ~50 in a row, which crawl when a downdate saturates the fp64
// pipe next to them).  Xin: T on entry (row stride ldw); Xout: V on exit (may not alias Xin).  Named barrier BAR is used with
// NW * 32 threads; every one of the NW warps must call.  Returns the partial dot products of row (lane / 4) with y over the
// warp's own columns (the caller sums over the 4 lanes of a row and over the warps).
template <int NB, int NW, int BAR>
__device__ __forceinline__ double warps_trsm_tile_rl(const double* Xin, double* Xout, int ldw, const double* const (&Lj)[NB / 32],
                                                     const int (&ldj)[NB / 32], const double* D, int ldd, const double* y, int w) {
  constexpr int NP = NB / 32, NC = 4 / NW;   // column tiles per block owned by a warp
  static_assert(4 % NW == 0, "NW must divide 4");
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  auto gbar = [] { asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NW * 32) : "memory"); };
  double t[NP][NC][2];
#pragma unroll
  for (int J = 0; J < NP; ++J)
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
      const double2 v = *reinterpret_cast<const double2*>(Xin + (size_t)g * ldw + 32 * J + 8 * (w + NW * ci) + 2 * t4);
      t[J][ci][0] = v.x; t[J][ci][1] = v.y;
    }
  double part = 0.0;
  double* Ts = const_cast<double*>(Xin);   // T_J of the current block is exchanged through the input tile
#pragma unroll
  for (int J = 0; J < NP; ++J) {
    if (J > 0) {
#pragma unroll
      for (int ci = 0; ci < NC; ++ci)
        *reinterpret_cast<double2*>(Ts + (size_t)g * ldw + 32 * J + 8 * (w + NW * ci) + 2 * t4) = make_double2(t[J][ci][0], t[J][ci][1]);
    }
    gbar();
    // X_J = T_J Dinv_J^T (Dinv lower triangular: column tile ct needs k < 8 ct + 8)
    const double* Dj = D + (size_t)J * 32 * ldd;
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
      const int ct = w + NW * ci;
      double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
      for (int kq = 0; kq < 2 * ct + 2; kq += 2) {
        dmma884f(d0, d1, Ts[(size_t)g * ldw + 32 * J + 4 * kq + t4], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + t4]);
        dmma884f(e0, e1, Ts[(size_t)g * ldw + 32 * J + 4 * kq + 4 + t4], Dj[(size_t)(8 * ct + g) * ldd + 4 * kq + 4 + t4]);
      }
      d0 += e0; d1 += e1;
      *reinterpret_cast<double2*>(Xout + (size_t)g * ldw + 32 * J + 8 * ct + 2 * t4) = make_double2(d0, d1);
      part += d0 * y[32 * J + 8 * ct + 2 * t4] + d1 * y[32 * J + 8 * ct + 2 * t4 + 1];
    }
    if (J + 1 < NP) {
      gbar();   // X_J is complete in Xout
      // own tiles of the later blocks: T_J' -= X_J L_{J'J}^T, K = 32, two accumulator chains per tile
      double af[8];
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) af[kq] = -Xout[(size_t)g * ldw + 32 * J + 4 * kq + t4];
#pragma unroll
      for (int Jp = J + 1; Jp < NP; ++Jp)
#pragma unroll
        for (int ci = 0; ci < NC; ++ci) {
          const double* lb = Lj[Jp] + (size_t)(8 * (w + NW * ci) + g) * ldj[Jp] + 32 * J + t4;
          double u0 = 0.0, u1 = 0.0;
#pragma unroll
          for (int kq = 0; kq < 8; kq += 2) {
            dmma884f(t[Jp][ci][0], t[Jp][ci][1], af[kq], lb[4 * kq]);
            dmma884f(u0, u1, af[kq + 1], lb[4 * kq + 4]);
          }
          t[Jp][ci][0] += u0; t[Jp][ci][1] += u1;
        }
    }
  }
  return part;
}
