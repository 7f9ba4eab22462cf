// csrc/ekf_chol128.cuh — K4b for the single-filter update: Cholesky S = L L^T of one 128 x 128 innovation block by ONE CTA
// with the whole matrix resident in REGISTERS, the inverses of its four 32 x 32 diagonal blocks, and y = L^-1 nu.
//
// Why a second kernel beside cta_chol_panel (ekf_factor.cuh, still used by the batched filters at NB = 64): that one keeps S
// in shared memory and was measured to be bound by shared-memory bandwidth — every rank-4 / rank-32 update moves the C tile
// in and out (profiles/r1k_factor_step_trace.txt: ~2 300 cycles per four columns, 40 us per block, 38 % of the cfg2 step's
// critical path).  Here every 8 x 8 tile of the lower triangle lives in the DMMA accumulator layout of the warp that owns it
// for the whole factorisation (152 tiles over 16 warps: <= 10 tiles = 20 doubles per thread), so an update costs two
// fragment loads of the 8-column PANEL per operand and two DMMA.8x8x4 — no C traffic at all.
//
//   warp 0        owns the 16 diagonal tiles: after the panel of step P it updates tile (P+1, P+1) first, factors it (every
//                 lane redundantly in registers: 8 rsqrt on the chain) and inverts it, then updates the remaining diagonal tiles;
//   warps 1..15   own the 136 off-diagonal tiles (the extra tile row 16 carries nu, so y = L^-1 nu falls out of the same
//                 recurrence), round-robin in column-major order so that the shrinking trailing matrix stays balanced.
//   step P (8 columns):  [panel]  L(I,P) = C(I,P) L(P,P)^-T as C x inv(L(P,P))^T on the tensor pipe -> shared panel buffer
//                        barrier
//                        [update] C(I,J) -= L(I,P) L(J,P)^T for every owned tile with J > P; warp 0: next pivot tile
//                        barrier
// Two block barriers per 8 columns; the critical path per step is warp 0's update + 8 x 8 factorisation + inversion
// (~1 100 cycles) and one panel product.  Output format identical to cta_chol_panel<128> (L row-major, four 32 x 32
// diagonal-block inverses for the blocked triangular solve of k_blk_V, y); entries of Lout above the diagonal are not written.
#pragma once
#include <cuda_runtime.h>

#include "ekf_factor.cuh"

#define CH_THREADS 512
#define CH_RS 12                 // row stride (doubles) of an 8 x 8 tile in shared memory: fragment loads hit every bank pair twice
#define CH_TS (8 * CH_RS)
#define CH_SLOTS 10              // off-diagonal tiles per bulk warp (136 tiles over 15 warps)
#define CH_NOFF 136              // sum_{J = 0..15} (16 - J): tiles (I, J), J < I <= 16

#ifdef CH_DEBUG   // tools/chol128_probe.cu: cycles per phase (lane 0 of warp 0 / warp 1)
__device__ long long g_ch_acc[16];
#define CHT(i) do { if (lane == 0) { const long long _t = clock64(); _cacc[i] += _t - _ct; _ct = _t; } } while (0)
#define CHT_INIT long long _cacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long _ct = clock64()
#define CHT_DUMP(base) do { if (lane == 0) for (int _i = 0; _i < 8; ++_i) g_ch_acc[base + _i] = _cacc[_i]; } while (0)
#else
#define CHT(i) do {} while (0)
#define CHT_INIT do {} while (0)
#define CHT_DUMP(base) do {} while (0)
#endif

struct __align__(16) Chol128Smem {
  double pan[2][17][CH_TS];      // panel tiles L(I, P) of the current / previous step (tile row 16 = nu row)
  double ldiag[CH_TS];           // L(P, P) of the current step (lower triangle; read by the panel solve)
  double rdiag[8];               // 1 / L(j, j) of the current step
  double dg[16][CH_TS];          // the 16 diagonal tiles (warp 0 works on them in place: accumulator layout <-> factorisation)
  double ld[4][32][33];          // the four diagonal 32 x 32 blocks of L (for their inverses)
  double rinv[128];              // 1 / L(j, j)
};
static_assert(sizeof(Chol128Smem) < 100 * 1024, "Chol128Smem");

// 8 x 8 Cholesky of the tile at `t` (row stride CH_RS, lower triangle read) by ONE warp, every lane redundantly in registers
// (8 rsqrt on the chain, no data exchange).  Lane i publishes row i of L (zeros above the diagonal) into `t` and `lcopy`,
// and 1 / L(i,i) into r8a[i] and r8b[i].  Returns false on a non-positive pivot.
__device__ __forceinline__ bool warp_chol8(double* __restrict__ t, double* __restrict__ lcopy, double* __restrict__ r8a,
                                           double* __restrict__ r8b) {
  const int lane = threadIdx.x & 31;
  double a[8][8], r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) a[i][j] = t[i * CH_RS + j];   // broadcast loads
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ok = ok && (a[j][j] > 0.0);
    r[j] = rsqrt(a[j][j]);
    a[j][j] = a[j][j] * r[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) a[i][j] = a[i][j] * r[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[i][k] = __fma_rn(-a[i][j], a[k][j], a[i][k]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (lane == i) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 v = make_double2(j <= i ? a[i][j] : 0.0, j + 1 <= i ? a[i][j + 1] : 0.0);
        *reinterpret_cast<double2*>(t + i * CH_RS + j) = v;
        *reinterpret_cast<double2*>(lcopy + i * CH_RS + j) = v;
      }
      r8a[i] = r[i]; r8b[i] = r[i];
    }
  }
  __syncwarp();
  return ok;
}

// Sb: 128 x 128 SPD (leading dimension lds, even; only the lower triangle is read), nu: 128.
// Lout (ldl): L on and below the diagonal (entries above are left untouched); Dout: 4 blocks of 32 rows x ldd: inverse of the J-th
// 32 x 32 diagonal block of L; yout: L^-1 nu.  blockDim.x must be CH_THREADS.  Ends with a block barrier.
__device__ __forceinline__ void cta_chol128(void* smem_raw, const double* __restrict__ Sb, int lds, const double* __restrict__ nu,
                                            double* __restrict__ Lout, int ldl, double* __restrict__ Dout, int ldd,
                                            double* __restrict__ yout, int* chol_fail) {
  Chol128Smem& sm = *reinterpret_cast<Chol128Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  // Warp 0 and the bulk warps run SEPARATE loops (their register needs differ: the 8 x 8 factorisation keeps ~45 doubles
  // live, the bulk warps their accumulator tiles) that meet at a named barrier: bar.sync with an explicit thread count is
  // defined for arrivals from different program locations.
  auto cta_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(CH_THREADS) : "memory"); };
  // C(J, J) -= L(J, P) L(J, P)^T for four diagonal tiles at a time (independent DMMAs in flight)
  auto diag_update = [&](const double (*pn)[CH_TS], int J0, int J1) {
    for (int Jb = J0; Jb < J1; Jb += 4) {
      double2 cv[4];
      double e[4][2], f[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int J = min(Jb + u, 15);
        const double* pp = pn[J] + g * CH_RS + t4;
        cv[u] = *reinterpret_cast<const double2*>(sm.dg[J] + g * CH_RS + 2 * t4);
        e[u][0] = e[u][1] = f[u][0] = f[u][1] = 0.0;
        dmma884f(e[u][0], e[u][1], -pp[0], pp[0]);
        dmma884f(f[u][0], f[u][1], -pp[4], pp[4]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (Jb + u < J1)
          *reinterpret_cast<double2*>(sm.dg[Jb + u] + g * CH_RS + 2 * t4) = make_double2(cv[u].x + (e[u][0] + f[u][0]), cv[u].y + (e[u][1] + f[u][1]));
    }
  };
  if (warp == 0) {
    // ---- load the 16 diagonal tiles straight from global / L2 into their shared-memory home -----------------------------
    {
      double2 v[16];
#pragma unroll
      for (int J = 0; J < 16; ++J) v[J] = *reinterpret_cast<const double2*>(Sb + (size_t)(8 * J + g) * lds + 8 * J + 2 * t4);
#pragma unroll
      for (int J = 0; J < 16; ++J) *reinterpret_cast<double2*>(sm.dg[J] + g * CH_RS + 2 * t4) = v[J];
    }
    CHT_INIT;
    CHT(0);
    // Step P of warp 0: [deferred] the diagonal tiles beyond the pivot get the update of panel P-1 while the bulk warps solve
    // panel P; [barrier A] panel P is complete; the next pivot tile (P+1, P+1) is updated with panel P and factored while the
    // bulk warps run their trailing update; [barrier B].  P = -1 is the prologue (tile (0, 0)): ONE call site for the factor.
#pragma unroll 1
    for (int P = -1; P < 16; ++P) {
      if (P >= 1) diag_update(sm.pan[(P - 1) & 1], P + 1, 16);
      CHT(1);
      if (P >= 0) cta_bar();                          // A: the panel of step P is in sm.pan[P & 1]
      CHT(2);
      const int J = P + 1;
      if (J < 16) {
        if (P >= 0) {                                 // the pivot tile alone: two independent DMMAs, nothing else on the chain
          const double* pp = sm.pan[P & 1][J] + g * CH_RS + t4;
          double2* cp = reinterpret_cast<double2*>(sm.dg[J] + g * CH_RS + 2 * t4);
          const double2 cv = *cp;
          double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
          dmma884f(e0, e1, -pp[0], pp[0]);
          dmma884f(f0, f1, -pp[4], pp[4]);
          *cp = make_double2(cv.x + (e0 + f0), cv.y + (e1 + f1));
        }
        __syncwarp();
        CHT(3);
        const bool ok = warp_chol8(sm.dg[J], sm.ldiag, sm.rdiag, sm.rinv + 8 * J);
        if (lane == 0 && !ok) *chol_fail = 1;
        const double2 lv = *reinterpret_cast<const double2*>(sm.dg[J] + g * CH_RS + 2 * t4);
        *reinterpret_cast<double2*>(Lout + (size_t)(8 * J + g) * ldl + 8 * J + 2 * t4) = lv;
        double* ldp = &sm.ld[J >> 2][8 * (J & 3) + g][8 * (J & 3) + 2 * t4];
        ldp[0] = lv.x; ldp[1] = lv.y;
        CHT(4);
      }
      cta_bar();                                      // B: L(P+1, P+1) published; the bulk warps are done with panel P
      CHT(5);
    }
    CHT_DUMP(0);
  } else {
    double acc[CH_SLOTS][2];    // off-diagonal tile of slot s in the DMMA accumulator layout
    int tI[CH_SLOTS], tJ[CH_SLOTS];
#pragma unroll
    for (int s = 0; s < CH_SLOTS; ++s) {
      const int t = (warp - 1) + 15 * s;
      int J = 0, off = 0;
      while (J < 15 && off + (16 - J) <= t) { off += 16 - J; ++J; }
      const int I = J + 1 + (t - off);
      tI[s] = (t < CH_NOFF) ? I : -1; tJ[s] = (t < CH_NOFF) ? J : -1;
      double2 v = make_double2(0.0, 0.0);
      if (t < CH_NOFF) {
        if (I < 16) v = *reinterpret_cast<const double2*>(Sb + (size_t)(8 * I + g) * lds + 8 * J + 2 * t4);
        else if (g == 0) v = *reinterpret_cast<const double2*>(nu + 8 * J + 2 * t4);
      }
      acc[s][0] = v.x; acc[s][1] = v.y;
    }
    CHT_INIT;
    CHT(0);
    cta_bar();                                        // B of the prologue: L(0, 0) published
    CHT(1);
#pragma unroll 1
    for (int P = 0; P < 16; ++P) {
      double (*pn)[CH_TS] = sm.pan[P & 1];
      // ---- panel: L(I, P) = C(I, P) L(P,P)^-T by forward substitution along each row.  The four lanes that share a row
      // of the accumulator layout gather the row's 8 entries with shuffles and solve it redundantly (28 FMA + 8 MUL on
      // broadcast loads of L(P,P) and 1 / diag); each keeps its own two columns. ---------------------------------------
#pragma unroll
      for (int s = 0; s < CH_SLOTS; ++s) {
        if (tJ[s] == P) {                             // warp-uniform
          const int I = tI[s];
          double row[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            row[2 * q] = __shfl_sync(0xffffffffu, acc[s][0], (lane & ~3) + q);
            row[2 * q + 1] = __shfl_sync(0xffffffffu, acc[s][1], (lane & ~3) + q);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            double v = row[j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = __fma_rn(-row[k], sm.ldiag[j * CH_RS + k], v);
            row[j] = v * sm.rdiag[j];
          }
          double x0 = row[0], x1 = row[1];
#pragma unroll
          for (int q = 1; q < 4; ++q)
            if (t4 == q) { x0 = row[2 * q]; x1 = row[2 * q + 1]; }
          *reinterpret_cast<double2*>(pn[I] + g * CH_RS + 2 * t4) = make_double2(x0, x1);
          if (I < 16) {
            *reinterpret_cast<double2*>(Lout + (size_t)(8 * I + g) * ldl + 8 * P + 2 * t4) = make_double2(x0, x1);
            if ((I >> 2) == (P >> 2)) {
              double* ldp = &sm.ld[P >> 2][8 * (I & 3) + g][8 * (P & 3) + 2 * t4];
              ldp[0] = x0; ldp[1] = x1;
            }
          } else if (g == 0) {
            *reinterpret_cast<double2*>(yout + 8 * P + 2 * t4) = make_double2(x0, x1);
          }
        }
      }
      CHT(2);
      cta_bar();                                      // A
      CHT(3);
      // ---- trailing update with the panel of step P ------------------------------------------------------------------
#pragma unroll
      for (int s = 0; s < CH_SLOTS; ++s) {
        if (tJ[s] > P) {                              // not a finished column, not an empty slot (-1)
          const double* pa = pn[tI[s]] + g * CH_RS + t4;
          const double* pb = pn[tJ[s]] + g * CH_RS + t4;
          double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
          dmma884f(e0, e1, -pa[0], pb[0]);
          dmma884f(f0, f1, -pa[4], pb[4]);
          acc[s][0] += e0 + f0; acc[s][1] += e1 + f1;
        }
      }
      CHT(4);
      cta_bar();                                      // B
      CHT(5);
    }
    if (warp == 1) CHT_DUMP(8);
  }
  __syncthreads();
  // ---- inverses of the four diagonal 32 x 32 blocks: warp J solves X L_JJ^T = I by substitution, lane = row r of
  // X = L_JJ^-T, i.e. column r of L_JJ^-1 (same recurrence as cta_chol_panel) -----------------------------------------
  if (warp < 4) {
    const int o = 32 * warp;
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      double s0 = (c == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int dd = 0; dd < c; ++dd) {
        const double pr = x[dd] * sm.ld[warp][c][dd];
        if ((dd & 3) == 0) s0 -= pr; else if ((dd & 3) == 1) s1 -= pr; else if ((dd & 3) == 2) s2 -= pr; else s3 -= pr;
      }
      x[c] = ((s0 + s1) + (s2 + s3)) * sm.rinv[o + c];
    }
    double* X = Dout + (size_t)warp * 32 * ldd;
#pragma unroll
    for (int c = 0; c < 32; ++c) X[c * ldd + lane] = x[c];
  }
  __syncthreads();
}
