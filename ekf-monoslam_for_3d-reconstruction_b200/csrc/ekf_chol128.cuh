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
//   warp 0        is the pivot chain: it receives tile (P+1, P+1) from its owner as soon as that tile has the update of panel P
//                 and factors it (every lane redundantly in registers: 8 rsqrt on the chain) while the bulk warps update;
//   bulk warps    (the 12 warps with warp & 3 != 0) own the 136 off-diagonal tiles (the extra tile row 16 carries nu, so
//                 y = L^-1 nu falls out of the same recurrence), round-robin in column-major order so that the shrinking
//                 trailing matrix stays balanced.  Warps 4, 8 and 12 only take part in the barriers: they share warp 0's SM
//                 sub-partition, and DMMAs of another warp in the same fp64 pipe stretch every DFMA / DMMA of the pivot
//                 chain (profiles/r1k_dmma_latency.txt: 150 -> 517 cycles per dependent DMMA with four warps per sub-partition;
//                 measured here: 3 000 -> see profiles/r2_chol128_probe.txt cycles per 8 x 8 factorisation).
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

#define CH_THREADS 512           // 16 warps: warp 0 = pivot chain, warps 4, 8, 12 = panel solvers, the other 12 = tile owners
#define CH_RS 12                 // row stride (doubles) of an 8 x 8 tile in shared memory: fragment loads hit every bank pair twice
#define CH_TS (8 * CH_RS)
#define CH_SOLVERS 3
#define CH_OWNERS 12
#define CH_SLOTS 13              // tiles per owner warp (151 tiles over 12 warps)
#define CH_NTILES 151            // tiles (I, J): J = 0: I = 1..16; J >= 1: I = J..16 (tile row 16 = nu); (0, 0) belongs to warp 0

#ifdef CH_DEBUG   // tools/chol128_probe.cu: cycles per phase (lane 0 of warp 0 / warp 1)
__device__ long long g_ch_acc[16];
#define CHT(i) do { _sink += *reinterpret_cast<volatile double*>(&sm.rinv[127]); if (lane == 0) { const long long _t = clock64(); _cacc[i] += _t - _ct; _ct = _t; } } while (0)
#define CHT_INIT long long _cacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; double _sink = 0; long long _ct = clock64()
#define CHT_DUMP(base) do { if (lane == 0) { for (int _i = 0; _i < 8; ++_i) g_ch_acc[base + _i] = _cacc[_i]; if (_sink == 12345.678) g_ch_acc[15] = 1; } } while (0)
#else
#define CHT(i) do {} while (0)
#define CHT_INIT do {} while (0)
#define CHT_DUMP(base) do {} while (0)
#endif

struct __align__(16) Chol128Smem {
  double pan[17][CH_TS];         // panel tiles L(I, P) of the current step (tile row 16 = nu row)
  double ldiag[CH_TS];           // L(P, P) of the current step (lower triangle; read by the panel solve)
  double rdiag[8];               // 1 / L(j, j) of the current step
  double dg[2][CH_TS];           // the next two pivot tiles in transit: accumulator layout (owner warp) -> factorisation (warp 0)
  double ld[4][32][33];          // the four diagonal 32 x 32 blocks of L (for their inverses)
  double rinv[128];              // 1 / L(j, j)
};
static_assert(sizeof(Chol128Smem) < 100 * 1024, "Chol128Smem");

// 8 x 8 Cholesky of the tile at `t` (row stride CH_RS, lower triangle read) by ONE warp, every lane redundantly in registers
// (8 rsqrt on the chain, no data exchange).  Lane 0 stores the lower triangle of L into `lout` (entries above the diagonal
// are NOT written) and 1 / L(i,i) into r8a[i] and r8b[i] (both 16-byte aligned).  Returns false on a non-positive pivot.
__device__ __forceinline__ bool warp_chol8(const double* __restrict__ t, double* __restrict__ lout, double* __restrict__ r8a,
                                           double* __restrict__ r8b, long long* stamps = nullptr) {
  const int lane = threadIdx.x & 31;
  if (stamps) stamps[0] = clock64();
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
  if (stamps) stamps[1] = clock64() + (long long)(a[7][7] * 0.0);
  __syncwarp();
  // publish: every lane holds the same values; lane 0 stores them (ONE divergent region; row i stored by lane i cost 1 650
  // cycles in eight divergent regions, the same stores issued by all 32 lanes 570)
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; j += 2)
        *reinterpret_cast<double2*>(lout + i * CH_RS + j) = make_double2(a[i][j], j + 1 <= i ? a[i][j + 1] : 0.0);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      *reinterpret_cast<double2*>(r8a + i) = make_double2(r[i], r[i + 1]);
      *reinterpret_cast<double2*>(r8b + i) = make_double2(r[i], r[i + 1]);
    }
  }
  __syncwarp();
  if (stamps) stamps[2] = clock64();
  return ok;
}

// Sb: 128 x 128 SPD (leading dimension lds, even; only the lower triangle is read), nu: 128.
// Lout (ldl): L on and below the diagonal (entries above are left untouched); Dout: 4 blocks of 32 rows x ldd: inverse of the J-th
// 32 x 32 diagonal block of L; yout: L^-1 nu.  blockDim.x must be CH_THREADS.  Ends with a block barrier.
// (Sb and nu carry no __restrict__: k_chain_factor assembles them in global memory inside the same kernel, so their loads must
// not take the non-coherent path.)
__device__ __forceinline__ void cta_chol128(void* smem_raw, const double* Sb, int lds, const double* nu,
                                            double* __restrict__ Lout, int ldl, double* __restrict__ Dout, int ldd,
                                            double* __restrict__ yout, int* chol_fail) {
  Chol128Smem& sm = *reinterpret_cast<Chol128Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  // Three roles with SEPARATE loops (their register needs differ) that meet at named barriers; bar.sync / bar.arrive with an
  // explicit thread count are defined for arrivals from different program locations.
  //   barrier 1 (all warps):          A = "the panel of step P is solved", B = "L(P+1, P+1) is published"
  //   barrier 3 (owners + solvers):   the owners arrive once their tiles of column P are in sm.pan, the solvers wait
  auto cta_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(CH_THREADS) : "memory"); };
  if (warp == 0) {
    {   // tile (0, 0) straight from global / L2
      const double2 v = *reinterpret_cast<const double2*>(Sb + (size_t)g * lds + 2 * t4);
      *reinterpret_cast<double2*>(sm.dg[0] + g * CH_RS + 2 * t4) = v;
    }
    CHT_INIT;
    CHT(0);
    // P = -1 is the prologue (tile (0, 0)).  Step P >= 0: [A] tile (P+1, P+1) already sits in sm.dg with every update up to
    // panel P-1 (its owner handed it over one step ahead, off the chain); warp 0 applies panel P itself — two DMMAs in an
    // otherwise idle sub-partition — factors the tile and publishes it [B].
#pragma unroll 1
    for (int P = -1; P < 16; ++P) {
      if (P >= 0) cta_bar();                          // A
      CHT(1);
      const int J = P + 1;
      if (J < 16) {
        double* dt = sm.dg[J & 1];
        if (P >= 0) {
          const double* pp = sm.pan[J] + g * CH_RS + t4;
          double2* cp = reinterpret_cast<double2*>(dt + g * CH_RS + 2 * t4);
          const double2 cv = *cp;
          double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
          dmma884f(e0, e1, -pp[0], pp[0]);
          dmma884f(f0, f1, -pp[4], pp[4]);
          *cp = make_double2(cv.x + (e0 + f0), cv.y + (e1 + f1));
        }
        __syncwarp();
        CHT(2);
#ifdef CH_DEBUG
        long long st3[3];
        const bool ok = warp_chol8(dt, sm.ldiag, sm.rdiag, sm.rinv + 8 * J, st3);
        if (lane == 0) { _cacc[6] += st3[1] - st3[0]; _cacc[7] += st3[2] - st3[1]; }
#else
        const bool ok = warp_chol8(dt, sm.ldiag, sm.rdiag, sm.rinv + 8 * J);
#endif
        if (lane == 0 && !ok) *chol_fail = 1;
        CHT(3);
      }
      cta_bar();                                      // B: L(P+1, P+1) published; sm.pan free again
      CHT(4);
      if (J < 16) {                                   // off the chain: the diagonal tile goes to Lout and to the block copy
        double2 lv = *reinterpret_cast<const double2*>(sm.ldiag + g * CH_RS + 2 * t4);
        if (2 * t4 > g) lv.x = 0.0;                   // above the diagonal: stale data, not part of L
        if (2 * t4 + 1 > g) lv.y = 0.0;
        *reinterpret_cast<double2*>(Lout + (size_t)(8 * J + g) * ldl + 8 * J + 2 * t4) = lv;
        double* ldp = &sm.ld[J >> 2][8 * (J & 3) + g][8 * (J & 3) + 2 * t4];
        ldp[0] = lv.x; ldp[1] = lv.y;
      }
      CHT(5);
    }
    CHT_DUMP(0);
  } else if ((warp & 3) == 0) {
    // ---- panel solvers: one THREAD per row of the panel: X(row, :) = C(row, :) L(P,P)^-T by forward substitution with
    // L(P,P) and 1 / diag in registers (36 doubles, loaded once per step with every load in flight; with the operands read
    // from shared memory inside the recurrence every FMA waited for its own load: 2 400 cycles per step) ---------------
    // (warps 4, 8, 12: the pivot chain's own sub-partition, idle while warp 0 factors — DMMA / DFMA traffic of another warp
    // in that sub-partition stretched the 8 x 8 factorisation from 820 to 1 420 cycles)
    const int st = ((warp >> 2) - 1) * 32 + lane;      // 0 .. 95
    cta_bar();                                        // B of the prologue
#pragma unroll 1
    for (int P = 0; P < 16; ++P) {
      double Lp[8][8], rp[8];
#pragma unroll
      for (int j = 1; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < j; k += 2) {
          const double2 v = *reinterpret_cast<const double2*>(sm.ldiag + j * CH_RS + k);
          Lp[j][k] = v.x; Lp[j][k + 1] = v.y;
        }
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(sm.rdiag + j);
        rp[j] = v.x; rp[j + 1] = v.y;
      }
      asm volatile("bar.sync 3, %0;" ::"n"((CH_SOLVERS + CH_OWNERS) * 32) : "memory");   // column P of C is in sm.pan
      const int nrows = 8 * (15 - P) + 1;             // tiles P+1 .. 15, and the nu row
#pragma unroll 1
      for (int rho = st; rho < nrows; rho += 32 * CH_SOLVERS) {
        const int I = P + 1 + (rho >> 3), r = rho & 7;
        double* rowp = sm.pan[I] + r * CH_RS;
        double row[8];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          const double2 v = *reinterpret_cast<const double2*>(rowp + q);
          row[q] = v.x; row[q + 1] = v.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double v = row[j];
#pragma unroll
          for (int k = 0; k < j; ++k) v = __fma_rn(-row[k], Lp[j][k], v);
          row[j] = v * rp[j];
        }
#pragma unroll
        for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2*>(rowp + q) = make_double2(row[q], row[q + 1]);
        if (I < 16) {
          double* lo = Lout + (size_t)(8 * I + r) * ldl + 8 * P;
#pragma unroll
          for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2*>(lo + q) = make_double2(row[q], row[q + 1]);
          if ((I >> 2) == (P >> 2)) {
            double* ldp = &sm.ld[P >> 2][8 * (I & 3) + r][8 * (P & 3)];
#pragma unroll
            for (int q = 0; q < 8; ++q) ldp[q] = row[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2*>(yout + 8 * P + q) = make_double2(row[q], row[q + 1]);
        }
      }
      cta_bar();                                      // A
      cta_bar();                                      // B
    }
  } else {
    double acc[CH_SLOTS][2];    // tile of slot s in the DMMA accumulator layout
    int tIJ[CH_SLOTS];          // I | J << 8, or -1
    const int bw = (warp >> 2) * 3 + (warp & 3) - 1;   // 0 .. CH_OWNERS - 1
#pragma unroll
    for (int s = 0; s < CH_SLOTS; ++s) {
      const int t = bw + CH_OWNERS * s;
      // column-major enumeration: column 0 holds I = 1..16 (16 tiles), column J >= 1 holds I = J..16 (17 - J tiles)
      int J = 0, off = 0, cnt = 16;
      while (J < 15 && off + cnt <= t) { off += cnt; ++J; cnt = 17 - J; }
      const int I = (J == 0 ? 1 : J) + (t - off);
      tIJ[s] = (t < CH_NTILES) ? (I | (J << 8)) : -1;
      double2 v = make_double2(0.0, 0.0);
      if (t < CH_NTILES) {
        if (I < 16) v = *reinterpret_cast<const double2*>(Sb + (size_t)(8 * I + g) * lds + 8 * J + 2 * t4);
        else if (g == 0) v = *reinterpret_cast<const double2*>(nu + 8 * J + 2 * t4);
      }
      acc[s][0] = v.x; acc[s][1] = v.y;
    }
#pragma unroll
    for (int s = 0; s < CH_SLOTS; ++s)                // tile (1, 1) goes to warp 0 right away
      if (tIJ[s] == (1 | (1 << 8))) *reinterpret_cast<double2*>(sm.dg[1] + g * CH_RS + 2 * t4) = make_double2(acc[s][0], acc[s][1]);
    CHT_INIT;
    CHT(0);
    cta_bar();                                        // B of the prologue: L(0, 0) published
    CHT(1);
#pragma unroll 1
    for (int P = 0; P < 16; ++P) {
      // ---- the tiles of column P go to the solvers (accumulator layout, in place in sm.pan) -----------------------------
#pragma unroll
      for (int s = 0; s < CH_SLOTS; ++s) {
        const int I = tIJ[s] & 255;
        if ((tIJ[s] >> 8) == P && I != P && tIJ[s] >= 0)
          *reinterpret_cast<double2*>(sm.pan[I] + g * CH_RS + 2 * t4) = make_double2(acc[s][0], acc[s][1]);
      }
      __syncwarp();
      asm volatile("bar.arrive 3, %0;" ::"n"((CH_SOLVERS + CH_OWNERS) * 32) : "memory");
      CHT(2);
      cta_bar();                                      // A: the panel is solved
      CHT(3);
      // ---- trailing update with the panel of step P, in groups of four slots (operands of finished / empty slots are zeroed
      // instead of branching around them, so that the DMMAs of a group are in flight together).  Tile (P+1, P+1) is already with
      // warp 0; tile (P+2, P+2) is handed over (through sm.dg) as soon as it has this update: one step ahead, off the chain. ----
#pragma unroll
      for (int s0 = 0; s0 < CH_SLOTS; s0 += 4) {
        bool any = false;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (s0 + u < CH_SLOTS) any = any || ((tIJ[s0 + u] >> 8) > P);
        if (any) {                                    // warp-uniform; slots are in column order: the live ones are a suffix
          // The two k-steps accumulate straight into the tile (a dependent pair per slot), and nothing reads the tile before
          // the next step: the DMMAs of ALL groups stay in flight together.  (With per-group temporaries that were summed
          // inside the branch every group paid the full DMMA latency — 517 cycles with four warps per sub-partition.)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (s0 + u < CH_SLOTS) {
              const int s = s0 + u;
              const int I = tIJ[s] & 255, J = tIJ[s] >> 8;
              const bool live = J > P && !(I == P + 1 && J == P + 1);
              const double* pa = sm.pan[live ? I : 16] + g * CH_RS + t4;
              const double* pb = sm.pan[live ? J : 16] + g * CH_RS + t4;
              const double a0 = live ? -pa[0] : 0.0, a1 = live ? -pa[4] : 0.0;
              dmma884f(acc[s][0], acc[s][1], a0, pb[0]);
              dmma884f(acc[s][0], acc[s][1], a1, pb[4]);
            }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < CH_SLOTS; ++s)
        if (tIJ[s] == ((P + 2) | ((P + 2) << 8)))
          *reinterpret_cast<double2*>(sm.dg[P & 1] + g * CH_RS + 2 * t4) = make_double2(acc[s][0], acc[s][1]);
      CHT(4);
      cta_bar();                                      // B
      CHT(5);
    }
    if (bw == 0) CHT_DUMP(8);
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
