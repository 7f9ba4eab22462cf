// csrc/ekf_batch.cu — batch of independent filters (BASELINE config 3).
//
// B filters x N <= 32 features (n <= 206).  One filter's covariance is n x ld fp64 (~310 KB at N = 30):
// it does not fit one SM's shared memory, but one CTA per filter keeps it L2-resident for the whole
// step (148 concurrent filters x 310 KB = 46 MB of the 126 MB L2), so HBM sees each Sigma once in and
// once out.  Three launches per step:
//   k_batch_predict  (B CTAs)       motion model, block covariance propagation, h / H / gate, S blocks
//   k_match_filter_batch (N x B)    active search, shared frame (ekf_match.cu)
//   k_batch_update   (B CTAs)       1-point RANSAC, Li update, Hi rescue, Hi update, book-keeping:
//                                   W = Sigma H^T (n x k <= 64) and its Cholesky gain live in shared
//                                   memory, the rank-k downdate Sigma -= V V^T runs on the fp64 tensor
//                                   pipe (DMMA.8x8x4) straight from shared-memory V.
// plus k_batch_compact when a filter flagged features for deletion.
// Every stage evaluates the same expressions as the single-filter kernels (same device functions).
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/ekf_b200.h"
#include "ekf_cta.cuh"
#include "ekf_factor.cuh"
#include "ekf_handle.h"
#include "ekf_kernels.h"
#include "ekf_math.cuh"

#define BK 64                 // measurement rows of one stacked update (<= 32 features)
#define BNCAP 32              // feature capacity of a batched filter
#define BNMAX (EKF_CAM + 6 * BNCAP + 2)  // 208: state rows rounded up to a multiple of 8
#define BLDW (BK + 4)         // row stride of W / V / Linv in shared memory (conflict-free fragment reads)
#define BUPD_THREADS FACT_THREADS

struct BatchView {
  double* Sigma;        // [B][ncap * ld]
  double* mu;           // [B][ld]
  FeatTab ft;           // every field [B][Ncap]
  DevCtl* ctl;          // [B]
  int* n;               // [B] state dimension
  int* N;               // [B] feature count
  long long sstride;    // doubles between consecutive filters' Sigma
  int ld, Ncap, tstride;
};

// ------------------------------------------------------------------------------------------------
// predict for one filter per CTA: same arithmetic as k_predict_cov / k_predict_features / k_predict_S2
// ------------------------------------------------------------------------------------------------
#define BPRED_THREADS 128
__global__ void __launch_bounds__(BPRED_THREADS, 4) k_batch_predict(BatchView bv, FrameView fr, DevCfg cfg, double dT, double3 dv,
                                                                 double3 dw, int vcontrol) {
  __shared__ double F[169], C[169], T[169], Q[169], cam[13];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = bv.n[b], N = bv.N[b], ld = bv.ld;
  double* Sigma = bv.Sigma + (size_t)b * bv.sstride;
  double* mu = bv.mu + (size_t)b * ld;
  const FeatTab ft = feattab_slice(bv.ft, b, bv.Ncap, bv.tstride);
  DevCtl* ctl = bv.ctl + b;
  if (tid == 0) {
    const double ctrl[3] = {dw.x, dw.y, dw.z};
    double mu13[13];
    for (int i = 0; i < 13; ++i) mu13[i] = mu[i];
    d_system_jacobian(mu13, dT, ctrl, F);
    const double a3[3] = {dv.x, dv.y, dv.z};
    d_predict_state(mu13, a3, ctrl, dT);
    for (int i = 0; i < 13; ++i) cam[i] = mu13[i];
  }
  for (int e = tid; e < 169; e += BPRED_THREADS) C[e] = Sigma[(size_t)(e / 13) * ld + (e % 13)];
  __syncthreads();
  for (int e = tid; e < 169; e += BPRED_THREADS) {
    const int a = e / 13, c = e % 13;
    double s = 0;
    for (int k = 0; k < 6; ++k) {
      const double vmax = vcontrol ? cfg.Vmax[k] : cfg.Vmax[k] * 2.0;
      const double vs = (vmax / dT) / dT;
      s += (F[a * 13 + 7 + k] * vs) * F[c * 13 + 7 + k];
    }
    Q[e] = s;
    double t = 0;
    for (int k = 0; k < 13; ++k) t += F[a * 13 + k] * C[k * 13 + c];
    T[e] = t;
  }
  __syncthreads();
  for (int e = tid; e < 169; e += BPRED_THREADS) {
    const int a = e / 13, c = e % 13;
    double s = 0;
    for (int k = 0; k < 13; ++k) s += T[a * 13 + k] * F[c * 13 + k];
    Sigma[(size_t)a * ld + c] = s + Q[e];
  }
  for (int j = 13 + tid; j < n; j += BPRED_THREADS) {
    double x[13], y[13];
    for (int c = 0; c < 13; ++c) x[c] = Sigma[(size_t)c * ld + j];
    for (int a = 0; a < 13; ++a) {
      double s = 0;
      for (int c = 0; c < 13; ++c) s += F[a * 13 + c] * x[c];
      y[a] = s;
    }
    for (int a = 0; a < 13; ++a) Sigma[(size_t)a * ld + j] = y[a];
    double* row = Sigma + (size_t)j * ld;
    for (int c = 0; c < 13; ++c) x[c] = row[c];
    for (int a = 0; a < 13; ++a) {
      double s = 0;
      for (int c = 0; c < 13; ++c) s += x[c] * F[a * 13 + c];
      y[a] = s;
    }
    for (int a = 0; a < 13; ++a) row[a] = y[a];
  }
  __syncthreads();  // Sigma of this filter is final for the rest of the kernel
  // per-feature prediction (k_predict_features)
  for (int i = tid; i < N; i += BPRED_THREADS) {
    int ok = 0;
    const int pos = ft.pos[i], coding = ft.coding[i];
    const int fsz = coding ? 3 : 6;
    double fs[6];
    for (int c = 0; c < fsz; ++c) fs[c] = mu[pos + c];
    bool skip = false;
    if (!coding && fs[5] <= 0) { ft.removef[i] = 1; skip = true; }
    if (!skip) {
      double cm[13];
      for (int c = 0; c < 13; ++c) cm[c] = cam[c];
      double qc[4] = {cm[3], -cm[4], -cm[5], -cm[6]}, Rcw[9], hi[2], Hc[26], hcz;
      d_quat2rot(qc, Rcw);
      d_feature_hH(cfg.cam, fs, coding, cm, qc, Rcw, hi, Hc, &hcz);
      const int half = cfg.window / 2;
      ok = (hi[0] > half && hi[1] > half && hi[0] < fr.w - half && hi[1] < fr.h - half) && (hcz >= 0);
      if (ok) {
        ft.h[2 * i] = hi[0]; ft.h[2 * i + 1] = hi[1];
        for (int c = 0; c < 26; ++c) ft.Hc[26 * i + c] = Hc[c];
      }
    }
    ft.innov[i] = ok; ft.li[i] = 0; ft.hi[i] = 0;
    if (ok) {
      const int chunks = cfg.tstride >> 4;
      const uint4* src = reinterpret_cast<const uint4*>(ft.patch + (size_t)i * cfg.tstride);
      uint4* dst = reinterpret_cast<uint4*>(ft.mpatch + (size_t)i * cfg.tstride);
      for (int c = 0; c < chunks; ++c) dst[c] = src[c];
    }
  }
  __syncthreads();
  if (tid < 13) mu[tid] = cam[tid];
  const int m = block_compact(ft.innov, N, ft.sel, ft.pos_in_z);
  if (tid == 0) ctl->m_innov = m;
  __syncthreads();
  // 2x2 innovation blocks (k_predict_S2): one warp per feature
  const int lane = tid & 31;
  for (int f = tid >> 5; f < N; f += BPRED_THREADS / 32) {
    if (!ft.innov[f]) continue;
    const int pos = ft.pos[f], nd = 7 + (ft.coding[f] ? 3 : 6);
    const double* hc = ft.Hc + 26 * f;
    double t0 = 0, t1 = 0, h0 = 0, h1 = 0;
    if (lane < nd) {
      const int jb = ekf_idx13(lane, pos);
      double sg[13];
#pragma unroll
      for (int c = 0; c < 13; ++c) sg[c] = (c < nd) ? Sigma[(size_t)ekf_idx13(c, pos) * ld + jb] : 0.0;  // independent loads in flight
#pragma unroll
      for (int c = 0; c < 13; ++c)
        if (c < nd) { t0 += hc[c] * sg[c]; t1 += hc[13 + c] * sg[c]; }
      h0 = hc[lane]; h1 = hc[13 + lane];
    }
    double s00 = t0 * h0, s01 = t0 * h1, s10 = t1 * h0, s11 = t1 * h1;
    for (int o = 8; o > 0; o >>= 1) {
      s00 += __shfl_down_sync(0xffffffffu, s00, o, 16); s01 += __shfl_down_sync(0xffffffffu, s01, o, 16);
      s10 += __shfl_down_sync(0xffffffffu, s10, o, 16); s11 += __shfl_down_sync(0xffffffffu, s11, o, 16);
    }
    if (lane == 0) {
      ft.S2[4 * f + 0] = s00 + cfg.sigma_pixel_2; ft.S2[4 * f + 1] = s01;
      ft.S2[4 * f + 2] = s10; ft.S2[4 * f + 3] = s11 + cfg.sigma_pixel_2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// update for one filter per CTA
// ------------------------------------------------------------------------------------------------
struct UpdSmem {
  double* W;      // [BNMAX][BLDW]   W = Sigma H^T, then V = W L^-T
  double* fact;   // Cholesky workspace / row staging of the W gather (kFactDoubles)
  double* Dv;     // [BK / 32][32][BLDD] inverses of the diagonal blocks of L
  double* Sb;     // [BK][BLDW]      S, then L
  double* nu;     // [BK]
  double* y;      // [BK]
  double* mu_i;   // [BNMAX]
  double* Hs;     // [BNCAP][27]
  double* qj;     // [48] quaternion-normalisation scratch: J[16], C[16], T[16]
  int* ibuf;      // cand[BNCAP], fids[BNCAP], poss[BNCAP], nds[BNCAP]
};
#define BLDD 36
static constexpr size_t kFactDoubles = 2 * 16 * BNMAX > cta_chol_panel_smem_doubles<BK>() ? 2 * 16 * BNMAX : cta_chol_panel_smem_doubles<BK>();
static constexpr size_t kUpdSmemDoubles =
    (size_t)BNMAX * BLDW + kFactDoubles + (BK / 32) * 32 * BLDD + (size_t)BK * BLDW + BK + BK + BNMAX + BNCAP * 27 + 48 + (4 * BNCAP) / 2;
static constexpr size_t kUpdSmemBytes = kUpdSmemDoubles * sizeof(double);

// One stacked update over the `cnt` features listed in ft.sel (V:1036-1064 / V:1245-1284).
__device__ void cta_stacked_update(const UpdSmem& sm, double* __restrict__ Sigma, int ld, int n, double* __restrict__ mu,
                                   const FeatTab& ft, int cnt, const DevCfg& cfg, DevCtl* ctl) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const int k = 2 * cnt;
  int* fids = sm.ibuf + BNCAP; int* poss = fids + BNCAP; int* nds = poss + BNCAP;
  for (int a = tid; a < BNCAP; a += BUPD_THREADS) {
    if (a < cnt) {
      const int f = ft.sel[a];
      fids[a] = f; poss[a] = ft.pos[f]; nds[a] = 7 + (ft.coding[f] ? 3 : 6);
    } else { fids[a] = -1; poss[a] = 0; nds[a] = 0; }
  }
  __syncthreads();
  for (int e = tid; e < BNCAP * 26; e += BUPD_THREADS) {
    const int a = e / 26, c = e % 26;
    sm.Hs[a * 27 + c] = (a < cnt) ? ft.Hc[26 * fids[a] + c] : 0.0;
  }
  if (tid < BK) {
    double v = 0.0;
    if (tid < k) { const int f = fids[tid >> 1]; v = ft.z[2 * f + (tid & 1)] - ft.h[2 * f + (tid & 1)]; }
    sm.nu[tid] = v;
  }
  __syncthreads();
  // W = Sigma H^T (n x k), zero-padded to BNMAX x BK.  Sigma rows are staged through shared memory
  // (the Cholesky workspace is idle here) in chunks of 16 rows with cp.async, double buffered, so the
  // 13-column gathers hit shared memory instead of 13 dependent L2 round trips; thread = (row, feature).
  {
    constexpr int CH = 16;
    double* stage = sm.fact;                      // 2 x CH x ld doubles (ld <= 208)
    const int nchunk = (n + CH - 1) / CH;
    const int vec_per_row = ld >> 1;              // 16-byte pieces per row (ld is a multiple of 8)
    auto issue = [&](int ck) {
      const int r0 = ck * CH;
      double* dst = stage + (size_t)(ck & 1) * CH * ld;
      for (int e = tid; e < CH * vec_per_row; e += BUPD_THREADS) {
        const int r = e / vec_per_row, v = e - r * vec_per_row;
        if (r0 + r < n) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * ld + 2 * v);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(Sigma + (size_t)(r0 + r) * ld + 2 * v));
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    issue(0);
    const int rl = tid >> 5, a = tid & 31;       // 16 rows x 32 features per chunk
    const int pos = poss[a], nd = nds[a];
    const double* hs = sm.Hs + a * 27;
    for (int ck = 0; ck < nchunk; ++ck) {
      if (ck + 1 < nchunk) { issue(ck + 1); asm volatile("cp.async.wait_group 1;\n" ::); }
      else asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();
      const int i = ck * CH + rl;
      double w0 = 0, w1 = 0;
      if (i < n && a < cnt) {
        const double* row = stage + (size_t)(ck & 1) * CH * ld + (size_t)rl * ld;
        for (int c = 0; c < nd; ++c) {
          const double s = row[ekf_idx13(c, pos)];
          w0 += s * hs[c]; w1 += s * hs[13 + c];
        }
      }
      if (i < BNMAX) *reinterpret_cast<double2*>(sm.W + (size_t)i * BLDW + 2 * a) = make_double2(w0, w1);
      __syncthreads();
    }
    for (int e = tid; e < (BNMAX - nchunk * CH) * (BK / 2); e += BUPD_THREADS) {   // zero rows past the last chunk
      const int i = nchunk * CH + e / (BK / 2), aa = e % (BK / 2);
      *reinterpret_cast<double2*>(sm.W + (size_t)i * BLDW + 2 * aa) = make_double2(0.0, 0.0);
    }
  }
  __syncthreads();
  // S = H W + sigma_px^2 I; rows / columns past k are identity
  for (int e = tid; e < BK * BK; e += BUPD_THREADS) {
    const int r = e / BK, s = e % BK;
    double v = (r == s) ? 1.0 : 0.0;
    if (r < k && s < k) {
      const int a = r >> 1;
      const int pos = poss[a], nd = nds[a];
      const double* hc = sm.Hs + a * 27 + 13 * (r & 1);
      double acc = 0;
      for (int c = 0; c < nd; ++c) acc += hc[c] * sm.W[(size_t)ekf_idx13(c, pos) * BLDW + s];
      v = acc + ((r == s) ? cfg.sigma_pixel_2 : 0.0);
    }
    sm.Sb[r * BLDW + s] = v;
  }
  __syncthreads();
  cta_chol_panel<BK>(sm.fact, sm.Sb, BLDW, sm.nu, sm.Sb, BLDW, sm.Dv, BLDD, sm.y, &ctl->chol_fail);
  // V = W L^-T in place: every warp owns 8-row tiles for the whole blocked triangular solve
  for (int rt = warp; rt < BNMAX / 8; rt += BUPD_THREADS / 32) {
    double part = warp_trsm_tile<BK>(sm.W + (size_t)rt * 8 * BLDW, BLDW, sm.Sb, BLDW, sm.Dv, BLDD, sm.y);
    // mu += V y: the four lanes of a row hold its partial sums
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    const int i = rt * 8 + g;
    if (t4 == 0 && i < n) mu[i] += part;
  }
  __syncthreads();
  // Sigma -= V V^T on the fp64 tensor pipe: 16 x 16 warp tiles (2 x 2 DMMA tiles), K = ceil(k / 4) * 4.
  // The C tile of a warp's next tile is loaded before the DMMAs of the current one (L2 latency hidden).
  {
    const int nt = (n + 15) >> 4, ksteps = (k + 3) >> 2, ntile = nt * nt;
    auto load_c = [&](int t, double (&c)[2][2][2]) {
      const int ti = t / nt, tj = t - ti * nt;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int r = ti * 16 + a * 8 + g, col = tj * 16 + cc * 8 + 2 * t4;
          double2 v = make_double2(0.0, 0.0);
          if (t < ntile) {
            if (r < n && col + 1 < n) v = *reinterpret_cast<const double2*>(Sigma + (size_t)r * ld + col);
            else if (r < n && col < n) v.x = Sigma[(size_t)r * ld + col];
          }
          c[a][cc][0] = v.x; c[a][cc][1] = v.y;
        }
    };
    double nxt[2][2][2];
    load_c(warp, nxt);
    for (int t = warp; t < ntile; t += BUPD_THREADS / 32) {
      const int ti = t / nt, tj = t - ti * nt;
      double acc[2][2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) { acc[a][cc][0] = nxt[a][cc][0]; acc[a][cc][1] = nxt[a][cc][1]; }
      load_c(t + BUPD_THREADS / 32, nxt);
      const double* va = sm.W + (size_t)(ti * 16 + g) * BLDW + t4;
      const double* vb = sm.W + (size_t)(tj * 16 + g) * BLDW + t4;
      for (int q = 0; q < ksteps; ++q) {
        const double a0 = va[4 * q], a1 = va[8 * BLDW + 4 * q];
        const double b0 = -vb[4 * q], b1 = -vb[8 * BLDW + 4 * q];
        dmma884f(acc[0][0][0], acc[0][0][1], a0, b0);
        dmma884f(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884f(acc[1][0][0], acc[1][0][1], a1, b0);
        dmma884f(acc[1][1][0], acc[1][1][1], a1, b1);
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int r = ti * 16 + a * 8 + g, col = tj * 16 + cc * 8 + 2 * t4;
          if (r < n && col + 1 < n) *reinterpret_cast<double2*>(Sigma + (size_t)r * ld + col) = make_double2(acc[a][cc][0], acc[a][cc][1]);
          else if (r < n && col < n) Sigma[(size_t)r * ld + col] = acc[a][cc][0];
        }
    }
  }
  __syncthreads();
  // normalizeQuaternion (V:1625-1642): 4 rows + 4 columns
  double* J = sm.qj; double* Cq = J + 16; double* Tq = Cq + 16;
  if (tid == 0) {
    double q[4];
    for (int i = 0; i < 4; ++i) q[i] = mu[3 + i];
    const double norma = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double sc = 1 / (norma * norma * norma);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) J[i * 4 + j] = ((norma * norma) * (i == j ? 1.0 : 0.0) - q[i] * q[j]) * sc;
    for (int i = 0; i < 4; ++i) mu[3 + i] = q[i] / norma;
  }
  if (tid >= 32 && tid < 48) Cq[tid - 32] = Sigma[(size_t)(3 + (tid - 32) / 4) * ld + 3 + ((tid - 32) % 4)];
  __syncthreads();
  if (tid < 16) {
    const int a = tid / 4, c = tid % 4;
    double t = 0;
    for (int kk = 0; kk < 4; ++kk) t += J[a * 4 + kk] * Cq[kk * 4 + c];
    Tq[tid] = t;
  }
  __syncthreads();
  if (tid < 16) {
    const int a = tid / 4, c = tid % 4;
    double s = 0;
    for (int kk = 0; kk < 4; ++kk) s += Tq[a * 4 + kk] * J[c * 4 + kk];
    Sigma[(size_t)(3 + a) * ld + 3 + c] = s;
  }
  for (int jj = tid; jj < n - 4; jj += BUPD_THREADS) {
    const int j = jj >= 3 ? jj + 4 : jj;
    double x[4], yv[4];
    for (int c = 0; c < 4; ++c) x[c] = Sigma[(size_t)(3 + c) * ld + j];
    for (int a = 0; a < 4; ++a) {
      double s = 0;
      for (int c = 0; c < 4; ++c) s += J[a * 4 + c] * x[c];
      yv[a] = s;
    }
    for (int a = 0; a < 4; ++a) Sigma[(size_t)(3 + a) * ld + j] = yv[a];
    double* row = Sigma + (size_t)j * ld + 3;
    for (int c = 0; c < 4; ++c) x[c] = row[c];
    for (int a = 0; a < 4; ++a) {
      double s = 0;
      for (int c = 0; c < 4; ++c) s += x[c] * J[a * 4 + c];
      yv[a] = s;
    }
    for (int a = 0; a < 4; ++a) row[a] = yv[a];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(BUPD_THREADS, 1) k_batch_update(BatchView bv, DevCfg cfg, const uint32_t* __restrict__ picks,
                                                                  int n_picks, double* __restrict__ out_mu14,
                                                                  double* __restrict__ out_S14, int* __restrict__ out_stats,
                                                                  int min_features, int max_features) {
  extern __shared__ __align__(16) double usm[];
  UpdSmem sm;
  sm.W = usm;
  sm.fact = sm.W + (size_t)BNMAX * BLDW;
  sm.Dv = sm.fact + kFactDoubles;
  sm.Sb = sm.Dv + (BK / 32) * 32 * BLDD;
  sm.nu = sm.Sb + (size_t)BK * BLDW;
  sm.y = sm.nu + BK;
  sm.mu_i = sm.y + BK;
  sm.Hs = sm.mu_i + BNMAX;
  sm.qj = sm.Hs + BNCAP * 27;
  sm.ibuf = reinterpret_cast<int*>(sm.qj + 48);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = bv.n[b], N = bv.N[b], ld = bv.ld;
  double* Sigma = bv.Sigma + (size_t)b * bv.sstride;
  double* mu = bv.mu + (size_t)b * ld;
  const FeatTab ft = feattab_slice(bv.ft, b, bv.Ncap, bv.tstride);
  DevCtl* ctl = bv.ctl + b;
  if (tid == 0) ctl->chol_fail = 0;
  // 1-point RANSAC, low-innovation update
  cta_ransac(Sigma, ld, n, mu, ft, N, ctl, cfg, picks, n_picks, sm.mu_i, sm.ibuf);
  __syncthreads();
  const int n_li = ctl->n_li;
  if (n_li > 0) cta_stacked_update(sm, Sigma, ld, n, mu, ft, n_li, cfg, ctl);
  // high-innovation rescue (k_hi_rescue): pose from the old mu, feature parameters from mu_tmp
  for (int i = tid; i < N; i += BUPD_THREADS) {
    int flag = 0;
    if (!ft.li[i] && ft.innov[i]) {
      const int pos = ft.pos[i], coding = ft.coding[i];
      const int fsz = coding ? 3 : 6, nd = 7 + fsz;
      double fs[6], r[3], qc[4], Rcw[9], hi[2], Hc[26], hcz;
      for (int c = 0; c < fsz; ++c) fs[c] = mu[pos + c];
      for (int c = 0; c < 3; ++c) r[c] = ctl->cam_old[c];
      qc[0] = ctl->cam_old[3]; qc[1] = -ctl->cam_old[4]; qc[2] = -ctl->cam_old[5]; qc[3] = -ctl->cam_old[6];
      d_quat2rot(qc, Rcw);
      d_feature_hH(cfg.cam, fs, coding, r, qc, Rcw, hi, Hc, &hcz);
      ft.h[2 * i] = hi[0]; ft.h[2 * i + 1] = hi[1];
      for (int c = 0; c < 26; ++c) ft.Hc[26 * i + c] = Hc[c];
      double Tm[26];
      for (int bb = 0; bb < nd; ++bb) {
        const int jb = ekf_idx13(bb, pos);
        double sg[13];
#pragma unroll
        for (int c = 0; c < 13; ++c) sg[c] = (c < nd) ? Sigma[(size_t)ekf_idx13(c, pos) * ld + jb] : 0.0;
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int c = 0; c < 13; ++c)
          if (c < nd) { t0 += Hc[c] * sg[c]; t1 += Hc[13 + c] * sg[c]; }
        Tm[bb] = t0; Tm[13 + bb] = t1;
      }
      double S[4] = {0, 0, 0, 0};
      for (int bb = 0; bb < nd; ++bb) {
        S[0] += Tm[bb] * Hc[bb]; S[1] += Tm[bb] * Hc[13 + bb];
        S[2] += Tm[13 + bb] * Hc[bb]; S[3] += Tm[13 + bb] * Hc[13 + bb];
      }
      const double det = S[0] * S[3] - S[2] * S[1];
      const double invdet = 1.0 / det;
      const double i00 = S[3] * invdet, i10 = -S[2] * invdet, i01 = -S[1] * invdet, i11 = S[0] * invdet;
      const double e0 = hi[0] - ft.z[2 * i], e1 = hi[1] - ft.z[2 * i + 1];
      const double t0 = e0 * i00 + e1 * i10;
      const double t1 = e0 * i01 + e1 * i11;
      const double chi = t0 * e0 + t1 * e1;
      flag = (chi <= cfg.th_hi) ? 1 : 0;
    }
    ft.hi[i] = flag;
  }
  __syncthreads();
  const int n_hi = block_compact(ft.hi, N, ft.sel, ft.pos_in_z);
  if (tid == 0) { ctl->n_hi = n_hi; ctl->k_rows = 2 * n_hi; }
  __syncthreads();
  if (n_hi > 0) cta_stacked_update(sm, Sigma, ld, n, mu, ft, n_hi, cfg, ctl);
  // book-keeping (k_bookkeeping) + visibility count (V:1296-1315)
  int my_remove = 0, my_vis = 0;
  for (int i = tid; i < N; i += BUPD_THREADS) {
    int nfind = ft.n_find[i];
    const int ntot = ft.n_tot[i];
    if (ft.hi[i] || ft.li[i]) nfind++;
    ft.n_find[i] = nfind;
    const float qi = (float)(ntot - nfind) / ((float)nfind);
    ft.quality[i] = qi;
    if (qi > cfg.quality_ratio) ft.removef[i] = 1;
    if (ft.removef[i]) my_remove++;
    else if (ft.innov[i]) my_vis++;
  }
  const int n_rem = __syncthreads_count(my_remove);   // N <= 32 <= blockDim: one feature per thread
  const int n_vis = __syncthreads_count(my_vis);
  for (int e = tid; e < 14; e += BUPD_THREADS) out_mu14[(size_t)b * 14 + e] = mu[e];
  if (out_S14)
    for (int e = tid; e < 196; e += BUPD_THREADS) out_S14[(size_t)b * 196 + e] = Sigma[(size_t)(e / 14) * ld + (e % 14)];
  if (tid == 0) {
    int* st = out_stats + (size_t)b * EKF_BATCH_STAT_FIELDS;
    int removed = n_rem, topup = 0;
    if (n_vis < min_features) {
      if (N - n_rem > max_features) removed += 1;   // removeFeature(0) (V:1313), applied by k_batch_compact
      topup = min_features - n_vis;
    }
    ctl->n_remove = removed;
    ctl->n_visible = n_vis;
    st[EKF_BSTAT_INNOV] = ctl->m_innov; st[EKF_BSTAT_MATCHED] = ctl->n_matched; st[EKF_BSTAT_LI] = n_li;
    st[EKF_BSTAT_HI] = n_hi; st[EKF_BSTAT_HYPS] = ctl->ransac_hyps; st[EKF_BSTAT_CHOL_FAIL] = ctl->chol_fail;
    st[EKF_BSTAT_REMOVED] = removed; st[EKF_BSTAT_TOPUP] = topup;
  }
}

// ------------------------------------------------------------------------------------------------
// removeFeature (V:373-421) for every flagged feature of every filter that has any: Sigma / mu are
// compacted through the filter's slice of the scratch buffer, the feature table through registers.
// ------------------------------------------------------------------------------------------------
#define BCMP_THREADS 256
__global__ void __launch_bounds__(BCMP_THREADS) k_batch_compact(BatchView bv, double* __restrict__ scratch, int min_features,
                                                                int max_features) {
  __shared__ int keep[BNCAP], newpos[BNCAP], map[BNMAX];
  __shared__ int s_N2, s_n2;
  const int b = blockIdx.x, tid = threadIdx.x;
  DevCtl* ctl = bv.ctl + b;
  if (ctl->n_remove <= 0) return;
  const int n = bv.n[b], N = bv.N[b], ld = bv.ld;
  double* Sigma = bv.Sigma + (size_t)b * bv.sstride;
  double* tmp = scratch + (size_t)b * bv.sstride;
  double* mu = bv.mu + (size_t)b * ld;
  const FeatTab ft = feattab_slice(bv.ft, b, bv.Ncap, bv.tstride);
  if (tid == 0) {
    int flagged = 0;
    for (int f = 0; f < N; ++f) flagged += ft.removef[f] ? 1 : 0;
    // removeFeature(0) after the flagged ones are gone (V:1312-1313): first survivor
    const bool drop_first = (ctl->n_visible < min_features) && (N - flagged > max_features);
    int N2 = 0, n2 = EKF_CAM;
    bool dropped = false;
    for (int i = 0; i < EKF_CAM; ++i) map[i] = i;
    for (int f = 0; f < N; ++f) {
      if (ft.removef[f]) continue;
      if (drop_first && !dropped) { dropped = true; continue; }
      const int fs = ft.coding[f] ? 3 : 6;
      keep[N2] = f; newpos[N2] = n2; ++N2;
      for (int c = 0; c < fs; ++c) map[n2++] = ft.pos[f] + c;
    }
    s_N2 = N2; s_n2 = n2;
  }
  __syncthreads();
  const int N2 = s_N2, n2 = s_n2;
  for (int e = tid; e < n2 * n2; e += BCMP_THREADS) {
    const int i = e / n2, j = e - i * n2;
    tmp[(size_t)i * ld + j] = Sigma[(size_t)map[i] * ld + map[j]];
  }
  double mv = 0.0;
  if (tid < n2) mv = mu[map[tid]];   // n2 <= 206 <= BCMP_THREADS
  __syncthreads();
  for (int e = tid; e < n2 * n2; e += BCMP_THREADS) {
    const int i = e / n2, j = e - i * n2;
    Sigma[(size_t)i * ld + j] = tmp[(size_t)i * ld + j];
  }
  if (tid < n2) mu[tid] = mv;
  // feature table: thread f < N2 moves record keep[f] -> f through registers
  const int w16 = bv.tstride >> 4;
  int o = -1;
  int iv[10]; float fv[4]; double dv[8 + 26];
  if (tid < N2) {
    o = keep[tid];
    iv[0] = ft.coding[o]; iv[1] = ft.innov[o]; iv[2] = ft.li[o]; iv[3] = ft.hi[o]; iv[4] = ft.removef[o];
    iv[5] = ft.n_tot[o]; iv[6] = ft.n_find[o]; iv[7] = ft.real_index[o]; iv[8] = ft.pos_in_z[o]; iv[9] = 0;
    fv[0] = ft.center[2 * o]; fv[1] = ft.center[2 * o + 1]; fv[2] = ft.quality[o]; fv[3] = ft.last_ncc[o];
    for (int c = 0; c < 2; ++c) { dv[c] = ft.z[2 * o + c]; dv[2 + c] = ft.h[2 * o + c]; }
    for (int c = 0; c < 4; ++c) dv[4 + c] = ft.S2[4 * o + c];
    for (int c = 0; c < 26; ++c) dv[8 + c] = ft.Hc[26 * o + c];
  }
  __syncthreads();
  if (tid < N2) {
    ft.pos[tid] = newpos[tid]; ft.coding[tid] = iv[0]; ft.innov[tid] = iv[1]; ft.li[tid] = iv[2]; ft.hi[tid] = iv[3];
    ft.removef[tid] = iv[4]; ft.n_tot[tid] = iv[5]; ft.n_find[tid] = iv[6]; ft.real_index[tid] = iv[7]; ft.pos_in_z[tid] = iv[8];
    ft.center[2 * tid] = fv[0]; ft.center[2 * tid + 1] = fv[1]; ft.quality[tid] = fv[2]; ft.last_ncc[tid] = fv[3];
    for (int c = 0; c < 2; ++c) { ft.z[2 * tid + c] = dv[c]; ft.h[2 * tid + c] = dv[2 + c]; }
    for (int c = 0; c < 4; ++c) ft.S2[4 * tid + c] = dv[4 + c];
    for (int c = 0; c < 26; ++c) ft.Hc[26 * tid + c] = dv[8 + c];
  }
  // templates: records only move towards lower indices, in ascending order
  for (int f = 0; f < N2; ++f) {
    const int src = keep[f];
    if (src == f) continue;
    uint4 a = make_uint4(0, 0, 0, 0), c = make_uint4(0, 0, 0, 0);
    if (tid < w16) {
      a = reinterpret_cast<const uint4*>(ft.patch + (size_t)src * bv.tstride)[tid];
      c = reinterpret_cast<const uint4*>(ft.mpatch + (size_t)src * bv.tstride)[tid];
    }
    __syncthreads();
    if (tid < w16) {
      reinterpret_cast<uint4*>(ft.patch + (size_t)f * bv.tstride)[tid] = a;
      reinterpret_cast<uint4*>(ft.mpatch + (size_t)f * bv.tstride)[tid] = c;
    }
    __syncthreads();
  }
  if (tid == 0) { bv.n[b] = n2; bv.N[b] = N2; ctl->n_remove = 0; }
}

// seed: every filter <- the single filter `src`
__global__ void __launch_bounds__(256) k_batch_seed(BatchView bv, const double* __restrict__ Ssrc, int ld_src,
                                                    const double* __restrict__ musrc, FeatTab src, int n, int N) {
  const int b = blockIdx.x, tid = threadIdx.x;
  double* Sigma = bv.Sigma + (size_t)b * bv.sstride;
  double* mu = bv.mu + (size_t)b * bv.ld;
  const FeatTab ft = feattab_slice(bv.ft, b, bv.Ncap, bv.tstride);
  for (int e = tid; e < n * n; e += 256) {
    const int i = e / n, j = e - i * n;
    Sigma[(size_t)i * bv.ld + j] = Ssrc[(size_t)i * ld_src + j];
  }
  for (int e = tid; e < n; e += 256) mu[e] = musrc[e];
  for (int f = tid; f < N; f += 256) {
    ft.pos[f] = src.pos[f]; ft.coding[f] = src.coding[f]; ft.innov[f] = src.innov[f]; ft.li[f] = src.li[f]; ft.hi[f] = src.hi[f];
    ft.removef[f] = src.removef[f]; ft.n_tot[f] = src.n_tot[f]; ft.n_find[f] = src.n_find[f]; ft.real_index[f] = src.real_index[f];
    ft.pos_in_z[f] = src.pos_in_z[f]; ft.quality[f] = src.quality[f]; ft.last_ncc[f] = src.last_ncc[f];
    for (int c = 0; c < 2; ++c) { ft.center[2 * f + c] = src.center[2 * f + c]; ft.z[2 * f + c] = src.z[2 * f + c]; ft.h[2 * f + c] = src.h[2 * f + c]; }
    for (int c = 0; c < 4; ++c) ft.S2[4 * f + c] = src.S2[4 * f + c];
    for (int c = 0; c < 26; ++c) ft.Hc[26 * f + c] = src.Hc[26 * f + c];
  }
  for (int e = tid; e < N * bv.tstride; e += 256) { ft.patch[e] = src.patch[e]; ft.mpatch[e] = src.mpatch[e]; }
  if (tid == 0) { bv.n[b] = n; bv.N[b] = N; bv.ctl[b] = DevCtl{}; }
}
__global__ void k_batch_set_cam(BatchView bv, const double* __restrict__ mu14, int B) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < B * 14) bv.mu[(size_t)(e / 14) * bv.ld + (e % 14)] = mu14[e];
}

// ================================================================================================
// host side
// ================================================================================================
struct ekf_batch {
  ekf_config cfg;
  DevCfg dcfg;
  int device = 0, B = 0, Ncap = 0, ncap = 0, ld = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  BatchView bv{};
  double* scratch = nullptr;
  uint8_t* frame = nullptr;
  size_t frame_cap = 0;
  FrameView fv{nullptr, 0, 0, 0};
  EkfTensorMap frame_map{};
  int* match_defer = nullptr;   // [B * Ncap] (filter, feature) pairs the warp matcher left to the CTA matcher, then the count
  uint32_t* picks_dev = nullptr;
  int picks_cap = 0;
  double *out_mu14 = nullptr, *out_S14 = nullptr;
  int* out_stats = nullptr;
  double *h_mu14 = nullptr, *h_S14 = nullptr;   // pinned
  int* h_stats = nullptr;                       // pinned
  int *h_n = nullptr, *h_N = nullptr;           // pinned mirrors of bv.n / bv.N
  double* stage_mu14 = nullptr;                 // device staging for set_camera_states
  double dT = 1.0, old_ts = -1.0;
  bool have_frame = false, seeded = false, results_valid = false;
  long long launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
};

static int bfail(ekf_batch* b, int code, const std::string& msg) {
  if (b) b->err = msg;
  return code;
}
#define BCHECK(expr)                                                                                    \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      char _buf[400];                                                                                   \
      snprintf(_buf, sizeof _buf, "CUDA error %d (%s) at %s:%d: %s", (int)_e, cudaGetErrorString(_e), __FILE__, __LINE__, #expr); \
      return bfail(b, EKF_ERR_CUDA, _buf);                                                              \
    }                                                                                                   \
  } while (0)

template <class T>
static cudaError_t balloc(T** p, size_t count) { return cudaMalloc((void**)p, sizeof(T) * std::max<size_t>(count, 1)); }

extern "C" {

const char* ekf_batch_last_error(const ekf_batch* b) { return b ? b->err.c_str() : "null batch"; }

int ekf_batch_destroy(ekf_batch* b) {
  if (!b) return EKF_OK;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  FeatTab& t = b->bv.ft;
  cudaFree(t.pos); cudaFree(t.coding); cudaFree(t.innov); cudaFree(t.li); cudaFree(t.hi); cudaFree(t.removef);
  cudaFree(t.n_tot); cudaFree(t.n_find); cudaFree(t.real_index); cudaFree(t.pos_in_z); cudaFree(t.sel);
  cudaFree(t.center); cudaFree(t.quality); cudaFree(t.last_ncc); cudaFree(t.z); cudaFree(t.h); cudaFree(t.Hc);
  cudaFree(t.S2); cudaFree(t.patch); cudaFree(t.mpatch);
  cudaFree(b->bv.Sigma); cudaFree(b->bv.mu); cudaFree(b->bv.ctl); cudaFree(b->bv.n); cudaFree(b->bv.N);
  cudaFree(b->match_defer);
  cudaFree(b->scratch); cudaFree(b->frame); cudaFree(b->picks_dev); cudaFree(b->out_mu14); cudaFree(b->out_S14);
  cudaFree(b->out_stats); cudaFree(b->stage_mu14);
  if (b->h_mu14) cudaFreeHost(b->h_mu14);
  if (b->h_S14) cudaFreeHost(b->h_S14);
  if (b->h_stats) cudaFreeHost(b->h_stats);
  if (b->h_n) cudaFreeHost(b->h_n);
  if (b->h_N) cudaFreeHost(b->h_N);
  for (auto& e : b->ev) if (e) cudaEventDestroy(e);
  if (b->own_stream) cudaStreamDestroy(b->own_stream);
  delete b;
  return EKF_OK;
}

int ekf_batch_create(const ekf_config* cfg, int n_filters, int feature_capacity, int device, ekf_batch** out) {
  if (!cfg || !out || n_filters < 1 || feature_capacity < 1) return EKF_ERR_ARG;
  *out = nullptr;
  if (feature_capacity > BNCAP) return EKF_ERR_CAPACITY;
  if (cfg->kernel_size < 100000 || cfg->scale != 1 || cfg->forsePlane != 0 || cfg->xyz_conversion != 0) return EKF_ERR_UNSUPPORTED;
  if (cfg->window_size < 3 || cfg->window_size > 31 || cfg->search_clamp > 20 || cfg->search_clamp < 0) return EKF_ERR_UNSUPPORTED;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return EKF_ERR_CUDA;
  ekf_batch* b = new ekf_batch;
  b->cfg = *cfg; b->device = device; b->B = n_filters; b->Ncap = feature_capacity;
  b->ncap = EKF_CAM + 6 * feature_capacity;
  b->ld = (b->ncap + 7) & ~7;
  const int w2 = (cfg->window_size * cfg->window_size + 15) & ~15;
  auto fail = [&](cudaError_t e, const char* what) {
    fprintf(stderr, "ekf_batch_create: %s: %s\n", what, cudaGetErrorString(e));
    ekf_batch_destroy(b);
    return EKF_ERR_CUDA;
  };
  cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) return fail(e, #x);
  TRY(cudaSetDevice(device))
  TRY(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking))
  b->stream = b->own_stream;
  TRY(cudaFuncSetAttribute(k_batch_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdSmemBytes))
  const size_t Bn = (size_t)n_filters, cap = Bn * feature_capacity;
  BatchView& v = b->bv;
  v.ld = b->ld; v.Ncap = feature_capacity; v.tstride = w2;
  v.sstride = (long long)b->ncap * b->ld;
  TRY(balloc(&v.Sigma, Bn * v.sstride)) TRY(balloc(&b->scratch, Bn * v.sstride)) TRY(balloc(&v.mu, Bn * b->ld))
  TRY(balloc(&v.ctl, Bn)) TRY(balloc(&v.n, Bn)) TRY(balloc(&v.N, Bn))
  TRY(balloc(&b->match_defer, cap + 1))
  FeatTab& t = v.ft;
  TRY(balloc(&t.pos, cap)) TRY(balloc(&t.coding, cap)) TRY(balloc(&t.innov, cap)) TRY(balloc(&t.li, cap)) TRY(balloc(&t.hi, cap))
  TRY(balloc(&t.removef, cap)) TRY(balloc(&t.n_tot, cap)) TRY(balloc(&t.n_find, cap)) TRY(balloc(&t.real_index, cap))
  TRY(balloc(&t.pos_in_z, cap)) TRY(balloc(&t.sel, cap)) TRY(balloc(&t.center, 2 * cap)) TRY(balloc(&t.quality, cap))
  TRY(balloc(&t.last_ncc, cap)) TRY(balloc(&t.z, 2 * cap)) TRY(balloc(&t.h, 2 * cap)) TRY(balloc(&t.Hc, 26 * cap))
  TRY(balloc(&t.S2, 4 * cap)) TRY(balloc(&t.patch, cap * w2)) TRY(balloc(&t.mpatch, cap * w2))
  TRY(balloc(&b->out_mu14, Bn * 14)) TRY(balloc(&b->out_S14, Bn * 196)) TRY(balloc(&b->out_stats, Bn * EKF_BATCH_STAT_FIELDS))
  TRY(balloc(&b->stage_mu14, Bn * 14))
  TRY(cudaMallocHost((void**)&b->h_mu14, sizeof(double) * Bn * 14)) TRY(cudaMallocHost((void**)&b->h_S14, sizeof(double) * Bn * 196))
  TRY(cudaMallocHost((void**)&b->h_stats, sizeof(int) * Bn * EKF_BATCH_STAT_FIELDS))
  TRY(cudaMallocHost((void**)&b->h_n, sizeof(int) * Bn)) TRY(cudaMallocHost((void**)&b->h_N, sizeof(int) * Bn))
  TRY(cudaMemsetAsync(v.Sigma, 0, sizeof(double) * Bn * v.sstride, b->stream))
  TRY(cudaMemsetAsync(v.mu, 0, sizeof(double) * Bn * b->ld, b->stream))
  TRY(cudaMemsetAsync(v.ctl, 0, sizeof(DevCtl) * Bn, b->stream))
  TRY(cudaMemsetAsync(v.n, 0, sizeof(int) * Bn, b->stream)) TRY(cudaMemsetAsync(v.N, 0, sizeof(int) * Bn, b->stream))
  for (auto& evn : b->ev) TRY(cudaEventCreate(&evn))
  TRY(cudaStreamSynchronize(b->stream))
#undef TRY
  for (size_t i = 0; i < Bn; ++i) { b->h_n[i] = 0; b->h_N[i] = 0; }
  DevCfg& d = b->dcfg;
  d.cam = CamParams{cfg->fx, cfg->fy, cfg->u0, cfg->v0, cfg->k1, cfg->k2, cfg->k3, cfg->p1, cfg->p2};
  d.Vmax[0] = cfg->sigma_vx * cfg->sigma_vx; d.Vmax[1] = cfg->sigma_vy * cfg->sigma_vy; d.Vmax[2] = cfg->sigma_vz * cfg->sigma_vz;
  d.Vmax[3] = cfg->sigma_wx * cfg->sigma_wx; d.Vmax[4] = cfg->sigma_wy * cfg->sigma_wy; d.Vmax[5] = cfg->sigma_wz * cfg->sigma_wz;
  d.sigma_pixel_2 = (double)(cfg->sigma_pixel * cfg->sigma_pixel);
  d.th_low = cfg->li_threshold_factor * cfg->sigma_pixel;
  d.th_hi = cfg->hi_chi2_threshold;
  d.ransac_p = cfg->ransac_p;
  d.linearity_threshold = cfg->linearity_threshold;
  d.rho_0 = cfg->rho_0; d.sigma_rho_0 = cfg->sigma_rho_0;
  d.T_camera = cfg->T_camera; d.kernel_min_size = cfg->kernel_size;
  d.ncc_threshold = (float)cfg->ncc_threshold; d.search_clamp = (float)cfg->search_clamp;
  d.sigma_size_f = (float)cfg->sigma_size; d.quality_ratio = (float)cfg->quality_ratio;
  d.window = cfg->window_size; d.sigma_pixel = cfg->sigma_pixel; d.nhyp0 = cfg->ransac_nhyp0;
  d.forsePlane = cfg->forsePlane; d.abs_int_quirk = cfg->abs_int_quirk;
  d.tstride = w2;
  *out = b;
  return EKF_OK;
}

int ekf_batch_describe(const ekf_batch* b, ekf_batch_desc* out) {
  if (!b || !out) return EKF_ERR_ARG;
  *out = ekf_batch_desc{b->B, b->Ncap, b->ncap, b->ld, b->device, {0, 0, 0}};
  return EKF_OK;
}
int ekf_batch_set_stream(ekf_batch* b, void* s) {
  if (!b) return EKF_ERR_ARG;
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  b->stream = s ? (cudaStream_t)s : b->own_stream;
  return EKF_OK;
}
int ekf_batch_sync(ekf_batch* b) {
  if (!b) return EKF_ERR_ARG;
  BCHECK(cudaSetDevice(b->device));
  BCHECK(cudaStreamSynchronize(b->stream));
  return EKF_OK;
}

int ekf_batch_seed_from(ekf_batch* b, ekf_handle* src) {
  if (!b || !src) return EKF_ERR_ARG;
  if (src->device != b->device) return bfail(b, EKF_ERR_ARG, "seed filter lives on another device");
  if (src->N > b->Ncap || src->n > b->ncap) return bfail(b, EKF_ERR_CAPACITY, "seed filter holds more features than the batch capacity");
  if (src->cfg.window_size != b->cfg.window_size) return bfail(b, EKF_ERR_ARG, "window_size differs");
  BCHECK(cudaSetDevice(b->device));
  BCHECK(cudaStreamSynchronize(src->stream));
  k_batch_seed<<<b->B, 256, 0, b->stream>>>(b->bv, src->Sigma, src->ld, src->mu, src->ft, src->n, src->N);
  b->launches += 1;
  BCHECK(cudaGetLastError());
  BCHECK(cudaStreamSynchronize(b->stream));
  for (int i = 0; i < b->B; ++i) { b->h_n[i] = src->n; b->h_N[i] = src->N; }
  b->dT = src->dT; b->old_ts = src->old_ts;
  b->seeded = true;
  b->results_valid = false;
  return EKF_OK;
}

int ekf_batch_set_camera_states(ekf_batch* b, const double* mu14) {
  if (!b || !mu14) return EKF_ERR_ARG;
  BCHECK(cudaSetDevice(b->device));
  BCHECK(cudaMemcpyAsync(b->stage_mu14, mu14, sizeof(double) * 14 * b->B, cudaMemcpyHostToDevice, b->stream));
  k_batch_set_cam<<<(b->B * 14 + 255) / 256, 256, 0, b->stream>>>(b->bv, b->stage_mu14, b->B);
  b->launches += 1;
  BCHECK(cudaGetLastError());
  BCHECK(cudaStreamSynchronize(b->stream));
  return EKF_OK;
}

static int bcapture(ekf_batch* b, const uint8_t* gray, int width, int height, int stride, double stamp, bool dev) {
  if (!b || !gray || width < 8 || height < 8 || stride < width) return EKF_ERR_ARG;
  BCHECK(cudaSetDevice(b->device));
  if (stamp >= 0) {
    if (b->old_ts > 0) b->dT = (stamp - b->old_ts);
    b->old_ts = stamp;
  }
  const int dstride = (width + 15) & ~15;
  const size_t need = (size_t)dstride * height;
  if (need > b->frame_cap) {
    BCHECK(cudaStreamSynchronize(b->stream));
    cudaFree(b->frame);
    b->frame = nullptr;
    BCHECK(cudaMalloc((void**)&b->frame, need));
    b->frame_cap = need;
  }
  BCHECK(cudaMemcpy2DAsync(b->frame, dstride, gray, stride, width, height, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           b->stream));
  if (b->fv.px != b->frame || b->fv.w != width || b->fv.h != height || b->fv.stride != dstride)
    match_make_tensor_map(&b->frame_map, b->frame, width, height, dstride, 1, b->cfg.window_size, (int)b->cfg.search_clamp);
  b->fv = FrameView{b->frame, width, height, dstride};
  b->have_frame = true;
  return EKF_OK;
}
int ekf_batch_capture_frame(ekf_batch* b, const uint8_t* g, int w, int h, int s, double stamp) { return bcapture(b, g, w, h, s, stamp, false); }
int ekf_batch_capture_frame_device(ekf_batch* b, const uint8_t* g, int w, int h, int s, double stamp) { return bcapture(b, g, w, h, s, stamp, true); }

int ekf_batch_step(ekf_batch* b, const double dv[3], const double dw[3], int vcontrol, const uint32_t* picks, int n_picks) {
  if (!b || n_picks < 0 || (n_picks > 0 && !picks)) return EKF_ERR_ARG;
  if (!b->seeded) return bfail(b, EKF_ERR_STATE, "batch step before ekf_batch_seed_from");
  if (!b->have_frame) return bfail(b, EKF_ERR_STATE, "batch step before captureNewFrame");
  BCHECK(cudaSetDevice(b->device));
  cudaStream_t st = b->stream;
  if (n_picks > b->picks_cap) {
    BCHECK(cudaStreamSynchronize(st));
    cudaFree(b->picks_dev);
    b->picks_dev = nullptr;
    BCHECK(cudaMalloc((void**)&b->picks_dev, sizeof(uint32_t) * n_picks));
    b->picks_cap = n_picks;
  }
  if (n_picks > 0) BCHECK(cudaMemcpyAsync(b->picks_dev, picks, sizeof(uint32_t) * n_picks, cudaMemcpyHostToDevice, st));
  const double z3[3] = {0, 0, 0};
  const double* a = dv ? dv : z3; const double* w = dw ? dw : z3;
  BCHECK(cudaEventRecord(b->ev[0], st));
  k_batch_predict<<<b->B, BPRED_THREADS, 0, st>>>(b->bv, b->fv, b->dcfg, b->dT, make_double3(a[0], a[1], a[2]),
                                                   make_double3(w[0], w[1], w[2]), vcontrol);
  b->launches += 1;
  BCHECK(cudaEventRecord(b->ev[1], st));
  launch_match_filter_batch(st, b->bv.ft, b->Ncap, b->bv.N, b->B, b->fv, b->dcfg, &b->frame_map, b->match_defer,
                            b->match_defer + (size_t)b->B * b->Ncap, &b->launches);
  BCHECK(cudaEventRecord(b->ev[2], st));
  k_batch_update<<<b->B, BUPD_THREADS, kUpdSmemBytes, st>>>(b->bv, b->dcfg, b->picks_dev, n_picks, b->out_mu14, b->out_S14,
                                                            b->out_stats, b->cfg.min_features, b->cfg.max_features);
  b->launches += 1;
  BCHECK(cudaEventRecord(b->ev[3], st));
  BCHECK(cudaGetLastError());
  const size_t Bn = (size_t)b->B;
  BCHECK(cudaMemcpyAsync(b->h_stats, b->out_stats, sizeof(int) * Bn * EKF_BATCH_STAT_FIELDS, cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(b->h_mu14, b->out_mu14, sizeof(double) * Bn * 14, cudaMemcpyDeviceToHost, st));
  BCHECK(cudaStreamSynchronize(st));
  b->results_valid = true;
  long long removed = 0, fails = 0;
  for (size_t i = 0; i < Bn; ++i) {
    removed += b->h_stats[i * EKF_BATCH_STAT_FIELDS + EKF_BSTAT_REMOVED];
    fails += b->h_stats[i * EKF_BATCH_STAT_FIELDS + EKF_BSTAT_CHOL_FAIL];
  }
  if (removed > 0) {
    k_batch_compact<<<b->B, BCMP_THREADS, 0, st>>>(b->bv, b->scratch, b->cfg.min_features, b->cfg.max_features);
    b->launches += 1;
    BCHECK(cudaGetLastError());
    BCHECK(cudaMemcpyAsync(b->h_n, b->bv.n, sizeof(int) * Bn, cudaMemcpyDeviceToHost, st));
    BCHECK(cudaMemcpyAsync(b->h_N, b->bv.N, sizeof(int) * Bn, cudaMemcpyDeviceToHost, st));
    BCHECK(cudaStreamSynchronize(st));
  }
  if (fails > 0) return bfail(b, EKF_ERR_STATE, "innovation covariance not positive definite in at least one filter");
  return EKF_OK;
}

int ekf_batch_get_camera_states(ekf_batch* b, double* mu14, double* sigma14, int32_t* stats) {
  if (!b) return EKF_ERR_ARG;
  if (!b->results_valid) return bfail(b, EKF_ERR_STATE, "no completed step");
  BCHECK(cudaSetDevice(b->device));
  const size_t Bn = (size_t)b->B;
  if (sigma14) {
    BCHECK(cudaMemcpyAsync(b->h_S14, b->out_S14, sizeof(double) * Bn * 196, cudaMemcpyDeviceToHost, b->stream));
    BCHECK(cudaStreamSynchronize(b->stream));
    std::copy(b->h_S14, b->h_S14 + Bn * 196, sigma14);
  }
  if (mu14) std::copy(b->h_mu14, b->h_mu14 + Bn * 14, mu14);
  if (stats) std::copy(b->h_stats, b->h_stats + Bn * EKF_BATCH_STAT_FIELDS, stats);
  return EKF_OK;
}

int ekf_batch_num_features(ekf_batch* b, int f) { return (!b || f < 0 || f >= b->B) ? EKF_ERR_ARG : b->h_N[f]; }
int ekf_batch_state_dim(ekf_batch* b, int f) { return (!b || f < 0 || f >= b->B) ? EKF_ERR_ARG : b->h_n[f]; }

int ekf_batch_get_full(ekf_batch* b, int f, double* mu, double* sigma, int ld) {
  if (!b || f < 0 || f >= b->B || !mu) return EKF_ERR_ARG;
  const int n = b->h_n[f];
  if (sigma && ld < n) return EKF_ERR_ARG;
  BCHECK(cudaSetDevice(b->device));
  BCHECK(cudaMemcpyAsync(mu, b->bv.mu + (size_t)f * b->ld, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
  if (sigma)
    BCHECK(cudaMemcpy2DAsync(sigma, sizeof(double) * ld, b->bv.Sigma + (size_t)f * b->bv.sstride, sizeof(double) * b->ld,
                             sizeof(double) * n, n, cudaMemcpyDeviceToHost, b->stream));
  BCHECK(cudaStreamSynchronize(b->stream));
  return EKF_OK;
}
int ekf_batch_set_full(ekf_batch* b, int f, const double* mu, const double* sigma, int ld) {
  if (!b || f < 0 || f >= b->B || !mu || !sigma) return EKF_ERR_ARG;
  const int n = b->h_n[f];
  if (ld < n) return EKF_ERR_ARG;
  BCHECK(cudaSetDevice(b->device));
  BCHECK(cudaMemcpyAsync(b->bv.mu + (size_t)f * b->ld, mu, sizeof(double) * n, cudaMemcpyHostToDevice, b->stream));
  BCHECK(cudaMemcpy2DAsync(b->bv.Sigma + (size_t)f * b->bv.sstride, sizeof(double) * b->ld, sigma, sizeof(double) * ld,
                           sizeof(double) * n, n, cudaMemcpyHostToDevice, b->stream));
  BCHECK(cudaStreamSynchronize(b->stream));
  return EKF_OK;
}

int ekf_batch_get_feature(ekf_batch* b, int f, int idx, ekf_feature_info* o) {
  if (!b || f < 0 || f >= b->B || !o || idx < 0 || idx >= b->h_N[f]) return EKF_ERR_ARG;
  BCHECK(cudaSetDevice(b->device));
  const FeatTab t = feattab_slice(b->bv.ft, f, b->Ncap, b->bv.tstride);
  memset(o, 0, sizeof *o);
  cudaStream_t st = b->stream;
  int iv[10];
  const int* isrc[10] = {t.pos, t.pos_in_z, t.coding, t.n_tot, t.n_find, t.real_index, t.innov, t.li, t.hi, t.removef};
  for (int c = 0; c < 10; ++c) BCHECK(cudaMemcpyAsync(&iv[c], isrc[c] + idx, sizeof(int), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(o->center, t.center + 2 * idx, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(&o->quality_index, t.quality + idx, sizeof(float), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(&o->last_ncc, t.last_ncc + idx, sizeof(float), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(o->z, t.z + 2 * idx, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(o->h, t.h + 2 * idx, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaMemcpyAsync(o->H, t.Hc + 26 * idx, 26 * sizeof(double), cudaMemcpyDeviceToHost, st));
  BCHECK(cudaStreamSynchronize(st));
  o->position_in_state = iv[0]; o->position_in_z = iv[1]; o->coding = iv[2]; o->n_tot = iv[3]; o->n_find = iv[4];
  o->real_index = iv[5]; o->is_in_innovation = iv[6]; o->is_in_li = iv[7]; o->is_in_hi = iv[8]; o->remove_flag = iv[9];
  const int pos = iv[0], fs = iv[2] ? 3 : 6;
  const double* mu = b->bv.mu + (size_t)f * b->ld;
  const double* Sg = b->bv.Sigma + (size_t)f * b->bv.sstride;
  BCHECK(cudaMemcpyAsync(o->state, mu + pos, sizeof(double) * fs, cudaMemcpyDeviceToHost, st));
  double blk[36];
  BCHECK(cudaMemcpy2DAsync(blk, sizeof(double) * fs, Sg + (size_t)pos * b->ld + pos, sizeof(double) * b->ld, sizeof(double) * fs, fs,
                           cudaMemcpyDeviceToHost, st));
  BCHECK(cudaStreamSynchronize(st));
  for (int a = 0; a < fs; ++a)
    for (int c = 0; c < fs; ++c) o->cov[a * 6 + c] = blk[a * fs + c];
  return EKF_OK;
}

int64_t ekf_batch_kernel_launches(const ekf_batch* b) { return b ? b->launches : 0; }

int ekf_batch_last_step_ms(ekf_batch* b, float out[3]) {
  if (!b || !out) return EKF_ERR_ARG;
  if (!b->results_valid) return bfail(b, EKF_ERR_STATE, "no completed step");
  for (int i = 0; i < 3; ++i)
    if (cudaEventElapsedTime(&out[i], b->ev[i], b->ev[i + 1]) != cudaSuccess) out[i] = -1.0f;
  return EKF_OK;
}

int ekf_batch_last_match_deferred(ekf_batch* b) {
  if (!b) return EKF_ERR_ARG;
  if (cudaSetDevice(b->device) != cudaSuccess) return EKF_ERR_CUDA;
  int v = 0;
  if (cudaMemcpyAsync(&v, b->match_defer + (size_t)b->B * b->Ncap, sizeof(int), cudaMemcpyDeviceToHost, b->stream) != cudaSuccess ||
      cudaStreamSynchronize(b->stream) != cudaSuccess)
    return EKF_ERR_CUDA;
  return v;
}

}  // extern "C"
