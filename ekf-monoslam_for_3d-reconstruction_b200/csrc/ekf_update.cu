// csrc/ekf_update.cu — K5 (1-point RANSAC), K4a-c (innovation covariance, Cholesky gain) and the
// high-innovation rescue / book-keeping of VSlamFilter::update (vslamRansac.cpp:964-1341).
// SURVEY.md §8(a) rows a14-a19.  V: = mono-slam/src/vslamRansac.cpp.
//
// Stacked update algebra.  The reference forms St = H Sigma H^T + R for all k selected rows, inverts
// it by LU and applies Kt = Sigma H^T St^-1, Sigma <- (I - Kt H) Sigma (V:1053-1060).  Here the k
// rows are consumed in blocks of EKF_UB rows: for each block b
//     W_b = Sigma H_b^T                       (gather GEMM: H has 13 non-zeros per row)
//     S_b = H_b W_b + sigma_px^2 I = L_b L_b^T (Cholesky, one CTA, shared memory)
//     V_b = W_b L_b^-T,  y_b = L_b^-1 (z_b - h_b - H_b delta)
//     Sigma <- Sigma - V_b V_b^T              (DMMA, ekf_gemm.cu),   delta <- delta + V_b y_b
// and mu <- mu + delta at the end.  S_b computed from the already-downdated Sigma is the Schur
// complement of the preceding blocks, so the sequence is exactly a right-looking blocked Cholesky
// of the full St whose trailing updates are carried by the covariance downdate; in exact arithmetic
// it equals the reference's one-shot update (R is diagonal, H is evaluated once at the prior mean).
#include <cstdlib>

#include "ekf_kernels.h"
#include "ekf_math.cuh"
#include "ekf_cta.cuh"
#include "ekf_factor.cuh"


// ------------------------------------------------------------------------------------------------
// K5: 1-point RANSAC (V:964-1034) — one persistent CTA runs the whole adaptive loop on the device.
// picks[] replaces rand() (V:970,989).  The Li flags left behind are those of the LAST hypothesis
// evaluated (quirk V:1022), and S_i equals the 2x2 block computed in predict (same Sigma, same H).
// ------------------------------------------------------------------------------------------------
#define RANSAC_THREADS 1024
__global__ void __launch_bounds__(RANSAC_THREADS) k_ransac(const double* __restrict__ Sigma, int ld, int n,
                                                           const double* __restrict__ mu, FeatTab ft, int N, DevCtl* ctl,
                                                           DevCfg cfg, const uint32_t* __restrict__ picks, int n_picks,
                                                           double* __restrict__ mu_i, int* __restrict__ cand) {
  cta_ransac(Sigma, ld, n, mu, ft, N, ctl, cfg, picks, n_picks, mu_i, cand);
}

// The same loop by a thread-block CLUSTER of RANSAC_CL CTAs (single filter, n >= RANSAC_CLUSTER_MIN_N): the two O(n) / O(N)
// phases of a hypothesis — mu_i = mu + (Sigma H_i^T) S_i^-1 (z_i - h_i), 13 scattered columns of Sigma for every row, and the
// re-projection of every feature — are split over the CTAs (one SM pulls ~1.25 MB of 32-byte sectors per hypothesis at
// n = 3014, which bounded the one-CTA version: 43 us); the candidate list stays with CTA 0, the pick and the inlier count
// travel through DevCtl, and cluster barriers (release / acquire at cluster scope) order the global-memory hand-overs.
// Every CTA keeps its own copy of the loop state (cnt, nhyp, best count), computed from the same inputs, so all of them
// leave the loop in the same iteration.  Arithmetic per row / per feature is that of cta_ransac: identical flags and counts.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define RANSAC_CL 8
#define RANSAC_CL_THREADS 512
#define RANSAC_CLUSTER_MIN_N 1000
__global__ void __launch_bounds__(RANSAC_CL_THREADS) k_ransac_cluster(const double* __restrict__ Sigma, int ld, int n,
                                                                      const double* __restrict__ mu, FeatTab ft, int N, DevCtl* ctl,
                                                                      DevCfg cfg, const uint32_t* __restrict__ picks, int n_picks,
                                                                      double* __restrict__ mu_i, int* __restrict__ cand) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
  const int nthr = blockDim.x, tid = threadIdx.x;
  __shared__ double Hs[26], Sinv[4], inn[2], rr[3], Rcw[9];
  __shared__ int s_pos, s_nd, s_nhyp, s_numzli, s_cnt;
  if (rank == 0) {
    const int c0 = block_compact(ft.innov, N, cand, nullptr);
    if (tid == 0) {
      ctl->n_matched = c0;
      for (int i = 0; i < 7; ++i) ctl->cam_old[i] = mu[i];
      ctl->rs_count = 0;
    }
  }
  cluster.sync();
  if (tid == 0) { s_cnt = ctl->n_matched; s_nhyp = cfg.nhyp0; s_numzli = 0; }
  __syncthreads();
  const int matched = s_cnt;
  int cnt = matched, it = 0;
  while (it < s_nhyp && cnt > 0) {
    if (rank == 0) {   // pick without replacement (V:989-991); the list is private to CTA 0
      __shared__ int s_p;
      if (tid == 0) {
        const uint32_t rv = n_picks > 0 ? picks[it % n_picks] : 0u;
        s_p = (int)(rv % (uint32_t)cnt);
        ctl->rs_sel = cand[s_p];
        ctl->rs_count = 0;
      }
      __syncthreads();
      const int p = s_p;
      // vector::erase (V:991): shift the tail down by one, in chunks of 8 * nthr entries (low to high: a chunk only
      // reads entries that later chunks write), so that any list length up to the feature capacity is handled
      for (int base = p; base < cnt - 1; base += 8 * nthr) {
        int tmp[8];
        int c = 0;
        for (int idx = base + tid; idx < cnt - 1 && c < 8; idx += nthr) tmp[c++] = cand[idx + 1];
        __syncthreads();
        c = 0;
        for (int idx = base + tid; idx < cnt - 1 && c < 8; idx += nthr) cand[idx] = tmp[c++];
        __syncthreads();
      }
    }
    cnt -= 1;
    cluster.sync();                                   // the pick is visible
    const int sel = ctl->rs_sel;
    if (tid < 26) Hs[tid] = ft.Hc[26 * sel + tid];
    if (tid == 32) {
      double S[4];
      for (int c = 0; c < 4; ++c) S[c] = ft.S2[4 * sel + c];
      double X[4];
      d_inv2_pplu(S, X);
      for (int c = 0; c < 4; ++c) Sinv[c] = X[c];
      inn[0] = ft.z[2 * sel] - ft.h[2 * sel];
      inn[1] = ft.z[2 * sel + 1] - ft.h[2 * sel + 1];
      s_pos = ft.pos[sel];
      s_nd = 7 + (ft.coding[sel] ? 3 : 6);
    }
    __syncthreads();
    {  // mu_i = mu + (Sigma H^T) S^-1 (z - h)   (V:995-996), rows interleaved over the cluster
      const int pos = s_pos, nd = s_nd;
      for (int i = rank * nthr + tid; i < n; i += ncta * nthr) {
        const double* row = Sigma + (size_t)i * ld;
        double sg[13];
#pragma unroll
        for (int c = 0; c < 13; ++c) sg[c] = (c < nd) ? row[ekf_idx13(c, pos)] : 0.0;
        double w0 = 0, w1 = 0;
#pragma unroll
        for (int c = 0; c < 13; ++c)
          if (c < nd) { w0 += sg[c] * Hs[c]; w1 += sg[c] * Hs[13 + c]; }
        const double k0 = w0 * Sinv[0] + w1 * Sinv[2];
        const double k1 = w0 * Sinv[1] + w1 * Sinv[3];
        mu_i[i] = mu[i] + (k0 * inn[0] + k1 * inn[1]);
      }
    }
    cluster.sync();                                   // mu_i complete
    if (tid == 0) {
      for (int c = 0; c < 3; ++c) rr[c] = mu_i[c];
      double q[4] = {mu_i[3], mu_i[4], mu_i[5], mu_i[6]};
      const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      double qc[4];
      qc[0] = q[0] / qn; qc[1] = -(q[1] / qn); qc[2] = -(q[2] / qn); qc[3] = -(q[3] / qn);
      double R[9];
      d_quat2rot(qc, R);
      for (int c = 0; c < 9; ++c) Rcw[c] = R[c];
    }
    __syncthreads();
    int local = 0;
    for (int start = rank * nthr; start < N; start += ncta * nthr) {
      const int i = start + tid;
      int flag = 0;
      if (i < N && ft.innov[i]) {
        const int pos = ft.pos[i], coding = ft.coding[i];
        double fs[6], hi[2], r3[3] = {rr[0], rr[1], rr[2]}, R[9];
        for (int c = 0; c < 9; ++c) R[c] = Rcw[c];
        if (!coding) for (int c = 0; c < 6; ++c) fs[c] = mu_i[pos + c];
        else for (int c = 0; c < 3; ++c) fs[c] = mu[pos + c];  // quirk V:1016: mu, not mu_i
        d_feature_h(cfg.cam, fs, coding, r3, R, hi);
        const double e0 = ft.z[2 * i] - hi[0], e1 = ft.z[2 * i + 1] - hi[1];
        flag = (sqrt(e0 * e0 + e1 * e1) <= cfg.th_low) ? 1 : 0;
        ft.li[i] = flag;
      }
      local += __syncthreads_count(flag);
    }
    if (tid == 0 && local > 0) atomicAdd(&ctl->rs_count, local);
    cluster.sync();                                   // the count is complete
    if (tid == 0) {
      const int actual = ctl->rs_count;
      if (actual > s_numzli) {
        s_numzli = actual;
        s_nhyp = (int)(log(1 - cfg.ransac_p) / (log(1 - (actual / (matched + 0.0)))));  // V:1030
      }
    }
    ++it;
    cluster.sync();                                   // everybody has read the count before CTA 0 clears it for the next pick
  }
  cluster.sync();
  if (rank == 0) {
    const int nli = block_compact(ft.li, N, ft.sel, ft.pos_in_z);  // V:1040-1048
    if (tid == 0) {
      ctl->ransac_hyps = it;
      ctl->n_li = nli;
      ctl->k_rows = 2 * nli;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// High-innovation rescue (V:1066-1130): camera pose from the OLD mu (ctl->cam_old), feature
// parameters from mu_tmp (= mu after the low-innovation update), S_hi = H Sigma_tmp H^T without R,
// chi^2 <= th_hi.  The last block compacts the Hi list.
// ------------------------------------------------------------------------------------------------
// Book-keeping of one feature (Patch::update_quality_index, Patch.cpp:143-150, and the flags of the packed result record) and the
// record's header: shared by k_bookkeeping and by the tail of k_hi_rescue (which does both itself when nothing was rescued).
__device__ __forceinline__ void bookkeeping_feature(int i, FeatTab ft, int N, const DevCfg& cfg, int* __restrict__ outi) {
  int nfind = ft.n_find[i];
  const int ntot = ft.n_tot[i];
  if (ft.hi[i] || ft.li[i]) nfind++;
  ft.n_find[i] = nfind;
  const float qi = (float)(ntot - nfind) / ((float)nfind);
  ft.quality[i] = qi;
  if (qi > cfg.quality_ratio) ft.removef[i] = 1;
  const int fl = (ft.innov[i] ? 1 : 0) | (ft.li[i] ? 2 : 0) | (ft.hi[i] ? 4 : 0) | (ft.removef[i] ? 8 : 0);
  outi[16 + i] = fl;
  outi[16 + N + i] = ntot;
  outi[16 + 2 * N + i] = nfind;
}
__device__ __forceinline__ void bookkeeping_record(const double* __restrict__ Sigma, int ld, const double* __restrict__ mu, const DevCtl* ctl,
                                                   double* __restrict__ outd, int* __restrict__ outi) {
  for (int e = threadIdx.x; e < 14; e += blockDim.x) outd[e] = mu[e];
  for (int e = threadIdx.x; e < 196; e += blockDim.x) outd[14 + e] = Sigma[(size_t)(e / 14) * ld + (e % 14)];
  if (threadIdx.x == 0) {
    outi[0] = ctl->m_innov; outi[1] = ctl->n_matched; outi[2] = ctl->n_li; outi[3] = ctl->n_hi;
    outi[4] = ctl->ransac_hyps; outi[5] = ctl->chol_fail; outi[6] = ctl->blur_count; outi[7] = ctl->blur_too_large;
  }
}

__global__ void __launch_bounds__(128) k_hi_rescue(const double* __restrict__ Sigma, int ld, const double* __restrict__ mu,
                                                   FeatTab ft, int N, DevCtl* ctl, DevCfg cfg, double* __restrict__ outd = nullptr,
                                                   int* __restrict__ outi = nullptr) {
  __shared__ int is_last;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
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
      for (int b = 0; b < nd; ++b) {
        const int jb = ekf_idx13(b, pos);
        double t0 = 0, t1 = 0;
        for (int c = 0; c < nd; ++c) {
          const double s = Sigma[(size_t)ekf_idx13(c, pos) * ld + jb];
          t0 += Hc[c] * s; t1 += Hc[13 + c] * s;
        }
        Tm[b] = t0; Tm[13 + b] = t1;
      }
      double S[4] = {0, 0, 0, 0};
      for (int b = 0; b < nd; ++b) {
        S[0] += Tm[b] * Hc[b]; S[1] += Tm[b] * Hc[13 + b];
        S[2] += Tm[13 + b] * Hc[b]; S[3] += Tm[13 + b] * Hc[13 + b];
      }
      // fixed-size 2x2 inverse(): closed form (V:1114)
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
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&ctl->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int nhi = block_compact(ft.hi, N, ft.sel, ft.pos_in_z);  // V:1122-1130
  if (threadIdx.x == 0) {
    ctl->n_hi = nhi;
    ctl->k_rows = 2 * nhi;
    ctl->ticket = 0;
  }
  // outd != null: nothing rescued means no second update, so the step's book-keeping and result record are final now — done here
  // by the last CTA instead of one more launch (when something WAS rescued the host runs the second update and k_bookkeeping)
  if (outd && nhi == 0) {
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) bookkeeping_feature(j, ft, N, cfg, outi);
    bookkeeping_record(Sigma, ld, mu, ctl, outd, outi);
  }
}

// ------------------------------------------------------------------------------------------------
// Peer-memory exchange (row-block partition, look-ahead pipeline).  A producing kernel stores its panel into
// every peer's buffer over NVLink, its last CTA publishes the block's epoch in every peer's flag word, and the
// consumer spins on its OWN flag words (bounded: a timeout sets ctl->chol_fail |= 16 instead of hanging).
// ------------------------------------------------------------------------------------------------
#define P2P_TIMEOUT_CYCLES (1ll << 32)   // ~2 s at 1.965 GHz
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// called by every thread of a kernel after its peer stores; the last CTA to arrive publishes `epoch` in slot
// `slot0 + rank` of every rank's flag words.  ticket: device counter, self-resetting.
__device__ __forceinline__ void p2p_publish(const P2PView& pv, int slot0, unsigned int* ticket) {
  __shared__ int last_cta;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    last_cta = (t == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!last_cta) return;
  __threadfence_system();
  if (threadIdx.x < pv.world) st_release_sys(pv.flags[threadIdx.x] + slot0 + pv.rank, pv.epoch);
  if (threadIdx.x == 0) *ticket = 0;
}
// lanes 0 .. world-1 of the calling warp wait until slot0 + r of the LOCAL flag words reached `epoch`
__device__ __forceinline__ void p2p_wait(const unsigned long long* local_flags, int slot0, int world, unsigned long long epoch, DevCtl* ctl) {
  const int lane = threadIdx.x & 31;
  if (lane < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(local_flags + slot0 + lane) < epoch) {
      if (clock64() - t0 > P2P_TIMEOUT_CYCLES) { atomicOr(&ctl->chol_fail, 16); break; }
      __nanosleep(100);
    }
  }
  __syncwarp();
}
__global__ void k_p2p_publish_only(P2PView pv, int slot0) {   // a rank without rows still has to report
  __threadfence_system();
  if (threadIdx.x < pv.world) st_release_sys(pv.flags[threadIdx.x] + slot0 + pv.rank, pv.epoch);
}
__global__ void k_p2p_wait(const unsigned long long* local_flags, int slot0, int world, unsigned long long epoch, DevCtl* ctl) {
  p2p_wait(local_flags, slot0, world, epoch, ctl);
}

// ------------------------------------------------------------------------------------------------
// K4a: W_b = Sigma H_b^T for the block of <= 64 selected features starting at sel[f0]
// (n x EKF_UB, row-major), and nu_b = (z - h) - H_b delta.  One pass over the needed columns of Sigma.
// ------------------------------------------------------------------------------------------------
// rows of Sigma per CTA.  Every CTA first stages the block's 64 x 26 measurement Jacobians (13 KB), so more rows per CTA would
// amortise that — measured the other way round (cfg2, frames/s): 4 rows 1 270, 8 rows 1 264, 16 rows 1 237, 32 rows 1 196: the
// kernel lives on parallelism (754 CTAs at n = 3014), not on the staging.  EKF_GATHER_ROWS overrides (multiple of 4).
static int gather_rows() {
  static const int v = [] { const char* e = getenv("EKF_GATHER_ROWS"); const int r = e ? atoi(e) : 4; return (r >= 4 && r % 4 == 0) ? r : 4; }();
  return v;
}
__global__ void __launch_bounds__(256) k_blk_gather(const double* __restrict__ Sigma, int ld, int row0, int n, FeatTab ft, int f0,
                                                    int cnt, const double* __restrict__ delta, double* __restrict__ W,
                                                    double* __restrict__ nu, double* __restrict__ W2, int rows_per_cta,
                                                    BlkTab bt = BlkTab{nullptr, nullptr, nullptr, nullptr}, const int* __restrict__ cnt_dev = nullptr,
                                                    unsigned int* pub_ticket = nullptr, unsigned int* pub_flag = nullptr, unsigned int pub_token = 0,
                                                    int hot_rows = 0) {
  // hot_rows (needs bt): the launch covers only the rows the block's own S_b reads — the 7 camera rows and the block's feature rows,
  // row index r of the launch -> state row r (r < 7) or pos[(r - 7) / 6] + (r - 7) % 6 — written at their true row index.
  if (cnt_dev) cnt = *cnt_dev;   // launched before the host has read the count back (see ekf_update_after_match)
  // Sigma is read past L1 (ld.global.cg): in the chain-short schedule this kernel runs BESIDE the downdate that still writes the
  // tiles it does not read (k_wait_tiles gates it), so no line of Sigma may be served from a stale L1 copy.
  __shared__ double Hs[EKF_UB / 2][27];
  __shared__ int poss[EKF_UB / 2], nds[EKF_UB / 2], fids[EKF_UB / 2];
  const int nb = min(EKF_UB / 2, cnt - f0);
  const int tid = threadIdx.x;
  if (bt.H) {   // prepared tables: one round trip instead of three
    for (int a = tid; a < EKF_UB / 2; a += blockDim.x) {
      fids[a] = (a < nb) ? 0 : -1;
      poss[a] = (a < nb) ? bt.pos[f0 + a] : 0;
      nds[a] = (a < nb) ? bt.nd[f0 + a] : 0;
    }
    for (int e = tid; e < (EKF_UB / 2) * 26; e += blockDim.x) {
      const int a = e / 26, c = e % 26;
      Hs[a][c] = (a < nb) ? bt.H[26 * (size_t)f0 + e] : 0.0;
    }
  } else {
    for (int a = tid; a < EKF_UB / 2; a += blockDim.x) {
      if (a < nb) {
        const int f = ft.sel[f0 + a];
        fids[a] = f; poss[a] = ft.pos[f]; nds[a] = 7 + (ft.coding[f] ? 3 : 6);
      } else { fids[a] = -1; poss[a] = 0; nds[a] = 0; }
    }
    __syncthreads();
    for (int e = tid; e < (EKF_UB / 2) * 26; e += blockDim.x) {
      const int a = e / 26, c = e % 26;
      Hs[a][c] = (a < nb) ? ft.Hc[26 * fids[a] + c] : 0.0;
    }
  }
  __syncthreads();
  const int a = tid & (EKF_UB / 2 - 1), rl = tid / (EKF_UB / 2);
  const int rstep = 256 / (EKF_UB / 2);
  const int pos = poss[a], nd = nds[a];
#pragma unroll 2
  for (int rq = rl; rq < rows_per_cta; rq += rstep) {
    int i = row0 + blockIdx.x * rows_per_cta + rq;   // rows [row0, n): the caller's row block
    if (hot_rows) {
      const int r = blockIdx.x * rows_per_cta + rq;
      if (r >= 7 + 6 * nb) break;
      if (r >= 7) {
        const int fa = (r - 7) / 6, k = (r - 7) % 6;
        if (7 + k >= nds[fa]) continue;   // an XYZ feature has three rows
        i = poss[fa] + k;
      } else i = r;
    }
    if (i >= n) break;   // (no barrier inside this loop: the publish below is reached by every thread)
    const double* row = Sigma + (size_t)i * ld;
    double sg[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) sg[c] = (c < nd) ? __ldcg(row + ekf_idx13(c, pos)) : 0.0;   // 13 independent loads in flight
    double w0 = 0, w1 = 0;
#pragma unroll
    for (int c = 0; c < 13; ++c)
      if (c < nd) { w0 += sg[c] * Hs[a][c]; w1 += sg[c] * Hs[a][13 + c]; }
    reinterpret_cast<double2*>(W + (size_t)i * EKF_UB)[a] = make_double2(w0, w1);
    if (W2) reinterpret_cast<double2*>(W2 + (size_t)i * EKF_UB)[a] = make_double2(w0, w1);   // second copy: see launch_blk_gather2
  }
  if (pub_flag) chain_publish_last_cta(pub_ticket, pub_flag, pub_token);   // resident-chain schedule: W' is complete
  if (nu && blockIdx.x == 0 && tid < EKF_UB / 2) {
    double v0 = 0, v1 = 0;
    if (tid < nb) {
      const int f = fids[tid];
      double hd0 = 0, hd1 = 0;
      for (int c = 0; c < nds[tid]; ++c) {
        const double d = delta[ekf_idx13(c, poss[tid])];
        hd0 += Hs[tid][c] * d; hd1 += Hs[tid][13 + c] * d;
      }
      v0 = (ft.z[2 * f] - ft.h[2 * f]) - hd0;
      v1 = (ft.z[2 * f + 1] - ft.h[2 * f + 1]) - hd1;
    }
    nu[2 * tid] = v0; nu[2 * tid + 1] = v1;
  }
}

// ------------------------------------------------------------------------------------------------
// K4b(1): S_b = H_b W_b + sigma_px^2 I (EKF_UB x EKF_UB, row-major), one CTA per row.  Unused rows of
// a partial block are identity so they contribute nothing downstream.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EKF_UB) k_blk_S(const double* __restrict__ W, FeatTab ft, int f0, int cnt,
                                                  double sigma_pixel_2, double* __restrict__ Sb, int plain,
                                                  const double* __restrict__ delta, double* __restrict__ nu,
                                                  const double* __restrict__ Gsub, int row0 = 0, int row1 = 0x7fffffff,
                                                  int add_diag = 1, P2PView pv = P2PView{}, unsigned int* ticket = nullptr) {
  // plain == 0: S_b = H_b W + sigma_px^2 I (identity past the block's rows).
  // plain != 0: G = H_b W with zero padding — W then holds the PREVIOUS block's V (look-ahead correction).
  // nu != null: CTA 0 also forms nu_b = (z - h) - H_b delta.
  // [row0, row1), add_diag: row-block partition (ekf_dist) — only the rows of W this rank owns contribute (the partial
  // S blocks are summed by an all-reduce) and one rank adds the diagonal term.
  // Gsub != null: W is the UNCORRECTED gather W' of the look-ahead pipeline and S_b = H_b W' + R - G G^T with
  // G = H_b V_prev (H_b (W' - V_prev G^T) = H_b W' - G G^T), so that S_b does not wait for the correction of W.
  __shared__ double Gr[EKF_UB];
  const int r = blockIdx.x, s = threadIdx.x;
  if (Gsub) { Gr[s] = Gsub[r * EKF_UB + s]; __syncthreads(); }
  const int nb = min(EKF_UB / 2, cnt - f0), kr = 2 * nb;
  double v = (!plain && r == s && add_diag) ? 1.0 : 0.0;
  if (r < kr && (plain || s < kr)) {
    const int f = ft.sel[f0 + (r >> 1)];
    const int pos = ft.pos[f], nd = 7 + (ft.coding[f] ? 3 : 6);
    const double* hc = ft.Hc + 26 * f + 13 * (r & 1);
    double acc = 0;
    // all 13 loads of W and of H in flight before the first use (a rolled loop paid the L2 latency per term: this kernel is
    // the G step on the critical chain of every update block); same terms, same order
    double hv[13], wv[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) {
      const int idx = ekf_idx13(c < nd ? c : 0, pos);
      const bool in = c < nd && idx >= row0 && idx < row1;
      hv[c] = in ? hc[c] : 0.0;
      wv[c] = in ? W[(size_t)idx * EKF_UB + s] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 13; ++c) {
      const int idx = ekf_idx13(c < nd ? c : 0, pos);
      if (c < nd && idx >= row0 && idx < row1) acc += hv[c] * wv[c];
    }
    v = acc + ((!plain && r == s && add_diag) ? sigma_pixel_2 : 0.0);
    if (Gsub) {
      // thread s streams row s of G (16-byte loads, 16 in flight) against row r in shared memory; a transposed copy of G
      // read coalesced over s was measured SLOWER (25 vs 15 us: twice the load instructions in this latency-bound kernel)
      const double2* gs = reinterpret_cast<const double2*>(Gsub + (size_t)s * EKF_UB);
      double g0 = 0, g1 = 0, g2 = 0, g3 = 0;
#pragma unroll 8
      for (int k = 0; k < EKF_UB / 2; k += 2) {
        const double2 t = gs[k], u = gs[k + 1];
        g0 += Gr[2 * k] * t.x; g1 += Gr[2 * k + 1] * t.y;
        g2 += Gr[2 * k + 2] * u.x; g3 += Gr[2 * k + 3] * u.y;
      }
      v -= (g0 + g1) + (g2 + g3);
    }
  }
  Sb[r * EKF_UB + s] = v;
  if (ticket) {   // the partial block goes into slot `rank` of every rank's partial-S buffer
    for (int q = 0; q < pv.world; ++q) pv.spart[q][(size_t)pv.rank * EKF_UB * EKF_UB + r * EKF_UB + s] = v;
  }
  if (nu && r == 0) {
    double out = 0.0;
    if (s < kr) {
      const int f = ft.sel[f0 + (s >> 1)];
      const int pos = ft.pos[f], nd = 7 + (ft.coding[f] ? 3 : 6);
      const double* hc = ft.Hc + 26 * f + 13 * (s & 1);
      double hd = 0;
      for (int c = 0; c < nd; ++c) hd += hc[c] * delta[ekf_idx13(c, pos)];
      out = (ft.z[2 * f + (s & 1)] - ft.h[2 * f + (s & 1)]) - hd;
    }
    nu[s] = out;
  }
  if (ticket) p2p_publish(pv, 0, ticket);
}

// ------------------------------------------------------------------------------------------------
// K4b(1'): S_b = H_b W'_b + sigma_px^2 I - G G^T for the factor-beside-downdate schedule, tiled.  k_blk_S forms the G G^T term
// with one thread per entry streaming a whole row of G (128 KB per CTA, 32-byte sectors half used: 16 us on the critical chain
// of every block).  Here one CTA owns a 32 x 32 block of the LOWER triangle (10 CTAs; the factor kernels read nothing else):
// the two 32-row slabs of G it needs are staged in shared memory with coalesced 16-byte loads and multiplied on the tensor pipe
// (DMMA.8x8x4, each warp an 8 x 32 strip), and the gather H_b W'_b (13 rows of W' per measurement row) lands directly in the
// accumulator layout.  CTA (0, 0) also forms nu_b = (z - h) - H_b delta.
// ------------------------------------------------------------------------------------------------
#define S2_LD (EKF_UB + 4)   // shared-memory row stride of a G slab: fragment loads (8 rows x 4 columns) hit every bank pair twice
__global__ void __launch_bounds__(128) k_blk_S_tiled(const double* __restrict__ W, FeatTab ft, int f0, int cnt, double sigma_pixel_2,
                                                     double* __restrict__ Sb, const double* __restrict__ delta, double* __restrict__ nu,
                                                     const double* __restrict__ Gsub, const double* __restrict__ gy = nullptr,
                                                     BlkTab bt = BlkTab{nullptr, nullptr, nullptr, nullptr}, const double* __restrict__ Sg = nullptr,
                                                     unsigned int* pub_ticket = nullptr, unsigned int* pub_flag = nullptr, unsigned int pub_token = 0,
                                                     const double* __restrict__ Sg2 = nullptr) {
  // Sg2 != null: W' was gathered one block earlier still and -G2 G2^T (G2 = H_b V_{b-2}) is the second correction term.
  // Sg != null (chain-short schedule): the term -G G^T was formed ahead of time by k_blk_Sg while the gather of W' was still
  // running; this launch then only adds the 13-row gather H_b W' and R — two round trips on the critical cycle instead of five.
  extern __shared__ __align__(16) double s2sm[];
  // lower-triangle block index -> (bi, bj), bi >= bj, 4 x 4 blocks of 32
  int bi = 0, rem = blockIdx.x;
  while (rem > bi) { rem -= bi + 1; ++bi; }
  const int bj = rem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const int nb = min(EKF_UB / 2, cnt - f0), kr = 2 * nb;
  double* Ga = s2sm;                    // rows 32 bi .. of G
  double* Gb = s2sm + 32 * S2_LD;       // rows 32 bj .. of G
  if (Gsub) {
    // both slabs with cp.async, every 16-byte copy of a thread in flight at once (a load -> store loop paid the L2 latency
    // 32 times in a row: 10 of the kernel's 14 us under ncu); the gather below runs while they land
    for (int e = tid; e < 32 * (EKF_UB / 2); e += 128) {
      const int r = e / (EKF_UB / 2), c = (e % (EKF_UB / 2)) * 2;
      const unsigned sa = (unsigned)__cvta_generic_to_shared(Ga + r * S2_LD + c), sb = (unsigned)__cvta_generic_to_shared(Gb + r * S2_LD + c);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(Gsub + (size_t)(32 * bi + r) * EKF_UB + c));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sb), "l"(Gsub + (size_t)(32 * bj + r) * EKF_UB + c));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }
  // gather part, straight into the accumulator layout: lane (g, t4) of warp q holds row 32 bi + 8 q + g, columns 32 bj + 8 ct + 2 t4 (+1)
  const int r = 32 * bi + 8 * warp + g;
  double acc[4][2];
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) { acc[ct][0] = 0.0; acc[ct][1] = 0.0; }
  double2 sgv[4];
  if (Sg) {
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) sgv[ct] = *reinterpret_cast<const double2*>(Sg + (size_t)r * EKF_UB + 32 * bj + 8 * ct + 2 * t4);
    if (Sg2) {
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        const double2 v2 = *reinterpret_cast<const double2*>(Sg2 + (size_t)r * EKF_UB + 32 * bj + 8 * ct + 2 * t4);
        sgv[ct].x += v2.x; sgv[ct].y += v2.y;
      }
    }
  }
  if (r < kr) {
    int pos, nd;
    const double* hc;
    if (bt.H) { pos = bt.pos[f0 + (r >> 1)]; nd = bt.nd[f0 + (r >> 1)]; hc = bt.H + 26 * (size_t)(f0 + (r >> 1)) + 13 * (r & 1); }
    else {
      const int f = ft.sel[f0 + (r >> 1)];
      pos = ft.pos[f]; nd = 7 + (ft.coding[f] ? 3 : 6);
      hc = ft.Hc + 26 * f + 13 * (r & 1);
    }
    double hv[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) hv[c] = (c < nd) ? hc[c] : 0.0;
    // 13 x 4 independent 16-byte loads per lane, issued before the first use (unrolled: one L2 latency, not thirteen)
#pragma unroll
    for (int c = 0; c < 13; ++c) {
      const double* wrow = W + (size_t)ekf_idx13(c < nd ? c : 0, pos) * EKF_UB + 32 * bj + 2 * t4;
      double2 v[4];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) v[ct] = *reinterpret_cast<const double2*>(wrow + 8 * ct);
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        if (c < nd) { acc[ct][0] += hv[c] * v[ct].x; acc[ct][1] += hv[c] * v[ct].y; }
      }
    }
  }
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int sc = 32 * bj + 8 * ct + 2 * t4 + u;
      // rows / columns past the block's measurements: identity (they contribute nothing downstream)
      if (r >= kr || sc >= kr) acc[ct][u] = (r == sc) ? 1.0 : 0.0;
      else if (r == sc) acc[ct][u] += sigma_pixel_2;
    }
  }
  if (Gsub) {
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    const double* ga = Ga + (8 * warp + g) * S2_LD + t4;
    double e[4][2];
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) { e[ct][0] = 0.0; e[ct][1] = 0.0; }
#pragma unroll 8
    for (int k4 = 0; k4 < EKF_UB / 4; ++k4) {
      const double a = -ga[4 * k4];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) dmma884f(e[ct][0], e[ct][1], a, Gb[(8 * ct + g) * S2_LD + 4 * k4 + t4]);
    }
#pragma unroll
    for (int ct = 0; ct < 4; ++ct)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int sc = 32 * bj + 8 * ct + 2 * t4 + u;
        if (r < kr && sc < kr) acc[ct][u] += e[ct][u];
      }
  }
  if (Sg) {
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      const int sc = 32 * bj + 8 * ct + 2 * t4;
      if (r < kr && sc < kr) acc[ct][0] += sgv[ct].x;
      if (r < kr && sc + 1 < kr) acc[ct][1] += sgv[ct].y;
    }
  }
#pragma unroll
  for (int ct = 0; ct < 4; ++ct)
    *reinterpret_cast<double2*>(Sb + (size_t)r * EKF_UB + 32 * bj + 8 * ct + 2 * t4) = make_double2(acc[ct][0], acc[ct][1]);
  if (nu && blockIdx.x == 0) {
    const int s = tid;
    double out = 0.0;
    if (s < kr && bt.H) {
      const int j = f0 + (s >> 1), pos = bt.pos[j], nd = bt.nd[j];
      const double* hc = bt.H + 26 * (size_t)j + 13 * (s & 1);
      double hv[13], dv[13];
#pragma unroll
      for (int c = 0; c < 13; ++c) { hv[c] = (c < nd) ? hc[c] : 0.0; dv[c] = (c < nd) ? delta[ekf_idx13(c, pos)] : 0.0; }
      double hd = 0;
#pragma unroll
      for (int c = 0; c < 13; ++c) if (c < nd) hd += hv[c] * dv[c];
      out = bt.zmh[2 * j + (s & 1)] - hd;
      if (gy) out -= gy[s];
    } else if (s < kr) {
      const int f = ft.sel[f0 + (s >> 1)];
      const int pos = ft.pos[f], nd = 7 + (ft.coding[f] ? 3 : 6);
      const double* hc = ft.Hc + 26 * f + 13 * (s & 1);
      double hd = 0;
      for (int c = 0; c < nd; ++c) hd += hc[c] * delta[ekf_idx13(c, pos)];
      out = (ft.z[2 * f + (s & 1)] - ft.h[2 * f + (s & 1)]) - hd;
      // gy = G_b y_{b-1} = H_b V_{b-1} y_{b-1}: `delta` then lacks the previous block's term (chain-short schedule)
      if (gy) out -= gy[s];
    }
    nu[s] = out;
  }
  if (pub_flag) chain_publish_last_cta(pub_ticket, pub_flag, pub_token);
}

// K4b(1-fin): S_b = H_b W'_b + R + Sg (+ Sg2) and nu_b for the chain-short schedule — the part of S_b that has to wait for the gather.
// k_blk_S_tiled in its Sg mode did this with 8 entries and 52 16-byte loads per lane; here one thread owns ONE column pair (13 loads,
// all in flight behind a single table look-up) and nu_b gets its own CTA, so the kernel is two round trips deep whatever its size:
// 10 CTAs x 512 threads for the lower 32 x 32 blocks + 1 CTA for nu.
__global__ void __launch_bounds__(512) k_blk_S_fin(const double* __restrict__ W, int f0, int cnt, double sigma_pixel_2, double* __restrict__ Sb,
                                                   const double* __restrict__ delta, double* __restrict__ nu, const double* __restrict__ gy,
                                                   BlkTab bt, const double* __restrict__ Sg, const double* __restrict__ Sg2,
                                                   unsigned int* pub_ticket = nullptr, unsigned int* pub_flag = nullptr, unsigned int pub_token = 0) {
  const int nb = min(EKF_UB / 2, cnt - f0), kr = 2 * nb, tid = threadIdx.x;
  if (blockIdx.x == 10) {   // nu_b = (z - h) - H_b delta - G_b y
    if (tid < EKF_UB) {
      double out = 0.0;
      if (tid < kr) {
        const int j = f0 + (tid >> 1), pos = bt.pos[j], nd = bt.nd[j];
        const double* hc = bt.H + 26 * (size_t)j + 13 * (tid & 1);
        double hv[13], dv[13];
#pragma unroll
        for (int c = 0; c < 13; ++c) { hv[c] = (c < nd) ? hc[c] : 0.0; dv[c] = (c < nd) ? delta[ekf_idx13(c, pos)] : 0.0; }
        double hd = 0;
#pragma unroll
        for (int c = 0; c < 13; ++c) if (c < nd) hd += hv[c] * dv[c];
        out = bt.zmh[2 * j + (tid & 1)] - hd;
        if (gy) out -= gy[tid];
      }
      nu[tid] = out;
    }
    if (pub_flag) chain_publish_last_cta(pub_ticket, pub_flag, pub_token);
    return;
  }
  int bi = 0, rem = blockIdx.x;
  while (rem > bi) { rem -= bi + 1; ++bi; }
  const int r = 32 * bi + (tid >> 4), c = 32 * rem + 2 * (tid & 15);
  double v0 = (r == c) ? 1.0 : 0.0, v1 = (r == c + 1) ? 1.0 : 0.0;   // rows / columns past the block's measurements: identity
  if (r < kr && c < kr) {
    const int j = f0 + (r >> 1), pos = bt.pos[j], nd = bt.nd[j];
    const double* hc = bt.H + 26 * (size_t)j + 13 * (r & 1);
    double2 g2 = make_double2(0.0, 0.0), g3 = make_double2(0.0, 0.0);
    if (Sg) g2 = *reinterpret_cast<const double2*>(Sg + (size_t)r * EKF_UB + c);
    if (Sg2) g3 = *reinterpret_cast<const double2*>(Sg2 + (size_t)r * EKF_UB + c);
    double2 wv[13];
    double hv[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) {
      hv[q] = (q < nd) ? hc[q] : 0.0;
      wv[q] = (q < nd) ? *reinterpret_cast<const double2*>(W + (size_t)ekf_idx13(q, pos) * EKF_UB + c) : make_double2(0.0, 0.0);
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int q = 0; q < 13; ++q) if (q < nd) { s0 += hv[q] * wv[q].x; s1 += hv[q] * wv[q].y; }
    if (r == c) s0 += sigma_pixel_2;
    if (r == c + 1) s1 += sigma_pixel_2;
    if (Sg) { s0 += g2.x; s1 += g2.y; }
    if (Sg2) { s0 += g3.x; s1 += g3.y; }
    v0 = s0; v1 = s1;   // kr is even and c is even: c + 1 < kr
  }
  *reinterpret_cast<double2*>(Sb + (size_t)r * EKF_UB + c) = make_double2(v0, v1);
  if (pub_flag) chain_publish_last_cta(pub_ticket, pub_flag, pub_token);
}

// K4b(1''): Sg = -G G^T on the lower 32 x 32 blocks (10 CTAs, DMMA), the part of S_b that does not need the gather of W'_b: in
// the chain-short schedule it runs right after k_blk_Gx, in the shadow of the downdate / gather the block waits for.
__global__ void __launch_bounds__(128) k_blk_Sg(const double* __restrict__ Gsub, double* __restrict__ Sg, unsigned int* pub_ticket = nullptr,
                                                unsigned int* pub_flag = nullptr, unsigned int pub_token = 0) {
  extern __shared__ __align__(16) double s2sm[];
  int bi = 0, rem = blockIdx.x;
  while (rem > bi) { rem -= bi + 1; ++bi; }
  const int bj = rem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  double* Ga = s2sm;
  double* Gb = s2sm + 32 * S2_LD;
  for (int e = tid; e < 32 * (EKF_UB / 2); e += 128) {
    const int r = e / (EKF_UB / 2), c = (e % (EKF_UB / 2)) * 2;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(Ga + r * S2_LD + c), sb = (unsigned)__cvta_generic_to_shared(Gb + r * S2_LD + c);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(Gsub + (size_t)(32 * bi + r) * EKF_UB + c));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sb), "l"(Gsub + (size_t)(32 * bj + r) * EKF_UB + c));
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  const double* ga = Ga + (8 * warp + g) * S2_LD + t4;
  // two accumulator sets per column tile (even / odd k-steps): 16 dependent DMMAs per chain instead of 32
  double e0[4][2], e1[4][2];
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) { e0[ct][0] = e0[ct][1] = e1[ct][0] = e1[ct][1] = 0.0; }
#pragma unroll 4
  for (int k4 = 0; k4 < EKF_UB / 4; k4 += 2) {
    const double a0 = -ga[4 * k4], a1 = -ga[4 * k4 + 4];
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      dmma884f(e0[ct][0], e0[ct][1], a0, Gb[(8 * ct + g) * S2_LD + 4 * k4 + t4]);
      dmma884f(e1[ct][0], e1[ct][1], a1, Gb[(8 * ct + g) * S2_LD + 4 * k4 + 4 + t4]);
    }
  }
  const int r = 32 * bi + 8 * warp + g;
#pragma unroll
  for (int ct = 0; ct < 4; ++ct)
    *reinterpret_cast<double2*>(Sg + (size_t)r * EKF_UB + 32 * bj + 8 * ct + 2 * t4) = make_double2(e0[ct][0] + e1[ct][0], e0[ct][1] + e1[ct][1]);
  if (pub_flag) chain_publish_last_cta(pub_ticket, pub_flag, pub_token);
}

// Compact tables of the selected features for the block kernels (BlkTab), once per stacked update.
__global__ void __launch_bounds__(128) k_blk_prep(FeatTab ft, int cnt, double* __restrict__ H, double* __restrict__ zmh, int* __restrict__ pos,
                                                  int* __restrict__ nd, const int* __restrict__ cnt_dev) {
  if (cnt_dev) cnt = *cnt_dev;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= cnt) return;
  const int f = ft.sel[j];
  if (lane < 26) H[26 * (size_t)j + lane] = ft.Hc[26 * f + lane];
  if (lane < 2) zmh[2 * j + lane] = ft.z[2 * f + lane] - ft.h[2 * f + lane];
  if (lane == 2) pos[j] = ft.pos[f];
  if (lane == 3) nd[j] = 7 + (ft.coding[f] ? 3 : 6);
}

// ------------------------------------------------------------------------------------------------
// forsePlane (V:1245-1263, 1272): three pseudo-measurements mu[1] = mu[4] = mu[6] = 0 with R = 1e-5 I3,
// appended to the second stacked update.  Handled as one more 128-row block whose first three rows are
// the unit rows e1, e4, e6: W[:, j] = Sigma[:, k_j], nu_j = -(mu[k_j] + delta[k_j]), S = that 3x3 block of
// W + 1e-5 I (identity elsewhere).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_plane_gather(const double* __restrict__ Sigma, int ld, int row0, int n,
                                                      const double* __restrict__ mu, const double* __restrict__ delta,
                                                      double* __restrict__ W, double* __restrict__ nu, DevCtl* ctl) {
  const int kidx[3] = {1, 4, 6};
  const int i = row0 + blockIdx.x, s = threadIdx.x;
  if (i < n) W[(size_t)i * EKF_UB + s] = (s < 3) ? Sigma[(size_t)i * ld + kidx[s]] : 0.0;
  if (blockIdx.x == 0) {
    nu[s] = (s < 3) ? (0.0 - mu[kidx[s]]) - delta[kidx[s]] : 0.0;
    if (s == 0) ctl->k_rows += 3;
  }
}
__global__ void __launch_bounds__(EKF_UB) k_plane_S(const double* __restrict__ W, double* __restrict__ Sb) {
  const int kidx[3] = {1, 4, 6};
  const int r = blockIdx.x, s = threadIdx.x;
  double v = (r == s) ? 1.0 : 0.0;
  if (r < 3 && s < 3) v = W[(size_t)kidx[r] * EKF_UB + s] + ((r == s) ? 0.00001 : 0.0);
  Sb[r * EKF_UB + s] = v;
}
void launch_plane_gather(cudaStream_t st, const double* Sigma, int ld, int row0, int row1, const double* mu, const double* delta,
                         double* W, double* nu, DevCtl* ctl, long long* launches) {
  const int nr = row1 > row0 ? row1 - row0 : 1;
  k_plane_gather<<<nr, 128, 0, st>>>(Sigma, ld, row0, row1, mu, delta, W, nu, ctl);
  *launches += 1;
}
void launch_plane_S(cudaStream_t st, const double* W, double* Sb, long long* launches) {
  k_plane_S<<<EKF_UB, EKF_UB, 0, st>>>(W, Sb);
  *launches += 1;
}

#include "ekf_factor.cuh"
#include "ekf_chol128.cuh"
// K4b: Cholesky gain of one 128-row block.  Default: cta_chol128 (ekf_chol128.cuh: the matrix lives in registers in the DMMA
// accumulator layout, 8-column steps, two barriers per step).  k_blk_factor_smem is the round-1 kernel (matrix in shared
// memory, bound by shared-memory bandwidth), kept selectable with EKF_CHOL_SMEM=1 for A/B timing and as a cross-check.
__global__ void __launch_bounds__(CH_THREADS, 1) k_blk_factor(const double* __restrict__ Sb, const double* __restrict__ nu,
                                                             double* __restrict__ Lout, double* __restrict__ Dblk,
                                                             double* __restrict__ yout, DevCtl* ctl) {
  extern __shared__ __align__(16) double fsm[];
  cta_chol128(fsm, Sb, EKF_UB, nu, Lout, EKF_UB, Dblk, 32, yout, &ctl->chol_fail);
}
// Pre-positioned factor (EKF_SCHED=3): launched on its own stream BEFORE S_b exists, so that it takes over the SM the previous
// block's factor kernel just gave back, and waits (flag word, acquire, bounded) for the last CTA of the S_b kernel.
__global__ void __launch_bounds__(CH_THREADS, 1) k_blk_factor_wait(const double* Sb, const double* nu, double* __restrict__ Lout,
                                                                  double* __restrict__ Dblk, double* __restrict__ yout, DevCtl* ctl,
                                                                  const unsigned int* flag, unsigned int token) {
  extern __shared__ __align__(16) double fsm[];
  if (threadIdx.x == 0 && !chain_wait(flag, token)) atomicOr(&ctl->chol_fail, 64 | 512);
  __syncthreads();
  cta_chol128(fsm, Sb, EKF_UB, nu, Lout, EKF_UB, Dblk, 32, yout, &ctl->chol_fail);
}
__global__ void __launch_bounds__(FACT_THREADS) k_blk_factor_smem(const double* __restrict__ Sb, const double* __restrict__ nu,
                                                                  double* __restrict__ Lout, double* __restrict__ Dblk,
                                                                  double* __restrict__ yout, DevCtl* ctl) {
  extern __shared__ __align__(16) double fsm[];
  cta_chol_panel<EKF_UB>(fsm, Sb, EKF_UB, nu, Lout, EKF_UB, Dblk, 32, yout, &ctl->chol_fail);
}

// Resident-chain schedule (ekf_api.cu::stacked_update_resident_chain): ONE CTA stays resident for the whole stacked update and
// factors every block.  Per block it waits (flag word, acquire) until k_blk_S_tiled has published S_b and nu_b, factors with
// cta_chol128 and publishes L_b, D_b, y_b.  Because its SM is never given back between blocks, the downdate no longer has to be
// held back until the next factor kernel has found a free SM (the release rule of the other two schedules), and the factor starts
// without a launch.  (A first version also assembled S_b inside this CTA — the 13-row gather of 10 K entries by one SM took 12 us,
// as long as the ten-CTA kernel plus its launch gap: profiles/r2k_trace_resident_chain.txt.)
__global__ void __launch_bounds__(CH_THREADS, 1) k_chain_factor(ChainFactorArgs a, DevCtl* ctl) {
  extern __shared__ __align__(16) double fsm[];
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int b = 0; b < a.nblk; ++b) {
    const int p = b & 1;
    if (tid == 0 && !chain_wait(a.fl.sg + b, a.fl.token0 + b)) atomicOr(&ctl->chol_fail, 64 | 512 | (b << 16));
    __syncthreads();
    cta_chol128(fsm, a.S, EKF_UB, a.nu, a.L[p], EKF_UB, a.D[p], 32, a.y[p], &ctl->chol_fail);
    __threadfence();   // every thread: its stores of L, D, y are performed before the flag goes up
    __syncthreads();
    if (tid == 0) chain_publish(a.fl.fact + b, a.fl.token0 + b);
  }
}
// one-warp gate: the kernels behind it in the stream start once *flag == token (the n-row solve V_b behind "factor_b is published")
__global__ void k_wait_flag(const unsigned int* flag, unsigned int token, DevCtl* ctl) {
  if (threadIdx.x == 0 && !chain_wait(flag, token)) atomicOr(&ctl->chol_fail, 64 | 2048);
}
void launch_chain_factor(cudaStream_t st, const ChainFactorArgs& a, DevCtl* ctl, long long* launches) {
  k_chain_factor<<<1, CH_THREADS, sizeof(Chol128Smem), st>>>(a, ctl);
  *launches += 1;
}
void launch_wait_flag(cudaStream_t st, const unsigned int* flag, unsigned int token, DevCtl* ctl, long long* launches) {
  k_wait_flag<<<1, 32, 0, st>>>(flag, token, ctl);
  *launches += 1;
}

// Row-block partition: S_b = sum over ranks of the partial blocks the peers stored into this rank's slots (fixed
// order: every rank forms the same bits), once every peer's epoch has arrived; then the same factorisation.
__global__ void __launch_bounds__(CH_THREADS, 1) k_blk_factor_p2p(const double* __restrict__ spart, const unsigned long long* __restrict__ flags,
                                                                 int world, unsigned long long epoch, double* __restrict__ Ssum,
                                                                 const double* __restrict__ nu, double* __restrict__ Lout,
                                                                 double* __restrict__ Dblk, double* __restrict__ yout, DevCtl* ctl) {
  extern __shared__ __align__(16) double fsm[];
  if (threadIdx.x < 32) p2p_wait(flags, 0, world, epoch, ctl);
  __syncthreads();
  for (int e = threadIdx.x; e < EKF_UB * EKF_UB; e += CH_THREADS) {
    double s = 0;
    for (int q = 0; q < world; ++q) s += __ldcg(spart + (size_t)q * EKF_UB * EKF_UB + e);
    Ssum[e] = s;
  }
  __syncthreads();   // Ssum was written by this CTA: visible to all of its threads after the barrier
  cta_chol128(fsm, Ssum, EKF_UB, nu, Lout, EKF_UB, Dblk, 32, yout, &ctl->chol_fail);
}

// ------------------------------------------------------------------------------------------------
// K4c: V_b = W_b L^-T in place (blocked triangular solve on the fp64 tensor pipe, DMMA.8x8x4) and
// delta += V_b y.  One CTA = 32 rows; each of its 4 warps owns one 8-row tile for the whole solve
// (warp_trsm_tile, ekf_factor.cuh), so there is no block barrier after the operands are staged.
// ------------------------------------------------------------------------------------------------
#define VT_ROWS 32
#define VT_THREADS 256   // 8 warps stage the operands; warps 0..3 each own one 8-row tile of the solve
#define VT_LD (EKF_UB + 4)
#define VT_LDD 36
// L in shared memory PACKED: the solve reads of block row J (rows 32 J ..) only columns 0 .. 32 J - 1, stored with row stride
// 32 J + 4 (fragment loads stay conflict-free: stride = 4 mod 16 doubles).  121 KB per CTA instead of 207 KB and <= 64 registers
// per thread: a CTA now fits into what ONE retiring downdate CTA leaves behind (the downdate runs 4 CTAs x 16 K registers x
// 20 KB per SM) instead of waiting for the downdate's last wave to drain.
#define VT_LPACK (32 * 36 + 32 * 68 + 32 * 100)   // doubles: block rows 1, 2, 3
__global__ void __launch_bounds__(VT_THREADS, 4) k_blk_V(double* __restrict__ W, int rbase, int n, const double* __restrict__ Lg,
                                                      const double* __restrict__ Dg, const double* __restrict__ yg,
                                                      double* __restrict__ delta, P2PView pv = P2PView{}, unsigned int* ticket = nullptr,
                                                      double* __restrict__ Vout = nullptr, const double* __restrict__ delta_in = nullptr) {
  // Vout != null: V_b goes there and W stays as it was (the chain-short schedule reads rows of W_b for the next block's G while
  // this kernel runs); delta_in != null: delta = delta_in + V_b y out of place (the next S_b reads delta_in meanwhile).
  extern __shared__ __align__(16) double vsm[];
  double* Lst = vsm;                         // packed block rows 1 .. 3 of L
  const double* const Lj[4] = {Lst, Lst, Lst + 32 * 36, Lst + 32 * 36 + 32 * 68};   // block row 0 is never read
  const int ldj[4] = {36, 36, 68, 100};
  double* Ds = Lst + VT_LPACK;               // [EKF_UB / 32][32][VT_LDD] inverses of the diagonal blocks
  double* Ws = Ds + (EKF_UB / 32) * 32 * VT_LDD;  // [VT_ROWS][VT_LD] W rows, then V rows
  double* ys = Ws + VT_ROWS * VT_LD;         // [EKF_UB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = rbase + blockIdx.x * VT_ROWS;   // rows [rbase, n): the caller's row block
  // everything is staged with cp.async, all copies in flight at once (the kernel is latency-bound: 95 CTAs, one wave).
  // Of L only the 32 x 32 blocks strictly below the diagonal are read by the solve — the diagonal blocks enter
  // through their inverses D — so 48 KB instead of 128 KB per CTA.
  auto cp16 = [](double* sdst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
  };
  for (int e = tid; e < VT_ROWS * EKF_UB / 2; e += VT_THREADS) {
    const int r = e >> 6, c = (e & 63) * 2;
    if (row0 + r < n) cp16(Ws + r * VT_LD + c, W + (size_t)(row0 + r) * EKF_UB + c);
    else *reinterpret_cast<double2*>(Ws + r * VT_LD + c) = make_double2(0.0, 0.0);
  }
  for (int e = tid; e < (EKF_UB / 32) * 32 * 16; e += VT_THREADS) {
    const int r = e >> 4, c = (e & 15) * 2;   // r = block * 32 + row
    cp16(Ds + r * VT_LDD + c, Dg + r * 32 + c);
  }
  for (int e = tid; e < (EKF_UB - 32) * (EKF_UB - 32) / 2; e += VT_THREADS) {
    const int r = 32 + e / ((EKF_UB - 32) / 2), c = (e % ((EKF_UB - 32) / 2)) * 2;
    if (c < (r & ~31)) {
      const int J = r >> 5;                                  // 1 .. 3 (arithmetic, not Lj[J]: keeps the tables in registers)
      const int offJ = J == 1 ? 0 : (J == 2 ? 32 * 36 : 32 * 36 + 32 * 68);
      cp16(Lst + offJ + (r & 31) * (32 * J + 4) + c, Lg + r * EKF_UB + c);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  for (int e = tid; e < EKF_UB; e += VT_THREADS) ys[e] = yg[e];
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  // (A version with 16-row CTAs and four warps per tile — warps_trsm_tile_rl, as in k_blk_Gx — staged L and D twice as often
  // and measured no faster inside the step: cfg2 0.738 against 0.728 ms, cfg4's V class 1.63 against 1.41 ms.)
  if (warp < VT_ROWS / 8) {
    double part = warp_trsm_tile_packed<EKF_UB>(Ws + (size_t)warp * 8 * VT_LD, VT_LD, Lj, ldj, Ds, VT_LDD, ys);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    const int i = row0 + warp * 8 + (lane >> 2);
    if ((lane & 3) == 0 && i < n && delta) delta[i] = (delta_in ? delta_in[i] : delta[i]) + part;
  }
  __syncthreads();
  double* const dst = Vout ? Vout : W;
  for (int e = tid; e < VT_ROWS * EKF_UB / 2; e += VT_THREADS) {
    const int r = e >> 6, c = (e & 63) * 2;
    if (row0 + r < n) {
      const double2 v = *reinterpret_cast<const double2*>(Ws + r * VT_LD + c);
      *reinterpret_cast<double2*>(dst + (size_t)(row0 + r) * EKF_UB + c) = v;
      if (ticket) {   // all-gather fused into the solve: the finished rows go straight into every peer's panel over NVLink
        for (int q = 0; q < pv.world; ++q)
          if (q != pv.rank) *reinterpret_cast<double2*>(pv.w[q] + (size_t)(row0 + r) * EKF_UB + c) = v;
      }
    }
  }
  if (ticket) p2p_publish(pv, 8, ticket);
}

// ------------------------------------------------------------------------------------------------
// K4b(0): G_b = H_b V_{b-1} WITHOUT waiting for V_{b-1} (chain-short schedule).  V_{b-1} = W_{b-1} L^-T, so
//   G_b = (H_b W_{b-1}) L^-T :  X = H_b W_{b-1} reads only the 7 camera rows and the block's own feature rows of the (corrected)
// panel W_{b-1}, and the triangular solve is the one of k_blk_V on 128 rows instead of n.  With it the n-row solve V_{b-1}
// leaves the critical chain  factor_{b-1} -> G_b -> S_b -> factor_b  and runs beside factor_b on its own stream.
// One CTA = 8 rows of G (4 features): the <= 31 rows of W it needs, L (packed), the diagonal-block inverses and y arrive by
// cp.async, all in flight at once and none through registers (the CTA must start in what ONE retiring downdate CTA leaves:
// 16 K registers); X is formed from shared memory and ONE warp solves the tile.  Also gy = G_b y_{b-1}, the term of nu_b that
// delta does not hold yet.
// ------------------------------------------------------------------------------------------------
#define GX_FEATS 4
#define GX_WROWS (7 + 6 * GX_FEATS)
__global__ void __launch_bounds__(VT_THREADS, 4) k_blk_Gx(const double* __restrict__ Wc, FeatTab ft, int f0, int cnt,
                                                          const double* __restrict__ Lg, const double* __restrict__ Dg,
                                                          const double* __restrict__ yg, double* __restrict__ G, double* __restrict__ gy,
                                                          BlkTab bt = BlkTab{nullptr, nullptr, nullptr, nullptr},
                                                          const unsigned int* wait_flag = nullptr, unsigned int wait_token = 0, DevCtl* ctl = nullptr) {
  extern __shared__ __align__(16) double gxsm[];
  __shared__ int poss[GX_FEATS], nds[GX_FEATS], fids[GX_FEATS];
  if (wait_flag) {   // resident-chain schedule: L, D, y of the previous block come from k_chain_factor, not from a kernel ahead in the stream
    if (threadIdx.x == 0 && !chain_wait(wait_flag, wait_token) && ctl) atomicOr(&ctl->chol_fail, 64 | 1024);
    __syncthreads();
  }
  double* Lst = gxsm;
  const double* const Lj[4] = {Lst, Lst, Lst + 32 * 36, Lst + 32 * 36 + 32 * 68};
  const int ldj[4] = {36, 36, 68, 100};
  double* Ds = Lst + VT_LPACK;
  double* Xs = Ds + (EKF_UB / 32) * 32 * VT_LDD;   // [8][VT_LD]: X rows, then G rows
  double* ys = Xs + 8 * VT_LD;                     // [EKF_UB]
  double* Wr = ys + EKF_UB;                        // [GX_WROWS][EKF_UB]: camera rows 0..6 of W, then 6 rows per feature
  double* Hs = Wr + GX_WROWS * EKF_UB;             // [GX_FEATS][26]
  double* Gs = Hs + GX_FEATS * 26 + 2;             // [8][VT_LD]: G rows (output tile of the cooperative solve; 16-byte aligned)
  double* red = Gs + 8 * VT_LD;                    // [4][8] partial G y per warp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = min(EKF_UB / 2, cnt - f0);
  const int a0 = blockIdx.x * GX_FEATS;
  if (tid < GX_FEATS) {
    const int a = a0 + tid;
    if (a < nb && bt.H) { fids[tid] = 0; poss[tid] = bt.pos[f0 + a]; nds[tid] = bt.nd[f0 + a]; }
    else if (a < nb) {
      const int f = ft.sel[f0 + a];
      fids[tid] = f; poss[tid] = ft.pos[f]; nds[tid] = 7 + (ft.coding[f] ? 3 : 6);
    } else { fids[tid] = -1; poss[tid] = 0; nds[tid] = 0; }
  }
  auto cp16 = [](double* sdst, const double* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
  };
  // operands that do not depend on the feature table first: their copies fly while the table look-ups return
  for (int e = tid; e < (EKF_UB / 32) * 32 * 16; e += VT_THREADS) {
    const int r = e >> 4, c = (e & 15) * 2;
    cp16(Ds + r * VT_LDD + c, Dg + r * 32 + c);
  }
  for (int e = tid; e < (EKF_UB - 32) * (EKF_UB - 32) / 2; e += VT_THREADS) {
    const int r = 32 + e / ((EKF_UB - 32) / 2), c = (e % ((EKF_UB - 32) / 2)) * 2;
    if (c < (r & ~31)) {
      const int J = r >> 5;
      const int offJ = J == 1 ? 0 : (J == 2 ? 32 * 36 : 32 * 36 + 32 * 68);
      cp16(Lst + offJ + (r & 31) * (32 * J + 4) + c, Lg + r * EKF_UB + c);
    }
  }
  for (int e = tid; e < 7 * (EKF_UB / 2); e += VT_THREADS) {
    const int r = e >> 6, c = (e & 63) * 2;
    cp16(Wr + r * EKF_UB + c, Wc + (size_t)r * EKF_UB + c);
  }
  for (int e = tid; e < EKF_UB; e += VT_THREADS) ys[e] = yg[e];
  __syncthreads();
  for (int e = tid; e < 6 * GX_FEATS * (EKF_UB / 2); e += VT_THREADS) {
    const int rr = e >> 6, c = (e & 63) * 2, a = rr / 6, k = rr % 6;
    double* d = Wr + (7 + rr) * EKF_UB + c;
    if (7 + k < nds[a]) cp16(d, Wc + (size_t)(poss[a] + k) * EKF_UB + c);
    else *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  for (int e = tid; e < GX_FEATS * 26; e += VT_THREADS)
    Hs[e] = fids[e / 26] < 0 ? 0.0 : (bt.H ? bt.H[26 * (size_t)(f0 + a0) + e] : ft.Hc[26 * fids[e / 26] + e % 26]);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  // X = H_b W_{b-1}: 8 rows x 64 column pairs, two per thread
  for (int e = tid; e < 8 * (EKF_UB / 2); e += VT_THREADS) {
    const int r = e >> 6, c = (e & 63) * 2, a = r >> 1;
    const double* hr = Hs + a * 26 + 13 * (r & 1);
    const int nd = nds[a];
    double x0 = 0.0, x1 = 0.0;
#pragma unroll
    for (int q = 0; q < 13; ++q) {
      if (q < nd) {
        const double2 v = *reinterpret_cast<const double2*>(Wr + (q < 7 ? q : 7 + 6 * a + (q - 7)) * EKF_UB + c);
        x0 += hr[q] * v.x; x1 += hr[q] * v.y;
      }
    }
    *reinterpret_cast<double2*>(Xs + r * VT_LD + c) = make_double2(x0, x1);
  }
  __syncthreads();
  if (warp < 4) {   // four warps solve the tile together (warps_trsm_tile_rl: <= 8 dependent DMMAs per 32-column block)
    double part = warps_trsm_tile_rl<EKF_UB, 4, 1>(Xs, Gs, VT_LD, Lj, ldj, Ds, VT_LDD, ys, warp);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if ((lane & 3) == 0) red[8 * warp + (lane >> 2)] = part;
  }
  __syncthreads();
  if (tid < 8) gy[8 * blockIdx.x + tid] = (red[tid] + red[8 + tid]) + (red[16 + tid] + red[24 + tid]);
  for (int e = tid; e < 8 * (EKF_UB / 2); e += VT_THREADS) {
    const int r = e >> 6, c = (e & 63) * 2;
    *reinterpret_cast<double2*>(G + (size_t)(8 * blockIdx.x + r) * EKF_UB + c) = *reinterpret_cast<const double2*>(Gs + r * VT_LD + c);
  }
}

// Chain-short schedule, gather beside the downdate: for every update block g one CTA lists the T (T + 1) / 2 tiles (64 x 64) of the
// lower triangle with the tiles the gather of block g reads FIRST — tile (tm, tn) is hot when tm or tn is an index range holding
// the camera columns or one of the block's features (with the mirrored store these are all entries of those columns) — and counts
// them.  k_gemm_nt_sub walks the list; the gather waits for the count (k_blk_gather, hot_counter).
__global__ void __launch_bounds__(1024) k_blk_tile_order(FeatTab ft, int cnt, int T, ushort2* __restrict__ order, int* __restrict__ n_hot,
                                                         unsigned int* __restrict__ hot_counters) {
  __shared__ unsigned long long m[2];
  __shared__ int c_hot, c_cold;
  const int g = blockIdx.x, tid = threadIdx.x, f0 = g * (EKF_UB / 2);
  if (tid < 2) m[tid] = (tid == 0) ? 1ull : 0ull;   // columns 0 .. 6 live in range 0
  if (tid == 0) { c_hot = 0; c_cold = 0; hot_counters[g] = 0u; }
  __syncthreads();
  if (tid < EKF_UB / 2 && f0 + tid < cnt) {
    const int f = ft.sel[f0 + tid];
    const int pos = ft.pos[f], last = pos + (ft.coding[f] ? 2 : 5);
    for (int t = pos >> 6; t <= (last >> 6); ++t) atomicOr(&m[(t >> 6) & 1], 1ull << (t & 63));
  }
  __syncthreads();
  const int L = T * (T + 1) / 2;
  ushort2* out = order + (size_t)g * L;
  for (int idx = tid; idx < L; idx += blockDim.x) {
    int tm = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((tm + 1) * (tm + 2) / 2 <= idx) ++tm;
    while (tm * (tm + 1) / 2 > idx) --tm;
    const int tn = idx - tm * (tm + 1) / 2;
    const bool hot = ((m[(tm >> 6) & 1] >> (tm & 63)) | (m[(tn >> 6) & 1] >> (tn & 63))) & 1ull;
    const int at = hot ? atomicAdd(&c_hot, 1) : L - 1 - atomicAdd(&c_cold, 1);
    out[at] = make_ushort2((unsigned short)tm, (unsigned short)tn);
  }
  __syncthreads();
  if (tid == 0) n_hot[g] = c_hot;
}
void launch_blk_tile_order(cudaStream_t st, FeatTab ft, int cnt, int T, ushort2* order, int* n_hot, unsigned int* hot_counters, long long* launches) {
  const int nblk = (cnt + EKF_UB / 2 - 1) / (EKF_UB / 2);
  if (nblk <= 0) return;
  k_blk_tile_order<<<nblk, 1024, 0, st>>>(ft, cnt, T, order, n_hot, hot_counters);
  *launches += 1;
}

// delta += V y for ALL rows, one warp per row (row-block partition: every rank runs this on the all-gathered V so that
// the replicas of delta stay bit-identical; the owner-only accumulation inside k_blk_V is switched off there)
__global__ void __launch_bounds__(256) k_delta_rows(const double* __restrict__ V, const double* __restrict__ y, double* __restrict__ delta,
                                                    int n) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const double2* v = reinterpret_cast<const double2*>(V + (size_t)row * EKF_UB);
  const double2* yy = reinterpret_cast<const double2*>(y);
  double s = 0;
#pragma unroll
  for (int k = 0; k < EKF_UB / 64; ++k) {
    const double2 a = v[lane + 32 * k], b = yy[lane + 32 * k];
    s += a.x * b.x + a.y * b.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) delta[row] += s;
}


// ------------------------------------------------------------------------------------------------
// Book-keeping (V:1296-1303, Patch.cpp:143-150) and the packed per-step output record.
//   out: [0,14) state, [14,210) Sigma 14x14 (as doubles), then ints: 16 counters, 3N per feature
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bookkeeping(const double* __restrict__ Sigma, int ld, const double* __restrict__ mu,
                                                     FeatTab ft, int N, DevCtl* ctl, DevCfg cfg, double* __restrict__ outd,
                                                     int* __restrict__ outi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) bookkeeping_feature(i, ft, N, cfg, outi);
  if (blockIdx.x == 0) bookkeeping_record(Sigma, ld, mu, ctl, outd, outi);
}

// ---- launch wrappers ---------------------------------------------------------------------------
static const size_t kFactSmemOld = (size_t)cta_chol_panel_smem_doubles<EKF_UB>() * sizeof(double);
static const size_t kFactSmem = sizeof(Chol128Smem);
static int g_chol_smem = 0;   // EKF_CHOL_SMEM=1: the round-1 shared-memory factor kernel
static const size_t kVSmem = (size_t)(VT_LPACK + (EKF_UB / 32) * 32 * VT_LDD + VT_ROWS * VT_LD + EKF_UB) * sizeof(double);
static const size_t kGxSmem = (size_t)(VT_LPACK + (EKF_UB / 32) * 32 * VT_LDD + 8 * VT_LD + EKF_UB + GX_WROWS * EKF_UB + GX_FEATS * 26 + 2 + 8 * VT_LD + 32) * sizeof(double);

int update_kernels_init() {
  const char* env = getenv("EKF_CHOL_SMEM");
  g_chol_smem = env ? atoi(env) : 0;
  cudaError_t e = cudaFuncSetAttribute(k_blk_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_factor_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactSmemOld);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_S_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 32 * S2_LD * sizeof(double)));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_Sg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 32 * S2_LD * sizeof(double)));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_factor_wait, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_chain_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_factor_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_V, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_blk_Gx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGxSmem);
  if (e != cudaSuccess) return (int)e;
  // CUDA loads a kernel's code lazily at its first launch, and that load can wait for running kernels to finish: a kernel launched
  // for the first time while k_chain_factor spins on a flag only that launch can raise would never start (measured: every wait of the
  // first resident-chain update ran into its time-out).  cudaFuncGetAttributes forces the load now.
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, k_chain_factor)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_wait_flag)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_Gx)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_Sg)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_V)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_gather)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_prep)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_S_tiled)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_S_fin)) != cudaSuccess) return (int)e;
  if ((e = cudaFuncGetAttributes(&fa, k_blk_factor_wait)) != cudaSuccess) return (int)e;
  return gemm_kernels_preload();
}

void launch_ransac(cudaStream_t st, const double* Sigma, int ld, int n, const double* mu, FeatTab ft, int N, DevCtl* ctl,
                   const DevCfg& cfg, const uint32_t* picks, int n_picks, double* mu_i, int* cand, long long* launches) {
  const char* e = getenv("EKF_RANSAC_CLUSTER");   // read per launch (tests switch kernels inside one process)
  const int use_cluster = e ? atoi(e) : 1;
  if (use_cluster && n >= RANSAC_CLUSTER_MIN_N) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(RANSAC_CL); lc.blockDim = dim3(RANSAC_CL_THREADS); lc.dynamicSmemBytes = 0; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = RANSAC_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (cudaLaunchKernelEx(&lc, k_ransac_cluster, Sigma, ld, n, mu, ft, N, ctl, cfg, picks, n_picks, mu_i, cand) == cudaSuccess) {
      *launches += 1;
      return;
    }
    cudaGetLastError();   // fall through to the one-CTA kernel
  }
  k_ransac<<<1, RANSAC_THREADS, 0, st>>>(Sigma, ld, n, mu, ft, N, ctl, cfg, picks, n_picks, mu_i, cand);
  *launches += 1;
}
void launch_hi_rescue(cudaStream_t st, const double* Sigma, int ld, const double* mu, FeatTab ft, int N, DevCtl* ctl,
                      const DevCfg& cfg, long long* launches, double* outd, int* outi) {
  const int fb = N > 0 ? (N + 127) / 128 : 1;
  k_hi_rescue<<<fb, 128, 0, st>>>(Sigma, ld, mu, ft, N, ctl, cfg, outd, outi);
  *launches += 1;
}
void launch_blk_gather(cudaStream_t st, const double* Sigma, int ld, int row0, int row1, FeatTab ft, int f0, int cnt,
                       const double* delta, double* W, double* nu, long long* launches) {
  const int nr = row1 > row0 ? row1 - row0 : 1;   // at least one CTA: block 0 also forms nu
  const int gr = gather_rows();
  k_blk_gather<<<(nr + gr - 1) / gr, 256, 0, st>>>(Sigma, ld, row0, row1, ft, f0, cnt, delta, W, nu, nullptr, gr);
  *launches += 1;
}
void launch_blk_factor(cudaStream_t st, const double* W, FeatTab ft, int f0, int cnt, const double* nu, const DevCfg& cfg,
                       double* Sb, double* Lb, double* Dblk, double* yb, DevCtl* ctl, long long* launches) {
  k_blk_S<<<EKF_UB, EKF_UB, 0, st>>>(W, ft, f0, cnt, cfg.sigma_pixel_2, Sb, 0, nullptr, nullptr, nullptr);
  *launches += 1;
  launch_blk_factor_only(st, Sb, nu, Lb, Dblk, yb, ctl, launches);
}
// look-ahead variants: S_b together with nu_b (delta is current only now), and G = H_b V_prev
void launch_blk_S_nu(cudaStream_t st, const double* W, FeatTab ft, int f0, int cnt, const DevCfg& cfg, const double* delta,
                     double* Sb, double* nu, long long* launches) {
  k_blk_S<<<EKF_UB, EKF_UB, 0, st>>>(W, ft, f0, cnt, cfg.sigma_pixel_2, Sb, 0, delta, nu, nullptr);
  *launches += 1;
}
// S_b from the uncorrected gather and G (see k_blk_S), and the gather with a second copy of W'
void launch_blk_S_nu_G(cudaStream_t st, const double* Wraw, FeatTab ft, int f0, int cnt, const DevCfg& cfg, const double* delta,
                       const double* G, double* Sb, double* nu, long long* launches, const double* gy, BlkTab bt, const double* Sg,
                       unsigned int* pub_ticket, unsigned int* pub_flag, unsigned int pub_token, const double* Sg2) {
  static const bool legacy = [] { const char* e = getenv("EKF_S_TILED"); return e && atoi(e) == 0; }();
  static const bool s_fin = [] { const char* e = getenv("EKF_S_FIN"); return !(e && atoi(e) == 0); }();
  if (s_fin && Sg && bt.H && nu) {
    k_blk_S_fin<<<11, 512, 0, st>>>(Wraw, f0, cnt, cfg.sigma_pixel_2, Sb, delta, nu, gy, bt, Sg, Sg2, pub_ticket, pub_flag, pub_token);
    *launches += 1;
    return;
  }
  if (legacy && !gy && !Sg && !pub_flag) k_blk_S<<<EKF_UB, EKF_UB, 0, st>>>(Wraw, ft, f0, cnt, cfg.sigma_pixel_2, Sb, 0, delta, nu, G);
  else k_blk_S_tiled<<<10, 128, Sg ? 0 : 2 * 32 * S2_LD * sizeof(double), st>>>(Wraw, ft, f0, cnt, cfg.sigma_pixel_2, Sb, delta, nu, Sg ? nullptr : G, gy, bt, Sg,
                                                                               pub_ticket, pub_flag, pub_token, Sg2);
  *launches += 1;
}
void launch_blk_Sg(cudaStream_t st, const double* G, double* Sg, long long* launches, unsigned int* pub_ticket, unsigned int* pub_flag,
                   unsigned int pub_token) {
  k_blk_Sg<<<10, 128, 2 * 32 * S2_LD * sizeof(double), st>>>(G, Sg, pub_ticket, pub_flag, pub_token);
  *launches += 1;
}
void launch_blk_prep(cudaStream_t st, FeatTab ft, int cnt, double* H, double* zmh, int* pos, int* nd, long long* launches, const int* cnt_dev) {
  if (cnt <= 0) return;   // with cnt_dev: cnt is the upper bound that sizes the grid
  k_blk_prep<<<(cnt + 3) / 4, 128, 0, st>>>(ft, cnt, H, zmh, pos, nd, cnt_dev);
  *launches += 1;
}
// G_b = (H_b W_{b-1}) L_{b-1}^-T and gy = G_b y_{b-1} (see k_blk_Gx)
void launch_blk_Gx(cudaStream_t st, const double* Wc, FeatTab ft, int f0, int cnt, const double* Lb, const double* Dblk, const double* yb,
                   double* G, double* gy, long long* launches, BlkTab bt, const unsigned int* wait_flag, unsigned int wait_token, DevCtl* ctl) {
  k_blk_Gx<<<EKF_UB / 8, VT_THREADS, kGxSmem, st>>>(Wc, ft, f0, cnt, Lb, Dblk, yb, G, gy, bt, wait_flag, wait_token, ctl);
  *launches += 1;
}
void launch_blk_gather2(cudaStream_t st, const double* Sigma, int ld, int n, FeatTab ft, int f0, int cnt, double* W, double* W2,
                        long long* launches, BlkTab bt, const int* cnt_dev, unsigned int* pub_ticket, unsigned int* pub_flag,
                        unsigned int pub_token) {
  const int nr = n > 0 ? n : 1;
  const int gr = gather_rows();
  k_blk_gather<<<(nr + gr - 1) / gr, 256, 0, st>>>(Sigma, ld, 0, n, ft, f0, cnt, nullptr, W, nullptr, W2, gr, bt, cnt_dev, pub_ticket, pub_flag, pub_token);
  *launches += 1;
}
// Gate of the gather that runs beside a downdate: one warp waits (acquire, bounded) until the downdate's first *n_hot tiles — the
// ones holding the columns the gather reads — are stored (k_blk_tile_order / k_gemm_nt_sub); the gather follows it in stream order.
// A gate instead of a wait inside the gather: a spinning gather grid would hold registers the downdate's own CTAs need.
__global__ void k_wait_tiles(const unsigned int* hot_counter, const int* __restrict__ n_hot, DevCtl* ctl) {
  if (threadIdx.x != 0) return;
  const unsigned int want = (unsigned int)*n_hot;
  unsigned int seen = 0;
  for (long long spins = 0;; ++spins) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(hot_counter) : "memory");
    if (seen >= want) break;
    if (spins > 4000000ll) { atomicOr(&ctl->chol_fail, 64); break; }   // seconds: report instead of hanging
    __nanosleep(100);
  }
}
void launch_blk_gather2_after_tiles(cudaStream_t st, const double* Sigma, int ld, int n, FeatTab ft, int f0, int cnt, double* W, double* W2,
                                    const unsigned int* hot_counter, const int* n_hot, DevCtl* ctl, long long* launches, BlkTab bt) {
  k_wait_tiles<<<1, 32, 0, st>>>(hot_counter, n_hot, ctl);
  *launches += 1;
  launch_blk_gather2(st, Sigma, ld, n, ft, f0, cnt, W, W2, launches, bt, nullptr, nullptr, nullptr, 0);
}
// the rows of W' = Sigma H_b^T that S_b itself reads (7 camera rows + the block's feature rows), at their true row index of W
void launch_blk_gather_hot(cudaStream_t st, const double* Sigma, int ld, int n, FeatTab ft, int f0, int cnt, double* W, long long* launches, BlkTab bt,
                           const int* cnt_dev) {
  const int rows = 7 + 6 * (EKF_UB / 2), gr = 4;
  k_blk_gather<<<(rows + gr - 1) / gr, 256, 0, st>>>(Sigma, ld, 0, n, ft, f0, cnt, nullptr, W, nullptr, nullptr, gr, bt, cnt_dev, nullptr, nullptr, 0, 1);
  *launches += 1;
}
void launch_blk_G(cudaStream_t st, const double* Vprev, FeatTab ft, int f0, int cnt, double* G, long long* launches) {
  k_blk_S<<<EKF_UB, EKF_UB, 0, st>>>(Vprev, ft, f0, cnt, 0.0, G, 1, nullptr, nullptr, nullptr);
  *launches += 1;
}
void launch_blk_factor_wait(cudaStream_t st, const double* Sb, const double* nu, double* Lb, double* Dblk, double* yb, DevCtl* ctl,
                            const unsigned int* flag, unsigned int token, long long* launches) {
  k_blk_factor_wait<<<1, CH_THREADS, kFactSmem, st>>>(Sb, nu, Lb, Dblk, yb, ctl, flag, token);
  *launches += 1;
}
void launch_blk_factor_only(cudaStream_t st, const double* Sb, const double* nu, double* Lb, double* Dblk, double* yb, DevCtl* ctl,
                            long long* launches) {
  if (g_chol_smem) k_blk_factor_smem<<<1, FACT_THREADS, kFactSmemOld, st>>>(Sb, nu, Lb, Dblk, yb, ctl);
  else k_blk_factor<<<1, CH_THREADS, kFactSmem, st>>>(Sb, nu, Lb, Dblk, yb, ctl);
  *launches += 1;
}
void launch_blk_V(cudaStream_t st, double* W, int row0, int row1, const double* Lb, const double* Dblk, const double* yb,
                  double* delta, long long* launches, double* Vout, const double* delta_in) {
  if (row1 <= row0) return;
  k_blk_V<<<(row1 - row0 + VT_ROWS - 1) / VT_ROWS, VT_THREADS, kVSmem, st>>>(W, row0, row1, Lb, Dblk, yb, delta, P2PView{}, nullptr, Vout, delta_in);
  *launches += 1;
}
void launch_blk_S_part(cudaStream_t st, const double* W, FeatTab ft, int f0, int cnt, const DevCfg& cfg, int row0, int row1, int add_diag,
                       const double* delta, double* nu, double* Sb, long long* launches) {
  k_blk_S<<<EKF_UB, EKF_UB, 0, st>>>(W, ft, f0, cnt, cfg.sigma_pixel_2, Sb, 0, delta, nu, nullptr, row0, row1, add_diag);
  *launches += 1;
}
// peer-memory variants of the partitioned look-ahead update (pv: the peer mappings of this block, passed by value)
void launch_blk_S_part_p2p(cudaStream_t st, const double* W, FeatTab ft, int f0, int cnt, const DevCfg& cfg, int row0, int row1, int add_diag,
                           const double* delta, double* nu, double* Sb, const P2PView& pv, unsigned int* ticket, long long* launches) {
  k_blk_S<<<EKF_UB, EKF_UB, 0, st>>>(W, ft, f0, cnt, cfg.sigma_pixel_2, Sb, 0, delta, nu, nullptr, row0, row1, add_diag, pv, ticket);
  *launches += 1;
}
void launch_blk_factor_p2p(cudaStream_t st, const double* spart, const unsigned long long* flags, int world, unsigned long long epoch,
                           double* Ssum, const double* nu, double* Lb, double* Dblk, double* yb, DevCtl* ctl, long long* launches) {
  k_blk_factor_p2p<<<1, CH_THREADS, kFactSmem, st>>>(spart, flags, world, epoch, Ssum, nu, Lb, Dblk, yb, ctl);
  *launches += 1;
}
void launch_blk_V_p2p(cudaStream_t st, double* W, int row0, int row1, const double* Lb, const double* Dblk, const double* yb,
                      const P2PView& pv, unsigned int* ticket, const unsigned long long* flags, int world, unsigned long long epoch,
                      DevCtl* ctl, long long* launches) {
  if (row1 > row0) {
    k_blk_V<<<(row1 - row0 + VT_ROWS - 1) / VT_ROWS, VT_THREADS, kVSmem, st>>>(W, row0, row1, Lb, Dblk, yb, nullptr, pv, ticket);
  } else {
    k_p2p_publish_only<<<1, 32, 0, st>>>(pv, 8);
  }
  *launches += 1;
  k_p2p_wait<<<1, 32, 0, st>>>(flags, 8, world, epoch, ctl);   // every rank's rows of V_b have landed in the local panel
  *launches += 1;
}
void launch_delta_rows(cudaStream_t st, const double* V, const double* y, double* delta, int n, long long* launches) {
  k_delta_rows<<<(n + 7) / 8, 256, 0, st>>>(V, y, delta, n);
  *launches += 1;
}
void launch_bookkeeping(cudaStream_t st, const double* Sigma, int ld, const double* mu, FeatTab ft, int N, DevCtl* ctl,
                        const DevCfg& cfg, double* outd, int* outi, long long* launches) {
  const int fb = N > 0 ? (N + 255) / 256 : 1;
  k_bookkeeping<<<fb, 256, 0, st>>>(Sigma, ld, mu, ft, N, ctl, cfg, outd, outi);
  *launches += 1;
}
