// csrc/ekf_cta.cuh — CTA-cooperative pieces of VSlamFilter::update shared by the single-filter kernels
// (ekf_update.cu) and the fused batched-filter kernel (ekf_batch.cu).  V: = mono-slam/src/vslamRansac.cpp.
#pragma once
#include "ekf_common.cuh"
#include "ekf_math.cuh"

// K5: 1-point RANSAC (V:964-1034) by ONE CTA (any multiple of 32 threads): the whole adaptive loop
// runs on the device.  picks[] replaces rand() (V:970,989).  The Li flags left behind are those of
// the LAST hypothesis evaluated (quirk V:1022), and S_i equals the 2x2 block computed in predict
// (same Sigma, same H).  mu_i: n doubles, cand: N ints of scratch (shared or global).
static __device__ __noinline__ void cta_ransac(const double* __restrict__ Sigma, int ld, int n, const double* __restrict__ mu, FeatTab ft,
                                        int N, DevCtl* ctl, const DevCfg& cfg, const uint32_t* __restrict__ picks, int n_picks,
                                        double* mu_i, int* cand) {
  const int nthr = blockDim.x;
  __shared__ double Hs[26], Sinv[4], inn[2], rr[3], Rcw[9];
  __shared__ int s_p, s_sel, s_pos, s_nd, s_nhyp, s_numzli;
  const int tid = threadIdx.x;
  int cnt = block_compact(ft.innov, N, cand, nullptr);
  const int matched = cnt;
  if (tid == 0) {
    ctl->n_matched = cnt;
    for (int i = 0; i < 7; ++i) ctl->cam_old[i] = mu[i];
    s_nhyp = cfg.nhyp0;
    s_numzli = 0;
  }
  int it = 0;
  while (true) {
    __syncthreads();
    if (!(it < s_nhyp && cnt > 0)) break;
    if (tid == 0) {
      const uint32_t rv = n_picks > 0 ? picks[it % n_picks] : 0u;
      s_p = (int)(rv % (uint32_t)cnt);
      s_sel = cand[s_p];
    }
    __syncthreads();
    const int p = s_p, sel = s_sel;
    {  // erase cand[p] (V:991): chunks of 8 * nthr entries, low to high, so any list length is handled
      for (int base = p; base < cnt - 1; base += 8 * nthr) {
        int tmp[8];
        int c = 0;
        for (int idx = base + tid; idx < cnt - 1 && c < 8; idx += nthr) tmp[c++] = cand[idx + 1];
        __syncthreads();
        c = 0;
        for (int idx = base + tid; idx < cnt - 1 && c < 8; idx += nthr) cand[idx] = tmp[c++];
        __syncthreads();
      }
      cnt -= 1;
    }
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
    {  // mu_i = mu + (Sigma H^T) S^-1 (z - h)   (V:995-996)
      const int pos = s_pos, nd = s_nd;
      for (int i = tid; i < n; i += nthr) {
        const double* row = Sigma + (size_t)i * ld;
        double sg[13];
#pragma unroll
        for (int c = 0; c < 13; ++c) sg[c] = (c < nd) ? row[ekf_idx13(c, pos)] : 0.0;  // independent loads in flight
        double w0 = 0, w1 = 0;
#pragma unroll
        for (int c = 0; c < 13; ++c)
          if (c < nd) { w0 += sg[c] * Hs[c]; w1 += sg[c] * Hs[13 + c]; }
        const double k0 = w0 * Sinv[0] + w1 * Sinv[2];
        const double k1 = w0 * Sinv[1] + w1 * Sinv[3];
        mu_i[i] = mu[i] + (k0 * inn[0] + k1 * inn[1]);
      }
    }
    __syncthreads();
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
    int actual = 0;
    for (int start = 0; start < N; start += nthr) {
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
      actual += __syncthreads_count(flag);
    }
    if (tid == 0 && actual > s_numzli) {
      s_numzli = actual;
      s_nhyp = (int)(log(1 - cfg.ransac_p) / (log(1 - (actual / (matched + 0.0)))));  // V:1030
    }
    ++it;
  }
  __syncthreads();
  const int nli = block_compact(ft.li, N, ft.sel, ft.pos_in_z);  // V:1040-1048
  if (tid == 0) {
    ctl->ransac_hyps = it;
    ctl->n_li = nli;
    ctl->k_rows = 2 * nli;
  }
}
