// csrc/ekf_match.cu — K3: active-search NCC matcher (Patch::findMatch, Patch.cpp:215-291, and
// computeCorrelation, Patch.cpp:293-329).  SURVEY.md §8(a) rows a11-a13.
//
// One CTA per feature.  The search window ((2*delta_u + w) x (2*delta_v + w) u8, at most 64 x 64 for
// w <= 23 at the reference's +-20 px clamp) and the template are staged in shared memory; each thread
// owns whole candidates and walks the w x w pixels in the reference's row-major order with
// __dadd_rn/__dmul_rn (never contracted), so every candidate's double-precision score is computed by
// the same IEEE operation sequence as the reference: template statistics are hoisted (identical
// values for every candidate), the pixel sums of pass 1 are exact integers.  The result is rounded to
// float and compared in float exactly as Patch.cpp:243,252,278 do.  Warp shuffles carry the arg-max
// reduction with the reference's tie-break: first candidate in scan order (u outer, v inner) wins.
// The kernel is fp64-ALU bound (5 DP ops per pixel per candidate), not HBM bound.
#include "ekf_kernels.h"
#include "ekf_math.cuh"

#define MATCH_THREADS 256
#define MATCH_MAX_W 31          // largest template side supported by the smem carve-up
#define MATCH_WIN_MAX (2 * 20 + 1 + MATCH_MAX_W)  // window side bound at clamp 20

struct MatchJob {
  const uint8_t* frame;  // frame base
  int fw, fh, fstride;
  const uint8_t* tmpl;   // w*w template
  double hu, hv;         // Patch::h
  double S[4];           // 2x2 block of St
};

struct MatchResult {
  float best;  // max NCC (-1 if no candidate)
  int bi, bj;  // argmax (valid if best > -1)
};

// Core search for one feature by one CTA.  All threads must call.  smem: dynamic buffer.
__device__ MatchResult match_one(const MatchJob& jb, int w, float sigma_size, float clampv, unsigned char* smem_raw) {
  const int tid = threadIdx.x;
  const int half = w / 2, w2 = w * w;
  // --- scalar setup, replicated per thread (Patch.cpp:218-241) ---
  const int uc = (int)jb.hu;
  const int vc = (int)jb.hv;
  double invS[4];
  d_inv2_pplu(jb.S, invS);
  const float x_2_coeff = (float)invS[0];
  const float y_2_coeff = (float)invS[3];
  const float yx_coeff = (float)(2 * invS[2]);
  const float sigma_2 = sigma_size * sigma_size;
  float delta_u = (float)(sigma_size * sqrt(jb.S[0]));
  float delta_v = (float)(sigma_size * sqrt(jb.S[3]));
  if (delta_u > clampv) delta_u = clampv;
  if (delta_v > clampv) delta_v = clampv;
  // for (int i = uc - delta_u; i <= uc + delta_u; i++): float arithmetic, truncation toward zero
  const int i0 = (int)((float)uc - delta_u);
  const int j0 = (int)((float)vc - delta_v);
  const float iu_hi = (float)uc + delta_u, jv_hi = (float)vc + delta_v;
  int i1 = (int)floorf(iu_hi), j1 = (int)floorf(jv_hi);
  // NaN covariance: loops do not run in the reference (comparisons false)
  int nu = (iu_hi == iu_hi) ? (i1 - i0 + 1) : 0;
  int nv = (jv_hi == jv_hi) ? (j1 - j0 + 1) : 0;
  if (nu < 0) nu = 0;
  if (nv < 0) nv = 0;
  // clip the candidate range to pixels that pass the in-image test (Patch.cpp:246) so the staged
  // window never leaves the frame; scan order and keys are unaffected.
  const int ilo = max(i0, half + 1), ihi = min(i1, jb.fw - half - 1);
  const int jlo = max(j0, half + 1), jhi = min(j1, jb.fh - half - 1);
  const int cw = ihi - ilo + 1, ch = jhi - jlo + 1;  // valid candidate grid
  MatchResult res;
  res.best = -1.0f; res.bi = 0; res.bj = 0;

  // --- smem carve-up ---
  double* d1 = reinterpret_cast<double*>(smem_raw);          // w2 doubles: (float)s1 - m1
  double* red_n1 = d1 + MATCH_MAX_W * MATCH_MAX_W;            // 1 double
  float* red_s = reinterpret_cast<float*>(red_n1 + 1);        // 8 floats
  int* red_k = reinterpret_cast<int*>(red_s + 8);             // 8 ints
  int* isum = red_k + 8;                                      // 1 int
  unsigned char* win = reinterpret_cast<unsigned char*>(isum + 4);
  const int ww = (cw > 0 ? cw + w - 1 : 0), wh = (ch > 0 ? ch + w - 1 : 0);
  const int wstride = (ww + 3) & ~3;

  // --- template statistics (hoisted out of computeCorrelation; identical for every candidate) ---
  if (tid == 0) *isum = 0;
  __syncthreads();
  {
    int part = 0;
    for (int e = tid; e < w2; e += MATCH_THREADS) part += jb.tmpl[e];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0 && part) atomicAdd(isum, part);  // integer sum: exact in any order
  }
  __syncthreads();
  const double m1 = __ddiv_rn((double)(*isum), (double)w2);
  for (int e = tid; e < w2; e += MATCH_THREADS) d1[e] = __dsub_rn((double)(float)jb.tmpl[e], m1);
  // --- stage the window ---
  if (cw > 0 && ch > 0) {
    const int x0 = ilo - half, y0 = jlo - half;
    for (int e = tid; e < wh * wstride; e += MATCH_THREADS) {
      const int yy = e / wstride, xx = e % wstride;
      win[e] = (xx < ww) ? jb.frame[(size_t)(y0 + yy) * jb.fstride + x0 + xx] : 0;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double n1 = 0;
    for (int e = 0; e < w2; ++e) n1 = __dadd_rn(n1, __dmul_rn(d1[e], d1[e]));  // sequential, reference order
    *red_n1 = n1;
  }
  __syncthreads();
  const double n1 = *red_n1;

  float best = -1.0f;
  int bestkey = 0x7fffffff;
  if (cw > 0 && ch > 0) {
    const int ncand = cw * ch;
    for (int c = tid; c < ncand; c += MATCH_THREADS) {
      // thread -> candidate: consecutive lanes take consecutive u so their window bytes share
      // 32-bit smem words (conflict-free); the scan-order key below is independent of this mapping
      const int jv = c / cw, iu = c % cw;
      const int i = ilo + iu, j = jlo + jv;
      const int di = i - uc, dj = j - vc;
      // ellipse gate in float, same association as Patch.cpp:247
      const float e = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(x_2_coeff, (float)di), (float)di),
                                          __fmul_rn(__fmul_rn(y_2_coeff, (float)dj), (float)dj)),
                                __fmul_rn(__fmul_rn(yx_coeff, (float)di), (float)dj));
      if (!(e <= sigma_2)) continue;
      const unsigned char* wp = win + (size_t)jv * wstride + iu;
      int s2sum = 0;
      for (int y = 0; y < w; ++y)
        for (int x = 0; x < w; ++x) s2sum += wp[y * wstride + x];
      const double m2 = __ddiv_rn((double)s2sum, (double)w2);
      double n2 = 0, corr = 0;
      for (int y = 0; y < w; ++y) {
#pragma unroll 4
        for (int x = 0; x < w; ++x) {
          const double d2 = __dsub_rn((double)(float)wp[y * wstride + x], m2);
          n2 = __dadd_rn(n2, __dmul_rn(d2, d2));
          corr = __dadd_rn(corr, __dmul_rn(d1[y * w + x], d2));
        }
      }
      const float sc = (float)__ddiv_rn(corr, __dsqrt_rn(__dmul_rn(n2, n1)));
      const int key = (i - i0) * nv + (j - j0);  // position in the reference's scan order
      if (sc > best || (sc == best && key < bestkey)) { best = sc; bestkey = key; }
    }
  }
  // --- arg-max: higher score wins, ties go to the earlier key (strict '>' in a sequential scan) ---
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, bestkey, o);
    if (ob > best || (ob == best && ok < bestkey)) { best = ob; bestkey = ok; }
  }
  if ((tid & 31) == 0) { red_s[tid >> 5] = best; red_k[tid >> 5] = bestkey; }
  __syncthreads();
  best = red_s[0]; bestkey = red_k[0];
  for (int wv = 1; wv < MATCH_THREADS / 32; ++wv) {
    const float ob = red_s[wv];
    const int ok = red_k[wv];
    if (ob > best || (ob == best && ok < bestkey)) { best = ob; bestkey = ok; }
  }
  res.best = best;
  if (bestkey != 0x7fffffff && nv > 0) {
    res.bi = i0 + bestkey / nv;
    res.bj = j0 + bestkey % nv;
  }
  (void)nu;
  return res;
}

static size_t match_smem_bytes() {
  return (MATCH_MAX_W * MATCH_MAX_W + 1) * sizeof(double) + 8 * sizeof(float) + 12 * sizeof(int) +
         (size_t)MATCH_WIN_MAX * ((MATCH_WIN_MAX + 3) & ~3);
}

// Filter-attached matcher: the loop V:870-880 with one CTA per feature.
__global__ void __launch_bounds__(MATCH_THREADS) k_match_filter(FeatTab ft, int N, FrameView fr, DevCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = blockIdx.x;
  if (f >= N || !ft.innov[f]) return;
  const int w = cfg.window, w2 = w * w;
  MatchJob jb;
  jb.frame = fr.px; jb.fw = fr.w; jb.fh = fr.h; jb.fstride = fr.stride;
  jb.tmpl = ft.mpatch + (size_t)f * w2;
  jb.hu = ft.h[2 * f]; jb.hv = ft.h[2 * f + 1];
  for (int c = 0; c < 4; ++c) jb.S[c] = ft.S2[4 * f + c];
  const MatchResult r = match_one(jb, w, cfg.sigma_size_f, cfg.search_clamp, smem_raw);
  __syncthreads();
  const bool accept = !(r.best < cfg.ncc_threshold);  // Patch.cpp:278
  if (accept) {
    // matching_patch <- matched ROI (Patch.cpp:285)
    const int x0 = r.bi - w / 2, y0 = r.bj - w / 2;
    for (int e = threadIdx.x; e < w2; e += MATCH_THREADS)
      ft.mpatch[(size_t)f * w2 + e] = fr.px[(size_t)(y0 + e / w) * fr.stride + x0 + (e % w)];
  }
  if (threadIdx.x == 0) {
    ft.n_tot[f] += 1;  // Patch.cpp:218
    ft.last_ncc[f] = r.best;
    if (!accept) {
      ft.center[2 * f] = -1.0f; ft.center[2 * f + 1] = -1.0f;
      ft.innov[f] = 0; ft.li[f] = 0; ft.hi[f] = 0;
    } else {
      ft.center[2 * f] = (float)r.bi; ft.center[2 * f + 1] = (float)r.bj;
      ft.z[2 * f] = (double)(float)r.bi; ft.z[2 * f + 1] = (double)(float)r.bj;
    }
  }
}

// Stateless batch (BASELINE config 5): grid = frames x features.
__global__ void __launch_bounds__(MATCH_THREADS) k_match_batch(const uint8_t* __restrict__ frames, int width, int height,
                                                               int stride, const uint8_t* __restrict__ templates, int fpf,
                                                               int w, const double* __restrict__ hh, const double* __restrict__ Sm,
                                                               float sigma_size, float thr, float clampv,
                                                               int32_t* __restrict__ out_uv, float* __restrict__ out_score, int total) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int idx = blockIdx.x;
  if (idx >= total) return;
  MatchJob jb;
  jb.frame = frames + (size_t)(idx / fpf) * height * stride;
  jb.fw = width; jb.fh = height; jb.fstride = stride;
  jb.tmpl = templates + (size_t)idx * w * w;
  jb.hu = hh[2 * idx]; jb.hv = hh[2 * idx + 1];
  for (int c = 0; c < 4; ++c) jb.S[c] = Sm[4 * idx + c];
  const MatchResult r = match_one(jb, w, sigma_size, clampv, smem_raw);
  if (threadIdx.x == 0) {
    const bool accept = !(r.best < thr);
    out_uv[2 * idx] = accept ? r.bi : -1;
    out_uv[2 * idx + 1] = accept ? r.bj : -1;
    out_score[idx] = r.best;
  }
}

void launch_match_filter(cudaStream_t st, FeatTab ft, int N, FrameView fr, const DevCfg& cfg, long long* launches) {
  if (N <= 0) return;
  static bool attr_done = false;
  const size_t smem = match_smem_bytes();
  if (!attr_done) {
    cudaFuncSetAttribute(k_match_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  k_match_filter<<<N, MATCH_THREADS, smem, st>>>(ft, N, fr, cfg);
  *launches += 1;
}

int launch_match_batch(cudaStream_t st, const uint8_t* frames, int n_frames, int width, int height, int stride,
                       const uint8_t* templates, int fpf, int w, const double* h, const double* S, float sigma_size,
                       float thr, float clampv, int32_t* out_uv, float* out_score) {
  if (w > MATCH_MAX_W || w < 1 || clampv > 20.0f) return -1;
  const int total = n_frames * fpf;
  if (total <= 0) return 0;
  static bool attr_done = false;
  const size_t smem = match_smem_bytes();
  if (!attr_done) {
    cudaFuncSetAttribute(k_match_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  k_match_batch<<<total, MATCH_THREADS, smem, st>>>(frames, width, height, stride, templates, fpf, w, h, S, sigma_size, thr,
                                                   clampv, out_uv, out_score, total);
  return 0;
}
