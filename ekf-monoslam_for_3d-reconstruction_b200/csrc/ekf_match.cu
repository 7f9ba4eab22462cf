// csrc/ekf_match.cu — K3: active-search NCC matcher (Patch::findMatch, Patch.cpp:215-291, and
// computeCorrelation, Patch.cpp:293-329).  SURVEY.md §8(a) rows a11-a13.
//
// One CTA per feature.  Staged in shared memory: the template as d1 = (float)pixel - mean (doubles),
// the search window ((2*delta_u + w) x (2*delta_v + w), at most 72 x 72 at the reference's +-20 px
// clamp) converted ONCE to doubles, and its summed-area table (exact integer sums give every
// candidate's mean without a pass over its pixels).  Each thread owns two vertically adjacent
// candidates and walks their w x w pixels in the reference's row-major order with
// __dadd_rn/__dmul_rn (never contracted), so every candidate's double-precision score is produced
// by the same IEEE operation sequence as the reference; template statistics are hoisted (identical
// for every candidate).  The score is rounded to float and compared in float exactly as
// Patch.cpp:243,252,278 do.  Warp shuffles carry the arg-max with the reference's tie-break: the
// first candidate in scan order (u outer, v inner) wins.  fp64-ALU bound: 5 DP ops per pixel per
// candidate, one 8-byte shared load per 10 DP ops.
#include "ekf_kernels.h"
#include "ekf_math.cuh"

#define MATCH_THREADS 256
#define MATCH_MAX_W 31  // largest template side

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

// shared-memory plan for template side w and search clamp cl (pixels)
struct MatchSmem {
  int side;      // max window side = 2 cl + w
  int wsd;       // window row stride in doubles
  size_t off_win, off_sat, off_red, total;
};
__host__ __device__ inline MatchSmem match_smem_plan(int w, int cl) {
  MatchSmem p;
  p.side = 2 * cl + w;
  p.wsd = p.side + 1;
  size_t o = (size_t)((w * w + 1) & ~1) * sizeof(double);               // d1
  p.off_win = o; o += (size_t)(p.side + 1) * p.wsd * sizeof(double);     // window (one spare row)
  p.off_sat = o; o += (size_t)(p.side + 1) * (p.side + 1) * sizeof(int); // summed-area table
  o = (o + 7) & ~(size_t)7;
  p.off_red = o; o += 2 * sizeof(double) + 8 * sizeof(float) + 8 * sizeof(int);
  p.total = o;
  return p;
}

// Core search for one feature by one CTA.  All threads must call.
__device__ MatchResult match_one(const MatchJob& jb, int w, float sigma_size, float clampv, unsigned char* smem_raw) {
  const int tid = threadIdx.x;
  const int half = w / 2, w2 = w * w;
  const MatchSmem pl = match_smem_plan(w, (int)clampv);
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
  const int i1 = (int)floorf(iu_hi), j1 = (int)floorf(jv_hi);
  // NaN covariance: the loops do not run in the reference (comparisons are false)
  const bool finite_ok = (iu_hi == iu_hi) && (jv_hi == jv_hi) && (delta_u == delta_u) && (delta_v == delta_v);
  int nv = finite_ok ? (j1 - j0 + 1) : 0;
  if (nv < 0) nv = 0;
  // clip the candidate range to pixels that pass the in-image test (Patch.cpp:246) so the staged
  // window never leaves the frame; scan order and keys are unaffected.
  const int ilo = max(i0, half + 1), ihi = min(i1, jb.fw - half - 1);
  const int jlo = max(j0, half + 1), jhi = min(j1, jb.fh - half - 1);
  int cw = finite_ok ? ihi - ilo + 1 : 0, ch = finite_ok ? jhi - jlo + 1 : 0;  // valid candidate grid
  if (cw > pl.side - w + 1) cw = pl.side - w + 1;  // cannot happen for delta <= clamp; keeps smem in bounds
  if (ch > pl.side - w + 1) ch = pl.side - w + 1;
  const bool any = cw > 0 && ch > 0;
  const int ww = any ? cw + w - 1 : 0, wh = any ? ch + w - 1 : 0;
  const int wsd = pl.wsd, sst = pl.side + 1;

  double* d1 = reinterpret_cast<double*>(smem_raw);
  double* wind = reinterpret_cast<double*>(smem_raw + pl.off_win);
  int* sat = reinterpret_cast<int*>(smem_raw + pl.off_sat);
  double* red_n1 = reinterpret_cast<double*>(smem_raw + pl.off_red);
  float* red_s = reinterpret_cast<float*>(red_n1 + 2);
  int* red_k = reinterpret_cast<int*>(red_s + 8);
  int* isum = reinterpret_cast<int*>(red_n1 + 1);

  // --- template mean (integer sum: exact in any order) ---
  if (tid == 0) *isum = 0;
  __syncthreads();
  {
    int part = 0;
    for (int e = tid; e < w2; e += MATCH_THREADS) part += jb.tmpl[e];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0 && part) atomicAdd(isum, part);
  }
  // --- stage the window as doubles (u8 -> float -> double is exact) ---
  if (any) {
    const int x0 = ilo - half, y0 = jlo - half;
    for (int yy = tid / 64; yy < wh; yy += MATCH_THREADS / 64)
      for (int xx = tid & 63; xx < wsd; xx += 64)
        wind[yy * wsd + xx] = (xx < ww) ? (double)(float)jb.frame[(size_t)(y0 + yy) * jb.fstride + x0 + xx] : 0.0;
    for (int xx = tid; xx < wsd; xx += MATCH_THREADS) wind[wh * wsd + xx] = 0.0;  // spare row
  }
  __syncthreads();
  const double m1 = __ddiv_rn((double)(*isum), (double)w2);
  for (int e = tid; e < w2; e += MATCH_THREADS) d1[e] = __dsub_rn((double)(float)jb.tmpl[e], m1);
  // --- summed-area table: row prefix sums, then column prefix sums (ints, exact) ---
  if (any) {
    for (int yy = tid; yy < wh; yy += MATCH_THREADS) {
      int run = 0;
      sat[(yy + 1) * sst] = 0;
      for (int xx = 0; xx < ww; ++xx) { run += (int)wind[yy * wsd + xx]; sat[(yy + 1) * sst + xx + 1] = run; }
    }
    for (int xx = tid; xx <= ww; xx += MATCH_THREADS) sat[xx] = 0;
  }
  __syncthreads();
  if (tid == MATCH_THREADS - 1) {
    // n1 = sum (s1 - m1)^2 sequentially in the reference order; overlaps the column pass below
    double n1 = 0;
    for (int e = 0; e < w2; ++e) n1 = __dadd_rn(n1, __dmul_rn(d1[e], d1[e]));
    *red_n1 = n1;
  } else if (any) {
    for (int xx = tid; xx <= ww; xx += MATCH_THREADS - 1) {
      int run = 0;
      for (int yy = 1; yy <= wh; ++yy) { run += sat[yy * sst + xx]; sat[yy * sst + xx] = run; }
    }
  }
  __syncthreads();
  const double n1 = *red_n1;

  float best = -1.0f;
  int bestkey = 0x7fffffff;
  if (any) {
    const int chp = (ch + 1) >> 1;
    const int npair = cw * chp;
    const double dw2 = (double)w2;
    for (int c = tid; c < npair; c += MATCH_THREADS) {
      // consecutive lanes take consecutive u (conflict-free 8-byte window reads); the two candidates
      // of a thread are vertical neighbours and share every window row they both touch
      const int jp = c / cw, iu = c - jp * cw;
      const int jv = 2 * jp;
      const int i = ilo + iu, ja = jlo + jv;
      const int di = i - uc;
      const bool hasB = (jv + 1 < ch);
      bool va, vb;
      {
        const float fdi = (float)di;
        const float ex = __fmul_rn(__fmul_rn(x_2_coeff, fdi), fdi);
        const float dja = (float)(ja - vc), djb = (float)(ja + 1 - vc);
        // ellipse gate in float, same association as Patch.cpp:247
        const float ea = __fadd_rn(__fadd_rn(ex, __fmul_rn(__fmul_rn(y_2_coeff, dja), dja)), __fmul_rn(__fmul_rn(yx_coeff, fdi), dja));
        const float eb = __fadd_rn(__fadd_rn(ex, __fmul_rn(__fmul_rn(y_2_coeff, djb), djb)), __fmul_rn(__fmul_rn(yx_coeff, fdi), djb));
        va = (ea <= sigma_2);
        vb = hasB && (eb <= sigma_2);
      }
      if (!va && !vb) continue;
      const int* sa = sat + jv * sst + iu;
      const int suma = sa[w * sst + w] - sa[w] - sa[w * sst] + sa[0];
      const int sumb = hasB ? sa[(w + 1) * sst + w] - sa[sst + w] - sa[(w + 1) * sst] + sa[sst] : 0;
      const double m2a = __ddiv_rn((double)suma, dw2);
      const double m2b = __ddiv_rn((double)sumb, dw2);
      double n2a = 0, ca = 0, n2b = 0, cb = 0;
      const double* wr = wind + jv * wsd + iu;
      {  // window row 0: candidate A only
        const double* t = d1;
#pragma unroll 4
        for (int x = 0; x < w; ++x) {
          const double da = __dsub_rn(wr[x], m2a);
          n2a = __dadd_rn(n2a, __dmul_rn(da, da));
          ca = __dadd_rn(ca, __dmul_rn(t[x], da));
        }
      }
      for (int r = 1; r < w; ++r) {  // rows shared by A (template row r) and B (template row r-1)
        const double* row = wr + r * wsd;
        const double* ta = d1 + r * w;
        const double* tb = ta - w;
#pragma unroll 4
        for (int x = 0; x < w; ++x) {
          const double pv = row[x];
          const double da = __dsub_rn(pv, m2a);
          n2a = __dadd_rn(n2a, __dmul_rn(da, da));
          ca = __dadd_rn(ca, __dmul_rn(ta[x], da));
          const double db = __dsub_rn(pv, m2b);
          n2b = __dadd_rn(n2b, __dmul_rn(db, db));
          cb = __dadd_rn(cb, __dmul_rn(tb[x], db));
        }
      }
      {  // window row w: candidate B only (the spare row keeps the read in bounds when B is absent)
        const double* row = wr + w * wsd;
        const double* tb = d1 + (w - 1) * w;
#pragma unroll 4
        for (int x = 0; x < w; ++x) {
          const double db = __dsub_rn(row[x], m2b);
          n2b = __dadd_rn(n2b, __dmul_rn(db, db));
          cb = __dadd_rn(cb, __dmul_rn(tb[x], db));
        }
      }
      if (va) {
        const float sc = (float)__ddiv_rn(ca, __dsqrt_rn(__dmul_rn(n2a, n1)));
        const int key = (i - i0) * nv + (ja - j0);  // position in the reference's scan order
        if (sc > best || (sc == best && key < bestkey)) { best = sc; bestkey = key; }
      }
      if (vb) {
        const float sc = (float)__ddiv_rn(cb, __dsqrt_rn(__dmul_rn(n2b, n1)));
        const int key = (i - i0) * nv + (ja + 1 - j0);
        if (sc > best || (sc == best && key < bestkey)) { best = sc; bestkey = key; }
      }
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
  MatchResult res;
  res.best = best; res.bi = 0; res.bj = 0;
  if (bestkey != 0x7fffffff && nv > 0) {
    res.bi = i0 + bestkey / nv;
    res.bj = j0 + bestkey % nv;
  }
  return res;
}

static size_t match_smem_bytes(int w, float clampv) { return match_smem_plan(w, (int)clampv).total; }

// Filter-attached matcher: the loop V:870-880 with one CTA per feature.
__device__ __forceinline__ void match_filter_feature(FeatTab ft, int f, FrameView fr, const DevCfg& cfg, unsigned char* smem_raw) {
  const int w = cfg.window, w2 = w * w;
  MatchJob jb;
  jb.frame = fr.px; jb.fw = fr.w; jb.fh = fr.h; jb.fstride = fr.stride;
  jb.tmpl = ft.mpatch + (size_t)f * cfg.tstride;
  jb.hu = ft.h[2 * f]; jb.hv = ft.h[2 * f + 1];
  for (int c = 0; c < 4; ++c) jb.S[c] = ft.S2[4 * f + c];
  const MatchResult r = match_one(jb, w, cfg.sigma_size_f, cfg.search_clamp, smem_raw);
  __syncthreads();
  const bool accept = !(r.best < cfg.ncc_threshold);  // Patch.cpp:278
  if (accept) {
    // matching_patch <- matched ROI (Patch.cpp:285)
    const int x0 = r.bi - w / 2, y0 = r.bj - w / 2;
    for (int e = threadIdx.x; e < w2; e += MATCH_THREADS)
      ft.mpatch[(size_t)f * cfg.tstride + e] = fr.px[(size_t)(y0 + e / w) * fr.stride + x0 + (e % w)];
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
__global__ void __launch_bounds__(MATCH_THREADS) k_match_filter(FeatTab ft, int N, FrameView fr, DevCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = blockIdx.x;
  if (f >= N || !ft.innov[f]) return;
  match_filter_feature(ft, f, fr, cfg, smem_raw);
}
// Batched filters (BASELINE config 3): grid = (feature capacity, filters); the frame is shared.
__global__ void __launch_bounds__(MATCH_THREADS) k_match_filter_batch(FeatTab base, int Ncap, const int* __restrict__ Nper,
                                                                      FrameView fr, DevCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = blockIdx.x, b = blockIdx.y;
  const FeatTab ft = feattab_slice(base, b, Ncap, cfg.tstride);
  if (f >= Nper[b] || !ft.innov[f]) return;
  match_filter_feature(ft, f, fr, cfg, smem_raw);
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
  static size_t attr_smem = 0;
  const size_t smem = match_smem_bytes(cfg.window, cfg.search_clamp);
  if (smem > attr_smem) {
    cudaFuncSetAttribute(k_match_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  k_match_filter<<<N, MATCH_THREADS, smem, st>>>(ft, N, fr, cfg);
  *launches += 1;
}

void launch_match_filter_batch(cudaStream_t st, FeatTab base, int Ncap, const int* Nper, int B, FrameView fr, const DevCfg& cfg,
                               long long* launches) {
  if (B <= 0 || Ncap <= 0) return;
  static size_t attr_smem = 0;
  const size_t smem = match_smem_bytes(cfg.window, cfg.search_clamp);
  if (smem > attr_smem) {
    cudaFuncSetAttribute(k_match_filter_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  k_match_filter_batch<<<dim3(Ncap, B), MATCH_THREADS, smem, st>>>(base, Ncap, Nper, fr, cfg);
  *launches += 1;
}

int launch_match_batch(cudaStream_t st, const uint8_t* frames, int n_frames, int width, int height, int stride,
                       const uint8_t* templates, int fpf, int w, const double* h, const double* S, float sigma_size,
                       float thr, float clampv, int32_t* out_uv, float* out_score) {
  if (w > MATCH_MAX_W || w < 1 || clampv > 20.0f || !(clampv >= 0.0f)) return -1;
  const int total = n_frames * fpf;
  if (total <= 0) return 0;
  static size_t attr_smem = 0;
  const size_t smem = match_smem_bytes(w, clampv);
  if (smem > attr_smem) {
    cudaFuncSetAttribute(k_match_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  k_match_batch<<<total, MATCH_THREADS, smem, st>>>(frames, width, height, stride, templates, fpf, w, h, S, sigma_size, thr,
                                                   clampv, out_uv, out_score, total);
  return 0;
}
