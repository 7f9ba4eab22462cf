// csrc/ekf_match_tile.cuh — the register-tiled scoring core of the warp-per-feature NCC matcher for FULL search windows
// (Patch::findMatch, Patch.cpp:236-262; computeCorrelation, Patch.cpp:293-329).  SURVEY.md §8(a) rows a11-a12.
//
// The CTA-per-feature matcher (match_one in ekf_match.cu) spends 22 k warp-instructions per 41 x 41-candidate feature, 8 % of them
// DP4A: separable box sums through shared memory, a double-precision rsqrt per candidate, seven block barriers.  Here ONE lane
// scores a TILE of 4 x 4 candidates and nothing but the u8 window sits in shared memory:
//   * the lane walks the R + W - 1 window rows of its tile once; per row four aligned words, nine funnel shifts, and the twelve
//     shifted words feed (a) Stp of every candidate row that overlaps this window row (template words in registers), (b) running
//     sums of p and p^2 whose differences between the first and last row of a candidate are its P and PP — all by DP4A, all
//     exact integers;
//   * the score that ranks the candidates is FLOAT: f = (n Stp - T P) * rsqrt(d1) * rsqrt(n PP - P^2), numerator and d2 exact in
//     int32 (n <= 144), relative error of f below 6e-7.  f only PRE-SELECTS: the candidates within 4e-6 of the largest f are the
//     only ones that can lie in the guard band of the double-precision ncc* (two float ulps + 4e-12 below its maximum, see
//     ekf_match.cu), so ncc* in double and then the reference's exact operation sequence are evaluated for those few (normally
//     one) and the result bits are the reference's.  Each lane keeps its two best candidates with their integer sums and the
//     value of its third; a third inside the band (or more than MT_LIST band candidates in the warp) hands the feature to the
//     CTA matcher.
// Everything here is __host__ __device__: tests/match_tile_emu.cu runs the same code lane by lane on the CPU against the oracle.
#pragma once
#include <cmath>
#include <cstdint>

#include "ekf_math.cuh"

#if defined(__CUDACC__)
#define MT_HD __host__ __device__ __forceinline__
#else
#define MT_HD inline
#endif

#define MT_R 4          // most candidate rows per tile (always 4 candidate columns: one aligned word of shifts); sizes the window buffer
#define MT_WSW 19       // window row stride in words: 4 rows apart = 76 words = 12 banks, so the three tile rows a warp works on at once do not collide
#define MT_NW 16        // words staged per window row (41 + 15 - 1 <= 56 bytes + one word of slack)
#define MT_MAXGRID 41   // candidate grid side at the reference's clamp of +-20 px
#define MT_LIST 16      // band candidates per feature
#define MT_BAND 4.0e-6f // pre-selection band below the largest float score

MT_HD unsigned mt_dp4a(unsigned a, unsigned b, unsigned c) {
#ifdef __CUDA_ARCH__
  return __dp4a(a, b, c);
#else
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
  return c;
#endif
}
MT_HD unsigned mt_fshr(unsigned lo, unsigned hi, int sh) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
MT_HD float mt_fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;   // one rounding, never contracted
  return r;
#endif
}
MT_HD float mt_fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
MT_HD float mt_rsqrtf(float x) {   // x = (float) of a positive int32: one MUFU.RSQ (2 ulp), no range fix-up needed
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}

struct CUtensorMap_st;
struct MatchJob {
  const uint8_t* frame;  // frame base
  int fw, fh, fstride;
  const uint8_t* tmpl;   // w*w template
  double hu, hv;         // Patch::h
  double S[4];           // 2x2 block of St
  const CUtensorMap_st* tmap;  // tensor map of the frame stack (null: stage with ordinary loads)
  int frame_index;       // z coordinate in that stack
};

// Scalar setup of Patch::findMatch (Patch.cpp:218-246), replicated per thread: the candidate range of the reference's two loops,
// the ellipse coefficients, and that range clipped to pixels that pass the in-image test.
struct MatchGeom {
  int uc, vc, i0, j0, nv, ilo, jlo, cw, ch;
  float x_2_coeff, y_2_coeff, yx_coeff, sigma_2;
  bool any;
};
MT_HD MatchGeom match_geometry(const MatchJob& jb, int w, float sigma_size, float clampv, int max_grid) {
  MatchGeom G;
  const int half = w / 2;
  G.uc = (int)jb.hu;
  G.vc = (int)jb.hv;
  double invS[4];
  d_inv2_pplu(jb.S, invS);
  G.x_2_coeff = (float)invS[0];
  G.y_2_coeff = (float)invS[3];
  G.yx_coeff = (float)(2 * invS[2]);
  G.sigma_2 = sigma_size * sigma_size;
  float delta_u = (float)(sigma_size * sqrt(jb.S[0]));
  float delta_v = (float)(sigma_size * sqrt(jb.S[3]));
  if (delta_u > clampv) delta_u = clampv;
  if (delta_v > clampv) delta_v = clampv;
  // for (int i = uc - delta_u; i <= uc + delta_u; i++): float arithmetic, truncation toward zero
  G.i0 = (int)((float)G.uc - delta_u);
  G.j0 = (int)((float)G.vc - delta_v);
  const float iu_hi = (float)G.uc + delta_u, jv_hi = (float)G.vc + delta_v;
  const int i1 = (int)floorf(iu_hi), j1 = (int)floorf(jv_hi);
  // NaN covariance: the loops do not run in the reference (comparisons are false)
  const bool finite_ok = (iu_hi == iu_hi) && (jv_hi == jv_hi) && (delta_u == delta_u) && (delta_v == delta_v);
  G.nv = finite_ok ? (j1 - G.j0 + 1) : 0;
  if (G.nv < 0) G.nv = 0;
  // clip the candidate range to pixels that pass the in-image test (Patch.cpp:246) so the staged
  // window never leaves the frame; scan order and keys are unaffected.
  G.ilo = max(G.i0, half + 1);
  G.jlo = max(G.j0, half + 1);
  const int ihi = min(i1, jb.fw - half - 1), jhi = min(j1, jb.fh - half - 1);
  G.cw = finite_ok ? ihi - G.ilo + 1 : 0;
  G.ch = finite_ok ? jhi - G.jlo + 1 : 0;  // valid candidate grid
  if (G.cw > max_grid) G.cw = max_grid;  // cannot happen for delta <= clamp; keeps smem in bounds
  if (G.ch > max_grid) G.ch = max_grid;
  G.any = G.cw > 0 && G.ch > 0;
  return G;
}


// Window staging of the tile matchers: MT_NW words per row (row stride MT_WSW), columns past ww zero, and MT_R - 1 zero rows
// below (the last tile row reads them).  Thread t of nthreads (a multiple of 16) owns word column t % 16 — everything that
// depends on the column is hoisted — and walks the rows t / 16, + nthreads / 16, ...
MT_HD void mt_stage_window(unsigned* winw, const MatchJob& jb, const MatchGeom& G, int W, int t, int nthreads) {
  const int half = W / 2, ww = G.cw + W - 1, wh = G.ch + W - 1;
  const int x0 = G.ilo - half, y0 = G.jlo - half;
  const int nrows = wh + MT_R - 1;
  if ((((size_t)jb.frame | (size_t)jb.fstride) & 3) == 0) {
    const int sh = x0 & 3, xa = x0 - sh;
    const int k = t & (MT_NW - 1), dr = nthreads / MT_NW;
    const int rem = ww - 4 * k;
    const bool colin = rem > 0;
    const unsigned cmask = rem >= 4 ? 0xffffffffu : (rem > 0 ? (1u << (8 * rem)) - 1u : 0u);
    const int last = (4 * k + 3 < ww - 1) ? 4 * k + 3 : ww - 1;                // last window column this word needs
    const bool need_hi = sh > 0 && last + sh >= 4 * k + 4;    // the word past the last needed byte is never touched
    const uint8_t* src0 = jb.frame + (size_t)y0 * jb.fstride + xa + 4 * k;
#pragma unroll 4
    for (int yy = t / MT_NW; yy < nrows; yy += dr) {
      unsigned v = 0;
      if (yy < wh && colin) {
        const unsigned* p = reinterpret_cast<const unsigned*>(src0 + (size_t)yy * jb.fstride);
        const unsigned lo = p[0];
        const unsigned hi = need_hi ? p[1] : 0u;
        v = mt_fshr(lo, hi, 8 * sh) & cmask;
      }
      winw[yy * MT_WSW + k] = v;
    }
  } else {
    uint8_t* winb = reinterpret_cast<uint8_t*>(winw);
    for (int e = t; e < nrows * MT_NW * 4; e += nthreads) {
      const int yy = e / (MT_NW * 4), xx = e - yy * (MT_NW * 4);
      winb[yy * MT_WSW * 4 + xx] = (yy < wh && xx < ww) ? jb.frame[(size_t)(y0 + yy) * jb.fstride + x0 + xx] : (uint8_t)0;
    }
  }
}

// what a lane needs of the feature's geometry to gate and rank a candidate
struct MTGate {
  float x2c, y2c, yxc, sigma2;   // ellipse (Patch.cpp:241-247)
  int du0, dv0;                  // ilo - uc, jlo - vc: candidate (iu, jv) sits at (du0 + iu, dv0 + jv) from the predicted centre
  int cw, ch;                    // candidate grid
  int T;                         // sum of the template bytes
  float rd1f;                    // (float) rsqrt(n TT - T^2)
};

// a lane's two best candidates (float score, index, integer sums) and the score of its third
struct MTTop {
  float a0, a1, a2;
  int i0, i1;
  unsigned s0, s1;   // Stp
  int p0, p1;        // P
  int q0, q1;        // PP
  MT_HD void reset() {
    a0 = a1 = a2 = -INFINITY;
    i0 = i1 = 0; s0 = s1 = 0; p0 = p1 = 0; q0 = q1 = 0;
  }
  MT_HD void insert(float f, int idx, unsigned s, int P, int PP) {
    if (f > a2) {
      if (f > a1) {
        a2 = a1;
        if (f > a0) {
          a1 = a0; i1 = i0; s1 = s0; p1 = p0; q1 = q0;
          a0 = f; i0 = idx; s0 = s; p0 = P; q0 = PP;
        } else {
          a1 = f; i1 = idx; s1 = s; p1 = P; q1 = PP;
        }
      } else {
        a2 = f;
      }
    }
  }
};

// One tile of 4 x R candidates: (4 tx + s, R ty + j), s = 0..3, j = 0..R-1.  winw: the staged window as words (row stride MT_WSW); T: the packed
// template (zero padded to whole words), row-major [W][TW].
//
// SLIDE: the window sums of the three right-hand candidates of a row come from the left-hand one by sliding — one byte leaves, one
// enters: d = in - out, P += d, PP += d (in + out) — on the integer ALU / IMAD pipes instead of 18 of the 24 DP4As per window row
// (DP4A is the instruction this kernel is bound by).  What runs down the rows is then cp[0] / cpp[0] and the three running
// DIFFERENCES between neighbouring candidates; a prefix sum rebuilds cp[s] where a candidate row starts or ends.
template <int W, int R, bool SLIDE = false>
MT_HD void mt_tile(const unsigned* winw, int tx, int ty, const unsigned (&T)[W * ((W + 3) / 4)], const MTGate& g, MTTop& top) {
  constexpr int TW = (W + 3) / 4;
  constexpr int ROWS = R + W - 1;
  constexpr int NN = W * W;
  constexpr unsigned LASTMASK = (W & 3) ? ((1u << (8 * (W & 3))) - 1u) : 0xffffffffu;
  static_assert(W <= 12, "int32 numerators need W <= 12; TW + 1 <= 4 words per row");
  unsigned stp[R][4], ps[R][4], pps[R][4], cp[4], cpp[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    cp[s] = 0; cpp[s] = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) { stp[j][s] = 0; ps[j][s] = 0; pps[j][s] = 0; }
  }
  const unsigned* base = winw + (R * ty) * MT_WSW + tx;
  // per column of the tile: inside the grid?, and the two ellipse terms that do not depend on the row (Patch.cpp:247, same
  // association: (x2c di) di, (yxc di) dj)
  bool colok[4];
  float ex[4], cxy[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int iu = 4 * tx + s;
    colok[s] = iu < g.cw;
    const float fdi = (float)(g.du0 + iu);
    ex[s] = mt_fmul(mt_fmul(g.x2c, fdi), fdi);
    cxy[s] = mt_fmul(g.yxc, fdi);
  }
#pragma unroll
  for (int y = 0; y < ROWS; ++y) {
    unsigned wv[TW + 1];
#pragma unroll
    for (int k = 0; k <= TW; ++k) wv[k] = base[y * MT_WSW + k];
    unsigned x[4][TW];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
      for (int k = 0; k < TW; ++k) x[s][k] = s ? mt_fshr(wv[k], wv[k + 1], 8 * s) : wv[k];
      x[s][TW - 1] &= LASTMASK;   // bytes past the template side of this candidate
    }
    if (y < R) {                  // candidate row y starts here: remember the running sums before this row
      if (SLIDE) {
        unsigned a = cp[0], b = cpp[0];
        ps[y][0] = a; pps[y][0] = b;
#pragma unroll
        for (int s = 1; s < 4; ++s) { a += cp[s]; b += cpp[s]; ps[y][s] = a; pps[y][s] = b; }
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s) { ps[y][s] = cp[s]; pps[y][s] = cpp[s]; }
      }
    }
    if (SLIDE) {
#pragma unroll
      for (int k = 0; k < TW; ++k) {
        cp[0] = mt_dp4a(x[0][k], 0x01010101u, cp[0]);
        cpp[0] = mt_dp4a(x[0][k], x[0][k], cpp[0]);
      }
#pragma unroll
      for (int s = 1; s < 4; ++s) {   // candidate s = candidate s - 1 without window column s - 1, with column s - 1 + W
        const unsigned bout = (wv[(s - 1) >> 2] >> (8 * ((s - 1) & 3))) & 255u;
        const unsigned bin = (wv[(s - 1 + W) >> 2] >> (8 * ((s - 1 + W) & 3))) & 255u;
        const unsigned d = bin - bout;
        cp[s] += d;
        cpp[s] += d * (bin + bout);
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int k = 0; k < TW; ++k) {
          cp[s] = mt_dp4a(x[s][k], 0x01010101u, cp[s]);
          cpp[s] = mt_dp4a(x[s][k], x[s][k], cpp[s]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = y - j;        // template row that window row y meets in candidate row j
      if (r >= 0 && r < W) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
          for (int k = 0; k < TW; ++k) stp[j][s] = mt_dp4a(x[s][k], T[r * TW + k], stp[j][s]);
        }
      }
    }
    if (y >= W - 1) {             // candidate row j = y - (W - 1) is complete: straight-line scoring, one rare branch
      const int j = y - (W - 1);
      const int jv = R * ty + j;
      const bool rowok = jv < g.ch;
      const float dj = (float)(g.dv0 + jv);
      const float ey = mt_fmul(mt_fmul(g.y2c, dj), dj);
      unsigned cps[4], cpps[4];   // the running sums of the four candidates after this window row
      if (SLIDE) {
        cps[0] = cp[0]; cpps[0] = cpp[0];
#pragma unroll
        for (int s = 1; s < 4; ++s) { cps[s] = cps[s - 1] + cp[s]; cpps[s] = cpps[s - 1] + cpp[s]; }
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s) { cps[s] = cp[s]; cpps[s] = cpp[s]; }
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float e = mt_fadd(mt_fadd(ex[s], ey), mt_fmul(cxy[s], dj));   // ellipse gate in float
        const int P = (int)(cps[s] - ps[j][s]), PP = (int)(cpps[s] - pps[j][s]);
        const int d2 = NN * PP - P * P;                   // exact: <= 144 * 144 * 255^2 < 2^31; 0 for a flat window
        const int num = NN * (int)stp[j][s] - g.T * P;    // exact
        const float f = (float)num * (g.rd1f * mt_rsqrtf((float)d2));
        const bool valid = rowok && colok[s] && (e <= g.sigma2) && (d2 > 0);   // flat window: 0/0 in the reference, never selected
#ifdef __CUDA_ARCH__
        if (__builtin_expect(valid && f > top.a2, 0))
#else
        if (valid && f > top.a2)
#endif
          top.insert(f, jv * g.cw + 4 * tx + s, stp[j][s], P, PP);
      }
    }
  }
}

// ncc* in double from the integer sums (the expression of match_one's fast pass)
MT_HD double mt_ncc_star(int w2, unsigned stp, int P, int PP, int T, double rd1) {
  const double dn = (double)w2;
  const double d2 = dn * (double)PP - (double)P * (double)P;
  const double num = dn * (double)stp - (double)T * (double)P;
#ifdef __CUDA_ARCH__
  return (num * rd1) * rsqrt(d2);
#else
  return (num * rd1) * (1.0 / sqrt(d2));
#endif
}
