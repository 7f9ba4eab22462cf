// tests/match_tile_emu.cu — TEST INFRASTRUCTURE.  Runs the scoring core of the full-window warp matcher
// (csrc/ekf_match_tile.cuh: mt_tile, MTTop, mt_ncc_star, match_geometry — the same __host__ __device__ code the kernels
// k_match_filter_batch_warp2 / k_match_batch_warp2 execute) on the CPU, lane by lane, so that tests/test_match_tile_emu.py can
// compare its decisions with the oracle's Patch::findMatch without a GPU.  The flow mirrors match_one_warp2 in ekf_match.cu:
// template sums, window staging, tiles round-robin over 32 lanes, band list, ncc* in double, guard band, exact pass.
// Compiled as HOST code by nvcc (tests build it on the fly); nothing here is shipped.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ekf_match_tile.cuh"

namespace {
constexpr int W = 11, TW = 3, W2 = 121;

// computeCorrelation (Patch.cpp:293-329) for one candidate, in the reference's operation order
float exact_score(const unsigned char* tb, double m1, const uint8_t* winb, int wsb, int roi, int P) {
  const double m2 = (double)P / (double)W2;
  volatile double n1 = 0, n2 = 0, corr = 0;
  for (int r = 0; r < W; ++r)
    for (int x = 0; x < W; ++x) {
      const double a = (double)(float)tb[r * W + x] - m1;
      const double b = (double)(float)winb[roi + r * wsb + x] - m2;
      volatile double aa = a * a, bb = b * b, ab = a * b;
      n1 = n1 + aa; n2 = n2 + bb; corr = corr + ab;
    }
  volatile double nn = n2 * n1;
  return (float)(corr / sqrt(nn));
}

// returns 1 decided, 0 handed to the CTA matcher
template <int R, bool SLIDE>
int emu_one(const MatchJob& jb, float sigma_size, float clampv, float* best_out, int* bi, int* bj, int* nlist_out) {
  *best_out = -1.0f; *bi = 0; *bj = 0; *nlist_out = 0;
  if (!(clampv <= 20.0f)) return 0;
  const MatchGeom G = match_geometry(jb, W, sigma_size, clampv, MT_MAXGRID);
  if (!G.any) return 1;
  const int cw = G.cw, ch = G.ch;
  std::vector<unsigned> win((MT_MAXGRID + W - 1 + MT_R - 1) * MT_WSW, 0xdeadbeefu);   // unstaged words hold junk on the GPU too
  unsigned tpk[W * TW];
  unsigned char tb[W2];
  int T = 0, TT = 0;
  for (int e = 0; e < W * TW; ++e) {
    const int r = e / TW, k4 = (e - r * TW) * 4;
    unsigned word = 0;
    for (int c = 0; c < 4; ++c)
      if (k4 + c < W) {
        const unsigned v = jb.tmpl[r * W + k4 + c];
        word |= v << (8 * c);
        T += (int)v; TT += (int)(v * v);
      }
    tpk[e] = word;
  }
  memcpy(tb, jb.tmpl, W2);
  // the kernels' own staging (word loads, funnel shift to the window's alignment, column masks), thread by thread: 32 threads as
  // in the warp kernels for R = 4, 256 as in the CTA kernel for R = 2
  const int nthreads = R == 2 ? 256 : 32;
  for (int t = 0; t < nthreads; ++t) mt_stage_window(win.data(), jb, G, W, t, nthreads);
  const double dn = (double)W2;
  const double m1 = (double)T / dn;
  const double d1 = dn * (double)TT - (double)T * (double)T;
  if (!(d1 > 0.0)) return 1;
  const double rd1 = 1.0 / sqrt(d1);
  MTTop top[32];
  MTGate g;
  g.x2c = G.x_2_coeff; g.y2c = G.y_2_coeff; g.yxc = G.yx_coeff; g.sigma2 = G.sigma_2;
  g.du0 = G.ilo - G.uc; g.dv0 = G.jlo - G.vc; g.cw = cw; g.ch = ch; g.T = T; g.rd1f = (float)rd1;
  const int ntx = (cw + 3) >> 2, nty = (ch + R - 1) / R;
  for (int lane = 0; lane < 32; ++lane) {
    top[lane].reset();
    int ty = lane / ntx, tx = lane - ty * ntx;
    const int dty = 32 / ntx, dtx = 32 - dty * ntx;
    while (ty < nty) {
      mt_tile<W, R, SLIDE>(win.data(), tx, ty, tpk, g, top[lane]);
      tx += dtx; ty += dty;
      if (tx >= ntx) { tx -= ntx; ++ty; }
    }
  }
  float F = -INFINITY;
  for (int lane = 0; lane < 32; ++lane) F = fmaxf(F, top[lane].a0);
  if (!(F > -INFINITY)) return 1;
  const float thrF = F - MT_BAND;
  struct Ent { int idx; unsigned s; int p, q; };
  std::vector<Ent> list;
  for (int lane = 0; lane < 32; ++lane) {
    if (top[lane].a2 >= thrF) return 0;
    if (top[lane].a0 >= thrF) list.push_back({top[lane].i0, top[lane].s0, top[lane].p0, top[lane].q0});
    if (top[lane].a1 >= thrF) list.push_back({top[lane].i1, top[lane].s1, top[lane].p1, top[lane].q1});
  }
  if ((int)list.size() > MT_LIST) return 0;
  *nlist_out = (int)list.size();
  double Mstar = -1.0e300;
  for (const Ent& c : list) Mstar = fmax(Mstar, mt_ncc_star(W2, c.s, c.p, c.q, T, rd1));
  const float fm = fabsf((float)Mstar);
  const double ulp = (double)(nextafterf(fm, 3.0e38f) - fm);
  const double thr = Mstar - (2.0 * ulp + 4.0e-12);
  float best = -1.0f;
  int bestkey = 0x7fffffff;
  const uint8_t* winb = reinterpret_cast<const uint8_t*>(win.data());
  for (const Ent& c : list) {
    if (!(mt_ncc_star(W2, c.s, c.p, c.q, T, rd1) >= thr)) continue;
    const int jv = c.idx / cw, iu = c.idx - jv * cw;
    const float s1 = exact_score(tb, m1, winb, MT_WSW * 4, jv * MT_WSW * 4 + iu, c.p);
    const int key = (G.ilo + iu - G.i0) * G.nv + (G.jlo + jv - G.j0);
    if (s1 > best || (s1 == best && key < bestkey)) { best = s1; bestkey = key; }
  }
  *best_out = best;
  if (bestkey != 0x7fffffff && G.nv > 0) {
    *bi = G.i0 + bestkey / G.nv;
    *bj = G.j0 + bestkey % G.nv;
  }
  return 1;
}
}  // namespace

// Same argument meaning as ekf_match_batch (include/ekf_b200.h); template side 11 only.  decided[i] = 0: the warp matcher would
// hand feature i to the CTA matcher (out_* untouched); nlist[i] = band candidates that saw double precision; tile_rows = 4 (the
// warp kernels) or 2 (the CTA-per-feature kernel: 231 tiles over 256 threads); x 10: with sliding window sums (mt_tile SLIDE).
extern "C" int emu_match_batch(const uint8_t* frames, int n_frames, int width, int height, int stride, const uint8_t* templates,
                               int fpf, const double* h, const double* S, float sigma_size, float thr, float clampv,
                               int32_t* out_uv, float* out_score, int32_t* decided, int32_t* nlist, int tile_rows) {
  for (int idx = 0; idx < n_frames * fpf; ++idx) {
    MatchJob jb;
    jb.tmap = nullptr; jb.frame_index = idx / fpf;
    jb.frame = frames + (size_t)(idx / fpf) * height * stride;
    jb.fw = width; jb.fh = height; jb.fstride = stride;
    jb.tmpl = templates + (size_t)idx * W2;
    jb.hu = h[2 * idx]; jb.hv = h[2 * idx + 1];
    for (int c = 0; c < 4; ++c) jb.S[c] = S[4 * idx + c];
    float best; int bi, bj, nl;
    switch (tile_rows) {
      case 2: decided[idx] = emu_one<2, false>(jb, sigma_size, clampv, &best, &bi, &bj, &nl); break;
      case 20: decided[idx] = emu_one<2, true>(jb, sigma_size, clampv, &best, &bi, &bj, &nl); break;
      case 40: decided[idx] = emu_one<4, true>(jb, sigma_size, clampv, &best, &bi, &bj, &nl); break;
      default: decided[idx] = emu_one<4, false>(jb, sigma_size, clampv, &best, &bi, &bj, &nl); break;
    }
    nlist[idx] = nl;
    if (decided[idx]) {
      const bool accept = !(best < thr);
      out_uv[2 * idx] = accept ? bi : -1;
      out_uv[2 * idx + 1] = accept ? bj : -1;
      out_score[idx] = best;
    }
  }
  return 0;
}
