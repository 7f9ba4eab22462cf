// oracle/oracle_capi.cpp — C interface over ekf_oracle.hpp for ctypes (tests / bench baseline).
// TEST INFRASTRUCTURE ONLY; see the header of ekf_oracle.hpp.
//
// kind: 0 = Filter<double,float>  (parity target: fp64 state, float matcher decisions)
//       1 = Filter<float,float>   (what the fp32 reference computes)
//       2 = Filter<double,double> (matches a reference build with every float widened)
#include <chrono>
#include <cstdio>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ekf_b200.h"
#include "ekf_oracle.hpp"

using namespace ekf_oracle;

namespace {

Config to_config(const ekf_config& c) {
  Config o;
  o.sigma_vx = c.sigma_vx; o.sigma_vy = c.sigma_vy; o.sigma_vz = c.sigma_vz;
  o.sigma_wx = c.sigma_wx; o.sigma_wy = c.sigma_wy; o.sigma_wz = c.sigma_wz;
  o.rho_0 = c.rho_0; o.sigma_rho_0 = c.sigma_rho_0; o.T_camera = c.T_camera;
  o.fx = c.fx; o.fy = c.fy; o.u0 = c.u0; o.v0 = c.v0;
  o.k1 = c.k1; o.k2 = c.k2; o.k3 = c.k3; o.p1 = c.p1; o.p2 = c.p2;
  o.ncc_threshold = c.ncc_threshold; o.search_clamp = c.search_clamp; o.ransac_p = c.ransac_p;
  o.li_threshold_factor = c.li_threshold_factor; o.hi_chi2_threshold = c.hi_chi2_threshold;
  o.quality_ratio = c.quality_ratio; o.linearity_threshold = c.linearity_threshold;
  o.window_size = c.window_size; o.sigma_pixel = c.sigma_pixel; o.kernel_size = c.kernel_size;
  o.sigma_size = c.sigma_size; o.scale = c.scale; o.nInitFeatures = c.nInitFeatures;
  o.min_features = c.min_features; o.max_features = c.max_features; o.forsePlane = c.forsePlane;
  o.ransac_nhyp0 = c.ransac_nhyp0; o.xyz_conversion = c.xyz_conversion; o.abs_int_quirk = c.abs_int_quirk;
  return o;
}

struct Base {
  virtual ~Base() {}
  virtual void capture(const uint8_t*, int, int, int, double) = 0;
  virtual int add_feature(float, float) = 0;
  virtual int add_features_structured(const float*, int) = 0;
  virtual void remove_feature(int) = 0;
  virtual void predict(const double*, const double*, int) = 0;
  virtual int match() = 0;
  virtual void update_after_match(const uint32_t*, int) = 0;
  virtual void inject_match(int, double, double, int) = 0;
  virtual void convert(int) = 0;
  virtual int n() = 0;
  virtual int nfeat() = 0;
  virtual void get_full(double*, double*, int) = 0;
  virtual void set_full(const double*, const double*, int) = 0;
  virtual void get_feature(int, ekf_feature_info*) = 0;
  virtual void get_template(int, int, uint8_t*) = 0;
  virtual void get_S_blocks(double*) = 0;
  virtual int get_St(double*, int) = 0;
  virtual void stats(ekf_step_stats*) = 0;
  virtual double cov_param() = 0;
  virtual double dt() = 0;
  virtual double min_margin() = 0;
  virtual void import_state(int, int, const double*, const double*, const int*, const int*, const int*, const int*,
                            const int*, const float*, const uint8_t*, int) = 0;
  virtual int num_deleted() = 0;
  virtual void get_deleted(int i, int* real_index, double* xyz, double* cov9) = 0;
  virtual int points_features(double* out, int cap_rows) = 0;
  virtual void rts_epoch(double* mu13, double* sg13, const double* mus13, const double* sgs13, const double* dts,
                         const double* drs, double dT) = 0;
};

template <class S, class MF>
struct Impl : Base {
  Filter<S, MF> f;
  int n_removed_last = 0;
  explicit Impl(const Config& c) : f(c) {}
  void capture(const uint8_t* g, int w, int h, int stride, double stamp) override {
    f.captureNewFrame(g, w, h, stride, stamp, stamp >= 0);
  }
  int add_feature(float u, float v) override { return f.addFeature(u, v); }
  int add_features_structured(const float* uv, int n) override { return f.addFeaturesStructured(uv, n); }
  void remove_feature(int i) override { f.removeFeature(i); }
  void predict(const double* dv, const double* dw, int vc) override {
    S a[3] = {S(dv[0]), S(dv[1]), S(dv[2])}, b[3] = {S(dw[0]), S(dw[1]), S(dw[2])};
    f.predict(a, b, vc != 0);
  }
  int match() override { return f.match(); }
  void update_after_match(const uint32_t* p, int np) override {
    const int before = f.numOfFeatures();
    f.update_after_match(p, np);
    n_removed_last = before - f.numOfFeatures();
  }
  void inject_match(int i, double zu, double zv, int acc) override {
    auto& p = f.patches[i];
    if (acc) {
      p.isInInnovation = true; p.isInLi = p.isInHi = false;
      p.z[0] = S(zu); p.z[1] = S(zv);
      p.center_x = float(zu); p.center_y = float(zv);
    } else {
      p.center_x = p.center_y = -1;
      p.setIsInInnovation(false);
    }
  }
  void convert(int i) override {
    if (i < 0) f.convert2XYZ_ifLinearAll();
    else if (!f.patches[i].coding) f.convert2XYZ_ifLinear(i);
  }
  int n() override { return f.mu.r; }
  int nfeat() override { return f.numOfFeatures(); }
  void get_full(double* mu, double* sg, int ld) override {
    const int nn = f.mu.r;
    for (int i = 0; i < nn; ++i) mu[i] = double(f.mu[i]);
    if (sg)
      for (int i = 0; i < nn; ++i)
        for (int j = 0; j < nn; ++j) sg[size_t(i) * ld + j] = double(f.Sigma(i, j));
  }
  void set_full(const double* mu, const double* sg, int ld) override {
    const int nn = f.mu.r;
    for (int i = 0; i < nn; ++i) f.mu[i] = S(mu[i]);
    for (int i = 0; i < nn; ++i)
      for (int j = 0; j < nn; ++j) f.Sigma(i, j) = S(sg[size_t(i) * ld + j]);
  }
  void get_feature(int i, ekf_feature_info* o) override {
    std::memset(o, 0, sizeof(*o));
    const auto& p = f.patches[i];
    const int pos = p.position_in_state, fs = p.coding ? 3 : 6;
    o->position_in_state = pos; o->position_in_z = p.position_in_z; o->coding = p.coding ? 1 : 0;
    o->n_tot = p.n_tot; o->n_find = p.n_find; o->real_index = p.real_index;
    o->is_in_innovation = p.isInInnovation; o->is_in_li = p.isInLi; o->is_in_hi = p.isInHi;
    o->remove_flag = p.removeFlag;
    o->center[0] = p.center_x; o->center[1] = p.center_y;
    o->quality_index = p.quality_index; o->last_ncc = p.last_ncc;
    for (int a = 0; a < 2; ++a) { o->z[a] = double(p.z[a]); o->h[a] = double(p.h[a]); }
    if (p.H.r == 2 && p.H.c >= pos + fs) {
      for (int a = 0; a < 2; ++a) {
        for (int c = 0; c < 7; ++c) o->H[a * 13 + c] = double(p.H(a, c));
        for (int c = 0; c < fs; ++c) o->H[a * 13 + 7 + c] = double(p.H(a, pos + c));
      }
    }
    for (int a = 0; a < fs; ++a) o->state[a] = double(f.mu[pos + a]);
    for (int a = 0; a < fs; ++a)
      for (int b = 0; b < fs; ++b) o->cov[a * 6 + b] = double(f.Sigma(pos + a, pos + b));
  }
  void get_template(int i, int which, uint8_t* out) override {
    const auto& p = f.patches[i];
    const auto& v = which ? p.matching_patch : p.patch;
    if (!v.empty()) std::memcpy(out, v.data(), v.size());
  }
  void get_S_blocks(double* out) override {
    for (size_t i = 0; i < f.patches.size(); ++i) {
      const auto& p = f.patches[i];
      for (int a = 0; a < 4; ++a) out[4 * i + a] = 0;
      if (p.isInInnovation && p.position_in_z + 1 < f.St.r) {
        const int z = p.position_in_z;
        out[4 * i + 0] = double(f.St(z, z)); out[4 * i + 1] = double(f.St(z, z + 1));
        out[4 * i + 2] = double(f.St(z + 1, z)); out[4 * i + 3] = double(f.St(z + 1, z + 1));
      }
    }
  }
  int get_St(double* out, int cap) override {
    const int k = f.St.r;
    if (out && cap >= k * k)
      for (int i = 0; i < k * k; ++i) out[i] = double(f.St.d[i]);
    return k;
  }
  void stats(ekf_step_stats* s) override {
    std::memset(s, 0, sizeof(*s));
    int ninn = 0;
    for (auto& p : f.patches) ninn += p.isInInnovation;
    s->n_in_innovation_predict = ninn;
    s->n_matched = f.last_n_matched; s->n_li = f.last_n_li; s->n_hi = f.last_n_hi;
    s->ransac_hypotheses = f.last_ransac_hyps; s->n_removed = n_removed_last;
    s->topup_request = f.topup_request; s->blur_requests = f.blur_requests;
  }
  double cov_param() override { return double(f.Covariance_Parameter()); }
  int num_deleted() override { return (int)f.deleted_patches.size(); }
  void get_deleted(int i, int* real_index, double* xyz, double* cov9) override {
    const auto& p = f.deleted_patches[i];
    *real_index = p.real_index;
    for (int c = 0; c < 3; ++c) xyz[c] = double(p.XYZ_pos[c]);
    for (int c = 0; c < 9; ++c) cov9[c] = double(p.cov_4_delete[c]);
  }
  int points_features(double* out, int cap_rows) override {
    Mat<S> pts = f.getPointsFeatures();
    if (out && cap_rows >= pts.r)
      for (int i = 0; i < pts.r * 12; ++i) out[i] = double(pts.d[i]);
    return pts.r;
  }
  void rts_epoch(double* mu13, double* sg13, const double* mus13, const double* sgs13, const double* dts, const double* drs,
                 double dT) override {
    Mat<S> MU(13, 1), SG(13, 13), MUS(13, 1), SGS(13, 13);
    for (int i = 0; i < 13; ++i) { MU[i] = S(mu13[i]); MUS[i] = S(mus13[i]); }
    for (int i = 0; i < 169; ++i) { SG.d[i] = S(sg13[i]); SGS.d[i] = S(sgs13[i]); }
    const S a[3] = {S(dts[0]), S(dts[1]), S(dts[2])}, b[3] = {S(drs[0]), S(drs[1]), S(drs[2])};
    f.rts_epoch(MU, SG, MUS, SGS, a, b, dT);
    for (int i = 0; i < 13; ++i) mu13[i] = double(MU[i]);
    for (int i = 0; i < 169; ++i) sg13[i] = double(SG.d[i]);
  }
  double dt() override { return f.dT; }
  double min_margin() override { return f.min_margin; }
  // Replace mu / Sigma / the patch list wholesale (tests and the bench baseline seed large maps this
  // way: the dense addFeature of the reference costs O(n^3) per feature).
  void import_state(int n, int N, const double* mu, const double* sg, const int* pos, const int* coding, const int* ntot,
                    const int* nfind, const int* real, const float* centers, const uint8_t* tmpl, int patchnumbre) override {
    f.mu = Mat<S>(n, 1);
    f.Sigma = Mat<S>(n, n);
    for (int i = 0; i < n; ++i) f.mu[i] = S(mu[i]);
    for (size_t i = 0; i < size_t(n) * n; ++i) f.Sigma.d[i] = S(sg[i]);
    f.patches.clear();
    const int w = f.windowsSize;
    for (int i = 0; i < N; ++i) {
      Patch<S> p;
      p.w = w;
      p.patch.assign(tmpl + size_t(i) * w * w, tmpl + size_t(i + 1) * w * w);
      p.position_in_state = pos[i]; p.coding = coding[i] != 0; p.n_tot = ntot[i]; p.n_find = nfind[i];
      p.real_index = real[i]; p.center_x = centers[2 * i]; p.center_y = centers[2 * i + 1];
      f.patches.push_back(p);
    }
    f.patchnumbre = patchnumbre;
  }
};

}  // namespace

extern "C" {

void* orc_create(const ekf_config* cfg, int kind) {
  Config c = to_config(*cfg);
  if (kind == 1) return new Impl<float, float>(c);
  if (kind == 2) return new Impl<double, double>(c);
  return new Impl<double, float>(c);
}
void orc_destroy(void* h) { delete static_cast<Base*>(h); }
void orc_capture(void* h, const uint8_t* g, int w, int hh, int stride, double stamp) { static_cast<Base*>(h)->capture(g, w, hh, stride, stamp); }
int orc_add_feature(void* h, float u, float v) { return static_cast<Base*>(h)->add_feature(u, v); }
int orc_add_features_structured(void* h, const float* uv, int n) { return static_cast<Base*>(h)->add_features_structured(uv, n); }
void orc_remove_feature(void* h, int i) { static_cast<Base*>(h)->remove_feature(i); }
void orc_predict(void* h, const double* dv, const double* dw, int vc) { static_cast<Base*>(h)->predict(dv, dw, vc); }
int orc_match(void* h) { return static_cast<Base*>(h)->match(); }
void orc_update_after_match(void* h, const uint32_t* p, int n) { static_cast<Base*>(h)->update_after_match(p, n); }
void orc_update(void* h, const uint32_t* p, int n) { static_cast<Base*>(h)->match(); static_cast<Base*>(h)->update_after_match(p, n); }
void orc_inject_match(void* h, int i, double zu, double zv, int acc) { static_cast<Base*>(h)->inject_match(i, zu, zv, acc); }
void orc_convert2xyz(void* h, int i) { static_cast<Base*>(h)->convert(i); }
int orc_state_dim(void* h) { return static_cast<Base*>(h)->n(); }
int orc_num_features(void* h) { return static_cast<Base*>(h)->nfeat(); }
void orc_get_full(void* h, double* mu, double* sg, int ld) { static_cast<Base*>(h)->get_full(mu, sg, ld); }
void orc_set_full(void* h, const double* mu, const double* sg, int ld) { static_cast<Base*>(h)->set_full(mu, sg, ld); }
void orc_get_feature(void* h, int i, ekf_feature_info* o) { static_cast<Base*>(h)->get_feature(i, o); }
void orc_get_template(void* h, int i, int which, uint8_t* o) { static_cast<Base*>(h)->get_template(i, which, o); }
void orc_get_S_blocks(void* h, double* o) { static_cast<Base*>(h)->get_S_blocks(o); }
int orc_get_St(void* h, double* o, int cap) { return static_cast<Base*>(h)->get_St(o, cap); }
void orc_get_step_stats(void* h, ekf_step_stats* s) { static_cast<Base*>(h)->stats(s); }
double orc_covariance_parameter(void* h) { return static_cast<Base*>(h)->cov_param(); }
double orc_get_dt(void* h) { return static_cast<Base*>(h)->dt(); }
double orc_min_margin(void* h) { return static_cast<Base*>(h)->min_margin(); }
int orc_num_deleted(void* h) { return static_cast<Base*>(h)->num_deleted(); }
void orc_get_deleted(void* h, int i, int* real_index, double* xyz, double* cov9) { static_cast<Base*>(h)->get_deleted(i, real_index, xyz, cov9); }
int orc_points_features(void* h, double* out, int cap_rows) { return static_cast<Base*>(h)->points_features(out, cap_rows); }
void orc_rts_epoch(void* h, double* mu13, double* sg13, const double* mus13, const double* sgs13, const double* dts,
                   const double* drs, double dT) { static_cast<Base*>(h)->rts_epoch(mu13, sg13, mus13, sgs13, dts, drs, dT); }

// Patch::findMatch for a batch (BASELINE config 5), host buffers.  Same argument meaning as
// ekf_match_batch of include/ekf_b200.h.  kind_mf: 0 float matcher, 1 double matcher.
void orc_match_batch(const uint8_t* frames, int n_frames, int width, int height, int stride,
                     const uint8_t* templates, int fpf, int w, const double* h, const double* Sm,
                     float sigma_size, float ncc_threshold, float search_clamp, int32_t* out_uv,
                     float* out_score, int kind_mf) {
  Config c;
  c.window_size = w; c.ncc_threshold = ncc_threshold; c.search_clamp = search_clamp;
#pragma omp parallel for schedule(dynamic, 4)
  for (int idx = 0; idx < n_frames * fpf; ++idx) {
    const int fr = idx / fpf;
    auto run = [&](auto& filt) {
      filt.frame.w = width; filt.frame.h = height; filt.frame.stride = stride;
      filt.frame.p = frames + size_t(fr) * height * stride;
      using PT = typename std::remove_reference<decltype(filt.patches)>::type::value_type;
      PT p;
      p.w = w; p.matching_w = w;
      p.matching_patch.assign(templates + size_t(idx) * w * w, templates + size_t(idx + 1) * w * w);
      p.patch = p.matching_patch;
      p.h[0] = h[2 * idx]; p.h[1] = h[2 * idx + 1];
      p.isInInnovation = true;
      Mat<double> cov(2, 2);
      for (int a = 0; a < 4; ++a) cov.d[a] = Sm[4 * idx + a];
      filt.findMatch(p, cov, sigma_size);
      out_uv[2 * idx] = int(p.center_x); out_uv[2 * idx + 1] = int(p.center_y);
      out_score[idx] = p.last_ncc;
    };
    if (kind_mf == 1) { Filter<double, double> filt(c); run(filt); }
    else { Filter<double, float> filt(c); run(filt); }
  }
}

void orc_import_state(void* h, int n, int N, const double* mu, const double* sg, const int* pos, const int* coding,
                      const int* ntot, const int* nfind, const int* real, const float* centers, const uint8_t* tmpl,
                      int patchnumbre) {
  static_cast<Base*>(h)->import_state(n, N, mu, sg, pos, coding, ntot, nfind, real, centers, tmpl, patchnumbre);
}

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
}
