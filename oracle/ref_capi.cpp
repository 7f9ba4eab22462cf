// oracle/ref_capi.cpp — TEST INFRASTRUCTURE ONLY.
//
// C harness around the REFERENCE'S OWN `class VSlamFilter`, compiled from the unmodified sources
// under /root/reference/mono-slam/src (vslamRansac.cpp, Patch.cpp, camModel.cpp, utils.cpp, libblur.cpp, RosVSLAMRansac.cpp) against
// the API stand-ins in oracle/shim/ (Eigen3, OpenCV, ROS and libconfig++ are not installed here).
// oracle/build_ref.py builds two variants into oracle/_ref/:
//   libref_f32.so  the sources as written (fp32 state, what the reference really runs)
//   libref_f64.so  -DEKF_SHIM_DOUBLE: every `float` of the reference re-typed to double — the fp64
//                  parity target BASELINE.json names; compared with the oracle's all-double kind
// tests/test_oracle_vs_ref.py and oracle/gen_golden.py drive it; nothing in the product does.
// Replaced pieces, all outside the EKF arithmetic: ConfigVSLAM's libconfig reader (values come from
// the ekf_config struct) and rand()/srand() (injected picks).
#include "../include/ekf_b200.h"  // before the shims: keeps `float` fields of the ABI structs float

#include <stdarg.h>

#include <vector>

// all shim / system headers first (include guards), then open up the reference class: `St` and
// friends are private members of VSlamFilter (vslamRansac.hpp:29-60) that the harness must read
#include <opencv2/opencv.hpp>
#include <eigen3/Eigen/Dense>
#define private public
#define protected public
#define class struct
#include "RosVSLAMRansac.hpp"   // includes vslamRansac.hpp; RosVSLAM::getPointsFeatures (RosVSLAMRansac.cpp:340-418)
#include "libblur.h"
#undef class
#undef private
#undef protected
#ifdef float
#undef float
#endif

using Eigen::shim_real;

// ---- pieces of the reference that are replaced -------------------------------------------------
static ekf_config g_cfg;  // consumed by the next ConfigVSLAM constructed

ConfigVSLAM::ConfigVSLAM(char*) {  // stands in for ConfigVSLAM.cpp:26-151 (libconfig++ reader)
  sigma_vx = (shim_real)g_cfg.sigma_vx; sigma_vy = (shim_real)g_cfg.sigma_vy; sigma_vz = (shim_real)g_cfg.sigma_vz;
  sigma_wx = (shim_real)g_cfg.sigma_wx; sigma_wy = (shim_real)g_cfg.sigma_wy; sigma_wz = (shim_real)g_cfg.sigma_wz;
  rho_0 = (shim_real)g_cfg.rho_0; sigma_rho_0 = (shim_real)g_cfg.sigma_rho_0;
  window_size = g_cfg.window_size; sigma_pixel = g_cfg.sigma_pixel; kernel_size = g_cfg.kernel_size;
  sigma_size = g_cfg.sigma_size; scale = g_cfg.scale; T_camera = (shim_real)g_cfg.T_camera;
  nInitFeatures = g_cfg.nInitFeatures; min_features = g_cfg.min_features; max_features = g_cfg.max_features;
  forsePlane = g_cfg.forsePlane;
  camParams.fx = (shim_real)g_cfg.fx; camParams.fy = (shim_real)g_cfg.fy; camParams.u0 = (shim_real)g_cfg.u0;
  camParams.v0 = (shim_real)g_cfg.v0; camParams.k1 = (shim_real)g_cfg.k1; camParams.k2 = (shim_real)g_cfg.k2;
  camParams.k3 = (shim_real)g_cfg.k3; camParams.p1 = (shim_real)g_cfg.p1; camParams.p2 = (shim_real)g_cfg.p2;
}


static std::vector<uint32_t> g_picks;
static size_t g_pick_pos = 0;
extern "C" int ekf_shim_rand(void) {
  if (g_picks.empty()) return 0;
  const uint32_t v = g_picks[g_pick_pos % g_picks.size()];
  ++g_pick_pos;
  return (int)(v & 0x7fffffffu);
}
extern "C" void ekf_shim_srand(unsigned int) {}
static int g_log_level = 3;  // 3 = silent
extern "C" void ekf_shim_log(int level, const char* fmt, ...) {
  if (level < g_log_level) return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
}

struct RefHandle {
  VSlamFilter* f;       // the RosVSLAM object seen through its base: predict / update resolve to VSlamFilter's (non-virtual)
  RosVSLAM* ros = nullptr;
  int n_hyp_last = 0;
};

extern "C" {

int ref_scalar_bytes(void) { return (int)sizeof(shim_real); }
void ref_set_log_level(int l) { g_log_level = l; }

void* ref_create(const ekf_config* cfg) {
  g_cfg = *cfg;
  RefHandle* h = new RefHandle;
  h->ros = new RosVSLAM(nullptr);
  h->f = h->ros;
  return h;
}
void ref_destroy(void* hh) {
  RefHandle* h = static_cast<RefHandle*>(hh);
  delete h->ros;
  delete h;
}
void ref_capture(void* hh, const uint8_t* gray, int w, int hgt, int stride, double stamp) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  cv::Mat view(hgt, w, CV_8UC1, (void*)gray, (size_t)stride);
  cv::Mat img = view.clone();  // the ROS node hands the filter its own copy (cv_bridge::toCvCopy)
  if (stamp >= 0) f->captureNewFrame(img, stamp);
  else f->captureNewFrame(img);
}
int ref_add_feature(void* hh, float u, float v) {
  return static_cast<RefHandle*>(hh)->f->addFeature(cv::Point2f(u, v));
}
void ref_remove_feature(void* hh, int i) { static_cast<RefHandle*>(hh)->f->removeFeature(i); }
void ref_predict(void* hh, const double* dv, const double* dw, int vcontrol) {
  Eigen::Vector3f a, b;
  for (int i = 0; i < 3; ++i) { a(i) = (shim_real)(dv ? dv[i] : 0.0); b(i) = (shim_real)(dw ? dw[i] : 0.0); }
  static_cast<RefHandle*>(hh)->f->predict(a, b, vcontrol != 0);
}
void ref_update(void* hh, const uint32_t* picks, int n) {
  g_picks.assign(picks, picks + (n > 0 ? n : 0));
  g_pick_pos = 0;
  static_cast<RefHandle*>(hh)->f->update();
  static_cast<RefHandle*>(hh)->n_hyp_last = (int)g_pick_pos;
}
int ref_last_hypotheses(void* hh) { return static_cast<RefHandle*>(hh)->n_hyp_last; }
void ref_convert2xyz(void* hh, int i) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  if (i < 0) f->convert2XYZ_ifLinearAll();
  else f->convert2XYZ_ifLinear(i);
}
int ref_state_dim(void* hh) { return static_cast<RefHandle*>(hh)->f->mu.rows(); }
int ref_num_features(void* hh) { return static_cast<RefHandle*>(hh)->f->numOfFeatures(); }
double ref_get_dt(void* hh) { return static_cast<RefHandle*>(hh)->f->getDt(); }
double ref_covariance_parameter(void* hh) { return (double)static_cast<RefHandle*>(hh)->f->Covariance_Parameter(); }

void ref_get_full(void* hh, double* mu, double* sg, int ld) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  const int n = f->mu.rows();
  for (int i = 0; i < n; ++i) mu[i] = (double)f->mu(i);
  if (sg)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) sg[(size_t)i * ld + j] = (double)f->Sigma(i, j);
}
void ref_set_full(void* hh, const double* mu, const double* sg, int ld) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  const int n = f->mu.rows();
  for (int i = 0; i < n; ++i) f->mu(i) = (shim_real)mu[i];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) f->Sigma(i, j) = (shim_real)sg[(size_t)i * ld + j];
}
// getState / getSigma as the ROS node calls them (vslamRansac.cpp:131-140)
void ref_get_state14(void* hh, double* mu14, double* sg14) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  Eigen::VectorXf m = f->getState();
  Eigen::MatrixXf s = f->getSigma();
  for (int i = 0; i < 14; ++i) mu14[i] = (double)m(i);
  for (int i = 0; i < 14; ++i)
    for (int j = 0; j < 14; ++j) sg14[i * 14 + j] = (double)s(i, j);
}

void ref_get_feature(void* hh, int i, ekf_feature_info* o) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  Patch& p = f->patches[i];
  memset(o, 0, sizeof *o);
  const int pos = p.position_in_state, fs = p.isXYZ() ? 3 : 6;
  o->position_in_state = pos; o->position_in_z = p.position_in_z; o->coding = p.isXYZ() ? 1 : 0;
  o->n_tot = p.n_tot; o->n_find = p.n_find; o->real_index = p.real_index;
  o->is_in_innovation = p.patchIsInInnovation(); o->is_in_li = p.patchIsInLi(); o->is_in_hi = p.patchIsInHi();
  o->remove_flag = p.mustBeRemove();
  o->center[0] = (float)p.center.x; o->center[1] = (float)p.center.y;
  o->quality_index = (float)p.get_quality_index();
  o->last_ncc = 0.0f;
  for (int a = 0; a < 2; ++a) { o->z[a] = (double)p.z(a); o->h[a] = (double)p.h(a); }
  if (p.H.rows() == 2 && p.H.cols() >= pos + fs) {
    for (int r = 0; r < 2; ++r) {
      for (int c = 0; c < 7; ++c) o->H[13 * r + c] = (double)p.H(r, c);
      for (int c = 0; c < fs; ++c) o->H[13 * r + 7 + c] = (double)p.H(r, pos + c);
    }
  }
  for (int a = 0; a < fs; ++a) o->state[a] = (double)f->mu(pos + a);
  for (int a = 0; a < fs; ++a)
    for (int b = 0; b < fs; ++b) o->cov[a * 6 + b] = (double)f->Sigma(pos + a, pos + b);
}
void ref_get_template(void* hh, int i, int which, uint8_t* out) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  const cv::Mat& m = which ? f->patches[i].matching_patch : f->patches[i].patch;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) out[r * m.cols + c] = m.at<uchar>(r, c);
}
// 2x2 diagonal blocks of St after predict (vslamRansac.cpp:598), zeros for features not in innovation
// archive of removed features (vslamRansac.cpp:394-404) and the RTS epoch (vslamRansac.cpp:423-449)
int ref_num_deleted(void* hh) { return (int)static_cast<RefHandle*>(hh)->f->deleted_patches.size(); }
void ref_get_deleted(void* hh, int i, int* real_index, double* xyz, double* cov9) {
  Patch& p = static_cast<RefHandle*>(hh)->f->deleted_patches[i];
  *real_index = p.real_index;
  for (int c = 0; c < 3; ++c) xyz[c] = (double)p.XYZ_pos(c);
  for (int c = 0; c < 9; ++c) cov9[c] = (double)p.cov_4_delete(c);
}
// RosVSLAM::getPointsFeatures, compiled from the reference's own RosVSLAMRansac.cpp (ROS message types are stand-ins)
int ref_points_features(void* hh, double* out, int cap_rows) {
  RosVSLAM* r = static_cast<RefHandle*>(hh)->ros;
  if (r->patches.empty()) return 0;   // the reference reads patches[size-1] unguarded (RosVSLAMRansac.cpp:349)
  MatrixXf pts = r->getPointsFeatures();
  const int rows = (int)pts.rows();
  if (out && cap_rows >= rows)
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < 12; ++j) out[i * 12 + j] = (double)pts(i, j);
  return rows;
}
void ref_rts_epoch(void* hh, double* mu13, double* sg13, const double* mus13, const double* sgs13, const double* dts,
                   const double* drs, double dT) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  VectorXf MU(13), MUS(13);
  MatrixXf SG(13, 13), SGS(13, 13);
  for (int i = 0; i < 13; ++i) { MU(i) = (shim_real)mu13[i]; MUS(i) = (shim_real)mus13[i]; }
  for (int i = 0; i < 13; ++i)
    for (int j = 0; j < 13; ++j) { SG(i, j) = (shim_real)sg13[i * 13 + j]; SGS(i, j) = (shim_real)sgs13[i * 13 + j]; }
  Vector3f a, b;
  for (int i = 0; i < 3; ++i) { a(i) = (shim_real)dts[i]; b(i) = (shim_real)drs[i]; }
  f->rts_epoch(MU, SG, MUS, SGS, a, b, dT);
  for (int i = 0; i < 13; ++i) mu13[i] = (double)MU(i);
  for (int i = 0; i < 13; ++i)
    for (int j = 0; j < 13; ++j) sg13[i * 13 + j] = (double)SG(i, j);
}
void ref_get_S_blocks(void* hh, double* out) {
  VSlamFilter* f = static_cast<RefHandle*>(hh)->f;
  const int N = (int)f->patches.size();
  for (int i = 0; i < N; ++i) {
    for (int c = 0; c < 4; ++c) out[4 * i + c] = 0.0;
    Patch& p = f->patches[i];
    if (!p.patchIsInInnovation()) continue;
    const int z = p.position_in_z;
    if (z + 2 > f->St.rows()) continue;
    out[4 * i + 0] = (double)f->St(z, z); out[4 * i + 1] = (double)f->St(z, z + 1);
    out[4 * i + 2] = (double)f->St(z + 1, z); out[4 * i + 3] = (double)f->St(z + 1, z + 1);
  }
}

// computeCorrelation / Patch::findMatch stand-alone (Patch.cpp:215-329): one feature
// blurPatch of the reference (libblur.cpp:57-81) stand-alone: w x w u8 patch, segment one -> two
void ref_blur_patch(const uint8_t* patch, int w, double x1, double y1, double x2, double y2, uint8_t* out) {
  cv::Mat pv(w, w, CV_8UC1, (void*)patch, (size_t)w);
  cv::Mat res = blurPatch(pv.clone(), cv::Point2f((shim_real)x1, (shim_real)y1), cv::Point2f((shim_real)x2, (shim_real)y2));
  for (int r = 0; r < w; ++r)
    for (int c = 0; c < w; ++c) out[r * w + c] = res.at<uchar>(r, c);
}

int ref_find_match(const uint8_t* frame, int w, int hgt, int stride, const uint8_t* tmpl, int win, const double* h2,
                   const double* S4, float sigma_size, int32_t* out_uv) {
  cv::Mat view(hgt, w, CV_8UC1, (void*)frame, (size_t)stride);
  cv::Mat img = view.clone();
  cv::Mat tv(win, win, CV_8UC1, (void*)tmpl, (size_t)win);
  Patch p(tv, cv::Point2f(0, 0), 0, 0);
  p.matching_patch = p.patch.clone();
  p.h(0) = (shim_real)h2[0]; p.h(1) = (shim_real)h2[1];
  p.setIsInInnovation(true);
  Eigen::MatrixXf S(2, 2);
  S(0, 0) = (shim_real)S4[0]; S(0, 1) = (shim_real)S4[1]; S(1, 0) = (shim_real)S4[2]; S(1, 1) = (shim_real)S4[3];
  const bool ok = p.findMatch(img, S, (shim_real)sigma_size, false);
  out_uv[0] = ok ? (int)p.center.x : -1;
  out_uv[1] = ok ? (int)p.center.y : -1;
  return ok ? 1 : 0;
}

}  // extern "C"
