// host/vslam_filter.hpp — C++ twin of the reference's `class VSlamFilter`
// (mono-slam/src/vslamRansac.hpp:27-141) over the C ABI of include/ekf_b200.h.
//
// Same method names, argument meaning and call order as the reference class, so that
// ImageConverter::imageCb (monoslam_ransac.cpp:382-851) and RosVSLAM (RosVSLAMRansac.cpp) keep
// working against it; Eigen / OpenCV types are replaced by plain structs and std::vector because
// neither library is a dependency of this repository (an adapter to cv::Mat / Eigen is three lines,
// see INTEGRATION.md).  Header-only; link with libekf_b200.so.  Errors throw std::runtime_error:
// there is no CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ekf_b200.h"

namespace ekf_b200 {

struct Point2f { float x, y; };
struct GrayImage { const uint8_t* data; int width, height, stride; int channels = 1; };  // stands in for cv::Mat (8-bit gray or BGR)

class VSlamFilter {
 public:
  // VSlamFilter(char* file) (vslamRansac.cpp:142): the libconfig file is replaced by the struct it fills
  explicit VSlamFilter(const ekf_config* cfg = nullptr, int feature_capacity = 128, int device = 0) {
    ekf_config c;
    if (cfg) c = *cfg; else { ekf_config_default(&c); c.xyz_conversion = 0; }
    const int rc = ekf_create(&c, feature_capacity, device, &h_);
    if (rc != EKF_OK) throw std::runtime_error("ekf_create failed with " + std::to_string(rc));
    patchnumbre = 1;
    noise_cov_factor = 0;
  }
  ~VSlamFilter() { ekf_destroy(h_); }
  VSlamFilter(const VSlamFilter&) = delete;
  VSlamFilter& operator=(const VSlamFilter&) = delete;

  int patchnumbre;       // vslamRansac.hpp:103
  int noise_cov_factor;  // vslamRansac.hpp:140
  // MatrixX3i Point4sba (vslamRansac.hpp:98), rows of (real_index, u, v): Zero(1,3) at construction
  // (vslamRansac.cpp:219); the loop that would fill it in update() is commented out in the reference (:1318-1340)
  std::vector<int> Point4sba = std::vector<int>(3, 0);

  int addFeature(Point2f pf) { int rc = ck(ekf_add_feature(h_, pf.x, pf.y)); patchnumbre += rc; return rc; }  // :309
  void removeFeature(int index) { ck(ekf_remove_feature(h_, index)); }                                           // :373
  void predict(const double dV[3] = nullptr, const double dW[3] = nullptr, bool Vcontrol = false) {             // :451
    if (Vcontrol) noise_cov_factor = 0; else noise_cov_factor++;
    ck(ekf_predict(h_, dV, dW, Vcontrol ? 1 : 0));
  }
  // update(float v_x, float w_z) (:868; both arguments are unused by the reference).  `picks` replaces rand().
  void update(const std::vector<uint32_t>& picks = {}) { ck(ekf_update(h_, picks.empty() ? nullptr : picks.data(), (int)picks.size())); }
  void captureNewFrame(const GrayImage& f) { captureNewFrame(f, -1.0); }                                          // :234
  void captureNewFrame(const GrayImage& f, double time_stamp) {                                                  // :226
    ck(f.channels == 3 ? ekf_capture_frame_bgr(h_, f.data, f.width, f.height, f.stride, time_stamp)
                       : ekf_capture_frame(h_, f.data, f.width, f.height, f.stride, time_stamp));
  }
  std::vector<double> getState() { std::vector<double> v(EKF_STATE_DIM); ck(ekf_get_state(h_, v.data())); return v; }          // :135
  std::vector<double> getSigma() { std::vector<double> v(EKF_STATE_DIM * EKF_STATE_DIM); ck(ekf_get_sigma(h_, v.data())); return v; }  // :131
  void convert2XYZ_ifLinear(int index) { ck(ekf_convert2xyz_if_linear(h_, index)); }                             // :741
  void convert2XYZ_ifLinearAll() { ck(ekf_convert2xyz_if_linear_all(h_)); }                                      // :775
  int numOfFeatures() { return ekf_num_features(h_); }                                                           // :127
  float Covariance_Parameter() { double p = 0; ck(ekf_covariance_parameter(h_, &p)); return (float)p; }          // :841
  Point2f returnCentrPatchIndx(int i) { float c[2]; ck(ekf_get_center(h_, i, c)); return Point2f{c[0], c[1]}; }  // vslamRansac.hpp:136
  double getDt() { return ekf_get_dt(h_); }                                                                      // :247
  int findNewFeatures(int num = -1) { return ck(ekf_find_new_features(h_, num)); }                              // :783
  // the `num` update() passed to findNewFeatures in its last call (:1314), 0 if it did not top up
  int topupRequest() { ekf_step_stats s; ck(ekf_get_step_stats(h_, &s)); return s.topup_request; }

  // what RosVSLAM reads from the protected members (RosVSLAMRansac.cpp:19-21,68,114,177-183)
  ekf_feature_info feature(int i) { ekf_feature_info f; ck(ekf_get_feature(h_, i, &f)); return f; }
  int stateDim() { return ekf_state_dim(h_); }
  void getFull(std::vector<double>& mu, std::vector<double>& Sigma) {
    const int n = stateDim();
    mu.resize(n); Sigma.resize((size_t)n * n);
    ck(ekf_get_full(h_, mu.data(), Sigma.data(), n));
  }
  // RosVSLAM::getPointsFeatures (RosVSLAMRansac.cpp:340-418): rows x 12, row-major
  std::vector<double> getPointsFeatures(int* rows_out = nullptr) {
    int rows = 0;
    ck(ekf_get_points_features(h_, nullptr, 0, &rows));
    std::vector<double> pts((size_t)rows * 12);
    ck(ekf_get_points_features(h_, pts.data(), rows, &rows));
    if (rows_out) *rows_out = rows;
    return pts;
  }
  // VSlamFilter::deleted_patches (vslamRansac.cpp:394-404)
  std::vector<ekf_deleted_info> deletedPatches() {
    std::vector<ekf_deleted_info> v(ekf_num_deleted(h_));
    for (size_t i = 0; i < v.size(); ++i) ck(ekf_get_deleted(h_, (int)i, &v[i]));
    return v;
  }
  // rts_epoch(MU, SIGMA, MU_S, SIGMA_S, dTspeed, dRspeed, deltaT) (:423): 13-dimensional camera states, in place
  void rts_epoch(double MU[13], double SIGMA[169], const double MU_S[13], const double SIGMA_S[169], const double dTspeed[3],
                 const double dRspeed[3], double deltaT) {
    ck(ekf_rts_epoch(h_, MU, SIGMA, MU_S, SIGMA_S, dTspeed, dRspeed, deltaT));
  }
  ekf_handle* handle() { return h_; }

 private:
  int ck(int rc) {
    if (rc < 0) throw std::runtime_error(std::string("ekf_b200: ") + ekf_last_error(h_) + " (" + std::to_string(rc) + ")");
    return rc;
  }
  ekf_handle* h_ = nullptr;
};

}  // namespace ekf_b200
