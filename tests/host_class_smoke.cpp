// Compile-and-link check of host/vslam_filter.hpp (tests/test_host_class.py).  With a GPU it also
// runs one predict/update; without one ekf_create fails loudly and the program reports that.
#include <cstdio>
#include <vector>

#include "../ekf-monoslam_for_3d-reconstruction_b200/host/vslam_filter.hpp"
#include "../ekf-monoslam_for_3d-reconstruction_b200/host/keyframe_recorder.hpp"

int main() {
  try {
    ekf_config c;
    ekf_config_default(&c);
    c.xyz_conversion = 0; c.window_size = 11; c.min_features = 0;
    ekf_b200::VSlamFilter slam(&c, 16, 0);
    std::vector<uint8_t> img(640 * 480);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
    ekf_b200::GrayImage g{img.data(), 640, 480, 640};
    slam.captureNewFrame(g, 1.0);
    int added = 0;
    for (int k = 0; k < 8; ++k) added += slam.addFeature(ekf_b200::Point2f{100.f + 50.f * k, 200.f});
    slam.captureNewFrame(g, 1.0 + 1.0 / 30);
    slam.predict();
    slam.update({1u, 2u, 3u});
    std::vector<double> mu = slam.getState();
    // the caller-side rows: key-frame recorder over the real class, points matrix, archive, one RTS epoch
    ekf_b200::KeyframeRecorder rec("/tmp");
    rec.onFrame(slam, 1, img.data(), 640, 480, 640, 1);
    rec.finish(slam);
    int rows = 0;
    const std::vector<double> pts = slam.getPointsFeatures(&rows);
    const size_t n_deleted = slam.deletedPatches().size();
    double MU[13] = {0, 0, 0, 1, 0, 0, 0, 0.1, 0, 0, 0, 0.01, 0}, SG[169] = {0}, SGS[169] = {0};
    for (int i = 0; i < 13; ++i) { SG[i * 14] = 1e-3; SGS[i * 14] = 5e-4; }
    const double dts[3] = {0, 0, 0}, drs[3] = {0, 0, 0};
    slam.rts_epoch(MU, SG, MU, SGS, dts, drs, 1.0 / 30);
    if (rows < 1 || pts.size() != (size_t)rows * 12 || n_deleted != 0 || !(SG[0] > 0)) return 4;
    printf("gpu ok: added %d features, n = %d, |q| part = %.6f, cov par = %g\n", added, slam.stateDim(), mu[6], slam.Covariance_Parameter());
    return added == 8 ? 0 : 2;
  } catch (const std::exception& e) {
    printf("no usable device: %s\n", e.what());
    return 3;
  }
}
