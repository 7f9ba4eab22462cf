// Compile-and-link check of host/vslam_filter.hpp (tests/test_host_class.py).  With a GPU it also
// runs one predict/update; without one ekf_create fails loudly and the program reports that.
#include <cstdio>
#include <vector>

#include "../ekf-monoslam_for_3d-reconstruction_b200/host/vslam_filter.hpp"

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
    printf("gpu ok: added %d features, n = %d, |q| part = %.6f, cov par = %g\n", added, slam.stateDim(), mu[6], slam.Covariance_Parameter());
    return added == 8 ? 0 : 2;
  } catch (const std::exception& e) {
    printf("no usable device: %s\n", e.what());
    return 3;
  }
}
