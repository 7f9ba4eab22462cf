// tests/keyframe_stub.cpp — drives host/keyframe_recorder.hpp with a stub filter (no GPU): reads a trajectory
// file written by tests/test_keyframes.py and writes the reference-format files into the given directory.
//   per frame: 14 state doubles, 196 covariance doubles, 1 covariance-parameter double; then rows, rows*12 points
#include <cstdio>
#include <vector>
#include "../ekf-monoslam_for_3d-reconstruction_b200/host/keyframe_recorder.hpp"
struct Stub {
  std::vector<double> state, sigma, points;
  double cov = 0;
  int rows = 0;
  std::vector<int> Point4sba = std::vector<int>(3, 0);
  std::vector<double> getState() { return state; }
  std::vector<double> getSigma() { return sigma; }
  double Covariance_Parameter() { return cov; }
  std::vector<double> getPointsFeatures(int* r) { *r = rows; return points; }
};
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int nframes = 0;
  if (fread(&nframes, sizeof(int), 1, f) != 1) return 4;
  Stub s;
  s.state.resize(14); s.sigma.resize(196);
  std::vector<std::string> saved;
  ekf_b200::KeyframeRecorder rec(argv[2], [&](const std::string& path, const uint8_t*, int, int, int, int) { saved.push_back(path); });
  const uint8_t img[4] = {1, 2, 3, 4};
  for (int t = 1; t <= nframes; ++t) {
    if (fread(s.state.data(), 8, 14, f) != 14 || fread(s.sigma.data(), 8, 196, f) != 196 || fread(&s.cov, 8, 1, f) != 1) return 5;
    rec.onFrame(s, t, img, 2, 2, 2, 1);
  }
  if (fread(&s.rows, sizeof(int), 1, f) != 1) return 6;
  s.points.resize((size_t)s.rows * 12);
  if (fread(s.points.data(), 8, s.points.size(), f) != s.points.size()) return 7;
  fclose(f);
  rec.finish(s);
  for (const auto& p : saved) printf("%s\n", p.c_str());
  return 0;
}
