// host/keyframe_recorder.hpp — the caller-side data formats of the reference's ROS node (SURVEY.md §8(f)
// item 4): the key-frame selector of ImageConverter::imageCb (monoslam_ransac.cpp:585-687; :689-752 is commented out there) and the
// writers of nodes_and_prjcts.txt / cams_cov.txt / cams_cov2.txt / points.txt (monoslam_ransac.cpp:232-236,
// 262-275) that sparse_bundle_adjustment/src/nodes/sba_add.cpp:76-180 consumes.  Host-side logic over the
// filter's accessors; Eigen's default stream format (precision 6, columns right-aligned to the widest
// coefficient, one row per line) is reproduced so the files are token-identical for the consumer's `>>`.
//
// Differences from the reference, all forced: poses are doubles (the fp64 parity target) formatted at the
// same 6 significant digits; images go through a caller-supplied writer (no OpenCV here: cv::imwrite in the
// ROS node); `Filter` is a template parameter so that tests can drive the selector without a GPU.
// Parity: PINNED against the reference's own selector — monoslam_ransac.cpp:585, 609-687 (+ quat2vec / poses_diff, :40-60)
// are compiled from the unmodified source into oracle/_ref/libref_selector_f64.so (oracle/build_ref_selector.py) and
// tests/test_keyframe_pinned.py requires byte-identical nodes_and_prjcts.txt / cams_cov.txt / cams_cov2.txt and the same
// image names on trajectories that take every branch; committed copies of the reference's outputs (tests/golden/keyframes)
// repeat the check where /root/reference is absent.  Eigen's number formatting itself stays unpinned (absent dependency).
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

namespace ekf_b200 {

// quat2vec (monoslam_ransac.cpp:40-50): rotation vector of a unit quaternion (w first)
inline void quat2vec(const double q[4], double v[3]) {
  const double n = std::acos(q[0]) * 2;
  if (n > 0.0001) {
    const double n1 = n / std::sin(n / 2);
    v[0] = q[1] * n1; v[1] = q[2] * n1; v[2] = q[3] * n1;
  } else {
    v[0] = v[1] = v[2] = 0.0;
  }
}
// poses_diff (monoslam_ransac.cpp:52-60): 3.33 * translation + rotation-vector change in degrees (L1)
inline double poses_diff(const double old7[7], const double new7[7], const double last_rot[3]) {
  double d2 = 0;
  for (int i = 0; i < 3; ++i) d2 += (old7[i] - new7[i]) * (old7[i] - new7[i]);
  const double a = std::sqrt(d2) * 3.33;
  double v[3];
  quat2vec(new7 + 3, v);
  double s = a;
  for (int i = 0; i < 3; ++i) s += std::fabs((last_rot[i] - v[i]) * 57.29577951308232);
  return s;
}

// Eigen::operator<<(ostream, DenseBase) with the default IOFormat: every coefficient printed at the stream's
// precision, padded to the widest one, " " between columns, "\n" between rows, nothing after the last row.
template <class T>
inline std::string eigen_format(const T* m, int rows, int cols, int precision = 6) {
  std::vector<std::string> cell((size_t)rows * cols);
  size_t width = 0;
  for (int i = 0; i < rows * cols; ++i) {
    std::ostringstream o;
    o.precision(precision);
    o << m[i];
    cell[i] = o.str();
    if (cell[i].size() > width) width = cell[i].size();
  }
  std::string out;
  for (int r = 0; r < rows; ++r) {
    if (r) out += "\n";
    for (int c = 0; c < cols; ++c) {
      if (c) out += " ";
      const std::string& s = cell[(size_t)r * cols + c];
      out.append(width - s.size(), ' ');
      out += s;
    }
  }
  return out;
}

class KeyframeRecorder {
 public:
  using ImageWriter = std::function<void(const std::string& path, const uint8_t* data, int w, int h, int stride, int channels)>;

  // ImageConverter::ImageConverter (monoslam_ransac.cpp:183-236): thresholds and the files it opens
  explicit KeyframeRecorder(const std::string& dir = ".", ImageWriter writer = nullptr) : dir_(dir), write_image_(writer) {
    node_proj.open(dir_ + "/nodes_and_prjcts.txt");
    cov_cams.open(dir_ + "/cams_cov.txt");
    cov_cams2.open(dir_ + "/cams_cov2.txt");
    f_points.open(dir_ + "/points.txt");
    min_projs.assign(3, 0);
  }

  float MoveThresh = 18;                   // :195
  int Num_of_points_thershold = 10;        // :186 (only read by the reference's commented-out rule)
  double min_cov_for_pose = 10000000;      // :187
  std::vector<int> key_frames;             // ids written so far (diagnostic, not in the reference)

  // One camera frame after slam.update(): monoslam_ransac.cpp:560, 585, 609-687.
  template <class Filter>
  void onFrame(Filter& slam, int frameId, const uint8_t* img, int w, int h, int stride, int channels = 1) {
    const std::vector<double> stat14 = slam.getState();
    const double DistWalked = poses_diff(last_image_pose, stat14.data(), last_vrot);          // :585
    if (DistWalked > (MoveThresh / 2) && DistWalked < MoveThresh) {                            // :609
      const double some_var = slam.Covariance_Parameter();
      if (some_var < min_cov_for_pose) candidate(slam, frameId, stat14, some_var, img, w, h, stride, channels, true);
    } else if (DistWalked >= MoveThresh) {                                                     // :627
      if (min_cov_for_pose < 1000000) {
        if ((slam.Covariance_Parameter() - min_cov_for_pose) < 0.000085) {                     // :637: the current frame is as good
          write_node(frameId, stat14.data(), nullptr, 0);
          const std::vector<double> S = slam.getSigma();
          write_cov(cov_cams2, S); write_cov(cov_cams, S);                                     // :642-644
          save_image(frameId, img, w, h, stride, channels);
        } else {                                                                               // :647: the best candidate since the last key frame
          write_node(Pose_id, min_stat, min_projs.data(), (int)min_projs.size() / 3);
          write_cov(cov_cams, min_camscov);
          if (!sel_img_.empty()) save_image(Pose_id, sel_img_.data(), sel_w_, sel_h_, sel_w_ * sel_c_, sel_c_);
        }
        set_last(stat14);                                                                      // :664-665
      } else if (frameId < 5) {                                                                // :668: first frames
        write_node(frameId, stat14.data(), nullptr, 0);
        min_camscov = slam.getSigma();
        write_cov(cov_cams, min_camscov);
        save_image(frameId, img, w, h, stride, channels);
        set_last(stat14);
      }
      min_cov_for_pose = 10000000;                                                             // :686
    }
    // monoslam_ransac.cpp:689-752 (the Point4sba.rows() >= Num_of_points_thershold candidate rule and the
    // take_image_every_x_frame rule) sits inside a comment block in the reference: dead code, not reproduced.
  }

  // ImageConverter::~ImageConverter (monoslam_ransac.cpp:262-275): closes the files and writes points.txt
  template <class Filter>
  void finish(Filter& slam) {
    node_proj.close(); cov_cams.close(); cov_cams2.close();
    int rows = 0;
    const std::vector<double> pts = slam.getPointsFeatures(&rows);
    f_points << eigen_format(pts.data(), rows, 12);
    f_points.close();
  }

 private:
  template <class Filter>
  void candidate(Filter& slam, int frameId, const std::vector<double>& stat14, double cov, const uint8_t* img, int w, int h, int stride,
                 int channels, bool with_cov) {                                                // :613-625
    min_cov_for_pose = cov;
    Pose_id = frameId;
    for (int i = 0; i < 7; ++i) min_stat[i] = stat14[i];
    if (with_cov) min_camscov = slam.getSigma();
    min_projs = slam.Point4sba;
    sel_w_ = w; sel_h_ = h; sel_c_ = channels;
    sel_img_.resize((size_t)w * h * channels);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w * channels; ++x) sel_img_[(size_t)y * w * channels + x] = img[(size_t)y * stride + x];
  }
  void set_last(const std::vector<double>& stat14) {
    quat2vec(stat14.data() + 3, last_vrot);
    for (int i = 0; i < 7; ++i) last_image_pose[i] = stat14[i];
  }
  // "P<id>", the 7 pose entries one per line (VectorXf <<), then the projections (MatrixX3i <<) or the literal
  // "0  0  0" the reference writes when it has none (:640, :672)
  void write_node(int id, const double* pose7, const int* projs, int nproj) {
    node_proj << "P" << id << std::endl;
    node_proj << eigen_format(pose7, 7, 1) << std::endl;
    if (projs) node_proj << eigen_format(projs, nproj, 3) << std::endl;
    else node_proj << "0  0  0" << std::endl;
    key_frames.push_back(id);
  }
  static void write_cov(std::ofstream& f, const std::vector<double>& S14) {                   // block<7,7>(0,0) of getSigma()
    double b[49];
    for (int i = 0; i < 7; ++i)
      for (int j = 0; j < 7; ++j) b[i * 7 + j] = S14.size() >= 196 ? S14[(size_t)i * 14 + j] : 0.0;
    f << eigen_format(b, 7, 7) << std::endl;
  }
  void save_image(int id, const uint8_t* img, int w, int h, int stride, int channels) {       // cv::imwrite(ToString(id) + ".png", ...)
    if (write_image_ && img) write_image_(dir_ + "/" + std::to_string(id) + ".png", img, w, h, stride, channels);
  }

  std::string dir_;
  ImageWriter write_image_;
  std::ofstream node_proj, cov_cams, cov_cams2, f_points;
  double last_vrot[3] = {0, 0, 0};                 // :196
  double last_image_pose[7] = {0, 0, 0, 0, 0, 0, 0};  // :198
  int Pose_id = 0;                                 // :202
  double min_stat[7] = {0, 0, 0, 0, 0, 0, 0};      // :203
  std::vector<double> min_camscov = std::vector<double>(196, 0.0);   // :204
  std::vector<int> min_projs;                      // :201 MatrixX3i::Zero(1,3)
  std::vector<uint8_t> sel_img_;                   // Selected_Pose
  int sel_w_ = 0, sel_h_ = 0, sel_c_ = 1;
};

}  // namespace ekf_b200
