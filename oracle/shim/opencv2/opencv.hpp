// oracle/shim/opencv2/opencv.hpp — TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the few OpenCV types the reference's hot-path sources touch (cv::Mat as an 8-bit
// image with ROI views, clone(), at<uchar>(); Point/Size/Rect/Scalar), so that those sources compile
// unmodified into oracle/_ref (OpenCV's C++ headers are not installed in this image).  On the path
// itself the reference uses OpenCV only for exact byte copies and pixel reads.
// Everything the path does not reach either does nothing (drawing, imshow) or aborts loudly
// (resize to another size, colour conversion, optical flow, filter2D): the harness never enables them.
#ifndef EKF_SHIM_OPENCV_HPP_
#define EKF_SHIM_OPENCV_HPP_
#include "../shim_prelude.h"

typedef unsigned char uchar;

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_BGR2GRAY 6
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

namespace cv {

[[noreturn]] inline void shim_unsupported(const char* what) {
  fprintf(stderr, "oracle/shim: cv::%s is outside the hot path and not provided\n", what);
  abort();
}

template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
#ifdef EKF_SHIM_DOUBLE
typedef Point_<double> Point2f;  // the reference re-typed to fp64
#else
typedef Point_<float> Point2f;
#endif
typedef Point_<int> Point2i;
typedef Point_<int> Point;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  template <class A, class B> Rect(A a, B b, int w, int h) : x((int)a), y((int)b), width(w), height(h) {}
};
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
struct KeyPoint { Point2f pt; };
struct TermCriteria {
  enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
  TermCriteria(int = 0, int = 0, double = 0) {}
};
struct NoArray {};
inline NoArray noArray() { return NoArray(); }
enum { FONT_HERSHEY_SIMPLEX = 0, FONT_HERSHEY_SCRIPT_SIMPLEX = 6 };

// 8-bit image, 1 or 3 channels, reference-counted buffer, ROI views share the buffer.
class Mat {
  std::shared_ptr<std::vector<uchar>> buf_;

 public:
  uchar* data = nullptr;
  int rows = 0, cols = 0;
  size_t step = 0;
  int type_ = CV_8UC1;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, void* ext, size_t stp) : data((uchar*)ext), rows(r), cols(c), step(stp), type_(type) {}
  Mat(const Mat& m, const Rect& roi) : buf_(m.buf_), rows(roi.height), cols(roi.width), step(m.step), type_(m.type_) {
    if (roi.x < 0 || roi.y < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows) {
      fprintf(stderr, "oracle/shim: cv::Mat ROI (%d,%d,%d,%d) outside %dx%d\n", roi.x, roi.y, roi.width, roi.height, m.cols, m.rows);
      abort();
    }
    data = m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.channels();
  }
  void create(int r, int c, int type) {
    type_ = type; rows = r; cols = c; step = (size_t)c * channels();
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step, 0);
    data = buf_->data();
  }
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return rows == 0 || cols == 0; }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * channels());
    return m;
  }
  void copyTo(Mat& dst) const { dst = clone(); }
  template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  Mat& setTo(const Scalar& s) {
    for (int r = 0; r < rows; ++r) memset(data + (size_t)r * step, (int)s.val[0], (size_t)cols * channels());
    return *this;
  }
};

inline void resize(const Mat& src, Mat& dst, Size sz) {
  if (sz.width != src.cols || sz.height != src.rows) shim_unsupported("resize to a different size (config.scale != 1)");
  Mat keep = src;  // src and dst may be the same object (vslamRansac.cpp:236)
  dst = keep;
}
inline void cvtColor(const Mat&, Mat&, int) { shim_unsupported("cvtColor"); }
template <class... A> void goodFeaturesToTrack(const Mat&, std::vector<Point2f>& out, A...) { out.clear(); }
template <class... A> void calcOpticalFlowPyrLK(A&&...) { shim_unsupported("calcOpticalFlowPyrLK"); }
template <class... A> void hconcat(A&&...) { shim_unsupported("hconcat"); }
template <class... A> void rectangle(A&&...) {}
template <class... A> void circle(A&&...) {}
template <class... A> void ellipse(A&&...) {}
template <class... A> void line(A&&...) {}
template <class... A> void putText(A&&...) {}
template <class... A> void imshow(A&&...) {}
template <class... A> bool imwrite(A&&...) { return true; }
inline int waitKey(int = 0) { return -1; }

}  // namespace cv
#include "../shim_retype.h"
#endif
