// oracle/shim/opencv2/opencv.hpp — TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the few OpenCV types the reference's hot-path sources touch (cv::Mat as an 8-bit
// image with ROI views, clone(), at<uchar>(); Point/Size/Rect/Scalar), so that those sources compile
// unmodified into oracle/_ref (OpenCV's C++ headers are not installed in this image).  On the path
// itself the reference uses OpenCV only for exact byte copies and pixel reads.
// Also the double-matrix subset libblur.cpp needs (zeros, convertTo, at<double>, Mat / scalar, sum, norm,
// filter2D in its direct form).  Everything the path does not reach either does nothing (drawing,
// imshow) or aborts loudly (resize to another size, colour conversion, optical flow).
#ifndef EKF_SHIM_OPENCV_HPP_
#define EKF_SHIM_OPENCV_HPP_
#include "../shim_prelude.h"

typedef unsigned char uchar;

#define CV_8UC1 0
#define CV_8U 0
#define CV_8UC3 16
#define CV_64F 6
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

enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
template <class T> struct DataType;
template <> struct DataType<double> { enum { type = CV_64F }; };
template <> struct DataType<uchar> { enum { type = CV_8U }; };

// 8-bit image (1 or 3 channels) or double matrix, reference-counted buffer, ROI views share the buffer.
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
    type_ = type; rows = r; cols = c; step = (size_t)c * channels() * elemSize1();
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step + 8, 0);
    data = buf_->data();
  }
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  size_t elemSize1() const { return type_ == CV_64F ? 8 : 1; }
  int type() const { return type_; }
  int depth() const { return type_ == CV_64F ? CV_64F : CV_8U; }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return rows == 0 || cols == 0; }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * channels() * elemSize1());
    return m;
  }
  // convertTo between 8-bit and double (scale 1): u8 -> double exact; double -> u8 = saturate_cast<uchar>(cvRound(v)),
  // cvRound rounding half to even
  void convertTo(Mat& dst, int rtype) const {
    Mat out(rows, cols, rtype);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        const double v = type_ == CV_64F ? *reinterpret_cast<const double*>(data + (size_t)r * step + (size_t)c * 8)
                                          : (double)data[(size_t)r * step + c];
        if (rtype == CV_64F) *reinterpret_cast<double*>(out.data + (size_t)r * out.step + (size_t)c * 8) = v;
        else {
          const long iv = lrint(v);
          out.data[(size_t)r * out.step + c] = (uchar)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
        }
      }
    dst = out;
  }
  void copyTo(Mat& dst) const { dst = clone(); }
  // out-of-range accesses (evaluateKernel can index one past its kernel, libblur.cpp:39-43: undefined behaviour in the
  // reference) go to a sink instead of corrupting the heap
  template <class T> T& at(int r, int c) {
    static T sink; if (r < 0 || c < 0 || r >= rows || c >= cols) { sink = T(); return sink; }
    return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T));
  }
  template <class T> const T& at(int r, int c) const {
    static T sink = T(); if (r < 0 || c < 0 || r >= rows || c >= cols) return sink;
    return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T));
  }
  Mat& setTo(const Scalar& s) {
    for (int r = 0; r < rows; ++r) memset(data + (size_t)r * step, (int)s.val[0], (size_t)cols * channels());
    return *this;
  }
};

inline Mat operator/(const Mat& a, double s) {
  Mat out = a.clone();
  if (a.type() != CV_64F) shim_unsupported("Mat / scalar on a non-double matrix");
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) out.at<double>(r, c) = a.at<double>(r, c) / s;
  return out;
}
inline Scalar sum(const Mat& a) {
  if (a.type() != CV_64F) shim_unsupported("sum on a non-double matrix");
  double s = 0;
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) s += a.at<double>(r, c);
  return Scalar(s);
}
template <class T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
inline int shim_reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}
// filter2D, direct form (correlation, anchor at the kernel centre by default, BORDER_REFLECT_101), double only.
// Coefficients are visited in row-major order, zeros skipped, as OpenCV's non-DFT engine does.
inline void filter2D(const Mat& src, Mat& dst, int /*ddepth*/, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0,
                     int /*borderType*/ = BORDER_DEFAULT) {
  if (src.type() != CV_64F || kernel.type() != CV_64F) shim_unsupported("filter2D on non-double matrices");
  const int ax = anchor.x < 0 ? kernel.cols / 2 : anchor.x, ay = anchor.y < 0 ? kernel.rows / 2 : anchor.y;
  Mat in = src.clone();   // src and dst may be the same object (libblur.cpp:70)
  Mat out(src.rows, src.cols, CV_64F);
  for (int y = 0; y < src.rows; ++y)
    for (int x = 0; x < src.cols; ++x) {
      double s = delta;
      for (int ky = 0; ky < kernel.rows; ++ky)
        for (int kx = 0; kx < kernel.cols; ++kx) {
          const double kf = kernel.at<double>(ky, kx);
          if (kf == 0) continue;
          s += kf * in.at<double>(shim_reflect101(y + ky - ay, src.rows), shim_reflect101(x + kx - ax, src.cols));
        }
      out.at<double>(y, x) = s;
    }
  dst = out;
}
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
// cv::imwrite: nothing is written; a harness that defines ekf_shim_imwrite (weak) is told the file name
extern "C" void ekf_shim_imwrite(const char* name) __attribute__((weak));
template <class... A> bool imwrite(const std::string& name, A&&...) {
  if (ekf_shim_imwrite) ekf_shim_imwrite(name.c_str());
  return true;
}
inline int waitKey(int = 0) { return -1; }

}  // namespace cv
#include "../shim_retype.h"
#endif
