// opencv2/core/core.hpp -- stand-in for the OpenCV core types the reference's ORBmatcher.cc touches, so that the
// UNMODIFIED reference source compiles here without OpenCV (test infrastructure: oracle/Makefile target `ref`, output
// oracle/_ref/liborbmatcher_ref.so, used by tests/test_ref_orbmatcher.py to pin the oracle's matchers against the
// reference's own code).
//
// Unlike oracle/ref_shim (a byte container), cv::Mat here does real arithmetic on small CV_32F matrices, with lazy
// expressions, because ORBmatcher.cc projects MapPoints with them.  OpenCV is an un-vendored third-party dependency
// of the reference; the numerics below restate what OpenCV (3.4 / 4.x, baseline build) does for these shapes and are
// pinned against the installed cv2 4.13 in tests/test_ref_orbmatcher.py::test_cv_shim_numerics_match_cv2:
//   * A * B [+ C] with no transpose flag and inner dimension 2..4 equal to a side of the result: cv::gemm's
//     small-matrix path -- products and sums in FLOAT, left to right, then (float)(t * alpha + c * beta) in double;
//   * any product with a transposed operand (A.t() * B): the generic path, accumulation in DOUBLE, then
//     (float)(alpha * acc [+ beta * c]);
//   * Mat::dot and cv::norm (L2): accumulation in double;
//   * Mat * scalar, Mat / scalar: convertTo(alpha) = float product with (float)alpha  [not reachable from Python, so
//     this one follows OpenCV's cvtScale 32f->32f source and is not pinned];
//   * A + B, A - B, -A: element-wise float.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;

class KeyPoint {
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

class MatExpr;

// Row-major matrix view on a shared buffer (step in bytes), CV_8U or CV_32F.
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t ext_step) {  // header over caller-owned rows (Frame.cc:609,630)
    rows = r; cols = c; type_ = type; step = ext_step; data = static_cast<unsigned char*>(ext);
  }
  Mat(const MatExpr& e);
  Mat& operator=(const MatExpr& e);

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * elemSize();
    const size_t bytes = (size_t)r * step;
    buf_ = std::shared_ptr<unsigned char>(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
    data = buf_.get();
  }
  static Mat ones(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) m.at<float>(i, j) = 1.f;
    return m;
  }
  // 8U -> 32F (may be in place, as Frame.cc:610,631 call it): exact
  void convertTo(Mat& dst, int rtype) const {
    assert(rtype == CV_32F);
    Mat m(rows, cols, CV_32F);
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) m.at<float>(i, j) = type_ == CV_32F ? at<float>(i, j) : (float)at<unsigned char>(i, j);
    dst = m;
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    std::memset(m.data, 0, (size_t)r * m.step);
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; r++) std::memcpy(m.data + r * m.step, data + r * step, (size_t)cols * elemSize());
    return m;
  }
  size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  // single-index access to a vector (one row or one column), as OpenCV defines it
  template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }

  Mat sub(int r0, int r1, int c0, int c1) const {
    Mat m;
    m.rows = r1 - r0; m.cols = c1 - c0; m.step = step; m.type_ = type_; m.buf_ = buf_;
    m.data = data + (size_t)r0 * step + (size_t)c0 * elemSize();
    return m;
  }
  Mat rowRange(int a, int b) const { return sub(a, b, 0, cols); }
  Mat colRange(int a, int b) const { return sub(0, rows, a, b); }
  Mat row(int r) const { return sub(r, r + 1, 0, cols); }
  Mat col(int c) const { return sub(0, rows, c, c + 1); }

  MatExpr t() const;
  double dot(const Mat& o) const {  // accumulation in double (dotProd_32f scalar path)
    assert(type_ == CV_32F && rows == o.rows && cols == o.cols);
    double r = 0;
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) r += (double)at<float>(i, j) * o.at<float>(i, j);
    return r;
  }

 private:
  int type_ = CV_8U;
  std::shared_ptr<unsigned char> buf_;
};

// D = alpha * op(A) * B [+ beta * C]; see the header comment for which accumulation each shape gets.
inline Mat gemm_(const Mat& A, bool aT, const Mat& B, double alpha, const Mat* C, double beta) {
  const int M = aT ? A.cols : A.rows, K = aT ? A.rows : A.cols, N = B.cols;
  assert(A.type() == CV_32F && B.type() == CV_32F && B.rows == K);
  Mat D(M, N, CV_32F);
  const bool small = !aT && K >= 2 && K <= 4 && (K == N || K == M);
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N; j++) {
      const float c = C ? C->at<float>(i, j) : 0.f;
      if (small) {
        float t = A.at<float>(i, 0) * B.at<float>(0, j);
        for (int k = 1; k < K; k++) t = t + A.at<float>(i, k) * B.at<float>(k, j);
        D.at<float>(i, j) = (float)((double)t * alpha + (double)c * beta);
      } else {
        double acc = 0;
        for (int k = 0; k < K; k++) acc += (double)(aT ? A.at<float>(k, i) : A.at<float>(i, k)) * (double)B.at<float>(k, j);
        D.at<float>(i, j) = C ? (float)(alpha * acc + beta * (double)c) : (float)(alpha * acc);
      }
    }
  return D;
}

// Lazy expression: either alpha * op(a) (kind 0) or alpha * op(a) * b + beta * c (kind 1).
class MatExpr {
 public:
  int kind = 0;
  Mat a, b, c;
  bool aT = false, has_c = false;
  double alpha = 1, beta = 0;

  Mat eval() const {
    if (kind == 1) return gemm_(a, aT, b, alpha, has_c ? &c : nullptr, beta);
    Mat m(aT ? a.cols : a.rows, aT ? a.rows : a.cols, CV_32F);
    const float fa = (float)alpha;  // convertTo(alpha) on CV_32F works in float
    for (int i = 0; i < m.rows; i++)
      for (int j = 0; j < m.cols; j++) {
        const float v = aT ? a.at<float>(j, i) : a.at<float>(i, j);
        m.at<float>(i, j) = alpha == 1 ? v : v * fa;
      }
    return m;
  }
  template <typename T> T at(int i) const { return eval().at<T>(i); }
};

inline Mat::Mat(const MatExpr& e) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) { *this = e.eval(); return *this; }
inline MatExpr Mat::t() const { MatExpr e; e.a = *this; e.aT = true; return e; }

inline MatExpr scaled(const Mat& m, double s) { MatExpr e; e.a = m; e.alpha = s; return e; }
inline MatExpr operator-(const Mat& m) { return scaled(m, -1); }
inline MatExpr operator-(const MatExpr& x) { MatExpr e = x; e.alpha = -e.alpha; if (e.kind == 1) e.beta = -e.beta; return e; }
inline MatExpr operator*(double s, const Mat& m) { return scaled(m, s); }
inline MatExpr operator*(const Mat& m, double s) { return scaled(m, s); }
inline MatExpr operator/(const Mat& m, double s) { return scaled(m, 1.0 / s); }
inline MatExpr operator*(double s, const MatExpr& x) { MatExpr e = x; e.alpha *= s; if (e.kind == 1) e.beta *= s; return e; }

inline MatExpr mul_(const Mat& a, bool aT, double alpha, const Mat& b) {
  MatExpr e; e.kind = 1; e.a = a; e.aT = aT; e.alpha = alpha; e.b = b; return e;
}
inline MatExpr operator*(const Mat& a, const Mat& b) { return mul_(a, false, 1, b); }
inline MatExpr operator*(const MatExpr& x, const Mat& b) {
  if (x.kind == 0) return mul_(x.a, x.aT, x.alpha, b);
  return mul_(x.eval(), false, 1, b);
}
inline MatExpr operator+(const MatExpr& x, const Mat& c) {
  if (x.kind == 1 && !x.has_c) { MatExpr e = x; e.c = c; e.has_c = true; e.beta = 1; return e; }
  Mat l = x.eval();
  Mat r(l.rows, l.cols, CV_32F);
  for (int i = 0; i < l.rows; i++)
    for (int j = 0; j < l.cols; j++) r.at<float>(i, j) = l.at<float>(i, j) + c.at<float>(i, j);
  return scaled(r, 1);
}
inline Mat operator+(const Mat& l, const Mat& c) {
  Mat r(l.rows, l.cols, CV_32F);
  for (int i = 0; i < l.rows; i++)
    for (int j = 0; j < l.cols; j++) r.at<float>(i, j) = l.at<float>(i, j) + c.at<float>(i, j);
  return r;
}
inline Mat operator-(const Mat& l, const Mat& c) {
  Mat r(l.rows, l.cols, CV_32F);
  for (int i = 0; i < l.rows; i++)
    for (int j = 0; j < l.cols; j++) r.at<float>(i, j) = l.at<float>(i, j) - c.at<float>(i, j);
  return r;
}

inline Mat operator-(const Mat& l, const MatExpr& x) { return l - x.eval(); }  // A - s * B: addWeighted in float

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
inline double norm(const Mat& a, const Mat& b, int type) {  // NORM_L1 of a difference, accumulation in double
  assert(type == NORM_L1 && a.rows == b.rows && a.cols == b.cols);
  double s = 0;
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) s += std::abs((double)a.at<float>(i, j) - (double)b.at<float>(i, j));
  return s;
}

namespace cuda {
// a view on caller-owned image rows: what Frame::ComputeStereoMatches uses of cv::cuda::GpuMat (Frame.cc:608,629)
class GpuMat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;
  GpuMat() {}
  GpuMat(int r, int c, unsigned char* d, size_t s) : rows(r), cols(c), step(s), data(d) {}
  int type() const { return CV_8U; }
  GpuMat rowRange(int a, int b) const { return GpuMat(b - a, cols, data + (ptrdiff_t)a * (ptrdiff_t)step, step); }
  GpuMat colRange(int a, int b) const { return GpuMat(rows, b - a, data + a, step); }
};
}  // namespace cuda

inline double norm(const Mat& m) {  // NORM_L2, accumulation in double
  double s = 0;
  for (int i = 0; i < m.rows; i++)
    for (int j = 0; j < m.cols; j++) s += (double)m.at<float>(i, j) * (double)m.at<float>(i, j);
  return std::sqrt(s);
}

}  // namespace cv
