// opencv2/core/core.hpp -- a stand-in for the handful of OpenCV core types the reference's vendored DBoW2
// (code/Thirdparty/DBoW2) touches, so that its UNMODIFIED sources compile here without OpenCV (test infrastructure
// only: oracle/Makefile target `ref`, output oracle/_ref/libdbow2_ref.so, used by tests/test_ref_dbow2.py to pin
// the oracle's restatement against the reference's own code).  cv::Mat here is a reference-counted row-major
// byte / float container with the members DBoW2 uses; FileStorage / FileNode exist only so that the YAML
// save()/load() members (virtual, hence instantiated) compile -- they are never called.
#pragma once
// the real header brings these in transitively and the DBoW2 sources rely on that
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;  // bytes per row (always dense here)
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * (type == CV_32F ? 4 : 1);
    const size_t bytes = (size_t)r * c * elemSize();
    buf_ = std::shared_ptr<unsigned char>(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
    data = buf_.get();
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    std::memset(m.data, 0, (size_t)r * c * m.elemSize());
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    if (data) std::memcpy(m.data, data, (size_t)rows * cols * elemSize());
    return m;
  }
  void release() { rows = cols = 0; data = nullptr; buf_.reset(); }
  size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  unsigned char* ptr(int r = 0) { return data + (size_t)r * cols * elemSize(); }
  const unsigned char* ptr(int r = 0) const { return data + (size_t)r * cols * elemSize(); }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * cols * elemSize()); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elemSize()); }
  Mat rowRange(int, int) const { return *this; }   // compile-only (never executed by the wrapper)
  Mat colRange(int, int) const { return *this; }
  Mat row(int) const { return *this; }
  int channels() const { return 1; }
  void copyTo(Mat& o) const { o = clone(); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }

 private:
  int type_ = CV_8U;
  std::shared_ptr<unsigned char> buf_;
};

// compile-only stand-ins (see the header comment)
class FileNode {
 public:
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator double() const { return 0; }
  operator float() const { return 0; }
  operator std::string() const { return std::string(); }
};

class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T>
inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv

// ---- additions for the reference's ORBextractor.cc (oracle/Makefile target `ref`): geometry types, cv::KeyPoint,
// Input/OutputArray proxies and the rounding helpers; the image-processing calls are declared in imgproc.hpp and
// defined as aborting stubs in oracle/ref_orbextractor_wrap.cpp (the wrapper only runs the constructor and
// DistributeOctTree, which need none of them).
namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  Point_& operator*=(T s) { x *= s; y *= s; return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  static Scalar all(double v) { return Scalar(v, v, v, v); }
};
struct Range {
  int start, end;
  Range(int s, int e) : start(s), end(e) {}
};

class KeyPoint {
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
  KeyPoint(Point2f p, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(p), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

template <typename T>
using Ptr = std::shared_ptr<T>;

namespace cuda { class GpuMat; }

// proxies: enough for signatures and for `image.getMat()` / `descriptors.create()` to compile
class _InputArray {
 public:
  _InputArray() : m_(nullptr), g_(nullptr) {}
  _InputArray(const Mat& m) : m_(&m), g_(nullptr) {}
  _InputArray(const cuda::GpuMat& g) : m_(nullptr), g_(&g) {}
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  const cuda::GpuMat& getGpuMatRef() const { return *g_; }
  bool empty() const { return !m_ || m_->empty(); }
  int type() const { return m_ ? m_->type() : 0; }
  const Mat* m_;
  const cuda::GpuMat* g_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() : out_(nullptr) {}
  _OutputArray(Mat& m) : _InputArray(m), out_(&m) {}
  _OutputArray(cuda::GpuMat& g) : _InputArray(g), out_(nullptr) {}
  void create(int r, int c, int type) const { if (out_) out_->create(r, c, type); }
  void release() const { if (out_) out_->release(); }
  Mat getMat() const { return out_ ? *out_ : Mat(); }
  Mat* out_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline InputArray noArray() { static _InputArray a; return a; }

inline int cvRound(double v) { return (int)std::lrint(v); }   // round half to even, like OpenCV's SSE2 cvRound
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
float fastAtan2(float y, float x);

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

}  // namespace cv
using cv::cvRound;
using cv::cvFloor;
using cv::cvCeil;
#define CV_Assert(expr) do { if (!(expr)) std::abort(); } while (0)
