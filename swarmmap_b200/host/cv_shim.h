// cv_shim.h -- the few OpenCV types the drop-in wrapper classes need, for builds WITHOUT OpenCV
// (the build and GPU images have no OpenCV C++ headers).  When the real OpenCV is available,
// compile with -DSWM_HAVE_OPENCV and this file is not used.  Layouts match OpenCV so the raw
// arrays cross the C ABI unchanged (cv::KeyPoint is 28 bytes: pt.x pt.y size angle response octave class_id).
#pragma once
#ifndef SWM_HAVE_OPENCV
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

struct Point2f {
  float x = 0, y = 0;
  Point2f() {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// Minimal row-major 8-bit matrix with shared ownership (enough for images and N x 32 descriptors).
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
  Mat(int r, int c, int /*type*/, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((uint8_t*)ext) {}
  void create(int r, int c, int /*type*/) {
    if (r == rows && c == cols && data) return;
    rows = r; cols = c; step = (size_t)c;
    buf_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * c + 1], std::default_delete<uint8_t[]>());
    data = buf_.get();
  }
  void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  bool isContinuous() const { return step == (size_t)cols; }
  uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.buf_ = buf_; return m; }

 private:
  std::shared_ptr<uint8_t> buf_;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif  // !SWM_HAVE_OPENCV
