#pragma once
// cv::cuda::GpuMat / Stream stand-ins: members of the reference's ORBextractor, never used by the wrapper
#include <opencv2/core/core.hpp>
namespace cv { namespace cuda {
class GpuMat {
 public:
  class Allocator {
   public:
    virtual ~Allocator() {}
    virtual bool allocate(GpuMat* mat, int rows, int cols, size_t elemSize) = 0;
    virtual void free(GpuMat* mat) = 0;
  };
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;
  int* refcount = nullptr;
  unsigned char* datastart = nullptr;
  unsigned char* dataend = nullptr;
  Allocator* allocator = nullptr;
  GpuMat() {}
  explicit GpuMat(Allocator* a) : allocator(a) {}
  GpuMat(int r, int c, int, Allocator* a = nullptr) : rows(r), cols(c), allocator(a) {}
  GpuMat(Size s, int, Allocator* a = nullptr) : rows(s.height), cols(s.width), allocator(a) {}
  GpuMat(const GpuMat& m, Rect) { *this = m; }
  GpuMat(int r, int c, int, void* d, size_t s = 0) : rows(r), cols(c), step(s), data((unsigned char*)d) {}
  explicit GpuMat(const Mat& m) : rows(m.rows), cols(m.cols) {}
  GpuMat rowRange(int, int) const { return *this; }
  GpuMat colRange(int, int) const { return *this; }
  GpuMat operator()(Rect) const { return *this; }
  void create(int r, int c, int) { rows = r; cols = c; }
  void create(Size s, int) { rows = s.height; cols = s.width; }
  void upload(InputArray) {}
  void upload(InputArray, class Stream&) {}
  void download(OutputArray) const {}
  void download(OutputArray, class Stream&) const {}
  GpuMat& setTo(Scalar) { return *this; }
  GpuMat& setTo(Scalar, class Stream&) { return *this; }
  void copyTo(OutputArray) const {}
  void copyTo(OutputArray, class Stream&) const {}
  bool empty() const { return data == nullptr; }
  Size size() const { return Size(cols, rows); }
  int type() const { return 0; }
  size_t elemSize() const { return 1; }
};
class Stream {
 public:
  Stream() {}
  void waitForCompletion() {}
  static Stream& Null() { static Stream s; return s; }
};
} }  // namespace cv::cuda
