// cvcuda_shim.hpp -- stand-in for the OpenCV (core + CUDA device-layer) declarations that the reference's
// code/src/cuda/Fast_gpu.cu and Orb_gpu.cu use, so that nvcc compiles those two files UNMODIFIED here without
// OpenCV (test infrastructure: oracle/Makefile target `ref`, output oracle/_ref/libref_cuda.so, run on the GPU box
// by tests/test_gpu_ref_cuda.py to pin the oracle's FAST / IC_Angle / rBRIEF restatements against the reference's own
// kernels).  Everything here is plumbing (device buffers, views, a warp sum); none of it restates a kernel:
//   GpuMat        pitched device buffer with ROI views, setTo(0) and download
//   PtrStep[Sz]   the (y, x) device accessors the kernels index with
//   reduce<32>    OpenCV's block/warp reduction is used by IC_Angle_kernel only for INTEGER sums over the 32 lanes of
//                 one row, so any exact sum gives the same bits (here: shuffle-down adds; lane 0 holds the total)
#pragma once
#include <cuda_runtime.h>
#include <thrust/tuple.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_32SC1 4
#define CV_PI 3.1415926535897932384626433832795
#define CV_PI_F ((float)CV_PI)

typedef unsigned char uchar;
typedef unsigned int uint;

namespace cv {

struct Point { int x, y; };
struct Point2f {
  float x, y;
  __host__ __device__ Point2f() : x(0), y(0) {}
  __host__ __device__ Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Scalar {
  double val[4];
  static Scalar all(double v) { Scalar s; s.val[0] = s.val[1] = s.val[2] = s.val[3] = v; return s; }
};

class KeyPoint {  // 28 bytes, the layout the reference cudaMemcpy's to the device (Fast_gpu.cu:490, Orb_gpu.cu:123)
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class Mat {  // host byte matrix, enough for GpuMat::download
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;
  void create(int r, int c, int) {
    rows = r; cols = c; step = (size_t)c;
    buf_.reset(new uchar[(size_t)r * c + 1], std::default_delete<uchar[]>());
    data = buf_.get();
  }
 private:
  std::shared_ptr<uchar> buf_;
};

namespace cuda {

template <typename T> struct PtrStep {
  T* data;
  size_t step;
  __host__ __device__ T* ptr(int y = 0) { return (T*)((char*)data + y * step); }
  __host__ __device__ const T* ptr(int y = 0) const { return (const T*)((const char*)data + y * step); }
  __host__ __device__ T& operator()(int y, int x) { return ptr(y)[x]; }
  __host__ __device__ const T& operator()(int y, int x) const { return ptr(y)[x]; }
};
template <typename T> struct PtrStepSz : public PtrStep<T> {
  int cols, rows;
};
typedef PtrStep<uchar> PtrStepb;
typedef PtrStep<int> PtrStepi;
typedef PtrStepSz<uchar> PtrStepSzb;

class Stream {
 public:
  cudaStream_t s = 0;
  Stream() {}
  ~Stream() {}  // the reference calls this destructor explicitly and lets it run again
};
struct StreamAccessor {
  static Stream wrapStream(cudaStream_t s) { Stream r; r.s = s; return r; }
  static cudaStream_t getStream(const Stream& st) { return st.s; }
};

class GpuMat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;
  GpuMat() {}
  GpuMat(int r, int c, int type) { create(r, c, type); }
  GpuMat(Size s, int type) { create(s.height, s.width, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; esz_ = type == CV_32SC1 ? 4 : 1;
    void* p = nullptr;
    if (cudaMallocPitch(&p, &step, (size_t)c * esz_, r) != cudaSuccess) { std::fprintf(stderr, "cvcuda_shim: cudaMallocPitch failed\n"); std::abort(); }
    buf_.reset((uchar*)p, [](uchar* q) { cudaFree(q); });
    data = (uchar*)p;
  }
  bool empty() const { return data == nullptr; }
  Size size() const { return Size(cols, rows); }
  GpuMat rowRange(int a, int b) const { GpuMat m = *this; m.rows = b - a; m.data = data + (size_t)a * step; return m; }
  GpuMat colRange(int a, int b) const { GpuMat m = *this; m.cols = b - a; m.data = data + (size_t)a * esz_; return m; }
  GpuMat& setTo(Scalar v, Stream& st) {  // only ever called with 0
    if (v.val[0] != 0) std::abort();
    cudaMemset2DAsync(data, step, 0, (size_t)cols * esz_, rows, st.s);
    return *this;
  }
  void upload(const uchar* src, size_t sstep, cudaStream_t s = 0) {
    cudaMemcpy2DAsync(data, step, src, sstep, (size_t)cols * esz_, rows, cudaMemcpyHostToDevice, s);
  }
  void download(Mat& m, Stream& st) const {
    m.create(rows, cols * (int)esz_, CV_8UC1);
    cudaMemcpy2DAsync(m.data, m.step, data, step, (size_t)cols * esz_, rows, cudaMemcpyDeviceToHost, st.s);
  }
  template <typename T> operator PtrStepSz<T>() const { PtrStepSz<T> p; p.data = (T*)data; p.step = step; p.cols = cols; p.rows = rows; return p; }
  template <typename T> operator PtrStep<T>() const { PtrStep<T> p; p.data = (T*)data; p.step = step; return p; }
 private:
  size_t esz_ = 1;
  std::shared_ptr<uchar> buf_;
};

namespace device {
template <typename T> struct plus {
  __device__ __forceinline__ T operator()(const T& a, const T& b) const { return a + b; }
};
static __host__ __device__ __forceinline__ int divUp(int total, int grain) { return (total + grain - 1) / grain; }

template <int N, typename T, class Op>
__device__ __forceinline__ void reduce(volatile T*, T& val, unsigned int, const Op& op) {
  static_assert(N == 32, "the reference reduces over one 32-lane row");
  for (int d = 16; d > 0; d >>= 1) val = op(val, __shfl_down_sync(0xffffffffu, val, d));
}
template <typename P0, typename P1>
__device__ __forceinline__ thrust::tuple<P0*, P1*> smem_tuple(P0* a, P1* b) { return thrust::make_tuple(a, b); }
template <int N, typename P0, typename P1, typename R0, typename R1, class Op0, class Op1>
__device__ __forceinline__ void reduce(const thrust::tuple<P0, P1>&, const thrust::tuple<R0, R1>& val, unsigned int,
                                       const thrust::tuple<Op0, Op1>& op) {
  static_assert(N == 32, "the reference reduces over one 32-lane row");
  for (int d = 16; d > 0; d >>= 1) {
    thrust::get<0>(val) = thrust::get<0>(op)(thrust::get<0>(val), __shfl_down_sync(0xffffffffu, thrust::get<0>(val), d));
    thrust::get<1>(val) = thrust::get<1>(op)(thrust::get<1>(val), __shfl_down_sync(0xffffffffu, thrust::get<1>(val), d));
  }
}
}  // namespace device
}  // namespace cuda

class _InputArray {
 public:
  _InputArray(const cuda::GpuMat& g) : g_(g) {}
  cuda::GpuMat getGpuMat() const { return g_; }
 private:
  cuda::GpuMat g_;
};
typedef const _InputArray& InputArray;

}  // namespace cv
