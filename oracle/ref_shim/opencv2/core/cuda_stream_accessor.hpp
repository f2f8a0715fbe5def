#pragma once
#include <opencv2/core/cuda.hpp>
#include <cuda_runtime.h>
namespace cv { namespace cuda {
struct StreamAccessor {
  static cudaStream_t getStream(const Stream&) { return nullptr; }
  static Stream wrapStream(cudaStream_t) { return Stream(); }
};
} }
