// stand-in for the CUDA samples' helper_cuda.h: the reference only uses checkCudaErrors (Fast_gpu.cu:346-352 ...)
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define checkCudaErrors(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)
