// ref_cuda_wrap.cu -- C entry points around the REFERENCE's own CUDA kernels: code/src/cuda/Fast_gpu.cu
// (tileCalcKeypoints_kernel, IC_Angle_kernel, addBorder_kernel and their host classes GpuFast / IC_Angle) and
// code/src/cuda/Orb_gpu.cu (calcOrb_kernel, GpuOrb), both #included below UNMODIFIED from where they lie under
// /root/reference and compiled by nvcc for sm_100a with the reference's own nvcc flag (-use_fast_math,
// CMakeLists.txt:34) against oracle/ref_shim_cuda's stand-ins for the OpenCV CUDA headers.
// -> oracle/_ref/libref_cuda.so (`make -C oracle ref`), run on the GPU box by tests/test_gpu_ref_cuda.py.
// Test infrastructure only; nothing in the product links or loads this.
#include <src/cuda/Fast_gpu.cu>
#include <src/cuda/Orb_gpu.cu>

#include <cstdint>

using namespace ORB_SLAM2::cuda;

namespace {
// The reference's per-pixel corner test + score (isKeyPoint2 -> cornerScore, Fast_gpu.cu:190-262) evaluated on every
// valid pixel of an image: the same device code tileCalcKeypoints_kernel runs, without its tile / retry logic.
__global__ void score_map_kernel(const cv::cuda::PtrStepSzb img, int threshold, cv::cuda::PtrStepi scoreMat) {
  const int j = threadIdx.x + blockIdx.x * blockDim.x + 3;
  const int i = threadIdx.y + blockIdx.y * blockDim.y + 3;
  if (i < img.rows - 3 && j < img.cols - 3) isKeyPoint2(img, i, j, threshold, scoreMat);
}

// tileCalcKeypoints_kernel's four phases (Fast_gpu.cu:296-340) run as separate launches, i.e. in LOCK-STEP over the
// whole image: every tile finishes a phase before any tile starts the next.  The per-pixel work is the reference's own
// device code (isKeyPoint2, isMax); only the phase boundaries are kernel boundaries here instead of __syncthreads(),
// which removes the inter-block race of the original launch and nothing else.  tile = the original kernel's block:
// pixels j in [3 + 32 bx, 35 + 32 bx), i in [3 + 32 by, 35 + 32 by).
__global__ void ls_score_kernel(const cv::cuda::PtrStepSzb img, int threshold, cv::cuda::PtrStepi scoreMat,
                                const uint8_t* has_kp, int tiles_x, uint8_t* is_kp) {
  const int j = threadIdx.x + blockIdx.x * blockDim.x + 3;
  const int i = threadIdx.y + blockIdx.y * blockDim.y + 3;
  if (!(i < img.rows - 3 && j < img.cols - 3)) return;
  if (has_kp && has_kp[((i - 3) / 32) * tiles_x + (j - 3) / 32]) return;  // `if (hasKp) return;` (:318)
  is_kp[(size_t)i * img.cols + j] = isKeyPoint2(img, i, j, threshold, scoreMat) ? 1 : 0;
}
__global__ void ls_nms_kernel(const cv::cuda::PtrStepSzb img, cv::cuda::PtrStepi scoreMat, const uint8_t* is_kp,
                              const uint8_t* skip_tiles, uint8_t* has_kp, int tiles_x, int32_t* out, int cap,
                              unsigned int* counter) {
  const int j = threadIdx.x + blockIdx.x * blockDim.x + 3;
  const int i = threadIdx.y + blockIdx.y * blockDim.y + 3;
  if (!(i < img.rows - 3 && j < img.cols - 3)) return;
  const int tile = ((i - 3) / 32) * tiles_x + (j - 3) / 32;
  if (skip_tiles && skip_tiles[tile]) return;
  if (!is_kp[(size_t)i * img.cols + j]) return;
  if (isMax(make_short2(j, i), scoreMat)) {
    if (has_kp) has_kp[tile] = 1;
    const unsigned int ind = atomicInc(counter, (unsigned int)(-1));
    if ((int)ind < cap) { out[3 * ind] = j; out[3 * ind + 1] = i; out[3 * ind + 2] = scoreMat(i, j); }
  }
}

cv::cuda::GpuMat upload(const uint8_t* img, int w, int h, int stride) {
  cv::cuda::GpuMat g(h, w, CV_8UC1);
  g.upload(img, (size_t)stride);
  cudaDeviceSynchronize();
  return g;
}
}  // namespace

extern "C" {

int refc_device_ok() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

// GpuFast::detect (Fast_gpu.cu:363-389) on one image (the extractor passes the level ROI [16, w-16) x [16, h-16),
// ORBextractor.cc:695-713).  out: n x (x, y, score) int32, in the kernel's atomicInc order (callers sort).
int refc_fast_detect(const uint8_t* img, int w, int h, int stride, int hi, int lo, int max_kp, int32_t* out, int cap) {
  cv::cuda::GpuMat g = upload(img, w, h, stride);
  GpuFast fast(hi, lo, max_kp);
  std::vector<cv::KeyPoint> kps;
  fast.detect(g, kps);
  const int n = (int)kps.size();
  for (int i = 0; i < n && i < cap; i++) {
    out[3 * i] = (int32_t)kps[i].pt.x; out[3 * i + 1] = (int32_t)kps[i].pt.y; out[3 * i + 2] = (int32_t)kps[i].response;
  }
  return n;
}

// The lock-step execution described above.  out: n x (x, y, score); has_kp_out (tiles_x * tiles_y, may be NULL)
// receives each tile's pass-1 outcome.
int refc_fast_detect_lockstep(const uint8_t* img, int w, int h, int stride, int hi, int lo, int32_t* out, int cap,
                              uint8_t* has_kp_out) {
  cv::cuda::GpuMat g = upload(img, w, h, stride);
  cv::cuda::GpuMat score(h, w, CV_32SC1);
  cv::cuda::Stream st;
  score.setTo(cv::Scalar::all(0), st);
  const int tiles_x = (w - 6 + 31) / 32, tiles_y = (h - 6 + 31) / 32;
  uint8_t *is_kp, *has_kp, *snap;
  int32_t* d_out;
  unsigned int* counter;
  cudaMalloc(&is_kp, (size_t)w * h); cudaMalloc(&has_kp, tiles_x * tiles_y); cudaMalloc(&snap, tiles_x * tiles_y);
  cudaMalloc(&d_out, (size_t)cap * 12); cudaMalloc(&counter, 4);
  cudaMemset(is_kp, 0, (size_t)w * h); cudaMemset(has_kp, 0, tiles_x * tiles_y); cudaMemset(counter, 0, 4);
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  ls_score_kernel<<<grid, block>>>(g, hi, score, nullptr, tiles_x, is_kp);
  ls_nms_kernel<<<grid, block>>>(g, score, is_kp, nullptr, has_kp, tiles_x, d_out, cap, counter);
  cudaMemcpy(snap, has_kp, tiles_x * tiles_y, cudaMemcpyDeviceToDevice);  // pass-1 outcome, frozen
  cudaMemset(is_kp, 0, (size_t)w * h);
  ls_score_kernel<<<grid, block>>>(g, lo, score, snap, tiles_x, is_kp);
  ls_nms_kernel<<<grid, block>>>(g, score, is_kp, snap, nullptr, tiles_x, d_out, cap, counter);
  unsigned int n = 0;
  cudaMemcpy(&n, counter, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(out, d_out, (size_t)(n < (unsigned)cap ? n : cap) * 12, cudaMemcpyDeviceToHost);
  if (has_kp_out) cudaMemcpy(has_kp_out, snap, tiles_x * tiles_y, cudaMemcpyDeviceToHost);
  cudaFree(is_kp); cudaFree(has_kp); cudaFree(snap); cudaFree(d_out); cudaFree(counter);
  return (int)n;
}

// score map at one threshold: 0 where isKeyPoint2 says "not a corner", cornerScore otherwise
void refc_fast_score_map(const uint8_t* img, int w, int h, int stride, int threshold, int32_t* out) {
  cv::cuda::GpuMat g = upload(img, w, h, stride);
  cv::cuda::GpuMat score(h, w, CV_32SC1);
  cv::cuda::Stream st;
  score.setTo(cv::Scalar::all(0), st);
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  score_map_kernel<<<grid, block>>>(g, threshold, score);
  cudaMemcpy2D(out, (size_t)w * 4, score.data, score.step, (size_t)w * 4, h, cudaMemcpyDeviceToHost);
}

// IC_Angle::launch_async + join (Fast_gpu.cu:476-509): kps (cv::KeyPoint layout, ROI coordinates) get pt += border,
// octave, size and the orientation in degrees.  img = the whole level image (no border needed beyond the patch).
void refc_ic_angle(const uint8_t* img, int w, int h, int stride, void* kps, int n, int half_k, int min_border_x,
                   int min_border_y, int octave, int size, const int* umax, int n_umax) {
  cv::cuda::GpuMat g = upload(img, w, h, stride);
  IC_Angle::loadUMax(umax, n_umax);
  IC_Angle ic(n > 0 ? n : 1);
  ic.launch_async(g, (cv::KeyPoint*)kps, n, half_k, min_border_x, min_border_y, octave, size);
  if (n > 0) ic.join((cv::KeyPoint*)kps, n);
}

// GpuOrb::launch_async + join (Orb_gpu.cu:102-136): kps in level coordinates with .angle in degrees; img = the
// blurred level image; pattern = 512 cv::Point (ORBextractor.cc:380-382).  desc: n x 32.
void refc_orb(const uint8_t* img, int w, int h, int stride, const void* kps, int n, const int32_t* pattern_xy,
              uint8_t* desc) {
  cv::cuda::GpuMat g = upload(img, w, h, stride);
  GpuOrb::loadPattern((const cv::Point*)pattern_xy);
  GpuOrb orb(n > 0 ? n : 1);
  orb.launch_async(g, (const cv::KeyPoint*)kps, n);
  if (n == 0) return;
  cv::Mat d;
  orb.join(d);
  for (int i = 0; i < n; i++) std::memcpy(desc + 32 * (size_t)i, d.data + (size_t)i * d.step, 32);
}

}  // extern "C"
