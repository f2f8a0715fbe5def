// swm_core.cuh -- per-pixel arithmetic shared by the extractor kernels, written host/device
// neutral so tests/host_harness.cpp can check it against the oracle on a CPU-only box.
// (The product only instantiates the device side; there is no CPU execution path in libswm_orb.)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SWM_HD __host__ __device__ __forceinline__
#else
#define SWM_HD inline
#endif

namespace swm {

constexpr int kEdge = 19;       // EDGE_THRESHOLD, ORBextractor.cc:78
constexpr int kBand = 16;       // minBorderX = EDGE_THRESHOLD-3, ORBextractor.cc:695
constexpr int kHalfPatch = 15;  // ORBextractor.cc:77
constexpr int kPatch = 31;      // ORBextractor.cc:76
constexpr int kPadX = 32;       // left padding of every plane row so ROI x=0 is 32-byte aligned (>= kEdge)

// One entry of the bilinear resize tables (cv::resize INTER_LINEAR 8U fixed point: source offset
// and the two 11-bit weights).  Built on the host with IEEE double/float arithmetic exactly as
// OpenCV does, so the device never re-derives a rounding decision.
struct ResizeTap {
  int16_t ofs, a0, a1, pad;
};

SWM_HD int reflect101(int p, int n) {
  if (p < 0) p = -p;
  else if (p >= n) p = 2 * n - 2 - p;
  return p < 0 ? 0 : (p >= n ? n - 1 : p);
}

// Level l pixel (dx,dy) from the un-blurred level l-1 (ORBextractor.cc:845 with cv::resize
// semantics).  src points at ROI pixel (0,0) of level l-1.
SWM_HD int resize_fetch(const uint8_t* __restrict__ src, int spitch, int sw, int sh, ResizeTap tx, ResizeTap ty) {
  int sy0 = ty.ofs, sy1 = ty.ofs + 1;
  sy0 = sy0 < 0 ? 0 : (sy0 > sh - 1 ? sh - 1 : sy0);
  sy1 = sy1 < 0 ? 0 : (sy1 > sh - 1 ? sh - 1 : sy1);
  const int sx0 = tx.ofs;
  const int sx1 = sx0 + 1 > sw - 1 ? sw - 1 : sx0 + 1;
  const uint8_t* r0 = src + (long long)sy0 * spitch;
  const uint8_t* r1 = src + (long long)sy1 * spitch;
  const int h0 = r0[sx0] * tx.a0 + r0[sx1] * tx.a1;
  const int h1 = r1[sx0] * tx.a0 + r1[sx1] * tx.a1;
  return (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
}

// FAST-9/16 score at threshold th (Fast_gpu.cu:70-266): largest t >= th for which 9 contiguous
// ring pixels are all brighter than v+t or all darker than v-t; 0 if the pixel is not a corner at th.
// c points at the centre pixel inside an image/tile with row pitch `pitch`.
//
// Formulated on the two saturated difference rings b = max(r-v,0), k = max(v-r,0) with min-only
// sliding windows.  Do NOT rewrite this as "max(min(d), -max(d))": ptxas 12.9 for sm_100a
// miscompiles a negated max feeding VIMNMX3 (verified on a B200: wrong scores at -O3, right at
// -Xptxas -O0); the saturated form has no negation and was checked bit-exact on the device.
SWM_HD int fast_score(const uint8_t* c, int pitch, int th) {
  const int v = c[0];
  int r[16];
  // opposite pairs first: any 9-arc holds one pixel of each pair, so a pair inside the
  // +-th band rejects the pixel (exact; same early-outs as calcMask :82-125).
  const unsigned band = 2u * (unsigned)th;
  r[0] = c[3 * pitch];
  r[8] = c[-3 * pitch];
  if ((unsigned)(r[0] - v + th) <= band && (unsigned)(r[8] - v + th) <= band) return 0;
  r[4] = c[3];
  r[12] = c[-3];
  if ((unsigned)(r[4] - v + th) <= band && (unsigned)(r[12] - v + th) <= band) return 0;
  r[2] = c[2 * pitch + 2];
  r[10] = c[-2 * pitch - 2];
  if ((unsigned)(r[2] - v + th) <= band && (unsigned)(r[10] - v + th) <= band) return 0;
  r[6] = c[-2 * pitch + 2];
  r[14] = c[2 * pitch - 2];
  if ((unsigned)(r[6] - v + th) <= band && (unsigned)(r[14] - v + th) <= band) return 0;
  r[1] = c[3 * pitch + 1];
  r[3] = c[pitch + 3];
  r[5] = c[-pitch + 3];
  r[7] = c[-3 * pitch + 1];
  r[9] = c[-3 * pitch - 1];
  r[11] = c[-pitch - 3];
  r[13] = c[pitch - 3];
  r[15] = c[3 * pitch - 1];
  int b[16], k_[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    b[k] = r[k] > v ? r[k] - v : 0;
    k_[k] = v > r[k] ? v - r[k] : 0;
  }
  // sliding minimum over every 9-long circular window via doubling (2,4,8,+1)
  int b2[16], k2[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    b2[k] = b[k] < b[(k + 1) & 15] ? b[k] : b[(k + 1) & 15];
    k2[k] = k_[k] < k_[(k + 1) & 15] ? k_[k] : k_[(k + 1) & 15];
  }
  int b4[16], k4[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    b4[k] = b2[k] < b2[(k + 2) & 15] ? b2[k] : b2[(k + 2) & 15];
    k4[k] = k2[k] < k2[(k + 2) & 15] ? k2[k] : k2[(k + 2) & 15];
  }
  int best = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    int mb = b4[k] < b4[(k + 4) & 15] ? b4[k] : b4[(k + 4) & 15];
    mb = mb < b[(k + 8) & 15] ? mb : b[(k + 8) & 15];
    int mk = k4[k] < k4[(k + 4) & 15] ? k4[k] : k4[(k + 4) & 15];
    mk = mk < k_[(k + 8) & 15] ? mk : k_[(k + 8) & 15];
    best = best > mb ? best : mb;
    best = best > mk ? best : mk;
  }
  return best > th ? best - 1 : 0;
}

// 7-tap Gaussian, sigma 2, 8 fractional bits (cv::GaussianBlur 8U fixed-point kernel, see oracle).
SWM_HD int gauss7_tap(int k) {
  return k == 3 ? 56 : ((k == 2 || k == 4) ? 48 : ((k == 1 || k == 5) ? 34 : 18));
}

}  // namespace swm
