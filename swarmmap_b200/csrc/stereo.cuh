// stereo.cuh -- Frame::ComputeStereoMatches on the device, straight from the resident outputs of two extractors
// (left and right view).  SURVEY section 8(f) rank 3, the last matcher of the reference's front-end.
//
// Reference behaviour restated (paths relative to /root/reference/code/):
//   Frame::ComputeStereoMatches   src/Frame.cc:516-690
//     row band of a right keypoint [floor(y - r), ceil(y + r)], r = 2 * scale(octave)            :531-541
//     best descriptor distance among the right keypoints of the left keypoint's row, octave +-1,
//     uR in [uL - bf/b, uL + 3], strict <, candidates in ascending index                         :549-594
//     11 x 11 SAD over +-5 px in mvImagePyramid[level of the left keypoint], centre-subtracted    :597-647
//       (operator() blurs mvImagePyramid[l] in place, ORBextractor.cc:765,788, so inside the level these are the
//       BLURRED pixels; a window column or row outside the level would read the un-blurred border of the bordered
//       buffer the level is a view of -- px() below makes the same choice)
//     parabola fit, disparity / depth                                                            :649-672
//     1.5 * 1.4 * median SAD filter                                                              :675-689
//
// One warp per left keypoint.  The reference's per-row candidate table is replaced by a scan of all right keypoints
// with the same predicate (a few thousand per frame): the winner is the minimum of (distance << 16 | index), i.e. the
// first of the smallest distances in ascending index, which is what the strict `<` scan of the row list keeps.  The SAD
// values are integers; float expressions are evaluated operation by operation as the reference writes them
// (__fmul_rn / __fadd_rn / __fdiv_rn, no contraction), so mvuRight and mvDepth are bit-identical to the oracle's.
#pragma once

namespace swm {

struct StereoArgs {
  const FrameLayout* L;
  const swm_keypoint* kl; const uint8_t* dl; const int32_t* nl;  // left view: [B][cap_l]
  const swm_keypoint* kr; const uint8_t* dr; const int32_t* nr;  // right view: [B][cap_r]
  int cap_l, cap_r;
  const uint8_t* plain_l; const uint8_t* plain_r;                // un-blurred plane sets (border pixels)
  const uint8_t* blur_l; const uint8_t* blur_r;                  // blurred plane sets (what mvImagePyramid holds after operator())
  float sf[SWM_MAX_LEVELS], inv_sf[SWM_MAX_LEVELS];
  float mbf, mb;
  float* u_right; float* depth; int32_t* sad;                    // [B][cap_l]
};

// pixel (x, y) of mvImagePyramid[level] as Frame::ComputeStereoMatches sees it: blurred inside the level, the reflect-101
// border of the un-blurred buffer outside (both planes share one layout; `roi` = offset of level pixel (0,0))
__device__ __forceinline__ int stereo_px(const uint8_t* __restrict__ plain, const uint8_t* __restrict__ blur, long long roi, int pitch,
                                         int w, int h, int x, int y) {
  const uint8_t* p = ((unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h) ? blur : plain;
  return p[roi + (long long)y * pitch + x];
}

__global__ void __launch_bounds__(256) stereo_match_kernel(const StereoArgs a) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int nl = min(a.nl[f], a.cap_l), nr = min(a.nr[f], a.cap_r);
  if (iL >= nl) return;
  const FrameLayout* __restrict__ L = a.L;
  const long long o = (long long)f * a.cap_l + iL;
  const swm_keypoint kp = a.kl[o];
  const int levelL = kp.octave;
  const float uL = kp.x, vL = kp.y;
  const int nRows = L->lv[0].h;
  const int row = (int)vL;
  const float maxD = __fdiv_rn(a.mbf, a.mb);
  const float minU = __fsub_rn(uL, maxD), maxU = __fadd_rn(uL, 3.0f);  // uL - minD, minD = -3
  float out_u = -1.0f, out_z = -1.0f;
  int out_sad = -1;
  uint32_t best = 0xFFFFFFFFu;
  if (row >= 0 && row < nRows && !(maxU < 0)) {
    const uint4* dlp = reinterpret_cast<const uint4*>(a.dl + o * 32);
    const uint4 d0 = __ldg(dlp), d1 = __ldg(dlp + 1);
    const swm_keypoint* kr = a.kr + (long long)f * a.cap_r;
    const uint4* drp = reinterpret_cast<const uint4*>(a.dr + (long long)f * a.cap_r * 32);
    for (int j = lane; j < nr; j += 32) {
      const float ry = kr[j].y, rx = kr[j].x;
      const int ro = kr[j].octave;
      const float r = __fmul_rn(2.0f, a.sf[ro]);
      const int maxr = (int)ceilf(__fadd_rn(ry, r)), minr = (int)floorf(__fsub_rn(ry, r));
      if (row < minr || row > maxr || ro < levelL - 1 || ro > levelL + 1 || !(rx >= minU && rx <= maxU)) continue;
      const uint4 e0 = __ldg(drp + 2 * j), e1 = __ldg(drp + 2 * j + 1);
      const int dist = __popc(d0.x ^ e0.x) + __popc(d0.y ^ e0.y) + __popc(d0.z ^ e0.z) + __popc(d0.w ^ e0.w) +
                       __popc(d1.x ^ e1.x) + __popc(d1.y ^ e1.y) + __popc(d1.z ^ e1.z) + __popc(d1.w ^ e1.w);
      best = min(best, ((uint32_t)dist << 16) | (uint32_t)j);
    }
  }
  best = __reduce_min_sync(0xffffffffu, best);
  const int bestDist = (int)(best >> 16);
  if (best != 0xFFFFFFFFu && bestDist < 100) {  // ORBmatcher::TH_HIGH
    const int bestIdxR = (int)(best & 0xFFFFu);
    const float uR0 = a.kr[(long long)f * a.cap_r + bestIdxR].x;
    const float scaleFactor = a.inv_sf[levelL];
    const float scaleduL = roundf(__fmul_rn(uL, scaleFactor)), scaledvL = roundf(__fmul_rn(vL, scaleFactor));
    const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
    const LevelGeom& g = L->lv[levelL];
    // iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1 with L = w = 5
    if (!(scaleduR0 < 0 || __fadd_rn(scaleduR0, 11.0f) >= (float)g.w)) {
      const long long roi = (long long)f * L->slab_bytes + g.plane_off + (long long)kEdge * g.pitch + kPadX;
      const int y0 = (int)scaledvL - 5, xL0 = (int)scaleduL - 5, xRm = (int)scaleduR0 - 10;  // xRm: window of incR = -5
      const int cL = stereo_px(a.plain_l, a.blur_l, roi, g.pitch, g.w, g.h, xL0 + 5, y0 + 5);
      int cR[11], sad[11];
#pragma unroll
      for (int k = 0; k < 11; k++) {
        cR[k] = stereo_px(a.plain_r, a.blur_r, roi, g.pitch, g.w, g.h, xRm + 5 + k, y0 + 5);
        sad[k] = 0;
      }
      for (int pos = lane; pos < 121; pos += 32) {
        const int yy = pos / 11, xx = pos - yy * 11;
        const int l = stereo_px(a.plain_l, a.blur_l, roi, g.pitch, g.w, g.h, xL0 + xx, y0 + yy) - cL;
#pragma unroll
        for (int k = 0; k < 11; k++)
          sad[k] += abs(l - (stereo_px(a.plain_r, a.blur_r, roi, g.pitch, g.w, g.h, xRm + xx + k, y0 + yy) - cR[k]));
      }
#pragma unroll
      for (int k = 0; k < 11; k++) sad[k] = __reduce_add_sync(0xffffffffu, sad[k]);
      int bestSad = INT_MAX, bestk = 0;
#pragma unroll
      for (int k = 0; k < 11; k++)
        if (sad[k] < bestSad) {  // (float)dist < bestDist with integers: strict, the first minimum wins
          bestSad = sad[k];
          bestk = k;
        }
      if (bestk != 0 && bestk != 10) {
        float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
        for (int k = 1; k < 10; k++)
          if (k == bestk) {
            dist1 = (float)sad[k - 1];
            dist2 = (float)sad[k];
            dist3 = (float)sad[k + 1];
          }
        const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
        const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), den);
        if (!(deltaR < -1 || deltaR > 1)) {
          float bestuR = __fmul_rn(a.sf[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)(bestk - 5)), deltaR));
          float disparity = __fsub_rn(uL, bestuR);
          if (disparity >= 0 && disparity < maxD) {
            if (disparity <= 0) {
              disparity = 0.01f;
              bestuR = (float)((double)uL - 0.01);
            }
            out_z = __fdiv_rn(a.mbf, disparity);
            out_u = bestuR;
            out_sad = bestSad;
          }
        }
      }
    }
  }
  if (lane == 0) {
    a.u_right[o] = out_u;
    a.depth[o] = out_z;
    a.sad[o] = out_sad;
  }
}

// The 1.5 * 1.4 * median filter (Frame.cc:675-689): the median is element size/2 of the ascending SAD list -- found by
// bisection on the value (SADs are below 2^16) -- and every match whose SAD is not below the threshold is cleared.
__global__ void __launch_bounds__(1024) stereo_filter_kernel(const int32_t* __restrict__ nl, int cap, float* __restrict__ u_right,
                                                             float* __restrict__ depth, const int32_t* __restrict__ sad) {
  __shared__ int s_cnt;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(nl[f], cap);
  const long long o = (long long)f * cap;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < n; i += blockDim.x) local += sad[o + i] >= 0;
  if (local) atomicAdd(&s_cnt, local);
  __syncthreads();
  const int m = s_cnt;
  if (m == 0) return;  // (the reference would index an empty vector here)
  const int k = m / 2;
  int lo = 0, hi = 1 << 16;  // smallest v with count(sad <= v) >= k + 1
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    __syncthreads();
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    local = 0;
    for (int i = tid; i < n; i += blockDim.x) {
      const int s = sad[o + i];
      local += s >= 0 && s <= mid;
    }
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (s_cnt >= k + 1) hi = mid;
    else lo = mid + 1;
  }
  const float median = (float)lo;
  const float thDist = __fmul_rn(1.5f * 1.4f, median);
  for (int i = tid; i < n; i += blockDim.x) {
    const int s = sad[o + i];
    if (s >= 0 && !((float)s < thDist)) {
      u_right[o + i] = -1.0f;
      depth[o + i] = -1.0f;
    }
  }
}

}  // namespace swm
