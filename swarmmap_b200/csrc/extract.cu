// extract.cu -- B200 (sm_100a) ORB extractor: fused pyramid/border/FAST-score/blur level kernels,
// tile-retry non-max suppression, block-parallel quadtree, orientation + rBRIEF, and the C ABI of
// the extractor (include/swm_orb.h).  Batched over frames: every launch covers all frames of a
// batch, nothing returns to the host between stages.
//
// Reference behaviour restated (paths relative to /root/reference/code/):
//   ORBextractor::operator()            src/ORBextractor.cc:746-819
//   ORBextractor::ComputePyramid        src/ORBextractor.cc:821-855   (cv::resize INTER_LINEAR 8U semantics)
//   ComputeKeyPointsOctTree             src/ORBextractor.cc:691-744
//   tileCalcKeypoints_kernel            src/cuda/Fast_gpu.cu:284-341  (lock-step deterministic form)
//   DistributeOctTree                   src/ORBextractor.cc:465-689   (octree_core.cuh)
//   IC_Angle_kernel / addBorder_kernel  src/cuda/Fast_gpu.cu:403-471
//   Gaussian 7x7 sigma 2                src/ORBextractor.cc:835,719,742 (cv::GaussianBlur 8U semantics)
//   calcOrb_kernel                      src/cuda/Orb_gpu.cu:67-100
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "octree_core.cuh"
#include "orb_pattern.inc"
#include "swm_internal.cuh"

namespace swm {

__constant__ signed char c_pattern[1024];
__constant__ int c_umax[16];

// ------------------------------------------------------------------------------------------------
// Level kernel: one CTA produces a 64x32 tile of level l (plus the mirrored border pixels that
// reflect into it), its FAST score map and its blurred tile, from a shared-memory staging tile
// with a 3-pixel halo.  Level 0 copies the input frame, levels >= 1 resize the un-blurred level l-1.
// ------------------------------------------------------------------------------------------------
constexpr int TW = 64, TH = 32, HALO = 3;
constexpr int SW = TW + 2 * HALO, SH = TH + 2 * HALO, SP = 72;

struct LevelArgs {
  LevelGeom dst, src;
  const uint8_t* img;
  int img_stride;
  long long img_frame_stride;
  uint8_t* plain;
  uint8_t* blur;
  uint8_t* score;
  long long slab_bytes;
  const ResizeTap* xtab;
  const ResizeTap* ytab;
  int min_th;
};

__device__ __forceinline__ void store4(uint8_t* row, int gx, int w, uint32_t word) {
  if (gx + 3 < w) {
    *reinterpret_cast<uint32_t*>(row + gx) = word;
  } else {
    for (int k = 0; k < 4; k++)
      if (gx + k < w) row[gx + k] = (uint8_t)(word >> (8 * k));
  }
}

template <bool kFirst>
__global__ void __launch_bounds__(256) level_kernel(const LevelArgs a) {
  __shared__ __align__(16) uint8_t s_px[SH * SP];
  __shared__ __align__(16) uint16_t s_h[SH * TW];
  const int tid = threadIdx.x;
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int w = a.dst.w, h = a.dst.h;
  const long long fo = (long long)f * a.slab_bytes;

  const uint8_t* src;
  int spitch;
  if (kFirst) {
    src = a.img + (long long)f * a.img_frame_stride;
    spitch = a.img_stride;
  } else {
    src = a.plain + fo + a.src.plane_off + (long long)kEdge * a.src.pitch + kPadX;
    spitch = a.src.pitch;
  }
  const ResizeTap* __restrict__ xt = a.xtab + a.dst.xtab_off;
  const ResizeTap* __restrict__ yt = a.ytab + a.dst.ytab_off;

  // ---- stage the un-blurred tile (+halo, reflect-101 at the image edge) in shared memory
  for (int i = tid; i < SH * SW; i += 256) {
    const int ly = i / SW, lx = i - ly * SW;
    const int gx = reflect101(x0 - HALO + lx, w), gy = reflect101(y0 - HALO + ly, h);
    int v;
    if (kFirst) v = __ldg(src + (long long)gy * spitch + gx);
    else v = resize_fetch(src, spitch, a.src.w, a.src.h, xt[gx], yt[gy]);
    s_px[ly * SP + lx] = (uint8_t)v;
  }
  __syncthreads();

  const long long roi0 = fo + a.dst.plane_off + (long long)kEdge * a.dst.pitch + kPadX;
  uint8_t* dplain = a.plain + roi0;
  uint8_t* dscore = a.score + roi0;
  uint8_t* dblur = a.blur + roi0;
  const int pitch = a.dst.pitch;
  const bool edge_tile = (x0 <= kEdge) || (x0 + TW >= w - 1 - kEdge) || (y0 <= kEdge) || (y0 + TH >= h - 1 - kEdge);

  // ---- horizontal blur pass for all staged rows (row pass of cv::GaussianBlur, 8.8 fixed point)
  for (int i = tid; i < SH * TW; i += 256) {
    const int r = i / TW, c = i - r * TW;
    const uint8_t* p = s_px + r * SP + c;
    const int acc = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
    s_h[i] = (uint16_t)acc;
  }

  // ---- un-blurred tile, border mirrors, FAST score
  const int c4 = (tid & 15) * 4;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int r = (tid >> 4) + 16 * pass;
    const int gy = y0 + r, gx = x0 + c4;
    if (gy < h && gx < w) {
      const uint8_t* p = s_px + (r + HALO) * SP + c4 + HALO;
      const uint32_t word = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
      store4(dplain + (long long)gy * pitch, gx, w, word);
      if (edge_tile) {
        // copyMakeBorder(BORDER_REFLECT_101, 19): every in-image pixel also lands at the border
        // positions that reflect onto it (ORBextractor.cc:846-851).
        int ys[3], ny = 0;
        ys[ny++] = gy;
        if (gy >= 1 && gy <= kEdge) ys[ny++] = -gy;
        if (gy >= h - 1 - kEdge && gy <= h - 2) ys[ny++] = 2 * (h - 1) - gy;
        for (int k = 0; k < 4; k++) {
          const int x = gx + k;
          if (x >= w) break;
          int xs[3], nx = 0;
          xs[nx++] = x;
          if (x >= 1 && x <= kEdge) xs[nx++] = -x;
          if (x >= w - 1 - kEdge && x <= w - 2) xs[nx++] = 2 * (w - 1) - x;
          for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++)
              if (ix | iy) dplain[(long long)ys[iy] * pitch + xs[ix]] = p[k];
        }
      }
      // FAST score at minThFAST on the band [19, w-19) x [19, h-19) (ORBextractor.cc:695-713 + 3 px ring)
      uint32_t sword = 0;
      if (gy >= kEdge && gy < h - kEdge) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int x = gx + k;
          if (x >= kEdge && x < w - kEdge) sword |= (uint32_t)fast_score(p + k, SP, a.min_th) << (8 * k);
        }
      }
      store4(dscore + (long long)gy * pitch, gx, w, sword);
    }
  }
  __syncthreads();

  // ---- vertical blur pass (16.16 fixed point, round half up)
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int r = (tid >> 4) + 16 * pass;
    const int gy = y0 + r, gx = x0 + c4;
    if (gy < h && gx < w) {
      uint32_t word = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint16_t* q = s_h + r * TW + c4 + k;
        const uint32_t acc = 18u * (q[0] + q[6 * TW]) + 34u * (q[TW] + q[5 * TW]) + 48u * (q[2 * TW] + q[4 * TW]) +
                             56u * q[3 * TW];
        word |= ((acc + 0x8000u) >> 16) << (8 * k);
      }
      store4(dblur + (long long)gy * pitch, gx, w, word);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile retry + non-max suppression (Fast_gpu.cu:284-341), deterministic lock-step form:
//   pass 1: keypoint iff S_hi(p) > S_hi(q) for the 8 neighbours; a tile without any pass-1
//           keypoint is flagged for retry;
//   pass 2 (retry tiles only): keypoint iff S(p) > S_eff(q), S_eff = S in retried tiles, S_hi elsewhere.
// Survivors are appended (unordered) to the per-(frame, level) candidate list; everything downstream
// orders by explicit keys, so the append order does not matter.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nms_kernel(const FrameLayout* __restrict__ L, const uint8_t* __restrict__ score,
                                                  int ini_th, int pass, uint8_t* __restrict__ retry,
                                                  uint32_t* __restrict__ cand, int* __restrict__ cand_count) {
  __shared__ uint8_t s[34][36];
  __shared__ int s_any;
  __shared__ uint8_t s_flag[9];
  const int tid = threadIdx.x;
  const int f = blockIdx.y;
  const int t = blockIdx.x;
  int lvl = 0;
  while (lvl + 1 < L->nlevels && t >= L->lv[lvl + 1].tile_off) lvl++;
  const LevelGeom& g = L->lv[lvl];
  uint8_t* fretry = retry + (long long)f * L->tiles_total;
  if (pass == 2 && !fretry[t]) return;
  const int tl = t - g.tile_off;
  const int by = tl / g.tiles_x, bx = tl - by * g.tiles_x;
  if (tid == 0) s_any = 0;
  if (pass == 2 && tid < 9) {
    const int ny = by + tid / 3 - 1, nx = bx + tid % 3 - 1;
    s_flag[tid] = (ny >= 0 && ny < g.tiles_y && nx >= 0 && nx < g.tiles_x) ? fretry[g.tile_off + ny * g.tiles_x + nx] : 0;
  }
  __syncthreads();
  const uint8_t* sc = score + (long long)f * L->slab_bytes + g.plane_off + (long long)kEdge * g.pitch + kPadX;
  const int ox = kEdge + 32 * bx - 1, oy = kEdge + 32 * by - 1;  // level coords of s[0][0]
  for (int i = tid; i < 34 * 34; i += 256) {
    const int ly = i / 34, lx = i - ly * 34;
    const int gx = ox + lx, gy = oy + ly;
    int v = (gx < g.w && gy < g.h) ? sc[(long long)gy * g.pitch + gx] : 0;
    bool raw = false;
    if (pass == 2) {
      const int fy = ly == 0 ? 0 : (ly == 33 ? 2 : 1), fx = lx == 0 ? 0 : (lx == 33 ? 2 : 1);
      raw = s_flag[fy * 3 + fx] != 0;
    }
    if (!raw && v < ini_th) v = 0;
    s[ly][lx] = (uint8_t)v;
  }
  __syncthreads();
  const int lx = (tid & 31) + 1;
  const int lane = tid & 31;
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ly = (tid >> 5) * 4 + k + 1;
    const int c = s[ly][lx];
    bool kp = c > 0;
    if (kp) {
      kp = c > s[ly - 1][lx - 1] && c > s[ly - 1][lx] && c > s[ly - 1][lx + 1] && c > s[ly][lx - 1] &&
           c > s[ly][lx + 1] && c > s[ly + 1][lx - 1] && c > s[ly + 1][lx] && c > s[ly + 1][lx + 1];
    }
    const unsigned m = __ballot_sync(0xffffffffu, kp);
    if (m) {
      any = true;
      int base = 0;
      if (lane == 0) base = atomicAdd(cand_count + f * L->nlevels + lvl, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (kp) {
        const int idx = base + __popc(m & ((1u << lane) - 1));
        if (idx < g.cand_cap) {
          const int gx = ox + lx, gy = oy + ly;
          cand[(long long)f * L->cand_total + g.cand_off + idx] = pack_pt(gx - kBand, gy - kBand, c);
        }
      }
    }
  }
  if (pass == 1) {
    if (any && lane == 0) s_any = 1;
    __syncthreads();
    if (tid == 0) fretry[t] = s_any ? 0 : 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Quadtree: one CTA per (level, frame).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) octree_kernel(const FrameLayout* __restrict__ L, const uint32_t* __restrict__ cand,
                                                     const int* __restrict__ cand_count, int max_pts,
                                                     uint32_t* __restrict__ sel, int* __restrict__ sel_count) {
  extern __shared__ __align__(16) uint8_t smem[];
  OtState& S = *reinterpret_cast<OtState*>(smem);
  uint32_t* pts = reinterpret_cast<uint32_t*>(smem + sizeof(OtState));
  uint16_t* pnode = reinterpret_cast<uint16_t*>(pts + max_pts);
  uint8_t* pchild = reinterpret_cast<uint8_t*>(pnode + max_pts);
  const int lvl = blockIdx.x, f = blockIdx.y;
  const LevelGeom& g = L->lv[lvl];
  const int tid = threadIdx.x;
  int n_all = cand_count[f * L->nlevels + lvl];
  if (n_all > g.cand_cap) n_all = g.cand_cap;
  const uint32_t* src = cand + (long long)f * L->cand_total + g.cand_off;
  int n;
  if (n_all <= max_pts) {
    for (int i = tid; i < n_all; i += blockDim.x) pts[i] = src[i];
    n = n_all;
  } else {
    // More survivors than the reference's buffer (Fast.hpp:30): keep the first max_pts in raster
    // order = the max_pts smallest packed words.  Bisect on the value.
    int* cnt = &S.scalars[OT_TMP0];
    uint32_t lo = 0, hi = 0xFFFFFFFFu;
    for (int it = 0; it < 32; it++) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (tid == 0) *cnt = 0;
      __syncthreads();
      int local = 0;
      for (int i = tid; i < n_all; i += blockDim.x) local += src[i] <= mid;
      if (local) atomicAdd(cnt, local);
      __syncthreads();
      if (*cnt >= max_pts) hi = mid;
      else lo = mid + 1;
      __syncthreads();
    }
    if (tid == 0) *cnt = 0;
    __syncthreads();
    for (int i = tid; i < n_all; i += blockDim.x) {
      const uint32_t v = src[i];
      if (v <= lo) {
        const int k = atomicAdd(cnt, 1);
        if (k < max_pts) pts[k] = v;
      }
    }
    n = max_pts;
  }
  __syncthreads();
  const int k = ot_distribute(S, pts, pnode, pchild, n, g.w - 2 * kBand, g.h - 2 * kBand, g.quota,
                              sel + (long long)f * L->sel_total + g.sel_off, g.sel_cap);
  if (tid == 0) sel_count[f * L->nlevels + lvl] = k;
}

// ------------------------------------------------------------------------------------------------
// Orientation (un-blurred plane) + rBRIEF (blurred plane) + keypoint record, one warp per keypoint.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) describe_kernel(const FrameLayout* __restrict__ L, const uint8_t* __restrict__ plain,
                                                       const uint8_t* __restrict__ blur, const uint32_t* __restrict__ sel,
                                                       const int* __restrict__ sel_count, swm_keypoint* __restrict__ kps,
                                                       uint8_t* __restrict__ desc, int cap, int32_t* __restrict__ n_out) {
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gidx = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int nl = L->nlevels;
  int total = 0, lvl = -1, in_level = 0;
  for (int l = 0; l < nl; l++) {
    const int c = sel_count[f * nl + l];
    if (lvl < 0 && gidx < total + c) {
      lvl = l;
      in_level = gidx - total;
    }
    total += c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) n_out[f] = total < cap ? total : cap;
  if (lvl < 0 || gidx >= cap) return;
  const LevelGeom& g = L->lv[lvl];
  const uint32_t p = sel[(long long)f * L->sel_total + g.sel_off + in_level];
  const int x = pt_x(p) + kBand, y = pt_y(p) + kBand;  // addBorder_kernel, Fast_gpu.cu:462-471
  const long long off = (long long)f * L->slab_bytes + g.plane_off + (long long)(y + kEdge) * g.pitch + kPadX + x;
  const int pitch = g.pitch;

  // IC_Angle (Fast_gpu.cu:403-460): integer moments over the radius-15 disc, one lane per column
  const uint8_t* c = plain + off;
  int m10 = 0, m01 = 0;
  if (lane <= 2 * kHalfPatch) m10 = (lane - kHalfPatch) * (int)c[lane - kHalfPatch];
  for (int v = 1; v <= kHalfPatch; ++v) {
    const int d = c_umax[v];
    if (lane <= 2 * d) {
      const int u = lane - d;
      const int vp = c[u + v * pitch], vm = c[u - v * pitch];
      m01 += v * (vp - vm);
      m10 += u * (vp + vm);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float kPiF = 3.14159265f;  // CV_PI_F
  float ang = atan2f((float)m01, (float)m10);
  if (ang < 0) ang = __fadd_rn(ang, 2.0f * kPiF);
  ang = __fmul_rn(ang, 180.0f / kPiF);

  // calcOrb_kernel (Orb_gpu.cu:67-100): lane t -> descriptor byte t
  const float factor_pi = (float)(3.1415926535897932384626433832795 / 180.f);
  const float rad = __fmul_rn(ang, factor_pi);
  const float ca = cosf(rad), sb = sinf(rad);
  const uint8_t* b = blur + off;
  const signed char* pat = c_pattern + 32 * lane;
  int val = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
    const int t0 = b[__float2int_rn(__fadd_rn(__fmul_rn(x0, sb), __fmul_rn(y0, ca))) * pitch +
                     __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sb)))];
    const int t1 = b[__float2int_rn(__fadd_rn(__fmul_rn(x1, sb), __fmul_rn(y1, ca))) * pitch +
                     __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sb)))];
    val |= (t0 < t1) << k;
  }
  const long long o = (long long)f * cap + gidx;
  desc[o * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    swm_keypoint kp;
    kp.x = lvl ? __fmul_rn((float)x, g.scale) : (float)x;  // ORBextractor.cc:808-814
    kp.y = lvl ? __fmul_rn((float)y, g.scale) : (float)y;
    kp.size = (float)g.kp_size;
    kp.angle = ang;
    kp.response = (float)pt_score(p);
    kp.octave = lvl;
    kp.class_id = -1;
    kps[o] = kp;
  }
}

}  // namespace swm

// =================================================================================================
// Host side: handle, tables, C ABI
// =================================================================================================
using namespace swm;

namespace {

thread_local std::string g_create_error;
std::once_flag g_const_once[64];

#define SWM_CK(h, call)                                                     \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      (h)->err = cuda_err(#call, e_);                                       \
      return SWM_E_CUDA;                                                    \
    }                                                                       \
  } while (0)

inline int cv_round_f(float v) { return (int)lrintf(v); }
inline int cv_floor_f(float v) {
  int i = (int)v;
  return i - (i > v);
}
inline short sat_s16(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }
inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

}  // namespace

int swm::check_device(int device, std::string* err) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    *err = "no CUDA device visible (" + std::string(cudaGetErrorString(e)) +
           "); libswm_orb has no CPU fallback";
    return SWM_E_NODEVICE;
  }
  if (device < 0 || device >= count) {
    *err = "device index out of range";
    return SWM_E_INVALID;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    *err = cuda_err("cudaGetDeviceProperties", e);
    return SWM_E_CUDA;
  }
  if (prop.major != 10) {
    *err = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
           "; libswm_orb is built for sm_100a only";
    return SWM_E_NODEVICE;
  }
  return SWM_OK;
}

struct swm_orb {
  swm_orb_cfg cfg;
  int device = 0;
  std::string err;
  cudaStream_t stream = nullptr;
  // tables (ORBextractor.cc:346-378)
  float sf[SWM_MAX_LEVELS], inv_sf[SWM_MAX_LEVELS], sig2[SWM_MAX_LEVELS], inv_sig2[SWM_MAX_LEVELS];
  int quota[SWM_MAX_LEVELS];
  int max_kp = 0;
  int max_pts = 10000;
  // geometry for the current frame size
  FrameLayout lay;
  bool allocated = false;
  FrameLayout* d_lay = nullptr;
  ResizeTap *d_xtab = nullptr, *d_ytab = nullptr;
  uint8_t *d_plain = nullptr, *d_blur = nullptr, *d_score = nullptr;
  uint8_t* d_retry = nullptr;
  uint32_t *d_cand = nullptr, *d_sel = nullptr;
  int *d_counts = nullptr;  // [2][B][nlevels]: candidate counts, selection counts
  // staging for the host-buffer entry points
  uint8_t* d_img = nullptr;
  int img_pitch = 0;
  swm_keypoint* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int32_t* d_n = nullptr;
  int last_batch = 0;
  int last_launches = 0;
  // last input (for swm_orb_run_stage)
  const uint8_t* last_img = nullptr;
  int last_stride = 0;
  long long last_frame_stride = 0;
  swm_keypoint* last_kps = nullptr;
  uint8_t* last_desc = nullptr;
  int32_t* last_n = nullptr;
  int last_cap = 0;
  // async host-buffer call in flight (swm_orb_extract_batch_async .. swm_orb_sync)
  int32_t* pending_n = nullptr;
  int pending_batch = 0;
  int pending_cap = 0;
};

namespace {

void free_frame_buffers(swm_orb* h) {
  cudaFree(h->d_lay); cudaFree(h->d_xtab); cudaFree(h->d_ytab);
  cudaFree(h->d_plain); cudaFree(h->d_blur); cudaFree(h->d_score);
  cudaFree(h->d_retry); cudaFree(h->d_cand); cudaFree(h->d_sel); cudaFree(h->d_counts);
  cudaFree(h->d_img); cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_n);
  h->d_lay = nullptr; h->d_xtab = h->d_ytab = nullptr;
  h->d_plain = h->d_blur = h->d_score = nullptr;
  h->d_retry = nullptr; h->d_cand = h->d_sel = nullptr; h->d_counts = nullptr;
  h->d_img = nullptr; h->d_kps = nullptr; h->d_desc = nullptr; h->d_n = nullptr;
  h->allocated = false;
}

// cv::resize INTER_LINEAR coefficient tables (OpenCV imgproc resize.cpp, restated): for each
// destination coordinate the source offset and two 11-bit fixed-point weights.
void build_taps(int s, int d, std::vector<ResizeTap>& out, bool horizontal) {
  const double inv_scale = (double)d / s;
  const double scale = 1. / inv_scale;
  for (int i = 0; i < d; i++) {
    float f = (float)((i + 0.5) * scale - 0.5);
    int o = cv_floor_f(f);
    f -= o;
    if (horizontal) {
      if (o < 0) { f = 0; o = 0; }
      if (o >= s - 1) { f = 0; o = s - 1; }
    }
    ResizeTap t;
    t.ofs = (int16_t)o;
    t.a0 = sat_s16(cv_round_f((1.f - f) * 2048));
    t.a1 = sat_s16(cv_round_f(f * 2048));
    t.pad = 0;
    out.push_back(t);
  }
}

int setup_geometry(swm_orb* h, int w, int hh) {
  if (h->allocated && h->lay.w == w && h->lay.h == hh) return SWM_OK;
  if (h->allocated) free_frame_buffers(h);
  const int nl = h->cfg.nlevels;
  FrameLayout& L = h->lay;
  memset(&L, 0, sizeof(L));
  L.nlevels = nl;
  L.w = w;
  L.h = hh;
  std::vector<ResizeTap> xt, yt;
  long long off = 0;
  int tile_off = 0, cand_off = 0, sel_off = 0;
  for (int l = 0; l < nl; l++) {
    LevelGeom& g = L.lv[l];
    g.w = cv_round_f((float)w * h->inv_sf[l]);
    g.h = cv_round_f((float)hh * h->inv_sf[l]);
    if (g.w < 2 * kEdge + 1 || g.h < 2 * kEdge + 1 || g.w > 4000 || g.h > 4000) {
      h->err = "image size unsupported: every pyramid level must be between 39 and 4000 pixels per side";
      return SWM_E_INVALID;
    }
    const int n_ini = (int)lroundf((float)(g.w - 2 * kBand) / (float)(g.h - 2 * kBand));
    if (n_ini < 1 || n_ini > kOtMaxIni) {
      h->err = "unsupported aspect ratio (DistributeOctTree needs 1 <= round(W/H) <= 16)";
      return SWM_E_INVALID;
    }
    g.pitch = (int)align_up(kPadX + g.w + kEdge, 128);
    g.rows = g.h + 2 * kEdge;
    g.plane_off = off;
    off += align_up((long long)g.pitch * g.rows, 256);
    g.xtab_off = (int)xt.size();
    g.ytab_off = (int)yt.size();
    if (l > 0) {
      build_taps(L.lv[l - 1].w, g.w, xt, true);
      build_taps(L.lv[l - 1].h, g.h, yt, false);
    }
    g.tiles_x = (g.w - 2 * kEdge + 31) / 32;
    g.tiles_y = (g.h - 2 * kEdge + 31) / 32;
    g.tile_off = tile_off;
    tile_off += g.tiles_x * g.tiles_y;
    // strict 8-neighbour maxima cannot be adjacent: at most one per 2x2 block
    g.cand_cap = ((g.w - 2 * kEdge + 1) / 2 + 1) * ((g.h - 2 * kEdge + 1) / 2 + 1);
    g.cand_off = cand_off;
    cand_off += (g.cand_cap + 31) / 32 * 32;
    g.quota = h->quota[l];
    g.sel_cap = std::max(g.quota + 3, 4 * n_ini) + 1;
    g.sel_off = sel_off;
    sel_off += (g.sel_cap + 31) / 32 * 32;
    g.scale = h->sf[l];
    g.kp_size = (int)(kPatch * h->sf[l]);
  }
  if (xt.empty()) { xt.push_back(ResizeTap{0, 2048, 0, 0}); yt.push_back(ResizeTap{0, 2048, 0, 0}); }
  L.slab_bytes = off;
  L.tiles_total = tile_off;
  L.cand_total = cand_off;
  L.sel_total = sel_off;
  int maxkp = 0;
  for (int l = 0; l < nl; l++) maxkp += L.lv[l].sel_cap;
  h->max_kp = maxkp;

  const int B = h->cfg.max_batch;
  SWM_CK(h, cudaSetDevice(h->device));
  SWM_CK(h, cudaMalloc(&h->d_lay, sizeof(FrameLayout)));
  SWM_CK(h, cudaMalloc(&h->d_xtab, xt.size() * sizeof(ResizeTap)));
  SWM_CK(h, cudaMalloc(&h->d_ytab, yt.size() * sizeof(ResizeTap)));
  SWM_CK(h, cudaMalloc(&h->d_plain, (size_t)L.slab_bytes * B));
  SWM_CK(h, cudaMalloc(&h->d_blur, (size_t)L.slab_bytes * B));
  SWM_CK(h, cudaMalloc(&h->d_score, (size_t)L.slab_bytes * B));
  SWM_CK(h, cudaMalloc(&h->d_retry, (size_t)L.tiles_total * B));
  SWM_CK(h, cudaMalloc(&h->d_cand, (size_t)L.cand_total * B * sizeof(uint32_t)));
  SWM_CK(h, cudaMalloc(&h->d_sel, (size_t)L.sel_total * B * sizeof(uint32_t)));
  SWM_CK(h, cudaMalloc(&h->d_counts, (size_t)2 * B * nl * sizeof(int)));
  h->img_pitch = (int)align_up(w, 128);
  SWM_CK(h, cudaMalloc(&h->d_img, (size_t)h->img_pitch * hh * B));
  SWM_CK(h, cudaMalloc(&h->d_kps, (size_t)h->max_kp * B * sizeof(swm_keypoint)));
  SWM_CK(h, cudaMalloc(&h->d_desc, (size_t)h->max_kp * B * 32));
  SWM_CK(h, cudaMalloc(&h->d_n, (size_t)B * sizeof(int32_t)));
  SWM_CK(h, cudaMemcpy(h->d_lay, &L, sizeof(L), cudaMemcpyHostToDevice));
  SWM_CK(h, cudaMemcpy(h->d_xtab, xt.data(), xt.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
  SWM_CK(h, cudaMemcpy(h->d_ytab, yt.data(), yt.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
  // planes start zeroed so padding bytes are deterministic
  SWM_CK(h, cudaMemset(h->d_plain, 0, (size_t)L.slab_bytes * B));
  SWM_CK(h, cudaMemset(h->d_blur, 0, (size_t)L.slab_bytes * B));
  SWM_CK(h, cudaMemset(h->d_score, 0, (size_t)L.slab_bytes * B));
  h->allocated = true;
  return SWM_OK;
}

size_t octree_smem(const swm_orb* h) { return sizeof(OtState) + (size_t)h->max_pts * 7 + 16; }

// Enqueue the stages selected by `mask` for `batch` frames on `st`.
int enqueue(swm_orb* h, int mask, const uint8_t* d_imgs, int batch, int stride, long long frame_stride,
            swm_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_n, cudaStream_t st) {
  const FrameLayout& L = h->lay;
  const int nl = L.nlevels;
  int launches = 0;
  int* d_cand_count = h->d_counts;
  int* d_sel_count = h->d_counts + (size_t)h->cfg.max_batch * nl;
  if (mask & SWM_STAGE_PYRAMID) {
    for (int l = 0; l < nl; l++) {
      LevelArgs a;
      a.dst = L.lv[l];
      a.src = L.lv[l ? l - 1 : 0];
      a.img = d_imgs;
      a.img_stride = stride;
      a.img_frame_stride = frame_stride;
      a.plain = h->d_plain;
      a.blur = h->d_blur;
      a.score = h->d_score;
      a.slab_bytes = L.slab_bytes;
      a.xtab = h->d_xtab;
      a.ytab = h->d_ytab;
      a.min_th = h->cfg.min_th_fast;
      dim3 grid((a.dst.w + TW - 1) / TW, (a.dst.h + TH - 1) / TH, batch);
      if (l == 0) level_kernel<true><<<grid, 256, 0, st>>>(a);
      else level_kernel<false><<<grid, 256, 0, st>>>(a);
      launches++;
    }
  }
  if (mask & SWM_STAGE_NMS) {
    SWM_CK(h, cudaMemsetAsync(h->d_counts, 0, (size_t)2 * h->cfg.max_batch * nl * sizeof(int), st));
    dim3 grid(L.tiles_total, batch);
    nms_kernel<<<grid, 256, 0, st>>>(h->d_lay, h->d_score, h->cfg.ini_th_fast, 1, h->d_retry, h->d_cand, d_cand_count);
    nms_kernel<<<grid, 256, 0, st>>>(h->d_lay, h->d_score, h->cfg.ini_th_fast, 2, h->d_retry, h->d_cand, d_cand_count);
    launches += 2;
  }
  if (mask & SWM_STAGE_OCTREE) {
    dim3 grid(nl, batch);
    octree_kernel<<<grid, 512, octree_smem(h), st>>>(h->d_lay, h->d_cand, d_cand_count, h->max_pts, h->d_sel,
                                                     d_sel_count);
    launches++;
  }
  if (mask & SWM_STAGE_DESCRIBE) {
    const int eff_cap = cap < h->max_kp ? cap : h->max_kp;
    dim3 grid((eff_cap + 7) / 8, batch);
    describe_kernel<<<grid, 256, 0, st>>>(h->d_lay, h->d_plain, h->d_blur, h->d_sel, d_sel_count, d_kps, d_desc, cap,
                                          d_n);
    launches++;
  }
  SWM_CK(h, cudaGetLastError());
  h->last_launches = launches;
  return SWM_OK;
}

}  // namespace

extern "C" {

const char* swm_version(void) { return "swm_orb 0.1 sm_100a " __DATE__; }

int swm_orb_create(const swm_orb_cfg* cfg, int device, swm_orb** out) {
  if (!cfg || !out) return SWM_E_INVALID;
  *out = nullptr;
  if (cfg->nfeatures <= 0 || cfg->nlevels < 1 || cfg->nlevels > SWM_MAX_LEVELS || !(cfg->scale_factor > 1.0f) ||
      cfg->min_th_fast < 1 || cfg->ini_th_fast < cfg->min_th_fast || cfg->ini_th_fast > 254 || cfg->max_batch < 1) {
    g_create_error = "invalid swm_orb_cfg";
    return SWM_E_INVALID;
  }
  int rc = check_device(device, &g_create_error);
  if (rc != SWM_OK) return rc;
  swm_orb* h = new swm_orb;
  h->cfg = *cfg;
  h->device = device;
  h->max_pts = cfg->max_fast_per_level > 0 ? cfg->max_fast_per_level : 10000;
  if (h->max_pts > 20000) h->max_pts = 20000;
  // scale tables: float chain with a double scale factor (ORBextractor.cc:346-362, ORBextractor.h:108)
  const double s = cfg->scale_factor;
  h->sf[0] = 1.0f;
  h->sig2[0] = 1.0f;
  for (int i = 1; i < cfg->nlevels; i++) {
    h->sf[i] = (float)(h->sf[i - 1] * s);
    h->sig2[i] = h->sf[i] * h->sf[i];
  }
  for (int i = 0; i < cfg->nlevels; i++) {
    h->inv_sf[i] = 1.0f / h->sf[i];
    h->inv_sig2[i] = 1.0f / h->sig2[i];
  }
  // per-level quotas (ORBextractor.cc:367-378)
  {
    const float factor = (float)(1.0f / s);
    float per = (float)(cfg->nfeatures * (1.0f - factor) / (1.0f - std::pow((double)factor, cfg->nlevels)));
    int sum = 0;
    for (int l = 0; l < cfg->nlevels - 1; l++) {
      h->quota[l] = cv_round_f(per);
      sum += h->quota[l];
      per *= factor;
    }
    h->quota[cfg->nlevels - 1] = std::max(cfg->nfeatures - sum, 0);
    for (int l = 0; l < cfg->nlevels; l++)
      if (h->quota[l] > kOtMaxQuota) {
        g_create_error = "nfeatures too large: per-level quota exceeds 2048";
        delete h;
        return SWM_E_INVALID;
      }
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    // constant tables are per device; identical for every handle (ORBextractor.cc:380-404)
    std::call_once(g_const_once[device & 63], [&]() {
      int umax[16];
      const int hp = kHalfPatch;
      int v, v0;
      const int vmax = cv_floor_f(hp * std::sqrt(2.f) / 2 + 1);
      const int vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
      for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(std::sqrt((double)hp * hp - v * v));
      for (v = hp, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
      }
      cudaError_t e2 = cudaMemcpyToSymbol(c_umax, umax, sizeof(umax));
      if (e2 == cudaSuccess) e2 = cudaMemcpyToSymbol(c_pattern, kSwmOrbPattern, 1024);
      if (e2 != cudaSuccess) e = e2;
    });
  }
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_smem(h));
  if (e != cudaSuccess) {
    g_create_error = cuda_err("swm_orb_create", e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SWM_E_CUDA;
  }
  *out = h;
  return SWM_OK;
}

void swm_orb_destroy(swm_orb* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->allocated) free_frame_buffers(h);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* swm_last_error(const swm_orb* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int swm_orb_extract_batch_device(swm_orb* h, const uint8_t* d_imgs, int batch, int w, int h_px, int stride,
                                 size_t frame_stride, swm_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_n,
                                 void* stream) {
  if (!h) return SWM_E_INVALID;
  if (batch < 0 || !d_kps || !d_desc || !d_n || cap < 1) { h->err = "bad argument"; return SWM_E_INVALID; }
  if (batch == 0) return SWM_OK;
  if (!d_imgs || w <= 0 || h_px <= 0 || stride < w) { h->err = "bad image argument"; return SWM_E_INVALID; }
  if (batch > h->cfg.max_batch) { h->err = "batch exceeds cfg.max_batch"; return SWM_E_CAPACITY; }
  SWM_CK(h, cudaSetDevice(h->device));
  int rc = setup_geometry(h, w, h_px);
  if (rc != SWM_OK) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  h->last_img = d_imgs; h->last_stride = stride; h->last_frame_stride = (long long)frame_stride;
  h->last_kps = d_kps; h->last_desc = d_desc; h->last_n = d_n; h->last_cap = cap;
  h->last_batch = batch;
  return enqueue(h, 15, d_imgs, batch, stride, (long long)frame_stride, d_kps, d_desc, cap, d_n, st);
}

static int enqueue_host_chunk(swm_orb* h, const uint8_t* src, int nb, int w, int h_px, int stride,
                              size_t frame_stride, swm_keypoint* kps, uint8_t* desc, int cap, int32_t* n) {
  const int kcap = h->max_kp;
  const int ccap = cap < kcap ? cap : kcap;
  if (frame_stride == (size_t)stride * h_px) {
    SWM_CK(h, cudaMemcpy2DAsync(h->d_img, h->img_pitch, src, stride, w, (size_t)h_px * nb, cudaMemcpyHostToDevice,
                                h->stream));
  } else {
    for (int f = 0; f < nb; f++)
      SWM_CK(h, cudaMemcpy2DAsync(h->d_img + (size_t)f * h->img_pitch * h_px, h->img_pitch,
                                  src + (size_t)f * frame_stride, stride, w, h_px, cudaMemcpyHostToDevice, h->stream));
  }
  int rc = swm_orb_extract_batch_device(h, h->d_img, nb, w, h_px, h->img_pitch, (size_t)h->img_pitch * h_px, h->d_kps,
                                        h->d_desc, kcap, h->d_n, h->stream);
  if (rc != SWM_OK) return rc;
  SWM_CK(h, cudaMemcpyAsync(n, h->d_n, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  SWM_CK(h, cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(swm_keypoint), h->d_kps, (size_t)kcap * sizeof(swm_keypoint),
                              (size_t)ccap * sizeof(swm_keypoint), nb, cudaMemcpyDeviceToHost, h->stream));
  SWM_CK(h, cudaMemcpy2DAsync(desc, (size_t)cap * 32, h->d_desc, (size_t)kcap * 32, (size_t)ccap * 32, nb,
                              cudaMemcpyDeviceToHost, h->stream));
  h->pending_n = n;
  h->pending_batch = nb;
  h->pending_cap = cap;
  return SWM_OK;
}

int swm_orb_sync(swm_orb* h) {
  if (!h) return SWM_E_INVALID;
  SWM_CK(h, cudaSetDevice(h->device));
  SWM_CK(h, cudaStreamSynchronize(h->stream));
  if (h->pending_n) {
    for (int f = 0; f < h->pending_batch; f++)
      if (h->pending_n[f] > h->pending_cap) h->pending_n[f] = h->pending_cap;
    h->pending_n = nullptr;
  }
  return SWM_OK;
}

int swm_orb_extract_batch_async(swm_orb* h, const uint8_t* imgs, int batch, int w, int h_px, int stride,
                                size_t frame_stride, swm_keypoint* kps, uint8_t* desc, int cap, int32_t* n) {
  if (!h) return SWM_E_INVALID;
  if (batch < 1 || !imgs || !kps || !desc || !n || cap < 1 || w <= 0 || h_px <= 0 || stride < w) {
    h->err = "bad argument";
    return SWM_E_INVALID;
  }
  if (batch > h->cfg.max_batch) { h->err = "batch exceeds cfg.max_batch"; return SWM_E_CAPACITY; }
  if (h->pending_n) { h->err = "previous async batch not synchronised"; return SWM_E_STATE; }
  SWM_CK(h, cudaSetDevice(h->device));
  int rc = setup_geometry(h, w, h_px);
  if (rc != SWM_OK) return rc;
  return enqueue_host_chunk(h, imgs, batch, w, h_px, stride, frame_stride, kps, desc, cap, n);
}

int swm_orb_extract_batch(swm_orb* h, const uint8_t* imgs, int batch, int w, int h_px, int stride,
                          size_t frame_stride, swm_keypoint* kps, uint8_t* desc, int cap, int32_t* n) {
  if (!h) return SWM_E_INVALID;
  if (batch < 0 || !kps || !desc || !n || cap < 1) { h->err = "bad argument"; return SWM_E_INVALID; }
  if (batch == 0) return SWM_OK;
  if (!imgs || w <= 0 || h_px <= 0) {  // empty image: silent return (ORBextractor.cc:750-751)
    for (int f = 0; f < batch; f++) n[f] = 0;
    return SWM_OK;
  }
  if (stride < w) { h->err = "stride < width"; return SWM_E_INVALID; }
  if (h->pending_n) { h->err = "previous async batch not synchronised"; return SWM_E_STATE; }
  SWM_CK(h, cudaSetDevice(h->device));
  int rc = setup_geometry(h, w, h_px);
  if (rc != SWM_OK) return rc;
  const int B = h->cfg.max_batch;
  for (int f0 = 0; f0 < batch; f0 += B) {
    const int nb = batch - f0 < B ? batch - f0 : B;
    rc = enqueue_host_chunk(h, imgs + (size_t)f0 * frame_stride, nb, w, h_px, stride, frame_stride,
                            kps + (size_t)f0 * cap, desc + (size_t)f0 * cap * 32, cap, n + f0);
    if (rc != SWM_OK) return rc;
    rc = swm_orb_sync(h);
    if (rc != SWM_OK) return rc;
  }
  return SWM_OK;
}

int swm_orb_extract(swm_orb* h, const uint8_t* img, int w, int h_px, int stride, swm_keypoint* kps, uint8_t* desc,
                    int cap, int* n) {
  if (!h || !n) return SWM_E_INVALID;
  int32_t nn = 0;
  if (!img || w <= 0 || h_px <= 0) { *n = 0; return SWM_OK; }
  int rc = swm_orb_extract_batch(h, img, 1, w, h_px, stride, (size_t)stride * h_px, kps, desc, cap, &nn);
  *n = nn;
  return rc;
}

int swm_orb_run_stage(swm_orb* h, int stage_mask, int batch, void* stream) {
  if (!h) return SWM_E_INVALID;
  if (!h->allocated || !h->last_img) { h->err = "run_stage before extract"; return SWM_E_STATE; }
  if (batch < 1 || batch > h->last_batch) { h->err = "bad batch"; return SWM_E_INVALID; }
  SWM_CK(h, cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  return enqueue(h, stage_mask, h->last_img, batch, h->last_stride, h->last_frame_stride, h->last_kps, h->last_desc,
                 h->last_cap, h->last_n, st);
}

int swm_orb_level_ptr(swm_orb* h, int frame, int level, int which, const uint8_t** dev, int* w, int* h_px, int* pitch) {
  if (!h || !dev) return SWM_E_INVALID;
  if (!h->allocated) { h->err = "level_ptr before the first extract"; return SWM_E_STATE; }
  if (frame < 0 || frame >= h->cfg.max_batch || level < 0 || level >= h->lay.nlevels || which < 0 || which > 1) {
    h->err = "bad frame/level/which";
    return SWM_E_INVALID;
  }
  const LevelGeom& g = h->lay.lv[level];
  const uint8_t* base = which == 0 ? h->d_plain : h->d_blur;
  *dev = base + (size_t)frame * h->lay.slab_bytes + g.plane_off + (size_t)kEdge * g.pitch + kPadX;
  if (w) *w = g.w;
  if (h_px) *h_px = g.h;
  if (pitch) *pitch = g.pitch;
  return SWM_OK;
}

int swm_orb_scale_tables(const swm_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  if (!h) return SWM_E_INVALID;
  for (int i = 0; i < h->cfg.nlevels; i++) {
    if (scale) scale[i] = h->sf[i];
    if (inv_scale) inv_scale[i] = h->inv_sf[i];
    if (sigma2) sigma2[i] = h->sig2[i];
    if (inv_sigma2) inv_sigma2[i] = h->inv_sig2[i];
  }
  return SWM_OK;
}

int swm_orb_level_quotas(const swm_orb* h, int32_t* quotas) {
  if (!h || !quotas) return SWM_E_INVALID;
  for (int i = 0; i < h->cfg.nlevels; i++) quotas[i] = h->quota[i];
  return SWM_OK;
}

int swm_orb_max_keypoints(const swm_orb* h) {
  if (!h) return SWM_E_INVALID;
  if (h->allocated) return h->max_kp;
  int m = 0;  // before the first frame n_ini is unknown: assume the worst aspect ratio allowed
  for (int l = 0; l < h->cfg.nlevels; l++) m += std::max(h->quota[l] + 3, 4 * kOtMaxIni) + 1;
  return m;
}

int swm_orb_last_launches(const swm_orb* h) { return h ? h->last_launches : SWM_E_INVALID; }

int swm_orb_debug_plane(swm_orb* h, int frame, int level, int which, uint8_t* out, int out_stride) {
  if (!h || !out) return SWM_E_INVALID;
  if (!h->allocated) { h->err = "debug_plane before extract"; return SWM_E_STATE; }
  if (frame < 0 || frame >= h->cfg.max_batch || level < 0 || level >= h->lay.nlevels) return SWM_E_INVALID;
  const LevelGeom& g = h->lay.lv[level];
  SWM_CK(h, cudaSetDevice(h->device));
  SWM_CK(h, cudaStreamSynchronize(h->stream));
  const size_t fo = (size_t)frame * h->lay.slab_bytes + g.plane_off;
  if (which == 0) {
    SWM_CK(h, cudaMemcpy2D(out, out_stride, h->d_plain + fo + (kPadX - kEdge), g.pitch, g.w + 2 * kEdge, g.rows,
                           cudaMemcpyDeviceToHost));
  } else if (which == 1 || which == 2) {
    const uint8_t* base = which == 1 ? h->d_blur : h->d_score;
    SWM_CK(h, cudaMemcpy2D(out, out_stride, base + fo + (size_t)kEdge * g.pitch + kPadX, g.pitch, g.w, g.h,
                           cudaMemcpyDeviceToHost));
  } else {
    return SWM_E_INVALID;
  }
  return SWM_OK;
}

int swm_orb_debug_points(swm_orb* h, int frame, int level, int which, int32_t* xys, int cap) {
  if (!h || !xys) return SWM_E_INVALID;
  if (!h->allocated) { h->err = "debug_points before extract"; return SWM_E_STATE; }
  if (frame < 0 || frame >= h->cfg.max_batch || level < 0 || level >= h->lay.nlevels || which < 0 || which > 1)
    return SWM_E_INVALID;
  const LevelGeom& g = h->lay.lv[level];
  const int nl = h->lay.nlevels;
  SWM_CK(h, cudaSetDevice(h->device));
  SWM_CK(h, cudaStreamSynchronize(h->stream));
  int count = 0;
  const int* d_count = h->d_counts + (which ? (size_t)h->cfg.max_batch * nl : 0) + (size_t)frame * nl + level;
  SWM_CK(h, cudaMemcpy(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost));
  const int lim = which ? g.sel_cap : g.cand_cap;
  if (count > lim) count = lim;
  std::vector<uint32_t> buf(count);
  const uint32_t* d_src = which ? h->d_sel + (size_t)frame * h->lay.sel_total + g.sel_off
                                : h->d_cand + (size_t)frame * h->lay.cand_total + g.cand_off;
  if (count) SWM_CK(h, cudaMemcpy(buf.data(), d_src, (size_t)count * 4, cudaMemcpyDeviceToHost));
  const int m = count < cap ? count : cap;
  for (int i = 0; i < m; i++) {
    xys[3 * i] = (int)((buf[i] >> 8) & 0xFFF);
    xys[3 * i + 1] = (int)(buf[i] >> 20);
    xys[3 * i + 2] = (int)(buf[i] & 0xFF);
  }
  return count;
}

}  // extern "C"
