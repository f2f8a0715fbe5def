// pyr_walk.cuh -- pyramid level + border + Gaussian blur as a register-resident column walk (round 2).
//
// Reference behaviour restated (paths relative to /root/reference/code/):
//   ORBextractor::ComputePyramid   src/ORBextractor.cc:821-855  (cv::resize INTER_LINEAR 8U, copyMakeBorder REFLECT_101)
//   GaussianBlur 7x7 sigma 2       src/ORBextractor.cc:835      (cv::GaussianBlur 8U fixed point, exact)
//
// One WARP produces a strip of <= 120 columns x R rows of level l.  The strips tile the BORDERED row, columns
// [-20, w + 20): a column outside [0, w) is its reflect-101 image, so the 19-px border columns come out of the same
// walk as aligned words (no mirrored byte stores).  A lane owns four adjacent columns (one 32-bit word) and walks
// down the rows; everything between the source pixels and the two output planes stays in registers:
//   * levels >= 1: the source rows of the next-but-one step are staged in shared memory by ONE tensor-map TMA copy
//     per step (cp.async.bulk.tensor.3d over (row byte, row, frame) of the level l-1 plane, box 224 B x 8 rows; a
//     3-stage ring per warp, each stage completing on its own mbarrier: UTMALDG + SYNCS in SASS), so the walk never
//     waits on DRAM.  The horizontal bilinear pass of a source row is three
//     aligned word loads, two funnel shifts and, per column, one PRMT (lane-constant selector) + one IDP.2A on the
//     packed 11-bit taps; the last two source rows are cached so each is computed once (1.2 rows per output row at
//     scale 1.2).  The vertical pass is two IMAD.HI per pixel.  Rounding steps are cv::resize's: (sum >> 4) per source
//     row, ((b0*h0)>>16 + (b1*h1)>>16 + 2) >> 2.
//   * level 0: the word is the frame's own (aligned load, or four reflected byte loads at the image edge).
//   * the un-blurred word is stored, and once more into the border row that mirrors it (rows 1..19, h-20..h-2);
//   * blur: ten rows of the walk are kept as two 16-bit lanes per register, the vertical 7-tap sum is exact in 16 bits,
//     neighbouring lanes' sums arrive by four warp shuffles, the horizontal 7 taps are IDP.2A dot products,
//     (acc + 0x8000) >> 16.  No block barrier anywhere: a CTA is one warp.
#pragma once

namespace swm {

struct PyrArgs {
  LevelGeom dst, src;
  const uint8_t* img;
  int img_stride;
  long long img_frame_stride;
  uint8_t* plain;
  uint8_t* blur;
  long long slab_bytes;
  const ResizeTap* xtab;
  const ResizeTap* ytab;
  const CUtensorMap* src_map;  // device copy of the level l-1 plane's tensor map: u8 (pitch, rows, frames), box 224 x 8 x 1
  int strip_lanes, nstrips;  // useful lanes (4 columns each) per warp, <= kWalkLanes, and strips per row block
  int rows_per_job, nrb;     // rows per warp (multiple of 4) and row blocks per level
};

constexpr int kWalkLanes = 30;      // lanes 0 and 31 carry the blur's left / right neighbours
constexpr int kWalkX0 = -20;        // first column of the bordered row that the strips cover (border = 19, word aligned)
constexpr int kWalkStages = 3;      // TMA ring depth (steps of 4 output rows)
constexpr int kWalkStageRows = 8;   // source rows per step: 4 * 1.5 + 1 at most (scale factor <= 1.5)
constexpr int kWalkRowBytes = 224;  // staged bytes per source row: 128 columns * 1.5 + 12 + alignment


__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// one TMA tile: box (kWalkRowBytes, kWalkStageRows, 1) of the tensor map at (c0, c1, c2) -> shared, completing on bar
__device__ __forceinline__ void tma_tile_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

template <bool kFirst>
__global__ void __launch_bounds__(32, 24) pyr_walk_kernel(const PyrArgs a) {
  // A CTA is ONE warp: the strip / row-block indices come from blockIdx, so nothing in the row loop depends on a
  // thread index except the column.
  __shared__ __align__(128) uint8_t s_stage[kFirst ? 16 : kWalkStages * kWalkStageRows * kWalkRowBytes];
  __shared__ __align__(8) uint64_t s_bar[kWalkStages];
  const int lane = threadIdx.x;
  const int f = blockIdx.z;
  const int rb = blockIdx.y, st = blockIdx.x;
  const int w = a.dst.w, h = a.dst.h, pitch = a.dst.pitch;
  const int y0 = rb * a.rows_per_job, y1 = min(y0 + a.rows_per_job, h);
  const int X = kWalkX0 + 4 * (st * a.strip_lanes + lane - 1);  // my four columns: X .. X+3
  const bool mine = lane >= 1 && lane <= a.strip_lanes;
  // un-blurred plane incl. the 19 border columns: whole words; the first / last word overhang the border by up to
  // three bytes, which land in the row padding (kPadX on the left, pitch slack or the next row's padding on the right)
  const bool pstore = mine && X + 3 >= -kEdge && X < w + kEdge;
  // blurred plane: columns [0, w) only; a last partial word is written byte by byte
  const int bcount = (mine && X >= 0) ? min(max(w - X, 0), 4) : 0;
  const bool has_partial = __any_sync(0xffffffffu, bcount > 0 && bcount < 4);  // warp-uniform: only the last strip, if at all
  const long long roi0 = (long long)f * a.slab_bytes + a.dst.plane_off + (long long)kEdge * pitch + kPadX;
  // running pointers: my word in un-blurred row v and in blurred row v - 8 (the row the blur emits while v is made)
  uint8_t* pr = a.plain + roi0 + (long long)(y0 - 5) * pitch + X;
  uint8_t* br = a.blur + roi0 + (long long)(y0 - 8) * pitch + X;

  // ---- per-lane source set-up
  const uint8_t* src = nullptr;
  int cx0 = 0, cx1 = 0, cx2 = 0, cx3 = 0;
  bool fastlane = false;
  uint32_t s8 = 0, sel0 = 0, sel1 = 0, sel2 = 0, sel3 = 0, a01_0 = 0, a01_1 = 0, a01_2 = 0, a01_3 = 0;
  int spitch = 0, sh = 0, xb = 0;
  const ResizeTap* __restrict__ yt = nullptr;
  uint32_t lane_addr = 0;                       // shared address of my 12 source bytes in row 0 of stage 0
  const uint32_t stage_u32 = smem_u32(s_stage), bar_u32 = smem_u32(s_bar);
  (void)stage_u32; (void)bar_u32; (void)lane_addr; (void)xb; (void)sh;
  if (kFirst) {
    src = a.img + (long long)f * a.img_frame_stride;
    spitch = a.img_stride;
    const bool aligned = ((reinterpret_cast<unsigned long long>(src) | (unsigned long long)a.img_stride) & 3ull) == 0;
    fastlane = aligned && X >= 0 && X + 3 < w;
    cx0 = reflect101(X, w);
    cx1 = reflect101(X + 1, w);
    cx2 = reflect101(X + 2, w);
    cx3 = reflect101(X + 3, w);
  } else {
    const ResizeTap* __restrict__ xt = a.xtab + a.dst.xtab_off;
    yt = a.ytab + a.dst.ytab_off;
    sh = a.src.h;
    ResizeTap t[4];
    int omin = 1 << 30;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      t[k] = xt[reflect101(X + k, w)];
      omin = min(omin, (int)t[k].ofs);
    }
    // a lane reads 12 source bytes from the aligned word below its left-most tap; after the funnel shift by s8 bits
    // byte 0 is that tap.  The warp's rows are staged from the 16-byte aligned column xb (the box is kWalkRowBytes
    // wide; the host checks that every lane's 12 bytes fall inside it).
    const int wbase = omin & ~3;
    xb = __reduce_min_sync(0xffffffffu, wbase) & ~15;
    lane_addr = stage_u32 + (uint32_t)(wbase - xb);
    s8 = (uint32_t)(omin & 3) * 8u;
    uint32_t sel[4], a01[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t d = (uint32_t)(t[k].ofs - omin);  // <= 6 for scale factors <= 1.5 (checked on the host)
      sel[k] = d | ((d + 1u) << 4) | 0x4400u;
      a01[k] = (uint32_t)(uint16_t)t[k].a0 | ((uint32_t)(uint16_t)t[k].a1 << 16);
    }
    sel0 = sel[0]; sel1 = sel[1]; sel2 = sel[2]; sel3 = sel[3];
    a01_0 = a01[0]; a01_1 = a01[1]; a01_2 = a01[2]; a01_3 = a01[3];
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kWalkStages; k++) mbar_init(&s_bar[k], 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
  const int nit = ((y1 - y0 + 3) >> 2) + 2;

  // Step j makes rows v .. v+3 and needs source rows lo .. lo+7 of level l-1: lanes 0-3 read the four row taps (they
  // are handed to the row loop by shuffle later), one lane issues the TMA tile.  Returns lo; tap = my row's tap.
  uint32_t psg = 0, csg = 0, cphase = 0;  // producer stage, consumer stage and its mbarrier phase
  auto prefetch = [&](int j, int2& tap) -> int {
    int s0 = 1 << 30;
    tap = make_int2(0, 0);
    if (lane < 4) {
      const int v = y0 - 5 + 4 * j + lane;
      const int ay = v < 0 ? -v : (v >= h ? 2 * h - 2 - v : v);
      tap = __ldg(reinterpret_cast<const int2*>(yt + ay));  // (ofs, a0), (a1, pad)
      s0 = (int)(short)(tap.x & 0xFFFF);
    }
    const int lo = __reduce_min_sync(0xffffffffu, s0);
    if (lane == 0) {
      const uint32_t bar = bar_u32 + 8u * psg;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kWalkStageRows * kWalkRowBytes) : "memory");
      tma_tile_3d(stage_u32 + psg * (kWalkStageRows * kWalkRowBytes), a.src_map, kPadX + xb, kEdge + lo, f, bar);
    }
    psg = psg == kWalkStages - 1 ? 0 : psg + 1;  // steps are prefetched in order: stage = step mod kWalkStages
    return lo;
  };
  int lo_cur = 0, lo_n1 = 0, lo_n2 = 0;
  int2 tap_cur = make_int2(0, 0), tap_n1 = tap_cur, tap_n2 = tap_cur;
  if (!kFirst) {
    lo_cur = prefetch(0, tap_cur);
    lo_n1 = prefetch(1, tap_n1);  // nit >= 3
  }

  // horizontal bilinear pass of staged source row r for my four columns (cv::resize: sum >> 4 kept, 11-bit taps)
  uint32_t row0_addr = 0;  // my bytes in the staged row that holds source row 0 (may lie below the stage: only rows >= lo are read)
  auto hrow = [&](int r, uint32_t (&hv)[4]) {
    const uint32_t p = row0_addr + (uint32_t)r * kWalkRowBytes;
    const uint32_t w0 = lds32(p), w1 = lds32(p + 4), w2 = lds32(p + 8);
    const uint32_t u0 = __funnelshift_r(w0, w1, s8), u1 = __funnelshift_r(w1, w2, s8);
    hv[0] = __dp2a_lo(a01_0, __byte_perm(u0, u1, sel0), 0u) >> 4;
    hv[1] = __dp2a_lo(a01_1, __byte_perm(u0, u1, sel1), 0u) >> 4;
    hv[2] = __dp2a_lo(a01_2, __byte_perm(u0, u1, sel2), 0u) >> 4;
    hv[3] = __dp2a_lo(a01_3, __byte_perm(u0, u1, sel3), 0u) >> 4;
  };
  // the two most recent source rows; which of them is the upper row of the current output row alternates, so the
  // vertical weights are swapped instead of the rows
  uint32_t H0[4] = {0, 0, 0, 0}, H1[4] = {0, 0, 0, 0};
  int r0 = -1, r1 = -1;

  // level 0: my word of frame row v (reflect-101 rows; -5 <= v <= h + 9 and h >= 39).  The words of the NEXT step are
  // loaded one step ahead, so the walk does not wait on the frame either.
  auto frame_word = [&](int v) -> uint32_t {
    const int ay = v < 0 ? -v : (v >= h ? 2 * h - 2 - v : v);
    const uint8_t* row = src + (long long)ay * spitch;
    if (fastlane) return __ldg(reinterpret_cast<const uint32_t*>(row + X));
    return (uint32_t)__ldg(row + cx0) | ((uint32_t)__ldg(row + cx1) << 8) | ((uint32_t)__ldg(row + cx2) << 16) |
           ((uint32_t)__ldg(row + cx3) << 24);
  };
  uint32_t cw[4] = {0, 0, 0, 0}, nw[4] = {0, 0, 0, 0};
  if (kFirst) {
#pragma unroll
    for (int i = 0; i < 4; i++) cw[i] = frame_word(y0 - 5 + i);
  }

  uint32_t e[10], o[10];
#pragma unroll
  for (int k = 0; k < 10; k++) e[k] = o[k] = 0;
#pragma unroll 1
  for (int it = 0; it < nit; it++) {
    const int vb = y0 - 5 + 4 * it;  // two warm-up steps fill the blur window (rows y0-5 .. y0+2)
    if (kFirst) {
#pragma unroll
      for (int i = 0; i < 4; i++) nw[i] = frame_word(vb + 4 + i);
    }
    if (!kFirst) {
      if (it + 2 < nit) lo_n2 = prefetch(it + 2, tap_n2);
      {  // wait for this step's tile (a lost transaction must abort, not hang the GPU)
        const uint32_t bar = bar_u32 + 8u * csg;
        uint32_t done;
        int spins = 0;
        do {
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                       : "=r"(done) : "r"(bar), "r"(cphase) : "memory");
          if (!done && ++spins > (1 << 22)) __trap();
        } while (!done);
      }
      row0_addr = lane_addr + (csg * kWalkStageRows - (uint32_t)lo_cur) * kWalkRowBytes;
      if (csg == kWalkStages - 1) { csg = 0; cphase ^= 1u; }
      else csg++;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int v = vb + i;
      uint32_t word, ev, ov;
      if (kFirst) {
        word = cw[i];
        ev = word & 0x00FF00FFu;
        ov = (word >> 8) & 0x00FF00FFu;
      } else {
        const uint32_t tx = (uint32_t)__shfl_sync(0xffffffffu, tap_cur.x, i), ty = (uint32_t)__shfl_sync(0xffffffffu, tap_cur.y, i);
        const int s0 = (int)(short)(tx & 0xFFFF);
        const int s1 = min(s0 + 1, sh - 1);  // second source row, clamped like cv::resize
        const uint32_t B0 = tx & 0xFFFF0000u, B1 = ty << 16;
        uint32_t W0, W1;
        if (r0 == s0) {
          if (r1 != s1) { hrow(s1, H1); r1 = s1; }
          W0 = B0; W1 = B1;
        } else if (r1 == s0) {
          if (r0 != s1) { hrow(s1, H0); r0 = s1; }
          W0 = B1; W1 = B0;
        } else {
          hrow(s0, H0); r0 = s0;
          if (r1 != s1) { hrow(s1, H1); r1 = s1; }
          W0 = B0; W1 = B1;
        }
        // vertical pass: ((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2, the products as high words of (b << 16) * h
        const uint32_t p0 = (__umulhi(W1, H1[0]) + (__umulhi(W0, H0[0]) + 2u)) >> 2;
        const uint32_t p1 = (__umulhi(W1, H1[1]) + (__umulhi(W0, H0[1]) + 2u)) >> 2;
        const uint32_t p2 = (__umulhi(W1, H1[2]) + (__umulhi(W0, H0[2]) + 2u)) >> 2;
        const uint32_t p3 = (__umulhi(W1, H1[3]) + (__umulhi(W0, H0[3]) + 2u)) >> 2;
        ev = p0 | (p2 << 16);
        ov = p1 | (p3 << 16);
        word = ev | (ov << 8);
      }
      if (v >= y0 && v < y1) {
        if (pstore) *reinterpret_cast<uint32_t*>(pr) = word;
        // copyMakeBorder(REFLECT_101, 19), rows: row v is also the border row that reflects onto it (ORBextractor.cc:846-851)
        if (v <= kEdge || v >= h - 1 - kEdge) {
          if (v >= 1 && v <= kEdge && pstore) *reinterpret_cast<uint32_t*>(pr - 2ll * v * pitch) = word;
          if (v >= h - 1 - kEdge && v <= h - 2 && pstore) *reinterpret_cast<uint32_t*>(pr + 2ll * (h - 1 - v) * pitch) = word;
        }
      }
      pr += pitch;
      e[6 + i] = ev;
      o[6 + i] = ov;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (it >= 2) {
        const int gy = vb - 3 + i;
        // vertical 7 taps (18 34 48 56 48 34 18) on (px0, px2) and (px1, px3): <= 255 * 256, exact in 16 bits
        const uint32_t ve = 18u * (e[i] + e[i + 6]) + 34u * (e[i + 1] + e[i + 5]) + 48u * (e[i + 2] + e[i + 4]) + 56u * e[i + 3];
        const uint32_t vo = 18u * (o[i] + o[i + 6]) + 34u * (o[i + 1] + o[i + 5]) + 48u * (o[i + 2] + o[i + 4]) + 56u * o[i + 3];
        const uint32_t P01 = __byte_perm(ve, vo, 0x5410), P23 = __byte_perm(ve, vo, 0x7632);  // (V0,V1), (V2,V3)
        const uint32_t A01 = __shfl_up_sync(0xffffffffu, P01, 1), A23 = __shfl_up_sync(0xffffffffu, P23, 1);
        const uint32_t C01 = __shfl_down_sync(0xffffffffu, P01, 1), C23 = __shfl_down_sync(0xffffffffu, P23, 1);
        // V index relative to my first pixel: A = (-4,-3),(-2,-1); P = (0,1),(2,3); C = (4,5),(6,7)
        const uint32_t m3m2 = __byte_perm(A01, A23, 0x5432);
        const uint32_t m1p0 = __byte_perm(A23, P01, 0x5432);
        const uint32_t p1p2 = __byte_perm(P01, P23, 0x5432);
        const uint32_t p3p4 = __byte_perm(P23, C01, 0x5432);
        const uint32_t p5p6 = __byte_perm(C01, C23, 0x5432);
        const uint32_t k01 = 18u | (34u << 8), k23 = 48u | (56u << 8), k45 = 48u | (34u << 8), k6 = 18u;
        uint32_t o0 = __dp2a_lo(m3m2, k01, 0x8000u);
        o0 = __dp2a_lo(m1p0, k23, o0);
        o0 = __dp2a_lo(p1p2, k45, o0);
        o0 = __dp2a_lo(p3p4, k6, o0);
        uint32_t o1 = __dp2a_lo(A23, k01, 0x8000u);
        o1 = __dp2a_lo(P01, k23, o1);
        o1 = __dp2a_lo(P23, k45, o1);
        o1 = __dp2a_lo(C01, k6, o1);
        uint32_t o2 = __dp2a_lo(m1p0, k01, 0x8000u);
        o2 = __dp2a_lo(p1p2, k23, o2);
        o2 = __dp2a_lo(p3p4, k45, o2);
        o2 = __dp2a_lo(p5p6, k6, o2);
        uint32_t o3 = __dp2a_lo(P01, k01, 0x8000u);
        o3 = __dp2a_lo(P23, k23, o3);
        o3 = __dp2a_lo(C01, k45, o3);
        o3 = __dp2a_lo(C23, k6, o3);
        const uint32_t bw = __byte_perm(__byte_perm(o0, o1, 0x0062), __byte_perm(o2, o3, 0x0062), 0x5410);
        if (gy < y1) {
          if (bcount == 4) *reinterpret_cast<uint32_t*>(br) = bw;
          if (has_partial && bcount > 0 && bcount < 4) {  // the word that holds column w - 1 when w is not a multiple of 4
            br[0] = (uint8_t)bw;
            if (bcount > 1) br[1] = (uint8_t)(bw >> 8);
            if (bcount > 2) br[2] = (uint8_t)(bw >> 16);
          }
        }
      }
      br += pitch;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
      e[k] = e[k + 4];
      o[k] = o[k + 4];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) cw[i] = nw[i];
    lo_cur = lo_n1;
    lo_n1 = lo_n2;
    tap_cur = tap_n1;
    tap_n1 = tap_n2;
  }
}

}  // namespace swm
