// fast_tile.cuh -- FAST-9/16 detection, score, tile retry and non-max suppression with one WARP per run of 32x32 tiles
// (round 2; replaces the 256-thread block kernel of round 1).
//
// Reference behaviour restated (paths relative to /root/reference/code/):
//   tileCalcKeypoints_kernel   src/cuda/Fast_gpu.cu:284-341  (deterministic lock-step form; tiles anchored at (19,19))
//   isKeyPoint2 / cornerScore  src/cuda/Fast_gpu.cu:70-266   (largest threshold at which the pixel is still a corner)
//   ComputeKeyPointsOctTree    src/ORBextractor.cc:691-744   (iniThFAST first, minThFAST where a tile found nothing)
//
// A CTA is ONE warp that walks a horizontal run of up to kFT tiles of one level.  Per tile:
//   * the 80 x 40 byte window (tile + 4 px halo, 16-byte aligned start) arrives by ONE tensor-map TMA copy
//     (cp.async.bulk.tensor.3d, UTMALDG in SASS); the next tile's copy is issued as soon as the scores of this one are
//     done and lands while it is non-max suppressed (a second window buffer measured slower: fewer warps per SM);
//   * packed quick reject, 8 pixels (an aligned word pair) per item: the four opposite ring pairs with VABSDIFF4 + carry-trick band test
//     (exact: every 9-arc holds one pixel of each pair); survivors are compacted with warp ballots (no block scan);
//   * survivors are scored on two 16-bit lanes (bright ring, dark ring): three-wide minima, then minima of three of
//     those = all 9-arcs, then the maximum (VIMNMX3.U16x2);
//   * strict 3x3 non-max suppression in the shared score tile; warp-aggregated append to the (frame, level) list.
//   pass 1: keypoint iff S_hi(p) > S_hi(q) for the 8 neighbours (S_hi: scores below iniThFAST count as 0), so it detects
//           at iniThFAST; a tile without a keypoint is flagged for retry and its run goes on the retry list;
//   pass 2 (persistent grid over the retry list): only flagged tiles are redone, at minThFAST, against
//           S_eff = S in retried tiles, S_hi elsewhere.
// No block barrier, no atomics in shared memory.  Append order is unordered; everything downstream orders by explicit keys.
#pragma once

namespace swm {

constexpr int kFT = 4;                       // most tiles per run (one warp); FastArgs::run_len <= kFT is what a handle uses
// Staged window: FT_PITCH x 40 bytes from level (X0 - 19, Y0 - 4), X0 - 19 = 32 * tile: the TMA needs a 16-byte aligned start
// (an unaligned first coordinate faults), so the first interior pixel sits at local (kFx, 4) = (19, 4).
#ifndef SWM_FT_PITCH
#define SWM_FT_PITCH 80  // 80: rows land 20 banks apart, fewer shared-memory conflicts than 64 (0.727 vs 0.741 ms); the window needs 56
#endif
constexpr int FT_PITCH = SWM_FT_PITCH, FT_ROWS = 40, kFx = 19;
constexpr int FT_WORDS = FT_PITCH / 4;
constexpr int FT_QW0 = 4, FT_QW = 9;         // words per row that hold scored pixels (local x 18 .. 51): words 4 .. 12
constexpr int FT_SCP = 48, FT_SCX = 16;      // score tile: pitch and the local x of its column 0
#ifndef SWM_FT_BUFS
#define SWM_FT_BUFS 1  // 1: single window buffer, the next tile's copy starts when the scores are done: more warps per SM (0.741 vs 0.765 ms); 2: double buffer
#endif
constexpr int FT_BUFS = SWM_FT_BUFS;
constexpr int FT_QP = 5;                     // quick-reject items per row: the aligned word pairs (4,5) .. (12,13)
constexpr int FT_PITEMS = 34 * FT_QP;        // items per tile (rows 3 .. 36)
constexpr int FT_PIT = (FT_PITEMS + 31) / 32;
constexpr int FT_LIST = 34 * 34 + 4;

// Quick-reject band test of four pixels: bit 7 of a byte of the result is set if |a - v| > th (c7 = (127 - th) *
// 0x01010101).  The carry of a byte sum into its neighbour is NOT masked off: it can only set the flag of a pixel with
// |a - v| == th, i.e. let one more pixel through to the exact score, never drop one -- and the integer ALU pipe, which
// bounds this kernel (ncu: 78 % busy), is spared one LOP3 per test.  Bits other than 7, 15, 23, 31 are garbage.
__device__ __forceinline__ uint32_t oob_bits(uint32_t a, uint32_t v, uint32_t c7) {
  const uint32_t ad = __vabsdiffu4(a, v);
  return (ad + c7) | ad;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32s(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// FAST score of the pixel at shared address c (row pitch FT_PITCH) on two 16-bit lanes per register: lane 0 = bright
// ring r - v, lane 1 = dark ring v - r, both biased by 0x4000 so that they stay positive and ONE IMAD (FMA pipe) makes
// both from a ring byte: r * 0xFFFF0001 = (r, -r), plus (0x4000 - v, 0x4000 + v); no carry crosses the lanes.  min over
// every 9-arc = min3 of three 3-wide minima; maximum over the 16 arcs; both polarities at once (VIMNMX3.U16x2).  A ring
// difference below zero needs no clamp: an arc that holds one cannot reach a positive minimum.  Same value as
// swm::fast_score (swm_core.cuh) without its early-outs; checked bit-exact on the device by the parity tests.
__device__ __forceinline__ int fast_score_x2(uint32_t c, int th) {
  constexpr int P = FT_PITCH;
  const uint32_t v = lds8(c);
  const uint32_t kv = (0x4000u - v) | ((0x4000u + v) << 16);
  uint32_t d[16];
#define SWM_RING(k, off) d[k] = lds8(c + (uint32_t)(off)) * 0xFFFF0001u + kv
  SWM_RING(0, 3 * P);      SWM_RING(1, 3 * P + 1);   SWM_RING(2, 2 * P + 2);   SWM_RING(3, P + 3);
  SWM_RING(4, 3);          SWM_RING(5, -P + 3);      SWM_RING(6, -2 * P + 2);  SWM_RING(7, -3 * P + 1);
  SWM_RING(8, -3 * P);     SWM_RING(9, -3 * P - 1);  SWM_RING(10, -2 * P - 2); SWM_RING(11, -P - 3);
  SWM_RING(12, -3);        SWM_RING(13, P - 3);      SWM_RING(14, 2 * P - 2);  SWM_RING(15, 3 * P - 1);
#undef SWM_RING
  uint32_t p3[16];
#pragma unroll
  for (int k = 0; k < 16; k++) p3[k] = __vimin3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  uint32_t best = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const uint32_t m0 = __vimin3_u16x2(p3[k], p3[(k + 3) & 15], p3[(k + 6) & 15]);
    const uint32_t m1 = __vimin3_u16x2(p3[k + 1], p3[(k + 4) & 15], p3[(k + 7) & 15]);
    best = __vimax3_u16x2(best, m0, m1);
  }
  const int b = max((int)(best & 0xFFFFu), (int)(best >> 16)) - 0x4000;
  return b > th ? b - 1 : 0;
}

struct FastArgs {
  const FrameLayout* L;
  const CUtensorMap* maps;   // per level: un-blurred plane (row byte, row, frame), box FT_PITCH x FT_ROWS x 1
  const int4* fblk_desc;     // per run: (level, bx, by, -)
  int ini_th, min_th;
  uint8_t* retry;            // per (frame, tile): 1 = no pass-1 keypoint
  int* retry_list;           // [0] = count, then frame * runs + run
  uint32_t* cand;
  int* cand_count;
  uint8_t* dbg_score;        // parity introspection: score map S at minThFAST (pass 1 only), or null
  int run_len;               // tiles per run: kFT for batches (set-up amortised), 1 for a single-frame handle (latency)
};

// kDbg (pass 1 only): keep the score map S at minThFAST for the parity tests; the product instantiation detects at
// iniThFAST, where every non-zero score already is >= iniThFAST and the S_hi clamp of the neighbours is a no-op.
template <int kPass, bool kDbg>
__global__ void __launch_bounds__(32, 24) fast_tile_kernel(const FastArgs a) {
  __shared__ __align__(128) uint8_t s_px[FT_BUFS][FT_ROWS * FT_PITCH];
  __shared__ __align__(16) uint8_t s_sc[FT_ROWS * FT_SCP];
  __shared__ uint16_t s_list[FT_LIST];
  __shared__ uint32_t s_colmask[2 * FT_QP];
  __shared__ uint8_t s_flag[3 * (kFT + 2)];
  __shared__ __align__(8) uint64_t s_bar[2];
  const FrameLayout* __restrict__ L = a.L;
  const int lane = threadIdx.x;
  const uint32_t lt = (1u << lane) - 1u;
  const int nrun = L->fblk_total;
  const uint32_t px_u32 = smem_u32(s_px), sc_u32 = smem_u32(s_sc), bar_u32 = smem_u32(s_bar), list_u32 = smem_u32(s_list);
  if (lane == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  uint32_t phbits = 0;
  bool cm_valid = false;  // s_colmask holds the masks of an unclipped tile
  const int th_run = (kPass == 1 && !kDbg) ? a.ini_th : a.min_th;
  const uint32_t c7 = (uint32_t)(127 - th_run) * 0x01010101u;

  for (int work = blockIdx.x;; work += gridDim.x) {
    int fb;
    if (kPass == 1) {
      if (work != (int)blockIdx.x) break;
      fb = blockIdx.y * nrun + blockIdx.x;
    } else {
      if (work >= a.retry_list[0]) break;
      fb = a.retry_list[1 + work];
    }
    const int f = kPass == 1 ? (int)blockIdx.y : fb / nrun;
    const int4 bd = __ldg(&a.fblk_desc[fb - f * nrun]);
    const int lvl = bd.x, bx = bd.y, by = bd.z;
    const LevelGeom& g = L->lv[lvl];
    const int w = g.w, h = g.h;
    uint8_t* fretry = a.retry + (long long)f * L->tiles_total + g.tile_off;
    const int tx0 = a.run_len * bx;  // first tile of the run
    const int t0 = by * g.tiles_x + tx0;
    const int ntile = min(a.run_len, g.tiles_x - tx0);
    const int Y0 = kEdge + 32 * by;
    if (kPass == 2) {
      __syncwarp();
      if (lane < 3 * (kFT + 2)) {
        const int ny = by + lane / (kFT + 2) - 1, nx = tx0 + lane % (kFT + 2) - 1;
        s_flag[lane] = (ny >= 0 && ny < g.tiles_y && nx >= 0 && nx < g.tiles_x) ? fretry[ny * g.tiles_x + nx] : 0;
      }
      __syncwarp();
    }
    const CUtensorMap* map = a.maps + lvl;
    auto active = [&](int j) -> int {  // first tile >= j of the run that has to be processed
      if (kPass == 2)
        while (j < ntile && s_flag[(kFT + 2) + 1 + j] == 0) j++;
      return j;
    };
    auto issue = [&](int j, int buf) {
      if (lane == 0) {
        const uint32_t bar = bar_u32 + 8u * buf;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(FT_ROWS * FT_PITCH) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                px_u32 + buf * (FT_ROWS * FT_PITCH)),
            "l"(map), "r"(kPadX + 32 * (tx0 + j)), "r"(kEdge + Y0 - 4), "r"(f), "r"(bar)
            : "memory");
      }
    };
    int anymask = 0;
    int buf = 0;
    int j = active(0);
    if (j < ntile) issue(j, buf);
    while (j < ntile) {
      const int jn = active(j + 1);
      __syncwarp();  // every lane is done with the other buffer (tile before this one)
      if (FT_BUFS == 2 && jn < ntile) issue(jn, buf ^ 1);
      const int X0 = kEdge + 32 * (tx0 + j);   // first interior pixel (level coords); staged origin (X0-19, Y0-4)
      // score tile cleared; scored columns of each word: interior +- 1 px and level x in [19, w - 19)
      for (int i = lane; i < FT_ROWS * FT_SCP / 16; i += 32) reinterpret_cast<uint4*>(s_sc)[i] = make_uint4(0, 0, 0, 0);
      // (only the first tile of a level row and one that reaches the right end of the FAST band have clipped masks)
      const bool clipped = X0 + 32 >= w - kEdge || X0 - 1 < kEdge;
      if (!cm_valid || clipped) {
        if (lane < 2 * FT_QP) {
          uint32_t m = 0;
          for (int k = 0; k < 4; k++) {
            const int lx = 4 * (FT_QW0 + lane) + k, gx = X0 - kFx + lx;
            if (lx >= kFx - 1 && lx <= kFx + 32 && gx >= kEdge && gx < w - kEdge) m |= 0x80u << (8 * k);
          }
          s_colmask[lane] = m;
        }
        cm_valid = !clipped;
      }
      mbar_wait(&s_bar[buf], (phbits >> buf) & 1u);
      phbits ^= 1u << buf;
      __syncwarp();
      const uint32_t pxb = px_u32 + buf * (FT_ROWS * FT_PITCH);

      // ---- packed quick reject on the scored region (interior + 1 px): rows ly 3..36, words 4..12.  An item is an
      // aligned PAIR of words (8 pixels): 11 loads (LDS.64 for the pair, LDS.32 for its left / right neighbour) serve
      // both words, all four opposite ring pairs are tested unconditionally (a warp hardly ever leaves early).
      int n_cand = 0;
#pragma unroll 1
      for (int it = 0; it < FT_PIT; it++) {
        const int i = lane + 32 * it;
        uint32_t m0 = 0, m1 = 0;
        int ly = 0, p = 0;
        if (i < FT_PITEMS) {
          const int r = i / FT_QP;
          p = i - r * FT_QP;
          ly = r + 3;
          const int gy = Y0 - 4 + ly;
          if (gy >= kEdge && gy < h - kEdge) {
            const uint32_t base = pxb + (uint32_t)(ly * FT_PITCH + 4 * FT_QW0 + 8 * p);  // first word of the pair
            const uint2 c = lds64(base), cm3 = lds64(base - 3 * FT_PITCH), cp3 = lds64(base + 3 * FT_PITCH);
            const uint2 cp = lds64(base + 2 * FT_PITCH), cm = lds64(base - 2 * FT_PITCH);
            const uint32_t l0 = lds32s(base - 4), r0 = lds32s(base + 8);
            const uint32_t lp = lds32s(base + 2 * FT_PITCH - 4), rp = lds32s(base + 2 * FT_PITCH + 8);
            const uint32_t lm = lds32s(base - 2 * FT_PITCH - 4), rm = lds32s(base - 2 * FT_PITCH + 8);
            const uint32_t v0 = c.x, v1 = c.y;
            const uint32_t fpp = __funnelshift_r(cp.x, cp.y, 16), fmm = __funnelshift_r(cm.x, cm.y, 16);  // rows +2 / -2, between the words
            m0 = (oob_bits(cm3.x, v0, c7) | oob_bits(cp3.x, v0, c7)) & s_colmask[2 * p];
            m0 &= oob_bits(__funnelshift_r(v0, v1, 24), v0, c7) | oob_bits(__funnelshift_r(l0, v0, 8), v0, c7);
            m0 &= oob_bits(fpp, v0, c7) | oob_bits(__funnelshift_r(lm, cm.x, 16), v0, c7);
            m0 &= oob_bits(fmm, v0, c7) | oob_bits(__funnelshift_r(lp, cp.x, 16), v0, c7);
            m1 = (oob_bits(cm3.y, v1, c7) | oob_bits(cp3.y, v1, c7)) & s_colmask[2 * p + 1];
            m1 &= oob_bits(__funnelshift_r(v1, r0, 24), v1, c7) | oob_bits(__funnelshift_r(v0, v1, 8), v1, c7);
            m1 &= oob_bits(__funnelshift_r(cp.y, rp, 16), v1, c7) | oob_bits(fmm, v1, c7);
            m1 &= oob_bits(__funnelshift_r(cm.y, rm, 16), v1, c7) | oob_bits(fpp, v1, c7);
          }
        }
        // survivors -> list: exclusive prefix of the per-lane counts (0..8) by a shuffle scan
        const int cnt = __popc(m0) + __popc(m1);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        uint32_t addr = list_u32 + 2u * (uint32_t)(n_cand + incl - cnt);
        uint32_t e = (uint32_t)((ly << 8) | (4 * FT_QW0 + 8 * p));
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if ((k < 4 ? m0 : m1) & (0x80u << (8 * (k & 3)))) {
            sts16(addr, e);
            addr += 2;
          }
          e++;
        }
        n_cand += __shfl_sync(0xffffffffu, incl, 31);
      }
      __syncwarp();

      // ---- score the survivors; those that can be keypoints are compacted to the front of the same list
      int n2 = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < n_cand; c0 += 32) {
        const int c = c0 + lane;
        bool q = false;
        uint32_t e = 0;
        if (c < n_cand) {
          e = lds16(list_u32 + 2u * c);
          const int ly = e >> 8, lx = e & 255;
          const int sc = fast_score_x2(pxb + (uint32_t)(ly * FT_PITCH + lx), th_run);
          sts8(sc_u32 + (uint32_t)(ly * FT_SCP + lx - FT_SCX), (uint32_t)sc);
          q = sc >= (kPass == 1 ? a.ini_th : 1) && (unsigned)(lx - kFx) < 32u && (unsigned)(ly - 4) < 32u;  // tile interior
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, q);
        __syncwarp();
        if (q) sts16(list_u32 + 2u * (uint32_t)(n2 + __popc(bal & lt)), e);
        n2 += __popc(bal);
      }
      __syncwarp();

      if (FT_BUFS == 1 && jn < ntile) issue(jn, 0);  // the window is no longer read (the score loop ended with a __syncwarp)
      // ---- non-max suppression (strict, 8 neighbours) of the short list
      bool tile_any = false;
#pragma unroll 1
      for (int c0 = 0; c0 < n2; c0 += 32) {
        const int c = c0 + lane;
        bool kp = false;
        int lx = 0, ly = 0, sc = 0;
        if (c < n2) {
          const int e = (int)lds16(list_u32 + 2u * c);
          ly = e >> 8;
          lx = e & 255;
          const uint32_t sp = sc_u32 + (uint32_t)(ly * FT_SCP + lx - FT_SCX);
          sc = (int)lds8(sp);
          if (kPass == 1 && !kDbg) {
            // every stored score is 0 or >= iniThFAST: strict maximum of the raw 3x3 neighbourhood
            const int m0 = max(max((int)lds8(sp - FT_SCP - 1), (int)lds8(sp - FT_SCP)), (int)lds8(sp - FT_SCP + 1));
            const int m1 = max(max((int)lds8(sp - 1), (int)lds8(sp + 1)), (int)lds8(sp + FT_SCP - 1));
            const int m2 = max((int)lds8(sp + FT_SCP), (int)lds8(sp + FT_SCP + 1));
            kp = sc > max(max(m0, m1), m2);
          } else {
            kp = true;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
              for (int dx = -1; dx <= 1; dx++) {
                if (dx == 0 && dy == 0) continue;
                int qv = (int)lds8(sp + dy * FT_SCP + dx);
                bool raw = false;
                if (kPass == 2) {
                  const int qx = lx + dx, qy = ly + dy;
                  const int fy = qy < 4 ? 0 : (qy > 35 ? 2 : 1);
                  const int fx = qx < kFx ? 0 : (qx > kFx + 31 ? 2 : 1);
                  raw = s_flag[fy * (kFT + 2) + j + fx] != 0;
                }
                if (!raw && qv < a.ini_th) qv = 0;
                kp = kp && sc > qv;
              }
          }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, kp);
        if (bal) {
          tile_any = true;
          int base = 0;
          if (lane == 0) base = atomicAdd(a.cand_count + f * L->nlevels + lvl, __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (kp) {
            const int idx = base + __popc(bal & lt);
            if (idx < g.cand_cap)
              a.cand[(long long)f * L->cand_total + g.cand_off + idx] = pack_pt(X0 - kFx + lx - kBand, Y0 - 4 + ly - kBand, sc);
          }
        }
      }
      if (tile_any) anymask |= 1 << j;
      if (kPass == 1 && kDbg) {  // parity introspection only: the score map S at minThFAST for the tile interior
        uint8_t* scp = a.dbg_score + (long long)f * L->slab_bytes + g.plane_off + (long long)kEdge * g.pitch + kPadX;
        for (int i = lane; i < 32 * 32; i += 32) {
          const int ly = (i >> 5) + 4, lx = (i & 31) + kFx;
          const int gx = X0 - kFx + lx, gy = Y0 - 4 + ly;
          if (gx < w - kEdge && gy < h - kEdge) scp[(long long)gy * g.pitch + gx] = s_sc[ly * FT_SCP + lx - FT_SCX];
        }
      }
      j = jn;
      if (FT_BUFS == 2) buf ^= 1;
    }
    if (kPass == 1) {
      // a tile with no pass-1 keypoint is retried at minThFAST (ORBextractor.cc:718-727); pass 2 re-runs only these runs
      const bool r = lane < ntile && !((anymask >> lane) & 1);
      if (lane < ntile) fretry[t0 + lane] = r ? 1 : 0;
      const uint32_t any_retry = __ballot_sync(0xffffffffu, r);
      if (any_retry && lane == 0) a.retry_list[1 + atomicAdd(a.retry_list, 1)] = fb;
    }
  }
}

}  // namespace swm
