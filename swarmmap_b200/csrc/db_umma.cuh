// Place-recognition shard scan on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract as db_top2_mma_kernel (match.cu): for every query descriptor the two nearest
// database descriptors of a slice, as 64-bit keys (dist << 48 | global index), ties -> lower index.
// The reference computes these distances one pair at a time with ORBmatcher::DescriptorDistance
// (/root/reference/code/src/ORBmatcher.cc:1845-1862); the brute-force scan is the all-pairs form of it.
//
// Formulation.  For 256-bit descriptors q, x:  hamming(q, x) = popc(q) + popc(x) - 2 q.x.  With
//   A[m][k] = +64 if bit k of query m is 0, -64 if it is 1          (s8, k < 256)
//   B[n][k] = 2 * bit k of database row n                          (u8, k < 256)
// the int8 product sums to 128 (popc(x) - 2 q.x).  A ninth K block of 32 constants adds the rest:
//   A[m][256..260] = bytes summing to 2 popc(q_m), A[m][261] = 1;  B[n][256..260] = 64, B[n][261] = n
// so the accumulator is D[m][n] = 128 hamming(q_m, x_n) + n  with n < 128 the column inside the tile:
// a ready-made 16-bit sort key (distance << 7 | column).  The epilogue never computes a distance: it
// packs the keys of two queries (rows m and m + 128 of the CTA's 256) into one register, finds chunk minima with
// 16x2 SIMD min, runs the exact 16x2 top-2 insertion only for chunks that can still matter, and folds the tile's
// top-2 into a 32-bit (distance << 20 | index) top-2 once per tile.
//
// CTA = 9 warps: 0-3 epilogue (TMEM lane quarter = warp id), 4 = MMA issuer, 5-8 = producers that
// expand database bits to bytes in shared memory (K-major, no swizzle: 8-row x 16-byte core matrices).
// Pipelines: full/empty mbarriers over kStages B tiles, tmem_full/tmem_empty over two accumulator
// stages of 256 TMEM columns (two 128 x 128 s32 tiles, one per A tile).  One CTA per SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace umma {

constexpr int kQPerCta = 256;                 // two A tiles of 128 queries
constexpr int kTileN = 128;                   // database rows per B tile (= MMA N)
constexpr int kKSteps = 9;                    // 8 x 32 descriptor bits + the constant block
constexpr int kLbo = 128;                     // bytes between the two 16-byte K chunks of one MMA
constexpr int kSbo = 18 * 128;                // bytes between 8-row groups: 18 K chunks of 128 B
constexpr int kTileBytes = 16 * kSbo;         // 128 rows
constexpr int kStages = 3;
constexpr int kThreads = 9 * 32;
constexpr int kSmemBytes = (2 + kStages) * kTileBytes + 256;
constexpr uint32_t kTmemCols = 512;
// instruction descriptor (kind::i8): D = s32, A = s8, B = u8, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (0u << 10) | ((kTileN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a pipeline bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  for (int spin = 0;; spin++) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 1023) == 1023 && clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | ((uint32_t)(kLbo >> 4) << 16);
  return ((uint64_t)((uint32_t)(kSbo >> 4) | (1u << 14)) << 32) | lo;
}

// Descriptors are passed as their low words; the high word (strides, version) is the same constant for all.
constexpr uint32_t kDescHi = (uint32_t)(kSbo >> 4) | (1u << 14);  // version 1 (sm_100), no swizzle, base offset 0
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, bool accumulate) {
  if (accumulate)
    asm volatile(
        "{\n.reg .b64 da, db;\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n}" ::"r"(tmem_d),
        "r"(adesc_lo), "r"(bdesc_lo), "r"(kDescHi), "r"(kIdesc)
        : "memory");
  else
    asm volatile(
        "{\n.reg .b64 da, db;\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n}" ::"r"(tmem_d),
        "r"(adesc_lo), "r"(bdesc_lo), "r"(kDescHi), "r"(kIdesc)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define SWM_LD16_OUT(v)                                                                                              \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define SWM_LD16_INOUT(v)                                                                                            \
  "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),       \
      "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])

// 16 consecutive accumulator columns of this thread's TMEM lane (asynchronous until tmem_wait).
__device__ __forceinline__ void tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : SWM_LD16_OUT(v)
      : "r"(taddr)
      : "memory");
}
// The loaded registers are in/out operands so that no use of them can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : SWM_LD16_INOUT(a), SWM_LD16_INOUT(b)::"memory");
}

// bits shift..shift+3 of w -> 4 bytes of value 0 / 2
__device__ __forceinline__ uint32_t expand_nibble2(uint32_t w, int shift) {
  return (((w >> shift) & 0xFu) * 0x00408102u) & 0x02020202u;
}

// One chunk = 16 accumulator columns of both A tiles.  The keys of rows m (low half) and m + 128 (high half) are
// packed into one register and a 16x2 SIMD min tree finds the chunk minimum; only when some lane of the warp sees
// a key below its running second-best distance (thr, packed per half) is the exact top-2 insertion run.  Keys that
// are skipped have distance >= the second-best of an earlier (lower-index) entry, so they can never be reported.
template <bool kMask>
__device__ __forceinline__ void consume16(const uint32_t (&a)[16], const uint32_t (&b)[16], int col0, int limit,
                                          uint32_t thr, uint32_t& k0, uint32_t& k1) {
  uint32_t p[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    p[i] = b[i] * 65536u + a[i];  // a < 65536: a multiply-add on the FMA pipe instead of a byte permute on the ALU
    if (kMask && col0 + i >= limit) p[i] = 0xFFFFFFFFu;  // last, partial tile: columns past the end never win
  }
  uint32_t m = __vminu2(p[0], p[1]);
#pragma unroll
  for (int i = 2; i < 16; i += 2) m = __vminu2(__vminu2(m, p[i]), p[i + 1]);
  if (__any_sync(0xFFFFFFFFu, __vminu2(m, thr) != thr)) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const uint32_t lo = __vminu2(p[i], p[i + 1]), hi = __vmaxu2(p[i], p[i + 1]);
      const uint32_t t = __vmaxu2(k0, lo);
      k1 = __vminu2(__vminu2(t, k1), hi);
      k0 = __vminu2(k0, lo);
    }
  }
}

template <bool kMask>
__device__ __forceinline__ void scan_tile(uint32_t taddr0, int limit, uint32_t thr, uint32_t& k0, uint32_t& k1) {
  // taddr0: accumulator of A tile 0 (queries m), +128 columns: A tile 1 (queries m + 128)
  uint32_t a0[16], a1[16], b0[16], b1[16];
  tmem_ld16(a0, taddr0);
  tmem_ld16(a1, taddr0 + 128);
#pragma unroll
  for (int c = 0; c < kTileN / 16; c += 2) {
    tmem_wait(a0, a1);
    tmem_ld16(b0, taddr0 + 16 * (c + 1));
    tmem_ld16(b1, taddr0 + 128 + 16 * (c + 1));
    consume16<kMask>(a0, a1, 16 * c, limit, thr, k0, k1);
    tmem_wait(b0, b1);
    if (c + 2 < kTileN / 16) {
      tmem_ld16(a0, taddr0 + 16 * (c + 2));
      tmem_ld16(a1, taddr0 + 128 + 16 * (c + 2));
    }
    consume16<kMask>(b0, b1, 16 * (c + 1), limit, thr, k0, k1);
  }
}

__device__ __forceinline__ void top2_insert32(uint32_t key, uint32_t& k0, uint32_t& k1) {
  k1 = min(k1, max(k0, key));
  k0 = min(k0, key);
}

// partial[(blockIdx.x * nq + query) * 2 + {0, 1}]: this slice's best / second-best key per query.
__global__ void __launch_bounds__(kThreads, 1)
db_top2_umma_kernel(const uint4* __restrict__ db, long long ndb, long long first_index, const uint4* __restrict__ q, int nq,
                    int tiles_per_cta, unsigned long long* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;                       // 2 A tiles
  uint8_t* s_b = smem + 2 * kTileBytes;      // kStages B tiles
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + (2 + kStages) * kTileBytes);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 16);
  const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + kStages);
  const uint32_t bar_tfull = smem_u32(s_bar + 2 * kStages), bar_tempty = smem_u32(s_bar + 2 * kStages + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total_tiles = (ndb + kTileN - 1) / kTileN;
  const long long tile0 = (long long)blockIdx.x * tiles_per_cta;
  const int ntiles = (int)min((long long)tiles_per_cta, total_tiles - tile0);
  const int qbase = blockIdx.y * kQPerCta;

  // ---- setup: barriers, TMEM, the A tiles and the constant K block of every B stage
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(bar_full + 8 * s, 4);   // one arrive per producer warp
      mbar_init(bar_empty + 8 * s, 1);  // tcgen05.commit
    }
    for (int t = 0; t < 2; t++) {
      mbar_init(bar_tfull + 8 * t, 1);   // tcgen05.commit
      mbar_init(bar_tempty + 8 * t, 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < kQPerCta) {  // one query row per thread
    const int row = qbase + tid;
    const bool live = row < nq;
    uint8_t* dst = s_a + (tid >> 7) * kTileBytes + ((tid & 127) >> 3) * kSbo + (tid & 7) * 16;
    int pq = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint4 w = live ? __ldg(q + 2 * (size_t)row + h) : make_uint4(0, 0, 0, 0);
      pq += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int c = 0; c < 2; c++) {  // K chunk 8 h + 2 k + c: bits 16 c .. 16 c + 15 of word k
          uint4 v;
          v.x = 0x40404040u ^ (expand_nibble2(ww[k], 16 * c) << 6);
          v.y = 0x40404040u ^ (expand_nibble2(ww[k], 16 * c + 4) << 6);
          v.z = 0x40404040u ^ (expand_nibble2(ww[k], 16 * c + 8) << 6);
          v.w = 0x40404040u ^ (expand_nibble2(ww[k], 16 * c + 12) << 6);
          if (!live) v = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(dst + (8 * h + 2 * k + c) * kLbo) = v;
        }
    }
    // constant block: bytes 0..4 sum to 2 popc(q) (each <= 127), byte 5 = 1
    int rem = live ? 2 * pq : 0;
    uint32_t e[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      e[i] = (uint32_t)min(rem, 127);
      rem -= (int)e[i];
    }
    *reinterpret_cast<uint4*>(dst + 16 * kLbo) =
        make_uint4(e[0] | (e[1] << 8) | (e[2] << 16) | (e[3] << 24), e[4] | (live ? 0x100u : 0u), 0, 0);
    *reinterpret_cast<uint4*>(dst + 17 * kLbo) = make_uint4(0, 0, 0, 0);
  }
  if (warp >= 5) {  // B constant block: bytes 0..4 = 64, byte 5 = column
    const int n = tid - 5 * 32;
    for (int s = 0; s < kStages; s++) {
      uint8_t* dst = s_b + s * kTileBytes + (n >> 3) * kSbo + (n & 7) * 16;
      *reinterpret_cast<uint4*>(dst + 16 * kLbo) = make_uint4(0x40404040u, 0x40u | ((uint32_t)n << 8), 0, 0);
      *reinterpret_cast<uint4*>(dst + 17 * kLbo) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;

  if (warp < 4) {
    // ===== epilogue: thread = TMEM lane m = queries qbase + m (low half) and qbase + 128 + m (high half)
    uint32_t r0[2] = {~0u, ~0u}, r1[2] = {~0u, ~0u};  // running 32-bit top-2 per half: dist << 20 | index in slice
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int i = 0; i < ntiles; i++) {
      const int t = i & 1;
      mbar_wait(bar_tfull + 8 * t, (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
      // packed per-half threshold: the running second-best distance << 7 (a later key must be strictly closer)
      const uint32_t thr = min((r1[0] >> 20) << 7, 0xFFFFu) | (min((r1[1] >> 20) << 7, 0xFFFFu) << 16);
      const long long remain = ndb - (tile0 + i) * kTileN;
      if (remain >= kTileN)
        scan_tile<false>(lane_addr + t * 256, kTileN, thr, k0, k1);
      else
        scan_tile<true>(lane_addr + t * 256, (int)remain, thr, k0, k1);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * t);
      // fold the tile's 16-bit keys (dist << 7 | column) into the running 32-bit keys
      const uint32_t tile_bits = (uint32_t)i << 7;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const uint32_t x0 = h ? (k0 >> 16) : (k0 & 0xFFFFu), x1 = h ? (k1 >> 16) : (k1 & 0xFFFFu);
        top2_insert32(((x0 & 0xFF80u) << 13) | (x0 & 0x7Fu) | tile_bits, r0[h], r1[h]);
        top2_insert32(((x1 & 0xFF80u) << 13) | (x1 & 0x7Fu) | tile_bits, r0[h], r1[h]);
      }
    }
    const unsigned long long slice_base = (unsigned long long)(first_index + tile0 * kTileN);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int row = qbase + 128 * h + tid;
      if (row >= nq) continue;
      auto widen = [&](uint32_t key) -> unsigned long long {
        const uint32_t d = key >> 20;
        if (d > 256) return ~0ull;
        return ((unsigned long long)d << 48) | (slice_base + (key & 0xFFFFFu));
      };
      partial[((size_t)blockIdx.x * nq + row) * 2] = widen(r0[h]);
      partial[((size_t)blockIdx.x * nq + row) * 2 + 1] = widen(r1[h]);
    }
  } else if (warp == 4) {
    // ===== MMA issuer.  The whole warp runs the loop so that descriptors and barrier addresses stay warp-uniform
    // (uniform registers feed UTCIMMA directly); one elected lane issues.
    const uint32_t a0 = (uint32_t)smem_desc(smem_u32(s_a)), a1 = (uint32_t)smem_desc(smem_u32(s_a + kTileBytes));
    for (int i = 0; i < ntiles; i++) {
      const int s = i % kStages, t = i & 1;
      mbar_wait(bar_tempty + 8 * t, ((i >> 1) & 1) ^ 1);
      mbar_wait(bar_full + 8 * s, (i / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t b = (uint32_t)smem_desc(smem_u32(s_b + s * kTileBytes));
      const uint32_t d0 = tmem_base + t * 256, d1 = d0 + 128;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kKSteps; ks++) {  // one k-step = two 16-byte K chunks = 2 kLbo bytes
          const uint32_t adv = (uint32_t)((2 * kLbo * ks) >> 4);
          mma_i8(d0, a0 + adv, b + adv, ks > 0);
          mma_i8(d1, a1 + adv, b + adv, ks > 0);
        }
        mma_commit(bar_empty + 8 * s);  // B stage free once these MMAs have read it
        mma_commit(bar_tfull + 8 * t);  // accumulators complete
      }
      __syncwarp();
    }
  } else {
    // ===== producers: thread = database row n of the tile
    const int n = tid - 5 * 32;
    const int pwarp_lane0 = lane == 0;
    uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
    {
      const long long row = tile0 * kTileN + n;
      if (ntiles > 0 && row < ndb) {
        w0 = __ldg(db + 2 * row);
        w1 = __ldg(db + 2 * row + 1);
      }
    }
    for (int i = 0; i < ntiles; i++) {
      const int s = i % kStages;
      const uint4 c0 = w0, c1 = w1;
      {  // next tile's bits: in flight while this one is expanded
        const long long row = (tile0 + i + 1) * kTileN + n;
        const bool live = i + 1 < ntiles && row < ndb;
        w0 = live ? __ldg(db + 2 * row) : make_uint4(0, 0, 0, 0);
        w1 = live ? __ldg(db + 2 * row + 1) : make_uint4(0, 0, 0, 0);
      }
      mbar_wait(bar_empty + 8 * s, ((i / kStages) & 1) ^ 1);
      uint8_t* dst = s_b + s * kTileBytes + (n >> 3) * kSbo + (n & 7) * 16;
      const uint32_t ww[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int k = 0; k < 8; k++)
#pragma unroll
        for (int c = 0; c < 2; c++)
          *reinterpret_cast<uint4*>(dst + (2 * k + c) * kLbo) =
              make_uint4(expand_nibble2(ww[k], 16 * c), expand_nibble2(ww[k], 16 * c + 4),
                         expand_nibble2(ww[k], 16 * c + 8), expand_nibble2(ww[k], 16 * c + 12));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (pwarp_lane0) mbar_arrive(bar_full + 8 * s);
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace umma
