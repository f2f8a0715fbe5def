// Place-recognition shard scan on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract as db_top2_mma_kernel (match.cu): for every query descriptor the two nearest
// database descriptors of a slice, as 64-bit keys (dist << 48 | global index), ties -> lower index.
// The reference computes these distances one pair at a time with ORBmatcher::DescriptorDistance
// (/root/reference/code/src/ORBmatcher.cc:1511-1525); the brute-force scan is the all-pairs form of it.
//
// Formulation.  For 256-bit descriptors q, x:  hamming(q, x) = popc(q) + popc(x) - 2 q.x.  With
//   A[m][k] = +64 if bit k of query m is 0, -64 if it is 1          (s8, k < 256)
//   B[n][k] = 2 * bit k of database row n                          (u8, k < 256)
// the int8 product sums to 128 (popc(x) - 2 q.x).  A ninth K block of 32 constants adds the rest:
//   A[m][256..260] = bytes summing to 2 popc(q_m), A[m][261] = 1;  B[n][256..260] = 64, B[n][261] = n
// so the accumulator is D[m][n] = 128 hamming(q_m, x_n) + n with n < 64 the row inside the B tile: a ready-made
// 16-bit sort key (distance << 7 | column).  The epilogue never computes a distance: it packs the keys of two
// queries (rows m and m + 128 of the CTA's 256) into one register, finds chunk minima with 16x2 SIMD min, runs
// the exact 16x2 top-2 insertion only for chunks that can still matter, and folds a tile's top-2 into a 32-bit
// (distance << 20 | index) top-2.
//
// Operands.  The descriptor bits of the two 128-query A tiles are written once into TMEM (tcgen05.st, 64 columns
// each) and K steps 0-7 take A from there (.ts MMAs): shared-memory bandwidth is spent on B only.  (With A in
// shared memory a 128x128x32 MMA reads 8 KB per 64 cycles, the full shared-memory rate; measured: tensor pipe
// 75 % with l1tex tc wavefronts at the same 75 %.)  The constant K block of A stays in shared memory (4 KB per A
// tile, K step 8 is an .ss MMA) so that TMEM holds 128 columns of A plus THREE accumulator stages of 128 columns
// (two 128 x 64 s32 tiles, one per A tile) = 512.  B tiles are 64 database rows expanded from bits to bytes by
// producer warps (K-major, no swizzle: 8-row x 16-byte core matrices); one tile = two 128x64x32 MMAs per K step.
//
// CTA = 19 warps.  0-11 epilogue in three groups of four (TMEM lane quarter = warp & 3): group g owns accumulator
// stage g = tiles i with i % 3 == g; it copies the tile to registers in two halves and releases the stage before
// doing the arithmetic of the second half.  12-14 MMA issuers, one per accumulator stage (the tensor pipe's queue
// is short: a single issuer's barrier waits between tiles drain it).  15-18 producers (15-16 even tiles, 17-18
// odd).  Pipelines: full/empty mbarriers over kStages B tiles, tmem_full/tmem_empty per accumulator stage.  With
// three stages the hand-back latency of one stage (commit -> epilogue wake-up -> TMEM loads -> arrive -> issuer
// wake-up) is covered by the MMAs of the other two.  One CTA per SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace umma {

constexpr int kQPerCta = 256;                 // two A tiles of 128 queries
constexpr int kTileN = 64;                    // database rows per B tile (= MMA N)
constexpr int kKSteps = 9;                    // 8 x 32 descriptor bits + the constant block
constexpr int kLbo = 128;                     // bytes between the two 16-byte K chunks of one MMA
constexpr int kSbo = 18 * 128;                // B: bytes between 8-row groups: 18 K chunks of 128 B
constexpr int kTileBytes = (kTileN / 8) * kSbo;
constexpr int kStages = 6;                    // B stages
constexpr int kAccStages = 3;
constexpr int kEpiWarps = 4 * kAccStages, kMmaWarp = kEpiWarps, kProdWarp0 = kMmaWarp + kAccStages;
constexpr int kThreads = (kProdWarp0 + 4) * 32;
constexpr int kAConstSbo = 256;               // A constant block: one K step = 2 chunks per 8-row group
constexpr int kAConstBytes = 16 * kAConstSbo; // per A tile
constexpr int kScratchBytes = (kAccStages - 1) * 128 * 16;  // epilogue groups 1.. -> group 0 hand-over
constexpr int kSmemBytes = kStages * kTileBytes + 2 * kAConstBytes + kScratchBytes + 256;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCols = 2 * kTileN;              // accumulator stage: A tile 0 | A tile 1
constexpr uint32_t kTmemA = kAccStages * kAccCols;     // A operands behind the accumulator stages
constexpr uint32_t kATileCols = 8 * 8;                 // 128 rows x 256 bytes = 64 columns
static_assert(kTmemA + 2 * kATileCols <= kTmemCols, "TMEM budget");
// instruction descriptor (kind::i8): D = s32, A = s8, B = u8, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (0u << 10) | ((kTileN >> 3) << 17) | ((128u >> 4) << 24);
// shared-memory matrix descriptors, high word: stride between 8-row groups, version 1 (sm_100), no swizzle
constexpr uint32_t kDescHiB = (uint32_t)(kSbo >> 4) | (1u << 14);
constexpr uint32_t kDescHiA = (uint32_t)(kAConstSbo >> 4) | (1u << 14);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Wait with a hardware suspend hint (the warp sleeps inside try_wait instead of spinning through the issue slots
// that producers and epilogue share).  Bounded: a pipeline bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(0x989680)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 4000000000ll) __trap();
}


// low word of a shared-memory matrix descriptor: start address and the K-chunk stride
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr) {
  return ((addr & 0x3FFFFu) >> 4) | ((uint32_t)(kLbo >> 4) << 16);
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, bool accumulate) {
  if (accumulate)
    asm volatile(
        "{\n.reg .b64 db;\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %4, p;\n}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(bdesc_lo), "r"(kDescHiB), "r"(kIdesc)
        : "memory");
  else
    asm volatile(
        "{\n.reg .b64 db;\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\nmov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %4, p;\n}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(bdesc_lo), "r"(kDescHiB), "r"(kIdesc)
        : "memory");
}
// D[tmem] += A[smem] . B[smem]
__device__ __forceinline__ void mma_i8_ss(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo) {
  asm volatile(
      "{\n.reg .b64 da, db;\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}" ::"r"(tmem_d),
      "r"(adesc_lo), "r"(bdesc_lo), "r"(kDescHiA), "r"(kDescHiB), "r"(kIdesc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

#define SWM_LD16_OUT(v)                                                                                              \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define SWM_LD16_INOUT(v)                                                                                            \
  "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),       \
      "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])

// 32 consecutive accumulator columns of this thread's TMEM lane, the low 16 bits of two adjacent columns packed
// into each register (asynchronous until tmem_wait).  Which column lands in which half does not matter here: a
// key carries its own column.
__device__ __forceinline__ void tmem_ld32p(uint32_t (&v)[16], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
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

// One chunk = 32 accumulator columns of one query, two 16-bit keys per register.  A 16x2 SIMD min tree finds the
// minimum of each half; only when some lane of the warp sees a key below its running second-best distance (thr, in
// both halves) is the exact top-2 insertion run (the two halves are independent streams, merged per tile).  Keys
// that are skipped have distance >= the second-best of an earlier (lower-index) entry: they can never be reported.
template <bool kMask>
__device__ __forceinline__ void consume32(uint32_t (&p)[16], int limit, uint32_t thr, uint32_t& k0, uint32_t& k1) {
  if (kMask) {  // last, partial tile: rows past the end of the database never win
#pragma unroll
    for (int i = 0; i < 16; i++) {
      uint32_t lo = p[i] & 0xFFFFu, hi = p[i] >> 16;
      if ((int)(lo & 0x7Fu) >= limit) lo = 0xFFFFu;
      if ((int)(hi & 0x7Fu) >= limit) hi = 0xFFFFu;
      p[i] = lo | (hi << 16);
    }
  }
  uint32_t m3[5];  // min tree of depth 3 (three-input SIMD min)
#pragma unroll
  for (int i = 0; i < 5; i++) m3[i] = __vminu2(__vminu2(p[3 * i], p[3 * i + 1]), p[3 * i + 2]);
  const uint32_t ma = __vminu2(__vminu2(m3[0], m3[1]), m3[2]), mb = __vminu2(__vminu2(m3[3], m3[4]), p[15]);
  const uint32_t m = __vminu2(ma, mb);
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

__device__ __forceinline__ void top2_insert32(uint32_t key, uint32_t& k0, uint32_t& k1) {
  k1 = min(k1, max(k0, key));
  k0 = min(k0, key);
}

// partial[(blockIdx.x * nq + query) * 2 + {0, 1}]: this slice's best / second-best key per query.
__global__ void __launch_bounds__(kThreads, 1)
db_top2_umma_kernel(const uint4* __restrict__ db, long long ndb, long long first_index, const uint4* __restrict__ q, int nq,
                    int tiles_per_cta, unsigned long long* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_b = smem;                                   // kStages B tiles
  uint8_t* s_ac = smem + kStages * kTileBytes;           // constant K block of the two A tiles
  uint4* s_scratch = reinterpret_cast<uint4*>(s_ac + 2 * kAConstBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_ac + 2 * kAConstBytes + kScratchBytes);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * kStages + 3 * kAccStages);
  const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + kStages);
  const uint32_t bar_tfull = smem_u32(s_bar + 2 * kStages), bar_tempty = smem_u32(s_bar + 2 * kStages + kAccStages);
  const uint32_t bar_token = smem_u32(s_bar + 2 * kStages + 2 * kAccStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total_tiles = (ndb + kTileN - 1) / kTileN;
  const long long tile0 = (long long)blockIdx.x * tiles_per_cta;
  const int ntiles = (int)min((long long)tiles_per_cta, total_tiles - tile0);
  const int qbase = blockIdx.y * kQPerCta;

  // ---- setup: barriers, TMEM, the A tiles and the constant K block of every B stage
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(bar_full + 8 * s, 2);   // one arrive per producer warp of the tile's pair
      mbar_init(bar_empty + 8 * s, 1);  // tcgen05.commit
    }
    for (int t = 0; t < kAccStages; t++) {
      mbar_init(bar_tfull + 8 * t, 1);   // tcgen05.commit
      mbar_init(bar_tempty + 8 * t, 4);  // one arrive per epilogue warp of the stage's group
      mbar_init(bar_token + 8 * t, 1);   // issue-order token, passed by the previous issuer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= kProdWarp0) {  // B constant block: bytes 0..4 = 64, byte 5 = row inside the tile
    const int n = (tid - kProdWarp0 * 32) & (kTileN - 1);
    for (int s = (tid - kProdWarp0 * 32) / kTileN; s < kStages; s += 2) {
      uint8_t* dst = s_b + s * kTileBytes + (n >> 3) * kSbo + (n & 7) * 16;
      *reinterpret_cast<uint4*>(dst + 16 * kLbo) = make_uint4(0x40404040u, 0x40u | ((uint32_t)n << 8), 0, 0);
      *reinterpret_cast<uint4*>(dst + 17 * kLbo) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;

  uint32_t r0[2] = {~0u, ~0u}, r1[2] = {~0u, ~0u};  // running 32-bit top-2 per half: dist << 20 | index in slice
  const int m = (warp & 3) * 32 + lane;               // epilogue: TMEM lane = query row inside each A tile
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  if (warp < 8) {
    // ---- this thread's query row: descriptor bits -> TMEM (lane m, 8 columns per K step, 4 consecutive k per
    // column), constant block -> shared memory
    const int atile = warp >> 2;
    const int row = qbase + 128 * atile + m;
    const bool live = row < nq;
    const uint32_t a_addr = lane_base + kTmemA + atile * kATileCols;
    int pq = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint4 w = live ? __ldg(q + 2 * (size_t)row + h) : make_uint4(0, 0, 0, 0);
      pq += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) v[c] = live ? (0x40404040u ^ (expand_nibble2(ww[k], 4 * c) << 6)) : 0u;
        tmem_st8(a_addr + 8 * (4 * h + k), v);
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
    uint8_t* dst = s_ac + atile * kAConstBytes + (m >> 3) * kAConstSbo + (m & 7) * 16;
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(e[0] | (e[1] << 8) | (e[2] << 16) | (e[3] << 24), e[4] | (live ? 0x100u : 0u), 0, 0);
    *reinterpret_cast<uint4*>(dst + kLbo) = make_uint4(0, 0, 0, 0);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp < kEpiWarps) {
    // ===== epilogue: thread = TMEM lane m = queries qbase + m (low half) and qbase + 128 + m (high half)
    const int grp = warp >> 2;
    const uint32_t taddr0 = lane_base + grp * kAccCols;  // A tile 0 (queries m); + kTileN columns: A tile 1
    uint32_t thr[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    for (int i = grp; i < ntiles; i += kAccStages) {
      mbar_wait(bar_tfull + 8 * grp, (i / kAccStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the whole tile (64 columns x 2 A tiles) goes to 64 registers, then the stage returns to its issuer
      uint32_t x[2][2][16];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        tmem_ld32p(x[h][0], taddr0 + h * kTileN);
        tmem_ld32p(x[h][1], taddr0 + h * kTileN + 32);
      }
      tmem_wait(x[0][0], x[0][1]);  // waits for all four loads; the second statement ties the rest for the compiler
      tmem_wait(x[1][0], x[1][1]);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * grp);
      const long long remain = ndb - (tile0 + i) * kTileN;
      const uint32_t tile_bits = (uint32_t)i << 6;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
        if (remain >= kTileN) {
          consume32<false>(x[h][0], kTileN, thr[h], k0, k1);
          consume32<false>(x[h][1], kTileN, thr[h], k0, k1);
        } else {
          consume32<true>(x[h][0], (int)remain, thr[h], k0, k1);
          consume32<true>(x[h][1], (int)remain, thr[h], k0, k1);
        }
        if (k0 != 0xFFFFFFFFu) {
          // fold the tile's 16-bit keys (dist << 7 | row in tile; two streams x top-2) into the running 32-bit keys
          const uint32_t c[4] = {k0 & 0xFFFFu, k0 >> 16, k1 & 0xFFFFu, k1 >> 16};
#pragma unroll
          for (int j = 0; j < 4; j++)
            top2_insert32(((c[j] & 0xFF80u) << 13) | (c[j] & 0x7Fu) | tile_bits, r0[h], r1[h]);
          // threshold: the running second-best distance << 7 in both halves (a later key must be strictly closer)
          thr[h] = min((r1[h] >> 20) << 7, 0xFFFFu) * 0x10001u;
        }
      }
    }
    // groups 1.. hand their result to group 0 through shared memory
    if (grp > 0) s_scratch[(grp - 1) * 128 + m] = make_uint4(r0[0], r1[0], r0[1], r1[1]);
  } else if (warp < kProdWarp0) {
    // ===== MMA issuers: warp kMmaWarp + g takes the tiles with i % 3 == g (accumulator stage g).  The whole warp
    // runs the loop so that descriptors and barrier addresses stay warp-uniform (uniform registers feed UTCIMMA
    // directly); one elected lane issues.
    const int grp = warp - kMmaWarp;
    const uint32_t a0 = tmem_base + kTmemA, a1 = a0 + kATileCols;
    const uint32_t d0 = tmem_base + grp * kAccCols, d1 = d0 + kTileN;
    const uint32_t b_base = smem_desc_lo(smem_u32(s_b));
    const uint32_t ac0 = smem_desc_lo(smem_u32(s_ac)), ac1 = smem_desc_lo(smem_u32(s_ac + kAConstBytes));
    for (int i = grp; i < ntiles; i += kAccStages) {
      const int s = i % kStages;
      mbar_wait(bar_tempty + 8 * grp, ((i / kAccStages) & 1) ^ 1);
      mbar_wait(bar_full + 8 * s, (i / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t b = b_base + (uint32_t)((s * kTileBytes) >> 4);
      // Issue order = tile order: the token arrives from the issuer of tile i - 1 once most of its MMAs are
      // queued.  (Free-running issuers interleave MMA by MMA, all three stages then complete together and sit in
      // their hand-back latency together.)  The operand waits above are already done when the token comes.
      if (i > 0) mbar_wait(bar_token + 8 * grp, ((i - (grp == 0 ? kAccStages : 0)) / kAccStages) & 1);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {  // one k-step = two 16-byte K chunks of B, 8 TMEM columns of A
          const uint32_t adv = (uint32_t)((2 * kLbo * ks) >> 4);
          mma_i8_ts(d0, a0 + 8 * ks, b + adv, ks > 0);
          mma_i8_ts(d1, a1 + 8 * ks, b + adv, ks > 0);
          if (ks == 6) mbar_arrive(bar_token + 8 * ((grp + 1) % kAccStages));
        }
        mma_i8_ss(d0, ac0, b + (uint32_t)((2 * kLbo * 8) >> 4));  // constant block: + 128 popc(q) + row in tile
        mma_i8_ss(d1, ac1, b + (uint32_t)((2 * kLbo * 8) >> 4));
        mma_commit(bar_empty + 8 * s);    // B stage free once these MMAs have read it
        mma_commit(bar_tfull + 8 * grp);  // accumulators complete
      }
      __syncwarp();
    }
  } else {
    // ===== producers: thread = database row n of the tile; two warps fill even tiles, two odd tiles
    const int n = (tid - kProdWarp0 * 32) & (kTileN - 1);
    const int first = (tid - kProdWarp0 * 32) / kTileN;
    uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
    {
      const long long row = (tile0 + first) * kTileN + n;
      if (first < ntiles && row < ndb) {
        w0 = __ldg(db + 2 * row);
        w1 = __ldg(db + 2 * row + 1);
      }
    }
    for (int i = first; i < ntiles; i += 2) {
      const int s = i % kStages;
      const uint4 c0 = w0, c1 = w1;
      {  // this thread's next tile: in flight while the current one is expanded
        const long long row = (tile0 + i + 2) * kTileN + n;
        const bool live = i + 2 < ntiles && row < ndb;
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
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
  if (warp < 4) {
#pragma unroll
    for (int g = 0; g < kAccStages - 1; g++) {
      const uint4 o = s_scratch[g * 128 + m];
      top2_insert32(o.x, r0[0], r1[0]);
      top2_insert32(o.y, r0[0], r1[0]);
      top2_insert32(o.z, r0[1], r1[1]);
      top2_insert32(o.w, r0[1], r1[1]);
    }
    const unsigned long long slice_base = (unsigned long long)(first_index + tile0 * kTileN);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int row = qbase + 128 * h + m;
      if (row >= nq) continue;
      auto widen = [&](uint32_t key) -> unsigned long long {
        const uint32_t d = key >> 20;
        if (d > 256) return ~0ull;
        return ((unsigned long long)d << 48) | (slice_base + (key & 0xFFFFFu));
      };
      partial[((size_t)blockIdx.x * nq + row) * 2] = widen(r0[h]);
      partial[((size_t)blockIdx.x * nq + row) * 2 + 1] = widen(r1[h]);
    }
  }
}

}  // namespace umma

// ------------------------------------------------------------------------------------------------
// Measured int8 peak of the tensor pipe (tcgen05 kind::i8): a pure MMA issue loop, no producers and no epilogue.
// One CTA per SM; one elected thread issues `iters` MMAs of the scan's own shapes back to back and commits once.
//   mode 0: 128 x 64 x 32, A from TMEM (.ts)   -- the shape db_top2_umma_kernel issues for K steps 0-7
//   mode 1: 128 x 256 x 32, A and B from shared memory (.ss) -- the largest single-CTA int8 MMA
// Operands are whatever the freshly allocated shared memory / TMEM holds (the values do not matter for the rate).
// The roofline of the shard scan is reported against this number instead of the nominal 4.5 POP/s.
// ------------------------------------------------------------------------------------------------
namespace umma {
constexpr int kPeakSmem = 32 * 1024;  // B at +0 (<= 18.4 KB), A at +24 KB (4 KB)
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int mode, int iters, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kPeakSmem / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0, 0x02020202u, 0);
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  if (warp == 0) {
    const uint32_t b_lo = smem_desc_lo(smem_u32(smem));
    const uint32_t a_lo = smem_desc_lo(smem_u32(smem + 24 * 1024));
    // N = 256 instruction descriptor for mode 1 (same fields as kIdesc)
    const uint32_t idesc256 = (2u << 4) | (1u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    if (elect_one()) {
      if (mode == 0) {
        // six accumulator tiles in rotation (the scan keeps 3 stages x 2 tiles in flight): back-to-back MMAs into ONE
        // accumulator would measure the accumulate latency, not the issue rate
        for (int i = 0; i < iters; i++) mma_i8_ts(tmem_base + (uint32_t)(i % 6) * kTileN, tmem_base + kTmemA, b_lo, i > 5);
      } else {
        for (int i = 0; i < iters; i++)
          asm volatile(
              "{\n.reg .b64 da, db;\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %4};\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}" ::"r"(tmem_base + (uint32_t)(i & 1) * 256u),
              "r"(a_lo), "r"(b_lo), "r"(kDescHiA), "r"(kDescHiA), "r"(idesc256)  // 256-byte 8-row groups: A 4 KB, B 8 KB
              : "memory");
      }
      mma_commit(smem_u32(&s_bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&s_bar), 0);
    if (tid == 0 && cycles) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}
}  // namespace umma
