// octree_core.cuh -- block-parallel DistributeOctTree (reference code/src/ORBextractor.cc:465-689,
// DivideNode :407-463), one CTA per (frame, level), all state in shared memory.
//
// The reference walks a std::list with push_front/erase; that order is reproduced in closed form:
//   * every node created after the initial ones is pushed to the FRONT, so at any time the list is
//     "non-initial nodes by descending creation sequence, then surviving initial nodes ascending";
//   * a breadth sweep (:534-591) splits every node holding >1 point, visiting them in list order;
//   * the "expand the biggest first" phase (:599-663) sorts (count, node) ascending and walks from
//     the back, stopping as soon as the list reaches N nodes -> a prefix of the order
//     (count desc, creation sequence desc), found with a prefix sum instead of a serial walk.
//   Ties of that sort are frozen as "later-created first" (the reference
//   compares heap addresses, which is not reproducible).
// Points never move: each carries the list position of its node; a split only re-labels them.
// The per-leaf winner (max response, first in raster order on ties, :667-686) is one atomicMax.
//
// The same source compiles as a serial host emulation when SWM_OCTREE_HOST is defined (each
// barrier-delimited phase is a data-parallel loop, so running the loops serially is one valid
// schedule).  tests/host_harness.cpp uses that to check the logic against the oracle on a CPU-only
// box; the shipped library only contains the device instantiation.
#pragma once
#include <stdint.h>

#ifdef SWM_OCTREE_HOST
#include <algorithm>
#include <cmath>
#define OT_FOR(i, n) for (int i = 0; i < (int)(n); ++i)
#define OT_SYNC() ((void)0)
#define OT_LEADER() (true)
#define OT_DEV
template <typename T>
static inline T ot_atomic_add(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
static inline void ot_atomic_max(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
static inline int ot_lroundf(float v) { return (int)std::lround(v); }
static inline int ot_ceil_half(int d) { return (int)std::ceil((float)d / 2); }
#else
#define OT_FOR(i, n) for (int i = threadIdx.x; i < (int)(n); i += blockDim.x)
#define OT_SYNC() __syncthreads()
#define OT_LEADER() (threadIdx.x == 0)
#define OT_DEV __device__ __forceinline__
static __device__ __forceinline__ uint32_t ot_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
static __device__ __forceinline__ int ot_atomic_add(int* p, int v) { return atomicAdd(p, v); }
static __device__ __forceinline__ void ot_atomic_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
static __device__ __forceinline__ int ot_lroundf(float v) { return (int)lroundf(v); }
static __device__ __forceinline__ int ot_ceil_half(int d) { return (int)ceilf((float)d / 2); }
#endif

namespace swm {

// Candidate / selected point packing: y:12 | x:12 | score:8 (ROI coordinates), so that the
// integer order of the packed word is raster order (y, then x).
OT_DEV uint32_t pack_pt(int x, int y, int score) { return ((uint32_t)y << 20) | ((uint32_t)x << 8) | (uint32_t)score; }
OT_DEV int pt_x(uint32_t p) { return (int)((p >> 8) & 0xFFF); }
OT_DEV int pt_y(uint32_t p) { return (int)(p >> 20); }
OT_DEV int pt_score(uint32_t p) { return (int)(p & 0xFF); }

constexpr int kOtMaxQuota = 2048;  // largest per-level quota any instantiation serves
constexpr int kOtMaxIni = 16;
constexpr int kOtNone = 0xFFFF;
// Live quadtree nodes never exceed max(quota + 3, 4 * n_ini); the state is a template on
// that bound so small quotas (the usual 1000/2000-feature extractors) leave room for several CTAs per SM.
constexpr int ot_max_live(int max_quota) { return max_quota + 64; }

struct OtNode {
  int16_t ulx, uly, brx, bry;
};

// Shared-memory working set of one (frame, level) quadtree.  Points and their labels live in global
// memory (L1/L2 resident, a few thousand entries).
template <int kMaxLive>
struct OtState {
  OtNode nodes[2][kMaxLive];
  uint32_t cnt[2][kMaxLive];      // points per node
  uint32_t ord[kMaxLive];         // old position -> processing index this round (kOtNone if not split-able)
  uint32_t cand[kMaxLive];        // processing index -> old position
  uint32_t ccnt[kMaxLive][4];     // per processing index: points per child; reused as per-leaf winner key
  uint32_t cbase[kMaxLive + 1];   // exclusive prefix of non-empty children over processing order
  uint32_t spos[kMaxLive + 1];    // old position -> new position if it survives un-split
  uint32_t scan_tmp[1024 + 32];
  int scalars[16];
};

enum { OT_L = 0, OT_E, OT_M, OT_K, OT_NEXP, OT_PHASE, OT_FINISH, OT_CUR, OT_NPTS, OT_TMP0, OT_TMP1 };

// Exclusive prefix sum of a[0..n) in place; total written to *total.  All threads participate.
OT_DEV void ot_exclusive_scan(uint32_t* a, int n, uint32_t* tmp, uint32_t* total) {
#ifdef SWM_OCTREE_HOST
  uint32_t run = 0;
  for (int i = 0; i < n; i++) { uint32_t v = a[i]; a[i] = run; run += v; }
  *total = run;
  (void)tmp;
#else
  const int T = blockDim.x, t = threadIdx.x;
  const int per = (n + T - 1) / T;
  const int b = t * per, e = min(b + per, n);
  uint32_t s = 0;
  for (int i = b; i < e; i++) s += a[i];
  tmp[t] = s;
  __syncthreads();
  // scan of the T partials by warp 0 (T <= 1024 -> 32 per lane)
  if (t < 32) {
    const int chunk = (T + 31) / 32;
    uint32_t loc = 0;
    for (int i = t * chunk; i < min((t + 1) * chunk, T); i++) loc += tmp[i];
    uint32_t inc = loc;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (t >= o) inc += v;
    }
    uint32_t run = inc - loc;
    for (int i = t * chunk; i < min((t + 1) * chunk, T); i++) { uint32_t v = tmp[i]; tmp[i] = run; run += v; }
    if (t == 31) *total = inc;
  }
  __syncthreads();
  uint32_t run = tmp[t];
  for (int i = b; i < e; i++) { uint32_t v = a[i]; a[i] = run; run += v; }
  __syncthreads();
#endif
}

// pts: packed candidates (ROI coords), n <= capacity of pnode.  Writes the selection (list order)
// to out[0..ret) and returns the count.  W,H = ROI extent (maxX-minX, maxY-minY), N = quota.
template <int kMaxLive>
OT_DEV int ot_distribute(OtState<kMaxLive>& S, const uint32_t* pts, uint16_t* pnode, uint8_t* pchild, int n, int W, int H, int N,
                         uint32_t* out, int out_cap) {
  int* sc = S.scalars;
  if (n == 0) return 0;
  const int n_ini = ot_lroundf((float)W / (float)H);  // :469
  if (n_ini < 1 || n_ini > kOtMaxIni) return 0;       // guarded by the host (SWM_E_INVALID)
  const float hX = (float)W / (float)n_ini;
  // ---- initial nodes (:479-511)
  OT_FOR(i, n_ini) {
    S.nodes[0][i].ulx = (int16_t)(int)(hX * (float)i);
    S.nodes[0][i].uly = 0;
    S.nodes[0][i].brx = (int16_t)(int)(hX * (float)(i + 1));
    S.nodes[0][i].bry = (int16_t)H;
    S.cnt[0][i] = 0;
  }
  OT_SYNC();
  OT_FOR(p, n) {
    const int ni = (int)((float)pt_x(pts[p]) / hX);  // :495
    pnode[p] = (uint16_t)ni;
    ot_atomic_add(&S.cnt[0][ni], 1u);
  }
  OT_SYNC();
  if (OT_LEADER()) {  // drop empty initial nodes, keep ascending order (:500-511)
    int L = 0;
    for (int i = 0; i < n_ini; i++) {
      S.spos[i] = (uint32_t)L;
      if (S.cnt[0][i]) {
        S.nodes[0][L] = S.nodes[0][i];
        S.cnt[0][L] = S.cnt[0][i];
        L++;
      }
    }
    sc[OT_L] = L;
    sc[OT_CUR] = 0;
    sc[OT_PHASE] = 0;
    sc[OT_FINISH] = 0;
  }
  OT_SYNC();
  OT_FOR(p, n) pnode[p] = (uint16_t)S.spos[pnode[p]];
  OT_SYNC();

  // ---- rounds
  for (int guard = 0; guard < 64; guard++) {
    const int cur = sc[OT_CUR], nxt = cur ^ 1;
    const int L = sc[OT_L];
    const int phase = sc[OT_PHASE];
    OtNode* nd = S.nodes[cur];
    uint32_t* cn = S.cnt[cur];
    // candidates = nodes with more than one point, numbered in list order
    OT_FOR(i, L) S.ord[i] = cn[i] > 1 ? 1u : 0u;
    OT_SYNC();
    uint32_t E_u;
    ot_exclusive_scan(S.ord, L, S.scan_tmp, (uint32_t*)&sc[OT_E]);
    OT_SYNC();
    E_u = (uint32_t)sc[OT_E];
    const int E = (int)E_u;
    if (E == 0) break;  // nothing can be split: size == prevSize (:595 / :660)
    OT_FOR(i, L) {
      if (cn[i] > 1) S.cand[S.ord[i]] = (uint32_t)i;
      else S.ord[i] = kOtNone;
    }
    OT_SYNC();
    if (phase == 1) {
      // processing order = (count desc, creation desc) == (count desc, list position asc); all
      // candidates were created in the previous round so position order is creation order reversed.
      OT_FOR(e, E) {
        const uint32_t pos = S.cand[e];
        const uint32_t key = (cn[pos] << 16) | (0xFFFFu - pos);
        int rank = 0;
        for (int k = 0; k < E; k++) {
          const uint32_t p2 = S.cand[k];
          const uint32_t key2 = (cn[p2] << 16) | (0xFFFFu - p2);
          rank += key2 > key;
        }
        S.ord[pos] = (uint32_t)rank;
      }
      OT_SYNC();
      OT_FOR(i, L) if (S.ord[i] != kOtNone) S.cand[S.ord[i]] = (uint32_t)i;
      OT_SYNC();
    }
    // per-candidate child populations (DivideNode :440-452)
    OT_FOR(e, E) { S.ccnt[e][0] = S.ccnt[e][1] = S.ccnt[e][2] = S.ccnt[e][3] = 0; }
    OT_SYNC();
    OT_FOR(p, n) {
      const uint32_t pos = pnode[p];
      const uint32_t j = S.ord[pos];
      if (j != kOtNone) {
        const OtNode q = nd[pos];
        const int cx = q.ulx + ot_ceil_half(q.brx - q.ulx);
        const int cy = q.uly + ot_ceil_half(q.bry - q.uly);
        const int x = pt_x(pts[p]), y = pt_y(pts[p]);
        const int c = x < cx ? (y < cy ? 0 : 2) : (y < cy ? 1 : 3);
        pchild[p] = (uint8_t)c;
        ot_atomic_add(&S.ccnt[j][c], 1u);
      }
    }
    OT_SYNC();
    // number of non-empty children per candidate, prefix over processing order
    OT_FOR(e, E) S.cbase[e] = (S.ccnt[e][0] > 0) + (S.ccnt[e][1] > 0) + (S.ccnt[e][2] > 0) + (S.ccnt[e][3] > 0);
    OT_SYNC();
    ot_exclusive_scan(S.cbase, E, S.scan_tmp, (uint32_t*)&sc[OT_TMP0]);
    OT_SYNC();
    if (OT_LEADER()) {
      S.cbase[E] = (uint32_t)sc[OT_TMP0];
      sc[OT_M] = E;
    }
    OT_SYNC();
    if (phase == 1) {
      // smallest prefix m with L + sum_{j<m}(nc_j - 1) >= N (:654-657); whole set if none.
      OT_FOR(e, E) {
        const int after = L + (int)S.cbase[e + 1] - (e + 1);
        const int before = L + (int)S.cbase[e] - e;
        if (after >= N && before < N) sc[OT_M] = e + 1;
      }
      OT_SYNC();
    }
    const int m = sc[OT_M];
    const int K = (int)S.cbase[m];
    // survivors keep their relative order behind the K new children
    OT_FOR(i, L) S.spos[i] = (S.ord[i] == kOtNone || (int)S.ord[i] >= m) ? 1u : 0u;
    OT_SYNC();
    ot_exclusive_scan(S.spos, L, S.scan_tmp, (uint32_t*)&sc[OT_TMP1]);
    OT_SYNC();
    const int Lnew = K + sc[OT_TMP1];
    OtNode* nn = S.nodes[nxt];
    uint32_t* cnn = S.cnt[nxt];
    OT_FOR(i, L) {
      const bool split = S.ord[i] != kOtNone && (int)S.ord[i] < m;
      if (!split) {
        const int np = K + (int)S.spos[i];
        nn[np] = nd[i];
        cnn[np] = cn[i];
        S.spos[i] = (uint32_t)np;
      }
    }
    OT_FOR(e, m) {
      const OtNode q = nd[S.cand[e]];
      const int hx = ot_ceil_half(q.brx - q.ulx), hy = ot_ceil_half(q.bry - q.uly);
      int rank = 0;
      for (int c = 0; c < 4; c++) {
        const uint32_t pc = S.ccnt[e][c];
        if (!pc) continue;
        const int created = (int)S.cbase[e] + rank;  // creation index within this round
        const int np = K - 1 - created;              // push_front => reversed
        OtNode ch;
        ch.ulx = (int16_t)((c & 1) ? q.ulx + hx : q.ulx);
        ch.brx = (int16_t)((c & 1) ? q.brx : q.ulx + hx);
        ch.uly = (int16_t)((c & 2) ? q.uly + hy : q.uly);
        ch.bry = (int16_t)((c & 2) ? q.bry : q.uly + hy);
        nn[np] = ch;
        cnn[np] = pc;
        S.ccnt[e][c] = (uint32_t)np | 0x80000000u;  // reuse as child -> new position
        rank++;
      }
    }
    OT_SYNC();
    OT_FOR(p, n) {
      const uint32_t pos = pnode[p];
      const uint32_t j = S.ord[pos];
      if (j != kOtNone && (int)j < m) pnode[p] = (uint16_t)(S.ccnt[j][pchild[p]] & 0xFFFFu);
      else pnode[p] = (uint16_t)S.spos[pos];
    }
    OT_SYNC();
    if (OT_LEADER()) sc[OT_NEXP] = 0;
    OT_SYNC();
    if (phase == 0) {
      OT_FOR(i, K) if (cnn[i] > 1) ot_atomic_add(&sc[OT_NEXP], 1);
      OT_SYNC();
    }
    if (OT_LEADER()) {
      sc[OT_CUR] = nxt;
      sc[OT_L] = Lnew;
      if (Lnew >= N || Lnew == L) sc[OT_FINISH] = 1;                      // :595 / :660
      else if (phase == 0 && Lnew + sc[OT_NEXP] * 3 > N) sc[OT_PHASE] = 1;  // :599
    }
    OT_SYNC();
    if (sc[OT_FINISH]) break;
  }
  OT_SYNC();
  // ---- best point per leaf (:667-686): max response, earliest in raster order on ties
  const int cur = sc[OT_CUR];
  const int L = sc[OT_L];
  (void)cur;
  uint32_t* best = &S.ccnt[0][0];  // the child counters are dead by now
  OT_FOR(i, L) best[i] = 0;
  OT_SYNC();
  OT_FOR(p, n) {
    const uint32_t v = pts[p];
    const uint32_t key = ((uint32_t)pt_score(v) << 24) | (0xFFFFFFu - (v >> 8));
    ot_atomic_max(&best[pnode[p]], key);
  }
  OT_SYNC();
  OT_FOR(i, L) {
    if (i < out_cap) {
      const uint32_t key = best[i];
      const uint32_t yx = 0xFFFFFFu - (key & 0xFFFFFFu);
      out[i] = (yx << 8) | (key >> 24);
    }
  }
  OT_SYNC();
  return L < out_cap ? L : out_cap;
}

}  // namespace swm
