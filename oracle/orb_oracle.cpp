// orb_oracle.cpp -- CPU ORACLE for the ORB extractor path.  TEST INFRASTRUCTURE ONLY
// (see orb_oracle.h for the parity status).  Every function cites the reference
// file:line it restates; paths are relative to /root/reference/code/.
//
// Build: make -C oracle   (g++ -std=c++17 -O2 -ffp-contract=off, no third-party deps)
#include "orb_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <list>
#include <utility>
#include <vector>

#include "orb_pattern_data.inc"

namespace {

constexpr int kPatchSize = 31;      // ORBextractor.cc:76
constexpr int kHalfPatch = 15;      // ORBextractor.cc:77
constexpr int kEdge = 19;           // ORBextractor.cc:78
constexpr int kFastBand = kEdge - 3;  // minBorderX, ORBextractor.cc:695
constexpr int kMaxFast = 10000;     // Fast.hpp:30 (GpuFast maxKeypoints default)

// cvRound / cvFloor semantics of OpenCV on x86: round-half-to-even via cvtss/cvtsd.
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_floor(float v) {
  int i = (int)v;
  return i - (i > v);
}
inline short sat_short(int v) { return (short)std::min(std::max(v, -32768), 32767); }

int g_umax[kHalfPatch + 1];
bool g_umax_ready = false;

// ORBextractor.cc:386-401
void build_umax() {
  if (g_umax_ready) return;
  int v, v0;
  const int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
  const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
  const double hp2 = kHalfPatch * kHalfPatch;
  for (v = 0; v <= vmax; ++v) g_umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (g_umax[v0] == g_umax[v0 + 1]) ++v0;
    g_umax[v] = v0;
    ++v0;
  }
  g_umax_ready = true;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// Tables
// ---------------------------------------------------------------------------------------------

// ORBextractor.cc:346-362.  scaleFactor is a double member initialised from a float
// (ORBextractor.h:108), so each step is float*double rounded back to float.
void orc_scale_tables(float scale_factor, int nlevels, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2) {
  const double s = scale_factor;
  std::vector<float> a(nlevels), b(nlevels);
  a[0] = 1.0f;
  b[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    a[i] = (float)(a[i - 1] * s);
    b[i] = a[i] * a[i];
  }
  for (int i = 0; i < nlevels; i++) {
    if (sf) sf[i] = a[i];
    if (sigma2) sigma2[i] = b[i];
    if (inv_sf) inv_sf[i] = 1.0f / a[i];
    if (inv_sigma2) inv_sigma2[i] = 1.0f / b[i];
  }
}

// ORBextractor.cc:367-378
void orc_level_quotas(int nfeatures, float scale_factor, int nlevels, int* quotas) {
  const double s = scale_factor;
  float factor = (float)(1.0f / s);
  float per_scale = (float)(nfeatures * (1.0f - factor) / (1.0f - std::pow((double)factor, nlevels)));
  int sum = 0;
  for (int level = 0; level < nlevels - 1; level++) {
    quotas[level] = cv_round(per_scale);
    sum += quotas[level];
    per_scale *= factor;
  }
  quotas[nlevels - 1] = std::max(nfeatures - sum, 0);
}

void orc_umax(int* umax16) {
  build_umax();
  std::memcpy(umax16, g_umax, sizeof(g_umax));
}

// ORBextractor.cc:825-826 / 841-842
void orc_level_sizes(int w, int h, float scale_factor, int nlevels, int* ws, int* hs) {
  std::vector<float> inv(nlevels);
  orc_scale_tables(scale_factor, nlevels, nullptr, inv.data(), nullptr, nullptr);
  for (int l = 0; l < nlevels; l++) {
    ws[l] = cv_round((float)w * inv[l]);
    hs[l] = cv_round((float)h * inv[l]);
  }
}

// ---------------------------------------------------------------------------------------------
// cv::resize(INTER_LINEAR) for CV_8UC1, the semantics north_star pins for ComputePyramid
// (call site ORBextractor.cc:845).  The arithmetic lives in OpenCV (imgproc/src/resize.cpp,
// un-vendored); this is a restatement of its published fixed-point algorithm:
//   11-bit coefficients, horizontal pass in int32, vertical pass
//   ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.
// Pinned bit-for-bit against cv2 4.13 in tests/test_oracle_cv2.py.
// ---------------------------------------------------------------------------------------------
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                          int dstride) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> alpha(2 * dw), beta(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    alpha[2 * dx] = sat_short(cv_round((1.f - fx) * 2048));
    alpha[2 * dx + 1] = sat_short(cv_round(fx * 2048));
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    beta[2 * dy] = sat_short(cv_round((1.f - fy) * 2048));
    beta[2 * dy + 1] = sat_short(cv_round(fy * 2048));
  }
  std::vector<int> r0(dw), r1(dw);
  auto hpass = [&](int sy, std::vector<int>& row) {
    sy = std::min(std::max(sy, 0), sh - 1);
    const uint8_t* S = src + (size_t)sy * sstride;
    for (int dx = 0; dx < dw; dx++) {
      const int sx = xofs[dx];
      const int sx1 = std::min(sx + 1, sw - 1);
      row[dx] = S[sx] * alpha[2 * dx] + S[sx1] * alpha[2 * dx + 1];
    }
  };
  for (int dy = 0; dy < dh; dy++) {
    hpass(yofs[dy], r0);
    hpass(yofs[dy] + 1, r1);
    const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
    uint8_t* D = dst + (size_t)dy * dstride;
    for (int dx = 0; dx < dw; dx++)
      D[dx] = (uint8_t)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

// cv::copyMakeBorder(BORDER_REFLECT_101) (call sites ORBextractor.cc:846-851).
void orc_border_reflect101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int border) {
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
  };
  for (int y = -border; y < h + border; y++) {
    const uint8_t* S = src + (size_t)refl(y, h) * sstride;
    uint8_t* D = dst + (size_t)(y + border) * dstride;
    for (int x = -border; x < w + border; x++) D[x + border] = S[refl(x, w)];
  }
}

// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) for CV_8U (filter created at ORBextractor.cc:835,
// applied on the level ROI at :719,:742).  OpenCV's 8-bit path is fixed point: the normalised
// kernel is quantised to 8 fractional bits with error diffusion -> {18,34,48,56,48,34,18}/256, the
// row pass keeps 8.8 values, the column pass 16.16, result = (acc + 0x8000) >> 16.
// Pinned bit-for-bit against cv2 4.13 in tests/test_oracle_cv2.py.
void orc_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  auto refl = [](int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
  };
  std::vector<uint16_t> tmp((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src + (size_t)y * sstride;
    for (int x = 0; x < w; x++) {
      int acc = 0;
      for (int k = 0; k < 7; k++) acc += K[k] * S[refl(x + k - 3, w)];
      tmp[(size_t)y * w + x] = (uint16_t)acc;
    }
  }
  for (int y = 0; y < h; y++) {
    uint8_t* D = dst + (size_t)y * dstride;
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int k = 0; k < 7; k++) acc += (uint32_t)K[k] * tmp[(size_t)refl(y + k - 3, h) * w + x];
      D[x] = (uint8_t)((acc + 0x8000u) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FAST-9/16 score (Fast_gpu.cu:63-68 diffType, :70-185 calcMask, :190-219 cornerScore,
// :221-266 isKeyPoint2).  A pixel is a corner at threshold t iff 9 contiguous ring pixels are all
// > v+t or all < v-t; its score is the largest such t (the binary search at :196-219 finds exactly
// that).  Equivalently score = max over the 16 arcs of min |diff| on the arc, minus 1.
// ---------------------------------------------------------------------------------------------
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, int min_th, uint8_t* score, int score_stride) {
  static const int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  for (int y = 0; y < h; y++) std::memset(score + (size_t)y * score_stride, 0, w);
  for (int y = 3; y < h - 3; y++) {
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* c = img + (size_t)y * stride + x;
      const int v = c[0];
      // Early-outs of calcMask (Fast_gpu.cu:82-125): every 9-arc contains one pixel of each
      // opposite pair, so a pair with both |diff| <= th rules the pixel out (exact, not a heuristic).
      bool reject = false;
      for (int k = 0; k < 8 && !reject; k += 2) {
        const int a = (int)c[RY[k] * stride + RX[k]] - v, b = (int)c[RY[k + 8] * stride + RX[k + 8]] - v;
        reject = (a <= min_th && a >= -min_th) && (b <= min_th && b >= -min_th);
      }
      if (reject) continue;
      int d[25];
      for (int k = 0; k < 16; k++) d[k] = (int)c[RY[k] * stride + RX[k]] - v;
      for (int k = 16; k < 25; k++) d[k] = d[k - 16];
      int best = 0;  // max over arcs of min(+d) and min(-d)
      for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int j = 1; j < 9; j++) {
          mn = std::min(mn, d[s + j]);
          mx = std::max(mx, d[s + j]);
        }
        best = std::max(best, std::max(mn, -mx));
      }
      // corner at min_th <=> best > min_th ; score = best - 1
      score[(size_t)y * score_stride + x] = (uint8_t)(best > min_th ? best - 1 : 0);
    }
  }
}

// Fast_gpu.cu:284-341 under the deterministic lock-step definition.
int orc_fast_tile_select(const uint8_t* score, int w, int h, int score_stride, int ini_th, orc_fastpt* out, int cap,
                         uint8_t* tile_retry) {
  const int tx = (w + 31) / 32, ty = (h + 31) / 32;  // Fast_gpu.cu:375
  auto S = [&](int x, int y) -> int {
    if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
    return score[(size_t)y * score_stride + x];
  };
  auto Shi = [&](int x, int y) -> int {
    const int s = S(x, y);
    return s >= ini_th ? s : 0;
  };
  std::vector<uint8_t> retry((size_t)tx * ty, 1);
  std::vector<uint8_t> kp1((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const int s = Shi(x, y);
      if (!s) continue;
      bool mx = true;
      for (int dy = -1; dy <= 1 && mx; dy++)
        for (int dx = -1; dx <= 1; dx++)
          if ((dx || dy) && !(s > Shi(x + dx, y + dy))) { mx = false; break; }
      if (mx) {
        kp1[(size_t)y * w + x] = 1;
        retry[(size_t)((y - 3) >> 5) * tx + ((x - 3) >> 5)] = 0;
      }
    }
  auto Seff = [&](int x, int y) -> int {
    if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
    return retry[(size_t)((y - 3) >> 5) * tx + ((x - 3) >> 5)] ? S(x, y) : Shi(x, y);
  };
  int n = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      bool is_kp = kp1[(size_t)y * w + x];
      if (!is_kp && retry[(size_t)((y - 3) >> 5) * tx + ((x - 3) >> 5)]) {
        const int s = S(x, y);
        if (s) {
          is_kp = true;
          for (int dy = -1; dy <= 1 && is_kp; dy++)
            for (int dx = -1; dx <= 1; dx++)
              if ((dx || dy) && !(s > Seff(x + dx, y + dy))) { is_kp = false; break; }
        }
      }
      if (is_kp) {
        if (n < cap) {
          out[n].x = (int16_t)x;
          out[n].y = (int16_t)y;
          out[n].score = S(x, y);
        }
        n++;
      }
    }
  if (tile_retry) std::memcpy(tile_retry, retry.data(), retry.size());
  return std::min(n, cap);
}

// ---------------------------------------------------------------------------------------------
// Quadtree distribution: ORBextractor.cc:407-463 (DivideNode) and :465-689 (DistributeOctTree).
// std::list order, push_front of non-empty children and the mid-sweep break are kept literally.
// Tie-break of the (count, node*) sort (:607-610) is frozen as (count, creation sequence) -- the
// reference's heap-address order is not reproducible.
// ---------------------------------------------------------------------------------------------
namespace {
struct QNode {
  int ulx, uly, brx, bry;  // UL.x, UL.y, BR.x (== UR.x), BR.y (== BL.y)
  std::vector<int> pts;    // indices into the input array, kept in input order
  bool no_more = false;
  int seq = 0;
  std::list<int>::iterator self;
};

struct QTree {
  const orc_fastpt* in;
  std::deque<QNode> nodes;   // deque: references stay valid while nodes are appended
  std::list<int> order;      // the reference's lNodes
  int next_seq = 0;

  int make(int ulx, int uly, int brx, int bry) {
    nodes.emplace_back();
    QNode& n = nodes.back();
    n.ulx = ulx; n.uly = uly; n.brx = brx; n.bry = bry;
    n.seq = next_seq++;
    return (int)nodes.size() - 1;
  }
  // DivideNode, :407-463
  void divide(int id, int child[4]) {
    const int ulx = nodes[id].ulx, uly = nodes[id].uly, brx = nodes[id].brx, bry = nodes[id].bry;
    const int hx = (int)std::ceil((float)(brx - ulx) / 2), hy = (int)std::ceil((float)(bry - uly) / 2);
    child[0] = make(ulx, uly, ulx + hx, uly + hy);
    child[1] = make(ulx + hx, uly, brx, uly + hy);
    child[2] = make(ulx, uly + hy, ulx + hx, bry);
    child[3] = make(ulx + hx, uly + hy, brx, bry);
    const float cx = (float)(ulx + hx), cy = (float)(uly + hy);
    for (int p : nodes[id].pts) {
      const float px = in[p].x, py = in[p].y;
      int c = px < cx ? (py < cy ? 0 : 2) : (py < cy ? 1 : 3);
      nodes[child[c]].pts.push_back(p);
    }
    for (int c = 0; c < 4; c++)
      if (nodes[child[c]].pts.size() == 1) nodes[child[c]].no_more = true;
  }
};
}  // namespace

int orc_octree(const orc_fastpt* pts, int n, int min_x, int max_x, int min_y, int max_y, int target,
               orc_fastpt* out, int out_cap) {
  const int n_ini = (int)std::round((float)(max_x - min_x) / (max_y - min_y));  // :469
  if (n_ini < 1) return 0;  // reference would divide by zero; portrait images are outside its envelope
  const float hX = (float)(max_x - min_x) / n_ini;
  QTree t;
  t.in = pts;
  std::vector<int> ini(n_ini);
  for (int i = 0; i < n_ini; i++) {
    const int id = t.make((int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), max_y - min_y);  // :481-484
    t.order.push_back(id);
    ini[i] = id;
  }
  for (int i = 0; i < n; i++) t.nodes[ini[(int)(pts[i].x / hX)]].pts.push_back(i);  // :495
  for (auto it = t.order.begin(); it != t.order.end();) {                             // :500-511
    QNode& nd = t.nodes[*it];
    if (nd.pts.size() == 1) { nd.no_more = true; ++it; }
    else if (nd.pts.empty()) it = t.order.erase(it);
    else ++it;
  }
  bool finish = false;
  std::vector<std::pair<int, int>> expand;  // (count, node id); sort key (count, seq)
  auto push_children = [&](int child[4], int* n_expand) {
    for (int c = 0; c < 4; c++) {
      QNode& ch = t.nodes[child[c]];
      if (ch.pts.empty()) continue;
      t.order.push_front(child[c]);
      ch.self = t.order.begin();
      if (ch.pts.size() > 1) {
        if (n_expand) (*n_expand)++;
        expand.emplace_back((int)ch.pts.size(), child[c]);
      }
    }
  };
  auto by_count_then_seq = [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
    if (a.first != b.first) return a.first < b.first;
    return t.nodes[a.second].seq < t.nodes[b.second].seq;
  };
  while (!finish) {
    const int prev = (int)t.order.size();
    int n_expand = 0;
    expand.clear();
    for (auto it = t.order.begin(); it != t.order.end();) {  // :534-591
      if (t.nodes[*it].no_more) { ++it; continue; }
      int child[4];
      t.divide(*it, child);
      push_children(child, &n_expand);
      it = t.order.erase(it);
    }
    if ((int)t.order.size() >= target || (int)t.order.size() == prev) {
      finish = true;
    } else if ((int)t.order.size() + n_expand * 3 > target) {  // :599
      while (!finish) {
        const int prev2 = (int)t.order.size();
        std::vector<std::pair<int, int>> cur = expand;
        expand.clear();
        std::sort(cur.begin(), cur.end(), by_count_then_seq);
        for (int j = (int)cur.size() - 1; j >= 0; j--) {
          int child[4];
          t.divide(cur[j].second, child);
          push_children(child, nullptr);
          t.order.erase(t.nodes[cur[j].second].self);
          if ((int)t.order.size() >= target) break;
        }
        if ((int)t.order.size() >= target || (int)t.order.size() == prev2) finish = true;
      }
    }
  }
  int k = 0;  // :667-686 best response per node, first wins on ties
  for (int id : t.order) {
    const QNode& nd = t.nodes[id];
    int best = nd.pts[0];
    for (size_t j = 1; j < nd.pts.size(); j++)
      if ((float)pts[nd.pts[j]].score > (float)pts[best].score) best = nd.pts[j];
    if (k < out_cap) out[k] = pts[best];
    k++;
  }
  return std::min(k, out_cap);
}

// ---------------------------------------------------------------------------------------------
// Orientation (Fast_gpu.cu:403-460) and rBRIEF (Orb_gpu.cu:63-100)
// ---------------------------------------------------------------------------------------------
void orc_ic_moments(const uint8_t* img, int stride, int x, int y, int* m10_out, int* m01_out) {
  build_umax();
  const uint8_t* c = img + (size_t)y * stride + x;
  int m01 = 0, m10 = 0;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int v_sum = 0;
    const int d = g_umax[v];
    for (int u = -d; u <= d; ++u) {
      const int p = c[u + v * stride], m = c[u - v * stride];
      v_sum += p - m;
      m10 += u * (p + m);
    }
    m01 += v * v_sum;
  }
  *m10_out = m10;
  *m01_out = m01;
}

float orc_ic_angle(const uint8_t* img, int stride, int x, int y) {
  int m10, m01;
  orc_ic_moments(img, stride, x, y, &m10, &m01);
  const float pi = 3.14159265f;  // CV_PI_F
  float a = atan2f((float)m01, (float)m10);
  a += (a < 0) * (2.0f * pi);
  a *= 180.0f / pi;
  return a;
}

void orc_rbrief(const uint8_t* img, int stride, int x, int y, float angle_deg, uint8_t* desc32) {
  const float factor_pi = (float)(3.1415926535897932384626433832795 / 180.f);
  const float ang = angle_deg * factor_pi;
  const float a = cosf(ang), b = sinf(ang);
  const uint8_t* c = img + (size_t)y * stride + x;
  auto sample = [&](int idx) -> int {
    const float px = kOracleOrbPattern[2 * idx], py = kOracleOrbPattern[2 * idx + 1];
    const int ry = (int)lrintf(px * b + py * a);
    const int rx = (int)lrintf(px * a - py * b);
    return c[ry * stride + rx];
  };
  for (int t = 0; t < 32; t++) {
    int val = 0;
    for (int k = 0; k < 8; k++) val |= (sample(16 * t + 2 * k) < sample(16 * t + 2 * k + 1)) << k;
    desc32[t] = (uint8_t)val;
  }
}

// ORBmatcher.cc:1511-1525 (the SWAR popcount is just popcount)
int orc_hamming256(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// ---------------------------------------------------------------------------------------------
// Full extractor: ORBextractor.cc:746-819 (operator()), :821-855 (ComputePyramid), :691-744
// ---------------------------------------------------------------------------------------------
struct orc_extractor {
  int nfeatures, nlevels, ini_th, min_th;
  float scale_factor;
  std::vector<float> sf, inv_sf;
  std::vector<int> quotas;
  int w = 0, h = 0;
  std::vector<int> lw, lh;
  std::vector<std::vector<uint8_t>> bordered, blurred, score;
  std::vector<std::vector<orc_fastpt>> fast, selected;
};

orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
  if (nfeatures <= 0 || nlevels <= 0 || nlevels > 32 || !(scale_factor > 1.0f)) return nullptr;
  orc_extractor* e = new orc_extractor;
  e->nfeatures = nfeatures;
  e->nlevels = nlevels;
  e->ini_th = ini_th;
  e->min_th = min_th;
  e->scale_factor = scale_factor;
  e->sf.resize(nlevels);
  e->inv_sf.resize(nlevels);
  e->quotas.resize(nlevels);
  orc_scale_tables(scale_factor, nlevels, e->sf.data(), e->inv_sf.data(), nullptr, nullptr);
  orc_level_quotas(nfeatures, scale_factor, nlevels, e->quotas.data());
  build_umax();
  return e;
}

void orc_extractor_destroy(orc_extractor* e) { delete e; }

int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, int stride, orc_keypoint* kps, uint8_t* desc,
                int cap) {
  if (!e) return -1;
  if (!img || w <= 0 || h <= 0) return 0;  // :750-751 silent return on empty image
  const int L = e->nlevels;
  e->w = w;
  e->h = h;
  e->lw.resize(L);
  e->lh.resize(L);
  orc_level_sizes(w, h, e->scale_factor, L, e->lw.data(), e->lh.data());
  if (e->lw[L - 1] < 2 * kEdge + 1 || e->lh[L - 1] < 2 * kEdge + 1) return -1;
  e->bordered.assign(L, {});
  e->blurred.assign(L, {});
  e->score.assign(L, {});
  e->fast.assign(L, {});
  e->selected.assign(L, {});
  // ComputePyramid (:839-853): level l is resized from the un-blurred ROI of level l-1.
  std::vector<std::vector<uint8_t>> plain(L);
  for (int l = 0; l < L; l++) {
    const int lw = e->lw[l], lh = e->lh[l];
    plain[l].resize((size_t)lw * lh);
    if (l == 0) {
      for (int y = 0; y < h; y++) std::memcpy(&plain[0][(size_t)y * w], img + (size_t)y * stride, w);
    } else {
      orc_resize_linear_u8(plain[l - 1].data(), e->lw[l - 1], e->lh[l - 1], e->lw[l - 1], plain[l].data(), lw, lh,
                           lw);
    }
    const int bw = lw + 2 * kEdge;
    e->bordered[l].resize((size_t)bw * (lh + 2 * kEdge));
    orc_border_reflect101(plain[l].data(), lw, lh, lw, e->bordered[l].data(), bw, kEdge);
  }
  int total = 0;
  std::vector<std::vector<orc_keypoint>> all(L);
  for (int l = 0; l < L; l++) {
    const int lw = e->lw[l], lh = e->lh[l];
    // FAST on the ROI [16, w-16) x [16, h-16) (:695-713)
    const int rw = lw - 2 * kFastBand, rh = lh - 2 * kFastBand;
    e->score[l].assign((size_t)lw * lh, 0);
    uint8_t* roi_score = e->score[l].data() + (size_t)kFastBand * lw + kFastBand;
    orc_fast_score_map(plain[l].data() + (size_t)kFastBand * lw + kFastBand, rw, rh, lw, e->min_th, roi_score, lw);
    e->fast[l].resize(kMaxFast);
    const int nf = orc_fast_tile_select(roi_score, rw, rh, lw, e->ini_th, e->fast[l].data(), kMaxFast, nullptr);
    e->fast[l].resize(nf);
    // DistributeOctTree (:726)
    e->selected[l].resize((size_t)nf + 8);
    const int ns = orc_octree(e->fast[l].data(), nf, kFastBand, lw - kFastBand, kFastBand, lh - kFastBand,
                              e->quotas[l], e->selected[l].data(), (int)e->selected[l].size());
    e->selected[l].resize(ns);
    // addBorder + IC_Angle on the un-blurred level (:717, Fast_gpu.cu:462-471)
    const int size = (int)(kPatchSize * e->sf[l]);
    all[l].resize(ns);
    for (int i = 0; i < ns; i++) {
      orc_keypoint& k = all[l][i];
      k.x = (float)(e->selected[l][i].x + kFastBand);
      k.y = (float)(e->selected[l][i].y + kFastBand);
      k.size = (float)size;
      k.response = (float)e->selected[l][i].score;
      k.octave = l;
      k.class_id = -1;
      k.angle = orc_ic_angle(plain[l].data(), lw, (int)k.x, (int)k.y);
    }
    total += ns;
    // Gaussian blur of the ROI after orientation (:719,:742)
    e->blurred[l].resize((size_t)lw * lh);
    orc_gauss7_u8(plain[l].data(), lw, lh, lw, e->blurred[l].data(), lw);
  }
  int off = 0;
  for (int l = 0; l < L; l++) {
    const int lw = e->lw[l];
    for (size_t i = 0; i < all[l].size(); i++) {
      if (off >= cap) return off;
      orc_keypoint k = all[l][i];
      orc_rbrief(e->blurred[l].data(), lw, (int)k.x, (int)k.y, k.angle, desc + (size_t)off * 32);
      if (l != 0) {  // :808-814
        k.x *= e->sf[l];
        k.y *= e->sf[l];
      }
      kps[off++] = k;
    }
  }
  (void)total;
  return off;
}

int orc_extractor_level(orc_extractor* e, int level, int which, const uint8_t** data, int* w, int* h, int* stride) {
  if (!e || level < 0 || level >= e->nlevels || e->lw.empty()) return -1;
  const int lw = e->lw[level], lh = e->lh[level];
  if (which == 0) {
    *data = e->bordered[level].data(); *w = lw + 2 * kEdge; *h = lh + 2 * kEdge; *stride = *w;
  } else if (which == 1) {
    *data = e->blurred[level].data(); *w = lw; *h = lh; *stride = lw;
  } else if (which == 2) {
    *data = e->score[level].data(); *w = lw; *h = lh; *stride = lw;
  } else {
    return -1;
  }
  return 0;
}

int orc_extractor_level_fast(orc_extractor* e, int level, const orc_fastpt** pts) {
  if (!e || level < 0 || level >= (int)e->fast.size()) return -1;
  *pts = e->fast[level].data();
  return (int)e->fast[level].size();
}

int orc_extractor_level_selected(orc_extractor* e, int level, const orc_fastpt** pts) {
  if (!e || level < 0 || level >= (int)e->selected.size()) return -1;
  *pts = e->selected[level].data();
  return (int)e->selected[level].size();
}

double orc_time_extract(orc_extractor* e, const uint8_t* imgs, int n_imgs, int w, int h, int iters) {
  std::vector<orc_keypoint> kps(4 * (size_t)e->nfeatures + 64);
  std::vector<uint8_t> desc(kps.size() * 32);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++)
    orc_extract(e, imgs + (size_t)(i % n_imgs) * w * h, w, h, w, kps.data(), desc.data(), (int)kps.size());
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
