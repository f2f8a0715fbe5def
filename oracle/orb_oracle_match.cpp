// orb_oracle_match.cpp -- CPU ORACLE for the ORBmatcher path on flat POD inputs.
// TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Loop order, greedy state and the threshold /
// ratio asymmetries of the reference are kept literally; every function cites the reference
// file:line it restates (paths relative to /root/reference/code/).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>
#include <map>
#include <chrono>

#include "orb_oracle.h"

namespace {

constexpr int TH_HIGH = 100;     // ORBmatcher.cc:37
constexpr int TH_LOW = 50;       // ORBmatcher.cc:38
constexpr int HISTO_LENGTH = 30; // ORBmatcher.cc:39
constexpr int GRID_COLS = 64;    // Frame.h:38
constexpr int GRID_ROWS = 48;    // Frame.h:37

inline int dist256(const uint8_t* a, const uint8_t* b) { return orc_hamming256(a, b); }

// ORBmatcher.cc:1475-1506 (histogram passed by value; strict > so the first bin wins ties)
void three_maxima(const std::vector<std::vector<int>>& histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) {
      max3 = max2; max2 = max1; max1 = s;
      ind3 = ind2; ind2 = ind1; ind1 = i;
    } else if (s > max2) {
      max3 = max2; max2 = s;
      ind3 = ind2; ind2 = i;
    } else if (s > max3) {
      max3 = s;
      ind3 = i;
    }
  }
  if (max2 < 0.1f * (float)max1) {
    ind2 = -1;
    ind3 = -1;
  } else if (max3 < 0.1f * (float)max1) {
    ind3 = -1;
  }
}

// The rotation-histogram bin, inlined six times in the reference (e.g. ORBmatcher.cc:218-227):
// factor = 1.0f/HISTO_LENGTH, so only bins 0..12 are ever hit (upstream quirk, kept).
inline int rot_bin(float a1, float a2) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

}  // namespace

struct orc_grid {
  std::vector<int32_t> cell[GRID_COLS][GRID_ROWS];
  float inv_w, inv_h;
};

extern "C" {

// Frame::AssignFeaturesToGrid (Frame.cc:277-292) + PosInGrid (:432-442, uses round)
orc_grid* orc_grid_build(const orc_frame* f) {
  orc_grid* g = new orc_grid;
  g->inv_w = (float)GRID_COLS / (float)(f->max_x - f->min_x);  // Frame.cc:259-260
  g->inv_h = (float)GRID_ROWS / (float)(f->max_y - f->min_y);
  for (int i = 0; i < f->n; i++) {
    const int px = (int)std::round((f->x[i] - f->min_x) * g->inv_w);
    const int py = (int)std::round((f->y[i] - f->min_y) * g->inv_h);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
    g->cell[px][py].push_back(i);
  }
  return g;
}

void orc_grid_destroy(orc_grid* g) { delete g; }

void orc_grid_csr(const orc_grid* g, int32_t* starts, int32_t* items) {
  int k = 0;
  for (int ix = 0; ix < GRID_COLS; ix++)
    for (int iy = 0; iy < GRID_ROWS; iy++) {
      starts[ix * GRID_ROWS + iy] = k;
      for (int32_t v : g->cell[ix][iy]) items[k++] = v;
    }
  starts[GRID_COLS * GRID_ROWS] = k;
}

// Frame::GetFeaturesInArea (Frame.cc:377-430); result order = for ix, for iy, for j in cell.
static void features_in_area(const orc_grid* g, const orc_frame* f, float x, float y, float r, int min_level,
                             int max_level, std::vector<int32_t>& out) {
  out.clear();
  const int min_cx = std::max(0, (int)std::floor((x - f->min_x - r) * g->inv_w));
  if (min_cx >= GRID_COLS) return;
  const int max_cx = std::min(GRID_COLS - 1, (int)std::ceil((x - f->min_x + r) * g->inv_w));
  if (max_cx < 0) return;
  const int min_cy = std::max(0, (int)std::floor((y - f->min_y - r) * g->inv_h));
  if (min_cy >= GRID_ROWS) return;
  const int max_cy = std::min(GRID_ROWS - 1, (int)std::ceil((y - f->min_y + r) * g->inv_h));
  if (max_cy < 0) return;
  const bool check_levels = (min_level > 0) || (max_level >= 0);
  for (int ix = min_cx; ix <= max_cx; ix++)
    for (int iy = min_cy; iy <= max_cy; iy++)
      for (int32_t j : g->cell[ix][iy]) {
        if (check_levels) {
          if (f->octave[j] < min_level) continue;
          if (max_level >= 0 && f->octave[j] > max_level) continue;
        }
        const float dx = f->x[j] - x, dy = f->y[j] - y;
        if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(j);
      }
}

int orc_grid_query(const orc_grid* g, const orc_frame* f, float x, float y, float r, int min_level, int max_level,
                   int32_t* out, int cap) {
  std::vector<int32_t> v;
  features_in_area(g, f, x, y, r, min_level, max_level, v);
  const int n = std::min((int)v.size(), cap);
  std::memcpy(out, v.data(), (size_t)n * sizeof(int32_t));
  return (int)v.size();
}

// ORBmatcher::SearchForInitialization, ORBmatcher.cc:375-479
int orc_search_for_initialization(const orc_frame* f1, const orc_frame* f2, float* prev_xy, int32_t* matches12,
                                  int window, float nnratio, int check_ori) {
  int nmatches = 0;
  for (int i = 0; i < f1->n; i++) matches12[i] = -1;
  std::vector<std::vector<int>> rot_hist(HISTO_LENGTH);
  std::vector<int> matched_dist(f2->n, INT_MAX), matches21(f2->n, -1);
  orc_grid* g2 = orc_grid_build(f2);
  std::vector<int32_t> cand;
  for (int i1 = 0; i1 < f1->n; i1++) {
    const int level1 = f1->octave[i1];
    if (level1 > 0) continue;
    features_in_area(g2, f2, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window, level1, level1, cand);
    if (cand.empty()) continue;
    const uint8_t* d1 = f1->desc + (size_t)i1 * 32;
    int best = INT_MAX, best2 = INT_MAX, best_idx2 = -1;
    for (int32_t i2 : cand) {
      const int dist = dist256(d1, f2->desc + (size_t)i2 * 32);
      if (matched_dist[i2] <= dist) continue;
      if (dist < best) {
        best2 = best;
        best = dist;
        best_idx2 = i2;
      } else if (dist < best2) {
        best2 = dist;
      }
    }
    if (best <= TH_LOW) {
      if (best < (float)best2 * nnratio) {
        if (matches21[best_idx2] >= 0) {
          matches12[matches21[best_idx2]] = -1;
          nmatches--;
        }
        matches12[i1] = best_idx2;
        matches21[best_idx2] = i1;
        matched_dist[best_idx2] = best;
        nmatches++;
        if (check_ori) rot_hist[rot_bin(f1->angle[i1], f2->angle[best_idx2])].push_back(i1);
      }
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rot_hist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rot_hist[i])
        if (matches12[idx1] >= 0) {
          matches12[idx1] = -1;
          nmatches--;
        }
    }
  }
  for (int i1 = 0; i1 < f1->n; i1++)
    if (matches12[i1] >= 0) {
      prev_xy[2 * i1] = f2->x[matches12[i1]];
      prev_xy[2 * i1 + 1] = f2->y[matches12[i1]];
    }
  orc_grid_destroy(g2);
  return nmatches;
}

// Generic windowed projection matcher.  One loop body covers
//   SearchByProjection(Frame&, vector<MapPoint*>&, th)            ORBmatcher.cc:44-121   (ratio_mode 1, TH_HIGH)
//   SearchByProjection(KeyFrame*, Scw, points, matched, th)       ORBmatcher.cc:264-373  (ratio_mode 0, TH_LOW)
//   SearchByProjection(Frame& cur, const Frame& last, th, mono)   ORBmatcher.cc:1223-1354 (ratio_mode 0, TH_HIGH, ori)
//   SearchByProjection(Frame&, KeyFrame*, found, th, ORBdist)     ORBmatcher.cc:1356-1473 (ratio_mode 0, ORBdist, ori)
// after the caller has projected every source point (see include/swm_orb.h swm_match_window).
int orc_match_window(const orc_frame* tgt, const orc_window_query* q, const uint8_t* tgt_blocked_init, int th_dist,
                     int ratio_mode, float nnratio, int check_ori, int32_t* assignment) {
  int nmatches = 0;
  std::vector<uint8_t> blocked(tgt->n, 0);
  if (tgt_blocked_init) std::memcpy(blocked.data(), tgt_blocked_init, tgt->n);
  std::vector<std::vector<int>> rot_hist(HISTO_LENGTH);
  orc_grid* g = orc_grid_build(tgt);
  std::vector<int32_t> cand;
  for (int s = 0; s < q->m; s++) {
    if (!q->valid[s]) continue;
    features_in_area(g, tgt, q->u[s], q->v[s], q->radius[s], q->min_level[s], q->max_level[s], cand);
    if (cand.empty()) continue;
    const uint8_t* d = q->desc + (size_t)s * 32;
    int best = 256, best2 = 256, best_level = -1, best_level2 = -1, best_idx = -1;
    for (int32_t i2 : cand) {
      if (blocked[i2]) continue;  // :67-69 / :306 / :1291-1293 / :1413
      const int dist = dist256(d, tgt->desc + (size_t)i2 * 32);
      if (dist < best) {
        best2 = best;
        best = dist;
        best_level2 = best_level;
        best_level = tgt->octave[i2];
        best_idx = i2;
      } else if (dist < best2) {
        best_level2 = tgt->octave[i2];
        best2 = dist;
      }
    }
    if (best <= th_dist) {
      if (ratio_mode == 1 && best_level == best_level2 && best > nnratio * best2) continue;  // :110-113
      assignment[best_idx] = s;
      if (q->blocks[s]) blocked[best_idx] = 1;
      nmatches++;
      if (check_ori) rot_hist[rot_bin(q->angle[s], tgt->angle[best_idx])].push_back(best_idx);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rot_hist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i != ind1 && i != ind2 && i != ind3) {
        for (int idx : rot_hist[i]) {
          assignment[idx] = -1;
          nmatches--;
        }
      }
    }
  }
  orc_grid_destroy(g);
  return nmatches;
}

// ORBmatcher::SearchByBoW: mode 0 KeyFrame->Frame (:150-262), mode 1 KeyFrame<->KeyFrame (:481-597)
int orc_search_by_bow(const orc_frame* f1, const orc_featvec* fv1, const uint8_t* valid1, const orc_frame* f2,
                      const orc_featvec* fv2, const uint8_t* valid2, int mode, float nnratio, int check_ori,
                      int32_t* matches) {
  int nmatches = 0;
  const int n_out = mode == 0 ? f2->n : f1->n;
  for (int i = 0; i < n_out; i++) matches[i] = -1;
  std::vector<uint8_t> taken2(f2->n, 0);  // mode 0: vpMapPointMatches[idxF] != NULL ; mode 1: vbMatched2
  std::vector<std::vector<int>> rot_hist(HISTO_LENGTH);
  int a = 0, b = 0;
  while (a < fv1->n_nodes && b < fv2->n_nodes) {
    if (fv1->node_ids[a] == fv2->node_ids[b]) {
      for (int i1 = fv1->offsets[a]; i1 < fv1->offsets[a + 1]; i1++) {
        const uint32_t idx1 = fv1->feats[i1];
        if (!valid1[idx1]) continue;  // !pMP || pMP->isBad()
        const uint8_t* d1 = f1->desc + (size_t)idx1 * 32;
        int best1 = 256, best_idx2 = -1, best2 = 256;
        for (int i2 = fv2->offsets[b]; i2 < fv2->offsets[b + 1]; i2++) {
          const uint32_t idx2 = fv2->feats[i2];
          if (taken2[idx2]) continue;
          if (mode == 1 && !valid2[idx2]) continue;
          const int dist = dist256(d1, f2->desc + (size_t)idx2 * 32);
          if (dist < best1) {
            best2 = best1;
            best1 = dist;
            best_idx2 = (int)idx2;
          } else if (dist < best2) {
            best2 = dist;
          }
        }
        const bool pass = mode == 0 ? (best1 <= TH_LOW) : (best1 < TH_LOW);  // :212 vs :550
        if (pass && (float)best1 < nnratio * (float)best2) {
          taken2[best_idx2] = 1;
          if (mode == 0) {
            matches[best_idx2] = (int32_t)idx1;
            if (check_ori) rot_hist[rot_bin(f1->angle[idx1], f2->angle[best_idx2])].push_back(best_idx2);
          } else {
            matches[idx1] = best_idx2;
            if (check_ori) rot_hist[rot_bin(f1->angle[idx1], f2->angle[best_idx2])].push_back((int)idx1);
          }
          nmatches++;
        }
      }
      a++;
      b++;
    } else if (fv1->node_ids[a] < fv2->node_ids[b]) {
      a = (int)(std::lower_bound(fv1->node_ids, fv1->node_ids + fv1->n_nodes, fv2->node_ids[b]) - fv1->node_ids);
    } else {
      b = (int)(std::lower_bound(fv2->node_ids, fv2->node_ids + fv2->n_nodes, fv1->node_ids[a]) - fv2->node_ids);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rot_hist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rot_hist[i]) {
        matches[idx] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// CheckDistEpipolarLine, :131-148
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, int octave2, const float* F12,
                                     const float* level_sigma2) {
  const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
  const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
  const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
  const float num = a * x2 + b * y2 + c;
  const float den = a * a + b * b;
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * level_sigma2[octave2];
}

// ORBmatcher::SearchForTriangulation, :599-749 (monocular)
int orc_search_for_triangulation(const orc_frame* f1, const orc_featvec* fv1, const uint8_t* valid1, const orc_frame* f2,
                                 const orc_featvec* fv2, const uint8_t* valid2, const float* F12, float ex, float ey,
                                 const float* scale_factors2, const float* level_sigma2, int check_ori,
                                 int32_t* matches12) {
  int nmatches = 0;
  for (int i = 0; i < f1->n; i++) matches12[i] = -1;
  std::vector<std::vector<int>> rot_hist(HISTO_LENGTH);
  int a = 0, b = 0;
  while (a < fv1->n_nodes && b < fv2->n_nodes) {
    if (fv1->node_ids[a] == fv2->node_ids[b]) {
      for (int i1 = fv1->offsets[a]; i1 < fv1->offsets[a + 1]; i1++) {
        const uint32_t idx1 = fv1->feats[i1];
        if (!valid1[idx1]) continue;  // already a MapPoint
        const uint8_t* d1 = f1->desc + (size_t)idx1 * 32;
        int best_dist = TH_LOW, best_idx2 = -1;
        for (int i2 = fv2->offsets[b]; i2 < fv2->offsets[b + 1]; i2++) {
          const uint32_t idx2 = fv2->feats[i2];
          if (!valid2[idx2]) continue;  // vbMatched2 is never set in the reference; pMP2 != NULL
          const int dist = dist256(d1, f2->desc + (size_t)idx2 * 32);
          if (dist > TH_LOW || dist > best_dist) continue;
          const float distex = ex - f2->x[idx2];
          const float distey = ey - f2->y[idx2];
          if (distex * distex + distey * distey < 100 * scale_factors2[f2->octave[idx2]]) continue;
          if (check_dist_epipolar_line(f1->x[idx1], f1->y[idx1], f2->x[idx2], f2->y[idx2], f2->octave[idx2], F12,
                                       level_sigma2)) {
            best_idx2 = (int)idx2;
            best_dist = dist;
          }
        }
        if (best_idx2 >= 0) {
          matches12[idx1] = best_idx2;
          nmatches++;
          if (check_ori) rot_hist[rot_bin(f1->angle[idx1], f2->angle[best_idx2])].push_back((int)idx1);
        }
      }
      a++;
      b++;
    } else if (fv1->node_ids[a] < fv2->node_ids[b]) {
      a = (int)(std::lower_bound(fv1->node_ids, fv1->node_ids + fv1->n_nodes, fv2->node_ids[b]) - fv1->node_ids);
    } else {
      b = (int)(std::lower_bound(fv2->node_ids, fv2->node_ids + fv2->n_nodes, fv1->node_ids[a]) - fv2->node_ids);
    }
  }
  if (check_ori) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rot_hist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rot_hist[i]) {
        matches12[idx] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}

// Fuse / SearchBySim3 inner loop (see header)
void orc_window_best(const orc_frame* tgt, int m, const uint8_t* desc, const float* u, const float* v, const float* radius,
                     const int32_t* pred_level, const uint8_t* valid, const float* inv_level_sigma2, float chi2,
                     int32_t* best_idx, int32_t* best_dist) {
  orc_grid* g = orc_grid_build(tgt);
  std::vector<int32_t> idx((size_t)std::max(tgt->n, 1));
  for (int i = 0; i < m; i++) {
    best_idx[i] = -1;
    best_dist[i] = 256;
    if (!valid[i]) continue;
    const int cnt = orc_grid_query(g, tgt, u[i], v[i], radius[i], -1, -1, idx.data(), (int)idx.size());  // no level args
    int bestDist = 256, bestIdx = -1;
    for (int c = 0; c < cnt; c++) {
      const int j = idx[c];
      const int kpLevel = tgt->octave[j];
      if (kpLevel < pred_level[i] - 1 || kpLevel > pred_level[i]) continue;
      if (chi2 > 0.0f) {
        const float ex = u[i] - tgt->x[j];
        const float ey = v[i] - tgt->y[j];
        const float e2 = ex * ex + ey * ey;
        if (e2 * inv_level_sigma2[kpLevel] > (double)chi2) continue;
      }
      const int dist = dist256(desc + (size_t)i * 32, tgt->desc + (size_t)j * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx = j;
      }
    }
    best_idx[i] = bestIdx;
    best_dist[i] = bestDist;
  }
  orc_grid_destroy(g);
}

// MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:361-391
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best_idx,
                                 int32_t* best_median) {
  for (int p = 0; p < npoints; p++) {
    const uint8_t* d = desc + (size_t)offsets[p] * 32;
    const size_t N = (size_t)(offsets[p + 1] - offsets[p]);
    best_idx[p] = -1;
    best_median[p] = -1;
    if (N == 0) continue;
    std::vector<float> distances(N * N);
    for (size_t i = 0; i < N; i++) {
      distances[i * N + i] = 0;
      for (size_t j = i + 1; j < N; j++) {
        const int distij = dist256(d + i * 32, d + j * 32);
        distances[i * N + j] = (float)distij;
        distances[j * N + i] = (float)distij;
      }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (size_t i = 0; i < N; i++) {
      std::vector<int> vDists(distances.begin() + i * N, distances.begin() + (i + 1) * N);
      std::sort(vDists.begin(), vDists.end());
      const int median = vDists[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) {
        BestMedian = median;
        BestIdx = (int)i;
      }
    }
    best_idx[p] = BestIdx;
    best_median[p] = BestMedian;
  }
}

// Frame::ComputeStereoMatches, Frame.cc:516-690.  Float expressions are evaluated left to right in float exactly as
// written there (the oracle is compiled with -ffp-contract=off); `round` is C's (half away from zero); the SAD values
// are integers, so cv::norm's double accumulation and the float it is stored in are exact.
int orc_stereo_matches(const orc_keypoint* kl, const uint8_t* dl, int nl, const orc_keypoint* kr, const uint8_t* dr, int nr,
                       const orc_pyramid* pl, const orc_pyramid* pr, const float* sf, const float* inv_sf, float mbf, float mb,
                       float* u_right, float* depth) {
  const int TH_HIGH = 100;  // ORBmatcher.cc:38
  for (int i = 0; i < nl; i++) u_right[i] = depth[i] = -1.0f;
  const int nRows = pl->h[0];
  std::vector<std::vector<int>> vRowIndices(nRows);  // :524-541
  for (int iR = 0; iR < nr; iR++) {
    const float kpY = kr[iR].y;
    const float r = 2.0f * sf[kr[iR].octave];
    const int maxr = (int)std::ceil(kpY + r);
    const int minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);  // (the reference indexes unchecked; keypoints keep >= 14 px from the edge)
  }
  const float minZ = mb, minD = -3, maxD = mbf / minZ;  // :544-546
  std::vector<std::pair<int, int>> vDistIdx;
  for (int iL = 0; iL < nl; iL++) {
    const int levelL = kl[iL].octave;
    const float vL = kl[iL].y, uL = kl[iL].x;
    const int row = (int)vL;  // vRowIndices[vL]: float -> size_t truncation
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cand = vRowIndices[row];
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH, bestIdxR = 0;
    for (int iR : cand) {  // :572-594
      if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
      const float uR = kr[iR].x;
      if (uR >= minU && uR <= maxU) {
        const int dist = dist256(dl + (size_t)iL * 32, dr + (size_t)iR * 32);
        if (dist < bestDist) {
          bestDist = dist;
          bestIdxR = iR;
        }
      }
    }
    if (!(bestDist < TH_HIGH)) continue;
    // sub-pixel match by correlation, :597-672
    const float uR0 = kr[bestIdxR].x;
    const float scaleFactor = inv_sf[levelL];
    const float scaleduL = std::round(uL * scaleFactor), scaledvL = std::round(vL * scaleFactor);
    const float scaleduR0 = std::round(uR0 * scaleFactor);
    const int w = 5, L = 5;
    const uint8_t* imL = pl->plane[levelL];
    const uint8_t* imR = pr->plane[levelL];
    const int sL = pl->stride[levelL], sR = pr->stride[levelL];
    const int y0 = (int)(scaledvL - w), xL0 = (int)(scaleduL - w);
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= pr->w[levelL]) continue;
    const int cL = imL[(y0 + w) * sL + xL0 + w];
    int bestSad = INT_MAX, bestincR = 0;
    float vDists[2 * 5 + 1];
    for (int incR = -L; incR <= +L; incR++) {
      const int xR0 = (int)(scaleduR0 + incR - w);
      const int cR = imR[(y0 + w) * sR + xR0 + w];
      int sad = 0;
      for (int yy = 0; yy < 2 * w + 1; yy++)
        for (int xx = 0; xx < 2 * w + 1; xx++)
          sad += std::abs((imL[(y0 + yy) * sL + xL0 + xx] - cL) - (imR[(y0 + yy) * sR + xR0 + xx] - cR));
      const float dist = (float)sad;
      if (dist < bestSad) {
        bestSad = (int)dist;
        bestincR = incR;
      }
      vDists[L + incR] = dist;
    }
    if (bestincR == -L || bestincR == L) continue;
    const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = sf[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
    float disparity = (uL - bestuR);
    if (disparity >= 0 && disparity < maxD) {
      if (disparity <= 0) {
        disparity = 0.01;
        bestuR = uL - 0.01;
      }
      depth[iL] = mbf / disparity;
      u_right[iL] = bestuR;
      vDistIdx.push_back(std::pair<int, int>(bestSad, iL));
    }
  }
  if (vDistIdx.empty()) return 0;  // (the reference would index an empty vector here)
  std::sort(vDistIdx.begin(), vDistIdx.end());  // :675-689
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)vDistIdx.size();
  for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
    if (vDistIdx[i].first < thDist) break;
    u_right[vDistIdx[i].second] = -1;
    depth[vDistIdx[i].second] = -1;
    kept--;
  }
  return kept;
}

// Brute-force top-2 (config 5): ties broken by the lower database index.
void orc_bruteforce_top2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int32_t* out4) {
  for (int i = 0; i < nq; i++) {
    int d0 = 257, d1 = 257;
    int64_t i0 = -1, i1 = -1;
    for (int64_t j = 0; j < ndb; j++) {
      const int d = dist256(q + (size_t)i * 32, db + (size_t)j * 32);
      if (d < d0) {
        d1 = d0; i1 = i0;
        d0 = d; i0 = j;
      } else if (d < d1) {
        d1 = d; i1 = j;
      }
    }
    out4[4 * i] = d0;
    out4[4 * i + 1] = (int32_t)i0;
    out4[4 * i + 2] = d1;
    out4[4 * i + 3] = (int32_t)i1;
  }
}

// cv::undistortPoints with P = K, R = I (see header).  One statement per OpenCV expression, evaluated left to right in
// double; compiled with -ffp-contract=off so that no multiply-add is fused.
static void undistort_one(float px, float py, const float* cam9, float* ox, float* oy) {
  const double fx = cam9[0], fy = cam9[1], cx = cam9[2], cy = cam9[3];
  const double k1 = cam9[4], k2 = cam9[5], p1 = cam9[6], p2 = cam9[7], k3 = cam9[8];
  const double ifx = 1. / fx, ify = 1. / fy;
  const double u = px, v = py;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);  // numerator 1 + ((k6 r2 + k5) r2 + k4) r2 = 1
    if (icdist < 0) {
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  // new camera matrix = K: xx = fx x + 0 y + cx, ww = 1 / (0 x + 0 y + 1)
  *ox = (float)(fx * x + cx);
  *oy = (float)(fy * y + cy);
}

void orc_undistort_points(const float* xy, int n, const float* cam9, float* out_xy) {
  if (cam9[4] == 0.0f) {
    for (int i = 0; i < 2 * n; i++) out_xy[i] = xy[i];
    return;
  }
  for (int i = 0; i < n; i++) undistort_one(xy[2 * i], xy[2 * i + 1], cam9, &out_xy[2 * i], &out_xy[2 * i + 1]);
}

void orc_image_bounds(int cols, int rows, const float* cam9, float* out4) {
  if (cam9[4] == 0.0f) {
    out4[0] = 0.0f; out4[1] = (float)cols; out4[2] = 0.0f; out4[3] = (float)rows;
    return;
  }
  const float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
  float m[8];
  orc_undistort_points(c, 4, cam9, m);
  out4[0] = std::min(m[0], m[4]);  // mnMinX = min(mat(0,0), mat(2,0))
  out4[1] = std::max(m[2], m[6]);  // mnMaxX = max(mat(1,0), mat(3,0))
  out4[2] = std::min(m[1], m[3]);  // mnMinY = min(mat(0,1), mat(1,1))
  out4[3] = std::max(m[5], m[7]);  // mnMaxY = max(mat(2,1), mat(3,1))
}

// ---------------------------------------------------------------------------------------------
// DBoW2 vocabulary + transform (see header for the reference lines)
// ---------------------------------------------------------------------------------------------
struct orc_vocab {
  struct Node {
    uint32_t parent = 0;
    std::vector<uint32_t> children;
    uint8_t desc[32] = {0};
    double weight = 0;  // WordValue is double; the file stores float (:1507)
    uint32_t word_id = 0;
  };
  int k = 0, L = 0, scoring = 0, weighting = 0;
  std::vector<Node> nodes;
  int n_words = 0;
};

orc_vocab* orc_vocab_load(const uint8_t* blob, size_t bytes) {
  if (bytes < 24) return nullptr;
  uint32_t nb_nodes, size_node;
  int32_t hdr[4];
  memcpy(&nb_nodes, blob, 4);
  memcpy(&size_node, blob + 4, 4);
  memcpy(hdr, blob + 8, 16);
  if (size_node != 41) return nullptr;
  orc_vocab* v = new orc_vocab();
  v->k = hdr[0]; v->L = hdr[1]; v->scoring = hdr[2]; v->weighting = hdr[3];
  const size_t n_rec = (bytes - 24) / size_node;
  v->nodes.resize(n_rec + 1);
  for (size_t r = 0; r < n_rec; r++) {
    const uint8_t* buf = blob + 24 + r * size_node;
    const uint32_t nid = (uint32_t)r + 1;
    int32_t parent;
    float w;
    memcpy(&parent, buf, 4);
    memcpy(&w, buf + 36, 4);
    if (parent < 0 || (size_t)parent >= v->nodes.size()) { delete v; return nullptr; }
    orc_vocab::Node& n = v->nodes[nid];
    n.parent = (uint32_t)parent;
    v->nodes[parent].children.push_back(nid);  // :1504
    memcpy(n.desc, buf + 4, 32);
    n.weight = w;
    if (buf[40]) n.word_id = (uint32_t)v->n_words++;  // :1509-1512
  }
  (void)nb_nodes;
  return v;
}

void orc_vocab_destroy(orc_vocab* v) { delete v; }

void orc_vocab_info(const orc_vocab* v, int32_t* k, int32_t* L, int32_t* n_nodes, int32_t* n_words) {
  *k = v->k; *L = v->L; *n_nodes = (int32_t)v->nodes.size(); *n_words = v->n_words;
}

static int forb_distance(const uint8_t* a, const uint8_t* b) {  // FORB.cpp:82-102
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    memcpy(&x, a + 4 * i, 4);
    memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// transform(feature, word_id, weight, nid, levelsup), :1242-1283.  When the leaf is shallower than nid_level the
// reference leaves *nid unset (an uninitialised read in the caller); defined here as the leaf itself.
static void transform_one(const orc_vocab* v, const uint8_t* f, int levelsup, uint32_t* word, double* weight, uint32_t* nid) {
  const int nid_level = v->L - levelsup;
  bool nid_set = false;
  if (nid_level <= 0) { *nid = 0; nid_set = true; }
  uint32_t final_id = 0;
  int current_level = 0;
  do {
    ++current_level;
    const std::vector<uint32_t>& nodes = v->nodes[final_id].children;
    final_id = nodes[0];
    double best_d = forb_distance(f, v->nodes[final_id].desc);
    for (size_t c = 1; c < nodes.size(); c++) {
      const double d = forb_distance(f, v->nodes[nodes[c]].desc);
      if (d < best_d) { best_d = d; final_id = nodes[c]; }
    }
    if (current_level == nid_level) { *nid = final_id; nid_set = true; }
  } while (!v->nodes[final_id].children.empty());
  if (!nid_set) *nid = final_id;
  *word = v->nodes[final_id].word_id;
  *weight = v->nodes[final_id].weight;
}

int orc_bow_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_ids,
                      double* word_values, uint32_t* node_ids, int32_t* offsets, uint32_t* feats, int32_t* n_nodes_out) {
  std::map<uint32_t, double> bow;                    // BowVector
  std::map<uint32_t, std::vector<uint32_t>> fv;      // FeatureVector
  *n_nodes_out = 0;
  offsets[0] = 0;
  if (v->nodes.size() <= 1) return 0;                // empty()
  const bool tf = v->weighting == 0 || v->weighting == 1;  // TF_IDF, TF -> addWeight; IDF, BINARY -> addIfNotExist
  const bool must = v->scoring != 5;                 // DotProductScoring does not normalise
  const bool l2 = v->scoring == 1;
  for (int i = 0; i < n; i++) {
    uint32_t id, nid;
    double w;
    transform_one(v, desc + (size_t)i * 32, levelsup, &id, &w, &nid);
    if (w > 0) {
      auto it = bow.lower_bound(id);
      if (it != bow.end() && it->first == id) {
        if (tf) it->second += w;
      } else {
        bow.insert(it, std::make_pair(id, w));
      }
      fv[nid].push_back((uint32_t)i);
    }
  }
  if (tf && !bow.empty() && !must) {
    const double nd = (double)bow.size();
    for (auto& e : bow) e.second /= nd;
  }
  if (must) {
    double norm = 0.0;
    if (!l2) {
      for (auto& e : bow) norm += fabs(e.second);
    } else {
      for (auto& e : bow) norm += e.second * e.second;
      norm = sqrt(norm);
    }
    if (norm > 0.0)
      for (auto& e : bow) e.second /= norm;
  }
  int nw = 0;
  for (auto& e : bow) { word_ids[nw] = e.first; word_values[nw] = e.second; nw++; }
  int nn = 0, nf = 0;
  for (auto& e : fv) {
    node_ids[nn] = e.first;
    for (uint32_t idx : e.second) feats[nf++] = idx;
    offsets[++nn] = nf;
  }
  *n_nodes_out = nn;
  return nw;
}

double orc_time_bow_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, int iters) {
  std::vector<uint32_t> a(n + 1), c(n + 1), e(n + 1);
  std::vector<double> b(n + 1);
  std::vector<int32_t> d(n + 2);
  int32_t nn;
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < iters; it++) orc_bow_transform(v, desc, n, levelsup, a.data(), b.data(), c.data(), d.data(), e.data(), &nn);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
