// ref_orbmatcher_wrap.cpp -- C entry points around the REFERENCE's own ORBmatcher.cc, compiled UNMODIFIED from
// /root/reference/code/src by `make -C oracle ref` (-> oracle/_ref/liborbmatcher_ref.so) against
//   * the reference's real code/include/ORBmatcher.h and the vendored DBoW2 FeatureVector,
//   * test doubles of Frame / KeyFrame / MapPoint (oracle/ref_shim_matcher/ref_doubles.h) whose grid and scale
//     member functions are the reference's own bodies, cut at build time from code/src/Frame.cc, KeyFrame.cc and
//     MapPoint.cc into oracle/_ref/gen/*.inc (never committed),
//   * a small-matrix stand-in for cv::Mat (oracle/ref_shim_matcher/opencv2/core/core.hpp, numerics pinned to cv2).
// The harness builds the doubles from flat arrays, calls the reference's member function, and flattens the result.
// Test infrastructure only (tests/test_ref_orbmatcher.py, tests/ref_vs_product_test.cpp); nothing in the product
// links or loads this.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <set>
#include <vector>

#include "ORBmatcher.h"  // the reference's own header (doubles are force-included by the Makefile)

using namespace std;

namespace ORB_SLAM2 {
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
// the reference's own function bodies (see the header comment)
#include "_ref/gen/frame_grid.inc"
#include "_ref/gen/frame_stereo.inc"
#include "_ref/gen/keyframe_grid.inc"
#include "_ref/gen/mappoint_scale.inc"

// KeyFrame's copy of AssignFeaturesToGrid / PosInGrid lives at KeyFrame.cc:1010-1040 and is identical to Frame's apart
// from the container; the harness fills KeyFrame::mGrid by running Frame's (reference) code on the same keypoints.
struct Exposed : public ORBmatcher {  // ComputeThreeMaxima is protected
  using ORBmatcher::ComputeThreeMaxima;
};
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

extern "C" {
#pragma GCC visibility push(default)

// Flat description of a Frame / KeyFrame.  All pointers may be NULL when the callee does not read the member.
struct refm_frame {
  int32_t n;
  const float* x;  // mvKeysUn[i].pt (mvKeys gets the same points: undistortion only rewrites pt, Frame.cc:479-482)
  const float* y;
  const int32_t* octave;
  const float* angle;
  const uint8_t* desc;  // n x 32
  float min_x, max_x, min_y, max_y;
  float fx, fy, cx, cy;
  int32_t nlevels;
  float log_scale_factor;
  const float* scale_factors;     // nlevels
  const float* level_sigma2;      // nlevels
  const float* inv_level_sigma2;  // nlevels
  const float* Tcw;               // 16, row-major, or NULL (identity)
  const int32_t* mp_index;        // n: index into the point table or -1 (mvpMapPoints)
  const uint8_t* outlier;         // n or NULL
  int32_t n_nodes;                // mFeatVec as CSR
  const uint32_t* node_ids;
  const int32_t* node_off;
  const uint32_t* node_feats;
};

struct refm_points {
  int32_t n;
  const float* pos;     // n x 3
  const float* normal;  // n x 3 or NULL
  const uint8_t* desc;  // n x 32
  const int32_t* nobs;  // Observations()
  const uint8_t* bad;   // isBad()
  const float* min_dist;  // mfMinDistance / mfMaxDistance (before the 0.8 / 1.2 factors)
  const float* max_dist;
  // fields Frame::isInFrustum writes (used by SearchByProjection(F, vpMapPoints, th)); NULL = not in view
  const uint8_t* track_in_view;
  const float* proj_x;
  const float* proj_y;
  const float* view_cos;
  const int32_t* track_level;
};

#pragma GCC visibility pop
}  // extern "C"

namespace {

cv::Mat vec3(const float* p) {
  cv::Mat m(3, 1, CV_32F);
  for (int i = 0; i < 3; i++) m.at<float>(i) = p[i];
  return m;
}

struct Points {
  vector<unique_ptr<MapPoint>> pts;
  explicit Points(const refm_points* d) {
    if (!d) return;
    for (int i = 0; i < d->n; i++) {
      unique_ptr<MapPoint> p(new MapPoint);
      p->index = i;
      p->mWorldPos = vec3(d->pos + 3 * i);
      const float up[3] = {0, 0, 1};
      p->mNormalVector = vec3(d->normal ? d->normal + 3 * i : up);
      p->mDescriptor = cv::Mat(1, 32, CV_8U);
      memcpy(p->mDescriptor.data, d->desc + 32 * (size_t)i, 32);
      p->nObs = d->nobs ? d->nobs[i] : 1;
      p->mbBad = d->bad ? d->bad[i] != 0 : false;
      p->mfMinDistance = d->min_dist ? d->min_dist[i] : 0.f;
      p->mfMaxDistance = d->max_dist ? d->max_dist[i] : 1e9f;
      if (d->track_in_view) {
        p->mbTrackInView = d->track_in_view[i] != 0;
        p->mTrackProjX = d->proj_x[i];
        p->mTrackProjY = d->proj_y[i];
        p->mTrackViewCos = d->view_cos[i];
        p->mnTrackScaleLevel = d->track_level[i];
      }
      pts.push_back(move(p));
    }
  }
  MapPoint* at(int i) { return i >= 0 && i < (int)pts.size() ? pts[i].get() : nullptr; }
};

void set_statics(const refm_frame* d) {
  Frame::fx = d->fx; Frame::fy = d->fy; Frame::cx = d->cx; Frame::cy = d->cy;
  Frame::mnMinX = d->min_x; Frame::mnMaxX = d->max_x; Frame::mnMinY = d->min_y; Frame::mnMaxY = d->max_y;
  // Frame.cc:99-100
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}

void fill_featvec(DBoW2::FeatureVector& fv, const refm_frame* d) {
  for (int a = 0; a < d->n_nodes; a++)
    for (int k = d->node_off[a]; k < d->node_off[a + 1]; k++) fv.addFeature(d->node_ids[a], d->node_feats[k]);
}

cv::Mat pose_of(const float* T) {
  cv::Mat m(4, 4, CV_32F);
  for (int i = 0; i < 16; i++) m.at<float>(i / 4, i % 4) = T ? T[i] : (i % 5 == 0 ? 1.f : 0.f);
  return m;
}

// Builds a Frame; sets the Frame statics from this description (callers build the "current" frame last when two
// frames with different cameras are involved -- the reference has one camera per process).
unique_ptr<Frame> make_frame(const refm_frame* d, Points& P) {
  set_statics(d);
  unique_ptr<Frame> f(new Frame);
  f->N = d->n;
  f->mvKeysUn.resize(d->n);
  for (int i = 0; i < d->n; i++) {
    cv::KeyPoint& k = f->mvKeysUn[i];
    k.pt.x = d->x[i]; k.pt.y = d->y[i];
    k.octave = d->octave ? d->octave[i] : 0;
    k.angle = d->angle ? d->angle[i] : 0.f;
  }
  f->mvKeys = f->mvKeysUn;
  f->mvuRight.assign(d->n, -1.f);
  f->mDescriptors = cv::Mat(d->n, 32, CV_8U);
  if (d->n) memcpy(f->mDescriptors.data, d->desc, (size_t)d->n * 32);
  f->mvpMapPoints.assign(d->n, nullptr);
  if (d->mp_index)
    for (int i = 0; i < d->n; i++) f->mvpMapPoints[i] = P.at(d->mp_index[i]);
  f->mvbOutlier.assign(d->n, false);
  if (d->outlier)
    for (int i = 0; i < d->n; i++) f->mvbOutlier[i] = d->outlier[i] != 0;
  f->mnScaleLevels = d->nlevels;
  f->mfLogScaleFactor = d->log_scale_factor;
  if (d->scale_factors) f->mvScaleFactors.assign(d->scale_factors, d->scale_factors + d->nlevels);
  f->mTcw = pose_of(d->Tcw);
  f->mRcw = f->mTcw.rowRange(0, 3).colRange(0, 3);  // Frame::UpdatePoseMatrices, Frame.cc:308-314
  f->mtcw = f->mTcw.rowRange(0, 3).col(3);
  f->mOw = -f->mRcw.t() * f->mtcw;
  if (d->node_ids) fill_featvec(f->mFeatVec, d);
  f->AssignFeaturesToGrid();
  return f;
}

unique_ptr<KeyFrame> make_keyframe(const refm_frame* d, Points& P) {
  unique_ptr<Frame> f = make_frame(d, P);  // runs the reference's AssignFeaturesToGrid
  unique_ptr<KeyFrame> k(new KeyFrame);
  k->N = f->N;
  k->mvKeysUn = f->mvKeysUn;
  k->mvuRight = f->mvuRight;
  k->mDescriptors = f->mDescriptors;
  k->mFeatVec = f->mFeatVec;
  k->fx = d->fx; k->fy = d->fy; k->cx = d->cx; k->cy = d->cy;
  k->mnScaleLevels = d->nlevels;
  k->mfLogScaleFactor = d->log_scale_factor;
  k->mvScaleFactors = f->mvScaleFactors;
  if (d->level_sigma2) k->mvLevelSigma2.assign(d->level_sigma2, d->level_sigma2 + d->nlevels);
  if (d->inv_level_sigma2) k->mvInvLevelSigma2.assign(d->inv_level_sigma2, d->inv_level_sigma2 + d->nlevels);
  // KeyFrame(Frame&) copies the bounds into ints (KeyFrame.cc:50-60) and the grid cell by cell (:67-72)
  k->mnMinX = Frame::mnMinX; k->mnMinY = Frame::mnMinY; k->mnMaxX = Frame::mnMaxX; k->mnMaxY = Frame::mnMaxY;
  k->mfGridElementWidthInv = Frame::mfGridElementWidthInv;
  k->mfGridElementHeightInv = Frame::mfGridElementHeightInv;
  k->mGrid.resize(k->mnGridCols);
  for (int i = 0; i < k->mnGridCols; i++) {
    k->mGrid[i].resize(k->mnGridRows);
    for (int j = 0; j < k->mnGridRows; j++) k->mGrid[i][j] = f->mGrid[i][j];
  }
  k->mvpMapPoints = f->mvpMapPoints;
  k->Rcw = f->mRcw.clone();
  k->tcw = f->mtcw.clone();
  k->Ow = f->mOw.clone();
  return k;
}

void flatten(const vector<MapPoint*>& v, int32_t* out) {
  for (size_t i = 0; i < v.size(); i++) out[i] = v[i] ? v[i]->index : -1;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int refm_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat A(1, 32, CV_8U), B(1, 32, CV_8U);
  memcpy(A.data, a, 32);
  memcpy(B.data, b, 32);
  return ORBmatcher::DescriptorDistance(A, B);
}

// Frame::GetFeaturesInArea (Frame.cc:377-430) on the grid built by Frame::AssignFeaturesToGrid (:277-292).
int refm_features_in_area(const refm_frame* fd, float x, float y, float r, int min_level, int max_level, int32_t* out,
                          int cap) {
  Points P(nullptr);
  unique_ptr<Frame> f = make_frame(fd, P);
  const vector<size_t> v = f->GetFeaturesInArea(x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// KeyFrame::GetFeaturesInArea (KeyFrame.cc:779-814).
int refm_kf_features_in_area(const refm_frame* fd, float x, float y, float r, int32_t* out, int cap) {
  Points P(nullptr);
  unique_ptr<KeyFrame> k = make_keyframe(fd, P);
  const vector<size_t> v = k->GetFeaturesInArea(x, y, r);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// cells as CSR (cell = ix * 48 + iy), items in the order AssignFeaturesToGrid appended them
void refm_grid_csr(const refm_frame* fd, int32_t* starts, int32_t* items) {
  Points P(nullptr);
  unique_ptr<Frame> f = make_frame(fd, P);
  int k = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
    for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
      starts[ix * FRAME_GRID_ROWS + iy] = k;
      for (size_t v : f->mGrid[ix][iy]) items[k++] = (int32_t)v;
    }
  starts[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
}

void refm_three_maxima(const int32_t* sizes, int L, int32_t* ind3) {
  vector<vector<int>> h(L);
  for (int i = 0; i < L; i++) h[i].assign(sizes[i], 0);
  int a = -1, b = -1, c = -1;
  Exposed::ComputeThreeMaxima(h, L, a, b, c);
  ind3[0] = a; ind3[1] = b; ind3[2] = c;
}

// MapPoint::PredictScale (MapPoint.cc:476-485)
int refm_predict_scale(float max_distance, float current_dist, float log_scale_factor, int nlevels) {
  MapPoint p;
  p.mfMaxDistance = max_distance;
  return p.PredictScale(current_dist, log_scale_factor, nlevels);
}

// Frame::isInFrustum (Frame.cc:316-375) for every point; writes the tracking fields back as flat arrays.
void refm_is_in_frustum(const refm_frame* fd, const refm_points* pd, float cos_limit, uint8_t* in_view, float* proj_x,
                        float* proj_y, float* view_cos, int32_t* level) {
  Points P(pd);
  unique_ptr<Frame> f = make_frame(fd, P);
  for (int i = 0; i < pd->n; i++) {
    MapPoint* p = P.at(i);
    in_view[i] = f->isInFrustum(p, cos_limit) ? 1 : 0;
    proj_x[i] = p->mTrackProjX; proj_y[i] = p->mTrackProjY;
    view_cos[i] = p->mTrackViewCos; level[i] = p->mnTrackScaleLevel;
  }
}

// SearchForInitialization (ORBmatcher.cc:375-479)
int refm_search_for_initialization(const refm_frame* f1d, const refm_frame* f2d, float* prev_xy, int32_t* matches12,
                                   int window, float nnratio, int check_ori) {
  Points P(nullptr);
  unique_ptr<Frame> F1 = make_frame(f1d, P), F2 = make_frame(f2d, P);
  vector<cv::Point2f> prev(f1d->n);
  for (int i = 0; i < f1d->n; i++) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
  vector<int> m12;
  ORBmatcher m(nnratio, check_ori != 0);
  const int n = m.SearchForInitialization(*F1, *F2, prev, m12, window);
  for (int i = 0; i < f1d->n; i++) { matches12[i] = m12[i]; prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
  return n;
}

// SearchByProjection(Frame& cur, const Frame& last, th, bMono) (:1223-1354).  assign[j] = point index held by
// cur.mvpMapPoints[j] after the call.
int refm_search_by_projection_last(const refm_frame* curd, const refm_frame* lastd, const refm_points* pd, float th,
                                   float nnratio, int check_ori, int32_t* assign) {
  Points P(pd);
  unique_ptr<Frame> last = make_frame(lastd, P), cur = make_frame(curd, P);
  ORBmatcher m(nnratio, check_ori != 0);
  const int n = m.SearchByProjection(*cur, *last, th, true);
  flatten(cur->mvpMapPoints, assign);
  return n;
}

// SearchByProjection(Frame& F, const vector<MapPoint*>&, th) (:44-121); order = indices into the point table.
int refm_search_by_projection_points(const refm_frame* fd, const refm_points* pd, const int32_t* order, int n_order,
                                     float th, float nnratio, int32_t* assign) {
  Points P(pd);
  unique_ptr<Frame> F = make_frame(fd, P);
  vector<MapPoint*> v;
  for (int i = 0; i < n_order; i++) v.push_back(P.at(order[i]));
  ORBmatcher m(nnratio, true);
  const int n = m.SearchByProjection(*F, v, th);
  flatten(F->mvpMapPoints, assign);
  return n;
}

// SearchByProjection(Frame& cur, KeyFrame*, sAlreadyFound, th, ORBdist, bGlobal) (:1356-1473)
int refm_search_by_projection_reloc(const refm_frame* curd, const refm_frame* kfd, const refm_points* pd,
                                    const int32_t* found, int n_found, float th, int orb_dist, float nnratio,
                                    int check_ori, int32_t* assign) {
  Points P(pd);
  unique_ptr<KeyFrame> kf = make_keyframe(kfd, P);
  unique_ptr<Frame> cur = make_frame(curd, P);
  set<MapPoint*> s;
  for (int i = 0; i < n_found; i++) s.insert(P.at(found[i]));
  ORBmatcher m(nnratio, check_ori != 0);
  const int n = m.SearchByProjection(*cur, kf.get(), s, th, orb_dist, false);
  flatten(cur->mvpMapPoints, assign);
  return n;
}

// SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (:264-373); matched: in/out point index per keypoint
int refm_search_by_projection_sim3(const refm_frame* kfd, const float* Scw, const refm_points* pd, const int32_t* order,
                                   int n_order, int32_t* matched, int th) {
  Points P(pd);
  unique_ptr<KeyFrame> kf = make_keyframe(kfd, P);
  vector<MapPoint*> v, vm(kfd->n);
  for (int i = 0; i < n_order; i++) v.push_back(P.at(order[i]));
  for (int i = 0; i < kfd->n; i++) vm[i] = P.at(matched[i]);
  ORBmatcher m(0.75f, true);
  const int n = m.SearchByProjection(kf.get(), pose_of(Scw), v, vm, th);
  flatten(vm, matched);
  return n;
}

// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (:150-262); out[j] = point index matched to F's keypoint j
int refm_search_by_bow_kf_f(const refm_frame* kfd, const refm_frame* fd, const refm_points* pd, float nnratio,
                            int check_ori, int32_t* out) {
  Points P(pd);
  unique_ptr<KeyFrame> kf = make_keyframe(kfd, P);
  unique_ptr<Frame> F = make_frame(fd, P);
  vector<MapPoint*> v;
  ORBmatcher m(nnratio, check_ori != 0);
  const int n = m.SearchByBoW(kf.get(), *F, v);
  flatten(v, out);
  return n;
}

// SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (:481-597); out[i1] = point index (of KF2's points) or -1
int refm_search_by_bow_kf_kf(const refm_frame* k1d, const refm_frame* k2d, const refm_points* pd, float nnratio,
                             int check_ori, int32_t* out) {
  Points P(pd);
  unique_ptr<KeyFrame> k1 = make_keyframe(k1d, P), k2 = make_keyframe(k2d, P);
  vector<MapPoint*> v;
  ORBmatcher m(nnratio, check_ori != 0);
  const int n = m.SearchByBoW(k1.get(), k2.get(), v);
  flatten(v, out);
  return n;
}

// SearchForTriangulation (:599-749), bOnlyStereo = false; pairs: (idx1, idx2) x n
int refm_search_for_triangulation(const refm_frame* k1d, const refm_frame* k2d, const refm_points* pd, const float* F12,
                                  int check_ori, int32_t* pairs, int cap) {
  Points P(pd);
  unique_ptr<KeyFrame> k1 = make_keyframe(k1d, P), k2 = make_keyframe(k2d, P);
  cv::Mat F(3, 3, CV_32F);
  for (int i = 0; i < 9; i++) F.at<float>(i / 3, i % 3) = F12[i];
  vector<pair<size_t, size_t>> vp;
  ORBmatcher m(0.6f, check_ori != 0);
  const int n = m.SearchForTriangulation(k1.get(), k2.get(), F, vp, false);
  for (size_t i = 0; i < vp.size() && (int)i < cap; i++) { pairs[2 * i] = (int32_t)vp[i].first; pairs[2 * i + 1] = (int32_t)vp[i].second; }
  return n;
}

// Fuse(KeyFrame*, vpMapPoints, th) (:751-893).  After the call: kf_assign[j] = point index in the keyframe's slot j;
// replaced_by[p] = index of the point that replaced p (-1: none).  order may contain -1 (NULL entries).
int refm_fuse(const refm_frame* kfd, const refm_points* pd, const int32_t* order, int n_order, float th,
              int32_t* kf_assign, int32_t* replaced_by) {
  Points P(pd);
  unique_ptr<KeyFrame> kf = make_keyframe(kfd, P);
  for (int j = 0; j < kfd->n; j++)
    if (kf->mvpMapPoints[j]) kf->mvpMapPoints[j]->mObservations[kf.get()] = j;
  vector<MapPoint*> v;
  for (int i = 0; i < n_order; i++) v.push_back(P.at(order[i]));
  ORBmatcher m(0.6f, true);
  const int n = m.Fuse(kf.get(), v, th);
  flatten(kf->mvpMapPoints, kf_assign);
  for (int i = 0; i < pd->n; i++) replaced_by[i] = P.at(i)->mpReplaced ? P.at(i)->mpReplaced->index : -1;
  return n;
}

// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:895-1009)
int refm_fuse_sim3(const refm_frame* kfd, const float* Scw, const refm_points* pd, const int32_t* order, int n_order,
                   float th, int32_t* kf_assign, int32_t* replace_point) {
  Points P(pd);
  unique_ptr<KeyFrame> kf = make_keyframe(kfd, P);
  vector<MapPoint*> v, rp(n_order, nullptr);
  for (int i = 0; i < n_order; i++) v.push_back(P.at(order[i]));
  ORBmatcher m(0.6f, true);
  const int n = m.Fuse(kf.get(), pose_of(Scw), v, th, rp);
  flatten(kf->mvpMapPoints, kf_assign);
  flatten(rp, replace_point);
  return n;
}

// SearchBySim3 (:1011-1221); matches12: in/out, point index (of KF2's points) per KF1 keypoint
int refm_search_by_sim3(const refm_frame* k1d, const refm_frame* k2d, const refm_points* pd, int32_t* matches12,
                        float s12, const float* R12, const float* t12, float th) {
  Points P(pd);
  unique_ptr<KeyFrame> k1 = make_keyframe(k1d, P), k2 = make_keyframe(k2d, P);
  for (int j = 0; j < k2d->n; j++)
    if (k2->mvpMapPoints[j]) k2->mvpMapPoints[j]->mObservations[k2.get()] = j;
  vector<MapPoint*> v(k1d->n);
  for (int i = 0; i < k1d->n; i++) v[i] = P.at(matches12[i]);
  cv::Mat R(3, 3, CV_32F);
  for (int i = 0; i < 9; i++) R.at<float>(i / 3, i % 3) = R12[i];
  ORBmatcher m(0.75f, true);
  const int n = m.SearchBySim3(k1.get(), k2.get(), v, s12, R, vec3(t12), th);
  flatten(v, matches12);
  return n;
}

// Frame::ComputeStereoMatches (code/src/Frame.cc:516-690, the reference's own body).  Keypoints as (x, y, octave)
// triples; plane[l] = ROI pixel (0,0) of the un-blurred pyramid level l (a view on a bordered plane, like the
// reference's mvImagePyramid: its window reads may reach into the border).
int refm_stereo_matches(const float* lx, const float* ly, const int32_t* loct, const uint8_t* ldesc, int nl,
                        const float* rx, const float* ry, const int32_t* roct, const uint8_t* rdesc, int nr,
                        const uint8_t* const* lplane, const uint8_t* const* rplane, const int32_t* stride, const int32_t* w,
                        const int32_t* h, int nlevels, const float* scale_factors, const float* inv_scale_factors, float mbf,
                        float mb, float* u_right, float* depth) {
  Frame F;
  ORBextractor exl, exr;
  for (int l = 0; l < nlevels; l++) {
    exl.mvImagePyramid.push_back(cv::cuda::GpuMat(h[l], w[l], const_cast<unsigned char*>(lplane[l]), (size_t)stride[l]));
    exr.mvImagePyramid.push_back(cv::cuda::GpuMat(h[l], w[l], const_cast<unsigned char*>(rplane[l]), (size_t)stride[l]));
  }
  F.mpORBextractorLeft = &exl;
  F.mpORBextractorRight = &exr;
  F.N = nl;
  F.mvKeys.resize(nl);
  for (int i = 0; i < nl; i++) {
    F.mvKeys[i].pt.x = lx[i]; F.mvKeys[i].pt.y = ly[i]; F.mvKeys[i].octave = loct[i];
  }
  F.mvKeysRight.resize(nr);
  for (int i = 0; i < nr; i++) {
    F.mvKeysRight[i].pt.x = rx[i]; F.mvKeysRight[i].pt.y = ry[i]; F.mvKeysRight[i].octave = roct[i];
  }
  F.mDescriptors = cv::Mat(nl, 32, CV_8U);
  if (nl) memcpy(F.mDescriptors.data, ldesc, (size_t)nl * 32);
  F.mDescriptorsRight = cv::Mat(nr, 32, CV_8U);
  if (nr) memcpy(F.mDescriptorsRight.data, rdesc, (size_t)nr * 32);
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  F.mvInvScaleFactors.assign(inv_scale_factors, inv_scale_factors + nlevels);
  F.mbf = mbf;
  F.mb = mb;
  F.ComputeStereoMatches();
  int kept = 0;
  for (int i = 0; i < nl; i++) {
    u_right[i] = F.mvuRight[i];
    depth[i] = F.mvDepth[i];
    kept += F.mvuRight[i] >= 0;
  }
  return kept;
}

#pragma GCC visibility pop
}  // extern "C"
