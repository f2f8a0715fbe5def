/*
 * orb_oracle.h -- C interface of the CPU ORACLE for the SwarmMap ORB front-end.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a plain-CPU restatement of the
 * reference's hot path (code/src/ORBextractor.cc, code/src/cuda/Fast_gpu.cu,
 * code/src/cuda/Orb_gpu.cu, code/src/ORBmatcher.cc, code/src/Frame.cc grid) and
 * exists only so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs can CHECK the CUDA product against it.
 * Nothing under swarmmap_b200/ may include, link or call it.
 *
 * PARITY STATUS: the reference ships no tests / golden vectors for this path and cannot be built as a whole here
 * (OpenCV-CUDA, Boost, ...) -- but every source file ON the path compiles on its own against
 * stand-in headers, and every function below is checked against output of the REFERENCE'S OWN CODE compiled
 * unmodified (`make ref` -> oracle/_ref/):
 *   ORBextractor.cc (constructor tables, DistributeOctTree)      tests/test_ref_orbextractor.py
 *   Thirdparty/DBoW2 (transform, FORB::distance)                 tests/test_ref_dbow2.py
 *   ORBmatcher.cc with its own ORBmatcher.h, the grid / scale bodies of Frame.cc, KeyFrame.cc, MapPoint.cc:
 *     every Search*, Fuse, SearchBySim3, SearchForTriangulation  tests/test_ref_orbmatcher.py
 *   Frame.cc:516-690 (ComputeStereoMatches, body cut at build time)  tests/test_ref_stereo.py
 *   cuda/Fast_gpu.cu, cuda/Orb_gpu.cu (nvcc, sm_100a, -use_fast_math): corner test + score, tile retry + NMS
 *     (single-tile launches of the real kernel and a lock-step run of its device functions), IC_Angle, rBRIEF
 *                                                                tests/test_gpu_ref_cuda.py (needs a GPU)
 * What the reference delegates to an un-vendored OpenCV is pinned to cv2 4.13 instead: resize INTER_LINEAR 8U,
 * copyMakeBorder REFLECT_101, GaussianBlur 7x7 sigma 2 8U, the FAST-9/16 score, undistortPoints and the cv::gemm /
 * cv::norm accumulation rules of the projections (tests/test_oracle_cv2.py, test_cv_shim_numerics_match_cv2).
 * Kept from round 1: closed-form tables, golden fixtures under tests/golden/, second independent numpy readings.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order/size as cv::KeyPoint (28 bytes). */
typedef struct orc_keypoint {
  float x, y;
  float size;
  float angle;
  float response;
  int32_t octave;
  int32_t class_id;
} orc_keypoint;

/* Raw FAST candidate in ROI coordinates (ROI origin = level pixel (16,16)). */
typedef struct orc_fastpt {
  int16_t x, y;
  int32_t score;
} orc_fastpt;

/* ---- tables (ORBextractor.cc:340-405) ---- */
void orc_scale_tables(float scale_factor, int nlevels, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2);
void orc_level_quotas(int nfeatures, float scale_factor, int nlevels, int* quotas);
void orc_umax(int* umax16);
void orc_level_sizes(int w, int h, float scale_factor, int nlevels, int* ws, int* hs);

/* ---- primitives ---- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_border_reflect101(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int border);
void orc_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
/* S(p): largest t >= min_th at which p is a FAST-9/16 corner, 0 if none.  Written for
 * 3 <= x < w-3, 3 <= y < h-3 of the given image; the 3-px rim is 0. */
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, int min_th, uint8_t* score, int score_stride);
/* Tile retry + NMS on a score map of an ROI (deterministic lock-step definition).
 * Returns the number of keypoints written (raster order, capped at `cap`). */
int orc_fast_tile_select(const uint8_t* score, int w, int h, int score_stride, int ini_th, orc_fastpt* out, int cap,
                         uint8_t* tile_retry /* optional, ceil(w/32)*ceil(h/32) */);
/* Quadtree distribution.  pts in ROI coords (raster order); returns count; out in final list order. */
int orc_octree(const orc_fastpt* pts, int n, int min_x, int max_x, int min_y, int max_y, int target, orc_fastpt* out,
               int out_cap);
float orc_ic_angle(const uint8_t* img, int stride, int x, int y);
void orc_ic_moments(const uint8_t* img, int stride, int x, int y, int* m10, int* m01);
void orc_rbrief(const uint8_t* img, int stride, int x, int y, float angle_deg, uint8_t* desc32);
int orc_hamming256(const uint8_t* a, const uint8_t* b);

/* ---- full extractor ---- */
typedef struct orc_extractor orc_extractor;
orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orc_extractor_destroy(orc_extractor* e);
/* Returns number of keypoints (<= cap) or -1 on bad arguments; 0 for an empty image. */
int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, int stride, orc_keypoint* kps, uint8_t* desc,
                int cap);
/* Introspection after orc_extract (for stage-by-stage parity). which: 0 = bordered un-blurred plane
 * ((w+38)x(h+38)), 1 = blurred ROI (w x h), 2 = FAST score map S at min_th on level coords (w x h). */
int orc_extractor_level(orc_extractor* e, int level, int which, const uint8_t** data, int* w, int* h, int* stride);
/* Per-level FAST candidates (pre-quadtree, raster order, ROI coords) and post-quadtree selection. */
int orc_extractor_level_fast(orc_extractor* e, int level, const orc_fastpt** pts);
int orc_extractor_level_selected(orc_extractor* e, int level, const orc_fastpt** pts);

/* ---- matcher support: Frame grid (Frame.cc:277-292,377-442) ---- */
typedef struct orc_frame {
  int n;
  const float* x;        /* undistorted keypoint coords */
  const float* y;
  const int32_t* octave;
  const float* angle;
  const uint8_t* desc;   /* n x 32 */
  float min_x, min_y, max_x, max_y; /* mnMinX.. image bounds */
} orc_frame;

typedef struct orc_grid orc_grid;
orc_grid* orc_grid_build(const orc_frame* f);
void orc_grid_destroy(orc_grid* g);
/* CSR export: cell index = ix*48+iy; starts has 64*48+1 entries. */
void orc_grid_csr(const orc_grid* g, int32_t* starts, int32_t* items);
int orc_grid_query(const orc_grid* g, const orc_frame* f, float x, float y, float r, int min_level, int max_level,
                   int32_t* out, int cap);

/* ---- matchers (ORBmatcher.cc) ---- */
/* M1 SearchForInitialization (:375-479).  prev_xy (n1 x 2) is read and updated. */
int orc_search_for_initialization(const orc_frame* f1, const orc_frame* f2, float* prev_xy, int32_t* matches12,
                                  int window, float nnratio, int check_ori);

/* Generic windowed projection matcher covering the SearchByProjection overloads
 * (:44-121 M3, :264-373 loop, :1223-1354 M2, :1356-1473 reloc).  The caller (C++ wrapper) has
 * already projected each source point; see include/swm_orb.h swm_match_window for field meaning. */
typedef struct orc_window_query {
  int m;
  const uint8_t* desc;       /* m x 32 */
  const float* u;
  const float* v;
  const float* radius;
  const int32_t* min_level;
  const int32_t* max_level;
  const uint8_t* valid;
  const float* angle;        /* may be NULL when check_ori == 0 */
  const uint8_t* blocks;     /* 1: once assigned, the target slot is unavailable to later sources */
} orc_window_query;

int orc_match_window(const orc_frame* tgt, const orc_window_query* q, const uint8_t* tgt_blocked_init, int th_dist,
                     int ratio_mode, float nnratio, int check_ori, int32_t* assignment /* n2, in/out */);

/* M4 SearchByBoW.  Feature vectors as CSR sorted by node id.  mode 0 = KF-Frame (:150-262, best <= TH_LOW),
 * mode 1 = KF-KF (:481-597, best < TH_LOW, side-2 validity + vbMatched2). */
typedef struct orc_featvec {
  int n_nodes;
  const uint32_t* node_ids; /* ascending */
  const int32_t* offsets;   /* n_nodes+1 */
  const uint32_t* feats;    /* feature indices, ascending within node */
} orc_featvec;

int orc_search_by_bow(const orc_frame* f1, const orc_featvec* fv1, const uint8_t* valid1, const orc_frame* f2,
                      const orc_featvec* fv2, const uint8_t* valid2, int mode, float nnratio, int check_ori,
                      int32_t* matches /* mode 0: size n2 -> idx1 or -1; mode 1: size n1 -> idx2 or -1 */);

/* ORBmatcher::SearchForTriangulation (:599-749), monocular (mvuRight < 0, bOnlyStereo = false), with
 * CheckDistEpipolarLine (:131-148).  valid1 / valid2: 1 = the keypoint has no MapPoint yet (:640-643, :662-666).
 * F12: 3x3 row-major; (ex, ey): epipole of KF1's centre in KF2 (:605-611); scale_factors2 / level_sigma2: pKF2's
 * tables.  Note the reference never sets vbMatched2, so rows are independent.  Float expressions are evaluated
 * left to right without contraction.  matches12: n1 entries (idx2 or -1).  Returns nmatches. */
int orc_search_for_triangulation(const orc_frame* f1, const orc_featvec* fv1, const uint8_t* valid1, const orc_frame* f2,
                                 const orc_featvec* fv2, const uint8_t* valid2, const float* F12, float ex, float ey,
                                 const float* scale_factors2, const float* level_sigma2, int check_ori,
                                 int32_t* matches12);

/* The search loop of ORBmatcher::Fuse (:824-870) / SearchBySim3 (:1098-1134), one independent row per query:
 * candidates = GetFeaturesInArea(u, v, radius) (Frame.cc:377-425) filtered to levels [pred - 1, pred] (:840-842); with
 * chi2 > 0 the monocular gate e2 * inv_level_sigma2[level] > chi2 -> skip (:857-864); dist < bestDist (first wins).
 * best_idx -1 / best_dist 256 when nothing survives. */
void orc_window_best(const orc_frame* tgt, int m, const uint8_t* desc, const float* u, const float* v, const float* radius,
                     const int32_t* pred_level, const uint8_t* valid, const float* inv_level_sigma2, float chi2,
                     int32_t* best_idx, int32_t* best_dist);

/* MapPoint::ComputeDistinctiveDescriptors (code/src/MapPoint.cc:361-391) for a batch of MapPoints (CSR offsets):
 * best_idx[p] = descriptor with the least median distance to the rest (-1 for an empty point). */
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* best_idx,
                                 int32_t* best_median);

/* ---- Frame::ComputeStereoMatches (Frame.cc:516-690): for every left keypoint the right keypoint of the same image row
 * band with the smallest descriptor distance (octave within +-1, disparity in [-3, mbf/mb]), refined by an 11x11 SAD
 * search over +-5 px in mvImagePyramid[level of the left keypoint] and a parabola fit, then the 1.5 * 1.4 * median
 * SAD filter.  plane[l] points at ROI pixel (0,0) of level l as operator() leaves it -- blurred in place inside its un-blurred
 * 19-px border (reads may reach into the border, like the reference's
 * colRange on a ROI GpuMat).  Outputs mvuRight / mvDepth (-1 = no match).  Returns the number of matches kept. */
typedef struct orc_pyramid {
  int nlevels;
  const uint8_t* const* plane;
  const int32_t* stride;
  const int32_t* w;
  const int32_t* h;
} orc_pyramid;
int orc_stereo_matches(const orc_keypoint* kl, const uint8_t* dl, int nl, const orc_keypoint* kr, const uint8_t* dr, int nr,
                       const orc_pyramid* pl, const orc_pyramid* pr, const float* scale_factors,
                       const float* inv_scale_factors, float mbf, float mb, float* u_right, float* depth);

/* Brute-force top-2 of each query against a descriptor database (config 5). out: per query
 * (dist0, idx0, dist1, idx1); ties broken by lower index. */
void orc_bruteforce_top2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int32_t* out4);

/* Frame::UndistortKeyPoints (code/src/Frame.cc:454-484): cv::undistortPoints(pts, pts, K, D, Mat(), K) on float32
 * points.  OpenCV is a third-party dependency absent from /root/reference (the reference links the system OpenCV 3.x);
 * the algorithm restated here is cvUndistortPointsInternal with the default criteria (5 fixed-point iterations, all
 * arithmetic in double, float in/out).  k4,k5,k6, thin-prism and tilt terms are zero (ORB-SLAM passes 4 or 5
 * coefficients).  cam9 = fx, fy, cx, cy, k1, k2, p1, p2, k3.  With k1 == 0 the points are copied (:456-460).  Pinned
 * against cv2.undistortPoints in tests/test_oracle_cv2.py. */
void orc_undistort_points(const float* xy, int n, const float* cam9, float* out_xy);
/* Frame::ComputeImageBounds (Frame.cc:486-514): out4 = mnMinX, mnMaxX, mnMinY, mnMaxY. */
void orc_image_bounds(int cols, int rows, const float* cam9, float* out4);

/* ---- DBoW2 (vendored in the reference: code/Thirdparty/DBoW2/DBoW2; needs OpenCV, so it cannot be compiled here) ----
 * Vocabulary from the binary layout read by TemplatedVocabulary::loadFromBinaryFile (TemplatedVocabulary.h:1478-1522):
 * header nb_nodes, size_node, k, L, scoring, weighting, then per node (ids 1..): int32 parent, 32 descriptor bytes,
 * float weight, uint8 is_leaf.  Children lists in file order, word ids = leaves in file order. */
typedef struct orc_vocab orc_vocab;
orc_vocab* orc_vocab_load(const uint8_t* blob, size_t bytes);
void orc_vocab_destroy(orc_vocab* v);
void orc_vocab_info(const orc_vocab* v, int32_t* k, int32_t* L, int32_t* n_nodes, int32_t* n_words);
/* TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) (:1151-1218, per feature :1242-1283;
 * FORB::distance FORB.cpp:82-102; BowVector::addWeight / addIfNotExist / normalize BowVector.cpp:33-90;
 * FeatureVector::addFeature FeatureVector.cpp:31-45).  Outputs sized for n entries; offsets n + 1.
 * Returns the number of words; *n_nodes_out the number of FeatureVector nodes. */
int orc_bow_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_ids,
                      double* word_values, uint32_t* node_ids, int32_t* offsets, uint32_t* feats, int32_t* n_nodes_out);
double orc_time_bow_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, int iters);

/* ---- timing helpers for bench.py's cpu_baseline (run entirely on the CPU) ---- */
double orc_time_extract(orc_extractor* e, const uint8_t* imgs, int n_imgs, int w, int h, int iters);

#ifdef __cplusplus
}
#endif
#endif
