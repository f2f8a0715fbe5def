// host_wrapper_test.cpp -- compiles the drop-in C++ wrapper classes (swarmmap_b200/host/) against the
// cv shim with small test doubles of Frame / KeyFrame / MapPoint that carry exactly the members the
// reference's ORBmatcher touches.  Linked against libswm_orb.so; run on the GPU box by tests/test_gpu_host.py.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "../swarmmap_b200/host/ORBextractor.h"
#include "../swarmmap_b200/host/ORBmatcher.h"
#include "../swarmmap_b200/host/ResidentFrame.h"
#include "../swarmmap_b200/host/ORBVocabulary.h"

struct Mat4 {  // stand-in for the CV_32F cv::Mat pose / position
  float v[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  int n = 16;
  template <typename T> T at(int r, int c) const { return v[r * 4 + c]; }
  template <typename T> T at(int i) const { return v[i]; }
};

struct MapPoint {
  Mat4 pos;
  cv::Mat desc;
  int nobs = 1;
  bool bad = false;
  bool mbTrackInView = true;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 1.0f;
  int mnTrackScaleLevel = 0;
  Mat4 GetWorldPos() { return pos; }
  Mat4 GetGlobalPos() { return pos; }
  Mat4 GetNormal() { Mat4 n; n.v[0] = 0; n.v[1] = 0; n.v[2] = 1; return n; }
  float GetMinDistanceInvariance() { return 0.1f; }
  float GetMaxDistanceInvariance() { return 100.f; }
  int PredictScale(float, float, int) { return mnTrackScaleLevel; }
  cv::Mat GetDescriptor() { return desc; }
  int Observations() { return nobs; }
  bool isBad() { return bad; }
  // Fuse bookkeeping
  std::map<void*, int> obs;
  MapPoint* replaced_by = nullptr;
  bool IsInKeyFrame(void* kf) { return obs.count(kf) > 0; }
  int GetIndexInKeyFrame(void* kf) { auto it = obs.find(kf); return it == obs.end() ? -1 : it->second; }
  void AddObservation(void* kf, int idx) { obs[kf] = idx; nobs = (int)obs.size(); }
  void Replace(MapPoint* other) { bad = true; replaced_by = other; }
};

struct Frame {
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  Mat4 mTcw;
  static constexpr float fx = 458.654f, fy = 457.296f, cx = 367.215f, cy = 248.375f;
  static constexpr float mnMinX = 0, mnMinY = 0, mnMaxX = 752, mnMaxY = 480;
  float mfLogScaleFactor = 0.1823f;
  int mnScaleLevels = 8;
  bool IsInImage(float x, float y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
  std::vector<float> mvScaleFactors;
  std::map<unsigned, std::vector<unsigned>> mFeatVec;
  const swm_frame* mpResident = nullptr;  // set when the features also live on the device (host/ResidentFrame.h)
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  // KeyFrame accessors used by SearchForTriangulation
  std::vector<float> mvLevelSigma2;
  Mat4 mOw, mRcw, mtcw;
  MapPoint* GetMapPoint(int i) { return mvpMapPoints[i]; }
  void AddMapPoint(MapPoint* p, int i) { mvpMapPoints[i] = p; }
  std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (auto* p : mvpMapPoints) if (p) s.insert(p); return s; }
  std::vector<float> mvInvLevelSigma2;
  Mat4 GetCameraCenter() { return mOw; }
  Mat4 GetRotation() { return mRcw; }
  Mat4 GetTranslation() { return mtcw; }
};

static void fill(Frame& f, ORB_SLAM2::ORBextractor& ex, const cv::Mat& img) {
  ex(img, cv::Mat(), f.mvKeys, f.mDescriptors);
  f.mvKeysUn = f.mvKeys;
  f.N = (int)f.mvKeys.size();
  f.mvpMapPoints.assign(f.N, nullptr);
  f.mvbOutlier.assign(f.N, false);
  f.mvScaleFactors = ex.GetScaleFactors();
  for (int i = 0; i < f.N; i++) f.mFeatVec[f.mDescriptors.ptr(i)[0] & 31].push_back(i);
}

int main(int argc, char** argv) {
  const int w = 752, h = 480;
  cv::Mat img(h, w, CV_8U);
  unsigned s = 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {  // blocky texture with corners
      s = s * 1664525u + 1013904223u;
      img.ptr(y)[x] = (unsigned char)(((x / 23 + y / 17) % 2 ? 60 : 180) + ((x / 7 * 31 + y / 5 * 17) % 40) + (s >> 29));
    }
  ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
  Frame f1, f2;
  fill(f1, ex, img);
  fill(f2, ex, img);
  std::printf("keypoints %d levels %d pyramid0 %dx%d allocated %d\n", f1.N, ex.GetLevels(), ex.mvImagePyramid[0].cols,
              ex.mvImagePyramid[0].rows, (int)ex.mvImagePyramidAllocatedFlag);
  if (f1.N < 300) return 1;
  ORB_SLAM2::ORBmatcher m(0.9f, true);
  std::vector<cv::Point2f> prev(f1.N);
  for (int i = 0; i < f1.N; i++) prev[i] = f1.mvKeysUn[i].pt;
  std::vector<int> m12;
  const int n_init = m.SearchForInitialization(f1, f2, prev, m12, 100);
  // give frame 1 map points in front of an identity camera that project onto their own keypoints
  std::vector<MapPoint> pts(f1.N);
  for (int i = 0; i < f1.N; i++) {
    const float z = 2.0f;
    pts[i].pos.v[0] = (f1.mvKeysUn[i].pt.x - f1.cx) / f1.fx * z;
    pts[i].pos.v[1] = (f1.mvKeysUn[i].pt.y - f1.cy) / f1.fy * z;
    pts[i].pos.v[2] = z;
    pts[i].desc = f1.mDescriptors.row(i);
    pts[i].mTrackProjX = f1.mvKeysUn[i].pt.x;
    pts[i].mTrackProjY = f1.mvKeysUn[i].pt.y;
    pts[i].mnTrackScaleLevel = f1.mvKeysUn[i].octave;
    f1.mvpMapPoints[i] = &pts[i];
  }
  const int n_proj = m.SearchByProjection(f2, f1, 15.0f, true);
  int self = 0;
  for (int j = 0; j < f2.N; j++) self += f2.mvpMapPoints[j] == &pts[j];
  Frame f3;
  fill(f3, ex, img);
  std::vector<MapPoint*> local;
  for (auto& p : pts) local.push_back(&p);
  ORB_SLAM2::ORBmatcher m3(0.8f, true);
  const int n_mp = m3.SearchByProjection(f3, local, 1.0f);
  std::vector<MapPoint*> bow;
  ORB_SLAM2::ORBmatcher mb(0.7f, true);
  const int n_bow = mb.SearchByBoW(&f1, f3, bow);
  std::vector<MapPoint*> bow2;
  const int n_bow2 = mb.SearchByBoW(&f1, &f2, bow2);
  Frame f4;
  fill(f4, ex, img);
  std::set<MapPoint*> already;
  ORB_SLAM2::ORBmatcher mr(0.9f, true);
  const int n_reloc = mr.SearchByProjection(f4, &f1, already, 10.0f, 100);
  Frame f5;
  fill(f5, ex, img);
  Mat4 Scw;
  std::vector<MapPoint*> matched(f5.N, nullptr);
  ORB_SLAM2::ORBmatcher ml(0.75f, true);
  const int n_loop = ml.SearchByProjection(&f5, Scw, local, matched, 10);
  std::printf("reloc %d loop %d\n", n_reloc, n_loop);
  if (n_reloc < f1.N / 2 || n_loop < f1.N / 2) { std::printf("HOST_WRAPPER_FAIL\n"); return 1; }
  // Resident path: the same frame built on the device from the extractor's output must give the host path's results
  Frame f6;
  fill(f6, ex, img);  // the extractor's last batch is now f6's image
  const swm_camera cam = {Frame::fx, Frame::fy, Frame::cx, Frame::cy, 0.f, 0.f, 0.f, 0.f, 0.f};
  float bounds[4];
  ORB_SLAM2::ResidentFrame::ComputeImageBounds(cam, w, h, bounds);
  ORB_SLAM2::ResidentFrame rf6, rf1;
  rf6.FromExtractor(ex, 0, &cam, bounds);
  std::vector<cv::KeyPoint> un;
  rf6.DownloadUndistorted(f6.mvKeys, un);
  bool same_un = rf6.size() == f6.N && bounds[0] == 0 && bounds[1] == w && bounds[2] == 0 && bounds[3] == h;
  for (int i = 0; i < f6.N && same_un; i++) same_un = un[i].pt.x == f6.mvKeys[i].pt.x && un[i].pt.y == f6.mvKeys[i].pt.y;
  f6.mpResident = rf6.get();
  ORB_SLAM2::ORBmatcher m6(0.9f, true);
  const int n_proj_res = m6.SearchByProjection(f6, f1, 15.0f, true);  // f6 (== f2's image) matched in place on the GPU
  int same_proj = 0;
  for (int j = 0; j < f6.N; j++) same_proj += f6.mvpMapPoints[j] == f2.mvpMapPoints[j];
  rf1.FromExtractor(ex, 0, &cam, bounds);  // same image as f1: both operands resident for the BoW matcher
  f1.mpResident = rf1.get();
  std::vector<MapPoint*> bow_res;
  const int n_bow_res = mb.SearchByBoW(&f1, f6, bow_res);
  f1.mpResident = nullptr;
  std::printf("resident: undistort-identity %d proj %d (host %d) same slots %d/%d bow %d (host %d)\n", (int)same_un, n_proj_res,
              n_proj, same_proj, f6.N, n_bow_res, n_bow);
  if (!same_un || n_proj_res != n_proj || same_proj != f6.N || n_bow_res != n_bow) { std::printf("HOST_WRAPPER_FAIL\n"); return 1; }
  // Stereo: the right view is the left view moved 12 px to the left; ComputeStereoMatches must recover that disparity
  {
    cv::Mat imgR(h, w, CV_8U);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) imgR.ptr(y)[x] = img.ptr(y)[x + 12 < w ? x + 12 : w - 1];
    ORB_SLAM2::ORBextractor exL(1000, 1.2f, 8, 20, 7), exR(1000, 1.2f, 8, 20, 7);
    Frame fl, fr;
    fill(fl, exL, img);
    fill(fr, exR, imgR);
    std::vector<float> uRight, depth;
    const float mbf = 47.9f, mb = 47.9f / 458.654f;
    exL.ComputeStereoMatches(exR, mbf, mb, fl.N, uRight, depth);
    int n_st = 0, n_ok = 0;
    for (int i = 0; i < fl.N; i++)
      if (uRight[i] >= 0) {
        n_st++;
        const float d = fl.mvKeys[i].pt.x - uRight[i];
        n_ok += d > 10.5f && d < 13.5f && depth[i] == mbf / d;
      }
    std::printf("stereo: %d of %d left keypoints matched, %d with the planted disparity\n", n_st, fl.N, n_ok);
    if (n_st < fl.N / 4 || n_ok < n_st * 9 / 10) { std::printf("HOST_WRAPPER_FAIL stereo\n"); return 1; }
  }
  // Vocabulary: a two-level synthetic tree in the ORBvoc.bin layout; transform() must fill DBoW2's containers
  {
    const int k = 4;
    std::vector<uint8_t> blob(24 + (size_t)(k + k * k) * 41, 0);
    const uint32_t nb = k + k * k, sz = 41;
    const int32_t hdr[4] = {k, 2, 0, 0};  // k, L, L1_NORM, TF_IDF
    std::memcpy(&blob[0], &nb, 4); std::memcpy(&blob[4], &sz, 4); std::memcpy(&blob[8], hdr, 16);
    unsigned s2 = 7;
    for (uint32_t r = 0; r < nb; r++) {
      uint8_t* rec = &blob[24 + (size_t)r * 41];
      const int32_t parent = r < (uint32_t)k ? 0 : (int32_t)((r - k) / k + 1);
      std::memcpy(rec, &parent, 4);
      for (int b = 0; b < 32; b++) { s2 = s2 * 1664525u + 1013904223u; rec[4 + b] = (uint8_t)(s2 >> 24); }
      const float w = 1.0f + (float)(r % 5);
      std::memcpy(rec + 36, &w, 4);
      rec[40] = r >= (uint32_t)k;
    }
    ORB_SLAM2::ORBVocabulary voc;
    if (!voc.loadFromMemory(blob.data(), blob.size()) || voc.size() != (unsigned)(k * k)) { std::printf("HOST_WRAPPER_FAIL vocab\n"); return 1; }
    std::vector<cv::Mat> feats;
    for (int i = 0; i < f1.N; i++) feats.push_back(f1.mDescriptors.row(i));
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fvv;
    voc.transform(feats, bv, fvv, 1);
    double sum = 0;
    size_t nfeat = 0;
    for (auto& e : bv) sum += e.second;
    for (auto& e : fvv) nfeat += e.second.size();
    std::printf("vocabulary: words %u bow entries %zu (L1 sum %.12f) nodes %zu features %zu\n", voc.size(), bv.size(), sum,
                fvv.size(), nfeat);
    if (bv.empty() || bv.size() > (size_t)(k * k) || fvv.size() > (size_t)k || nfeat != (size_t)f1.N || sum < 0.999999 || sum > 1.000001) {
      std::printf("HOST_WRAPPER_FAIL vocab transform\n");
      return 1;
    }
  }
  // SearchForTriangulation: two keyframes of the same image, a fundamental matrix for a sideways translation; every
  // keypoint lies on its own epipolar line (identical images), the epipole is far outside the image
  {
    Frame k1, k2;
    fill(k1, ex, img);
    fill(k2, ex, img);
    k2.mvLevelSigma2 = ex.GetScaleSigmaSquares();
    k1.mOw.v[0] = 1.0f; k1.mOw.v[1] = 0.0f; k1.mOw.v[2] = 0.0f;  // camera centre of KF1 in the world
    k2.mtcw.v[0] = 0.0f; k2.mtcw.v[1] = 0.0f; k2.mtcw.v[2] = 0.001f;
    // pure x-translation: F = K^-T [t]x K^-1 with t = (1, 0, 0): the epipolar line of (u, v) is the row v
    Mat4 F;
    for (int i = 0; i < 16; i++) F.v[i] = 0;
    F.v[1 * 4 + 2] = -1.0f / Frame::fy;           // F(1,2)
    F.v[2 * 4 + 1] = 1.0f / Frame::fy;            // F(2,1)
    std::vector<std::pair<size_t, size_t>> pairs;
    ORB_SLAM2::ORBmatcher mt(0.6f, false);
    const int n_tri = mt.SearchForTriangulation(&k1, &k2, F, pairs, false);
    int self_tri = 0;
    for (auto& p : pairs) self_tri += p.first == p.second;
    std::printf("triangulation: %d pairs, %d on the same keypoint of the identical image\n", n_tri, self_tri);
    if (n_tri != (int)pairs.size() || n_tri < k1.N / 2 || self_tri < n_tri * 9 / 10) { std::printf("HOST_WRAPPER_FAIL triangulation\n"); return 1; }
  }
  // Fuse: project frame 1's MapPoints (which sit exactly on its keypoints) into a keyframe of the same image: half of
  // the keyframe's slots are empty (-> AddObservation / AddMapPoint), half hold another point (-> Replace)
  {
    Frame kf;
    fill(kf, ex, img);
    kf.mvInvLevelSigma2 = ex.GetInverseScaleSigmaSquares();
    for (int i = 0; i < 3; i++) { kf.mtcw.v[i] = 0.0f; kf.mOw.v[i] = 0.0f; }  // identity pose: the points project onto their keypoints
    std::vector<MapPoint> fresh(pts.begin(), pts.end());  // copies: no observations yet
    for (auto& p : fresh) { p.obs.clear(); p.nobs = 0; p.bad = false; }
    std::vector<MapPoint> resident_pts(kf.N);
    for (int j = 0; j < kf.N; j += 2) { resident_pts[j].nobs = 3; kf.mvpMapPoints[j] = &resident_pts[j]; }
    std::vector<MapPoint*> cand;
    for (auto& p : fresh) cand.push_back(&p);
    ORB_SLAM2::ORBmatcher mf(0.8f, true);
    const int n_fused = mf.Fuse(&kf, cand, 3.0f);
    int added = 0, replaced = 0;
    for (int j = 0; j < kf.N; j++) added += (j % 2 == 1) && kf.mvpMapPoints[j] != nullptr;
    for (auto& p : fresh) replaced += p.bad && p.replaced_by != nullptr;
    std::printf("fuse: %d fused, %d added to empty slots, %d replaced by the keyframe's point\n", n_fused, added, replaced);
    if (n_fused < kf.N * 8 / 10 || added < kf.N / 2 * 7 / 10 || replaced < kf.N / 2 * 7 / 10 || added + replaced != n_fused) {
      std::printf("HOST_WRAPPER_FAIL fuse\n");
      return 1;
    }
  }
  // Fuse with a Sim3 pose (loop closing): identity similarity scaled by 2 (Scw = 2 [I|0]) projects the points onto
  // their keypoints; even slots hold a point (-> vpReplacePoint), odd slots are empty (-> AddMapPoint)
  {
    Frame kf;
    fill(kf, ex, img);
    std::vector<MapPoint> fresh(pts.begin(), pts.end()), held(kf.N);
    for (auto& p : fresh) { p.obs.clear(); p.nobs = 0; p.bad = false; }
    for (int j = 0; j < kf.N; j += 2) kf.mvpMapPoints[j] = &held[j];
    std::vector<MapPoint*> cand, repl(kf.N, nullptr);
    for (auto& p : fresh) cand.push_back(&p);
    Mat4 Scw;
    for (int i = 0; i < 16; i++) Scw.v[i] = 0;
    Scw.v[0] = Scw.v[5] = Scw.v[10] = 2.0f; Scw.v[15] = 1.0f;
    ORB_SLAM2::ORBmatcher mf2(0.8f, true);
    const int n_fused2 = mf2.Fuse(&kf, Scw, cand, 4.0f, repl);
    int n_repl = 0, n_added = 0;
    for (int j = 0; j < kf.N; j++) { n_repl += repl[j] != nullptr; n_added += (j % 2 == 1) && kf.mvpMapPoints[j] != nullptr; }
    std::printf("fuse(Scw): %d fused, %d to replace, %d added\n", n_fused2, n_repl, n_added);
    if (n_fused2 < kf.N * 8 / 10 || n_repl + n_added != n_fused2 || n_repl < kf.N / 2 * 7 / 10) { std::printf("HOST_WRAPPER_FAIL fuse sim3\n"); return 1; }
  }
  // SearchBySim3 with the identity similarity between two keyframes of the same image: every MapPoint of one keyframe
  // must find the same keypoint in the other, in both directions
  {
    Frame a, b;
    fill(a, ex, img);
    fill(b, ex, img);
    for (int i = 0; i < 3; i++) { a.mtcw.v[i] = 0.0f; b.mtcw.v[i] = 0.0f; }
    std::vector<MapPoint> pa(pts.begin(), pts.end()), pb(pts.begin(), pts.end());
    for (int i = 0; i < a.N; i++) { a.mvpMapPoints[i] = &pa[i]; b.mvpMapPoints[i] = &pb[i]; }
    std::vector<MapPoint*> m12(a.N, nullptr);
    Mat4 R12, t12;
    for (int i = 0; i < 3; i++) t12.v[i] = 0.0f;
    const float s12 = 1.0f;
    ORB_SLAM2::ORBmatcher ms(0.75f, true);
    const int n_sim3 = ms.SearchBySim3(&a, &b, m12, s12, R12, t12, 7.5f);
    int same = 0;
    for (int i = 0; i < a.N; i++) same += m12[i] == &pb[i];
    std::printf("sim3: %d mutual matches, %d onto the same keypoint\n", n_sim3, same);
    if (n_sim3 < a.N * 8 / 10 || same < n_sim3 * 9 / 10) { std::printf("HOST_WRAPPER_FAIL sim3\n"); return 1; }
  }
  const int d0 = ORB_SLAM2::ORBmatcher::DescriptorDistance(f1.mDescriptors.row(0), f1.mDescriptors.row(0));
  const int d1 = ORB_SLAM2::ORBmatcher::DescriptorDistance(f1.mDescriptors.row(0), f1.mDescriptors.row(1));
  std::printf("init %d proj %d self %d mappoints %d bow %d bowkf %d d0 %d d1 %d\n", n_init, n_proj, self, n_mp, n_bow,
              n_bow2, d0, d1);
  const bool ok = n_init > 50 && n_proj > f1.N / 2 && self > f1.N / 2 && n_mp > f1.N / 2 && n_bow > 20 && d0 == 0 && d1 > 0;
  std::printf("%s\n", ok ? "HOST_WRAPPER_OK" : "HOST_WRAPPER_FAIL");
  return ok ? 0 : 1;
}
