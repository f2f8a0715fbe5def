// ref_doubles.h -- test doubles of ORB_SLAM2::MapPoint / Frame / KeyFrame for compiling the reference's
// code/src/ORBmatcher.cc UNMODIFIED (test infrastructure, see oracle/Makefile target `ref`).
//
// How the build uses it: the compiler is given -DFRAME_H -DKEYFRAME_H -DMAPPOINT_H (the include guards of the
// reference's own headers, code/include/Frame.h:21, KeyFrame.h:21, MapPoint.h:21) and `-include ref_doubles.h`, so the
// reference's real ORBmatcher.h (code/include/ORBmatcher.h) is used as it is and its three #includes come up empty.
// The doubles declare exactly the members ORBmatcher.cc reads or calls, with the reference's names and types.
//
// The grid / scale functions are NOT restated here: their bodies are taken verbatim from the reference at build time
// (oracle/Makefile cuts code/src/Frame.cc:277-292,377-442, code/src/KeyFrame.cc:779-818 and
// code/src/MapPoint.cc:466-485 into oracle/_ref/gen/*.inc, git-ignored) and compiled against these declarations by
// oracle/ref_orbmatcher_wrap.cpp.
#pragma once
#ifndef FRAME_H
#define FRAME_H
#endif
#ifndef KEYFRAME_H
#define KEYFRAME_H
#endif
#ifndef MAPPOINT_H
#define MAPPOINT_H
#endif

#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48  // code/include/Frame.h:37-38
#define FRAME_GRID_COLS 64

// The reference's headers leak `using namespace std` (ORBmatcher.h:72,80,83 write `pair<>` / `vector<>` unqualified).
using namespace std;

namespace ORB_SLAM2 {

class KeyFrame;
class Frame;

class MapPoint {
 public:
  // tracking fields written by Frame::isInFrustum (Frame.cc:362-367)
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
  bool mbTrackInView = false;
  int mnTrackScaleLevel = 0;
  float mTrackViewCos = 0;

  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  cv::Mat GetGlobalPos() { return mWorldPos.clone(); }
  cv::Mat GetNormal() { return mNormalVector.clone(); }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  int Observations() { return nObs; }
  bool isBad() { return mbBad; }
  bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) > 0; }
  int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
  void AddObservation(KeyFrame* pKF, size_t idx) {
    if (mObservations.count(pKF)) return;
    mObservations[pKF] = idx;
    nObs++;
  }
  void Replace(MapPoint* pMP) { mbBad = true; mpReplaced = pMP; }
  // bodies from code/src/MapPoint.cc:466-485
  float GetMinDistanceInvariance();
  float GetMaxDistanceInvariance();
  int PredictScale(const float& currentDist, const float& logScaleFactor, const int mnScaleLevel);

  cv::Mat mWorldPos, mNormalVector, mDescriptor;
  int nObs = 0;
  bool mbBad = false;
  MapPoint* mpReplaced = nullptr;
  std::map<KeyFrame*, size_t> mObservations;
  float mfMinDistance = 0, mfMaxDistance = 0;
  std::mutex mMutexPos;
  int index = -1;  // position in the harness' point table
};

// what Frame::ComputeStereoMatches reads of the extractors (include/ORBextractor.h:90): the un-bordered pyramid views
class ORBextractor {
 public:
  std::vector<cv::cuda::GpuMat> mvImagePyramid;
};

class Frame {
 public:
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  // stereo members (code/include/Frame.h), read and written by ComputeStereoMatches, body from code/src/Frame.cc:516-690
  std::vector<cv::KeyPoint> mvKeysRight;
  cv::Mat mDescriptorsRight;
  std::vector<float> mvDepth, mvInvScaleFactors;
  ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
  void ComputeStereoMatches();
  std::vector<float> mvuRight;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mTcw, mRcw, mtcw, mOw;
  float mbf = 0, mb = 0;
  static float fx, fy, cx, cy;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  static float mfGridElementWidthInv, mfGridElementHeightInv;
  int mnScaleLevels = 0;
  float mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

  // bodies from code/src/Frame.cc:277-292 and :316-442
  void AssignFeaturesToGrid();
  bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                        const int maxLevel = -1) const;
  bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
};

class KeyFrame {
 public:
  int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvuRight;
  cv::Mat mDescriptors;
  DBoW2::FeatureVector mFeatVec;
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
  int mnScaleLevels = 0;
  float mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;  // ints in KeyFrame (code/include/KeyFrame.h), floats in Frame
  int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  std::vector<std::vector<std::vector<size_t> > > mGrid;

  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints)
      if (p && !p->isBad()) s.insert(p);
    return s;
  }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  cv::Mat GetRotation() { return Rcw.clone(); }
  cv::Mat GetTranslation() { return tcw.clone(); }
  // bodies from code/src/KeyFrame.cc:779-818
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
  bool IsInImage(const float& x, const float& y) const;

  std::vector<MapPoint*> mvpMapPoints;
  cv::Mat Ow, Rcw, tcw;
};

}  // namespace ORB_SLAM2
