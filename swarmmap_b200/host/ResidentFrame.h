// ResidentFrame.h -- device-resident counterpart of the feature part of ORB_SLAM2::Frame.
//
// The reference's Frame constructor (code/src/Frame.cc:236-275) runs the extractor, copies mvKeys to the host,
// undistorts them on the CPU with cv::undistortPoints (UndistortKeyPoints, :454-484) and fills mGrid on the CPU
// (AssignFeaturesToGrid, :277-292); every SearchByProjection then starts from those host arrays.  With this class
// the extractor's output stays on the GPU: undistortion (bit-identical to cv::undistortPoints), SoA split and the
// grid are one kernel pair on the extractor's stream, and the ORBmatcher wrapper (host/ORBmatcher.h) matches against
// the frame in place when the Frame object carries `const swm_frame* mpResident`.
//
//   ORB_SLAM2::ResidentFrame rf;                        // member of Frame
//   (*mpORBextractorLeft)(imGray, cv::Mat(), mvKeys, mDescriptors);
//   rf.FromExtractor(*mpORBextractorLeft, 0, cam, bounds);   // replaces UndistortKeyPoints + AssignFeaturesToGrid
//   rf.DownloadUndistorted(mvKeys, mvKeysUn);           // only if the CPU side needs mvKeysUn (pose optimisation)
//   mpResident = rf.get();
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/swm_orb.h"
#include "ORBextractor.h"

namespace ORB_SLAM2 {

class ResidentFrame {
 public:
  explicit ResidentFrame(int device = 0) {
    if (swm_frame_create(device, &f_) != SWM_OK)
      throw std::runtime_error(std::string("ResidentFrame: ") + swm_frame_last_error(nullptr));
  }
  ~ResidentFrame() { swm_frame_destroy(f_); }
  ResidentFrame(const ResidentFrame&) = delete;
  ResidentFrame& operator=(const ResidentFrame&) = delete;

  // Frame::ComputeImageBounds (Frame.cc:486-514): mnMinX, mnMaxX, mnMinY, mnMaxY.
  static void ComputeImageBounds(const swm_camera& cam, int cols, int rows, float bounds4[4], int device = 0) {
    if (swm_camera_bounds(device, &cam, cols, rows, bounds4) != SWM_OK)
      throw std::runtime_error("ResidentFrame::ComputeImageBounds failed (no CUDA device? there is no CPU fallback)");
  }

  // Frame `index` of the extractor's most recent call -> undistorted keypoints + descriptors + grid on the device.
  void FromExtractor(ORBextractor& extractor, int index, const swm_camera* cam, const float bounds4[4]) {
    check(swm_frame_from_extractor(f_, extractor.handle(), index, cam, bounds4));
  }

  // mvKeysUn = mvKeys with pt replaced by the undistorted position (Frame.cc:475-483).
  template <class KeyPointT>
  void DownloadUndistorted(const std::vector<KeyPointT>& mvKeys, std::vector<KeyPointT>& mvKeysUn) {
    const int n = swm_frame_size(f_);
    if ((int)mvKeys.size() != n) throw std::runtime_error("ResidentFrame: mvKeys does not belong to this frame");
    std::vector<float> x(n), y(n);
    check(swm_frame_download(f_, x.data(), y.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
    mvKeysUn = mvKeys;
    for (int i = 0; i < n; i++) { mvKeysUn[i].pt.x = x[i]; mvKeysUn[i].pt.y = y[i]; }
  }

  int size() const { return swm_frame_size(f_); }
  const swm_frame* get() const { return f_; }

 private:
  void check(int rc) {
    if (rc != SWM_OK) throw std::runtime_error(std::string("ResidentFrame: ") + swm_frame_last_error(f_));
  }
  swm_frame* f_ = nullptr;
};

}  // namespace ORB_SLAM2
