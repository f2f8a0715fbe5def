// opencv_build_check.cpp -- compile-only check of the SWM_HAVE_OPENCV branches of the drop-in headers
// (swarmmap_b200/host/ORBextractor.h: cv::InputArray / OutputArray, mvImagePyramid as cv::cuda::GpuMat headers).
// The images have no OpenCV C++ headers, so it is compiled against oracle/ref_shim's stand-ins, which declare the
// OpenCV signatures the header uses (tests/test_abi.py::test_opencv_branch_compiles, -fsyntax-only).
#include "../swarmmap_b200/host/ORBextractor.h"

int check(ORB_SLAM2::ORBextractor& ex, const cv::Mat& im, std::vector<cv::KeyPoint>& k, cv::Mat& d) {
  ex(im, cv::Mat(), k, d);
  const cv::cuda::GpuMat& lvl = ex.mvImagePyramid[0];      // what Frame.cc:611-612 takes
  const cv::cuda::GpuMat& brd = ex.mvImagePyramidBorder[0];
  return lvl.rows + lvl.cols + brd.rows + (int)ex.mvImagePyramidAllocatedFlag + ex.GetLevels();
}
