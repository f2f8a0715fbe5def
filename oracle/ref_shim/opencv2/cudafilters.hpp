#pragma once
#include <opencv2/core/cuda.hpp>
namespace cv { namespace cuda {
class Filter {
 public:
  virtual ~Filter() {}
  virtual void apply(InputArray src, OutputArray dst, Stream& stream = Stream::Null()) = 0;
};
Ptr<Filter> createGaussianFilter(int srcType, int dstType, Size ksize, double sigma1, double sigma2 = 0,
                                 int rowBorderMode = BORDER_REFLECT_101, int columnBorderMode = -1);
} }
