#pragma once
#include <opencv2/core/cuda.hpp>
namespace cv { namespace cuda {
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR,
            Stream& stream = Stream::Null());
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    Scalar value = Scalar(), Stream& stream = Stream::Null());
} }
