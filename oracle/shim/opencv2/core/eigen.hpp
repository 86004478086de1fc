// oracle/shim/opencv2/core/eigen.hpp -- TEST INFRASTRUCTURE ONLY.
// cv::eigen2cv as PnPsolver::Refine calls it (PnPsolver.cpp:230-231): the destination is never read;
// cv::cv2eigen as MLPnPsolver uses it.
#pragma once
#include "core.hpp"
namespace cv {
template <class M> inline void eigen2cv(const M &src, Mat &dst) { dst.rows = src.rows(); dst.cols = src.cols(); }
// cv::cv2eigen (MLPnPsolver.cpp:140-141): element-wise copy with conversion to the destination's scalar
template <class M> inline void cv2eigen(const Mat &src, M &dst)
{
    for (int i = 0; i < src.rows; ++i)
        for (int j = 0; j < src.cols; ++j) dst(i, j) = static_cast<typename M::Scalar>(src.get(i, j));
}
}  // namespace cv
