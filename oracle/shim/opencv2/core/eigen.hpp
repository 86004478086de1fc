// oracle/shim/opencv2/core/eigen.hpp -- TEST INFRASTRUCTURE ONLY.
// cv::eigen2cv as PnPsolver::Refine calls it (PnPsolver.cpp:230-231): the destination is never read.
#pragma once
#include "core.hpp"
namespace cv {
template <class M> inline void eigen2cv(const M &src, Mat &dst) { dst.rows = src.rows(); dst.cols = src.cols(); }
}  // namespace cv
