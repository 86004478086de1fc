// oracle/shim/orbslam/Frame.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/Frame.hpp with the members PnPsolver's constructor reads
// (Frame.hpp:102-105 fx, fy, cx, cy as float; :127 mvKeysUn; :142 mvpMapPoints; mvLevelSigma2) and the two
// DetectRelocalizationCandidates reads (:135 mBowVec, :157 mnId).
#pragma once
#include <algorithm>
#include <limits>
#include "MapPoint.hpp"
#include <opencv2/core/eigen.hpp>
#include "Thirdparty/DBoW2/DUtils/Random.h"     // the reference's own header (reached through its real Frame.hpp include chain)
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"    // the reference's own (vendored) DBoW2 header

namespace ORB_SLAM_CUSTOM {

class Frame {
public:
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<std::shared_ptr<MapPoint>> mvpMapPoints;
    std::vector<float> mvLevelSigma2;
    DBoW2::BowVector mBowVec;
    long unsigned int mnId = 0;
};

}  // namespace ORB_SLAM_CUSTOM
