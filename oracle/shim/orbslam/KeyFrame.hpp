// oracle/shim/orbslam/KeyFrame.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/KeyFrame.hpp with the members Sim3Solver's constructor reads
// (GetMapPointMatches, GetRotation, GetTranslation, mvKeysUn, mvLevelSigma2, mK).
#pragma once
#include "MapPoint.hpp"
#include "Thirdparty/DBoW2/DUtils/Random.h"   // the reference's own header (reached through its real KeyFrame.hpp include chain)

namespace ORB_SLAM_CUSTOM {

class KeyFrame {
public:
    Eigen::Matrix3f mRcw, mK;
    Eigen::Vector3f mtcw;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvLevelSigma2;
    std::vector<std::shared_ptr<MapPoint>> mvpMapPoints;

    std::vector<std::shared_ptr<MapPoint>> GetMapPointMatches() { return mvpMapPoints; }
    Eigen::Matrix3f GetRotation() { return mRcw; }
    Eigen::Vector3f GetTranslation() { return mtcw; }
};

}  // namespace ORB_SLAM_CUSTOM
