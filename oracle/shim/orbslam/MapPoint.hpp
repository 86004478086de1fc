// oracle/shim/orbslam/MapPoint.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/MapPoint.hpp with the members its PnPsolver / Sim3Solver sources touch
// (MapPoint.hpp:31 GetWorldPos, :46 isBad, GetIndexInKeyFrame); the real header pulls in the whole SLAM system
// (OpenCV, DBoW2, the map).  The build recipe places it beside a link to the reference's PnPsolver.hpp so that
// the quoted include there resolves to this file.
#pragma once
#include <memory>
#include <mutex>
#include <vector>
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>

namespace ORB_SLAM_CUSTOM {
using namespace std;   // the reference's headers name vector<> unqualified

class KeyFrame;

class MapPoint {
public:
    Eigen::Vector3f mWorldPos;
    bool mbBad = false;
    int mIndexKF1 = -1, mIndexKF2 = -1;          // Sim3Solver: index of this point in the two keyframes
    const KeyFrame *mpKF1 = nullptr, *mpKF2 = nullptr;

    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    bool isBad() { return mbBad; }
    int GetIndexInKeyFrame(std::shared_ptr<KeyFrame> pKF)
    {
        if (pKF.get() == mpKF1) return mIndexKF1;
        if (pKF.get() == mpKF2) return mIndexKF2;
        return -1;
    }
};

}  // namespace ORB_SLAM_CUSTOM
