// oracle/shim/orbslam/MapPoint.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/MapPoint.hpp with the members its PnPsolver / Sim3Solver / MLPnPsolver /
// ORBmatcher sources touch (MapPoint.hpp:31 GetWorldPos, :46 isBad, GetIndexInKeyFrame, GetDescriptor, the scale
// invariance distances, PredictScale); the real header pulls in the whole SLAM system (OpenCV, DBoW2, the map).
// The build recipe places it beside a link to the reference's headers so that their quoted includes resolve here.
// MapPoint::PredictScale (src/MapPoint.cpp:367-399) forwards to the oracle's orc_predict_scale: it is one of the
// helpers the matcher calls, not the matcher.
#pragma once
#include <memory>
#include <mutex>
#include <set>
#include <vector>
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>

extern "C" int orc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);

namespace ORB_SLAM_CUSTOM {
using namespace std;   // the reference's headers name vector<> unqualified

class KeyFrame;
class Frame;

class MapPoint {
public:
    Eigen::Vector3f mWorldPos, mNormalVector;
    bool mbBad = false;
    int mIndexKF1 = -1, mIndexKF2 = -1;          // index of this point in (up to) two keyframes
    const KeyFrame *mpKF1 = nullptr, *mpKF2 = nullptr;
    cv::Mat mDescriptor;
    float mfMinDistance = 0.f, mfMaxDistance = 0.f;
    int mnObs = 1;
    // Tracking's per-frame marks (ORBmatcher::SearchByProjection(Frame&, vpMapPoints), not on the path)
    bool mbTrackInView = false;
    float mTrackProjX = 0.f, mTrackProjY = 0.f, mTrackProjXR = 0.f, mTrackViewCos = 0.f;
    int mnTrackScaleLevel = 0;

    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    bool isBad() { return mbBad; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }      // MapPoint.cpp:355-365
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int Observations() { return mnObs; }
    void AddObservation(std::shared_ptr<KeyFrame>, size_t) { ++mnObs; }
    void Replace(std::shared_ptr<MapPoint>) {}
    int GetIndexInKeyFrame(std::shared_ptr<KeyFrame> pKF)
    {
        if (pKF.get() == mpKF1) return mIndexKF1;
        if (pKF.get() == mpKF2) return mIndexKF2;
        return -1;
    }
    bool IsInKeyFrame(std::shared_ptr<KeyFrame> pKF) { return GetIndexInKeyFrame(pKF) >= 0; }
    inline int PredictScale(const float &currentDist, std::shared_ptr<KeyFrame> pKF);
    inline int PredictScale(const float &currentDist, Frame *pF);
};

}  // namespace ORB_SLAM_CUSTOM
