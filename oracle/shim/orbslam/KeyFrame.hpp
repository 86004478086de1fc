// oracle/shim/orbslam/KeyFrame.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/KeyFrame.hpp with the members Sim3Solver's constructor reads
// (GetMapPointMatches, GetRotation, GetTranslation, mvKeysUn, mvLevelSigma2, mK), the ones KeyFrameDatabase.cpp
// reads and writes (KeyFrame.hpp:109 mnId, :129-134 the loop / relocalisation query marks, :155 mBowVec,
// :53-55 GetConnectedKeyFrames / GetBestCovisibilityKeyFrames) and the ones ORBmatcher.cpp reads (descriptors,
// feature vector, scale pyramid, image bounds, the feature grid), with the reference's own types.
// KeyFrame::GetFeaturesInArea (src/KeyFrame.cpp:560-599) forwards to the oracle's orc_features_in_area over the same
// grid: it is one of the helpers the matcher calls, not the matcher.
#pragma once
#include <set>
#include "MapPoint.hpp"
#include "orc.h"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"       // the reference's own (vendored) DBoW2 headers
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "Thirdparty/DBoW2/DUtils/Random.h"         // (reached through the real KeyFrame.hpp include chain)

namespace ORB_SLAM_CUSTOM {

class KeyFrame {
public:
    Eigen::Matrix3f mRcw, mK;
    Eigen::Vector3f mtcw;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0.f;
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f, mbf = 0.f, mb = 0.f;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    std::vector<std::shared_ptr<MapPoint>> mvpMapPoints;
    const orc_kf_view *mpView = nullptr;           // the feature grid lives in the oracle's view of this keyframe

    long unsigned int mnId = 0;
    long unsigned int mnLoopQuery = 0;
    int mnLoopWords = 0;
    float mLoopScore = 0.f;
    long unsigned int mnRelocQuery = 0;
    int mnRelocWords = 0;
    float mRelocScore = 0.f;
    DBoW2::BowVector mBowVec;
    std::vector<std::weak_ptr<KeyFrame>> mvpOrderedConnected;      // best covisibility first
    std::vector<std::weak_ptr<KeyFrame>> mvpConnected;

    std::set<std::shared_ptr<KeyFrame>> GetConnectedKeyFrames()
    {
        std::set<std::shared_ptr<KeyFrame>> s;
        for (auto &w : mvpConnected) s.insert(w.lock());
        return s;
    }
    std::vector<std::shared_ptr<KeyFrame>> GetBestCovisibilityKeyFrames(const int &N_)
    {
        std::vector<std::shared_ptr<KeyFrame>> v;
        for (size_t i = 0; i < mvpOrderedConnected.size() && (int)i < N_; ++i) v.push_back(mvpOrderedConnected[i].lock());
        return v;
    }
    std::vector<std::shared_ptr<MapPoint>> GetMapPointMatches() { return mvpMapPoints; }
    std::shared_ptr<MapPoint> GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
    std::set<std::shared_ptr<MapPoint>> GetMapPoints()
    {
        std::set<std::shared_ptr<MapPoint>> s;
        for (auto &p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
        return s;
    }
    void AddMapPoint(std::shared_ptr<MapPoint> pMP, const size_t &idx) { mvpMapPoints[idx] = pMP; }
    Eigen::Matrix3f GetRotation() { return mRcw; }
    Eigen::Vector3f GetTranslation() { return mtcw; }
    Eigen::Vector3f GetCameraCenter() { return -(mRcw.transpose() * mtcw); }
    bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const
    {
        std::vector<int32_t> tmp((size_t)(N > 0 ? N : 1));
        const int n = orc_features_in_area(mpView, x, y, r, tmp.data());
        return std::vector<size_t>(tmp.begin(), tmp.begin() + n);
    }
};

inline int MapPoint::PredictScale(const float &currentDist, std::shared_ptr<KeyFrame> pKF)
{
    return orc_predict_scale(mfMaxDistance, currentDist, pKF->mfLogScaleFactor, pKF->mnScaleLevels);
}

}  // namespace ORB_SLAM_CUSTOM
