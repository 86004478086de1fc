// oracle/shim/orbslam/Frame.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/Frame.hpp with the members PnPsolver's / MLPnPsolver's constructors read
// (Frame.hpp:102-105 fx, fy, cx, cy as float; :127 mvKeysUn; :142 mvpMapPoints; mvLevelSigma2), the two
// DetectRelocalizationCandidates reads (:135 mBowVec, :157 mnId) and the ones ORBmatcher.cpp reads (descriptors,
// feature vector, pose, scale pyramid, image bounds, the feature grid).
// Frame::GetFeaturesInArea (src/Frame.cpp:393-446) forwards to the oracle's orc_features_in_area over the same grid
// and applies the level filter: it is one of the helpers the matcher calls, not the matcher.
#pragma once
#include <algorithm>
#include <limits>
#include "MapPoint.hpp"
#include "orc.h"
#include <opencv2/core/eigen.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"       // the reference's own (vendored) DBoW2 headers
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "Thirdparty/DBoW2/DUtils/Random.h"         // (reached through the real Frame.hpp include chain)

namespace ORB_SLAM_CUSTOM {

class Frame {
public:
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f, mbf = 0.f, mb = 0.f;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<std::shared_ptr<MapPoint>> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0.f;
    float mnMinX = 0.f, mnMinY = 0.f, mnMaxX = 0.f, mnMaxY = 0.f;
    Eigen::Isometry3f mTcw;
    DBoW2::BowVector mBowVec;
    long unsigned int mnId = 0;
    const orc_kf_view *mpView = nullptr;

    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const
    {
        std::vector<int32_t> tmp((size_t)(N > 0 ? N : 1));
        const int n = orc_features_in_area(mpView, x, y, r, tmp.data());
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        std::vector<size_t> out;
        for (int i = 0; i < n; ++i) {
            const int o = mvKeysUn[tmp[i]].octave;
            if (bCheckLevels && (o < minLevel || (maxLevel >= 0 && o > maxLevel))) continue;
            out.push_back((size_t)tmp[i]);
        }
        return out;
    }
};

inline int MapPoint::PredictScale(const float &currentDist, Frame *pF)
{
    return orc_predict_scale(mfMaxDistance, currentDist, pF->mfLogScaleFactor, pF->mnScaleLevels);
}

}  // namespace ORB_SLAM_CUSTOM
