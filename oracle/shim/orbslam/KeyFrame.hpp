// oracle/shim/orbslam/KeyFrame.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/KeyFrame.hpp with the members Sim3Solver's constructor reads
// (GetMapPointMatches, GetRotation, GetTranslation, mvKeysUn, mvLevelSigma2, mK) and the ones KeyFrameDatabase.cpp
// reads and writes (KeyFrame.hpp:109 mnId, :129-134 the loop / relocalisation query marks, :155 mBowVec,
// :53-55 GetConnectedKeyFrames / GetBestCovisibilityKeyFrames), with the reference's own types.
#pragma once
#include "MapPoint.hpp"
#include <set>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"    // the reference's own (vendored) DBoW2 header
#include "Thirdparty/DBoW2/DUtils/Random.h"   // the reference's own header (reached through its real KeyFrame.hpp include chain)

namespace ORB_SLAM_CUSTOM {

class KeyFrame {
public:
    Eigen::Matrix3f mRcw, mK;
    Eigen::Vector3f mtcw;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvLevelSigma2;
    std::vector<std::shared_ptr<MapPoint>> mvpMapPoints;

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
    std::vector<std::shared_ptr<KeyFrame>> GetBestCovisibilityKeyFrames(const int &N)
    {
        std::vector<std::shared_ptr<KeyFrame>> v;
        for (size_t i = 0; i < mvpOrderedConnected.size() && (int)i < N; ++i) v.push_back(mvpOrderedConnected[i].lock());
        return v;
    }
    std::vector<std::shared_ptr<MapPoint>> GetMapPointMatches() { return mvpMapPoints; }
    Eigen::Matrix3f GetRotation() { return mRcw; }
    Eigen::Vector3f GetTranslation() { return mtcw; }
};

}  // namespace ORB_SLAM_CUSTOM
