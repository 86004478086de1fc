// oracle/shim/orbslam/ORBVocabulary.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/ORBVocabulary.hpp (a DBoW2::TemplatedVocabulary over OpenCV descriptors) with
// the two members KeyFrameDatabase.cpp calls: size() and score().  score() is what TemplatedVocabulary::score does --
// it forwards to the scoring object -- and the scoring object IS the reference's: DBoW2::L1Scoring from the vendored
// Thirdparty/DBoW2/DBoW2/ScoringObject.cpp (ORB-SLAM's vocabulary is TF-IDF / L1), compiled unmodified.
#pragma once
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/ScoringObject.h"

namespace ORB_SLAM_CUSTOM {

class ORBVocabulary {
public:
    explicit ORBVocabulary(unsigned int n_words) : m_words(n_words) {}
    unsigned int size() const { return m_words; }
    double score(const DBoW2::BowVector &a, const DBoW2::BowVector &b) const { return m_scoring.score(a, b); }

private:
    unsigned int m_words;
    DBoW2::L1Scoring m_scoring;
};

}  // namespace ORB_SLAM_CUSTOM
