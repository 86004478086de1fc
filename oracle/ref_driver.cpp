// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the REFERENCE's own classes, compiled from the reference's sources where they lie:
//   /root/reference/src/PnPsolver.cpp, Sim3Solver.cpp, MLPnPsolver.cpp,
//   KeyFrameDatabase.cpp                                                       (unmodified, against oracle/shim/)
//   /root/reference/Thirdparty/DBoW2/DUtils/Random.cpp, Timestamp.cpp          (unmodified, no stand-ins needed)
// Built by `make -C oracle ref` into oracle/_ref/libref_solvers.so; loaded only by tests/test_cpu_reference_build.py
// through tests/ref_api.py.  The stand-in headers (oracle/shim/) and what a build against them does and does not pin
// are described in oracle/shim/Eigen/Dense.
//
// Private / protected members are reached by redefining the access keywords for this translation unit only (the
// solver sources themselves are compiled as they are): the tests call compute_pose / CheckInliers / ComputeSim3 on
// chosen inputs and read the RANSAC state.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <vector>
#include <Eigen/Dense>

#define private public
#define protected public
#include "PnPsolver.hpp"
#include "Sim3Solver.hpp"
#include "MLPnPsolver.hpp"
#undef private
#undef protected

using namespace ORB_SLAM_CUSTOM;

namespace {

struct PnpBox {
    Frame frame;
    std::vector<std::shared_ptr<MapPoint>> matches;
    std::unique_ptr<PnPsolver> solver;
};

struct Sim3Box {
    std::shared_ptr<KeyFrame> kf1, kf2;
    std::vector<std::shared_ptr<MapPoint>> matched12;
    std::unique_ptr<Sim3Solver> solver;
};

// recorded 12 x 12 eigen-problems of the current thread's solver calls (shim hook)
std::vector<double> g_eig_basis;   // 48 doubles per call
int g_eig_calls = 0;
bool g_eig_record = false;
void eig_hook(const double *, double *U, double *)
{
    ++g_eig_calls;
    if (g_eig_record) g_eig_basis.insert(g_eig_basis.end(), U, U + 48);
}

void store_T(const Eigen::Matrix4f &T, float *out)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = T(i, j);
}

}  // namespace

extern "C" {

// ---- DUtils::Random (the reference's own Random.cpp) ----
void ref_seed(int seed) { DUtils::Random::SeedRand(seed); }
int ref_random_int(int lo, int hi) { return DUtils::Random::RandomInt(lo, hi); }

// ---- 12 x 12 eigen-problem recorder ----
void ref_eig_record(int on)
{
    Eigen::shim_eig12_hook() = eig_hook;
    g_eig_record = on != 0;
    g_eig_basis.clear();
    g_eig_calls = 0;
}
int ref_eig_calls(void) { return g_eig_calls; }
int ref_eig_take(double *out, int max_calls)
{
    const int n = std::min<int>(max_calls, (int)(g_eig_basis.size() / 48));
    if (out && n > 0) std::memcpy(out, g_eig_basis.data(), sizeof(double) * 48 * (size_t)n);
    return n;
}

// ---- PnPsolver ----
// state[i]: 0 = keypoint i has no map point, 1 = map point, 2 = bad map point (PnPsolver.cpp:23-29)
void *ref_pnp_create(int n_kp, const float *kp_xy, const int *octave, const float *level_sigma2, int n_levels,
                     const float *mp_xyz, const uint8_t *state, float fx, float fy, float cx, float cy)
{
    PnpBox *b = new PnpBox;
    b->frame.fx = fx; b->frame.fy = fy; b->frame.cx = cx; b->frame.cy = cy;
    b->frame.mvLevelSigma2.assign(level_sigma2, level_sigma2 + n_levels);
    b->frame.mvKeysUn.resize(n_kp);
    b->frame.mvpMapPoints.resize(n_kp);
    b->matches.resize(n_kp);
    for (int i = 0; i < n_kp; ++i) {
        b->frame.mvKeysUn[i].pt = cv::Point2f(kp_xy[2 * i], kp_xy[2 * i + 1]);
        b->frame.mvKeysUn[i].octave = octave[i];
        if (state[i]) {
            auto mp = std::make_shared<MapPoint>();
            mp->mWorldPos = Eigen::Vector3f(mp_xyz[3 * i], mp_xyz[3 * i + 1], mp_xyz[3 * i + 2]);
            mp->mbBad = state[i] == 2;
            b->matches[i] = mp;
        }
    }
    b->solver.reset(new PnPsolver(b->frame, b->matches));
    return b;
}
void ref_pnp_destroy(void *h) { delete static_cast<PnpBox *>(h); }

void ref_pnp_set_params(void *h, double prob, int min_inliers, int max_its, int min_set, float eps, float th2)
{
    static_cast<PnpBox *>(h)->solver->SetRansacParameters(prob, min_inliers, max_its, min_set, eps, th2);
}
// N, adjusted mRansacMinInliers / mRansacMaxIts / mRansacEpsilon, mvMaxError[N], mvKeyPointIndices[N]
void ref_pnp_get_params(void *h, int *N, int *min_inliers, int *max_its, float *eps, float *max_err, int *kp_index)
{
    PnPsolver &s = *static_cast<PnpBox *>(h)->solver;
    *N = (int)s.mvP2D.size();
    *min_inliers = s.mRansacMinInliers;
    *max_its = s.mRansacMaxIts;
    *eps = s.mRansacEpsilon;
    if (max_err)
        for (size_t i = 0; i < s.mvMaxError.size(); ++i) max_err[i] = s.mvMaxError[i];
    if (kp_index)
        for (size_t i = 0; i < s.mvKeyPointIndices.size(); ++i) kp_index[i] = (int)s.mvKeyPointIndices[i];
}
int ref_pnp_iterate(void *h, int n_iterations, int *no_more, uint8_t *inliers /* n_kp */, int *n_inliers, float *T16)
{
    PnpBox *b = static_cast<PnpBox *>(h);
    bool bNoMore = false;
    std::vector<bool> vb;
    int nInl = 0;
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const bool ok = b->solver->iterate(n_iterations, bNoMore, vb, nInl, T);
    *no_more = bNoMore ? 1 : 0;
    *n_inliers = nInl;
    std::memset(inliers, 0, b->matches.size());
    for (size_t i = 0; i < vb.size() && i < b->matches.size(); ++i) inliers[i] = vb[i] ? 1 : 0;
    store_T(T, T16);
    return ok ? 1 : 0;
}
// RANSAC state after a call: mnIterations, mnBestInliers, mnRefinedInliers, mBestTcw, mvbBestInliers[N]
void ref_pnp_state(void *h, int *iterations, int *best_inliers, int *refined_inliers, float *best_T16, uint8_t *best_mask)
{
    PnPsolver &s = *static_cast<PnpBox *>(h)->solver;
    *iterations = s.mnIterations;
    *best_inliers = s.mnBestInliers;
    *refined_inliers = s.mnRefinedInliers;
    if (best_T16) store_T(s.mBestTcw, best_T16);
    if (best_mask)
        for (size_t i = 0; i < s.mvbBestInliers.size(); ++i) best_mask[i] = s.mvbBestInliers[i] ? 1 : 0;
}
// PnPsolver::compute_pose on the subset idx[0..m) of the solver's correspondences, driven exactly as iterate() /
// Refine() drive it (PnPsolver.cpp:109,122,131 / :208-217)
double ref_pnp_compute_pose(void *h, const int *idx, int m, float *R9, float *t3)
{
    PnPsolver &s = *static_cast<PnpBox *>(h)->solver;
    s.set_maximum_number_of_correspondences(m);
    s.reset_correspondences();
    for (int i = 0; i < m; ++i) s.add_correspondence(s.mvP3Dw[idx[i]], s.mvP2D[idx[i]]);
    const double err = s.compute_pose(s.mRi, s.mti);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R9[i * 3 + j] = s.mRi(i, j);
        t3[i] = s.mti(i);
    }
    return err;
}
// PnPsolver::CheckInliers for a given pose; returns mnInliersi, mask = mvbInliersi
int ref_pnp_check_inliers(void *h, const float *R9, const float *t3, uint8_t *mask)
{
    PnPsolver &s = *static_cast<PnpBox *>(h)->solver;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) s.mRi(i, j) = R9[i * 3 + j];
        s.mti(i) = t3[i];
    }
    s.CheckInliers();
    for (size_t i = 0; i < s.mvbInliersi.size(); ++i) mask[i] = s.mvbInliersi[i] ? 1 : 0;
    return s.mnInliersi;
}

// ---- Sim3Solver ----
// Both keyframes sit at the world origin with identity rotation, so the constructor's camera-frame points
// (Sim3Solver.cpp:57-63) are the given ones exactly.  state[i] as above, for the matched point of keypoint i of KF1;
// keypoint i of KF1 always holds map point i.
void *ref_sim3_create(int n, const float *x1, const float *x2, const int *octave1, const int *octave2,
                      const float *level_sigma2, int n_levels, const uint8_t *state, const float *K1 /* fx fy cx cy */,
                      const float *K2)
{
    Sim3Box *b = new Sim3Box;
    b->kf1 = std::make_shared<KeyFrame>();
    b->kf2 = std::make_shared<KeyFrame>();
    KeyFrame *kfs[2] = {b->kf1.get(), b->kf2.get()};
    const float *Ks[2] = {K1, K2};
    for (int k = 0; k < 2; ++k) {
        kfs[k]->mRcw.setIdentity();
        kfs[k]->mtcw.setZero();
        kfs[k]->mK.setIdentity();
        kfs[k]->mK(0, 0) = Ks[k][0]; kfs[k]->mK(1, 1) = Ks[k][1]; kfs[k]->mK(0, 2) = Ks[k][2]; kfs[k]->mK(1, 2) = Ks[k][3];
        kfs[k]->mvLevelSigma2.assign(level_sigma2, level_sigma2 + n_levels);
        kfs[k]->mvKeysUn.resize(n);
        kfs[k]->mvpMapPoints.resize(n);
    }
    b->matched12.resize(n);
    for (int i = 0; i < n; ++i) {
        b->kf1->mvKeysUn[i].octave = octave1[i];
        b->kf2->mvKeysUn[i].octave = octave2[i];
        auto mp1 = std::make_shared<MapPoint>();
        mp1->mWorldPos = Eigen::Vector3f(x1[3 * i], x1[3 * i + 1], x1[3 * i + 2]);
        mp1->mpKF1 = b->kf1.get(); mp1->mIndexKF1 = i;
        b->kf1->mvpMapPoints[i] = mp1;
        if (state[i]) {
            auto mp2 = std::make_shared<MapPoint>();
            mp2->mWorldPos = Eigen::Vector3f(x2[3 * i], x2[3 * i + 1], x2[3 * i + 2]);
            mp2->mpKF2 = b->kf2.get(); mp2->mIndexKF2 = i;
            mp2->mbBad = state[i] == 2;
            b->kf2->mvpMapPoints[i] = mp2;
            b->matched12[i] = mp2;
        }
    }
    b->solver.reset(new Sim3Solver(b->kf1, b->kf2, b->matched12));
    return b;
}
void ref_sim3_destroy(void *h) { delete static_cast<Sim3Box *>(h); }
void ref_sim3_set_params(void *h, double prob, int min_inliers, int max_its)
{
    static_cast<Sim3Box *>(h)->solver->SetRansacParameters(prob, min_inliers, max_its);
}
// N, mRansacMaxIts, the integer thresholds mvnMaxError1/2 (Sim3Solver.hpp: vector<size_t>)
void ref_sim3_get_params(void *h, int *N, int *max_its, uint64_t *max_err1, uint64_t *max_err2)
{
    Sim3Solver &s = *static_cast<Sim3Box *>(h)->solver;
    *N = s.N;
    *max_its = s.mRansacMaxIts;
    for (size_t i = 0; i < s.mvnMaxError1.size(); ++i) {
        if (max_err1) max_err1[i] = s.mvnMaxError1[i];
        if (max_err2) max_err2[i] = s.mvnMaxError2[i];
    }
}
int ref_sim3_iterate(void *h, int n_iterations, int *no_more, uint8_t *inliers /* n */, int *n_inliers)
{
    Sim3Box *b = static_cast<Sim3Box *>(h);
    bool bNoMore = false;
    std::vector<bool> vb;
    int nInl = 0;
    const bool ok = b->solver->iterate(n_iterations, bNoMore, vb, nInl);
    *no_more = bNoMore ? 1 : 0;
    *n_inliers = nInl;
    std::memset(inliers, 0, b->matched12.size());
    for (size_t i = 0; i < vb.size() && i < b->matched12.size(); ++i) inliers[i] = vb[i] ? 1 : 0;
    return ok ? 1 : 0;
}
// mnIterations, mnBestInliers, GetEstimatedRotation / GetEstimatedTranslation, mvbBestInliers[N]
void ref_sim3_state(void *h, int *iterations, int *best_inliers, float *R9, float *t3, uint8_t *best_mask)
{
    Sim3Solver &s = *static_cast<Sim3Box *>(h)->solver;
    *iterations = s.mnIterations;
    *best_inliers = s.mnBestInliers;
    const Eigen::Matrix3f R = s.GetEstimatedRotation();
    const Eigen::Vector3f t = s.GetEstimatedTranslation();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R(i, j);
        t3[i] = t(i);
    }
    if (best_mask)
        for (size_t i = 0; i < s.mvbBestInliers.size(); ++i) best_mask[i] = s.mvbBestInliers[i] ? 1 : 0;
}
// Sim3Solver::ComputeSim3 on three pairs (columns of P are points, Sim3Solver.cpp:142-143) followed by CheckInliers
int ref_sim3_compute_and_check(void *h, const int *idx3, float *R9, float *t3, uint8_t *mask)
{
    Sim3Solver &s = *static_cast<Sim3Box *>(h)->solver;
    Eigen::Matrix3f P1, P2;
    for (int i = 0; i < 3; ++i) {
        P1.col(i) = s.mvX3Dc1[idx3[i]];
        P2.col(i) = s.mvX3Dc2[idx3[i]];
    }
    s.ComputeSim3(P1, P2);
    s.CheckInliers();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R9[i * 3 + j] = s.mR12i(i, j);
        t3[i] = s.mt12i(i);
    }
    for (size_t i = 0; i < s.mvbInliersi.size(); ++i) mask[i] = s.mvbInliersi[i] ? 1 : 0;
    return s.mnInliersi;
}

}  // extern "C"

// ---- KeyFrameDatabase ----
#include "KeyFrameDatabase.hpp"

namespace {
struct KfdbBox {
    std::shared_ptr<ORBVocabulary> voc;
    std::vector<std::shared_ptr<KeyFrame>> kfs;
    std::unique_ptr<KeyFrameDatabase> db;
};
int emit(const std::vector<std::shared_ptr<KeyFrame>> &cands, const KfdbBox *b, int32_t *out, int cap)
{
    int n = 0;
    for (const auto &kf : cands) {
        int idx = -1;
        for (size_t k = 0; k < b->kfs.size(); ++k)
            if (b->kfs[k] == kf) { idx = (int)k; break; }
        if (n < cap) out[n] = idx;
        ++n;
    }
    return n;
}
}  // namespace

extern "C" {

// keyframe k: BowVector = (bow_word, bow_val)[bow_off[k] .. bow_off[k+1]), GetBestCovisibilityKeyFrames(10) = covis[k][0..10)
// (-1 padded); keyframes are add()-ed in index order (KeyFrameDatabase.cpp:15-22)
void *ref_kfdb_create(int K, const int64_t *bow_off, const uint32_t *bow_word, const double *bow_val, const int32_t *covis, unsigned n_words)
{
    KfdbBox *b = new KfdbBox;
    b->voc = std::make_shared<ORBVocabulary>(n_words);
    b->kfs.resize(K);
    for (int k = 0; k < K; ++k) {
        b->kfs[k] = std::make_shared<KeyFrame>();
        b->kfs[k]->mnId = (unsigned long)k + 1000000000ul;
        for (int64_t i = bow_off[k]; i < bow_off[k + 1]; ++i) b->kfs[k]->mBowVec.addWeight(bow_word[i], bow_val[i]);
    }
    for (int k = 0; k < K; ++k)
        for (int j = 0; j < 10 && covis[k * 10 + j] >= 0; ++j) b->kfs[k]->mvpOrderedConnected.push_back(b->kfs[covis[k * 10 + j]]);
    b->db.reset(new KeyFrameDatabase(b->voc));
    for (int k = 0; k < K; ++k) b->db->add(b->kfs[k]);
    return b;
}
void ref_kfdb_destroy(void *h) { delete static_cast<KfdbBox *>(h); }

// KeyFrameDatabase::DetectRelocalizationCandidates(Frame*): candidate keyframe indices in the returned order
int ref_kfdb_reloc(void *h, int nq, const uint32_t *qword, const double *qval, unsigned long frame_id, int32_t *out, int cap)
{
    KfdbBox *b = static_cast<KfdbBox *>(h);
    Frame F;
    F.mnId = frame_id;
    for (int i = 0; i < nq; ++i) F.mBowVec.addWeight(qword[i], qval[i]);
    return emit(b->db->DetectRelocalizationCandidates(&F), b, out, cap);
}
// KeyFrameDatabase::DetectLoopCandidates(pKF, minScore) for database keyframe q, which takes the id query_id for the
// call; conn = GetConnectedKeyFrames()
int ref_kfdb_loop(void *h, int q, unsigned long query_id, int n_conn, const int32_t *conn, float min_score, int32_t *out, int cap)
{
    KfdbBox *b = static_cast<KfdbBox *>(h);
    std::shared_ptr<KeyFrame> kf = b->kfs[q];
    kf->mnId = query_id;
    kf->mvpConnected.clear();
    for (int i = 0; i < n_conn; ++i) kf->mvpConnected.push_back(b->kfs[conn[i]]);
    return emit(b->db->DetectLoopCandidates(kf, min_score), b, out, cap);
}
void ref_kfdb_reloc_scores(void *h, float *scores)
{
    KfdbBox *b = static_cast<KfdbBox *>(h);
    for (size_t k = 0; k < b->kfs.size(); ++k) scores[k] = b->kfs[k]->mRelocScore;
}
double ref_bow_l1_score(int n1, const uint32_t *w1, const double *v1, int n2, const uint32_t *w2, const double *v2)
{
    DBoW2::BowVector a, c;
    for (int i = 0; i < n1; ++i) a.addWeight(w1[i], v1[i]);
    for (int i = 0; i < n2; ++i) c.addWeight(w2[i], v2[i]);
    return DBoW2::L1Scoring().score(a, c);
}

}  // extern "C"

// ---- MLPnPsolver (commented out of the reference's own CMakeLists.txt:75; compiled here all the same) ----
namespace {
struct MlpnpBox {
    Frame frame;
    std::vector<std::shared_ptr<MapPoint>> matches;
    std::unique_ptr<MLPnPsolver> solver;
};
}  // namespace

extern "C" {

void *ref_mlpnp_create(int n_kp, const float *kp_xy, const int *octave, const float *level_sigma2, int n_levels,
                       const float *mp_xyz, const uint8_t *state, float fx, float fy, float cx, float cy)
{
    MlpnpBox *b = new MlpnpBox;
    b->frame.fx = fx; b->frame.fy = fy; b->frame.cx = cx; b->frame.cy = cy;
    b->frame.mvLevelSigma2.assign(level_sigma2, level_sigma2 + n_levels);
    b->frame.mvKeysUn.resize(n_kp);
    b->frame.mvpMapPoints.resize(n_kp);
    b->matches.resize(n_kp);
    for (int i = 0; i < n_kp; ++i) {
        b->frame.mvKeysUn[i].pt = cv::Point2f(kp_xy[2 * i], kp_xy[2 * i + 1]);
        b->frame.mvKeysUn[i].octave = octave[i];
        if (state[i]) {
            auto mp = std::make_shared<MapPoint>();
            mp->mWorldPos = Eigen::Vector3f(mp_xyz[3 * i], mp_xyz[3 * i + 1], mp_xyz[3 * i + 2]);
            mp->mbBad = state[i] == 2;
            b->matches[i] = mp;
        }
    }
    b->solver.reset(new MLPnPsolver(b->frame, b->matches));
    return b;
}
void ref_mlpnp_destroy(void *h) { delete static_cast<MlpnpBox *>(h); }
void ref_mlpnp_set_params(void *h, double prob, int min_inliers, int max_its, int min_set, float eps, float th2)
{
    static_cast<MlpnpBox *>(h)->solver->SetRansacParameters(prob, min_inliers, max_its, min_set, eps, th2);
}
void ref_mlpnp_get_params(void *h, int *N, int *min_inliers, int *max_its, float *max_err, int *kp_index)
{
    MLPnPsolver &s = *static_cast<MlpnpBox *>(h)->solver;
    *N = s.N;
    *min_inliers = s.mRansacMinInliers;
    *max_its = s.mRansacMaxIts;
    if (max_err)
        for (size_t i = 0; i < s.mvMaxError.size(); ++i) max_err[i] = s.mvMaxError[i];
    if (kp_index)
        for (size_t i = 0; i < s.mvKeyPointIndices.size(); ++i) kp_index[i] = (int)s.mvKeyPointIndices[i];
}
int ref_mlpnp_iterate(void *h, int n_iterations, int *no_more, uint8_t *inliers, int *n_inliers, float *T16)
{
    MlpnpBox *b = static_cast<MlpnpBox *>(h);
    bool bNoMore = false;
    std::vector<bool> vb;
    int nInl = 0;
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const bool ok = b->solver->iterate(n_iterations, bNoMore, vb, nInl, T);
    *no_more = bNoMore ? 1 : 0;
    *n_inliers = nInl;
    std::memset(inliers, 0, b->matches.size());
    for (size_t i = 0; i < vb.size() && i < b->matches.size(); ++i) inliers[i] = vb[i] ? 1 : 0;
    store_T(T, T16);
    return ok ? 1 : 0;
}
void ref_mlpnp_state(void *h, int *iterations, int *best_inliers, int *refined_inliers)
{
    MLPnPsolver &s = *static_cast<MlpnpBox *>(h)->solver;
    *iterations = s.mnIterations;
    *best_inliers = s.mnBestInliers;
    *refined_inliers = s.mnRefinedInliers;
}
// MLPnPsolver::computePose on the subset idx[0..m), driven as iterate() / Refine() drive it (MLPnPsolver.cpp:80-105);
// cov != NULL: m bearing covariances (3 x 3 each) -- the use_cov branch (:375-388) iterate() itself never takes
void ref_mlpnp_compute_pose(void *h, const int *idx, int m, const double *cov, double *R9, double *t3)
{
    MLPnPsolver &s = *static_cast<MlpnpBox *>(h)->solver;
    MLPnPsolver::bearingVectors_t f(m);
    MLPnPsolver::points_t p(m);
    std::vector<int> indexes(m);
    for (int i = 0; i < m; ++i) { f[i] = s.mvBearingVecs[idx[i]]; p[i] = s.mvP3Dw[idx[i]]; indexes[i] = i; }
    MLPnPsolver::cov3_mats_t covs(cov ? m : 1);
    if (cov)
        for (int i = 0; i < m; ++i)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) covs[i](r, c) = cov[9 * i + 3 * r + c];
    MLPnPsolver::transformation_t result;
    s.computePose(f, p, covs, indexes, result);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R9[r * 3 + c] = result(r, c);
        t3[r] = result(r, 3);
    }
}
int ref_mlpnp_check_inliers(void *h, const double *R9, const double *t3, uint8_t *mask)
{
    MLPnPsolver &s = *static_cast<MlpnpBox *>(h)->solver;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) s.mRi[r][c] = R9[r * 3 + c];
        s.mti[r] = t3[r];
    }
    s.CheckInliers();
    for (size_t i = 0; i < s.mvbInliersi.size(); ++i) mask[i] = s.mvbInliersi[i] ? 1 : 0;
    return s.mnInliersi;
}
// MLPnPsolver::mlpnpJacs (:773-1020, the generated polynomial) and the residual pair of :736-742
void ref_mlpnp_res_jac(void *h, const double *pt, const double *nr, const double *ns, const double *w, const double *t, double *r2, double *J12)
{
    MLPnPsolver &s = *static_cast<MlpnpBox *>(h)->solver;
    const Eigen::Vector3d P(pt[0], pt[1], pt[2]), NR(nr[0], nr[1], nr[2]), NS(ns[0], ns[1], ns[2]), W(w[0], w[1], w[2]), T(t[0], t[1], t[2]);
    Eigen::MatrixXd jacs(2, 6);
    s.mlpnpJacs(P, NR, NS, W, T, jacs);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 6; ++j) J12[i * 6 + j] = jacs(i, j);
    Eigen::Vector3d ptCam = s.rodrigues2rot(W) * P + T;
    ptCam /= ptCam.norm();
    r2[0] = NR.transpose() * ptCam;
    r2[1] = NS.transpose() * ptCam;
}
void ref_mlpnp_rodrigues(void *h, const double *w, double *R9, double *w_back)
{
    MLPnPsolver &s = *static_cast<MlpnpBox *>(h)->solver;
    const Eigen::Matrix3d R = s.rodrigues2rot(Eigen::Vector3d(w[0], w[1], w[2]));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R(i, j);
    const Eigen::Vector3d wb = s.rot2rodrigues(R);
    for (int i = 0; i < 3; ++i) w_back[i] = wb(i);
}

}  // extern "C"

// ---- ORBmatcher: SearchByBoW (both overloads), SearchBySim3, SearchByProjection(Frame, KeyFrame) ----
#include "ORBmatcher.hpp"

namespace {

cv::Mat make_descriptors(const uint32_t *desc, int n)
{
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(m.data, desc, (size_t)n * 32);
    return m;
}
std::shared_ptr<MapPoint> make_point(const KeyFrame *owner, int idx, const uint32_t *desc8)
{
    auto mp = std::make_shared<MapPoint>();
    mp->mpKF1 = owner;
    mp->mIndexKF1 = idx;
    mp->mDescriptor = make_descriptors(desc8, 1);
    mp->mWorldPos.setZero();
    mp->mNormalVector.setZero();
    return mp;
}
void fill_feature_vector(DBoW2::FeatureVector &fv, const orc_bow_features *f)
{
    for (int k = 0; k < f->n_nodes; ++k)
        for (int j = f->node_off[k]; j < f->node_off[k + 1]; ++j) fv.addFeature(f->node_ids[k], f->node_feat[j]);
}
// a keyframe as SearchByBoW reads it: descriptors, angles, map points (where valid), feature vector
std::shared_ptr<KeyFrame> kf_from_bow(const orc_bow_features *f)
{
    auto kf = std::make_shared<KeyFrame>();
    kf->N = f->n_feat;
    kf->mvKeysUn.resize(f->n_feat);
    kf->mvKeys.resize(f->n_feat);
    kf->mvpMapPoints.resize(f->n_feat);
    kf->mDescriptors = make_descriptors(f->desc, f->n_feat);
    for (int i = 0; i < f->n_feat; ++i) {
        kf->mvKeysUn[i].angle = kf->mvKeys[i].angle = f->angle[i];
        if (!f->valid || f->valid[i]) kf->mvpMapPoints[i] = make_point(kf.get(), i, f->desc + 8 * (size_t)i);
    }
    fill_feature_vector(kf->mFeatVec, f);
    return kf;
}
// a keyframe / frame as the guided searches read it (orc_kf_view)
void fill_view_common(const orc_kf_view *v, const float *K, std::vector<cv::KeyPoint> &keys, cv::Mat &desc, std::vector<float> &uright,
                      std::vector<float> &scale, std::vector<float> &sig2, std::vector<float> &isig2)
{
    keys.resize(v->n_feat);
    for (int i = 0; i < v->n_feat; ++i) {
        keys[i].pt = cv::Point2f(v->kp_xy[2 * i], v->kp_xy[2 * i + 1]);
        keys[i].octave = v->kp_octave[i];
        keys[i].angle = v->kp_angle ? v->kp_angle[i] : 0.f;
    }
    desc = make_descriptors(v->desc, v->n_feat);
    uright.assign(v->n_feat, -1.f);                         // monocular: no right coordinate anywhere
    scale.assign(v->scale_factors, v->scale_factors + v->n_levels);
    sig2.resize(v->n_levels);
    isig2.resize(v->n_levels);
    for (int l = 0; l < v->n_levels; ++l) { sig2[l] = scale[l] * scale[l]; isig2[l] = 1.0f / sig2[l]; }
    (void)K;
}
std::shared_ptr<KeyFrame> kf_from_view(const orc_kf_view *v, const float *K)
{
    auto kf = std::make_shared<KeyFrame>();
    kf->N = v->n_feat;
    kf->mpView = v;
    fill_view_common(v, K, kf->mvKeysUn, kf->mDescriptors, kf->mvuRight, kf->mvScaleFactors, kf->mvLevelSigma2, kf->mvInvLevelSigma2);
    kf->mvKeys = kf->mvKeysUn;
    kf->mnScaleLevels = v->n_levels;
    kf->mfLogScaleFactor = v->log_scale_factor;
    kf->fx = K[0]; kf->fy = K[1]; kf->cx = K[2]; kf->cy = K[3];
    kf->mnMinX = (int)v->bounds[0]; kf->mnMaxX = (int)v->bounds[1]; kf->mnMinY = (int)v->bounds[2]; kf->mnMaxY = (int)v->bounds[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) kf->mRcw(i, j) = v->Rcw[i * 3 + j];
        kf->mtcw(i) = v->tcw[i];
    }
    kf->mvpMapPoints.resize(v->n_feat);
    for (int i = 0; i < v->n_feat; ++i) {
        if (!v->mp_valid[i]) continue;
        auto mp = make_point(kf.get(), i, v->mp_desc + 8 * (size_t)i);
        mp->mWorldPos = Eigen::Vector3f(v->mp_xyz[3 * i], v->mp_xyz[3 * i + 1], v->mp_xyz[3 * i + 2]);
        mp->mfMaxDistance = v->mp_maxdist[i];
        mp->mfMinDistance = v->mp_mindist[i];
        kf->mvpMapPoints[i] = mp;
    }
    return kf;
}

}  // namespace

extern "C" {

// ORBmatcher::SearchByBoW: mode 0 (KeyFrame q, Frame t): match_out[frame feature] = keyframe feature or -1;
// mode 1 (KeyFrame q, KeyFrame t): match_out[q feature] = t feature or -1.  Returns nmatches.
int ref_search_by_bow(const orc_bow_features *q, const orc_bow_features *t, float nn_ratio, int check_orientation, int mode, int32_t *match_out)
{
    ORBmatcher matcher(nn_ratio, check_orientation != 0);
    std::shared_ptr<KeyFrame> kq = kf_from_bow(q);
    std::vector<std::shared_ptr<MapPoint>> matches;
    int n = 0;
    if (mode == 0) {
        Frame F;
        F.N = t->n_feat;
        F.mvKeys.resize(t->n_feat);
        F.mvKeysUn.resize(t->n_feat);
        for (int i = 0; i < t->n_feat; ++i) F.mvKeys[i].angle = F.mvKeysUn[i].angle = t->angle[i];
        F.mDescriptors = make_descriptors(t->desc, t->n_feat);
        F.mvpMapPoints.resize(t->n_feat);
        fill_feature_vector(F.mFeatVec, t);
        n = matcher.SearchByBoW(kq, F, matches);
        for (int i = 0; i < t->n_feat; ++i) match_out[i] = (i < (int)matches.size() && matches[i]) ? matches[i]->mIndexKF1 : -1;
    } else {
        std::shared_ptr<KeyFrame> kt = kf_from_bow(t);
        n = matcher.SearchByBoW(kq, kt, matches);
        for (int i = 0; i < q->n_feat; ++i) match_out[i] = (i < (int)matches.size() && matches[i]) ? matches[i]->mIndexKF1 : -1;
    }
    return n;
}
int ref_descriptor_distance(const uint32_t *a, const uint32_t *b)
{
    return ORBmatcher::DescriptorDistance(make_descriptors(a, 1), make_descriptors(b, 1));
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, R12, t12, th) (the reference has no scale argument: s12 = 1)
int ref_search_by_sim3(const orc_kf_view *v1, const orc_kf_view *v2, const float *K, const float *R12, const float *t12, float th,
                       const int32_t *matched12_in, int32_t *match12_out)
{
    ORBmatcher matcher(0.75f, true);
    std::shared_ptr<KeyFrame> k1 = kf_from_view(v1, K), k2 = kf_from_view(v2, K);
    std::vector<std::shared_ptr<MapPoint>> m12(v1->n_feat);
    for (int i = 0; i < v1->n_feat; ++i)
        if (matched12_in && matched12_in[i] >= 0) m12[i] = k2->mvpMapPoints[matched12_in[i]];
    Eigen::Matrix3f R;
    Eigen::Vector3f t;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R(i, j) = R12[i * 3 + j];
        t(i) = t12[i];
    }
    const int n = matcher.SearchBySim3(k1, k2, m12, R, t, th);
    for (int i = 0; i < v1->n_feat; ++i) match12_out[i] = m12[i] ? m12[i]->mIndexKF1 : -1;
    return n;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, pKF, sAlreadyFound, th, ORBdist): occupied[i] = the frame's
// keypoint i already holds a map point; already_found[j] = keyframe map point j is in sAlreadyFound;
// frame_match[i] = keyframe feature whose map point the call assigned to frame keypoint i, else -1
int ref_search_by_projection(const orc_kf_view *vf, const orc_kf_view *vk, const float *K, const float *Rcw, const float *tcw, float th, int orb_dist,
                             int check_orientation, const uint8_t *occupied, const uint8_t *already_found, int32_t *frame_match)
{
    ORBmatcher matcher(0.9f, check_orientation != 0);
    std::shared_ptr<KeyFrame> kf = kf_from_view(vk, K);
    Frame F;
    F.N = vf->n_feat;
    F.mpView = vf;
    fill_view_common(vf, K, F.mvKeysUn, F.mDescriptors, F.mvuRight, F.mvScaleFactors, F.mvLevelSigma2, F.mvInvLevelSigma2);
    F.mvKeys = F.mvKeysUn;
    F.mnScaleLevels = vf->n_levels;
    F.mfLogScaleFactor = vf->log_scale_factor;
    F.fx = K[0]; F.fy = K[1]; F.cx = K[2]; F.cy = K[3];
    F.mnMinX = vf->bounds[0]; F.mnMaxX = vf->bounds[1]; F.mnMinY = vf->bounds[2]; F.mnMaxY = vf->bounds[3];
    F.mvpMapPoints.resize(vf->n_feat);
    F.mvbOutlier.assign(vf->n_feat, false);
    auto placeholder = std::make_shared<MapPoint>();          // "this keypoint is taken" (any non-null pointer)
    for (int i = 0; i < vf->n_feat; ++i)
        if (occupied && occupied[i]) F.mvpMapPoints[i] = placeholder;
    Eigen::Matrix3f R;
    Eigen::Vector3f t;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R(i, j) = Rcw[i * 3 + j];
        t(i) = tcw[i];
    }
    F.mTcw.setIdentity();
    F.mTcw.linear() = R;
    F.mTcw.translation() = t;
    std::set<std::shared_ptr<MapPoint>> found;
    for (int j = 0; j < vk->n_feat; ++j)
        if (already_found && already_found[j] && kf->mvpMapPoints[j]) found.insert(kf->mvpMapPoints[j]);
    const int n = matcher.SearchByProjection(F, kf, found, th, orb_dist);
    for (int i = 0; i < vf->n_feat; ++i)
        frame_match[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != placeholder) ? F.mvpMapPoints[i]->mIndexKF1 : -1;
    return n;
}

}  // extern "C"
