// ransac_b200/solvers.hpp -- the reference's solver classes over the B200 engine.
//
// Same class names, method names, argument meaning and return/`bNoMore` behaviour as
//   PnPsolver    reference include/PnPsolver.hpp:21-31,   src/PnPsolver.cpp
//   MLPnPsolver  reference include/MLPnPsolver.hpp:10-21, src/MLPnPsolver.cpp
//   Sim3Solver   reference include/Sim3Solver.hpp:16-30,  src/Sim3Solver.cpp
// plus what BASELINE.json's north_star adds: GetEstimatedScale()/bFixScale and batched entry
// points (PnPsolver::SolveBatch, Sim3Solver::SolveBatch) that Tracking::Relocalization
// (src/Tracking.cpp:1225-1255) and LoopClosing::ComputeSim3 (src/LoopClosing.cpp:260-308) call
// once with all candidates before their iterate() loops.
//
// The reference classes take ORB-SLAM objects (Frame, KeyFrame, MapPoint) and Eigen types.
// Neither Eigen nor OpenCV is available to this repository, so the constructors take plain
// "views" holding exactly the fields the reference constructors read (PnPsolver.cpp:11-55,
// Sim3Solver.cpp:6-85, MLPnPsolver.cpp:5-53); INTEGRATION.md shows the ten-line adapters that
// build them from the real objects and how Matrix4f/Matrix3f/Vector3f map onto Eigen.
// Header-only; links against libransac_b200.so (include/ransac_b200.h).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ransac_b200.h"

namespace ransac_b200 {

// ---- minimal fixed-size types (row-major); with Eigen: Eigen::Map<Eigen::Matrix<float,4,4,Eigen::RowMajor>>(T.m)
struct Matrix4f {
    float m[16];
    Matrix4f() { setIdentity(); }
    void setIdentity() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float& operator()(int r, int c) { return m[r * 4 + c]; }
    float operator()(int r, int c) const { return m[r * 4 + c]; }
};
struct Matrix3f {
    float m[9];
    Matrix3f() { for (int i = 0; i < 9; ++i) m[i] = (i % 4 == 0) ? 1.f : 0.f; }
    float& operator()(int r, int c) { return m[r * 3 + c]; }
    float operator()(int r, int c) const { return m[r * 3 + c]; }
};
struct Vector3f {
    float v[3];
    Vector3f() { v[0] = v[1] = v[2] = 0.f; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};

// ---- engine handle shared by the solvers of a process (one per device) -------------------
class Engine {
public:
    explicit Engine(int device = 0)
    {
        const int rc = rsac_create(device, &h_);
        if (rc != RSAC_OK) throw std::runtime_error(rc == RSAC_ERR_NO_DEVICE ? "ransac_b200: no CUDA device (there is no CPU fallback)"
                                                                              : "ransac_b200: rsac_create failed");
    }
    ~Engine() { rsac_destroy(h_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    rsac_engine* handle() { return h_; }
    std::mutex& mutex() { return mu_; }          // Tracking and LoopClosing threads may share one engine
    uint64_t next_epoch() { return ++epoch_; }
    uint64_t epoch() const { return epoch_; }
    static Engine& Default()
    {
        static Engine e(0);
        return e;
    }
private:
    rsac_engine* h_ = nullptr;
    std::mutex mu_;
    uint64_t epoch_ = 0;
};

inline void check(int rc, rsac_engine* e, const char* what)
{
    if (rc != RSAC_OK) throw std::runtime_error(std::string("ransac_b200: ") + what + ": " + rsac_last_error(e));
}

// ---- views: the fields the reference constructors read -----------------------------------
struct FrameView {                 // Frame (PnPsolver.cpp:16-50, MLPnPsolver.cpp:15-47)
    int n_keypoints = 0;           // F.mvKeysUn.size()
    const float* keys_xy = nullptr;        // [n][2] mvKeysUn[i].pt
    const int* octave = nullptr;           // [n]    mvKeysUn[i].octave
    const float* level_sigma2 = nullptr;   // mvLevelSigma2
    float fx = 0, fy = 0, cx = 0, cy = 0;  // Frame::fx.. (static floats, Frame.hpp:102-105)
};
struct MapPointMatches {           // vpMapPointMatches: one slot per keypoint
    int n = 0;
    const unsigned char* valid = nullptr;  // [n] pMP != nullptr && !pMP->isBad()
    const float* world_pos = nullptr;      // [n][3] pMP->GetWorldPos()
};
struct KeyFrameView {              // KeyFrame (Sim3Solver.cpp:9-79)
    float Rcw[9];                  // GetRotation(), row-major
    float tcw[3];                  // GetTranslation()
    int n_keypoints = 0;
    const float* keys_xy = nullptr;        // unused by the solver (only octaves are read)
    const int* octave = nullptr;           // mvKeysUn[i].octave
    const float* level_sigma2 = nullptr;   // mvLevelSigma2
    float fx = 0, fy = 0, cx = 0, cy = 0;  // mK
};
struct Sim3Matches {               // per keypoint i1 of KF1 (length = vpMatched12.size() = mN1)
    int n = 0;
    const unsigned char* valid1 = nullptr;   // vpKeyFrameMP1[i1] non-null and not bad
    const float* world_pos1 = nullptr;       // [n][3]
    const int* index_in_kf1 = nullptr;       // pMP1->GetIndexInKeyFrame(pKF1)  (< 0: not observed)
    const unsigned char* valid2 = nullptr;   // vpMatched12[i1] non-null (and not bad)
    const float* world_pos2 = nullptr;       // [n][3]
    const int* index_in_kf2 = nullptr;       // pMP2->GetIndexInKeyFrame(pKF2)
};

inline void unpack_mask(const uint32_t* words, int n, const std::vector<size_t>& scatter, std::vector<bool>& out)
{
    for (int i = 0; i < n; ++i)
        if ((words[i >> 5] >> (i & 31)) & 1u) out[scatter[i]] = true;
}

// =========================================================================== PnPsolver
class PnPsolver {
public:
    // PnPsolver.cpp:11-55
    PnPsolver(const FrameView& F, const MapPointMatches& vpMapPointMatches, Engine* engine = nullptr)
        : eng_(engine ? engine : &Engine::Default())
    {
        N_points = vpMapPointMatches.n;
        for (int i = 0; i < vpMapPointMatches.n; ++i) {
            if (!vpMapPointMatches.valid[i]) continue;
            p2d_.push_back(F.keys_xy[2 * i]); p2d_.push_back(F.keys_xy[2 * i + 1]);
            sigma2_.push_back(F.level_sigma2[F.octave[i]]);
            for (int c = 0; c < 3; ++c) p3d_.push_back(vpMapPointMatches.world_pos[3 * i + c]);
            mvKeyPointIndices.push_back((size_t)i);
        }
        K_[0] = F.fx; K_[1] = F.fy; K_[2] = F.cx; K_[3] = F.cy;   // widened to double (PnPsolver.hpp:71)
        SetRansacParameters();   // the reference leaves the parameters unset until the caller does this
    }

    // PnPsolver.hpp:26-27 / PnPsolver.cpp:58-94
    void SetRansacParameters(double probability = 0.99, int minInliers = 8, int maxIterations = 300, int minSet = 4,
                             float epsilon = 0.4f, float th2 = 5.991f)
    {
        prm_.prob = probability; prm_.min_inliers = minInliers; prm_.max_its = maxIterations;
        prm_.min_set = minSet; prm_.eps = epsilon; prm_.th2 = th2;
        N = (int)sigma2_.size();
        if (N > 0) rsac_pnp_ransac_setup(N, &prm_, &mRansacMinInliers, &mRansacMaxIts);
        else { mRansacMinInliers = std::max(minInliers, minSet); mRansacMaxIts = 0; }
        cursor_ = 0; epoch_ = 0; cached_ = false;
    }

    // per-solver index stream: srand(seed) semantics, private to this solver (the reference shares
    // one unseeded global rand() between threads)
    void SetSeed(uint32_t seed) { seed_ = seed; table_.clear(); cached_ = false; epoch_ = 0; }
    // explicit minimal sets, H x minSet indices into the compact correspondence list
    void SetIndexTable(const std::vector<uint32_t>& t) { table_ = t; cached_ = false; epoch_ = 0; }
    // engine mode bits ORed into every run of this solver, e.g. RSAC_FLAG_EPNP_EIGEN: the reference's own 12 x 12
    // eigen-solve per hypothesis instead of the default QR null space -- with it the class returns
    // what the reference's PnPsolver returns bit for bit (tests/cpp/dropin_vs_reference.cpp)
    void SetEngineFlags(int flags) { flags_ = flags; cached_ = false; epoch_ = 0; }

    // PnPsolver.cpp:96-100
    bool find(std::vector<bool>& vbInliers, int& nInliers, Matrix4f& T)
    {
        bool bFlag;
        return iterate(mRansacMaxIts, bFlag, vbInliers, nInliers, T);
    }

    // PnPsolver.cpp:102-191.  The first call consumes the whole budget (`||` at :119): here it is
    // one batched device pass over all mRansacMaxIts hypotheses followed by the sequential replay;
    // later calls resume the replay where the previous one returned.
    bool iterate(int /*nIterations*/, bool& bNoMore, std::vector<bool>& vbInliers, int& nInliers, Matrix4f& T)
    {
        bNoMore = false;
        vbInliers.clear();
        nInliers = 0;
        if (N < mRansacMinInliers) { bNoMore = true; return false; }
        // Budget spent by earlier calls: the reference's loop (`mnIterations<mRansacMaxIts || nCurrentIterations<nIterations`,
        // PnPsolver.cpp:119) would draw nIterations FRESH minimal sets beyond the budget and then fall back to the best
        // set so far (:173-188).  The extra draws are not performed here (deviations); the fallback is:
        // the rerun below resumes at the end of the table, finds nothing and returns mBestTcw / mvbBestInliers with
        // bNoMore = true when mnBestInliers >= mRansacMinInliers.
        std::lock_guard<std::mutex> lock(eng_->mutex());
        rsac_result r;
        std::vector<uint32_t> words((size_t)(N + 31) / 32);
        if (cached_ && cursor_ == 0) {
            r = cache_; words = cache_mask_;
        } else {
            ensure_resident();
            if (cursor_ > 0) {
                int32_t resume = cursor_;
                check(rsac_pnp_rerun(eng_->handle(), flags_, &resume, nullptr), eng_->handle(), "rsac_pnp_rerun");
            }
            check(rsac_pnp_download(eng_->handle(), &r, words.data()), eng_->handle(), "rsac_pnp_download");
        }
        cached_ = false;
        cursor_ = r.n_hyp;
        bNoMore = r.no_more != 0;
        if (!r.ok) return false;
        nInliers = r.n_inliers;
        vbInliers.assign((size_t)N_points, false);
        unpack_mask(words.data(), N, mvKeyPointIndices, vbInliers);
        T.setIdentity();
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T(i, j) = r.R[3 * i + j]; T(i, 3) = r.t[i]; }
        return true;
    }

    // Batched entry point for Tracking::Relocalization: all candidate solvers in ONE device pass
    // (one problem per solver); each solver's first iterate()/find() then returns its cached result.
    static void SolveBatch(const std::vector<PnPsolver*>& solvers)
    {
        if (solvers.empty()) return;
        Engine* eng = solvers[0]->eng_;
        std::lock_guard<std::mutex> lock(eng->mutex());
        const int C = (int)solvers.size();
        std::vector<int32_t> offsets(C + 1, 0);
        std::vector<float> p3d, p2d, s2;
        std::vector<double> K;
        std::vector<rsac_ransac_params> prm;
        std::vector<uint32_t> seeds, tables;
        std::vector<int64_t> toff(C + 1, 0);
        bool any_table = false;
        for (auto* s : solvers) any_table = any_table || !s->table_.empty();
        for (int c = 0; c < C; ++c) {
            PnPsolver* s = solvers[c];
            offsets[c + 1] = offsets[c] + s->N;
            p3d.insert(p3d.end(), s->p3d_.begin(), s->p3d_.end());
            p2d.insert(p2d.end(), s->p2d_.begin(), s->p2d_.end());
            s2.insert(s2.end(), s->sigma2_.begin(), s->sigma2_.end());
            K.insert(K.end(), s->K_, s->K_ + 4);
            prm.push_back(s->prm_);
            seeds.push_back(s->seed_);
            if (any_table) {
                std::vector<uint32_t> t = s->table_;
                const size_t need = (size_t)std::max(0, s->mRansacMaxIts) * s->prm_.min_set;
                if (t.size() < need && s->N >= s->prm_.min_set) {
                    t.resize(need);
                    rsac_index_table(s->seed_, s->N, s->prm_.min_set, s->mRansacMaxIts, t.data());
                }
                tables.insert(tables.end(), t.begin(), t.end());
                toff[c + 1] = (int64_t)tables.size();
            }
        }
        rsac_pnp_batch b;
        std::memset(&b, 0, sizeof(b));
        b.C = C; b.offsets = offsets.data(); b.p3d = p3d.data(); b.p2d = p2d.data(); b.sigma2 = s2.data(); b.K = K.data();
        b.params = prm.data(); b.n_params = C; b.seeds = seeds.data();
        if (any_table) { b.tables = tables.data(); b.table_offsets = toff.data(); }
        std::vector<rsac_result> res(C);
        int64_t nwords = 0;
        for (auto* s : solvers) nwords += (s->N + 31) / 32;
        std::vector<uint32_t> masks((size_t)std::max<int64_t>(nwords, 1));
        // early exit in phases: identical records, hypotheses behind the reference's stopping point are not computed
        // (batches too small to fill one solver wave run all hypotheses at once)
        check(rsac_pnp_solve(eng->handle(), &b, RSAC_FLAG_EARLY_EXIT | solvers[0]->flags_, res.data(), masks.data()), eng->handle(), "rsac_pnp_solve");
        eng->next_epoch();
        size_t w0 = 0;
        for (int c = 0; c < C; ++c) {
            PnPsolver* s = solvers[c];
            const size_t nw = (size_t)(s->N + 31) / 32;
            s->cache_ = res[c];
            s->cache_mask_.assign(masks.begin() + w0, masks.begin() + w0 + nw);
            s->cached_ = true; s->cursor_ = 0; s->epoch_ = 0;   // a later resumed call re-uploads this solver alone
            w0 += nw;
        }
    }

    int GetIterations() const { return mRansacMaxIts; }
    int GetMinInliers() const { return mRansacMinInliers; }
    int GetNumCorrespondences() const { return N; }

private:
    void ensure_resident()
    {
        if (epoch_ != 0 && epoch_ == eng_->epoch()) return;   // the engine still holds this solver's hypotheses
        int32_t offsets[2] = {0, N};
        rsac_pnp_batch b;
        std::memset(&b, 0, sizeof(b));
        b.C = 1; b.offsets = offsets; b.p3d = p3d_.data(); b.p2d = p2d_.data(); b.sigma2 = sigma2_.data(); b.K = K_;
        b.params = &prm_; b.n_params = 1; b.seeds = &seed_;
        int64_t toff[2] = {0, (int64_t)table_.size()};
        if (!table_.empty()) { b.tables = table_.data(); b.table_offsets = toff; }
        check(rsac_pnp_upload(eng_->handle(), &b), eng_->handle(), "rsac_pnp_upload");
        check(rsac_pnp_run(eng_->handle(), flags_, nullptr), eng_->handle(), "rsac_pnp_run");
        epoch_ = eng_->next_epoch();
    }

    Engine* eng_;
    int N_points = 0, N = 0;
    std::vector<float> p3d_, p2d_, sigma2_;
    std::vector<size_t> mvKeyPointIndices;
    double K_[4];
    rsac_ransac_params prm_;
    int mRansacMinInliers = 0, mRansacMaxIts = 0;
    uint32_t seed_ = 1;     // glibc: an unseeded rand() behaves like srand(1)
    int flags_ = 0;         // SetEngineFlags
    std::vector<uint32_t> table_;
    int cursor_ = 0;        // hypotheses consumed so far (mnIterations)
    uint64_t epoch_ = 0;
    bool cached_ = false;
    rsac_result cache_;
    std::vector<uint32_t> cache_mask_;
};

// ========================================================================= MLPnPsolver
class MLPnPsolver {
public:
    // MLPnPsolver.cpp:5-53 (calls SetRansacParameters() itself, :52)
    MLPnPsolver(const FrameView& F, const MapPointMatches& vpMapPointMatches, Engine* engine = nullptr)
        : eng_(engine ? engine : &Engine::Default())
    {
        N_points = vpMapPointMatches.n;
        for (int i = 0; i < vpMapPointMatches.n; ++i) {
            if (!vpMapPointMatches.valid[i]) continue;
            if (i >= F.n_keypoints) continue;                               // :26
            p2d_.push_back(F.keys_xy[2 * i]); p2d_.push_back(F.keys_xy[2 * i + 1]);
            sigma2_.push_back(F.level_sigma2[F.octave[i]]);
            for (int c = 0; c < 3; ++c) p3d_.push_back(vpMapPointMatches.world_pos[3 * i + c]);
            mvKeyPointIndices.push_back((size_t)i);
        }
        K_[0] = F.fx; K_[1] = F.fy; K_[2] = F.cx; K_[3] = F.cy;
        SetRansacParameters();
    }

    // MLPnPsolver.hpp:16-17
    void SetRansacParameters(double probability = 0.99, int minInliers = 8, int maxIterations = 300, int minSet = 6,
                             float epsilon = 0.4f, float th2 = 5.991f)
    {
        prm_.prob = probability; prm_.min_inliers = minInliers; prm_.max_its = maxIterations;
        prm_.min_set = minSet; prm_.eps = epsilon; prm_.th2 = th2;
        N = (int)sigma2_.size();
        if (N > 0) rsac_pnp_ransac_setup(N, &prm_, &mRansacMinInliers, &mRansacMaxIts);
        else { mRansacMinInliers = std::max(minInliers, minSet); mRansacMaxIts = 0; }
        cursor_ = 0; epoch_ = 0;
    }
    void SetSeed(uint32_t seed) { seed_ = seed; epoch_ = 0; }
    // optional bearing covariances, one 3x3 per keypoint slot (use_cov branch, MLPnPsolver.cpp:375-388);
    // the reference passes covs(1), i.e. none
    void SetBearingCovariances(const double* cov_per_keypoint)
    {
        cov_.clear();
        for (size_t k : mvKeyPointIndices) cov_.insert(cov_.end(), cov_per_keypoint + 9 * k, cov_per_keypoint + 9 * k + 9);
        epoch_ = 0;
    }
    // reproduce MLPnPsolver::Refine as shipped (its pose is never stored, MLPnPsolver.cpp:290-296)
    void SetDiscardRefine(bool on) { flags_ = on ? RSAC_FLAG_MLPNP_DISCARD_REFINE : 0; }

    // MLPnPsolver.cpp:56-183
    bool iterate(int /*nIterations*/, bool& bNoMore, std::vector<bool>& vbInliers, int& nInliers, Matrix4f& Tout)
    {
        Tout.setIdentity();
        bNoMore = false;
        vbInliers.clear();
        nInliers = 0;
        if (N < mRansacMinInliers) { bNoMore = true; return false; }
        // budget spent: best-so-far fallback with bNoMore = true (see PnPsolver::iterate above; MLPnPsolver.cpp:71,165-180)
        std::lock_guard<std::mutex> lock(eng_->mutex());
        if (!(epoch_ != 0 && epoch_ == eng_->epoch())) {
            int32_t offsets[2] = {0, N};
            rsac_mlpnp_batch b;
            std::memset(&b, 0, sizeof(b));
            b.C = 1; b.offsets = offsets; b.p3d = p3d_.data(); b.p2d = p2d_.data(); b.sigma2 = sigma2_.data(); b.K = K_;
            b.cov = cov_.empty() ? nullptr : cov_.data();
            b.params = &prm_; b.n_params = 1; b.seeds = &seed_;
            check(rsac_mlpnp_upload(eng_->handle(), &b), eng_->handle(), "rsac_mlpnp_upload");
            check(rsac_mlpnp_run(eng_->handle(), flags_, nullptr), eng_->handle(), "rsac_mlpnp_run");
            epoch_ = eng_->next_epoch();
        }
        if (cursor_ > 0) {
            int32_t resume = cursor_;
            check(rsac_mlpnp_rerun(eng_->handle(), flags_, &resume, nullptr), eng_->handle(), "rsac_mlpnp_rerun");
        }
        rsac_result r;
        std::vector<uint32_t> words((size_t)(N + 31) / 32);
        check(rsac_mlpnp_download(eng_->handle(), &r, words.data()), eng_->handle(), "rsac_mlpnp_download");
        cursor_ = r.n_hyp;
        bNoMore = r.no_more != 0;
        if (!r.ok) return false;
        nInliers = r.n_inliers;
        vbInliers.assign((size_t)N_points, false);
        unpack_mask(words.data(), N, mvKeyPointIndices, vbInliers);
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Tout(i, j) = r.R[3 * i + j]; Tout(i, 3) = r.t[i]; }
        return true;
    }
    int GetIterations() const { return mRansacMaxIts; }

private:
    Engine* eng_;
    int N_points = 0, N = 0;
    std::vector<float> p3d_, p2d_, sigma2_;
    std::vector<double> cov_;
    std::vector<size_t> mvKeyPointIndices;
    float K_[4];
    rsac_ransac_params prm_;
    int mRansacMinInliers = 0, mRansacMaxIts = 0, flags_ = 0;
    uint32_t seed_ = 1;
    int cursor_ = 0;
    uint64_t epoch_ = 0;
};

// ========================================================================== Sim3Solver
class Sim3Solver {
public:
    // Sim3Solver.cpp:6-85; bFixScale as in upstream ORB-SLAM2 (the reference is fixed-scale only, :250)
    Sim3Solver(const KeyFrameView& KF1, const KeyFrameView& KF2, const Sim3Matches& vpMatched12, bool bFixScale = true,
               Engine* engine = nullptr)
        : eng_(engine ? engine : &Engine::Default()), mbFixScale(bFixScale)
    {
        mN1 = vpMatched12.n;
        for (int i1 = 0; i1 < mN1; ++i1) {
            if (!vpMatched12.valid2[i1]) continue;                          // :28
            if (!vpMatched12.valid1[i1]) continue;                          // :33-37
            const int indexKF1 = vpMatched12.index_in_kf1[i1], indexKF2 = vpMatched12.index_in_kf2[i1];
            if (indexKF1 < 0 || indexKF2 < 0) continue;                     // :42-43
            sigma2_1_.push_back(KF1.level_sigma2[KF1.octave[indexKF1]]);    // :48-52 (thresholds formed on the device)
            sigma2_2_.push_back(KF2.level_sigma2[KF2.octave[indexKF2]]);
            mvnIndices1.push_back((size_t)i1);
            const float* w1 = vpMatched12.world_pos1 + 3 * i1;
            const float* w2 = vpMatched12.world_pos2 + 3 * i1;
            for (int r = 0; r < 3; ++r)                                     // X3Dc = Rcw*X3Dw + tcw in f32 (:57-63)
                x1c_.push_back((KF1.Rcw[3 * r] * w1[0] + KF1.Rcw[3 * r + 1] * w1[1] + KF1.Rcw[3 * r + 2] * w1[2]) + KF1.tcw[r]);
            for (int r = 0; r < 3; ++r)
                x2c_.push_back((KF2.Rcw[3 * r] * w2[0] + KF2.Rcw[3 * r + 1] * w2[1] + KF2.Rcw[3 * r + 2] * w2[2]) + KF2.tcw[r]);
        }
        K1_[0] = KF1.fx; K1_[1] = KF1.fy; K1_[2] = KF1.cx; K1_[3] = KF1.cy;
        K2_[0] = KF2.fx; K2_[1] = KF2.fy; K2_[2] = KF2.cx; K2_[3] = KF2.cy;
        SetRansacParameters();                                              // :84
    }

    // Sim3Solver.hpp:23 / Sim3Solver.cpp:87-111 (resets mnIterations, :110)
    void SetRansacParameters(double probability = 0.99, int minInliers = 6, int maxIterations = 300)
    {
        prm_.prob = probability; prm_.min_inliers = minInliers; prm_.max_its = maxIterations; prm_.fix_scale = mbFixScale ? 1 : 0;
        N = (int)sigma2_1_.size();
        mRansacMinInliers = minInliers;
        mRansacMaxIts = 0;
        if (N > 0) rsac_sim3_ransac_setup(N, &prm_, &mRansacMaxIts);
        mnIterations = 0; mnBestInliers = 0; have_hyp_ = false;
    }
    void SetSeed(uint32_t seed) { seed_ = seed; have_hyp_ = false; }

    // Sim3Solver.cpp:113-178: nIterations hypotheses per call, state persisting across calls
    bool iterate(int nIterations, bool& bNoMore, std::vector<bool>& vbInliers, int& nInliers)
    {
        bNoMore = false;
        vbInliers.assign((size_t)mN1, false);
        nInliers = 0;
        if (N < mRansacMinInliers) { bNoMore = true; return false; }
        if (!have_hyp_) { std::vector<Sim3Solver*> me{this}; SolveBatch(me); }
        int nCurrentIterations = 0;
        const int words = (N + 31) / 32;
        while (mnIterations < mRansacMaxIts && nCurrentIterations < nIterations) {
            const int h = mnIterations;
            nCurrentIterations++;
            mnIterations++;
            const int cnt = counts_[h];
            if (cnt >= mnBestInliers) {                                    // :155
                mnBestInliers = cnt;
                best_h_ = h;
                if (cnt > mRansacMinInliers) {                             // :163
                    nInliers = cnt;
                    unpack_mask(hmasks_.data() + (size_t)h * words, N, mvnIndices1, vbInliers);
                    return true;
                }
            }
        }
        if (mnIterations >= mRansacMaxIts) bNoMore = true;                 // :174-175
        return false;
    }

    // Sim3Solver.cpp:180-184
    bool find(std::vector<bool>& vbInliers12, int& nInliers)
    {
        bool bFlag;
        return iterate(mRansacMaxIts, bFlag, vbInliers12, nInliers);
    }

    // Sim3Solver.cpp:296-304 (+ scale, upstream)
    Matrix3f GetEstimatedRotation() const
    {
        Matrix3f R;
        if (best_h_ >= 0) std::memcpy(R.m, poses_.data() + (size_t)best_h_ * 13, 9 * sizeof(float));
        return R;
    }
    Vector3f GetEstimatedTranslation() const
    {
        Vector3f t;
        if (best_h_ >= 0) std::memcpy(t.v, poses_.data() + (size_t)best_h_ * 13 + 9, 3 * sizeof(float));
        return t;
    }
    float GetEstimatedScale() const { return best_h_ >= 0 ? poses_[(size_t)best_h_ * 13 + 12] : 1.0f; }

    // Batched entry point for LoopClosing::ComputeSim3: every candidate's hypotheses in one device pass
    static void SolveBatch(const std::vector<Sim3Solver*>& solvers)
    {
        if (solvers.empty()) return;
        Engine* eng = solvers[0]->eng_;
        std::lock_guard<std::mutex> lock(eng->mutex());
        const int C = (int)solvers.size();
        std::vector<int32_t> offsets(C + 1, 0);
        std::vector<float> x1, x2, s1, s2, K1, K2;
        std::vector<rsac_sim3_params> prm;
        std::vector<uint32_t> seeds;
        int64_t hyp = 0, hwords = 0;
        for (int c = 0; c < C; ++c) {
            Sim3Solver* s = solvers[c];
            offsets[c + 1] = offsets[c] + s->N;
            x1.insert(x1.end(), s->x1c_.begin(), s->x1c_.end());
            x2.insert(x2.end(), s->x2c_.begin(), s->x2c_.end());
            s1.insert(s1.end(), s->sigma2_1_.begin(), s->sigma2_1_.end());
            s2.insert(s2.end(), s->sigma2_2_.begin(), s->sigma2_2_.end());
            K1.insert(K1.end(), s->K1_, s->K1_ + 4);
            K2.insert(K2.end(), s->K2_, s->K2_ + 4);
            prm.push_back(s->prm_);
            seeds.push_back(s->seed_);
            const int H = (s->N >= std::max(s->mRansacMinInliers, 3)) ? s->mRansacMaxIts : 0;
            hyp += H;
            hwords += (int64_t)H * ((s->N + 31) / 32);
        }
        rsac_sim3_batch b;
        std::memset(&b, 0, sizeof(b));
        b.C = C; b.offsets = offsets.data(); b.x1c = x1.data(); b.x2c = x2.data(); b.sigma2_1 = s1.data(); b.sigma2_2 = s2.data();
        b.K1 = K1.data(); b.K2 = K2.data(); b.params = prm.data(); b.n_params = C; b.seeds = seeds.data();
        check(rsac_sim3_upload(eng->handle(), &b), eng->handle(), "rsac_sim3_upload");
        check(rsac_sim3_run(eng->handle(), 0, nullptr), eng->handle(), "rsac_sim3_run");
        std::vector<float> poses((size_t)std::max<int64_t>(hyp, 1) * 13);
        std::vector<int32_t> counts((size_t)std::max<int64_t>(hyp, 1));
        std::vector<uint32_t> hm((size_t)std::max<int64_t>(hwords, 1));
        check(rsac_sim3_get_hypotheses(eng->handle(), poses.data(), counts.data(), hm.data()), eng->handle(), "rsac_sim3_get_hypotheses");
        eng->next_epoch();
        size_t h0 = 0, w0 = 0;
        for (int c = 0; c < C; ++c) {
            Sim3Solver* s = solvers[c];
            const int H = (s->N >= std::max(s->mRansacMinInliers, 3)) ? s->mRansacMaxIts : 0;
            const size_t nw = (size_t)H * ((s->N + 31) / 32);
            s->poses_.assign(poses.begin() + h0 * 13, poses.begin() + (h0 + H) * 13);
            s->counts_.assign(counts.begin() + h0, counts.begin() + h0 + H);
            s->hmasks_.assign(hm.begin() + w0, hm.begin() + w0 + nw);
            s->have_hyp_ = true;
            h0 += H; w0 += nw;
        }
    }
    int GetIterations() const { return mRansacMaxIts; }

private:
    Engine* eng_;
    bool mbFixScale;
    int mN1 = 0, N = 0;
    std::vector<float> x1c_, x2c_, sigma2_1_, sigma2_2_;
    std::vector<size_t> mvnIndices1;
    float K1_[4], K2_[4];
    rsac_sim3_params prm_;
    int mRansacMinInliers = 0, mRansacMaxIts = 0;
    int mnIterations = 0, mnBestInliers = 0, best_h_ = -1;
    uint32_t seed_ = 1;
    bool have_hyp_ = false;
    std::vector<float> poses_;
    std::vector<int32_t> counts_;
    std::vector<uint32_t> hmasks_;
};

// =========================================================================== Optimizer
// Mirror of `int Optimizer::PoseOptimization(Frame* pFrame)` (include/Optimizer.hpp, src/Optimizer.cpp:205-424):
// reads pFrame->mTcw, mvpMapPoints, mvKeysUn, mvuRight, mvInvLevelSigma2, fx/fy/cx/cy/mbf; writes pFrame->mvbOutlier
// for every keypoint that has a MapPoint, sets the pose (Frame::SetPose) and returns nInitialCorrespondences - nBad.
// Tracking::Relocalization calls it for every candidate RANSAC accepts (Tracking.cpp:1284,1300,1315):
// `PoseOptimizationBatch` optimises all of them in one device pass.
struct PoseOptFrame {
    // in
    int n_keypoints = 0;                       // pFrame->N
    const float* keys_xy = nullptr;            // [N][2] mvKeysUn[i].pt
    const int* octave = nullptr;               // [N]    mvKeysUn[i].octave
    const float* u_right = nullptr;            // [N]    mvuRight[i] (< 0: monocular); nullptr = all monocular
    const float* inv_level_sigma2 = nullptr;   // mvInvLevelSigma2
    const unsigned char* has_map_point = nullptr;  // [N] pFrame->mvpMapPoints[i] != nullptr
    const float* world_pos = nullptr;          // [N][3] pMP->GetWorldPos()
    float fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
    // in/out
    Matrix4f Tcw;                              // pFrame->mTcw (SetPose on return)
    std::vector<bool>* outlier = nullptr;      // pFrame->mvbOutlier (size N; entries with a MapPoint are written)
    // out
    int n_inliers = 0;                         // return value
};

class Optimizer {
public:
    static int PoseOptimization(PoseOptFrame* pFrame, Engine& engine = Engine::Default())
    {
        std::vector<PoseOptFrame*> v{pFrame};
        PoseOptimizationBatch(v, engine);
        return pFrame->n_inliers;
    }

    static void PoseOptimizationBatch(const std::vector<PoseOptFrame*>& frames, Engine& engine = Engine::Default())
    {
        if (frames.empty()) return;
        std::lock_guard<std::mutex> lock(engine.mutex());
        const int C = (int)frames.size();
        std::vector<int32_t> offsets(C + 1, 0);
        std::vector<float> p3d, obs, isig, K, T;
        std::vector<std::vector<int>> index(C);
        for (int c = 0; c < C; ++c) {
            const PoseOptFrame& f = *frames[c];
            for (int i = 0; i < f.n_keypoints; ++i) {
                if (!f.has_map_point[i]) continue;
                index[c].push_back(i);
                p3d.insert(p3d.end(), f.world_pos + 3 * i, f.world_pos + 3 * i + 3);
                obs.push_back(f.keys_xy[2 * i]);
                obs.push_back(f.keys_xy[2 * i + 1]);
                obs.push_back(f.u_right ? f.u_right[i] : -1.0f);
                isig.push_back(f.inv_level_sigma2[f.octave[i]]);
            }
            offsets[c + 1] = offsets[c] + (int32_t)index[c].size();
            const float k[5] = {f.fx, f.fy, f.cx, f.cy, f.bf};
            K.insert(K.end(), k, k + 5);
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q) T.push_back(f.Tcw(r, q));
            for (int r = 0; r < 3; ++r) T.push_back(f.Tcw(r, 3));
        }
        rsac_poseopt_batch b;
        std::memset(&b, 0, sizeof(b));
        b.C = C; b.offsets = offsets.data(); b.p3d = p3d.data(); b.obs = obs.data(); b.inv_sigma2 = isig.data();
        b.K = K.data(); b.Tcw = T.data();
        std::vector<rsac_poseopt_result> res(C);
        std::vector<uint8_t> out((size_t)std::max(offsets[C], 1));
        check(rsac_poseopt_solve(engine.handle(), &b, res.data(), out.data()), engine.handle(), "rsac_poseopt_solve");
        for (int c = 0; c < C; ++c) {
            PoseOptFrame& f = *frames[c];
            f.n_inliers = res[c].n_inliers;
            if (f.outlier)
                for (size_t k = 0; k < index[c].size(); ++k) (*f.outlier)[(size_t)index[c][k]] = out[(size_t)offsets[c] + k] != 0;
            if (index[c].size() < 3) continue;      // `return 0` before the pose is touched (Optimizer.cpp:326-327)
            for (int r = 0; r < 3; ++r) {
                for (int q = 0; q < 3; ++q) f.Tcw.m[4 * r + q] = res[c].Rf[3 * r + q];
                f.Tcw.m[4 * r + 3] = res[c].tf[r];
            }
        }
    }
};

// Mirror of `int Optimizer::OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
// const float th2)` (src/Optimizer.cpp:1054-1249; LoopClosing.cpp:311 calls it with th2 = 10 for every candidate whose
// Sim3Solver succeeded).  Reads both keyframes' poses, calibrations, undistorted keypoints and level sigmas and the two
// map points of every match; nulls the matches it rejects, writes the refined Sim3 back, returns the inlier count.
struct Sim3Value {                 // g2o::Sim3 as the caller holds it
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double t[3] = {0, 0, 0};
    double s = 1.0;
};
struct Sim3OptPair {
    // in
    KeyFrameView kf1, kf2;                     // GetRotation/GetTranslation, mK, mvKeysUn (keys_xy), octave, level_sigma2
    const float* inv_level_sigma2_1 = nullptr; // pKF1->mvInvLevelSigma2
    const float* inv_level_sigma2_2 = nullptr;
    int n = 0;                                 // vpMatches1.size()
    const unsigned char* valid1 = nullptr;     // [n] pKF1's map point i exists and is not bad
    const float* world_pos1 = nullptr;         // [n][3]
    const unsigned char* valid2 = nullptr;     // [n] vpMatches1[i] exists and is not bad
    const float* world_pos2 = nullptr;         // [n][3]
    const int* index_in_kf2 = nullptr;         // [n] pMP2->GetIndexInKeyFrame(pKF2) (< 0: skipped)
    float th2 = 10.0f;
    bool fix_scale = true;                     // the reference hard-codes true (Optimizer.cpp:1076)
    // in/out
    Sim3Value S12;                             // g2oS12
    std::vector<bool>* match_alive = nullptr;  // [n] false where the reference sets vpMatches1[i] = nullptr
    // out
    int n_inliers = 0;
};

class Sim3Optimizer {
public:
    static int OptimizeSim3(Sim3OptPair* pair, Engine& engine = Engine::Default())
    {
        std::vector<Sim3OptPair*> v{pair};
        OptimizeSim3Batch(v, engine);
        return pair->n_inliers;
    }

    static void OptimizeSim3Batch(const std::vector<Sim3OptPair*>& pairs, Engine& engine = Engine::Default())
    {
        if (pairs.empty()) return;
        std::lock_guard<std::mutex> lock(engine.mutex());
        const int C = (int)pairs.size();
        std::vector<int32_t> offsets(C + 1, 0), fix(C);
        std::vector<float> x1, x2, o1, o2, is1, is2, K1, K2, S, th;
        std::vector<std::vector<int>> index(C);
        auto to_cam = [](const float* R, const float* t, const float* X, std::vector<float>& out) {
            for (int r = 0; r < 3; ++r) out.push_back(R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2] + t[r]);   // f32, Optimizer.cpp:1114,1122
        };
        for (int c = 0; c < C; ++c) {
            const Sim3OptPair& p = *pairs[c];
            for (int i = 0; i < p.n; ++i) {
                if (!p.valid2[i] || !p.valid1[i] || p.index_in_kf2[i] < 0) continue;      // Optimizer.cpp:1096-1143
                const int i2 = p.index_in_kf2[i];
                index[c].push_back(i);
                to_cam(p.kf1.Rcw, p.kf1.tcw, p.world_pos1 + 3 * i, x1);
                to_cam(p.kf2.Rcw, p.kf2.tcw, p.world_pos2 + 3 * i, x2);
                o1.push_back(p.kf1.keys_xy[2 * i]); o1.push_back(p.kf1.keys_xy[2 * i + 1]);
                o2.push_back(p.kf2.keys_xy[2 * i2]); o2.push_back(p.kf2.keys_xy[2 * i2 + 1]);
                is1.push_back(p.inv_level_sigma2_1[p.kf1.octave[i]]);
                is2.push_back(p.inv_level_sigma2_2[p.kf2.octave[i2]]);
            }
            offsets[c + 1] = offsets[c] + (int32_t)index[c].size();
            const float k1[4] = {p.kf1.fx, p.kf1.fy, p.kf1.cx, p.kf1.cy}, k2[4] = {p.kf2.fx, p.kf2.fy, p.kf2.cx, p.kf2.cy};
            K1.insert(K1.end(), k1, k1 + 4);
            K2.insert(K2.end(), k2, k2 + 4);
            for (int k = 0; k < 9; ++k) S.push_back((float)p.S12.R[k]);
            for (int k = 0; k < 3; ++k) S.push_back((float)p.S12.t[k]);
            S.push_back((float)p.S12.s);
            th.push_back(p.th2);
            fix[c] = p.fix_scale ? 1 : 0;
        }
        rsac_sim3opt_batch b;
        std::memset(&b, 0, sizeof(b));
        b.C = C; b.offsets = offsets.data(); b.x1c = x1.data(); b.x2c = x2.data(); b.obs1 = o1.data(); b.obs2 = o2.data();
        b.inv_sigma2_1 = is1.data(); b.inv_sigma2_2 = is2.data(); b.K1 = K1.data(); b.K2 = K2.data(); b.S12 = S.data();
        b.th2 = th.data(); b.fix_scale = fix.data();
        std::vector<rsac_sim3opt_result> res(C);
        std::vector<uint8_t> removed((size_t)std::max(offsets[C], 1));
        check(rsac_sim3opt_solve(engine.handle(), &b, res.data(), removed.data()), engine.handle(), "rsac_sim3opt_solve");
        for (int c = 0; c < C; ++c) {
            Sim3OptPair& p = *pairs[c];
            p.n_inliers = res[c].n_inliers;
            if (p.match_alive)
                for (size_t k = 0; k < index[c].size(); ++k)
                    if (removed[(size_t)offsets[c] + k]) (*p.match_alive)[(size_t)index[c][k]] = false;
            if (!res[c].optimized) continue;          // `return 0` before g2oS12 is written (Optimizer.cpp:1203-1204)
            for (int k = 0; k < 9; ++k) p.S12.R[k] = res[c].R[k];
            for (int k = 0; k < 3; ++k) p.S12.t[k] = res[c].t[k];
            p.S12.s = res[c].s;
        }
    }
};

}  // namespace ransac_b200
