// dropin_vs_reference.cpp -- the drop-in claim, literally: the SAME ORB-SLAM objects (Frame, KeyFrame, MapPoint) are
// handed to the REFERENCE's own solver classes (src/PnPsolver.cpp, src/Sim3Solver.cpp, compiled unmodified into
// oracle/_ref/libref_solvers.so against the stand-in headers of oracle/shim/) and, through the adapters of
// INTEGRATION.md section 2, to the drop-in classes of include/ransac_b200/solvers.hpp (CUDA engine behind the C ABI);
// both are driven with the same call sequence -- SetRansacParameters, then iterate(5, ...) until bFound or bNoMore, the way
// Tracking::Relocalization (Tracking.cpp:1226-1255) and LoopClosing::ComputeSim3 (LoopClosing.cpp:260-286) do -- and what
// every call returns is compared bit for bit.  tests/test_gpu_dropin.py feeds it seeded frames.  TEST INFRASTRUCTURE.
//
// The drop-in runs with RSAC_FLAG_EPNP_EIGEN (the reference's own 12 x 12 eigen-solve per hypothesis); rand() is
// seeded identically on both sides (the reference: srand through DUtils::Random::SeedRand; the drop-in: SetSeed).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <string>
#include <fstream>
#include <memory>
#include <vector>

#include "PnPsolver.hpp"      // the reference's header (oracle/_ref/inc/ -> /root/reference/include/PnPsolver.hpp)
#include "Sim3Solver.hpp"     // the reference's header
#include "MLPnPsolver.hpp"    // the reference's header
#include "ransac_b200/solvers.hpp"

namespace rb = ransac_b200;
using ORB_SLAM_CUSTOM::Frame;
using ORB_SLAM_CUSTOM::KeyFrame;
using ORB_SLAM_CUSTOM::MapPoint;

// ---- INTEGRATION.md section 2, verbatim (namespace of this fork: ORB_SLAM_CUSTOM) ----
struct FrameSnapshot {                       // owns the flat copies the views point into
    std::vector<float> xy; std::vector<int> oct; std::vector<unsigned char> valid; std::vector<float> wp;
    rb::FrameView F; rb::MapPointMatches M;
    FrameSnapshot(const ORB_SLAM_CUSTOM::Frame& f, const std::vector<std::shared_ptr<ORB_SLAM_CUSTOM::MapPoint>>& mps) {
        const int n = (int)mps.size();                       // PnPsolver.cpp:16-22
        xy.resize(2 * n); oct.resize(n); valid.assign(n, 0); wp.assign(3 * n, 0.f);
        for (int i = 0; i < n; ++i) {
            xy[2*i] = f.mvKeysUn[i].pt.x; xy[2*i+1] = f.mvKeysUn[i].pt.y; oct[i] = f.mvKeysUn[i].octave;
            if (mps[i] && !mps[i]->isBad()) {                // PnPsolver.cpp:24-28
                valid[i] = 1;
                const Eigen::Vector3f p = mps[i]->GetWorldPos();
                wp[3*i] = p(0); wp[3*i+1] = p(1); wp[3*i+2] = p(2);
            }
        }
        F.n_keypoints = n; F.keys_xy = xy.data(); F.octave = oct.data();
        F.level_sigma2 = f.mvLevelSigma2.data(); F.fx = f.fx; F.fy = f.fy; F.cx = f.cx; F.cy = f.cy;
        M.n = n; M.valid = valid.data(); M.world_pos = wp.data();
    }
};
struct KeyFramePairSnapshot {                // Sim3Solver.cpp:9-79
    std::vector<int> oct1, oct2, idx1, idx2; std::vector<unsigned char> v1, v2; std::vector<float> wp1, wp2;
    rb::KeyFrameView K1, K2; rb::Sim3Matches M;
    static void view(const std::shared_ptr<KeyFrame>& kf, std::vector<int>& oct, rb::KeyFrameView& V) {
        const Eigen::Matrix3f R = kf->GetRotation(); const Eigen::Vector3f t = kf->GetTranslation();
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) V.Rcw[3*i+j] = R(i, j); V.tcw[i] = t(i); }
        oct.resize(kf->mvKeysUn.size());
        for (size_t i = 0; i < oct.size(); ++i) oct[i] = kf->mvKeysUn[i].octave;
        V.n_keypoints = (int)oct.size(); V.octave = oct.data(); V.level_sigma2 = kf->mvLevelSigma2.data();
        V.fx = kf->mK(0, 0); V.fy = kf->mK(1, 1); V.cx = kf->mK(0, 2); V.cy = kf->mK(1, 2);
    }
    KeyFramePairSnapshot(std::shared_ptr<KeyFrame> kf1, std::shared_ptr<KeyFrame> kf2, const std::vector<std::shared_ptr<MapPoint>>& m12) {
        view(kf1, oct1, K1); view(kf2, oct2, K2);
        const std::vector<std::shared_ptr<MapPoint>> mp1 = kf1->GetMapPointMatches();
        const int n = (int)m12.size();
        v1.assign(n, 0); v2.assign(n, 0); idx1.assign(n, -1); idx2.assign(n, -1); wp1.assign(3 * n, 0.f); wp2.assign(3 * n, 0.f);
        for (int i = 0; i < n; ++i) {
            if (mp1[i] && !mp1[i]->isBad()) {
                v1[i] = 1; idx1[i] = mp1[i]->GetIndexInKeyFrame(kf1);
                const Eigen::Vector3f p = mp1[i]->GetWorldPos(); wp1[3*i] = p(0); wp1[3*i+1] = p(1); wp1[3*i+2] = p(2);
            }
            if (m12[i] && !m12[i]->isBad()) {
                v2[i] = 1; idx2[i] = m12[i]->GetIndexInKeyFrame(kf2);
                const Eigen::Vector3f p = m12[i]->GetWorldPos(); wp2[3*i] = p(0); wp2[3*i+1] = p(1); wp2[3*i+2] = p(2);
            }
        }
        M.n = n; M.valid1 = v1.data(); M.world_pos1 = wp1.data(); M.index_in_kf1 = idx1.data();
        M.valid2 = v2.data(); M.world_pos2 = wp2.data(); M.index_in_kf2 = idx2.data();
    }
};

template <typename T> static void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(T) * n); }

static int fail(const char* what, int call) { std::printf("{\"equal\":0,\"where\":\"%s\",\"call\":%d}\n", what, call); return 1; }

// MODE 0: PnPsolver (bit for bit); MODE 1: MLPnPsolver (return values, counts and inlier vectors exact; the pose to 1e-6 relative:
// the 6-point solve runs sin / cos / acos of the device's libm, tests/test_gpu_mlpnp.py states the same tolerance)
template <int MODE> static int run_pnp(std::ifstream& f)
{
    int n; float K[4];
    rd(f, &n, 1); rd(f, K, 4);
    std::vector<float> xy(2 * n), world(3 * n), sigma2(8); std::vector<int> octave(n); std::vector<unsigned char> state(n);
    rd(f, xy.data(), xy.size()); rd(f, octave.data(), n); rd(f, state.data(), n); rd(f, world.data(), world.size()); rd(f, sigma2.data(), 8);
    double prob; int minInl, maxIts, minSet; float eps, th2; unsigned seed; int step;
    rd(f, &prob, 1); rd(f, &minInl, 1); rd(f, &maxIts, 1); rd(f, &minSet, 1); rd(f, &eps, 1); rd(f, &th2, 1); rd(f, &seed, 1); rd(f, &step, 1);
    // the ORB-SLAM objects
    Frame F;
    F.fx = K[0]; F.fy = K[1]; F.cx = K[2]; F.cy = K[3];
    F.mvLevelSigma2 = sigma2;
    F.mvKeysUn.resize(n); F.mvpMapPoints.resize(n);
    std::vector<std::shared_ptr<MapPoint>> mps(n);
    for (int i = 0; i < n; ++i) {
        F.mvKeysUn[i].pt = cv::Point2f(xy[2*i], xy[2*i+1]); F.mvKeysUn[i].octave = octave[i];
        if (state[i]) {
            mps[i] = std::make_shared<MapPoint>();
            mps[i]->mWorldPos = Eigen::Vector3f(world[3*i], world[3*i+1], world[3*i+2]);
            mps[i]->mbBad = state[i] == 2;
        }
    }
    FrameSnapshot snap(F, mps);
    // the first iterate(step) consumes the whole budget on both sides (the `||` of PnPsolver.cpp:119 / MLPnPsolver.cpp:71); calls
    // after a Refine() are not compared: there the as-shipped PnPsolver sums stale rows, the drop-in does not
    std::vector<bool> in_r, in_d; int n_r = 0, n_d = 0; bool nm_r = false, nm_d = false, ok_r, ok_d;
    Eigen::Matrix4f T_r; rb::Matrix4f T_d;
    if (MODE == 0) {
        ORB_SLAM_CUSTOM::PnPsolver ref(F, mps);
        ref.SetRansacParameters(prob, minInl, maxIts, minSet, eps, th2);
        rb::PnPsolver dr(snap.F, snap.M);
        dr.SetRansacParameters(prob, minInl, maxIts, minSet, eps, th2);
        dr.SetSeed(seed);
        dr.SetEngineFlags(RSAC_FLAG_EPNP_EIGEN);
        DUtils::Random::SeedRand((int)seed);
        ok_r = ref.iterate(step, nm_r, in_r, n_r, T_r);
        ok_d = dr.iterate(step, nm_d, in_d, n_d, T_d);
    } else {
        ORB_SLAM_CUSTOM::MLPnPsolver ref(F, mps);
        ref.SetRansacParameters(prob, minInl, maxIts, minSet, eps, th2);
        rb::MLPnPsolver dr(snap.F, snap.M);
        dr.SetRansacParameters(prob, minInl, maxIts, minSet, eps, th2);
        dr.SetSeed(seed);
        dr.SetDiscardRefine(true);                          // MLPnPsolver::Refine as shipped never stores its pose
        DUtils::Random::SeedRand((int)seed);
        ok_r = ref.iterate(step, nm_r, in_r, n_r, T_r);
        ok_d = dr.iterate(step, nm_d, in_d, n_d, T_d);
    }
    if (ok_r != ok_d || nm_r != nm_d || n_r != n_d) return fail("return values", 0);
    double worst = 0.0;
    if (ok_r) {
        if (in_r.size() != in_d.size()) return fail("inlier vector length", 0);
        for (size_t i = 0; i < in_r.size(); ++i) if (in_r[i] != in_d[i]) return fail("inlier vector", 0);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) {
            if (MODE == 0) { if (T_r(i, j) != T_d(i, j)) return fail("pose", 0); }
            else {
                const double d = std::fabs((double)T_r(i, j) - (double)T_d(i, j)) / (1.0 + std::fabs((double)T_r(i, j)));
                if (d > worst) worst = d;
            }
        }
        if (worst > 1e-6) return fail("pose tolerance", 0);
    }
    std::printf("{\"equal\":1,\"ok\":%d,\"noMore\":%d,\"nInliers\":%d,\"pose_dev\":%.3g}\n", (int)ok_r, (int)nm_r, n_r, worst);
    return 0;
}

static int run_sim3(std::ifstream& f)
{
    int n; float K[4];
    rd(f, &n, 1); rd(f, K, 4);
    std::vector<float> x1(3 * n), x2(3 * n), sigma2(8); std::vector<int> o1(n), o2(n); std::vector<unsigned char> state(n);
    rd(f, x1.data(), x1.size()); rd(f, x2.data(), x2.size()); rd(f, o1.data(), n); rd(f, o2.data(), n); rd(f, state.data(), n); rd(f, sigma2.data(), 8);
    double prob; int minInl, maxIts; unsigned seed; int step;
    rd(f, &prob, 1); rd(f, &minInl, 1); rd(f, &maxIts, 1); rd(f, &seed, 1); rd(f, &step, 1);
    auto kf1 = std::make_shared<KeyFrame>(), kf2 = std::make_shared<KeyFrame>();
    for (auto* kf : {kf1.get(), kf2.get()}) {
        kf->mRcw.setIdentity(); kf->mtcw.setZero(); kf->mK.setIdentity();
        kf->mK(0, 0) = K[0]; kf->mK(1, 1) = K[1]; kf->mK(0, 2) = K[2]; kf->mK(1, 2) = K[3];
        kf->mvLevelSigma2 = sigma2; kf->mvKeysUn.resize(n); kf->mvpMapPoints.resize(n);
    }
    std::vector<std::shared_ptr<MapPoint>> m12(n);
    for (int i = 0; i < n; ++i) {
        kf1->mvKeysUn[i].octave = o1[i]; kf2->mvKeysUn[i].octave = o2[i];
        auto p1 = std::make_shared<MapPoint>();
        p1->mWorldPos = Eigen::Vector3f(x1[3*i], x1[3*i+1], x1[3*i+2]); p1->mpKF1 = kf1.get(); p1->mIndexKF1 = i;
        kf1->mvpMapPoints[i] = p1;
        if (state[i]) {
            auto p2 = std::make_shared<MapPoint>();
            p2->mWorldPos = Eigen::Vector3f(x2[3*i], x2[3*i+1], x2[3*i+2]); p2->mpKF2 = kf2.get(); p2->mIndexKF2 = i; p2->mbBad = state[i] == 2;
            kf2->mvpMapPoints[i] = p2; m12[i] = p2;
        }
    }
    ORB_SLAM_CUSTOM::Sim3Solver ref(kf1, kf2, m12);
    ref.SetRansacParameters(prob, minInl, maxIts);
    KeyFramePairSnapshot snap(kf1, kf2, m12);
    rb::Sim3Solver dr(snap.K1, snap.K2, snap.M, true);
    dr.SetRansacParameters(prob, minInl, maxIts);
    dr.SetSeed(seed);
    DUtils::Random::SeedRand((int)seed);
    int calls = 0;
    for (; calls < 400; ++calls) {                          // LoopClosing.cpp:275-290: iterate(5, ...) until found or no more
        std::vector<bool> in_r, in_d; int n_r = 0, n_d = 0; bool nm_r = false, nm_d = false;
        const bool ok_r = ref.iterate(step, nm_r, in_r, n_r);
        const bool ok_d = dr.iterate(step, nm_d, in_d, n_d);
        if (ok_r != ok_d || nm_r != nm_d || n_r != n_d) return fail("sim3 return values", calls);
        if (in_r.size() != in_d.size()) return fail("sim3 inlier vector length", calls);
        for (size_t i = 0; i < in_r.size(); ++i) if (in_r[i] != in_d[i]) return fail("sim3 inlier vector", calls);
        if (ok_r) {
            const Eigen::Matrix3f R = ref.GetEstimatedRotation(); const Eigen::Vector3f t = ref.GetEstimatedTranslation();
            const rb::Matrix3f Rd = dr.GetEstimatedRotation(); const rb::Vector3f td = dr.GetEstimatedTranslation();
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) if (R(i, j) != Rd(i, j)) return fail("sim3 rotation", calls); if (t(i) != td(i)) return fail("sim3 translation", calls); }
        }
        if (ok_r || nm_r) { std::printf("{\"equal\":1,\"ok\":%d,\"noMore\":%d,\"nInliers\":%d,\"calls\":%d}\n", (int)ok_r, (int)nm_r, n_r, calls + 1); return 0; }
    }
    return fail("sim3 did not terminate", calls);
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: dropin_vs_reference pnp|mlpnp|sim3 file\n"); return 2; }
    std::ifstream f(argv[2], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
    try {
        const std::string mode(argv[1]);
        return mode == "pnp" ? run_pnp<0>(f) : mode == "mlpnp" ? run_pnp<1>(f) : run_sim3(f);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
