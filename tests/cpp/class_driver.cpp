// class_driver.cpp -- exercises the C++ class API (include/ransac_b200/solvers.hpp) the way
// Tracking::Relocalization / LoopClosing::ComputeSim3 drive the reference's solvers, and prints
// what every call returned as JSON lines.  tests/test_gpu_classes.py feeds it seeded problems and
// checks the output against the CPU oracle.  (The oracle is NOT linked here.)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "ransac_b200/solvers.hpp"

using namespace ransac_b200;

template <typename T> static void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(T) * n); }

static void print_inliers(const std::vector<bool>& v)
{
    std::printf("[");
    bool first = true;
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i]) { std::printf(first ? "%zu" : ",%zu", i); first = false; }
    std::printf("]");
}

static void print_T(const Matrix4f& T)
{
    std::printf("[");
    for (int i = 0; i < 16; ++i) std::printf(i ? ",%.9g" : "%.9g", T.m[i]);
    std::printf("]");
}

struct FrameData {
    int n = 0;
    float K[4];
    std::vector<float> xy, world, sigma2;
    std::vector<int> octave;
    std::vector<unsigned char> valid;
    FrameView frame() const
    {
        FrameView F;
        F.n_keypoints = n; F.keys_xy = xy.data(); F.octave = octave.data(); F.level_sigma2 = sigma2.data();
        F.fx = K[0]; F.fy = K[1]; F.cx = K[2]; F.cy = K[3];
        return F;
    }
    MapPointMatches matches() const
    {
        MapPointMatches M;
        M.n = n; M.valid = valid.data(); M.world_pos = world.data();
        return M;
    }
};

static FrameData read_frame(std::ifstream& f)
{
    FrameData d;
    rd(f, &d.n, 1);
    rd(f, d.K, 4);
    d.xy.resize(2 * d.n); d.world.resize(3 * d.n); d.octave.resize(d.n); d.valid.resize(d.n); d.sigma2.resize(8);
    rd(f, d.xy.data(), d.xy.size());
    rd(f, d.octave.data(), d.octave.size());
    rd(f, d.valid.data(), d.valid.size());
    rd(f, d.world.data(), d.world.size());
    rd(f, d.sigma2.data(), 8);
    return d;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: class_driver pnp|pnp_batch|mlpnp|sim3|poseopt|sim3opt file\n"); return 2; }
    const std::string mode = argv[1];
    std::ifstream f(argv[2], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
    try {
        if (mode == "pnp" || mode == "mlpnp") {
            FrameData d = read_frame(f);
            double prob; int minInl, maxIts, minSet; float eps, th2; unsigned seed; int step;
            rd(f, &prob, 1); rd(f, &minInl, 1); rd(f, &maxIts, 1); rd(f, &minSet, 1); rd(f, &eps, 1); rd(f, &th2, 1); rd(f, &seed, 1); rd(f, &step, 1);
            std::vector<bool> inl; int n; bool noMore; Matrix4f T;
            if (mode == "pnp") {
                PnPsolver s(d.frame(), d.matches());
                s.SetRansacParameters(prob, minInl, maxIts, minSet, eps, th2);
                s.SetSeed(seed);
                std::printf("{\"H\":%d,\"minInl\":%d,\"N\":%d}\n", s.GetIterations(), s.GetMinInliers(), s.GetNumCorrespondences());
                for (int call = 0; call < 400; ++call) {     // Tracking.cpp:1255 keeps calling iterate(5,...)
                    const bool ok = s.iterate(step, noMore, inl, n, T);
                    std::printf("{\"call\":%d,\"ok\":%d,\"noMore\":%d,\"nInliers\":%d,\"T\":", call, (int)ok, (int)noMore, n);
                    print_T(T);
                    std::printf(",\"inliers\":");
                    print_inliers(inl);
                    std::printf("}\n");
                    if (noMore) break;
                }
                {   // one call past the budget: the best-so-far fallback with bNoMore (PnPsolver.cpp:119,173-188)
                    const bool ok = s.iterate(step, noMore, inl, n, T);
                    std::printf("{\"call\":-1,\"ok\":%d,\"noMore\":%d,\"nInliers\":%d,\"T\":", (int)ok, (int)noMore, n);
                    print_T(T);
                    std::printf(",\"inliers\":");
                    print_inliers(inl);
                    std::printf("}\n");
                }
            } else {
                MLPnPsolver s(d.frame(), d.matches());
                s.SetRansacParameters(prob, minInl, maxIts, minSet, eps, th2);
                s.SetSeed(seed);
                for (int call = 0; call < 400; ++call) {
                    const bool ok = s.iterate(step, noMore, inl, n, T);
                    std::printf("{\"call\":%d,\"ok\":%d,\"noMore\":%d,\"nInliers\":%d,\"T\":", call, (int)ok, (int)noMore, n);
                    print_T(T);
                    std::printf(",\"inliers\":");
                    print_inliers(inl);
                    std::printf("}\n");
                    if (noMore) break;
                }
            }
        } else if (mode == "pnp_batch") {
            int C;
            rd(f, &C, 1);
            std::vector<FrameData> ds;
            std::vector<unsigned> seeds(C);
            for (int c = 0; c < C; ++c) { ds.push_back(read_frame(f)); rd(f, &seeds[c], 1); }
            std::vector<std::unique_ptr<PnPsolver>> solvers;
            std::vector<PnPsolver*> ptrs;
            for (int c = 0; c < C; ++c) {
                solvers.emplace_back(new PnPsolver(ds[c].frame(), ds[c].matches()));
                solvers.back()->SetRansacParameters(0.99, 10, 300, 4, 0.2f, 5.991f);
                solvers.back()->SetSeed(seeds[c]);
                ptrs.push_back(solvers.back().get());
            }
            PnPsolver::SolveBatch(ptrs);          // one device pass for all candidates
            for (int c = 0; c < C; ++c) {
                std::vector<bool> inl; int n; Matrix4f T;
                const bool ok = solvers[c]->find(inl, n, T);
                std::printf("{\"cand\":%d,\"ok\":%d,\"nInliers\":%d,\"T\":", c, (int)ok, n);
                print_T(T);
                std::printf(",\"inliers\":");
                print_inliers(inl);
                std::printf("}\n");
            }
        } else if (mode == "sim3") {
            int n, fix; unsigned seed; int minInl, maxIts, step; double prob;
            KeyFrameView K1, K2;
            rd(f, &n, 1); rd(f, &fix, 1); rd(f, &seed, 1); rd(f, &prob, 1); rd(f, &minInl, 1); rd(f, &maxIts, 1); rd(f, &step, 1);
            float Kc[4];
            rd(f, K1.Rcw, 9); rd(f, K1.tcw, 3); rd(f, K2.Rcw, 9); rd(f, K2.tcw, 3); rd(f, Kc, 4);
            std::vector<float> w1(3 * n), w2(3 * n), sig(8);
            std::vector<int> o1(n), o2(n), i1(n), i2(n);
            std::vector<unsigned char> v1(n), v2(n);
            rd(f, w1.data(), w1.size()); rd(f, w2.data(), w2.size()); rd(f, o1.data(), n); rd(f, o2.data(), n);
            rd(f, i1.data(), n); rd(f, i2.data(), n); rd(f, v1.data(), n); rd(f, v2.data(), n); rd(f, sig.data(), 8);
            K1.n_keypoints = K2.n_keypoints = n;
            K1.octave = o1.data(); K2.octave = o2.data(); K1.level_sigma2 = K2.level_sigma2 = sig.data();
            K1.fx = K2.fx = Kc[0]; K1.fy = K2.fy = Kc[1]; K1.cx = K2.cx = Kc[2]; K1.cy = K2.cy = Kc[3];
            Sim3Matches M;
            M.n = n; M.valid1 = v1.data(); M.valid2 = v2.data(); M.world_pos1 = w1.data(); M.world_pos2 = w2.data();
            M.index_in_kf1 = i1.data(); M.index_in_kf2 = i2.data();
            Sim3Solver s(K1, K2, M, fix != 0);
            s.SetRansacParameters(prob, minInl, maxIts);
            s.SetSeed(seed);
            std::vector<bool> inl; int nin; bool noMore;
            for (int call = 0; call < 400; ++call) {          // LoopClosing.cpp:286 iterate(5,...)
                const bool ok = s.iterate(step, noMore, inl, nin);
                const Matrix3f R = s.GetEstimatedRotation();
                const Vector3f t = s.GetEstimatedTranslation();
                std::printf("{\"call\":%d,\"ok\":%d,\"noMore\":%d,\"nInliers\":%d,\"s\":%.9g,\"R\":[", call, (int)ok, (int)noMore, nin, s.GetEstimatedScale());
                for (int i = 0; i < 9; ++i) std::printf(i ? ",%.9g" : "%.9g", R.m[i]);
                std::printf("],\"t\":[%.9g,%.9g,%.9g],\"inliers\":", t.v[0], t.v[1], t.v[2]);
                print_inliers(inl);
                std::printf("}\n");
                if (ok || noMore) break;
            }
        }
        else if (mode == "poseopt") {
            // Optimizer::PoseOptimization(&frame) per candidate, as Tracking.cpp:1284 does, in one batch
            int C;
            rd(f, &C, 1);
            struct Data { int n; float K[5]; float T[12]; std::vector<float> xy, ur, world, isig; std::vector<int> oct; std::vector<unsigned char> has; std::vector<bool> outl; };
            std::vector<Data> ds(C);
            std::vector<PoseOptFrame> frames(C);
            std::vector<PoseOptFrame*> ptrs;
            for (int c = 0; c < C; ++c) {
                Data& d = ds[c];
                rd(f, &d.n, 1); rd(f, d.K, 5); rd(f, d.T, 12);
                d.xy.resize(2 * d.n); d.ur.resize(d.n); d.world.resize(3 * d.n); d.isig.resize(8); d.oct.resize(d.n); d.has.resize(d.n);
                rd(f, d.xy.data(), d.xy.size()); rd(f, d.ur.data(), d.n); rd(f, d.oct.data(), d.n); rd(f, d.has.data(), d.n);
                rd(f, d.world.data(), d.world.size()); rd(f, d.isig.data(), 8);
                d.outl.assign(d.n, true);             // stale flags: only keypoints with a MapPoint are rewritten
                PoseOptFrame& F = frames[c];
                F.n_keypoints = d.n; F.keys_xy = d.xy.data(); F.octave = d.oct.data(); F.u_right = d.ur.data();
                F.inv_level_sigma2 = d.isig.data(); F.has_map_point = d.has.data(); F.world_pos = d.world.data();
                F.fx = d.K[0]; F.fy = d.K[1]; F.cx = d.K[2]; F.cy = d.K[3]; F.bf = d.K[4];
                for (int r = 0; r < 3; ++r) { for (int q = 0; q < 3; ++q) F.Tcw(r, q) = d.T[3 * r + q]; F.Tcw(r, 3) = d.T[9 + r]; }
                F.outlier = &d.outl;
                ptrs.push_back(&F);
            }
            Optimizer::PoseOptimizationBatch(ptrs);
            const int single = Optimizer::PoseOptimization(ptrs[0]) ;   // the per-frame form, from the optimised pose
            for (int c = 0; c < C; ++c) {
                std::printf("{\"cand\":%d,\"nInliers\":%d,\"T\":", c, c == 0 ? -1 : frames[c].n_inliers);
                print_T(frames[c].Tcw);
                std::printf(",\"outliers\":");
                print_inliers(ds[c].outl);
                std::printf("}\n");
            }
            std::printf("{\"single\":%d}\n", single);
        }
        else if (mode == "sim3opt") {
            // Optimizer::OptimizeSim3 for every loop candidate (LoopClosing.cpp:311), one batch
            int C;
            rd(f, &C, 1);
            struct Data { int n, nk2; KeyFrameView k1, k2; std::vector<float> xy1, xy2, w1, w2, isig; std::vector<int> o1, o2, i2;
                          std::vector<unsigned char> v1, v2; std::vector<bool> alive; float S[13]; float K[4]; };
            std::vector<Data> ds(C);
            std::vector<Sim3OptPair> pairs(C);
            std::vector<Sim3OptPair*> ptrs;
            for (int c = 0; c < C; ++c) {
                Data& d = ds[c];
                rd(f, &d.n, 1); rd(f, &d.nk2, 1);
                rd(f, d.k1.Rcw, 9); rd(f, d.k1.tcw, 3); rd(f, d.k2.Rcw, 9); rd(f, d.k2.tcw, 3); rd(f, d.K, 4); rd(f, d.S, 13);
                d.xy1.resize(2 * d.n); d.xy2.resize(2 * d.nk2); d.w1.resize(3 * d.n); d.w2.resize(3 * d.n); d.isig.resize(8);
                d.o1.resize(d.n); d.o2.resize(d.nk2); d.i2.resize(d.n); d.v1.resize(d.n); d.v2.resize(d.n);
                rd(f, d.xy1.data(), d.xy1.size()); rd(f, d.xy2.data(), d.xy2.size()); rd(f, d.w1.data(), d.w1.size()); rd(f, d.w2.data(), d.w2.size());
                rd(f, d.o1.data(), d.n); rd(f, d.o2.data(), d.nk2); rd(f, d.i2.data(), d.n); rd(f, d.v1.data(), d.n); rd(f, d.v2.data(), d.n);
                rd(f, d.isig.data(), 8);
                d.alive.assign(d.n, true);
                Sim3OptPair& P = pairs[c];
                d.k1.keys_xy = d.xy1.data(); d.k1.octave = d.o1.data(); d.k1.n_keypoints = d.n;
                d.k2.keys_xy = d.xy2.data(); d.k2.octave = d.o2.data(); d.k2.n_keypoints = d.nk2;
                d.k1.fx = d.k2.fx = d.K[0]; d.k1.fy = d.k2.fy = d.K[1]; d.k1.cx = d.k2.cx = d.K[2]; d.k1.cy = d.k2.cy = d.K[3];
                P.kf1 = d.k1; P.kf2 = d.k2; P.inv_level_sigma2_1 = P.inv_level_sigma2_2 = d.isig.data();
                P.n = d.n; P.valid1 = d.v1.data(); P.valid2 = d.v2.data(); P.world_pos1 = d.w1.data(); P.world_pos2 = d.w2.data();
                P.index_in_kf2 = d.i2.data(); P.th2 = 10.0f;
                for (int k = 0; k < 9; ++k) P.S12.R[k] = d.S[k];
                for (int k = 0; k < 3; ++k) P.S12.t[k] = d.S[9 + k];
                P.S12.s = d.S[12];
                P.match_alive = &d.alive;
                ptrs.push_back(&P);
            }
            Sim3Optimizer::OptimizeSim3Batch(ptrs);
            for (int c = 0; c < C; ++c) {
                std::printf("{\"cand\":%d,\"nIn\":%d,\"s\":%.17g,\"R\":[", c, pairs[c].n_inliers, pairs[c].S12.s);
                for (int i = 0; i < 9; ++i) std::printf(i ? ",%.17g" : "%.17g", pairs[c].S12.R[i]);
                std::printf("],\"t\":[%.17g,%.17g,%.17g],\"alive\":", pairs[c].S12.t[0], pairs[c].S12.t[1], pairs[c].S12.t[2]);
                print_inliers(ds[c].alive);
                std::printf("}\n");
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
