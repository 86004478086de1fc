// common.cuh -- device-side records shared by the kernels of the RANSAC engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rsac {

// One problem (= one candidate keyframe / loop candidate) of a batch.
struct ProblemMeta {
    int32_t corr_off;   // first correspondence in the concatenated arrays
    int32_t n;          // correspondences
    int32_t hyp_off;    // first hypothesis in the concatenated per-hypothesis arrays
    int32_t H;          // hypotheses (= mRansacMaxIts)
    int32_t min_inl;    // adjusted mRansacMinInliers
    int32_t min_set;    // 4 (EPnP) / 6 (MLPnP) / 3 (Sim3)
    int32_t word_off;   // first word of this problem in the final-mask array
    int32_t words;      // ceil(n/32)
    int64_t hmask_off;  // first word of this problem in the per-hypothesis mask array
    int64_t table_off;  // first entry of this problem's minimal-set table
    uint32_t seed;
    int32_t fix_scale;  // Sim3
    double fx, fy, cx, cy;       // PnP: double intrinsics (PnPsolver.hpp:71)
    float k1[4], k2[4];          // MLPnP: k1 = float intrinsics; Sim3: both cameras
};

constexpr float kUnitRoundoff = 5.9604644775390625e-08f;   // 2^-24

#ifdef __CUDACC__
// problem that owns global hypothesis g (binary search over hyp_off)
__device__ __forceinline__ int find_problem(const ProblemMeta* metas, int C, int64_t g)
{
    int lo = 0, hi = C - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)metas[mid].hyp_off <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}
#endif

}  // namespace rsac
