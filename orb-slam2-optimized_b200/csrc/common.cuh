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

// Checked build (-DRSAC_CHECKED; scripts/run_checked.sh): compute-sanitizer is closed on this GPU pool, so the hand-written
// mbarrier ring, the list-driven launches and the index arithmetic carry their own assertions -- a violated one prints its
// location and traps (the CUDA error surfaces through the C ABI and fails the test), and the scoring ring is poison-filled
// before every bulk copy so that a short or misplaced copy cannot go unnoticed.  The default build compiles all of it away.
#if defined(RSAC_CHECKED) && defined(__CUDA_ARCH__)
#include <cstdio>
#define RSAC_ASSERT(cond)                                                                                        \
    do {                                                                                                         \
        if (!(cond)) {                                                                                           \
            printf("RSAC_CHECKED: %s failed at %s:%d (block %d, thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                                            \
        }                                                                                                        \
    } while (0)
#else
#define RSAC_ASSERT(cond) ((void)0)
#endif

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
