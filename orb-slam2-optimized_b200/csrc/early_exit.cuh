// early_exit.cuh -- device-side state of the staged early exit (RSAC_FLAG_EARLY_EXIT), shared by the PnP and MLPnP pipelines:
// both reference solvers return from iterate() at the first hypothesis whose Refine() succeeds (PnPsolver.cpp:225-236,
// MLPnPsolver.cpp:144-160), so the hypotheses behind it never influence the result.
#pragma once
#include "common.cuh"

namespace rsac {

// ---- early exit in phases (RSAC_FLAG_EARLY_EXIT) ----
// The sequential reference stops at the first hypothesis whose Refine() succeeds (PnPsolver.cpp:225-236), on
// average after 35-40 of cfg4's 300; hypotheses behind it never influence the result.  The batched engine keeps
// that property without a host round trip:
//   stage 0   every problem: hypotheses [0, b0)          (b0 sized so that C * b0 is about one wave of the solver)
//   flag      problems WITHOUT a hypothesis of cnt >= minInliers in [0, b0) go on (list 1); the others are
//             predicted to finish inside what they have
//   stage j   list j: hypotheses [b(j-1), bj);  flag: those still without an acceptable hypothesis -> list j+1
//             (j = 1 .. K-1, bK-1 = H; up to kMaxStages stages, the lists alternate between two buffers)
//   replay    all problems, each over the upto[p] hypotheses it has; a problem whose refines all failed before
//             upto[p] < H is not decided yet: it is appended to list C instead of reporting "budget exhausted"
//   clean-up  list C: hypotheses [b0, H) (again where they exist: the scoring kernel stores whole-group counts, so
//             recomputation is idempotent), then the replay resumes those problems where they stopped (normally
//             empty: the kernels find a zero count and return)
// ee[] layout: [upto: C][listX: C][listY: C][listC: C][counters: 16]   counters[j] = |list j|, counters[15] = |list C|
// upto[p] = hypotheses of p computed so far; the replay stores -(stop + 1) there for a problem it hands to the clean-up
constexpr int kMaxStages = 8;
constexpr int kCleanupCounter = 15;
struct EarlyExit {
    int32_t* upto;
    int32_t* listX;
    int32_t* listY;
    int32_t* listC;
    int32_t* counters;
    __host__ __device__ int32_t* list(int stage) const { return (stage & 1) ? listX : listY; }
};
__host__ __device__ inline EarlyExit early_exit_view(int32_t* ee, int C)
{
    EarlyExit v;
    v.upto = ee; v.listX = ee + C; v.listY = ee + 2 * (size_t)C; v.listC = ee + 3 * (size_t)C; v.counters = ee + 4 * (size_t)C;
    return v;
}
constexpr size_t early_exit_words(int C) { return 4 * (size_t)C + 16; }

// one warp per problem (stage 0, complete) or per list entry (stage >= 1).
//   mode 0, after stage `stage` (hypotheses up to `lim` exist for its members): upto = min(H, lim); members with
//           H > lim and nothing acceptable in [0, lim) -> list stage+1
//   mode 2 (before a later iterate() call resumes): every problem with upto < H -> list C, upto = H
static __global__ void __launch_bounds__(128) early_exit_flag_kernel(const ProblemMeta* metas, int C, const int32_t* counts, int stage, int lim,
                                                              int32_t* ee, int mode)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const EarlyExit v = early_exit_view(ee, C);
    int p = w;
    if (mode == 0 && stage >= 1) { if (w >= v.counters[stage]) return; p = v.list(stage)[w]; }
    else if (w >= C) return;
    const ProblemMeta& m = metas[p];
    if (mode == 0) {
        if (m.H > lim) {
            bool any = false;
            for (int h = lane; h < lim; h += 32) any = any || (counts[m.hyp_off + h] >= m.min_inl);
            any = __any_sync(0xffffffffu, any);
            if (!any && lane == 0) {
                const int slot = atomicAdd(v.counters + stage + 1, 1);
                RSAC_ASSERT(slot >= 0 && slot < C && stage + 1 < kMaxStages && p >= 0 && p < C);
                v.list(stage + 1)[slot] = p;
            }
        }
        if (lane == 0) v.upto[p] = min(m.H, lim);
    } else {
        if (lane == 0 && v.upto[p] < m.H) {
            v.upto[p] = m.H;
            v.listC[atomicAdd(v.counters + kCleanupCounter, 1)] = p;   // the counter was reset by the host
        }
    }
}

}  // namespace rsac
