// pnp_pipeline.cuh -- PnPsolver::iterate as a batched device pipeline.
//
// Reference control flow: src/PnPsolver.cpp:102-191 (iterate; the `||` at :119 makes the
// first call consume the whole budget), :193-238 (Refine).  All H hypotheses of all
// problems are solved and scored in parallel; the sequential semantics (first strict
// maximum, refine on the prefix-best set, first refine with > minInliers wins, otherwise
// the unrefined best) are then replayed per problem.
#pragma once
#include "common.cuh"
#include "epnp.cuh"
#include "rng.cuh"
#include "score.cuh"

namespace rsac {

// ---- EPnP minimal solve: one thread per hypothesis (PnPsolver.cpp:125-141) ----
#ifndef RSAC_SOLVE_SMEM
#define RSAC_SOLVE_SMEM 0    // what lives in shared memory: 0 the basis U4 (48 doubles/thread), 1 L (60), 2 both (108)
#endif
constexpr int kSolveSmemDoubles = RSAC_SOLVE_SMEM == 0 ? 48 : (RSAC_SOLVE_SMEM == 1 ? 60 : 108);
#ifndef RSAC_SOLVE_THREADS
#define RSAC_SOLVE_THREADS 128
#endif
#ifndef RSAC_SOLVE_BLOCKS
#define RSAC_SOLVE_BLOCKS 3   // with the basis in shared memory: 168 registers x 384 threads/SM measured best (blocks/SM 2: 0.99 ms, 3: 0.83, 4: 0.96 per 307k solves)
#endif
// QR = true: null space of the 4-point system by Householder QR (default); false: 12x12 eigen-solve
// (RSAC_FLAG_EPNP_EIGEN, the reference's structure)
template <bool QR>
__global__ void __launch_bounds__(QR ? RSAC_SOLVE_THREADS : 128, QR ? RSAC_SOLVE_BLOCKS : 2) epnp_minimal_kernel(const ProblemMeta* metas, int C, int64_t sumH,
                                                           const uint32_t* tables, const float4* cA,
                                                           const float4* cC, float* poses)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sumH) return;
    const int p = find_problem(metas, C, g);
    const ProblemMeta& m = metas[p];
    const int h = (int)(g - m.hyp_off);
    const uint32_t* idx = tables + m.table_off + (size_t)h * 4;
    double pw[12], us[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t ci = (size_t)m.corr_off + idx[i];
        const float4 a = cA[ci];
        const float4 q = cC[ci];
        pw[3 * i] = (double)a.x; pw[3 * i + 1] = (double)a.y; pw[3 * i + 2] = (double)a.z;   // add_correspondence (:288-294)
        us[2 * i] = (double)q.x; us[2 * i + 1] = (double)q.y;
    }
    const Cam k = {m.fx, m.fy, m.cx, m.cy};
    float R[9], t[3];
    if constexpr (QR) {
        extern __shared__ double s_cols[];               // per thread one column of doubles: basis U4 [48] (and L [60])
#if RSAC_SOLVE_SMEM == 0
        epnp_compute_pose_small<4, true>(pw, us, k, R, t, s_cols + threadIdx.x, (int)blockDim.x);
#elif RSAC_SOLVE_SMEM == 1
        epnp_compute_pose_small<4, true>(pw, us, k, R, t, nullptr, 1, s_cols + threadIdx.x, (int)blockDim.x);
#else
        epnp_compute_pose_small<4, true>(pw, us, k, R, t, s_cols + threadIdx.x, (int)blockDim.x,
                                         s_cols + 48 * blockDim.x + threadIdx.x, (int)blockDim.x);
#endif
    } else {
        epnp_compute_pose_small<4, false>(pw, us, k, R, t);
    }
    float* out = poses + g * 12;
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    out[9] = t[0]; out[10] = t[1]; out[11] = t[2];
}

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
__global__ void __launch_bounds__(128) early_exit_flag_kernel(const ProblemMeta* metas, int C, const int32_t* counts, int stage, int lim,
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
            if (!any && lane == 0) v.list(stage + 1)[atomicAdd(v.counters + stage + 1, 1)] = p;
        }
        if (lane == 0) v.upto[p] = min(m.H, lim);
    } else {
        if (lane == 0 && v.upto[p] < m.H) {
            v.upto[p] = m.H;
            v.listC[atomicAdd(v.counters + kCleanupCounter, 1)] = p;   // the counter was reset by the host
        }
    }
}

// EPnP minimal solves of hypotheses [h_lo, h_lo + span) of the listed problems (list == nullptr: all C problems).
// Persistent grid-stride form: the amount of work is only known on the device.
// QR = false: the 12x12 eigen-solve of M^T M per hypothesis (RSAC_FLAG_EPNP_EIGEN, the reference's structure)
template <bool QR>
__global__ void __launch_bounds__(QR ? RSAC_SOLVE_THREADS : 128, QR ? RSAC_SOLVE_BLOCKS : 2)
epnp_minimal_range_kernel(const ProblemMeta* metas, int C, const int32_t* list, const int32_t* list_count, int h_lo, int span,
                          const uint32_t* tables, const float4* cA, const float4* cC, float* poses)
{
    extern __shared__ double s_cols[];
    const int np = list ? *list_count : C;
    const int64_t total = (int64_t)np * span;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t / span);
        const int h = h_lo + (int)(t - (int64_t)k * span);
        const int p = list ? list[k] : k;
        const ProblemMeta& m = metas[p];
        if (h >= m.H) continue;
        const uint32_t* idx = tables + m.table_off + (size_t)h * 4;
        double pw[12], us[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t ci = (size_t)m.corr_off + idx[i];
            const float4 a = cA[ci];
            const float4 q = cC[ci];
            pw[3 * i] = (double)a.x; pw[3 * i + 1] = (double)a.y; pw[3 * i + 2] = (double)a.z;
            us[2 * i] = (double)q.x; us[2 * i + 1] = (double)q.y;
        }
        const Cam kk = {m.fx, m.fy, m.cx, m.cy};
        float R[9], tr[3];
        if constexpr (!QR) {
            epnp_compute_pose_small<4, false>(pw, us, kk, R, tr);
        } else {
#if RSAC_SOLVE_SMEM == 0
        epnp_compute_pose_small<4, true>(pw, us, kk, R, tr, s_cols + threadIdx.x, (int)blockDim.x);
#elif RSAC_SOLVE_SMEM == 1
        epnp_compute_pose_small<4, true>(pw, us, kk, R, tr, nullptr, 1, s_cols + threadIdx.x, (int)blockDim.x);
#else
        epnp_compute_pose_small<4, true>(pw, us, kk, R, tr, s_cols + threadIdx.x, (int)blockDim.x,
                                         s_cols + 48 * blockDim.x + threadIdx.x, (int)blockDim.x);
#endif
        }
        float* out = poses + ((int64_t)m.hyp_off + h) * 12;
#pragma unroll
        for (int i = 0; i < 9; ++i) out[i] = R[i];
        out[9] = tr[0]; out[10] = tr[1]; out[11] = tr[2];
    }
}

}  // namespace rsac
