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
#include "early_exit.cuh"

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
