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
#include "mlpnp.cuh"

namespace rsac {

// ---- minimal-set tables from per-problem seeds: one thread per problem ----
__global__ void rng_tables_kernel(const ProblemMeta* metas, int C, uint32_t* tables)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= C) return;
    const ProblemMeta& m = metas[p];
    if (m.n < m.min_set) return;
    GlibcRand g;
    g.seed(m.seed);
    uint32_t* out = tables + m.table_off;
    for (int h = 0; h < m.H; ++h) draw_minimal_set<8>(g, m.n, m.min_set, out + (size_t)h * m.min_set);
}

__device__ __forceinline__ int find_problem(const ProblemMeta* metas, int C, int64_t g)
{
    int lo = 0, hi = C - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)metas[mid].hyp_off <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- EPnP minimal solve: one thread per hypothesis (PnPsolver.cpp:125-141) ----
__global__ void __launch_bounds__(128) epnp_minimal_kernel(const ProblemMeta* metas, int C, int64_t sumH,
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
    epnp_compute_pose_small<4>(pw, us, k, R, t);
    float* out = poses + g * 12;
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    out[9] = t[0]; out[10] = t[1]; out[11] = t[2];
}

// ---- MLPnP minimal solve: one thread per hypothesis (MLPnPsolver.cpp:76-120) ----
__global__ void __launch_bounds__(128) mlpnp_minimal_kernel(const ProblemMeta* metas, int C, int64_t sumH,
                                                            const uint32_t* tables, const float4* cA,
                                                            const float4* cC, const double* cov, double* poses)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sumH) return;
    const int p = find_problem(metas, C, g);
    const ProblemMeta& m = metas[p];
    const int h = (int)(g - m.hyp_off);
    const uint32_t* idx = tables + m.table_off + (size_t)h * 6;
    double f[18], pw[18], cv[54];
    for (int i = 0; i < 6; ++i) {
        const size_t ci = (size_t)m.corr_off + idx[i];
        const float4 a = cA[ci];
        const float4 q = cC[ci];
        mlpnp_bearing(q.x, q.y, m.k1, f + 3 * i);                         // MLPnPsolver.cpp:33-37
        pw[3 * i] = (double)a.x; pw[3 * i + 1] = (double)a.y; pw[3 * i + 2] = (double)a.z;
        if (cov)
            for (int k = 0; k < 9; ++k) cv[9 * i + k] = cov[9 * ci + k];
    }
    double R[9], t[3];
    double2 rec[kMaxSweepsRec * 66];
    mlpnp_compute_pose_small<6>(f, pw, cov ? cv : nullptr, R, t, rec);
    double* out = poses + g * 12;
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    out[9] = t[0]; out[10] = t[1]; out[11] = t[2];
}

}  // namespace rsac
