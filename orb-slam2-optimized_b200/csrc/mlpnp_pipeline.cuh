// mlpnp_pipeline.cuh -- MLPnPsolver::iterate's minimal solves as a batched kernel (MLPnPsolver.cpp:76-120).
#pragma once
#include "common.cuh"
#include "mlpnp.cuh"

namespace rsac {

// ---- MLPnP minimal solve: one thread per hypothesis (MLPnPsolver.cpp:76-120) ----
static __global__ void __launch_bounds__(128) mlpnp_minimal_kernel(const ProblemMeta* metas, int C, int64_t sumH,
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
