// mlpnp_pipeline.cuh -- MLPnPsolver::iterate's minimal solves as a batched kernel (MLPnPsolver.cpp:76-120).
#pragma once
#include "common.cuh"
#include "mlpnp.cuh"

namespace rsac {

// diagnostic: clock64() stamps of the 6-point solve's phases (hypothesis 0 of the last exhaustive launch; rsac_debug_mlpnp_clocks)
static __device__ long long g_mlpnp_clocks[8];

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
    mlpnp_compute_pose_small<6>(f, pw, cov ? cv : nullptr, R, t, rec, g == 0 ? g_mlpnp_clocks : nullptr);
    double* out = poses + g * 12;
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    out[9] = t[0]; out[10] = t[1]; out[11] = t[2];
}

// The same solves for hypotheses [h_lo, h_lo + span) of the listed problems (list == nullptr: all C problems): the staged
// early exit (MLPnPsolver::iterate returns at the first successful Refine, MLPnPsolver.cpp:144-160).  Grid-stride: the
// amount of work is only known on the device.
static __global__ void __launch_bounds__(128) mlpnp_minimal_range_kernel(const ProblemMeta* metas, int C, const int32_t* list,
                                                                  const int32_t* list_count, int h_lo, int span,
                                                                  const uint32_t* tables, const float4* cA, const float4* cC,
                                                                  const double* cov, double* poses)
{
    const int np = list ? *list_count : C;
    const int64_t total = (int64_t)np * span;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t / span);
        const int h = h_lo + (int)(t - (int64_t)k * span);
        const int p = list ? list[k] : k;
        const ProblemMeta& m = metas[p];
        if (h >= m.H) continue;
        const uint32_t* idx = tables + m.table_off + (size_t)h * 6;
        double f[18], pw[18], cv[54];
        for (int i = 0; i < 6; ++i) {
            const size_t ci = (size_t)m.corr_off + idx[i];
            const float4 a = cA[ci];
            const float4 q = cC[ci];
            mlpnp_bearing(q.x, q.y, m.k1, f + 3 * i);
            pw[3 * i] = (double)a.x; pw[3 * i + 1] = (double)a.y; pw[3 * i + 2] = (double)a.z;
            if (cov)
                for (int c = 0; c < 9; ++c) cv[9 * i + c] = cov[9 * ci + c];
        }
        double R[9], tr[3];
        double2 rec[kMaxSweepsRec * 66];
        mlpnp_compute_pose_small<6>(f, pw, cov ? cv : nullptr, R, tr, rec);
        double* out = poses + ((int64_t)m.hyp_off + h) * 12;
        for (int i = 0; i < 9; ++i) out[i] = R[i];
        out[9] = tr[0]; out[10] = tr[1]; out[11] = tr[2];
    }
}

}  // namespace rsac
