// score.cuh -- CheckInliers for every hypothesis x correspondence pair.
//
// Reference semantics (bit-exact target):
//   PnPsolver::CheckInliers   src/PnPsolver.cpp:241-268   (f32 transform, f32 1/z, f64 projection)
//   MLPnPsolver::CheckInliers src/MLPnPsolver.cpp:222-255 (f64 transform narrowed to f32, two f32 divisions)
//
// Design (DESIGN.md "scoring kernel"):
//   * lane <-> hypothesis: each lane keeps HPL poses in registers as 3x4 projective rows
//     with the intrinsics folded in and the whole matrix scaled by 1/(|R|_1 + |t|_inf);
//   * correspondences are staged once per CTA in shared memory by a 1-D TMA bulk copy and
//     read back as warp-broadcast LDS.128: two loads serve 32*HPL evaluations;
//   * two-tier evaluation: a 14-FMA-pipe-instruction fast path (FFMA + MUFU.RCP) decides
//     every pair whose squared error is further from the threshold than a rigorous bound on
//     the fast-vs-reference rounding difference; the rare remainder (|e-thr| <= band) is
//     re-evaluated with the reference's exact operation sequence.  Result: the reference's
//     inlier bit for every pair, at FMA speed;
//   * inlier bits are shifted into a per-lane 32-bit word (one SHF per evaluation), counted
//     with POPC, optionally stored as the hypothesis' bitmask.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace rsac {

struct ScoreArgs {
    const ProblemMeta* metas;
    const ScoreTile* tiles;
    const float4* cA;        // (X, Y, Z, cx-u)
    const float4* cB;        // (cy-v, thr, band, 0)
    const float2* uv;        // exact (u, v)
    const void* poses;       // PnP: float[sumH][12]; MLPnP: double[sumH][12]
    int32_t* counts;         // [sumH], atomically accumulated (zeroed before the launch)
    uint32_t* hmasks;        // optional per-hypothesis masks
    unsigned long long* exact_counter;   // optional diagnostic
    int32_t chunk_cap;       // capacity of the shared-memory tile in correspondences
};

// ---- exact (reference-arithmetic) evaluations; noinline keeps the hot loop's registers tight ----
__device__ __noinline__ bool pnp_exact_inlier(const float* __restrict__ pose, float X, float Y, float Z,
                                              float u, float v, float thr, const ProblemMeta* m)
{
    // PnPsolver.cpp:250-258
    const float xc = (pose[0] * X + pose[1] * Y + pose[2] * Z) + pose[9];
    const float yc = (pose[3] * X + pose[4] * Y + pose[5] * Z) + pose[10];
    const float zc = (pose[6] * X + pose[7] * Y + pose[8] * Z) + pose[11];
    const float invZc = 1 / zc;
    const float ue = (float)(m->cx + m->fx * (double)xc * (double)invZc);
    const float ve = (float)(m->cy + m->fy * (double)yc * (double)invZc);
    const float du = ue - u, dv = ve - v;
    const float error2 = du * du + dv * dv;
    return error2 < thr;
}

__device__ __noinline__ bool mlpnp_exact_inlier(const double* __restrict__ pose, float X, float Y, float Z,
                                                float u, float v, float thr, const ProblemMeta* m)
{
    // MLPnPsolver.cpp:231-245
    const float xc = (float)(pose[0] * X + pose[1] * Y + pose[2] * Z + pose[9]);
    const float yc = (float)(pose[3] * X + pose[4] * Y + pose[5] * Z + pose[10]);
    const float zc = (float)(pose[6] * X + pose[7] * Y + pose[8] * Z + pose[11]);
    const float ue = m->k1[0] * xc / zc + m->k1[2];
    const float ve = m->k1[1] * yc / zc + m->k1[3];
    const float distX = u - ue, distY = v - ve;
    const float error2 = distX * distX + distY * distY;
    return error2 < thr;
}

__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// MODEL 0: PnPsolver, MODEL 1: MLPnPsolver
template <int MODEL> struct ScoreModel;
template <> struct ScoreModel<0> {
    using pose_t = float;
    __device__ static void intrinsics(const ProblemMeta& m, float& fx, float& fy) { fx = (float)m.fx; fy = (float)m.fy; }
    __device__ static bool exact(const pose_t* p, float X, float Y, float Z, float u, float v, float thr, const ProblemMeta* m)
    { return pnp_exact_inlier(p, X, Y, Z, u, v, thr, m); }
};
template <> struct ScoreModel<1> {
    using pose_t = double;
    __device__ static void intrinsics(const ProblemMeta& m, float& fx, float& fy) { fx = m.k1[0]; fy = m.k1[1]; }
    __device__ static bool exact(const pose_t* p, float X, float Y, float Z, float u, float v, float thr, const ProblemMeta* m)
    { return mlpnp_exact_inlier(p, X, Y, Z, u, v, thr, m); }
};

// Folded, scaled projective rows of one hypothesis: c[0..3] = fx*[r0|t0]/B, c[4..7] = fy*[r1|t1]/B,
// c[8..11] = [r2|t2]/B with B = max_row|r|_1 + |t|_inf.  A non-finite or zero pose gets a zero
// third row, which drives every evaluation onto the exact path (1/z = inf => band = inf).
template <typename PT>
__device__ __forceinline__ void fold_pose(const PT* __restrict__ p, float fx, float fy, float* c)
{
    float r[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i] = (float)p[i];
    const float l0 = fabsf(r[0]) + fabsf(r[1]) + fabsf(r[2]);
    const float l1 = fabsf(r[3]) + fabsf(r[4]) + fabsf(r[5]);
    const float l2 = fabsf(r[6]) + fabsf(r[7]) + fabsf(r[8]);
    const float rho = fmaxf(l0, fmaxf(l1, l2));
    const float T = fmaxf(fabsf(r[9]), fmaxf(fabsf(r[10]), fabsf(r[11])));
    const float B = rho + T;
    bool sane = (B > 1e-30f) && (B < 1e30f);
#pragma unroll
    for (int i = 0; i < 12; ++i) sane = sane && (fabsf(r[i]) <= 1e30f);   // false for NaN
    const float invB = 1.0f / B;
    if (sane) {
        c[0] = (fx * r[0]) * invB; c[1] = (fx * r[1]) * invB; c[2] = (fx * r[2]) * invB; c[3] = (fx * r[9]) * invB;
        c[4] = (fy * r[3]) * invB; c[5] = (fy * r[4]) * invB; c[6] = (fy * r[5]) * invB; c[7] = (fy * r[10]) * invB;
        c[8] = r[6] * invB;        c[9] = r[7] * invB;        c[10] = r[8] * invB;       c[11] = r[11] * invB;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = 0.0f;
    }
}

// One fast evaluation; shifts the provisional inlier bit and the "uncertain" bit into the
// lane's words (most significant position first: call for i = 31 .. 0).
__device__ __forceinline__ void eval_fast(const float* c, const float4& a, const float4& b, uint32_t& inl, uint32_t& unc)
{
    const float x = fmaf(c[0], a.x, fmaf(c[1], a.y, fmaf(c[2], a.z, c[3])));
    const float y = fmaf(c[4], a.x, fmaf(c[5], a.y, fmaf(c[6], a.z, c[7])));
    const float z = fmaf(c[8], a.x, fmaf(c[9], a.y, fmaf(c[10], a.z, c[11])));
    const float iz = rcp_fast(z);
    const float dx = fmaf(x, iz, a.w);                    // fx*x/z + (cx - u)
    const float dy = fmaf(y, iz, b.x);
    const float d = fmaf(dy, dy, fmaf(dx, dx, -b.y));     // e - thr
    const float dc = fminf(d, b.y);                       // NaN -> thr (outlier); clamp enables the guard
    const float gb = b.z * fabsf(iz);                     // rounding band at this depth
    const float tt = fabsf(dc) - gb;                      // < 0  <=>  inside the band
    inl = __funnelshift_l(__float_as_uint(dc), inl, 1);   // sign(dc): e < thr
    unc = __funnelshift_l(__float_as_uint(tt), unc, 1);
}

template <int HPL, int MODEL>
__global__ void __launch_bounds__(256) score_kernel(ScoreArgs args)
{
    using Model = ScoreModel<MODEL>;
    using PT = typename Model::pose_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* sA = reinterpret_cast<float4*>(smem_raw);
    float4* sB = sA + args.chunk_cap;
    __shared__ __align__(8) uint64_t bar;

    const ScoreTile tile = args.tiles[blockIdx.x];
    const ProblemMeta* mp = args.metas + tile.problem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)tile.nc * 16u;
        const size_t off = (size_t)mp->corr_off + (size_t)tile.corr0;
        mbar_arrive_expect_tx(&bar, 2u * bytes);
        tma_load_1d(sA, args.cA + off, bytes, &bar);
        tma_load_1d(sB, args.cB + off, bytes, &bar);
    }

    // hypotheses of this lane (overlaps the bulk copy)
    const int H = mp->H;
    float fx, fy;
    Model::intrinsics(*mp, fx, fy);
    float c[HPL][12];
    int hyp[HPL];
    uint32_t live[HPL];
    int cnt[HPL];
    const PT* poses = reinterpret_cast<const PT*>(args.poses);
#pragma unroll
    for (int s = 0; s < HPL; ++s) {
        hyp[s] = tile.hyp0 + (warp * HPL + s) * 32 + lane;
        live[s] = (hyp[s] < H) ? 0xffffffffu : 0u;
        cnt[s] = 0;
        if (live[s]) {
            fold_pose<PT>(poses + (size_t)(mp->hyp_off + hyp[s]) * 12, fx, fy, c[s]);
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) c[s][i] = 0.0f;
        }
    }
    // whole warp beyond H: nothing to do (warp-uniform)
    if (tile.hyp0 + warp * HPL * 32 >= H) return;

    mbar_wait(&bar, 0);

    const int nwords = (tile.nc + 31) >> 5;
    for (int w = 0; w < nwords; ++w) {
        uint32_t inl[HPL], unc[HPL];
#pragma unroll
        for (int s = 0; s < HPL; ++s) { inl[s] = 0u; unc[s] = 0u; }
        const float4* pa = sA + w * 32;
        const float4* pb = sB + w * 32;
#pragma unroll
        for (int i = 31; i >= 0; --i) {
            const float4 a = pa[i];
            const float4 b = pb[i];
#pragma unroll
            for (int s = 0; s < HPL; ++s) eval_fast(c[s], a, b, inl[s], unc[s]);
        }
        const int rem = tile.nc - w * 32;
        const uint32_t valid = (rem >= 32) ? 0xffffffffu : ((1u << rem) - 1u);
#pragma unroll
        for (int s = 0; s < HPL; ++s) {
            inl[s] &= valid & live[s];
            uint32_t u = unc[s] & valid & live[s];
            if (u) {
                if (args.exact_counter) atomicAdd(args.exact_counter, (unsigned long long)__popc(u));
                const PT* pose = poses + (size_t)(mp->hyp_off + hyp[s]) * 12;
                while (u) {
                    const int i = __ffs(u) - 1;
                    u &= u - 1;
                    const int ci = w * 32 + i;
                    const float4 a = sA[ci];
                    const float4 b = sB[ci];
                    const float2 p2 = args.uv[(size_t)mp->corr_off + tile.corr0 + ci];
                    const bool in = Model::exact(pose, a.x, a.y, a.z, p2.x, p2.y, b.y, mp);
                    inl[s] = in ? (inl[s] | (1u << i)) : (inl[s] & ~(1u << i));
                }
            }
            cnt[s] += __popc(inl[s]);
            if (args.hmasks && live[s])
                args.hmasks[mp->hmask_off + (int64_t)hyp[s] * mp->words + (tile.corr0 >> 5) + w] = inl[s];
        }
    }
#pragma unroll
    for (int s = 0; s < HPL; ++s)
        if (live[s]) atomicAdd(args.counts + mp->hyp_off + hyp[s], cnt[s]);
}

// ---- packing: raw correspondences -> (cA, cB) tiles with thresholds and rounding bands ----
// band_i bounds |e_fast - e_reference| * |z/B| for an evaluation whose error is near thr_i
// (derivation in DESIGN.md): with u = 2^-24, M = 1 + |X|_inf, q = 1 + (|c|_inf + sqrt(thr))/f_min,
//   A = 3.83 sqrt(thr) * 12u * f_max * M * q,  C = 3.83 sqrt(thr) * u (4|c|_inf + |uv|_inf + 6 sqrt(thr)) + 4u thr
//   band = 2 * (A + C*M)
__device__ __forceinline__ float score_band(float X, float Y, float Z, float cu, float cv, float u, float v,
                                            float thr, float fmin_, float fmax_)
{
    const double uro = (double)kUnitRoundoff;
    const double M = 1.0 + fmax(fabs((double)X), fmax(fabs((double)Y), fabs((double)Z)));
    const double st = sqrt(fmax((double)thr, 0.0));
    const double cinf = fmax(fabs((double)cu), fabs((double)cv));
    const double q = 1.0 + (cinf + st) / (double)fmin_;
    const double A = 3.83 * st * 12.0 * uro * (double)fmax_ * M * q;
    const double Cc = 3.83 * st * uro * (4.0 * cinf + fmax(fabs((double)u), fabs((double)v)) + 6.0 * st) + 4.0 * uro * (double)thr;
    const double band = 2.0 * (A + Cc * M);
    // round up to float; NaN/inf inputs give a NaN/inf band => evaluations stay on the exact path or are outliers
    return __double2float_ru(band);
}

// One thread per correspondence; blockIdx.y = problem.  thr = sigma2*th2 as an f32 product
// (PnPsolver.cpp:93) unless a ready-made max_err array is supplied (scoring stress).
__global__ void pack_pnp_kernel(const ProblemMeta* metas, const float* p3d, const float* p2d, const float* sigma2,
                                const float* th2_per_problem, const float* max_err, int model,
                                float4* cA, float4* cB, float2* uv)
{
    const ProblemMeta& m = metas[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.n; i += gridDim.x * blockDim.x) {
        const size_t g = (size_t)m.corr_off + i;
        const float X = p3d[3 * g], Y = p3d[3 * g + 1], Z = p3d[3 * g + 2];
        const float u = p2d[2 * g], v = p2d[2 * g + 1];
        const float thr = max_err ? max_err[g] : sigma2[g] * th2_per_problem[blockIdx.y];
        float cu, cv, fmin_, fmax_;
        if (model == 0) {
            cu = (float)(m.cx - (double)u);
            cv = (float)(m.cy - (double)v);
            fmin_ = (float)fmin(fabs(m.fx), fabs(m.fy));
            fmax_ = (float)fmax(fabs(m.fx), fabs(m.fy));
        } else {
            cu = m.k1[2] - u;
            cv = m.k1[3] - v;
            fmin_ = fminf(fabsf(m.k1[0]), fabsf(m.k1[1]));
            fmax_ = fmaxf(fabsf(m.k1[0]), fabsf(m.k1[1]));
        }
        cA[g] = make_float4(X, Y, Z, cu);
        cB[g] = make_float4(cv, thr, score_band(X, Y, Z, cu, cv, u, v, thr, fmin_, fmax_), 0.0f);
        uv[g] = make_float2(u, v);
    }
}

}  // namespace rsac
