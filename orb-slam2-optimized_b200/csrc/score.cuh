// score.cuh -- CheckInliers for every hypothesis x correspondence pair.
//
// Reference semantics (bit-exact target):
//   PnPsolver::CheckInliers   src/PnPsolver.cpp:241-268   (f32 transform, f32 1/z, f64 projection)
//   MLPnPsolver::CheckInliers src/MLPnPsolver.cpp:222-255 (f64 transform narrowed to f32, two f32 divisions)
//
// Design (DESIGN.md "scoring kernel"):
//   * lane <-> hypothesis: each lane keeps HPL poses in registers as 3x4 projective rows with the
//     intrinsics folded in and the whole matrix scaled by 1/(|R|_1 + |t|_inf);
//   * persistent, warp-specialised CTAs: a producer warp pulls (hypothesis tile, correspondence chunk)
//     work items from per-group atomic counters -- SMs stay balanced at word (32-correspondence)
//     granularity -- and stages each chunk in a 4-slot shared-memory ring with 1-D TMA bulk copies
//     (full/empty mbarriers); consumer warps never meet at a CTA-wide barrier;
//   * correspondences are read back as warp-broadcast LDS.128: two loads serve 32*HPL evaluations;
//   * two-tier evaluation.  Fast path, division-free: with (x,y,z) = P X,
//         D = (x + (cx-u) z)^2 + (y + (cy-v) z)^2 - thr z^2      (= z^2 (e - thr))
//     is formed with 15 FFMA/FMUL; sign(D) is the provisional inlier bit.  A rigorous bound
//     band*|z| on |D_fast - D_reference| (pack kernel below, derivation in DESIGN.md) marks the
//     evaluations that are too close to call: |D| <= band |z|, or thr z^2 <= band |z| (point on
//     the camera plane).  Only those are re-evaluated with the reference's exact operation
//     sequence, from shared memory.  Result: the reference's inlier bit for every pair;
//   * bits are shifted into per-lane 32-bit words (one SHF per evaluation), counted with POPC and
//     optionally stored as the hypothesis' bitmask.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace rsac {

// a group = (problem, hypothesis tile); its correspondences are cut into chunks of whole mask words
struct ScoreGroup {
    int32_t problem;
    int32_t hyp0;        // first hypothesis of the tile
    int32_t nchunks;
    int32_t chunk_words; // words per chunk (the last chunk may be shorter)
};

struct ScoreArgs {
    const ProblemMeta* metas;
    const ScoreGroup* groups;
    int32_t* group_next;     // [ngroups] chunk counters, zeroed before the launch
    int32_t ngroups;
    const int32_t* cta_first; // [grid+1] range of this CTA in `visit`
    const int32_t* visit;     // group ids in the order each CTA works through them
    const float4* cA;        // (X, Y, Z, cx-u)
    const float4* cB;        // (cy-v, thr, band, 0)
    const float4* cC;        // (u, v, 0, 0)   exact pixel coordinates (exact path, minimal solvers)
    const void* poses;       // PnP: float[sumH][12]; MLPnP: double[sumH][12]
    int32_t* counts;         // [sumH], atomically accumulated (zeroed before the launch)
    uint32_t* hmasks;        // optional per-hypothesis masks
    unsigned long long* exact_counter;   // optional diagnostic
    int32_t chunk_cap;       // capacity of one shared-memory buffer in correspondences
    int32_t tile_hyps;       // hypotheses per tile = warps * 32 * HPL
};

// ---- exact (reference-arithmetic) evaluations ----
template <typename PT>
__device__ __forceinline__ bool pnp_exact_core(const PT* __restrict__ pose, float X, float Y, float Z,
                                               float u, float v, float thr, double fx, double fy, double cx, double cy)
{
    // PnPsolver.cpp:250-258
    const float xc = (pose[0] * X + pose[1] * Y + pose[2] * Z) + pose[9];
    const float yc = (pose[3] * X + pose[4] * Y + pose[5] * Z) + pose[10];
    const float zc = (pose[6] * X + pose[7] * Y + pose[8] * Z) + pose[11];
    const float invZc = 1 / zc;
    const float ue = (float)(cx + fx * (double)xc * (double)invZc);
    const float ve = (float)(cy + fy * (double)yc * (double)invZc);
    const float du = ue - u, dv = ve - v;
    const float error2 = du * du + dv * dv;
    return error2 < thr;
}

__device__ __forceinline__ bool mlpnp_exact_core(const double* __restrict__ pose, float X, float Y, float Z,
                                                 float u, float v, float thr, float fx, float fy, float cx, float cy)
{
    // MLPnPsolver.cpp:231-245
    const float xc = (float)(pose[0] * X + pose[1] * Y + pose[2] * Z + pose[9]);
    const float yc = (float)(pose[3] * X + pose[4] * Y + pose[5] * Z + pose[10]);
    const float zc = (float)(pose[6] * X + pose[7] * Y + pose[8] * Z + pose[11]);
    const float ue = fx * xc / zc + cx;
    const float ve = fy * yc / zc + cy;
    const float distX = u - ue, distY = v - ve;
    const float error2 = distX * distX + distY * distY;
    return error2 < thr;
}

// MODEL 0: PnPsolver, MODEL 1: MLPnPsolver
template <int MODEL> struct ScoreModel;
template <> struct ScoreModel<0> {
    using pose_t = float;
    struct Intr { double fx, fy, cx, cy; };
    __device__ static Intr intr(const ProblemMeta& m) { return {m.fx, m.fy, m.cx, m.cy}; }
    __device__ static void fold_f(const ProblemMeta& m, float& fx, float& fy) { fx = (float)m.fx; fy = (float)m.fy; }
    __device__ static bool exact(const pose_t* p, float X, float Y, float Z, float u, float v, float thr, const Intr& k)
    { return pnp_exact_core<float>(p, X, Y, Z, u, v, thr, k.fx, k.fy, k.cx, k.cy); }
    // global-memory variant used by the replay kernel
    __device__ static bool exact(const pose_t* p, float X, float Y, float Z, float u, float v, float thr, const ProblemMeta* m)
    { return pnp_exact_core<float>(p, X, Y, Z, u, v, thr, m->fx, m->fy, m->cx, m->cy); }
};
template <> struct ScoreModel<1> {
    using pose_t = double;
    struct Intr { float fx, fy, cx, cy; };
    __device__ static Intr intr(const ProblemMeta& m) { return {m.k1[0], m.k1[1], m.k1[2], m.k1[3]}; }
    __device__ static void fold_f(const ProblemMeta& m, float& fx, float& fy) { fx = m.k1[0]; fy = m.k1[1]; }
    __device__ static bool exact(const pose_t* p, float X, float Y, float Z, float u, float v, float thr, const Intr& k)
    { return mlpnp_exact_core(p, X, Y, Z, u, v, thr, k.fx, k.fy, k.cx, k.cy); }
    __device__ static bool exact(const pose_t* p, float X, float Y, float Z, float u, float v, float thr, const ProblemMeta* m)
    { return mlpnp_exact_core(p, X, Y, Z, u, v, thr, m->k1[0], m->k1[1], m->k1[2], m->k1[3]); }
};

// Folded, scaled projective rows of one hypothesis: c[0..3] = fx*[r0|t0]/B, c[4..7] = fy*[r1|t1]/B,
// c[8..11] = [r2|t2]/B with B = max_row|r|_1 + |t|_inf.  A non-finite or zero pose gets an all-zero
// matrix: z = 0 for every point, which the band test sends to the exact path.
template <typename PT>
__device__ __forceinline__ void fold_pose(const PT* __restrict__ r, float fx, float fy, float* c)
{
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = (float)r[i];
    const float l0 = fabsf(f[0]) + fabsf(f[1]) + fabsf(f[2]);
    const float l1 = fabsf(f[3]) + fabsf(f[4]) + fabsf(f[5]);
    const float l2 = fabsf(f[6]) + fabsf(f[7]) + fabsf(f[8]);
    const float rho = fmaxf(l0, fmaxf(l1, l2));
    const float T = fmaxf(fabsf(f[9]), fmaxf(fabsf(f[10]), fabsf(f[11])));
    const float B = rho + T;
    bool sane = (B > 1e-30f) && (B < 1e30f);
#pragma unroll
    for (int i = 0; i < 12; ++i) sane = sane && (fabsf(f[i]) <= 1e30f);   // false for NaN
    const float invB = 1.0f / B;
    if (sane) {
        c[0] = (fx * f[0]) * invB; c[1] = (fx * f[1]) * invB; c[2] = (fx * f[2]) * invB; c[3] = (fx * f[9]) * invB;
        c[4] = (fy * f[3]) * invB; c[5] = (fy * f[4]) * invB; c[6] = (fy * f[5]) * invB; c[7] = (fy * f[10]) * invB;
        c[8] = f[6] * invB;        c[9] = f[7] * invB;        c[10] = f[8] * invB;       c[11] = f[11] * invB;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = 0.0f;
    }
}

// One fast evaluation; shifts sign(D) (provisional inlier bit) and sign(t) (t >= 0: too close to
// call) into the lane's words, most significant position first: call for i = 31 .. 0.
__device__ __forceinline__ void eval_fast(const float* c, const float4& a, const float4& b, uint32_t& inl, uint32_t& cert)
{
    const float x = fmaf(c[0], a.x, fmaf(c[1], a.y, fmaf(c[2], a.z, c[3])));
    const float y = fmaf(c[4], a.x, fmaf(c[5], a.y, fmaf(c[6], a.z, c[7])));
    const float z = fmaf(c[8], a.x, fmaf(c[9], a.y, fmaf(c[10], a.z, c[11])));
    const float N = fmaf(a.w, z, x);                      // z * (u_est - u)
    const float M = fmaf(b.x, z, y);
    const float q = b.y * (z * z);                        // thr z^2
    const float D = fmaf(M, M, fmaf(N, N, -q));           // z^2 (e - thr)
    const float m = fminf(fabsf(D), q);                   // q <= band|z|: point on the camera plane
    const float t = fmaf(b.z, fabsf(z), -m);              // >= 0 (or NaN): inside the rounding band
    inl = __funnelshift_l(__float_as_uint(D), inl, 1);    // sign(D): e < thr   (NaN results are +qNaN: bit clear)
    cert = __funnelshift_l(__float_as_uint(t), cert, 1);  // sign(t) set: decision is certain
}

constexpr int kScoreStages = 4;        // shared-memory ring of correspondence chunks
constexpr int kScoreMaxThreads = 288;  // 8 consumer warps + 1 producer warp

// Warp-specialised persistent kernel.  blockDim.x = (consumer warps + 1) * 32.
//   producer (last warp, one lane): walks this CTA's groups, pulls chunk ids from the group's atomic
//     counter (the next id is requested before the current TMA is issued, so the atomic's latency is
//     off the critical path), waits for a free ring slot (empty barrier) and bulk-copies the chunk's
//     three 16-byte record arrays into it (full barrier, transaction bytes);
//   consumers: wait for the slot to fill, evaluate 32*HPL hypotheses per warp against the chunk, release
//     the slot.  Warps never meet at a CTA barrier inside the loop; they drift up to kScoreStages-1
//     chunks apart, which absorbs the rare exact-path excursions.
template <int HPL, int MODEL>
__global__ void __launch_bounds__(kScoreMaxThreads, 2) score_kernel(ScoreArgs args)
{
    using Model = ScoreModel<MODEL>;
    using PT = typename Model::pose_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: kScoreStages x [A | B | C] (cap records each), then raw poses: tile_hyps x 12 PT
    const int cap = args.chunk_cap;
    float4* sbuf = reinterpret_cast<float4*>(smem_raw);
    PT* sraw = reinterpret_cast<PT*>(smem_raw + (size_t)cap * 48 * kScoreStages);
    __shared__ __align__(8) uint64_t full_bar[kScoreStages], empty_bar[kScoreStages];
    __shared__ int s_chunk[kScoreStages];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncons = (blockDim.x >> 5) - 1;            // consumer warps
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kScoreStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], (uint32_t)ncons);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int v_begin = args.cta_first[blockIdx.x], v_end = args.cta_first[blockIdx.x + 1];

    if (warp == ncons) {
        // ------------------------------------------------------------- producer
        if (lane == 0) {
            uint32_t pit = 0;
            for (int k = v_begin; k < v_end; ++k) {
                const int g = args.visit[k];
                const ScoreGroup grp = args.groups[g];
                const ProblemMeta* mp = args.metas + grp.problem;
                const int words_total = mp->words, n = mp->n;
                const size_t base = (size_t)mp->corr_off;
                int next_id = atomicAdd(args.group_next + g, 1);
                for (;;) {
                    const int c = next_id;
                    const uint32_t stage = pit % kScoreStages;
                    mbar_wait(&empty_bar[stage], ((pit / kScoreStages) & 1u) ^ 1u);
                    if (c < grp.nchunks) {
                        s_chunk[stage] = c;
                        next_id = atomicAdd(args.group_next + g, 1);     // consumed next iteration
                        const int w0 = c * grp.chunk_words;
                        const int nw = min(grp.chunk_words, words_total - w0);
                        const int nc = min(nw * 32, n - w0 * 32);
                        const uint32_t bytes = (uint32_t)nc * 16u;
                        float4* dst = sbuf + (size_t)stage * cap * 3;
                        const size_t off = base + (size_t)w0 * 32;
                        mbar_arrive_expect_tx(&full_bar[stage], 3u * bytes);
                        tma_load_1d(dst, args.cA + off, bytes, &full_bar[stage]);
                        tma_load_1d(dst + cap, args.cB + off, bytes, &full_bar[stage]);
                        tma_load_1d(dst + 2 * cap, args.cC + off, bytes, &full_bar[stage]);
                        ++pit;
                    } else {
                        s_chunk[stage] = -1;                             // end of this group
                        mbar_arrive(&full_bar[stage]);
                        ++pit;
                        break;
                    }
                }
            }
        }
        return;
    }

    // --------------------------------------------------------------- consumers
    const PT* poses = reinterpret_cast<const PT*>(args.poses);
    uint32_t cit = 0;
    for (int k = v_begin; k < v_end; ++k) {
        const int g = args.visit[k];
        const ScoreGroup grp = args.groups[g];
        const ProblemMeta* mp = args.metas + grp.problem;
        const int words_total = mp->words;
        const int H = mp->H;
        float c[HPL][12];
        int hyp[HPL];
        uint32_t live[HPL];
        int cnt[HPL];
        bool any_live = false;
#pragma unroll
        for (int s = 0; s < HPL; ++s) {
            hyp[s] = grp.hyp0 + (warp * HPL + s) * 32 + lane;
            live[s] = (hyp[s] < H) ? 0xffffffffu : 0u;
            cnt[s] = 0;
            any_live = any_live || (live[s] != 0u);
        }
        const bool warp_live = __any_sync(0xffffffffu, any_live);
        bool folded = false;
        // poses are loaded and folded when the first chunk of the group arrives (a CTA may find its group
        // already finished by others)
        auto fold_all = [&]() {
            float fx, fy;
            Model::fold_f(*mp, fx, fy);
#pragma unroll
            for (int s = 0; s < HPL; ++s) {
                const int local = (warp * HPL + s) * 32 + lane;
                if (live[s]) {
                    const PT* src = poses + (size_t)(mp->hyp_off + hyp[s]) * 12;
                    PT raw[12];
                    if constexpr (sizeof(PT) == 4) {
                        const float4* s4 = reinterpret_cast<const float4*>(src);
                        const float4 r0 = s4[0], r1 = s4[1], r2 = s4[2];
                        raw[0] = r0.x; raw[1] = r0.y; raw[2] = r0.z; raw[3] = r0.w; raw[4] = r1.x; raw[5] = r1.y;
                        raw[6] = r1.z; raw[7] = r1.w; raw[8] = r2.x; raw[9] = r2.y; raw[10] = r2.z; raw[11] = r2.w;
                    } else {
                        const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
                        for (int i = 0; i < 6; ++i) { const double2 v = s2[i]; raw[2 * i] = v.x; raw[2 * i + 1] = v.y; }
                    }
                    fold_pose<PT>(raw, fx, fy, c[s]);
                    // each thread keeps the raw poses of its own slots for the exact path (no other thread reads them)
#pragma unroll
                    for (int i = 0; i < 12; ++i) sraw[(size_t)local * 12 + i] = raw[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 12; ++i) c[s][i] = 0.0f;
                }
            }
        };

        for (;;) {
            const uint32_t stage = cit % kScoreStages;
            mbar_wait(&full_bar[stage], (cit / kScoreStages) & 1u);
            const int chunk = s_chunk[stage];
            if (chunk >= 0 && warp_live) {
                if (!folded) { fold_all(); folded = true; }
                const float4* sA = sbuf + (size_t)stage * cap * 3;
                const float4* sB = sA + cap;
                const float4* sC = sA + 2 * cap;
                const int w0 = chunk * grp.chunk_words;
                const int nw = min(grp.chunk_words, words_total - w0);
                const int nc = min(nw * 32, mp->n - w0 * 32);
                for (int w = 0; w < nw; ++w) {
                    uint32_t inl[HPL], cert[HPL];
#pragma unroll
                    for (int s = 0; s < HPL; ++s) { inl[s] = 0u; cert[s] = 0u; }
                    const float4* pa = sA + w * 32;
                    const float4* pb = sB + w * 32;
#pragma unroll
                    for (int i = 31; i >= 0; --i) {
                        const float4 a = pa[i];
                        const float4 b = pb[i];
#pragma unroll
                        for (int s = 0; s < HPL; ++s) eval_fast(c[s], a, b, inl[s], cert[s]);
                    }
                    const int rem = nc - w * 32;
                    const uint32_t valid = (rem >= 32) ? 0xffffffffu : ((1u << rem) - 1u);
#pragma unroll
                    for (int s = 0; s < HPL; ++s) {
                        inl[s] &= valid & live[s];
                        uint32_t u = ~cert[s] & valid & live[s];
                        if (u) {
                            if (args.exact_counter) atomicAdd(args.exact_counter, (unsigned long long)__popc(u));
                            const PT* pose = sraw + (size_t)((warp * HPL + s) * 32 + lane) * 12;
                            while (u) {
                                const int i = __ffs(u) - 1;
                                u &= u - 1;
                                const int ci = w * 32 + i;
                                const float4 a = sA[ci];
                                const float4 b = sB[ci];
                                const float4 p2 = sC[ci];
                                const bool in = Model::exact(pose, a.x, a.y, a.z, p2.x, p2.y, b.y, mp);
                                inl[s] = in ? (inl[s] | (1u << i)) : (inl[s] & ~(1u << i));
                            }
                        }
                        cnt[s] += __popc(inl[s]);
                        if (args.hmasks && live[s])
                            args.hmasks[mp->hmask_off + (int64_t)hyp[s] * words_total + w0 + w] = inl[s];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);     // slot may be refilled
            ++cit;
            if (chunk < 0) break;
        }
#pragma unroll
        for (int s = 0; s < HPL; ++s)
            if (live[s] && cnt[s]) atomicAdd(args.counts + mp->hyp_off + hyp[s], cnt[s]);
    }
}

// ---- packing: raw correspondences -> (cA, cB, cC) records with thresholds and rounding bands ----
// band_i * |z| bounds |D_fast - z^2 (e_reference - thr)| for an evaluation that is near the threshold
// (derivation in DESIGN.md).  With u = 2^-24, M = 1 + |X|_inf, s = sqrt(thr),
// q = 1 + (|c|_inf + s)/f_min, c = (cx-u, cy-v):
//   band = 2 u M [ s ((12 + 8 q)(fx + fy) + 16 (|cu| + |cv|) + 2 (|u| + |v|)) + 33 thr ]
// and at least thr * 2^-10 * M, which makes "thr z^2 <= band |z|" cover |z| <= 2^-10 M (cancellation guard).
__device__ __forceinline__ float score_band(float X, float Y, float Z, float cu, float cv, float u, float v,
                                            float thr, float fx, float fy)
{
    const double uro = (double)kUnitRoundoff;
    const double M = 1.0 + fmax(fabs((double)X), fmax(fabs((double)Y), fabs((double)Z)));
    const double th = fmax((double)thr, 0.0);
    const double s = sqrt(th);
    const double fmin_ = fmin(fabs((double)fx), fabs((double)fy));
    const double cinf = fmax(fabs((double)cu), fabs((double)cv));
    const double q = 1.0 + (cinf + s) / fmin_;
    double band = 2.0 * uro * M * (s * ((12.0 + 8.0 * q) * (fabs((double)fx) + fabs((double)fy)) +
                                        16.0 * (fabs((double)cu) + fabs((double)cv)) + 2.0 * (fabs((double)u) + fabs((double)v))) +
                                   33.0 * th);
    band = fmax(band, th * 0x1p-10 * M);
    // round up to float; NaN/inf inputs give a NaN/inf band => those evaluations take the exact path
    return __double2float_ru(band);
}

// One thread per correspondence; blockIdx.y = problem.  thr = sigma2*th2 as an f32 product
// (PnPsolver.cpp:93) unless a ready-made max_err array is supplied (scoring stress).
__global__ void pack_pnp_kernel(const ProblemMeta* metas, const float* p3d, const float* p2d, const float* sigma2,
                                const float* th2_per_problem, const float* max_err, int model,
                                float4* cA, float4* cB, float4* cC)
{
    const ProblemMeta& m = metas[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.n; i += gridDim.x * blockDim.x) {
        const size_t g = (size_t)m.corr_off + i;
        const float X = p3d[3 * g], Y = p3d[3 * g + 1], Z = p3d[3 * g + 2];
        const float u = p2d[2 * g], v = p2d[2 * g + 1];
        const float thr = max_err ? max_err[g] : sigma2[g] * th2_per_problem[blockIdx.y];
        float cu, cv, fx, fy;
        if (model == 0) {
            cu = (float)(m.cx - (double)u);
            cv = (float)(m.cy - (double)v);
            fx = (float)m.fx; fy = (float)m.fy;
        } else {
            cu = m.k1[2] - u;
            cv = m.k1[3] - v;
            fx = m.k1[0]; fy = m.k1[1];
        }
        cA[g] = make_float4(X, Y, Z, cu);
        cB[g] = make_float4(cv, thr, score_band(X, Y, Z, cu, cv, u, v, thr, fx, fy), 0.0f);
        cC[g] = make_float4(u, v, 0.0f, 0.0f);
    }
}

}  // namespace rsac
