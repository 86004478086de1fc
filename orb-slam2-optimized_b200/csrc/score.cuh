// score.cuh -- CheckInliers for every hypothesis x correspondence pair.
//
// Reference semantics (bit-exact target):
//   PnPsolver::CheckInliers   src/PnPsolver.cpp:241-268   (f32 transform, f32 1/z, f64 projection)
//   MLPnPsolver::CheckInliers src/MLPnPsolver.cpp:222-255 (f64 transform narrowed to f32, two f32 divisions)
//
// Design (DESIGN.md "scoring kernel"):
//   * lane <-> hypothesis: each lane keeps HPL poses in registers as 3x4 projective rows with the
//     intrinsics folded in and the whole matrix scaled by 1/(|R|_1 + |t|_inf);
//   * persistent, warp-specialised CTAs, one per SM (up to 16 consumer warps + a producer warp): the
//     producer walks the CTA's statically dealt (hypothesis tile, correspondence chunk) work items and
//     stages each chunk in a 4-slot shared-memory ring with 1-D TMA bulk copies (full/empty mbarriers);
//     consumer warps never meet at a CTA-wide barrier.  (Measured alternatives: two 8-warp CTAs per SM
//     starve one another -- the warp arbiter favours one CTA, which then leaves the SM half empty for the
//     last third of the launch; a per-chunk atomic work counter costs an L2 round trip per 32
//     correspondences and only papers over that imbalance);
//   * correspondences are read back as warp-broadcast LDS.128: two loads serve 32*HPL evaluations;
//   * two-tier evaluation.  Fast path, division-free: with (x,y,z) = P X,
//         D = (x + (cx-u) z)^2 + (y + (cy-v) z)^2 - thr z^2      (= z^2 (e - thr))
//     is formed with 15 FFMA/FMUL; sign(D) is the provisional inlier bit.  A rigorous bound
//     band*|z| + eps2 on |D_fast - D_reference| (pack kernel below, derivation in DESIGN.md) marks the
//     evaluations that are too close to call.  Only those are re-evaluated with the reference's exact
//     operation sequence, from shared memory.  Result: the reference's inlier bit for every pair;
//   * bits are shifted into per-lane 32-bit words (one SHF per evaluation), counted with POPC and
//     optionally stored as the hypothesis' bitmask.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace rsac {

// a group = (problem, hypothesis tile); its correspondences are cut into chunks of whole mask words.
// The record is self-contained (everything the kernel needs about the problem), so a CTA starts with ONE
// global load before its first TMA copy; each CTA owns a list of such records.
struct __align__(16) ScoreGroup {
    int32_t gid;         // group index; < 0 terminates a CTA's list
    int32_t hyp0;        // first hypothesis of the tile
    int32_t nchunks;
    int32_t chunk_words; // words per chunk (the last chunk may be shorter)
    int32_t corr_off, n, words, H;     // copied from the problem's ProblemMeta
    int32_t hyp_off, word_off;
    int64_t hmask_off;
    float fx, fy;        // f32 focal lengths folded into the fast-path rows
    int32_t problem;
    int32_t first_stride; // this CTA's chunks of the group: first | (stride << 16) -- static round-robin deal
};
static_assert(sizeof(ScoreGroup) == 64, "ScoreGroup is copied as four 16-byte words");

struct ScoreArgs {
    const ProblemMeta* metas;        // exact path only (f64 intrinsics)
    const ScoreGroup* work;          // [grid][vlen] per-CTA lists of group records
    int32_t vlen;
    const float4* cP;                // pair-packed records, 4 x float4 per two correspondences (see pack kernel)
    const float4* cC;                // (u, v, 0, 0)   exact pixel coordinates (exact path)
    const void* poses;               // PnP: float[sumH][12]; MLPnP: double[sumH][12]
    int32_t* counts;                 // [sumH], atomically accumulated (zeroed before the launch)
    uint32_t* hmasks;                // optional per-hypothesis masks
    unsigned long long* exact_counter;   // optional diagnostic
    int32_t chunk_cap;               // capacity of one ring slot in correspondences
    int32_t tile_hyps;               // hypotheses per tile = warps * 32 * HPL
    // early-exit phases (pnp_pipeline.cuh): when `phase` is set, only the groups of problems with
    // phase[problem] == phase_want are scored; producer and consumers read the same stable flag
    const int32_t* phase = nullptr;
    int32_t phase_want = 0;
    // list mode (phases B and C): the problems to score are only known on the device.  `work` is then
    // [problem][tiles_per_problem] and the CTAs stride over list[0 .. *list_count) x tiles
    const int32_t* list = nullptr;
    const int32_t* list_count = nullptr;
    int32_t tiles_per_problem = 1;
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
// matrix: D = 0 for every point, which the certainty test (|D| > eps2) sends to the exact path.
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

// Two fast evaluations of one hypothesis (correspondences 2p and 2p+1) with the packed FP32 pipe
// (FFMA2 / FMUL2: one issue slot, two results).  C[k] = (c[k], c[k]).  Shifts sign(D) (provisional inlier
// bit) and sign(t) (t >= 0: too close to call) into the lane's words, most significant position first:
// call for p = 15 .. 0.  Component-wise the operation sequence is
//   x = fma(c0,X,fma(c1,Y,fma(c2,Z,c3))) (same for y, z);  N = fma(cu,z,x);  M = fma(cv,z,y);
//   q = thr*(z*z);  D = fma(M,M,fma(N,N,-q));  t = fma(band,|z|,eps2-|D|)
// which is what the band derivation in DESIGN.md assumes.
__device__ __forceinline__ void eval_pair(const float2* C, const float4& q0, const float4& q1, const float4& q2,
                                          const float4& q3, uint32_t& inl, uint32_t& cert)
{
    const float2 X = make_float2(q0.x, q0.y), Y = make_float2(q0.z, q0.w), Z = make_float2(q1.x, q1.y);
    const float2 cu = make_float2(q1.z, q1.w), cv = make_float2(q2.x, q2.y), nthr = make_float2(q2.z, q2.w);
    const float2 band = make_float2(q3.x, q3.y), eps2 = make_float2(q3.z, q3.w);
    const float2 x = __ffma2_rn(C[0], X, __ffma2_rn(C[1], Y, __ffma2_rn(C[2], Z, C[3])));
    const float2 y = __ffma2_rn(C[4], X, __ffma2_rn(C[5], Y, __ffma2_rn(C[6], Z, C[7])));
    const float2 z = __ffma2_rn(C[8], X, __ffma2_rn(C[9], Y, __ffma2_rn(C[10], Z, C[11])));
    const float2 N = __ffma2_rn(cu, z, x);                 // z * (u_est - u)
    const float2 M = __ffma2_rn(cv, z, y);
    const float2 nq = __fmul2_rn(nthr, __fmul2_rn(z, z));  // -thr z^2
    const float2 D = __ffma2_rn(M, M, __ffma2_rn(N, N, nq));   // z^2 (e - thr)
    const float t1 = fmaf(band.y, fabsf(z.y), eps2.y - fabsf(D.y));   // >= 0 (or NaN): inside the rounding band
    inl = __funnelshift_l(__float_as_uint(D.y), inl, 1);   // sign(D): e < thr   (NaN results are +qNaN: bit clear)
    cert = __funnelshift_l(__float_as_uint(t1), cert, 1);  // sign(t) set: decision is certain
    const float t0 = fmaf(band.x, fabsf(z.x), eps2.x - fabsf(D.x));
    inl = __funnelshift_l(__float_as_uint(D.x), inl, 1);
    cert = __funnelshift_l(__float_as_uint(t0), cert, 1);
}

// diagnostic: globaltimer (ns) stamps of consumer warp 0 in a few CTAs (rsac_debug_score_clocks):
// [cta][0] kernel entry, [1] first chunk landed, [2] poses folded, [3] last chunk done, [4] exit, [5] chunks
static __device__ unsigned long long g_score_clocks[8][8];
static __device__ unsigned long long g_score_all[1024][4];   // per CTA (first 1024): entry, exit (globaltimer ns), chunks, SM id
__device__ __forceinline__ unsigned long long rsac_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RSAC_SCORE_MARK(i) do { if (dbg_slot >= 0 && threadIdx.x == 0) g_score_clocks[dbg_slot][i] = rsac_globaltimer(); } while (0)

constexpr int kScoreStages = 4;        // shared-memory ring of correspondence chunks
constexpr int kScoreMaxThreads = 544;  // up to 16 consumer warps + 1 producer warp (one CTA per SM)

// Warp-specialised persistent kernel.  blockDim.x = (consumer warps + 1) * 32.
//   producer (last warp, one lane): walks this CTA's group records and its statically dealt chunks of each
//     group (a per-chunk atomic counter was measured to BE the critical path: one L2 round trip per 32
//     correspondences), waits for a free ring slot (empty barrier), writes the slot's header (chunk id,
//     list position, group record) and bulk-copies the chunk's two record arrays into it (full barrier,
//     transaction bytes);
//   consumers: wait for the slot to fill, switch group when the header says so (flush counts, fold the new
//     tile's poses), evaluate 32*HPL hypotheses per warp against the chunk, release the slot.  Warps never
//     meet at a CTA barrier inside the loop; they drift up to kScoreStages-1 chunks apart, which absorbs
//     the rare exact-path excursions.
template <int HPL, int MODEL>
__global__ void __launch_bounds__(kScoreMaxThreads, 1) score_kernel(ScoreArgs args)
{
    using Model = ScoreModel<MODEL>;
    using PT = typename Model::pose_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: kScoreStages x [P: cap/2 pairs x 64 B | C: cap x 16 B], then raw poses: tile_hyps x 12 PT
    const int cap = args.chunk_cap;
    float4* sbuf = reinterpret_cast<float4*>(smem_raw);
    PT* sraw = reinterpret_cast<PT*>(smem_raw + (size_t)cap * 48 * kScoreStages);
    __shared__ __align__(8) uint64_t full_bar[kScoreStages], empty_bar[kScoreStages];
    __shared__ __align__(16) ScoreGroup s_grp[kScoreStages];
    __shared__ int2 s_hdr[kScoreStages];                // (chunk id or -1, position in the CTA's list)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncons = (blockDim.x >> 5) - 1;            // consumer warps
    const int dbg_slot = (blockIdx.x < 4) ? (int)blockIdx.x : ((blockIdx.x + 4 >= gridDim.x) ? (int)(blockIdx.x + 8 - gridDim.x) : -1);
    RSAC_SCORE_MARK(0);
    const unsigned long long t_entry = rsac_globaltimer();
    int dbg_chunks = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kScoreStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], (uint32_t)ncons);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == ncons) {
        // ------------------------------------------------------------- producer
        if (lane == 0) {
            const bool by_list = args.list != nullptr;
            const int4* work = reinterpret_cast<const int4*>(by_list ? args.work : args.work + (size_t)blockIdx.x * args.vlen);
            const int T = args.tiles_per_problem;
            const int k_end = by_list ? *args.list_count * T : args.vlen;
            uint32_t pit = 0;
            for (int k = by_list ? (int)blockIdx.x : 0; k < k_end; k += by_list ? (int)gridDim.x : 1) {
                union { ScoreGroup g; int4 q[4]; } rec;
                const int4* src = by_list ? work + 4 * ((size_t)args.list[k / T] * T + k % T) : work + 4 * k;
#pragma unroll
                for (int i = 0; i < 4; ++i) rec.q[i] = src[i];
                const ScoreGroup& grp = rec.g;
                if (grp.gid < 0) { if (by_list) continue; else break; }
                if (!by_list && args.phase && args.phase[grp.problem] != args.phase_want) continue;   // not in this phase
                // this CTA's share of the group's chunks: first, first + stride, ... (dealt by the host; no
                // global round trip between chunks, so the ring runs kScoreStages chunks ahead of the consumers)
                const int c_first = grp.first_stride & 0xffff, c_stride = max(1, grp.first_stride >> 16);
                for (int c = c_first; c < grp.nchunks; c += c_stride) {
                    const uint32_t stage = pit % kScoreStages;
                    mbar_wait(&empty_bar[stage], ((pit / kScoreStages) & 1u) ^ 1u);
                    s_hdr[stage] = make_int2(c, k);
                    int4* gdst = reinterpret_cast<int4*>(&s_grp[stage]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) gdst[i] = rec.q[i];
                    const int w0 = c * grp.chunk_words;
                    const int nw = min(grp.chunk_words, grp.words - w0);
                    const int nc = min(nw * 32, grp.n - w0 * 32);
                    float4* dst = sbuf + (size_t)stage * cap * 3;
                    RSAC_ASSERT(nw > 0 && nc > 0 && nw * 32 <= cap && grp.chunk_words * 32 <= cap && grp.words == (grp.n + 31) / 32);
                    RSAC_ASSERT(grp.hyp0 >= 0 && grp.hyp0 < grp.H && grp.nchunks == (grp.words + grp.chunk_words - 1) / grp.chunk_words);
#ifdef RSAC_CHECKED
                    // poison the slot: whatever the bulk copies do not overwrite would be scored as NaN
                    for (int q = 0; q < cap * 3; ++q) dst[q] = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)nw * 1024u + (uint32_t)nc * 16u);
                    tma_load_1d(dst, args.cP + ((size_t)grp.word_off + w0) * 64, (uint32_t)nw * 1024u, &full_bar[stage]);
                    tma_load_1d(dst + 2 * cap, args.cC + (size_t)grp.corr_off + (size_t)w0 * 32, (uint32_t)nc * 16u, &full_bar[stage]);
                    ++pit;
                }
            }
            const uint32_t stage = pit % kScoreStages;             // end-of-work marker
            mbar_wait(&empty_bar[stage], ((pit / kScoreStages) & 1u) ^ 1u);
            s_hdr[stage] = make_int2(-1, -1);
            mbar_arrive(&full_bar[stage]);
        }
        return;
    }

    // --------------------------------------------------------------- consumers
    const PT* poses = reinterpret_cast<const PT*>(args.poses);
    float2 C[HPL][12];
    int cnt[HPL];
    uint32_t live[HPL];
    uint32_t act[HPL];       // live and the pose holds no NaN (see enter_group)
    int g_hyp0 = 0, g_cw = 0, g_n = 0, g_words = 0, g_hypoff = 0, g_problem = 0;
    int64_t g_hmask = 0;
    int cur_k = -1;
#pragma unroll
    for (int s = 0; s < HPL; ++s) { cnt[s] = 0; live[s] = 0u; act[s] = 0u; }

    bool g_owned = false;    // this CTA scores every chunk of the current group: counts are stored, not accumulated
    auto flush = [&]() {
#pragma unroll
        for (int s = 0; s < HPL; ++s) {
            int32_t* dst = args.counts + g_hypoff + g_hyp0 + (warp * HPL + s) * 32 + lane;
            if (g_owned) { if (live[s]) *dst = cnt[s]; }
            else if (live[s] && cnt[s]) atomicAdd(dst, cnt[s]);
            cnt[s] = 0;
        }
    };

    // take a group record: tile geometry, then load and fold this warp's poses
    auto enter_group = [&](const ScoreGroup& grp) {
        g_hyp0 = grp.hyp0; g_cw = grp.chunk_words; g_n = grp.n; g_words = grp.words;
        g_hypoff = grp.hyp_off; g_problem = grp.problem; g_hmask = grp.hmask_off;
        g_owned = grp.first_stride == (1 << 16);      // first chunk 0, stride 1
        const int H = grp.H;
        const float fx = grp.fx, fy = grp.fy;
#pragma unroll
        for (int s = 0; s < HPL; ++s) {
            const int local = (warp * HPL + s) * 32 + lane;
            const int hyp = g_hyp0 + local;
            live[s] = (hyp < H) ? 0xffffffffu : 0u;
            float c[12];
            if (live[s]) {
                const PT* src = poses + (size_t)(g_hypoff + hyp) * 12;
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
                fold_pose<PT>(raw, fx, fy, c);
                // A NaN anywhere in the pose makes xc, yc or zc NaN for every point, hence error2 NaN and the
                // reference's `error2 < mvMaxError` false (PnPsolver.cpp:250-262, MLPnPsolver.cpp:231-249): count 0,
                // empty mask, without sending n evaluations down the exact path one by one (a minimal set that holds
                // one map point twice -- two keypoints matched to the same wrong point -- yields such a pose, and one
                // of them used to stretch a whole scoring launch from 0.045 to 0.15 ms).  Inf entries are NOT
                // shortcut: 1/Inf = 0 can leave a finite error2.
                bool has_nan = false;
#pragma unroll
                for (int i = 0; i < 12; ++i) has_nan = has_nan || (raw[i] != raw[i]);
                act[s] = has_nan ? 0u : 0xffffffffu;
                // each thread keeps the raw poses of its own slots for the exact path (no other thread reads them)
#pragma unroll
                for (int i = 0; i < 12; ++i) sraw[(size_t)local * 12 + i] = raw[i];
            } else {
                act[s] = 0u;
#pragma unroll
                for (int i = 0; i < 12; ++i) c[i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) C[s][i] = make_float2(c[i], c[i]);
        }
    };

    // the CTA's first group is known without the producer: fold its poses while the first chunk is in flight
    if (args.list) {
        const int T = args.tiles_per_problem, k = (int)blockIdx.x;
        if (k < *args.list_count * T) {
            union { ScoreGroup g; int4 q[4]; } rec;
            const int4* w0 = reinterpret_cast<const int4*>(args.work + ((size_t)args.list[k / T] * T + k % T));
#pragma unroll
            for (int i = 0; i < 4; ++i) rec.q[i] = w0[i];
            if (rec.g.gid >= 0) { enter_group(rec.g); cur_k = k; }
        }
    } else {
        union { ScoreGroup g; int4 q[4]; } rec;
        const int4* w0 = reinterpret_cast<const int4*>(args.work + (size_t)blockIdx.x * args.vlen);
#pragma unroll
        for (int i = 0; i < 4; ++i) rec.q[i] = w0[i];
        if (rec.g.gid >= 0 && !(args.phase && args.phase[rec.g.problem] != args.phase_want)) { enter_group(rec.g); cur_k = 0; }
    }

    for (uint32_t cit = 0;; ++cit) {
        const uint32_t stage = cit % kScoreStages;
        mbar_wait(&full_bar[stage], (cit / kScoreStages) & 1u);
        const int2 hdr = s_hdr[stage];
        if (hdr.x < 0) break;
        if (cit == 0) RSAC_SCORE_MARK(1);
        ++dbg_chunks;
        RSAC_ASSERT(hdr.y >= 0 && (hdr.y == cur_k || hdr.y > cur_k || args.list != nullptr));     // the producer walks its list forwards
        if (hdr.y != cur_k) {
            // ---- new group: flush the finished tile, take the record from the slot header
            if (cur_k >= 0) flush();
            cur_k = hdr.y;
            enter_group(s_grp[stage]);
        }
        if (cit == 0) RSAC_SCORE_MARK(2);
        const bool warp_live = __any_sync(0xffffffffu, (live[0] != 0u));   // slot 0 holds the lowest hypotheses
        if (warp_live) {
            const float4* sP = sbuf + (size_t)stage * cap * 3;
            const float4* sC = sP + 2 * cap;
            const int w0 = hdr.x * g_cw;
            const int nw = min(g_cw, g_words - w0);
            const int nc = min(nw * 32, g_n - w0 * 32);
            RSAC_ASSERT(hdr.x >= 0 && w0 < g_words && nw > 0 && nc > 0 && nw * 32 <= cap);
            for (int w = 0; w < nw; ++w) {
                uint32_t inl[HPL], cert[HPL];
#pragma unroll
                for (int s = 0; s < HPL; ++s) { inl[s] = 0u; cert[s] = 0u; }
                const float4* pp = sP + w * 64;
#pragma unroll
                for (int p = 15; p >= 0; --p) {
                    const float4 q0 = pp[4 * p], q1 = pp[4 * p + 1], q2 = pp[4 * p + 2], q3 = pp[4 * p + 3];
#pragma unroll
                    for (int s = 0; s < HPL; ++s) eval_pair(C[s], q0, q1, q2, q3, inl[s], cert[s]);
                }
                const int rem = nc - w * 32;
                const uint32_t valid = (rem >= 32) ? 0xffffffffu : ((1u << rem) - 1u);
#pragma unroll
                for (int s = 0; s < HPL; ++s) {
                    inl[s] &= valid & act[s];
                    uint32_t u = ~cert[s] & valid & act[s];
                    if (u) {
                        if (args.exact_counter) atomicAdd(args.exact_counter, (unsigned long long)__popc(u));
                        const PT* pose = sraw + (size_t)((warp * HPL + s) * 32 + lane) * 12;
                        const ProblemMeta* mp = args.metas + g_problem;
                        while (u) {
                            const int i = __ffs(u) - 1;
                            u &= u - 1;
                            const int ci = w * 32 + i;
                            const float* rec = reinterpret_cast<const float*>(sP + (size_t)(ci >> 1) * 4) + (ci & 1);
                            const float4 p2 = sC[ci];
                            const bool in = Model::exact(pose, rec[0], rec[2], rec[4], p2.x, p2.y, -rec[10], mp);
                            inl[s] = in ? (inl[s] | (1u << i)) : (inl[s] & ~(1u << i));
                        }
                    }
                    cnt[s] += __popc(inl[s]);
                    if (args.hmasks && live[s])
                        args.hmasks[g_hmask + (int64_t)(g_hyp0 + (warp * HPL + s) * 32 + lane) * g_words + w0 + w] = inl[s];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);     // slot may be refilled
    }
    RSAC_SCORE_MARK(3);
    if (cur_k >= 0) flush();
    RSAC_SCORE_MARK(4);
    if (dbg_slot >= 0 && threadIdx.x == 0) g_score_clocks[dbg_slot][5] = (unsigned long long)dbg_chunks;
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_score_all[blockIdx.x][0] = t_entry;
        g_score_all[blockIdx.x][1] = rsac_globaltimer();
        g_score_all[blockIdx.x][2] = (unsigned long long)dbg_chunks;
        g_score_all[blockIdx.x][3] = smid;
    }
}

// ---- packing: raw correspondences -> records with thresholds and rounding bounds ----
// With u = 2^-24, M = 1 + |X|_inf, s = sqrt(thr), q = 1 + (|c|_inf + s)/f_min, c = (cx-u, cy-v), and all
// quantities in the kernel's scaled units (rows of K[R|t] divided by B, so |z| <= M, |x| <= fx M, |y| <= fy M):
//   * both sides evaluate N = x + cu z (resp. M) with an ABSOLUTE error <= eps_a = 12 u M (f_max + |c|_inf +
//     |uv|_inf + s), whatever z is: the reference's z_ref^2 * error2 equals (N_ref')^2 + (M_ref')^2 with
//     N_ref' = fx xc + cu zc + O(u)(fx|xc| + |zc||ue|) -- the division by zc cancels;
//   * if R = sqrt(N^2 + M^2) >= s|z| + 3 eps_a both D_fast and D_ref are positive (outlier, certain);
//   * otherwise |N|, |M| <= s|z| + 3 eps_a and
//         |D_fast - z^2 (e_ref - thr)| <= band |z| + eps2,
//     band = 2 u M [ s((12 + 8 q)(fx + fy) + 16(|cu| + |cv|) + 2(|u| + |v|)) + 33 thr ]   (first-order terms,
//     >= 30 % head-room over the term-by-term derivation in DESIGN.md),  eps2 = 10 eps_a^2  (second-order).
// An evaluation is certain iff |D| > band |z| + eps2.  No separate guard for points on the camera plane is
// needed: there |z| -> 0 and the test degenerates to |D| > eps2, i.e. N, M must be clearly non-zero.
struct ScoreBounds { float band, eps2; };

__device__ __forceinline__ ScoreBounds score_bounds(float X, float Y, float Z, float cu, float cv, float u, float v,
                                                    float thr, float fx, float fy)
{
    const double uro = (double)kUnitRoundoff;
    const double M = 1.0 + fmax(fabs((double)X), fmax(fabs((double)Y), fabs((double)Z)));
    const double th = fmax((double)thr, 0.0);
    const double s = sqrt(th);
    const double fmin_ = fmin(fabs((double)fx), fabs((double)fy));
    const double fmax_ = fmax(fabs((double)fx), fabs((double)fy));
    const double cinf = fmax(fabs((double)cu), fabs((double)cv));
    const double uvinf = fmax(fabs((double)u), fabs((double)v));
    const double q = 1.0 + (cinf + s) / fmin_;
    const double band = 2.0 * uro * M * (s * ((12.0 + 8.0 * q) * (fabs((double)fx) + fabs((double)fy)) +
                                              16.0 * (fabs((double)cu) + fabs((double)cv)) + 2.0 * (fabs((double)u) + fabs((double)v))) +
                                         33.0 * th);
    const double eps_a = 12.0 * uro * M * (fmax_ + cinf + uvinf + s);
    // round up to float; NaN/inf inputs give NaN/inf bounds => those evaluations take the exact path
    ScoreBounds r;
    r.band = __double2float_ru(band);
    r.eps2 = __double2float_ru(fmax(10.0 * eps_a * eps_a, 1e-30));
    return r;
}

// One thread per correspondence pair; blockIdx.y = problem.  thr = sigma2*th2 as an f32 product
// (PnPsolver.cpp:93) unless a ready-made max_err array is supplied (scoring stress).
// Outputs:
//   cA (X,Y,Z,cx-u), cB (cy-v,thr,band,0), cC (u,v,0,0): one record per correspondence, indexed by
//       corr_off + i -- minimal solvers, refinement and the exact path read these;
//   cP: the scoring kernel's stream, 64 B per PAIR of correspondences (2p, 2p+1) so that every operand of
//       the packed-FP32 evaluation is an aligned register pair straight out of LDS.128:
//         [X0 X1 Y0 Y1] [Z0 Z1 cu0 cu1] [cv0 cv1 -thr0 -thr1] [band0 band1 eps2_0 eps2_1]
//       indexed by (word_off*16 + p); a problem's tail is zero-padded to a whole 32-correspondence word.
struct PackedPoint { float X, Y, Z, cu, cv, thr, band, eps2, u, v; };

__device__ __forceinline__ PackedPoint pack_point(const ProblemMeta& m, size_t g, const float* p3d, const float* p2d,
                                                  const float* sigma2, float th2, const float* max_err, int model)
{
    PackedPoint r;
    r.X = p3d[3 * g]; r.Y = p3d[3 * g + 1]; r.Z = p3d[3 * g + 2];
    r.u = p2d[2 * g]; r.v = p2d[2 * g + 1];
    r.thr = max_err ? max_err[g] : sigma2[g] * th2;
    float fx, fy;
    if (model == 0) {
        r.cu = (float)(m.cx - (double)r.u);
        r.cv = (float)(m.cy - (double)r.v);
        fx = (float)m.fx; fy = (float)m.fy;
    } else {
        r.cu = m.k1[2] - r.u;
        r.cv = m.k1[3] - r.v;
        fx = m.k1[0]; fy = m.k1[1];
    }
    const ScoreBounds sb = score_bounds(r.X, r.Y, r.Z, r.cu, r.cv, r.u, r.v, r.thr, fx, fy);
    r.band = sb.band; r.eps2 = sb.eps2;
    return r;
}

static __global__ void pack_pnp_kernel(const ProblemMeta* metas, const float* p3d, const float* p2d, const float* sigma2,
                                const float* th2_per_problem, const float* max_err, int model,
                                float4* cA, float4* cB, float4* cC, float4* cP, int C)
{
  for (int pr = blockIdx.y; pr < C; pr += gridDim.y) {       // grid.y is capped at 65535 problems
    const ProblemMeta& m = metas[pr];
    const float th2 = th2_per_problem ? th2_per_problem[pr] : 0.0f;
    const int npairs = m.words * 16;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
        PackedPoint a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b = a;
        const int i0 = 2 * p, i1 = 2 * p + 1;
        if (i0 < m.n) {
            const size_t g = (size_t)m.corr_off + i0;
            a = pack_point(m, g, p3d, p2d, sigma2, th2, max_err, model);
            cA[g] = make_float4(a.X, a.Y, a.Z, a.cu);
            cB[g] = make_float4(a.cv, a.thr, a.band, 0.0f);
            cC[g] = make_float4(a.u, a.v, 0.0f, 0.0f);
        }
        if (i1 < m.n) {
            const size_t g = (size_t)m.corr_off + i1;
            b = pack_point(m, g, p3d, p2d, sigma2, th2, max_err, model);
            cA[g] = make_float4(b.X, b.Y, b.Z, b.cu);
            cB[g] = make_float4(b.cv, b.thr, b.band, 0.0f);
            cC[g] = make_float4(b.u, b.v, 0.0f, 0.0f);
        }
        float4* dst = cP + ((size_t)m.word_off * 16 + p) * 4;
        dst[0] = make_float4(a.X, b.X, a.Y, b.Y);
        dst[1] = make_float4(a.Z, b.Z, a.cu, b.cu);
        dst[2] = make_float4(a.cv, b.cv, -a.thr, -b.thr);
        dst[3] = make_float4(a.band, b.band, a.eps2, b.eps2);
    }
  }
}

}  // namespace rsac
