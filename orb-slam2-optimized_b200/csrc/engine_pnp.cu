// engine_pnp.cu -- C ABI for the batched PnPsolver (EPnP RANSAC sweep, early exit in stages) and the cfg5 scoring
// entry points (include/ransac_b200.h, "PnPsolver").
#include "engine_shared.cuh"
#include "engine_early.cuh"
#include "pnp_pipeline.cuh"
#include "epnp_subwarp.cuh"
#include "select.cuh"
#include "bow.cuh"

// The 4-point minimal solver: three lanes per hypothesis (epnp_subwarp.cuh); RSAC_SOLVE_IMPL=thread selects the
// round-1 one-thread-per-hypothesis kernel (A/B timing only)
static bool solve_subwarp()
{
    static const bool v = [] { const char* s = getenv("RSAC_SOLVE_IMPL"); return !(s && strcmp(s, "thread") == 0); }();
    return v;
}

// hypotheses of one resident wave of the minimal solver
static int64_t solve_wave_hyps(rsac_engine* e)
{
    if (e->pnp.run_eigen) return (int64_t)2 * e->sm_count * 128;
    return solve_subwarp() ? (int64_t)RSAC_SW_BLOCKS * e->sm_count * kSwHypsPerBlock
                           : (int64_t)RSAC_SOLVE_BLOCKS * e->sm_count * RSAC_SOLVE_THREADS;
}

static int solve_range_setup(rsac_engine* e);

// minimal solves of hypotheses [lo, lo + span) of the listed problems (list == nullptr: all C problems; `most` = upper
// bound of the work items, for the grid)
static int launch_solve_range(rsac_engine* e, PnpState& s, const int32_t* list, const int32_t* list_count, int lo, int span, int64_t most)
{
    const BatchDims& d = s.d;
    RSAC_TRY(solve_range_setup(e));
    e->stage_begin(RSAC_STAGE_SOLVE);
    if (s.run_eigen) {
        const int resident = 2 * e->sm_count;
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((most + 127) / 128, resident));
        epnp_minimal_range_kernel<false><<<blocks, 128, 0, e->stream>>>(
            (const ProblemMeta*)s.d_metas.p, d.C, list, list_count, lo, span, (const uint32_t*)s.d_tables.p,
            (const float4*)s.d_cA.p, (const float4*)s.d_uv.p, (float*)s.d_poses.p);
    } else if (solve_subwarp()) {
        const int resident = RSAC_SW_BLOCKS * e->sm_count;
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((most + kSwHypsPerBlock - 1) / kSwHypsPerBlock, resident));
        epnp_minimal_subwarp_kernel<<<blocks, kSwThreads, kSwSmemBytes, e->stream>>>(
            (const ProblemMeta*)s.d_metas.p, d.C, list, list_count, lo, span, (const uint32_t*)s.d_tables.p,
            (const float4*)s.d_cA.p, (const float4*)s.d_uv.p, (float*)s.d_poses.p);
    } else {
        const int resident = RSAC_SOLVE_BLOCKS * e->sm_count;
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((most + RSAC_SOLVE_THREADS - 1) / RSAC_SOLVE_THREADS, resident));
        epnp_minimal_range_kernel<true><<<blocks, RSAC_SOLVE_THREADS, sizeof(double) * kSolveSmemDoubles * RSAC_SOLVE_THREADS, e->stream>>>(
            (const ProblemMeta*)s.d_metas.p, d.C, list, list_count, lo, span, (const uint32_t*)s.d_tables.p,
            (const float4*)s.d_cA.p, (const float4*)s.d_uv.p, (float*)s.d_poses.p);
    }
    e->stage_end(RSAC_STAGE_SOLVE);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

static int pnp_first_phase(rsac_engine* e, const BatchDims& d)
{
    // first stage: three quarters of one wave of the minimal solver unless the caller chose (whole blocks: 1024
    // problems x 55 hypotheses = 440 blocks of 128 <= 444 resident is a full wave).  Measured on cfg4 with six sweeps in
    // flight, ms per sweep resident / end to end: (55,177) 0.449 / 0.504, (55,110,220) 0.440 / 0.512,
    // (40,80,160) 0.421 / 0.474, (28,55,110,220) 0.420 / 0.498, (20,40,80,160) 0.428 / 0.500
    int HA = e->first_phase > 0 ? e->first_phase : env_int("RSAC_EE_HA", 0);
    if (HA <= 0) {
        // measured with the three-lane solver (wave = 160 hypotheses per SM), six sweeps in flight, M candidates/s
        // resident / end to end: (17,34,...) 2.54 / 2.21, (23,46,92,184) 2.65 / 2.32, (46,92,184) 2.66 / 2.38,
        // (41,82,164) 2.69 / 2.40, (35,70,140) 2.68 / 2.38, (30,60,120,240) 2.66 / 2.36: about 1.8 waves first
        const int64_t wave_h = solve_wave_hyps(e);
        const int wave = (int)(wave_h / std::max(d.C, 1));
        // (round 2, final) two FULL waves: 46 at 1024 candidates -- the second round of the launch is then as full as the first
        // (41 left it 77 % full): 2.675 against 2.698 M candidates/s (-0.9 %), the stage-0 launch 17.6 -> 19.4 % of the FP64 peak
        const int first = (int)((solve_subwarp() ? wave_h * 2 : wave_h * 3 / 4) / std::max(d.C, 1));
        HA = wave >= d.maxH ? wave : std::max(16, first);
    }
    return HA;
}

// stage boundaries b0 < b1 < ... < b(K-1) = maxH: hypotheses [0, b0) for every problem, [b(j-1), bj) for the problems
// still without an acceptable hypothesis.  Caller's choice (rsac_set_stages / rsac_set_phases / RSAC_EE_STAGES), else
// b0 from the solver's wave and every further stage doubles what exists (cfg4: 41, 82, 164, 300)
static std::vector<int> pnp_stage_bounds(rsac_engine* e, const BatchDims& d)
{
    return early_stage_bounds(e, d, pnp_first_phase(e, d), "RSAC_EE_STAGES");
}

// indexed wire format (rsac_pnp_upload_indexed): flat p3d / p2d / sigma2 arrays gathered on the device from the frame's
// keypoint table and the map-point table through (keypoint index, map-point index) pairs
static __global__ void pnp_gather_indexed_kernel(int64_t total, const uint16_t* __restrict__ kp_idx, const uint32_t* __restrict__ mp_idx,
                                                 const float* __restrict__ kp_uv, const float* __restrict__ kp_sigma2,
                                                 const float* __restrict__ mp_xyz, float* __restrict__ p3d, float* __restrict__ p2d,
                                                 float* __restrict__ sigma2, int n_kp, int n_mp)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = kp_idx[i], m = mp_idx[i];
        if (k >= (uint32_t)n_kp || m >= (uint32_t)n_mp) {
            // an index outside the tables (a caller bug) never reads out of bounds: the correspondence becomes NaN, i.e.
            // a certain outlier (CheckInliers: NaN compares false), and poisons any minimal set that draws it
            const float qn = __int_as_float(0x7fc00000);
            p2d[2 * i] = p2d[2 * i + 1] = sigma2[i] = p3d[3 * i] = p3d[3 * i + 1] = p3d[3 * i + 2] = qn;
            continue;
        }
        p2d[2 * i] = kp_uv[2 * k]; p2d[2 * i + 1] = kp_uv[2 * k + 1];
        sigma2[i] = kp_sigma2[k];
        p3d[3 * i] = mp_xyz[3 * (size_t)m]; p3d[3 * i + 1] = mp_xyz[3 * (size_t)m + 1]; p3d[3 * i + 2] = mp_xyz[3 * (size_t)m + 2];
    }
}

// Indexed wire format, fused path (one RANSAC parameter set for the whole batch -- the reference's call shape): everything
// of a packed record that depends on the keypoint only -- u, v, cx-u, cy-v, the threshold sigma2*th2, and the two f64
// factors of the scoring kernel's rounding bounds (score_bounds: the square root and the division live here) -- is
// computed once per KEYPOINT (2000 per frame instead of 512 000 correspondences per sweep) ...
struct KpRecord { double A, E; float thr, pad; };   // band = ru(2u M A), eps_a = 12u M E  (score_bounds, operation for operation)
static __global__ void pnp_keypoint_table_kernel(int n_kp, const float* __restrict__ kp_uv, const float* __restrict__ kp_sigma2, float th2,
                                                 double fxd, double fyd, double cxd, double cyd, float4* __restrict__ kpA,
                                                 KpRecord* __restrict__ kpB)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_kp) return;
    const float u = kp_uv[2 * k], v = kp_uv[2 * k + 1];
    const float thr = kp_sigma2[k] * th2;                       // PnPsolver.cpp:93
    const float cu = (float)(cxd - (double)u), cv = (float)(cyd - (double)v);
    const float fx = (float)fxd, fy = (float)fyd;
    // the keypoint-only part of score_bounds (score.cuh), same expressions in the same order
    const double th = fmax((double)thr, 0.0);
    const double s = sqrt(th);
    const double fmin_ = fmin(fabs((double)fx), fabs((double)fy));
    const double fmax_ = fmax(fabs((double)fx), fabs((double)fy));
    const double cinf = fmax(fabs((double)cu), fabs((double)cv));
    const double uvinf = fmax(fabs((double)u), fabs((double)v));
    const double q = 1.0 + (cinf + s) / fmin_;
    KpRecord r;
    r.A = (s * ((12.0 + 8.0 * q) * (fabs((double)fx) + fabs((double)fy)) +
                16.0 * (fabs((double)cu) + fabs((double)cv)) + 2.0 * (fabs((double)u) + fabs((double)v))) +
           33.0 * th);
    r.E = (fmax_ + cinf + uvinf + s);
    r.thr = thr; r.pad = 0.0f;
    kpA[k] = make_float4(u, v, cu, cv);
    kpB[k] = r;
}

// ... and the packing kernel gathers: per correspondence the map point (12 B, L2-resident table), the keypoint record, two
// f64 products for the bounds.  Writes exactly what pack_pnp_kernel writes (bit for bit: tests compare records, masks
// and the number of exact-tier evaluations of the two paths); the flat p3d / p2d / sigma2 arrays are not materialised
// (rsac_poseopt_from_pnp gathers them on demand).
static __device__ __forceinline__ PackedPoint pack_point_indexed(size_t g, const uint16_t* kp_idx, const uint32_t* mp_idx,
                                                                 const float4* kpA, const KpRecord* kpB, const float* mp_xyz,
                                                                 int n_kp, int n_mp, float th2, const ProblemMeta& m)
{
    PackedPoint r;
    const uint32_t k = kp_idx[g], mi = mp_idx[g];
    if (k >= (uint32_t)n_kp || mi >= (uint32_t)n_mp) {
        // an index outside the tables (a caller bug) never reads out of bounds: the correspondence becomes NaN -- a certain
        // outlier (CheckInliers: NaN compares false) that poisons any minimal set drawing it; same record as the unfused path
        const float qn = __int_as_float(0x7fc00000);
        r.X = r.Y = r.Z = r.u = r.v = qn;
        r.thr = qn * th2;
        r.cu = (float)(m.cx - (double)r.u); r.cv = (float)(m.cy - (double)r.v);
        const ScoreBounds sb = score_bounds(r.X, r.Y, r.Z, r.cu, r.cv, r.u, r.v, r.thr, (float)m.fx, (float)m.fy);
        r.band = sb.band; r.eps2 = sb.eps2;
        return r;
    }
    const float4 a = kpA[k];
    const KpRecord b = kpB[k];
    r.X = mp_xyz[3 * (size_t)mi]; r.Y = mp_xyz[3 * (size_t)mi + 1]; r.Z = mp_xyz[3 * (size_t)mi + 2];
    r.u = a.x; r.v = a.y; r.cu = a.z; r.cv = a.w; r.thr = b.thr;
    const double uro = (double)kUnitRoundoff;
    const double M = 1.0 + fmax(fabs((double)r.X), fmax(fabs((double)r.Y), fabs((double)r.Z)));
    const double band = 2.0 * uro * M * b.A;
    const double eps_a = 12.0 * uro * M * b.E;
    r.band = __double2float_ru(band);
    r.eps2 = __double2float_ru(fmax(10.0 * eps_a * eps_a, 1e-30));
    return r;
}

static __global__ void pack_pnp_indexed_kernel(const ProblemMeta* metas, int C, const uint16_t* __restrict__ kp_idx,
                                               const uint32_t* __restrict__ mp_idx, const float4* __restrict__ kpA,
                                               const KpRecord* __restrict__ kpB, const float* __restrict__ mp_xyz, int n_kp, int n_mp,
                                               float th2, float4* cA, float4* cB, float4* cC, float4* cP)
{
    for (int pr = blockIdx.y; pr < C; pr += gridDim.y) {
        const ProblemMeta& m = metas[pr];
        const int npairs = m.words * 16;
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
            PackedPoint a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b = a;
            const int i0 = 2 * p, i1 = 2 * p + 1;
            if (i0 < m.n) {
                const size_t g = (size_t)m.corr_off + i0;
                a = pack_point_indexed(g, kp_idx, mp_idx, kpA, kpB, mp_xyz, n_kp, n_mp, th2, m);
                cA[g] = make_float4(a.X, a.Y, a.Z, a.cu);
                cB[g] = make_float4(a.cv, a.thr, a.band, 0.0f);
                cC[g] = make_float4(a.u, a.v, 0.0f, 0.0f);
            }
            if (i1 < m.n) {
                const size_t g = (size_t)m.corr_off + i1;
                b = pack_point_indexed(g, kp_idx, mp_idx, kpA, kpB, mp_xyz, n_kp, n_mp, th2, m);
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

// flat p3d / p2d / sigma2 of an indexed batch, on demand (the chained PoseOptimization reads them)
int rsac_internal_pnp_ensure_flat(rsac_engine* e)
{
    PnpState& s = e->pnp;
    if (s.flat_valid || s.d.total <= 0) return RSAC_OK;
    const int blocks = (int)std::min<int64_t>(((int64_t)s.d.total + 255) / 256, (int64_t)e->sm_count * 8);
    ++e->launches;
    pnp_gather_indexed_kernel<<<blocks, 256, 0, e->stream>>>(s.d.total, (const uint16_t*)s.d_kp_idx.p, (const uint32_t*)s.d_mp_idx.p,
                                                               (const float*)s.d_kp_uv.p, (const float*)s.d_kp_s2.p, (const float*)s.d_mp_xyz.p,
                                                               (float*)s.d_p3d.p, (float*)s.d_p2d.p, (float*)s.d_sigma2.p,
                                                               s.n_keypoints, s.n_mappoints);
    RSAC_CUDA(e, cudaGetLastError());
    s.flat_valid = true;
    return RSAC_OK;
}

static int pnp_pack(rsac_engine* e)
{
    PnpState& s = e->pnp;
    const BatchDims& d = s.d;
    if (s.packed) return RSAC_OK;
    if (d.total > 0 && d.C > 0 && s.indexed_fused) {
        dim3 grid((unsigned)std::max(1, std::min(64, (d.maxN / 2 + 255) / 256)), (unsigned)std::min(d.C, 65535));
        e->stage_begin(RSAC_STAGE_PACK);
        if (s.kp_table_dirty) {
            ++e->launches;
            pnp_keypoint_table_kernel<<<(s.n_keypoints + 255) / 256, 256, 0, e->stream>>>(
                s.n_keypoints, (const float*)s.d_kp_uv.p, (const float*)s.d_kp_s2.p, s.kp_th2, s.metas[0].fx, s.metas[0].fy,
                s.metas[0].cx, s.metas[0].cy, (float4*)s.d_kpA.p, (KpRecord*)s.d_kpB.p);
            s.kp_table_dirty = false;
        }
        pack_pnp_indexed_kernel<<<grid, 256, 0, e->stream>>>((const ProblemMeta*)s.d_metas.p, d.C, (const uint16_t*)s.d_kp_idx.p,
                                                             (const uint32_t*)s.d_mp_idx.p, (const float4*)s.d_kpA.p,
                                                             (const KpRecord*)s.d_kpB.p, (const float*)s.d_mp_xyz.p, s.n_keypoints,
                                                             s.n_mappoints, s.kp_th2, (float4*)s.d_cA.p, (float4*)s.d_cB.p,
                                                             (float4*)s.d_uv.p, (float4*)s.d_cP.p);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
    } else if (d.total > 0 && d.C > 0) {
        dim3 grid((unsigned)std::max(1, std::min(64, (d.maxN + 255) / 256)), (unsigned)std::min(d.C, 65535));
        e->stage_begin(RSAC_STAGE_PACK);
        pack_pnp_kernel<<<grid, 256, 0, e->stream>>>((const ProblemMeta*)s.d_metas.p, (const float*)s.d_p3d.p, (const float*)s.d_p2d.p,
                                                     (const float*)s.d_sigma2.p, (const float*)s.d_th2.p, nullptr, 0,
                                                     (float4*)s.d_cA.p, (float4*)s.d_cB.p, (float4*)s.d_uv.p, (float4*)s.d_cP.p, d.C);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
    }
    s.packed = true;
    return RSAC_OK;
}

static int pnp_upload_impl(rsac_engine* e, const rsac_pnp_batch* b, const rsac_pnp_indexed_batch* ib, bool idx_on_device = false);

int rsac_pnp_upload(rsac_engine* e, const rsac_pnp_batch* b) { return pnp_upload_impl(e, b, nullptr); }

// (keypoint, map point) index pairs of every candidate straight from SearchByBoW's match arrays: one CTA per pair walks the
// frame's features in order and compacts the matched ones (ballot + prefix) to offsets[c] ..; a candidate the host discarded
// (fewer than min_matches: offsets[c + 1] == offsets[c]) writes nothing
static __global__ void __launch_bounds__(128) pnp_pairs_from_bow_kernel(int C, const int64_t* __restrict__ t2q_off, const int32_t* __restrict__ t2q,
                                                                 const BowSet* __restrict__ sets, const int32_t* __restrict__ qset,
                                                                 const int32_t* __restrict__ tset, const uint32_t* __restrict__ mp_index,
                                                                 const int32_t* __restrict__ offsets, uint16_t* __restrict__ kp_idx,
                                                                 uint32_t* __restrict__ mp_idx)
{
    __shared__ int s_base, s_warp[4];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = offsets[c + 1] - offsets[c];
    if (cap <= 0) return;
    const BowSet Q = sets[qset[c]], T = sets[tset[c]];
    const int32_t* m = t2q + t2q_off[c];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int j0 = 0; j0 < T.n_feat; j0 += blockDim.x) {
        const int j = j0 + tid;
        const int q = j < T.n_feat ? m[j] : -1;
        const unsigned b = __ballot_sync(0xffffffffu, q >= 0);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int k = before + __popc(b & ((1u << lane) - 1u));
        if (q >= 0 && k < cap) {
            kp_idx[offsets[c] + k] = (uint16_t)j;
            mp_idx[offsets[c] + k] = mp_index[Q.feat_off + q];
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 4; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
}

int rsac_pnp_upload_from_bow(rsac_engine* e, const rsac_pnp_from_bow* fb)
{
    if (!e || !fb || !fb->n_matches || !fb->K || !fb->params || !fb->seeds) return RSAC_ERR_INVALID;
    BowState& bw = e->bow;
    if (!bw.ran) { e->err = "rsac_pnp_upload_from_bow before rsac_bow_run"; return RSAC_ERR_STATE; }
    if (bw.mode != 0 || !bw.one_target || !bw.have_mp_index) {
        e->err = "rsac_pnp_upload_from_bow needs a mode-0 SearchByBoW batch against ONE frame whose keyframe sets carry mp_index";
        return RSAC_ERR_STATE;
    }
    const int C = bw.C;
    if (C > 0 && bw.target_n_feat[0] > 65536) { e->err = "the frame has more than 65536 keypoints"; return RSAC_ERR_INVALID; }
    std::vector<int32_t> offsets((size_t)C + 1, 0);
    for (int c = 0; c < C; ++c) {
        const int n = fb->n_matches[c];
        if (n < 0 || n > bw.target_n_feat[c]) { e->err = "n_matches does not belong to the last SearchByBoW run"; return RSAC_ERR_INVALID; }
        if ((int64_t)offsets[c] + n > INT32_MAX) { e->err = "batch too large"; return RSAC_ERR_INVALID; }
        offsets[c + 1] = offsets[c] + (n >= fb->min_matches ? n : 0);            // Tracking.cpp:1215-1219: nmatches < 15 -> vbDiscarded
    }
    rsac_pnp_indexed_batch ib;
    memset(&ib, 0, sizeof(ib));
    ib.n_keypoints = fb->n_keypoints; ib.kp_uv = fb->kp_uv; ib.kp_sigma2 = fb->kp_sigma2;
    ib.n_mappoints = fb->n_mappoints; ib.mp_xyz = fb->mp_xyz;
    ib.C = C; ib.offsets = offsets.data(); ib.K = fb->K; ib.params = fb->params; ib.n_params = 1; ib.seeds = fb->seeds;
    PnpState& s = e->pnp;
    if ((!ib.kp_uv && s.n_keypoints == 0) || (!ib.mp_xyz && s.n_mappoints == 0)) { e->err = "no resident keypoint / map-point table"; return RSAC_ERR_STATE; }
    if ((ib.kp_uv && (ib.n_keypoints <= 0 || ib.n_keypoints > 65536 || !ib.kp_sigma2)) || (ib.mp_xyz && ib.n_mappoints <= 0)) {
        e->err = "bad keypoint / map-point table"; return RSAC_ERR_INVALID;
    }
    std::vector<double> K((size_t)std::max(C, 1) * 4);
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < 4; ++k) K[4 * (size_t)c + k] = fb->K[k];
    rsac_pnp_batch b;
    memset(&b, 0, sizeof(b));
    b.C = C; b.offsets = offsets.data(); b.K = K.data(); b.params = fb->params; b.n_params = 1; b.seeds = fb->seeds;
    RSAC_TRY(pnp_upload_impl(e, &b, &ib, true));
    if (!s.indexed_fused) { e->err = "rsac_pnp_upload_from_bow needs the fused indexed path (RSAC_INDEXED_FUSED=0 is set)"; return RSAC_ERR_STATE; }
    if (C > 0 && s.d.total > 0) {
        RSAC_TRY(e->d_scratch.ensure(e, 4 * ((size_t)C + 1)));
        int32_t* h = (int32_t*)s.h_fromBow.ensure(4 * ((size_t)C + 1));
        if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
        memcpy(h, offsets.data(), 4 * ((size_t)C + 1));
        RSAC_CUDA(e, cudaMemcpyAsync(e->d_scratch.p, h, 4 * ((size_t)C + 1), cudaMemcpyHostToDevice, e->stream));
        s.h_fromBow.mark(e->stream);
        ++e->launches;
        pnp_pairs_from_bow_kernel<<<C, 128, 0, e->stream>>>(C, (const int64_t*)bw.d_t2q_off.p, (const int32_t*)bw.d_t2q.p, (const BowSet*)bw.d_sets.p,
                                                           (const int32_t*)bw.d_qset.p, (const int32_t*)bw.d_tset.p, (const uint32_t*)bw.d_mp_index.p,
                                                           (const int32_t*)e->d_scratch.p, (uint16_t*)s.d_kp_idx.p, (uint32_t*)s.d_mp_idx.p);
        RSAC_CUDA(e, cudaGetLastError());
    }
    return RSAC_OK;
}

int rsac_pnp_upload_indexed(rsac_engine* e, const rsac_pnp_indexed_batch* ib)
{
    if (!e || !ib || ib->C < 0 || !ib->offsets || !ib->K) return RSAC_ERR_INVALID;
    if (ib->n_params < 1 || !ib->params) return RSAC_ERR_INVALID;
    const int64_t total = ib->offsets[ib->C];
    if (total > 0 && (!ib->kp_idx || !ib->mp_idx)) { e->err = "kp_idx / mp_idx is NULL"; return RSAC_ERR_INVALID; }
    PnpState& s = e->pnp;
    if ((!ib->kp_uv && s.n_keypoints == 0) || (!ib->mp_xyz && s.n_mappoints == 0)) { e->err = "no resident keypoint / map-point table"; return RSAC_ERR_STATE; }
    if ((ib->kp_uv && (ib->n_keypoints <= 0 || ib->n_keypoints > 65536 || !ib->kp_sigma2)) || (ib->mp_xyz && ib->n_mappoints <= 0)) {
        e->err = "bad keypoint / map-point table"; return RSAC_ERR_INVALID;
    }
    std::vector<double> K((size_t)std::max(ib->C, 1) * 4);
    for (int c = 0; c < ib->C; ++c)
        for (int k = 0; k < 4; ++k) K[4 * (size_t)c + k] = ib->K[k];
    rsac_pnp_batch b;
    memset(&b, 0, sizeof(b));
    b.C = ib->C; b.offsets = ib->offsets; b.K = K.data(); b.params = ib->params; b.n_params = ib->n_params;
    b.seeds = ib->seeds; b.tables = ib->tables; b.table_offsets = ib->table_offsets;
    return pnp_upload_impl(e, &b, ib);
}

static int pnp_upload_impl(rsac_engine* e, const rsac_pnp_batch* b, const rsac_pnp_indexed_batch* ib, bool idx_on_device)
{
    if (!e || !b || b->C < 0 || !b->offsets || !b->params || b->n_params < 1) return RSAC_ERR_INVALID;
    if (!b->seeds && !b->tables) { if (e) e->err = "need seeds or tables"; return RSAC_ERR_INVALID; }
    if (b->C > 0 && !b->K) { e->err = "K is NULL"; return RSAC_ERR_INVALID; }
    // the minimal solver is the reference's call shape: mRansacMinSet = 4 (Tracking.cpp:1226; PnPsolver.hpp:27 default).
    // Other sizes would need an n-point minimal kernel and a different table stride: refused, not silently misread
    for (int c = 0; c < b->n_params; ++c)
        if (b->params[c].min_set != 4) { e->err = "PnP needs min_set = 4"; return RSAC_ERR_INVALID; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    PnpState& s = e->pnp;
    s.uploaded = false; s.ran = false; s.ee_mode = false;
    std::vector<float> th2;
    int rc = pnp_build_metas(e, b->C, b->offsets, b->params, b->n_params, b->seeds, b->table_offsets, b->tables != nullptr, s.metas, th2, s.d);
    if (rc) return rc;
    for (int c = 0; c < b->C; ++c) {
        s.metas[c].fx = b->K[4 * c]; s.metas[c].fy = b->K[4 * c + 1]; s.metas[c].cx = b->K[4 * c + 2]; s.metas[c].cy = b->K[4 * c + 3];
    }
    const BatchDims& d = s.d;
    const size_t tot = (size_t)std::max(d.total, 1);
    // The scoring plans (tiles, chunking, per-CTA work lists) depend only on the batch's shape -- n, H and the focal
    // lengths of every problem, and the stage boundaries -- so a batch shaped like the previous one reuses them,
    // including the work arrays already on the device (a relocalisation loop with a fixed match budget per candidate)
    const std::vector<int> bounds_now = pnp_stage_bounds(e, d);
    std::vector<int32_t> sig;
    sig.reserve(4 * (size_t)b->C + 4);
    for (int c = 0; c < b->C; ++c) {
        const float fx = (float)s.metas[c].fx, fy = (float)s.metas[c].fy;
        int32_t fxb, fyb;
        memcpy(&fxb, &fx, 4); memcpy(&fyb, &fy, 4);
        sig.push_back(s.metas[c].n); sig.push_back(s.metas[c].H); sig.push_back(fxb); sig.push_back(fyb);
    }
    for (int v : bounds_now) sig.push_back(v);
    sig.push_back(b->C);
    const bool same_shape = s.plans_valid && sig == s.shape_sig;
    if (!same_shape) {
        s.plans_valid = false;
        s.ee_planned = false;
        ++s.plan_version;
        RSAC_TRY(plan_score<0>(e, s.metas, d.maxH, s.groups, s.plan));
    }

    RSAC_TRY(s.d_metas.ensure(e, sizeof(ProblemMeta) * std::max(d.C, 1)));
    RSAC_TRY(s.d_th2.ensure(e, sizeof(float) * std::max(d.C, 1)));
    RSAC_TRY(s.d_p3d.ensure(e, tot * 12));
    RSAC_TRY(s.d_p2d.ensure(e, tot * 8));
    RSAC_TRY(s.d_sigma2.ensure(e, tot * 4));
    RSAC_TRY(s.d_cA.ensure(e, tot * 16));
    RSAC_TRY(s.d_cB.ensure(e, tot * 16));
    RSAC_TRY(s.d_cP.ensure(e, (size_t)std::max<int64_t>(d.total_words, 1) * 1024));
    RSAC_TRY(s.d_uv.ensure(e, tot * 16));
    RSAC_TRY(s.d_tables.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.table_len, 1)));
    RSAC_TRY(s.d_poses.ensure(e, sizeof(float) * 12 * (size_t)std::max<int64_t>(d.sumH, 1)));
    RSAC_TRY(s.d_counts.ensure(e, sizeof(int32_t) * ((size_t)std::max<int64_t>(d.sumH, 1) + 8 + s.groups.size())));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_result) * std::max(d.C, 1)));
    RSAC_TRY(s.d_masks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_words, 1)));
    RSAC_TRY(s.d_sel.ensure(e, tot * 4));
    RSAC_TRY(s.d_pw.ensure(e, tot * 24));
    RSAC_TRY(s.d_us.ensure(e, tot * 16));
    RSAC_TRY(s.d_al.ensure(e, tot * 32));
    RSAC_TRY(s.d_extra.ensure(e, tot * 96));   // refine scratch: 12 doubles per correspondence

    cudaStream_t st = e->stream;
    RSAC_TRY(stage_small_tables(e, s, th2, !same_shape));
    if (ib) {
        // tables (when given) replace the resident ones; per sweep only 6 bytes per correspondence cross PCIe
        if (ib->kp_uv) {
            RSAC_TRY(s.d_kp_uv.ensure(e, (size_t)ib->n_keypoints * 8));
            RSAC_TRY(s.d_kp_s2.ensure(e, (size_t)ib->n_keypoints * 4));
            RSAC_CUDA(e, cudaMemcpyAsync(s.d_kp_uv.p, ib->kp_uv, (size_t)ib->n_keypoints * 8, cudaMemcpyHostToDevice, st));
            RSAC_CUDA(e, cudaMemcpyAsync(s.d_kp_s2.p, ib->kp_sigma2, (size_t)ib->n_keypoints * 4, cudaMemcpyHostToDevice, st));
            s.n_keypoints = ib->n_keypoints;
        }
        if (ib->mp_xyz) {
            RSAC_TRY(s.d_mp_xyz.ensure(e, (size_t)ib->n_mappoints * 12));
            RSAC_CUDA(e, cudaMemcpyAsync(s.d_mp_xyz.p, ib->mp_xyz, (size_t)ib->n_mappoints * 12, cudaMemcpyHostToDevice, st));
            s.n_mappoints = ib->n_mappoints;
        }
        // one parameter set for the whole batch (the reference's call shape): gather and packing are one kernel, fed by a
        // per-keypoint table (pnp_pack); otherwise the flat arrays are gathered first and packed like a flat upload
        s.indexed_fused = ib->n_params == 1 && env_int("RSAC_INDEXED_FUSED", 1) != 0;
        s.flat_valid = false;
        if (s.indexed_fused) {
            const float th2v = th2.empty() ? 0.0f : th2[0];
            const double kk[4] = {ib->K[0], ib->K[1], ib->K[2], ib->K[3]};
            if (ib->kp_uv || th2v != s.kp_th2 || memcmp(kk, s.kp_K, sizeof(kk)) != 0) s.kp_table_dirty = true;
            s.kp_th2 = th2v;
            memcpy(s.kp_K, kk, sizeof(kk));
            RSAC_TRY(s.d_kpA.ensure(e, (size_t)s.n_keypoints * 16));
            RSAC_TRY(s.d_kpB.ensure(e, (size_t)s.n_keypoints * sizeof(KpRecord)));
        }
        if (d.total > 0) {
            RSAC_TRY(s.d_kp_idx.ensure(e, (size_t)d.total * 2));
            RSAC_TRY(s.d_mp_idx.ensure(e, (size_t)d.total * 4));
            if (!idx_on_device) {       // (rsac_pnp_upload_from_bow writes the pairs with a kernel, after this function)
                RSAC_CUDA(e, cudaMemcpyAsync(s.d_kp_idx.p, ib->kp_idx, (size_t)d.total * 2, cudaMemcpyHostToDevice, st));
                RSAC_CUDA(e, cudaMemcpyAsync(s.d_mp_idx.p, ib->mp_idx, (size_t)d.total * 4, cudaMemcpyHostToDevice, st));
                if (!s.indexed_fused) RSAC_TRY(rsac_internal_pnp_ensure_flat(e));
            }
        }
    } else if (d.total > 0) {
        s.indexed_fused = false;
        s.flat_valid = true;
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p3d.p, b->p3d, (size_t)d.total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p2d.p, b->p2d, (size_t)d.total * 8, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_sigma2.p, b->sigma2, (size_t)d.total * 4, cudaMemcpyHostToDevice, st));
    }
    s.have_tables = b->tables != nullptr;
    if (s.have_tables && d.table_len > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_tables.p, b->tables, sizeof(uint32_t) * (size_t)d.table_len, cudaMemcpyHostToDevice, st));
    s.h2d_bytes = (uint64_t)d.total * (ib ? 6 : 24) + sizeof(ProblemMeta) * (uint64_t)d.C + sizeof(float) * (uint64_t)d.C +
                  sizeof(ScoreGroup) * s.groups.size() + (s.have_tables ? sizeof(uint32_t) * (uint64_t)d.table_len : 0);
    // the packing kernel is launched by the first run: a caller that uploads sweep k+1 while sweep k computes
    // (two engines) then overlaps only DMA with the kernels of sweep k -- a concurrent packing kernel takes block
    // slots from the solver and replay kernels, whose grids are sized to exactly one wave
    s.packed = false;
    s.tables_ready = false;
    {
        // the early-exit plans travel with the upload: no H2D copy is left for the run (copies issued by a run wait
        // for the previous sweep in the copy queue, in front of the next sweep's inputs)
        if (d.sumH > 0 && bounds_now.size() > 1 && !s.ee_planned) RSAC_TRY(early_plan<0>(e, s, bounds_now));
    }
    s.shape_sig.swap(sig);
    s.plans_valid = true;
    s.uploaded = true;
    return RSAC_OK;
}

// ---- early exit in phases (pnp_pipeline.cuh) ----
static int solve_range_setup(rsac_engine* e)
{
    if (e->pnp.run_eigen) return RSAC_OK;
    if (solve_subwarp()) {
        if (kSwSmemBytes > 32 * 1024)
            RSAC_TRY(set_func_attr_max(e, (const void*)epnp_minimal_subwarp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmemBytes));
        const size_t need = (kSwSmemBytes + 1024) * RSAC_SW_BLOCKS;
        const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024) + 1);
        RSAC_TRY(set_func_attr_max(e, (const void*)epnp_minimal_subwarp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        return RSAC_OK;
    }
    const size_t smem = sizeof(double) * kSolveSmemDoubles * RSAC_SOLVE_THREADS;
    if (smem > 32 * 1024)
        RSAC_TRY(set_func_attr_max(e, (const void*)epnp_minimal_range_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t need = (smem + 1024) * RSAC_SOLVE_BLOCKS;
    const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024) + 2);
    RSAC_TRY(set_func_attr_max(e, (const void*)epnp_minimal_range_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    return RSAC_OK;
}

static int pnp_launch_select(rsac_engine* e, int flags, const int32_t* d_resume, void* d_results_out, int only_phase);

template <> struct EarlyHooks<0> {
    static int solve_range(rsac_engine* e, PnpState& s, const int32_t* list, const int32_t* list_count, int lo, int span, int64_t most)
    { return launch_solve_range(e, s, list, list_count, lo, span, most); }
    static int select(rsac_engine* e, int flags, const int32_t* d_resume, void* d_results_out, int only_phase)
    { return pnp_launch_select(e, flags, d_resume, d_results_out, only_phase); }
    static int setup(rsac_engine* e) { return solve_range_setup(e); }
    static int stage0_hpl() { return 1; }
    // later stages and chunk size: one hypothesis per lane and 128-correspondence chunks (12 KB of ring per CTA instead of
    // 48 KB: more scoring CTAs resident beside the solver's) -- 0.364 -> 0.350 ms per cfg4 sweep with six in flight
    static int stage_hpl() { return 1; }
    static int stage_chunk_words() { return 4; }
};

static int pnp_launch_select(rsac_engine* e, int flags, const int32_t* d_resume, void* d_results_out, int only_phase)
{
    PnpState& s = e->pnp;
    const BatchDims& d = s.d;
    SelectArgs a;
    if (s.ee_mode) { a.ee = (int32_t*)s.d_ee.p; a.C = d.C; a.first_phase = s.ee_HA; a.only_phase = only_phase; }
    a.metas = (const ProblemMeta*)s.d_metas.p; a.cA = (const float4*)s.d_cA.p; a.cB = (const float4*)s.d_cB.p; a.cC = (const float4*)s.d_uv.p;
    a.poses = s.d_poses.p; a.counts = (const int32_t*)s.d_counts.p; a.cov = nullptr; a.flags = flags; a.resume = d_resume;
    a.sel = (uint32_t*)s.d_sel.p; a.pw_s = (double*)s.d_pw.p; a.us_s = (double*)s.d_us.p; a.al_s = (double*)s.d_al.p; a.tm_s = (double*)s.d_extra.p;
    a.results = s.d_results.p; a.results2 = d_results_out; a.masks = (uint32_t*)s.d_masks.p;
    a.problem_base = e->problem_base;
    a.problem_ids = (e->n_problem_ids == d.C && d.C > 0) ? (const int32_t*)e->d_problem_ids.p : nullptr;
    const size_t smem = (size_t)(3 * d.maxWords + 1) * 4 + 16;
    if (smem > 32 * 1024) RSAC_TRY(set_func_attr_max(e, (const void*)ransac_select_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        cudaFuncAttributes fa;
        RSAC_CUDA(e, cudaFuncGetAttributes(&fa, ransac_select_kernel<0>));
        const size_t need = (fa.sharedSizeBytes + smem + 1024) * kSelectCtasPerSm;
        const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024) + 2);
        RSAC_TRY(set_func_attr_max(e, (const void*)ransac_select_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    // the main replay of a sweep split around the first Refine's eigen-solve (select.cuh, SelectArgs::split_phase); later
    // iterate() calls and the clean-up replay run in one launch
    static const bool split_default = env_int("RSAC_SELECT_SPLIT", 1) != 0;
    if (split_default && only_phase < 0 && !d_resume && d.C > 0) {
        RSAC_TRY(s.d_carry.ensure(e, sizeof(SelectCarry) * (size_t)d.C));
        if (kSelEigSmemPerWarp * kSelEigWarps > 48 * 1024)
            RSAC_TRY(set_func_attr_max(e, (const void*)select_eigen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelEigSmemPerWarp * kSelEigWarps)));
        a.carry = (SelectCarry*)s.d_carry.p;
        a.split_phase = 1;
        e->stage_begin(RSAC_STAGE_SELECT);
        ransac_select_kernel<0><<<d.C, kSelectThreads, smem, e->stream>>>(a);
        e->stage_end(RSAC_STAGE_SELECT);
        e->stage_begin(RSAC_STAGE_SELECT);
        select_eigen_kernel<<<(d.C + kSelEigWarps - 1) / kSelEigWarps, kSelEigWarps * 32, kSelEigSmemPerWarp * kSelEigWarps, e->stream>>>(a.carry, d.C);
        e->stage_end(RSAC_STAGE_SELECT);
        a.split_phase = 2;
        e->stage_begin(RSAC_STAGE_SELECT);
        ransac_select_kernel<0><<<d.C, kSelectThreads, smem, e->stream>>>(a);
        e->stage_end(RSAC_STAGE_SELECT);
        RSAC_CUDA(e, cudaGetLastError());
        return RSAC_OK;
    }
    e->stage_begin(RSAC_STAGE_SELECT);
    ransac_select_kernel<0><<<d.C, kSelectThreads, smem, e->stream>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

// later iterate() calls on solvers that already ran: only the replay stage, from per-problem cursors
int rsac_pnp_rerun(rsac_engine* e, int flags, const int32_t* resume_from, void* d_results_out)
{
    if (!e || !resume_from) return RSAC_ERR_INVALID;
    PnpState& s = e->pnp;
    if (!s.ran) { e->err = "rsac_pnp_rerun before rsac_pnp_run"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    if (s.d.C == 0) return RSAC_OK;
    RSAC_TRY(e->d_resume.ensure(e, sizeof(int32_t) * (size_t)s.d.C));
    RSAC_CUDA(e, cudaMemcpyAsync(e->d_resume.p, resume_from, sizeof(int32_t) * (size_t)s.d.C, cudaMemcpyHostToDevice, e->stream));
    RSAC_TRY(early_complete<0>(e, s));     // the last run stopped early: compute what the resumed scan may need
    return pnp_launch_select(e, flags, (const int32_t*)e->d_resume.p, d_results_out, -1);
}

int rsac_pnp_run(rsac_engine* e, int flags, void* d_results_out)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& s = e->pnp;
    if (!s.uploaded) { e->err = "rsac_pnp_run before rsac_pnp_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    const BatchDims& d = s.d;
    cudaStream_t st = e->stream;
    const ProblemMeta* metas = (const ProblemMeta*)s.d_metas.p;
    if (d.C == 0) { s.ran = true; return RSAC_OK; }
    // the minimal-set tables depend on the seeds only: generated once per upload, later runs of the same
    // batch reuse them (like tables passed in by the host)
    const bool need_rng = !s.have_tables && d.table_len > 0 && !s.tables_ready;
    if (need_rng && !s.packed && e->aux_stream && !e->profile) {
        // first run after an upload: table generation (one warp per problem, latency-bound) and packing
        // (bandwidth-bound) are independent: fork the former onto the side stream
        RSAC_CUDA(e, cudaEventRecord(e->ev_fork, st));
        RSAC_CUDA(e, cudaStreamWaitEvent(e->aux_stream, e->ev_fork, 0));
        ++e->launches;
        rng_tables_kernel<<<(d.C + kRngWarps - 1) / kRngWarps, kRngWarps * 32, 0, e->aux_stream>>>(metas, d.C, (uint32_t*)s.d_tables.p);
        RSAC_CUDA(e, cudaGetLastError());
        RSAC_CUDA(e, cudaEventRecord(e->ev_join, e->aux_stream));
        RSAC_TRY(pnp_pack(e));
        RSAC_CUDA(e, cudaStreamWaitEvent(st, e->ev_join, 0));
        s.tables_ready = true;
    } else {
        RSAC_TRY(pnp_pack(e));
        if (need_rng) {
            e->stage_begin(RSAC_STAGE_RNG);
            rng_tables_kernel<<<(d.C + kRngWarps - 1) / kRngWarps, kRngWarps * 32, 0, st>>>(metas, d.C, (uint32_t*)s.d_tables.p);
            e->stage_end(RSAC_STAGE_RNG);
            RSAC_CUDA(e, cudaGetLastError());
            s.tables_ready = true;
        }
    }
    s.ee_mode = false;
    s.run_eigen = (flags & RSAC_FLAG_EPNP_EIGEN) != 0;
    if ((flags & RSAC_FLAG_EARLY_EXIT) && d.sumH > 0) {
        const std::vector<int> bounds = pnp_stage_bounds(e, d);
        if (bounds.size() > 1) return early_run<0>(e, s, flags, d_results_out, bounds);
    }
    if (d.sumH > 0) {
        const bool eigen = (flags & RSAC_FLAG_EPNP_EIGEN) != 0;
        if (sizeof(double) * kSolveSmemDoubles * RSAC_SOLVE_THREADS > 32 * 1024)
            RSAC_TRY(set_func_attr_max(e, (const void*)epnp_minimal_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(sizeof(double) * kSolveSmemDoubles * RSAC_SOLVE_THREADS)));
        {
            // ask for exactly the shared-memory carve-out that keeps RSAC_SOLVE_BLOCKS blocks resident (the rest of
            // the 228 KB stays L1 for the kernel's local memory); left to its own devices the driver was seen to pick
            // a smaller carve-out in some processes, which silently drops a block per SM (0.96 vs 0.82 ms)
            const size_t need = (sizeof(double) * kSolveSmemDoubles * RSAC_SOLVE_THREADS + 1024) * RSAC_SOLVE_BLOCKS;
            const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024) + 2);
            RSAC_TRY(set_func_attr_max(e, (const void*)epnp_minimal_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
        const int threads = eigen ? 128 : RSAC_SOLVE_THREADS;
        const unsigned blocks = (unsigned)((d.sumH + threads - 1) / threads);
        if (eigen || solve_subwarp()) {
            RSAC_TRY(launch_solve_range(e, s, nullptr, nullptr, 0, d.maxH, (int64_t)d.C * d.maxH));
        } else {
            e->stage_begin(RSAC_STAGE_SOLVE);
            epnp_minimal_kernel<true><<<blocks, threads, sizeof(double) * kSolveSmemDoubles * threads, st>>>(metas, d.C, d.sumH, (const uint32_t*)s.d_tables.p,
                                                                  (const float4*)s.d_cA.p, (const float4*)s.d_uv.p, (float*)s.d_poses.p);
            e->stage_end(RSAC_STAGE_SOLVE);
            RSAC_CUDA(e, cudaGetLastError());
        }

        ScoreArgs sa;
        RSAC_TRY(zero_score_region(e, s.d_counts, d.sumH, (int)s.groups.size(), sa));
        sa.metas = metas;
        sa.cP = (const float4*)s.d_cP.p; sa.cC = (const float4*)s.d_uv.p;
        sa.poses = s.d_poses.p;
        sa.hmasks = nullptr;
        if (flags & RSAC_FLAG_KEEP_MASKS) {
            RSAC_TRY(s.d_hmasks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_hwords, 1)));
            sa.hmasks = (uint32_t*)s.d_hmasks.p;
        }
        int rc = launch_score<0>(e, sa, s.plan, (int)s.groups.size(), s.d_visit);
        if (rc) return rc;
    }
    {
        int rc = pnp_launch_select(e, flags, nullptr, d_results_out, -1);
        if (rc) return rc;
    }
    s.ran = true;
    return RSAC_OK;
}

int rsac_pnp_phase_stats(rsac_engine* e, int64_t out[4])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    PnpState& s = e->pnp;
    out[0] = out[1] = out[2] = 0;
    out[3] = s.d.sumH;
    if (!s.ran) { e->err = "rsac_pnp_phase_stats before rsac_pnp_run"; return RSAC_ERR_STATE; }
    return early_stats(e, s, out);
}

int rsac_pnp_download_async(rsac_engine* e, rsac_result* results, uint32_t* masks)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& s = e->pnp;
    if (!s.ran) { e->err = "rsac_pnp_download before rsac_pnp_run"; return RSAC_ERR_STATE; }
    const BatchDims& d = s.d;
    if (results && d.C > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(results, s.d_results.p, sizeof(rsac_result) * d.C, cudaMemcpyDeviceToHost, e->stream));
    if (masks && d.total_words > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(masks, s.d_masks.p, sizeof(uint32_t) * (size_t)d.total_words, cudaMemcpyDeviceToHost, e->stream));
    return RSAC_OK;
}

int rsac_pnp_download(rsac_engine* e, rsac_result* results, uint32_t* masks)
{
    int rc = rsac_pnp_download_async(e, results, masks);
    if (rc) return rc;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_pnp_solve(rsac_engine* e, const rsac_pnp_batch* b, int flags, rsac_result* results, uint32_t* masks)
{
    int rc = rsac_pnp_upload(e, b);
    if (rc) return rc;
    rc = rsac_pnp_run(e, flags, nullptr);
    if (rc) return rc;
    return rsac_pnp_download(e, results, masks);
}

int64_t rsac_pnp_total_hypotheses(rsac_engine* e) { return e ? e->pnp.d.sumH : 0; }

int rsac_pnp_get_hypotheses(rsac_engine* e, float* poses, int32_t* counts)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& s = e->pnp;
    if (!s.ran) return RSAC_ERR_STATE;
    if (s.d.sumH > 0) {
        if (poses) RSAC_CUDA(e, cudaMemcpyAsync(poses, s.d_poses.p, sizeof(float) * 12 * (size_t)s.d.sumH, cudaMemcpyDeviceToHost, e->stream));
        if (counts) RSAC_CUDA(e, cudaMemcpyAsync(counts, s.d_counts.p, sizeof(int32_t) * (size_t)s.d.sumH, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

// ------------------------------------------------------- scoring stress (cfg5)
int rsac_score_pnp_upload(rsac_engine* e, int H, const float* poses, int n, const float* p3d, const float* p2d,
                          const float* max_err, const double K[4])
{
    if (!e || H < 0 || n < 0 || !K || (H > 0 && !poses) || (n > 0 && (!p3d || !p2d || !max_err))) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    ScoreState& s = e->score;
    s.uploaded = false; s.ran = false;
    s.metas.assign(1, ProblemMeta());
    ProblemMeta& m = s.metas[0];
    memset(&m, 0, sizeof(m));
    m.n = n; m.H = H; m.words = (n + 31) / 32;
    m.fx = K[0]; m.fy = K[1]; m.cx = K[2]; m.cy = K[3];
    s.H = H; s.n = n;
    RSAC_TRY(plan_score<0>(e, s.metas, H, s.groups, s.plan));
    const size_t tot = (size_t)std::max(n, 1), hh = (size_t)std::max(H, 1);
    RSAC_TRY(s.d_metas.ensure(e, sizeof(ProblemMeta)));
    RSAC_TRY(s.d_p3d.ensure(e, tot * 12));
    RSAC_TRY(s.d_p2d.ensure(e, tot * 8));
    RSAC_TRY(s.d_maxerr.ensure(e, tot * 4));
    RSAC_TRY(s.d_cA.ensure(e, tot * 16));
    RSAC_TRY(s.d_cB.ensure(e, tot * 16));
    RSAC_TRY(s.d_cP.ensure(e, (size_t)std::max<int64_t>((int64_t)m.words, 1) * 1024));
    RSAC_TRY(s.d_uv.ensure(e, tot * 16));
    RSAC_TRY(s.d_poses.ensure(e, hh * 48));
    RSAC_TRY(s.d_counts.ensure(e, (hh + 8 + s.groups.size()) * 4));
    RSAC_TRY(s.d_hmasks.ensure(e, hh * (size_t)std::max(m.words, 1) * 4));
    cudaStream_t st = e->stream;
    RSAC_CUDA(e, cudaMemcpyAsync(s.d_metas.p, s.metas.data(), sizeof(ProblemMeta), cudaMemcpyHostToDevice, st));
    RSAC_TRY(s.d_visit.ensure(e, sizeof(ScoreGroup) * std::max<size_t>(s.plan.work.size(), 1)));
    RSAC_CUDA(e, cudaMemcpyAsync(s.d_visit.p, s.plan.work.data(), sizeof(ScoreGroup) * s.plan.work.size(), cudaMemcpyHostToDevice, st));
    if (n > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p3d.p, p3d, (size_t)n * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p2d.p, p2d, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_maxerr.p, max_err, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    if (H > 0) RSAC_CUDA(e, cudaMemcpyAsync(s.d_poses.p, poses, (size_t)H * 48, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        dim3 grid((unsigned)std::max(1, std::min(256, (n + 255) / 256)), 1);
        e->stage_begin(RSAC_STAGE_PACK);
        pack_pnp_kernel<<<grid, 256, 0, st>>>((const ProblemMeta*)s.d_metas.p, (const float*)s.d_p3d.p, (const float*)s.d_p2d.p,
                                              nullptr, nullptr, (const float*)s.d_maxerr.p, 0,
                                              (float4*)s.d_cA.p, (float4*)s.d_cB.p, (float4*)s.d_uv.p, (float4*)s.d_cP.p, 1);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
    }
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_score_pnp_run(rsac_engine* e, int want_masks)
{
    if (!e) return RSAC_ERR_INVALID;
    ScoreState& s = e->score;
    if (!s.uploaded) { e->err = "rsac_score_pnp_run before upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    ScoreArgs sa;
    RSAC_TRY(zero_score_region(e, s.d_counts, s.H, (int)s.groups.size(), sa));
    (void)st;
    if (s.H > 0 && s.n > 0) {
        sa.metas = (const ProblemMeta*)s.d_metas.p;
        sa.cP = (const float4*)s.d_cP.p; sa.cC = (const float4*)s.d_uv.p;
        sa.poses = s.d_poses.p;
        sa.hmasks = want_masks ? (uint32_t*)s.d_hmasks.p : nullptr;
        int rc = launch_score<0>(e, sa, s.plan, (int)s.groups.size(), s.d_visit);
        if (rc) return rc;
    }
    s.ran = true;
    s.with_masks = want_masks != 0;
    return RSAC_OK;
}

int rsac_score_pnp_download(rsac_engine* e, uint32_t* masks, int32_t* counts)
{
    if (!e) return RSAC_ERR_INVALID;
    ScoreState& s = e->score;
    if (!s.ran) return RSAC_ERR_STATE;
    const int words = (s.n + 31) / 32;
    if (counts && s.H > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(counts, s.d_counts.p, sizeof(int32_t) * (size_t)s.H, cudaMemcpyDeviceToHost, e->stream));
    if (masks && s.with_masks && s.H > 0 && words > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(masks, s.d_hmasks.p, sizeof(uint32_t) * (size_t)s.H * words, cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_score_pnp(rsac_engine* e, int H, const float* poses, int n, const float* p3d, const float* p2d,
                   const float* max_err, const double K[4], uint32_t* masks, int32_t* counts)
{
    int rc = rsac_score_pnp_upload(e, H, poses, n, p3d, p2d, max_err, K);
    if (rc) return rc;
    rc = rsac_score_pnp_run(e, masks != nullptr);
    if (rc) return rc;
    return rsac_score_pnp_download(e, masks, counts);
}

int64_t rsac_score_exact_evals(rsac_engine* e)
{
    if (!e || !e->last_exact) return -1;
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, e->last_exact, sizeof(v), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) return -1;
    cudaStreamSynchronize(e->stream);
    return (int64_t)v;
}

// ------------------------------------------------------- host debug hooks
int rsac_debug_host_epnp4(const double K[4], const float p3d[12], const float p2d[8], float R[9], float t[3])
{
    double pw[12], us[8];
    for (int i = 0; i < 12; ++i) pw[i] = (double)p3d[i];
    for (int i = 0; i < 8; ++i) us[i] = (double)p2d[i];
    const Cam k = {K[0], K[1], K[2], K[3]};
    epnp_compute_pose_small<4, false>(pw, us, k, R, t);
    return RSAC_OK;
}

int rsac_debug_host_epnp4_qr(const double K[4], const float p3d[12], const float p2d[8], float R[9], float t[3])
{
    double pw[12], us[8];
    for (int i = 0; i < 12; ++i) pw[i] = (double)p3d[i];
    for (int i = 0; i < 8; ++i) us[i] = (double)p2d[i];
    const Cam k = {K[0], K[1], K[2], K[3]};
    epnp_compute_pose_small<4, true>(pw, us, k, R, t);
    return RSAC_OK;
}

int rsac_debug_score_clocks(rsac_engine* e, unsigned long long out[64])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    RSAC_CUDA(e, cudaMemcpyFromSymbol(out, g_score_clocks, sizeof(unsigned long long) * 64));
    return RSAC_OK;
}

int rsac_debug_solve_clocks(rsac_engine* e, long long out[16])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    RSAC_CUDA(e, cudaMemcpyFromSymbol(out, g_solve_clocks, sizeof(long long) * 16));
    return RSAC_OK;
}

int rsac_debug_score_all(rsac_engine* e, unsigned long long out[4096])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    RSAC_CUDA(e, cudaMemcpyFromSymbol(out, g_score_all, sizeof(unsigned long long) * 4096));
    return RSAC_OK;
}

int rsac_debug_select_clocks(rsac_engine* e, long long out[16])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    RSAC_CUDA(e, cudaMemcpyFromSymbol(out, g_select_clocks, sizeof(long long) * 16));
    return RSAC_OK;
}

int rsac_debug_host_jacobi12(const double a[144], double w[4], double v[48])
{
    double A[78];
    for (int i = 0; i < 12; ++i)
        for (int j = i; j < 12; ++j) A[tri_idx(12, i, j)] = a[i * 12 + j];
    std::vector<double2> rec(kMaxSweepsRec * 66);
    jacobi_lowest<12, 4>(A, w, v, rec.data());
    return RSAC_OK;
}
