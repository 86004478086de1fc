// engine_sim3.cu -- C ABI for the batched Sim3Solver (include/ransac_b200.h, "Sim3Solver").
#include "engine_shared.cuh"
#include "sim3.cuh"
#include "guided.cuh"

// what the Sim3Solver constructor computes per match (Sim3Solver.cpp:39-66), gathered from the resident keyframe views:
// X3Dc = Rcw * X3Dw + tcw in float for both keyframes, sigma^2 = mvLevelSigma2[octave] = scale_factor[octave]^2
static __global__ void __launch_bounds__(256) sim3_gather_from_views_kernel(int64_t total, const int32_t* __restrict__ pair_of,
                                                                    const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2,
                                                                    const KfViewDev* __restrict__ views, const int32_t* __restrict__ kf1,
                                                                    const int32_t* __restrict__ kf2, const float* __restrict__ mp_xyz,
                                                                    const int32_t* __restrict__ kp_octave, float* __restrict__ x1c,
                                                                    float* __restrict__ x2c, float* __restrict__ s1, float* __restrict__ s2)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int c = pair_of[k];
        const KfViewDev& v1 = views[kf1[c]];
        const KfViewDev& v2 = views[kf2[c]];
        const size_t g1 = (size_t)v1.feat_off + idx1[k], g2 = (size_t)v2.feat_off + idx2[k];
        const float p1[3] = {mp_xyz[3 * g1], mp_xyz[3 * g1 + 1], mp_xyz[3 * g1 + 2]};
        const float p2[3] = {mp_xyz[3 * g2], mp_xyz[3 * g2 + 1], mp_xyz[3 * g2 + 2]};
        float c1[3], c2[3];
        guided_mat3_vec(v1.Rcw, p1, v1.tcw, c1);
        guided_mat3_vec(v2.Rcw, p2, v2.tcw, c2);
        x1c[3 * k] = c1[0]; x1c[3 * k + 1] = c1[1]; x1c[3 * k + 2] = c1[2];
        x2c[3 * k] = c2[0]; x2c[3 * k + 1] = c2[1]; x2c[3 * k + 2] = c2[2];
        const float f1 = v1.scale_factors[kp_octave[g1]], f2 = v2.scale_factors[kp_octave[g2]];
        s1[k] = f1 * f1;
        s2[k] = f2 * f2;
    }
}

struct Sim3FromViews { const int32_t* idx1; const int32_t* idx2; const int32_t* pair_of; const int32_t* kf1; const int32_t* kf2; int C; };
static int sim3_upload_impl(rsac_engine* e, const rsac_sim3_batch* b, const Sim3FromViews* fv);

int rsac_sim3_upload(rsac_engine* e, const rsac_sim3_batch* b) { return sim3_upload_impl(e, b, nullptr); }

int rsac_sim3_upload_from_views(rsac_engine* e, const rsac_sim3_from_views* b, int32_t* offsets_out, int32_t* idx1_out, int32_t* idx2_out)
{
    if (!e || !b || b->C < 0 || !b->params || b->n_params < 1 || !b->seeds || !offsets_out) return RSAC_ERR_INVALID;
    if (b->C > 0 && (!b->kf1 || !b->kf2 || !b->matches12 || !b->K1 || !b->K2)) return RSAC_ERR_INVALID;
    GuidedState& g = e->guided;
    if (g.n_views <= 0) { e->err = "rsac_sim3_upload_from_views needs resident views (rsac_views_upload)"; return RSAC_ERR_STATE; }
    const int C = b->C;
    std::vector<int32_t> offsets((size_t)C + 1, 0), idx1, idx2, pair_of;
    int64_t moff = 0;
    for (int c = 0; c < C; ++c) {
        const int a = b->kf1[c], q = b->kf2[c];
        if (a < 0 || a >= g.n_views || q < 0 || q >= g.n_views) { e->err = "view index out of range"; return RSAC_ERR_INVALID; }
        const int n1 = g.view_n_feat[a], n2 = g.view_n_feat[q];
        const uint8_t* ok1 = g.view_mp_valid.data() + g.view_feat_off[a];
        const uint8_t* ok2 = g.view_mp_valid.data() + g.view_feat_off[q];
        for (int i1 = 0; i1 < n1; ++i1) {                       // Sim3Solver.cpp:26-70
            const int i2 = b->matches12[moff + i1];
            if (i2 < 0 || i2 >= n2) continue;                   // no match / pMP2 not observed by KF2 (indexKF2 < 0)
            if (!ok1[i1] || !ok2[i2]) continue;                 // !pMP1 || pMP1->isBad() || pMP2->isBad()
            idx1.push_back(i1); idx2.push_back(i2); pair_of.push_back(c);
        }
        moff += n1;
        if (idx1.size() > (size_t)INT32_MAX) { e->err = "batch too large"; return RSAC_ERR_INVALID; }
        offsets[c + 1] = (int32_t)idx1.size();
    }
    memcpy(offsets_out, offsets.data(), sizeof(int32_t) * ((size_t)C + 1));
    if (idx1_out && !idx1.empty()) memcpy(idx1_out, idx1.data(), sizeof(int32_t) * idx1.size());     // mvnIndices1
    if (idx2_out && !idx2.empty()) memcpy(idx2_out, idx2.data(), sizeof(int32_t) * idx2.size());
    rsac_sim3_batch sb;
    memset(&sb, 0, sizeof(sb));
    sb.C = C; sb.offsets = offsets.data(); sb.K1 = b->K1; sb.K2 = b->K2; sb.params = b->params; sb.n_params = b->n_params; sb.seeds = b->seeds;
    Sim3FromViews fv{idx1.data(), idx2.data(), pair_of.data(), b->kf1, b->kf2, C};
    return sim3_upload_impl(e, &sb, &fv);
}

static int sim3_upload_impl(rsac_engine* e, const rsac_sim3_batch* b, const Sim3FromViews* fv)
{
    if (!e || !b || b->C < 0 || !b->offsets || !b->params || b->n_params < 1) return RSAC_ERR_INVALID;
    if (!b->seeds && !b->tables) { e->err = "need seeds or tables"; return RSAC_ERR_INVALID; }
    if (b->tables && !b->table_offsets) { e->err = "tables without table_offsets"; return RSAC_ERR_INVALID; }
    if (b->C > 0 && (!b->K1 || !b->K2)) { e->err = "K1/K2 is NULL"; return RSAC_ERR_INVALID; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    Sim3State& s = e->sim3;
    s.uploaded = false; s.ran = false; s.tables_ready = false;
    const int C = b->C;
    s.metas.assign(C, ProblemMeta());
    BatchDims d;
    d.C = C;
    for (int c = 0; c < C; ++c) {
        const rsac_sim3_params& p = b->params[b->n_params == 1 ? 0 : c];
        ProblemMeta& m = s.metas[c];
        memset(&m, 0, sizeof(m));
        m.corr_off = b->offsets[c];
        m.n = b->offsets[c + 1] - b->offsets[c];
        if (m.n < 0) { e->err = "bad offsets"; return RSAC_ERR_INVALID; }
        int H = 0;
        if (m.n > 0) rsac_sim3_ransac_setup(m.n, &p, &H);
        if (m.n < p.min_inliers || m.n < 3) H = 0;         // iterate() returns at once (Sim3Solver.cpp:119-123)
        m.H = H;
        m.min_inl = p.min_inliers;
        m.min_set = 3;
        m.fix_scale = p.fix_scale;
        m.hyp_off = (int32_t)d.sumH;
        m.words = (m.n + 31) / 32;
        m.word_off = (int32_t)d.total_words;
        m.hmask_off = d.total_hwords;
        m.seed = b->seeds ? b->seeds[c] : 0u;
        if (b->tables) {
            m.table_off = b->table_offsets[c];
            if (b->table_offsets[c + 1] - b->table_offsets[c] < (int64_t)H * 3) { e->err = "index table too short"; return RSAC_ERR_INVALID; }
        } else {
            m.table_off = d.table_len;
        }
        for (int k = 0; k < 4; ++k) { m.k1[k] = b->K1[4 * c + k]; m.k2[k] = b->K2[4 * c + k]; }
        d.table_len += (int64_t)H * 3;
        d.sumH += H;
        d.total_words += m.words;
        d.total_hwords += (int64_t)H * m.words;
        d.maxH = std::max(d.maxH, H);
        d.maxN = std::max(d.maxN, m.n);
        d.maxWords = std::max(d.maxWords, m.words);
    }
    d.total = b->offsets[C];
    if (b->tables) d.table_len = b->table_offsets[C];
    s.d = d;
    const size_t tot = (size_t)std::max(d.total, 1);
    RSAC_TRY(s.d_metas.ensure(e, sizeof(ProblemMeta) * std::max(C, 1)));
    RSAC_TRY(s.d_x1.ensure(e, tot * 12));
    RSAC_TRY(s.d_x2.ensure(e, tot * 12));
    RSAC_TRY(s.d_s1.ensure(e, tot * 4));
    RSAC_TRY(s.d_s2.ensure(e, tot * 4));
    RSAC_TRY(s.d_c1.ensure(e, tot * 48));     // c1 | c2 | c3 back to back
    RSAC_TRY(s.d_tables.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.table_len, 1)));
    RSAC_TRY(s.d_poses.ensure(e, sizeof(float) * 13 * (size_t)std::max<int64_t>(d.sumH, 1)));
    RSAC_TRY(s.d_counts.ensure(e, sizeof(int32_t) * (size_t)std::max<int64_t>(d.sumH, 1)));
    RSAC_TRY(s.d_hmasks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_hwords, 1)));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_result) * std::max(C, 1)));
    RSAC_TRY(s.d_masks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_words, 1)));
    cudaStream_t st = e->stream;
    if (C > 0) RSAC_CUDA(e, cudaMemcpyAsync(s.d_metas.p, s.metas.data(), sizeof(ProblemMeta) * C, cudaMemcpyHostToDevice, st));
    if (d.total > 0 && fv) {
        // 12 B per correspondence (KF1 feature, KF2 feature, pair) instead of 32: the points and octaves are in the resident views
        GuidedState& g = e->guided;
        const size_t n = (size_t)d.total, c1 = (size_t)std::max(C, 1);
        RSAC_TRY(s.d_idx1.ensure(e, 4 * (3 * n + 2 * c1)));
        int32_t* h = (int32_t*)s.h_idx.ensure(4 * (3 * n + 2 * c1));
        if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
        memcpy(h, fv->idx1, 4 * n); memcpy(h + n, fv->idx2, 4 * n); memcpy(h + 2 * n, fv->pair_of, 4 * n);
        memcpy(h + 3 * n, fv->kf1, 4 * (size_t)C); memcpy(h + 3 * n + c1, fv->kf2, 4 * (size_t)C);
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_idx1.p, h, 4 * (3 * n + 2 * c1), cudaMemcpyHostToDevice, st));
        s.h_idx.mark(st);
        const int32_t* di = (const int32_t*)s.d_idx1.p;
        ++e->launches;
        sim3_gather_from_views_kernel<<<(unsigned)std::min<int64_t>((d.total + 255) / 256, (int64_t)e->sm_count * 8), 256, 0, st>>>(
            d.total, di + 2 * n, di, di + n, (const KfViewDev*)g.d_views.p, di + 3 * n, di + 3 * n + c1, (const float*)g.d_mp_xyz.p,
            (const int32_t*)g.d_kp_octave.p, (float*)s.d_x1.p, (float*)s.d_x2.p, (float*)s.d_s1.p, (float*)s.d_s2.p);
        RSAC_CUDA(e, cudaGetLastError());
    } else if (d.total > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_x1.p, b->x1c, (size_t)d.total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_x2.p, b->x2c, (size_t)d.total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_s1.p, b->sigma2_1, (size_t)d.total * 4, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_s2.p, b->sigma2_2, (size_t)d.total * 4, cudaMemcpyHostToDevice, st));
    }
    s.have_tables = b->tables != nullptr;
    if (s.have_tables && d.table_len > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_tables.p, b->tables, sizeof(uint32_t) * (size_t)d.table_len, cudaMemcpyHostToDevice, st));
    if (d.total > 0 && C > 0) {
        float4* c1 = (float4*)s.d_c1.p;
        dim3 grid((unsigned)std::max(1, std::min(64, (d.maxN + 255) / 256)), (unsigned)std::min(C, 65535));
        e->stage_begin(RSAC_STAGE_PACK);
        sim3_pack_kernel<<<grid, 256, 0, st>>>((const ProblemMeta*)s.d_metas.p, (const float*)s.d_x1.p, (const float*)s.d_x2.p,
                                               (const float*)s.d_s1.p, (const float*)s.d_s2.p, c1, c1 + tot, c1 + 2 * tot, C);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
    }
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_sim3_run(rsac_engine* e, int flags, void* d_results_out)
{
    (void)flags;
    if (!e) return RSAC_ERR_INVALID;
    Sim3State& s = e->sim3;
    if (!s.uploaded) { e->err = "rsac_sim3_run before rsac_sim3_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    const BatchDims& d = s.d;
    cudaStream_t st = e->stream;
    if (d.C == 0) { s.ran = true; return RSAC_OK; }
    const ProblemMeta* metas = (const ProblemMeta*)s.d_metas.p;
    if (!s.have_tables && d.table_len > 0 && !s.tables_ready) {     // once per upload: the tables depend on the seeds only
        s.tables_ready = true;
        e->stage_begin(RSAC_STAGE_RNG);
        rng_tables_kernel<<<(d.C + kRngWarps - 1) / kRngWarps, kRngWarps * 32, 0, st>>>(metas, d.C, (uint32_t*)s.d_tables.p);
        e->stage_end(RSAC_STAGE_RNG);
        RSAC_CUDA(e, cudaGetLastError());
    }
    const size_t tot = (size_t)std::max(d.total, 1);
    Sim3Args a;
    a.metas = metas; a.tables = (const uint32_t*)s.d_tables.p;
    a.c1 = (const float4*)s.d_c1.p; a.c2 = a.c1 + tot; a.c3 = a.c1 + 2 * tot;
    a.poses = (float*)s.d_poses.p; a.counts = (int32_t*)s.d_counts.p; a.hmasks = (uint32_t*)s.d_hmasks.p;
    a.results = s.d_results.p; a.results2 = d_results_out; a.masks = (uint32_t*)s.d_masks.p;
    a.problem_base = e->problem_base;
    a.tile = std::max(32, std::min(1024, ((d.maxN + 31) / 32) * 32));
    const int lanes = d.sumH <= 8192 ? 8 : 2;
    const int hpc = sim3_hyps_per_cta(lanes);
    a.tiles_h = std::max(1, (d.maxH + hpc - 1) / hpc);
    if (s.d_done.cap < sizeof(int32_t) * (size_t)d.C) {
        RSAC_TRY(s.d_done.ensure(e, sizeof(int32_t) * (size_t)d.C));
        RSAC_CUDA(e, cudaMemsetAsync(s.d_done.p, 0, s.d_done.cap, st));     // the kernel leaves the counters at zero
    }
    a.done = (int32_t*)s.d_done.p;
    const size_t smem = (size_t)a.tile * 48;
    const void* kern = lanes == 8 ? (const void*)sim3_kernel<8> : (const void*)sim3_kernel<2>;
    if (smem > 32 * 1024) RSAC_TRY(set_func_attr_max(e, kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    e->stage_begin(RSAC_STAGE_SOLVE);
    if (lanes == 8) sim3_kernel<8><<<(unsigned)d.C * (unsigned)a.tiles_h, kSim3Threads, smem, st>>>(a);
    else sim3_kernel<2><<<(unsigned)d.C * (unsigned)a.tiles_h, kSim3Threads, smem, st>>>(a);
    e->stage_end(RSAC_STAGE_SOLVE);
    RSAC_CUDA(e, cudaGetLastError());
    s.ran = true;
    return RSAC_OK;
}

int rsac_sim3_download(rsac_engine* e, rsac_result* results, uint32_t* masks)
{
    if (!e) return RSAC_ERR_INVALID;
    Sim3State& s = e->sim3;
    if (!s.ran) { e->err = "rsac_sim3_download before rsac_sim3_run"; return RSAC_ERR_STATE; }
    const BatchDims& d = s.d;
    if (results && d.C > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(results, s.d_results.p, sizeof(rsac_result) * d.C, cudaMemcpyDeviceToHost, e->stream));
    if (masks && d.total_words > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(masks, s.d_masks.p, sizeof(uint32_t) * (size_t)d.total_words, cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_sim3_solve(rsac_engine* e, const rsac_sim3_batch* b, int flags, rsac_result* results, uint32_t* masks)
{
    int rc = rsac_sim3_upload(e, b);
    if (rc) return rc;
    rc = rsac_sim3_run(e, flags, nullptr);
    if (rc) return rc;
    return rsac_sim3_download(e, results, masks);
}

int64_t rsac_sim3_total_hypotheses(rsac_engine* e) { return e ? e->sim3.d.sumH : 0; }

int rsac_sim3_get_hypotheses(rsac_engine* e, float* poses, int32_t* counts, uint32_t* masks)
{
    if (!e) return RSAC_ERR_INVALID;
    Sim3State& s = e->sim3;
    if (!s.ran) return RSAC_ERR_STATE;
    if (s.d.sumH > 0) {
        if (poses) RSAC_CUDA(e, cudaMemcpyAsync(poses, s.d_poses.p, sizeof(float) * 13 * (size_t)s.d.sumH, cudaMemcpyDeviceToHost, e->stream));
        if (counts) RSAC_CUDA(e, cudaMemcpyAsync(counts, s.d_counts.p, sizeof(int32_t) * (size_t)s.d.sumH, cudaMemcpyDeviceToHost, e->stream));
        if (masks && s.d.total_hwords > 0)
            RSAC_CUDA(e, cudaMemcpyAsync(masks, s.d_hmasks.p, sizeof(uint32_t) * (size_t)s.d.total_hwords, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_debug_host_sim3(const float P1[9], const float P2[9], int fix_scale, float R[9], float t[3], float* s)
{
    sim3_compute(P1, P2, fix_scale, R, t, s);
    return RSAC_OK;
}
