// engine_mlpnp.cu -- C ABI for the batched MLPnPsolver (include/ransac_b200.h, "MLPnPsolver").
#include "engine_shared.cuh"
#include "engine_early.cuh"
#include "mlpnp_pipeline.cuh"
#include "select.cuh"

int rsac_mlpnp_upload(rsac_engine* e, const rsac_mlpnp_batch* b)
{
    if (!e || !b || b->C < 0 || !b->offsets || !b->params || b->n_params < 1) return RSAC_ERR_INVALID;
    if (!b->seeds && !b->tables) { e->err = "need seeds or tables"; return RSAC_ERR_INVALID; }
    if (b->C > 0 && !b->K) { e->err = "K is NULL"; return RSAC_ERR_INVALID; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    PnpState& s = e->mlpnp;
    s.uploaded = false; s.ran = false; s.tables_ready = false; s.ee_mode = false; s.ee_planned = false;
    std::vector<float> th2;
    for (int c = 0; c < b->C; ++c)
        if (b->params[b->n_params == 1 ? 0 : c].min_set != 6) { e->err = "MLPnP needs min_set = 6"; return RSAC_ERR_INVALID; }
    int rc = pnp_build_metas(e, b->C, b->offsets, b->params, b->n_params, b->seeds, b->table_offsets, b->tables != nullptr, s.metas, th2, s.d);
    if (rc) return rc;
    for (int c = 0; c < b->C; ++c)
        for (int k = 0; k < 4; ++k) s.metas[c].k1[k] = b->K[4 * c + k];
    const BatchDims& d = s.d;
    const size_t tot = (size_t)std::max(d.total, 1);
    RSAC_TRY(plan_score<1>(e, s.metas, d.maxH, s.groups, s.plan));

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
    RSAC_TRY(s.d_poses.ensure(e, sizeof(double) * 12 * (size_t)std::max<int64_t>(d.sumH, 1)));
    RSAC_TRY(s.d_counts.ensure(e, sizeof(int32_t) * ((size_t)std::max<int64_t>(d.sumH, 1) + 8 + s.groups.size())));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_result) * std::max(d.C, 1)));
    RSAC_TRY(s.d_masks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_words, 1)));
    RSAC_TRY(s.d_sel.ensure(e, tot * 4));
    RSAC_TRY(s.d_pw.ensure(e, tot * sizeof(double) * kMlpnpScratch));
    RSAC_TRY(s.d_extra.ensure(e, tot * 96));   // refine scratch: 12 doubles per correspondence
    if (b->cov) RSAC_TRY(s.d_cov.ensure(e, tot * 72));

    cudaStream_t st = e->stream;
    RSAC_TRY(stage_small_tables(e, s, th2));
    if (d.total > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p3d.p, b->p3d, (size_t)d.total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p2d.p, b->p2d, (size_t)d.total * 8, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_sigma2.p, b->sigma2, (size_t)d.total * 4, cudaMemcpyHostToDevice, st));
        if (b->cov) RSAC_CUDA(e, cudaMemcpyAsync(s.d_cov.p, b->cov, (size_t)d.total * 72, cudaMemcpyHostToDevice, st));
    }
    s.have_cov = b->cov != nullptr;
    s.have_tables = b->tables != nullptr;
    if (s.have_tables && d.table_len > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_tables.p, b->tables, sizeof(uint32_t) * (size_t)d.table_len, cudaMemcpyHostToDevice, st));
    if (d.total > 0 && d.C > 0) {
        dim3 grid((unsigned)std::max(1, std::min(64, (d.maxN + 255) / 256)), (unsigned)std::min(d.C, 65535));
        e->stage_begin(RSAC_STAGE_PACK);
        pack_pnp_kernel<<<grid, 256, 0, st>>>((const ProblemMeta*)s.d_metas.p, (const float*)s.d_p3d.p, (const float*)s.d_p2d.p,
                                              (const float*)s.d_sigma2.p, (const float*)s.d_th2.p, nullptr, 1,
                                              (float4*)s.d_cA.p, (float4*)s.d_cB.p, (float4*)s.d_uv.p, (float4*)s.d_cP.p, d.C);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
    }
    s.uploaded = true;
    return RSAC_OK;
}

// resident threads of the minimal solver: the staged early exit sizes its first stage from one wave
static int64_t mlpnp_wave_hyps(rsac_engine* e)
{
    static int per_sm = 0;
    if (per_sm == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mlpnp_minimal_range_kernel, 128, 0) != cudaSuccess) { cudaGetLastError(); nb = 2; }
        per_sm = std::max(1, nb) * 128;
    }
    return (int64_t)per_sm * e->sm_count;
}

// Stage boundaries (rsac_set_stages / rsac_set_phases / RSAC_MLPNP_EE_STAGES); automatic: a batch that fits one wave of the
// solver runs every hypothesis at once (a stage costs one thread latency whatever its size -- cfg2, 64 frames x 300, is a
// single partial wave), otherwise the first stage is three quarters of a wave and every further stage doubles
static std::vector<int> mlpnp_stage_bounds(rsac_engine* e, const BatchDims& d)
{
    const int64_t wave_h = mlpnp_wave_hyps(e);
    const int wave = (int)std::min<int64_t>(wave_h / std::max(d.C, 1), INT32_MAX);
    const int first = e->first_phase > 0 ? e->first_phase : (wave >= d.maxH ? d.maxH : std::max(16, (int)(wave_h * 3 / 4 / std::max(d.C, 1))));
    return early_stage_bounds(e, d, first, "RSAC_MLPNP_EE_STAGES");
}

static int mlpnp_launch_solve_range(rsac_engine* e, PnpState& s, const int32_t* list, const int32_t* list_count, int lo, int span, int64_t most)
{
    const BatchDims& d = s.d;
    const int64_t resident = mlpnp_wave_hyps(e) / 128;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((most + 127) / 128, resident));
    e->stage_begin(RSAC_STAGE_SOLVE);
    mlpnp_minimal_range_kernel<<<blocks, 128, 0, e->stream>>>((const ProblemMeta*)s.d_metas.p, d.C, list, list_count, lo, span,
                                                             (const uint32_t*)s.d_tables.p, (const float4*)s.d_cA.p, (const float4*)s.d_uv.p,
                                                             s.have_cov ? (const double*)s.d_cov.p : nullptr, (double*)s.d_poses.p);
    e->stage_end(RSAC_STAGE_SOLVE);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

static int mlpnp_select_setup(rsac_engine* e)
{
    const BatchDims& d = e->mlpnp.d;
    const size_t smem = (size_t)(3 * d.maxWords + 1) * 4 + 16;
    if (smem > 32 * 1024) RSAC_TRY(set_func_attr_max(e, (const void*)ransac_select_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    RSAC_CUDA(e, cudaFuncGetAttributes(&fa, ransac_select_kernel<1>));
    const size_t need = (fa.sharedSizeBytes + smem + 1024) * kSelectCtasPerSm;
    const int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024) + 2);
    RSAC_TRY(set_func_attr_max(e, (const void*)ransac_select_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    return RSAC_OK;
}

static int mlpnp_launch_select(rsac_engine* e, int flags, const int32_t* d_resume, void* d_results_out, int only_phase = -1)
{
    PnpState& s = e->mlpnp;
    const BatchDims& d = s.d;
    SelectArgs a;
    if (s.ee_mode) { a.ee = (int32_t*)s.d_ee.p; a.C = d.C; a.first_phase = s.ee_HA; a.only_phase = only_phase; }
    a.metas = (const ProblemMeta*)s.d_metas.p; a.cA = (const float4*)s.d_cA.p; a.cB = (const float4*)s.d_cB.p; a.cC = (const float4*)s.d_uv.p;
    a.poses = s.d_poses.p; a.counts = (const int32_t*)s.d_counts.p; a.cov = s.have_cov ? (const double*)s.d_cov.p : nullptr;
    a.sel = (uint32_t*)s.d_sel.p; a.pw_s = (double*)s.d_pw.p; a.us_s = nullptr; a.al_s = nullptr; a.tm_s = (double*)s.d_extra.p;
    a.results = s.d_results.p; a.results2 = d_results_out; a.masks = (uint32_t*)s.d_masks.p;
    a.problem_base = e->problem_base; a.flags = flags; a.resume = d_resume;
    a.problem_ids = (e->n_problem_ids == d.C && d.C > 0) ? (const int32_t*)e->d_problem_ids.p : nullptr;
    const size_t smem = (size_t)(3 * d.maxWords + 1) * 4 + 16;
    // kernel attributes: a no-op after the first call of a shape; inside a graph capture nothing is left to set
    // (early_run calls the set-up before it captures)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(e->stream, &cap);
    if (cap == cudaStreamCaptureStatusNone) RSAC_TRY(mlpnp_select_setup(e));
    e->stage_begin(RSAC_STAGE_SELECT);
    ransac_select_kernel<1><<<d.C, kSelectThreadsMlpnp, smem, e->stream>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

template <> struct EarlyHooks<1> {
    static int solve_range(rsac_engine* e, PnpState& s, const int32_t* list, const int32_t* list_count, int lo, int span, int64_t most)
    { return mlpnp_launch_solve_range(e, s, list, list_count, lo, span, most); }
    static int select(rsac_engine* e, int flags, const int32_t* d_resume, void* d_results_out, int only_phase)
    { return mlpnp_launch_select(e, flags, d_resume, d_results_out, only_phase); }
    static int setup(rsac_engine* e) { return mlpnp_select_setup(e); }
    static int stage0_hpl() { return 0; }
    static int stage_hpl() { return 0; }
    static int stage_chunk_words() { return 0; }
};

int rsac_mlpnp_rerun(rsac_engine* e, int flags, const int32_t* resume_from, void* d_results_out)
{
    if (!e || !resume_from) return RSAC_ERR_INVALID;
    PnpState& s = e->mlpnp;
    if (!s.ran) { e->err = "rsac_mlpnp_rerun before rsac_mlpnp_run"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    if (s.d.C == 0) return RSAC_OK;
    RSAC_TRY(e->d_resume.ensure(e, sizeof(int32_t) * (size_t)s.d.C));
    RSAC_CUDA(e, cudaMemcpyAsync(e->d_resume.p, resume_from, sizeof(int32_t) * (size_t)s.d.C, cudaMemcpyHostToDevice, e->stream));
    RSAC_TRY(early_complete<1>(e, s));     // the last run stopped early: compute what the resumed scan may need
    return mlpnp_launch_select(e, flags, (const int32_t*)e->d_resume.p, d_results_out);
}

int rsac_mlpnp_run(rsac_engine* e, int flags, void* d_results_out)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& s = e->mlpnp;
    if (!s.uploaded) { e->err = "rsac_mlpnp_run before rsac_mlpnp_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    const BatchDims& d = s.d;
    cudaStream_t st = e->stream;
    const ProblemMeta* metas = (const ProblemMeta*)s.d_metas.p;
    if (d.C == 0) { s.ran = true; return RSAC_OK; }
    const double* cov = s.have_cov ? (const double*)s.d_cov.p : nullptr;

    if (!s.have_tables && d.table_len > 0 && !s.tables_ready) {   // once per upload: the tables depend on the seeds only
        e->stage_begin(RSAC_STAGE_RNG);
        rng_tables_kernel<<<(d.C + kRngWarps - 1) / kRngWarps, kRngWarps * 32, 0, st>>>(metas, d.C, (uint32_t*)s.d_tables.p);
        e->stage_end(RSAC_STAGE_RNG);
        RSAC_CUDA(e, cudaGetLastError());
        s.tables_ready = true;
    }
    s.ee_mode = false;
    if ((flags & RSAC_FLAG_EARLY_EXIT) && d.sumH > 0) {
        const std::vector<int> bounds = mlpnp_stage_bounds(e, d);
        if (bounds.size() > 1) return early_run<1>(e, s, flags, d_results_out, bounds);
    }
    if (d.sumH > 0) {
        const int threads = 128;
        const unsigned blocks = (unsigned)((d.sumH + threads - 1) / threads);
        e->stage_begin(RSAC_STAGE_SOLVE);
        mlpnp_minimal_kernel<<<blocks, threads, 0, st>>>(metas, d.C, d.sumH, (const uint32_t*)s.d_tables.p,
                                                         (const float4*)s.d_cA.p, (const float4*)s.d_uv.p, cov, (double*)s.d_poses.p);
        e->stage_end(RSAC_STAGE_SOLVE);
        RSAC_CUDA(e, cudaGetLastError());

        ScoreArgs sa;
        RSAC_TRY(zero_score_region(e, s.d_counts, d.sumH, (int)s.groups.size(), sa));
        sa.metas = metas;
        sa.cP = (const float4*)s.d_cP.p; sa.cC = (const float4*)s.d_uv.p;
        sa.poses = s.d_poses.p;
        sa.hmasks = nullptr;
        int rc = launch_score<1>(e, sa, s.plan, (int)s.groups.size(), s.d_visit);
        if (rc) return rc;
    }
    {
        int rc = mlpnp_launch_select(e, flags, nullptr, d_results_out);
        if (rc) return rc;
    }
    s.ran = true;
    return RSAC_OK;
}

int rsac_mlpnp_phase_stats(rsac_engine* e, int64_t out[4])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    PnpState& s = e->mlpnp;
    out[0] = out[1] = out[2] = 0;
    out[3] = s.d.sumH;
    if (!s.ran) { e->err = "rsac_mlpnp_phase_stats before rsac_mlpnp_run"; return RSAC_ERR_STATE; }
    return early_stats(e, s, out);
}

int rsac_mlpnp_download(rsac_engine* e, rsac_result* results, uint32_t* masks)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& s = e->mlpnp;
    if (!s.ran) { e->err = "rsac_mlpnp_download before rsac_mlpnp_run"; return RSAC_ERR_STATE; }
    const BatchDims& d = s.d;
    if (results && d.C > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(results, s.d_results.p, sizeof(rsac_result) * d.C, cudaMemcpyDeviceToHost, e->stream));
    if (masks && d.total_words > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(masks, s.d_masks.p, sizeof(uint32_t) * (size_t)d.total_words, cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_mlpnp_solve(rsac_engine* e, const rsac_mlpnp_batch* b, int flags, rsac_result* results, uint32_t* masks)
{
    int rc = rsac_mlpnp_upload(e, b);
    if (rc) return rc;
    rc = rsac_mlpnp_run(e, flags, nullptr);
    if (rc) return rc;
    return rsac_mlpnp_download(e, results, masks);
}

int64_t rsac_mlpnp_total_hypotheses(rsac_engine* e) { return e ? e->mlpnp.d.sumH : 0; }

int rsac_mlpnp_get_hypotheses(rsac_engine* e, double* poses, int32_t* counts)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& s = e->mlpnp;
    if (!s.ran) return RSAC_ERR_STATE;
    if (s.d.sumH > 0) {
        if (poses) RSAC_CUDA(e, cudaMemcpyAsync(poses, s.d_poses.p, sizeof(double) * 12 * (size_t)s.d.sumH, cudaMemcpyDeviceToHost, e->stream));
        if (counts) RSAC_CUDA(e, cudaMemcpyAsync(counts, s.d_counts.p, sizeof(int32_t) * (size_t)s.d.sumH, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_debug_host_mlpnp6(const float K[4], const float p3d[18], const float p2d[12], const double* cov54, double R[9], double t[3])
{
    double f[18], pw[18];
    for (int i = 0; i < 6; ++i) {
        mlpnp_bearing(p2d[2 * i], p2d[2 * i + 1], K, f + 3 * i);
        for (int c = 0; c < 3; ++c) pw[3 * i + c] = (double)p3d[3 * i + c];
    }
    std::vector<double2> rec(kMaxSweepsRec * 66);
    mlpnp_compute_pose_small<6>(f, pw, cov54, R, t, rec.data());
    return RSAC_OK;
}

int rsac_debug_mlpnp_clocks(rsac_engine* e, long long out[8])
{
    if (!e || !out) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    RSAC_CUDA(e, cudaMemcpyFromSymbol(out, g_mlpnp_clocks, sizeof(long long) * 8));
    return RSAC_OK;
}
