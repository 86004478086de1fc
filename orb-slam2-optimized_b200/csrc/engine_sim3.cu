// engine_sim3.cu -- C ABI for the batched Sim3Solver (include/ransac_b200.h, "Sim3Solver").
#include "engine_shared.cuh"
#include "sim3.cuh"

int rsac_sim3_upload(rsac_engine* e, const rsac_sim3_batch* b)
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
    if (d.total > 0) {
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
