// engine_sim3opt.cu -- C ABI for the batched Optimizer::OptimizeSim3 (include/ransac_b200.h).
#include "engine_shared.cuh"
#include "sim3opt.cuh"

static void sim3opt_fill_meta(Sim3OptMeta& m, int64_t off, int n, const float* K1, const float* K2, const float* S12, float th2, int fix)
{
    m.off = off;
    m.n = n;
    m.fix_scale = fix;
    m.th2 = th2;
    for (int k = 0; k < 4; ++k) { m.K1[k] = K1[k]; m.K2[k] = K2[k]; }
    for (int k = 0; k < 9; ++k) m.R12[k] = S12[k];
    for (int k = 0; k < 3; ++k) m.t12[k] = S12[9 + k];
    m.s12 = S12[12];
}

int rsac_sim3opt_upload(rsac_engine* e, const rsac_sim3opt_batch* b)
{
    if (!e || !b || b->C < 0 || !b->offsets || !b->K1 || !b->K2 || !b->S12 || !b->th2) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    Sim3OptState& s = e->sim3opt;
    s.uploaded = false; s.ran = false; s.chained = false;
    const int C = b->C;
    const int64_t total = b->offsets[C];
    if (total < 0 || (total > 0 && (!b->x1c || !b->x2c || !b->obs1 || !b->obs2 || !b->inv_sigma2_1 || !b->inv_sigma2_2))) return RSAC_ERR_INVALID;
    Sim3OptMeta* hm = (Sim3OptMeta*)s.h_metas.ensure(sizeof(Sim3OptMeta) * (size_t)std::max(C, 1));
    if (!hm) { e->err = "pinned allocation failed"; return RSAC_ERR_ALLOC; }
    for (int c = 0; c < C; ++c) {
        const int n = b->offsets[c + 1] - b->offsets[c];
        if (n < 0) { e->err = "bad offsets"; return RSAC_ERR_INVALID; }
        sim3opt_fill_meta(hm[c], b->offsets[c], n, b->K1 + 4 * c, b->K2 + 4 * c, b->S12 + 13 * c, b->th2[c], b->fix_scale ? b->fix_scale[c] : 1);
    }
    s.C = C;
    s.total = total;
    const size_t tot = (size_t)std::max<int64_t>(total, 1);
    RSAC_TRY(s.d_metas.ensure(e, sizeof(Sim3OptMeta) * (size_t)std::max(C, 1)));
    RSAC_TRY(s.d_x1.ensure(e, tot * 12));
    RSAC_TRY(s.d_x2.ensure(e, tot * 12));
    RSAC_TRY(s.d_o1.ensure(e, tot * 8));
    RSAC_TRY(s.d_o2.ensure(e, tot * 8));
    RSAC_TRY(s.d_is1.ensure(e, tot * 4));
    RSAC_TRY(s.d_is2.ensure(e, tot * 4));
    RSAC_TRY(s.d_removed.ensure(e, tot));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_sim3opt_result) * (size_t)std::max(C, 1)));
    cudaStream_t st = e->stream;
    if (C > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_metas.p, hm, sizeof(Sim3OptMeta) * (size_t)C, cudaMemcpyHostToDevice, st));
        s.h_metas.mark(st);
    }
    if (total > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_x1.p, b->x1c, (size_t)total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_x2.p, b->x2c, (size_t)total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_o1.p, b->obs1, (size_t)total * 8, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_o2.p, b->obs2, (size_t)total * 8, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_is1.p, b->inv_sigma2_1, (size_t)total * 4, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_is2.p, b->inv_sigma2_2, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    }
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_sim3opt_run(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    Sim3OptState& s = e->sim3opt;
    if (!s.uploaded) { e->err = "rsac_sim3opt_run before rsac_sim3opt_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    if (s.C > 0) {
        // large batches: one warp per pair, four pairs per CTA; otherwise (a loop closure hands over a few candidates)
        // one pair per CTA with four warps on its matches
        const bool wide = s.C < 4 * e->sm_count;
        const size_t smem = sizeof(double) * (size_t)(wide ? so::sim_smem_doubles(128) : kSim3OptWarps * so::sim_smem_doubles(32));
        const void* kern = wide ? (const void*)sim3opt_kernel<128> : (const void*)sim3opt_kernel<32>;
        if (smem > 48 * 1024) RSAC_TRY(set_func_attr_max(e, kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        e->stage_begin(RSAC_STAGE_SELECT);
        if (wide)
            sim3opt_kernel<128><<<s.C, 128, smem, e->stream>>>(
                (const Sim3OptMeta*)s.d_metas.p, s.C, (const float*)s.d_x1.p, (const float*)s.d_x2.p, (const float*)s.d_o1.p,
                (const float*)s.d_o2.p, (const float*)s.d_is1.p, (const float*)s.d_is2.p, (uint8_t*)s.d_removed.p,
                (rsac_sim3opt_result*)s.d_results.p, e->problem_base);
        else
            sim3opt_kernel<32><<<(s.C + kSim3OptWarps - 1) / kSim3OptWarps, kSim3OptWarps * 32, smem, e->stream>>>(
                (const Sim3OptMeta*)s.d_metas.p, s.C, (const float*)s.d_x1.p, (const float*)s.d_x2.p, (const float*)s.d_o1.p,
                (const float*)s.d_o2.p, (const float*)s.d_is1.p, (const float*)s.d_is2.p, (uint8_t*)s.d_removed.p,
                (rsac_sim3opt_result*)s.d_results.p, e->problem_base);
        e->stage_end(RSAC_STAGE_SELECT);
        RSAC_CUDA(e, cudaGetLastError());
    }
    s.ran = true;
    return RSAC_OK;
}

int rsac_sim3opt_download(rsac_engine* e, rsac_sim3opt_result* results, uint8_t* removed)
{
    if (!e) return RSAC_ERR_INVALID;
    Sim3OptState& s = e->sim3opt;
    if (!s.ran) { e->err = "rsac_sim3opt_download before rsac_sim3opt_run"; return RSAC_ERR_STATE; }
    if (results && s.C > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(results, s.d_results.p, sizeof(rsac_sim3opt_result) * (size_t)s.C, cudaMemcpyDeviceToHost, e->stream));
    if (removed && s.total > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(removed, s.d_removed.p, (size_t)s.total, cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_sim3opt_solve(rsac_engine* e, const rsac_sim3opt_batch* b, rsac_sim3opt_result* results, uint8_t* removed)
{
    int rc = rsac_sim3opt_upload(e, b);
    if (rc) return rc;
    rc = rsac_sim3opt_run(e);
    if (rc) return rc;
    return rsac_sim3opt_download(e, results, removed);
}

int rsac_debug_host_sim3opt(int n, const float* x1c, const float* x2c, const float* obs1, const float* obs2,
                            const float* inv_sigma2_1, const float* inv_sigma2_2, const float K1[4], const float K2[4],
                            const float S12[13], float th2, int fix_scale, rsac_sim3opt_result* result, uint8_t* removed)
{
    if (n < 0 || !K1 || !K2 || !S12 || !result) return RSAC_ERR_INVALID;
    Sim3OptMeta m;
    sim3opt_fill_meta(m, 0, n, K1, K2, S12, th2, fix_scale);
    std::vector<double> scratch((size_t)so::sim_smem_doubles(32));
    so::optimize_sim3<1>(m, x1c, x2c, obs1, obs2, inv_sigma2_1, inv_sigma2_2, removed, 0, scratch.data(), 0, result);
    return RSAC_OK;
}
