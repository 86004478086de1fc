// engine_poseopt.cu -- C ABI for the batched Optimizer::PoseOptimization (include/ransac_b200.h).
#include "engine_shared.cuh"
#include "poseopt.cuh"

int rsac_poseopt_upload(rsac_engine* e, const rsac_poseopt_batch* b)
{
    if (!e || !b || b->C < 0 || !b->offsets || !b->K || !b->Tcw) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    PoseOptState& s = e->poseopt;
    s.uploaded = false; s.ran = false;
    const int C = b->C;
    const int64_t total = b->offsets[C];
    if (total < 0 || (total > 0 && (!b->p3d || !b->obs || !b->inv_sigma2))) return RSAC_ERR_INVALID;
    PoseOptMeta* hm = (PoseOptMeta*)s.h_metas.ensure(sizeof(PoseOptMeta) * (size_t)std::max(C, 1));
    if (!hm) { e->err = "pinned allocation failed"; return RSAC_ERR_ALLOC; }
    for (int c = 0; c < C; ++c) {
        PoseOptMeta& m = hm[c];
        m.off = b->offsets[c];
        m.n = b->offsets[c + 1] - b->offsets[c];
        if (m.n < 0) { e->err = "bad offsets"; return RSAC_ERR_INVALID; }
        for (int k = 0; k < 5; ++k) m.K[k] = b->K[5 * c + k];
        for (int k = 0; k < 9; ++k) m.Rcw[k] = b->Tcw[12 * c + k];
        for (int k = 0; k < 3; ++k) m.tcw[k] = b->Tcw[12 * c + 9 + k];
    }
    s.C = C;
    s.chained = false;
    s.total = total;
    const size_t tot = (size_t)std::max<int64_t>(total, 1);
    RSAC_TRY(s.d_metas.ensure(e, sizeof(PoseOptMeta) * (size_t)std::max(C, 1)));
    RSAC_TRY(s.d_p3d.ensure(e, tot * 12));
    RSAC_TRY(s.d_obs.ensure(e, tot * 12));
    RSAC_TRY(s.d_isig.ensure(e, tot * 4));
    RSAC_TRY(s.d_outlier.ensure(e, tot));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_poseopt_result) * (size_t)std::max(C, 1)));
    cudaStream_t st = e->stream;
    if (C > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_metas.p, hm, sizeof(PoseOptMeta) * (size_t)C, cudaMemcpyHostToDevice, st));
        s.h_metas.mark(st);
    }
    if (total > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_p3d.p, b->p3d, (size_t)total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_obs.p, b->obs, (size_t)total * 12, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_isig.p, b->inv_sigma2, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    }
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_poseopt_from_pnp(rsac_engine* e, float bf)
{
    if (!e) return RSAC_ERR_INVALID;
    PnpState& p = e->pnp;
    if (!p.ran) { e->err = "rsac_poseopt_from_pnp before rsac_pnp_run"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    PoseOptState& s = e->poseopt;
    s.uploaded = false; s.ran = false;
    const int C = p.d.C;
    const int64_t total = p.d.total;
    s.C = C;
    s.total = total;
    s.chained = true;
    const size_t tot = (size_t)std::max<int64_t>(total, 1);
    RSAC_TRY(s.d_metas.ensure(e, sizeof(PoseOptMeta) * (size_t)std::max(C, 1)));
    RSAC_TRY(s.d_p3d.ensure(e, tot * 12));
    RSAC_TRY(s.d_obs.ensure(e, tot * 12));
    RSAC_TRY(s.d_isig.ensure(e, tot * 4));
    RSAC_TRY(s.d_outlier.ensure(e, tot));
    RSAC_TRY(s.d_src.ensure(e, tot * 4));
    RSAC_TRY(s.d_full.ensure(e, tot));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_poseopt_result) * (size_t)std::max(C, 1)));
    if (C > 0) {
        RSAC_TRY(rsac_internal_pnp_ensure_flat(e));      // an indexed batch keeps no flat arrays until somebody asks
        e->stage_begin(RSAC_STAGE_PACK);
        poseopt_from_pnp_kernel<<<C, 128, 0, e->stream>>>((const ProblemMeta*)p.d_metas.p, C, (const rsac_result*)p.d_results.p,
                                                          (const uint32_t*)p.d_masks.p, (const float*)p.d_p3d.p, (const float*)p.d_p2d.p,
                                                          (const float*)p.d_sigma2.p, bf, (PoseOptMeta*)s.d_metas.p, (float*)s.d_p3d.p,
                                                          (float*)s.d_obs.p, (float*)s.d_isig.p, (int32_t*)s.d_src.p);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
    }
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_poseopt_run(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    PoseOptState& s = e->poseopt;
    if (!s.uploaded) { e->err = "rsac_poseopt_run before rsac_poseopt_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    if (s.C > 0) {
        // large batches: one warp per frame, four frames per CTA (throughput); batches that cannot give every SM four
        // CTAs: one frame per CTA with four warps on its edges (latency -- a relocalisation hands over a few frames)
        const bool wide = s.C < 4 * e->sm_count;
        e->stage_begin(RSAC_STAGE_SELECT);
        if (wide)
            poseopt_kernel<128><<<s.C, 128, sizeof(double) * po::red_doubles(128), e->stream>>>(
                (const PoseOptMeta*)s.d_metas.p, s.C, (const float*)s.d_p3d.p, (const float*)s.d_obs.p, (const float*)s.d_isig.p,
                (uint8_t*)s.d_outlier.p, (rsac_poseopt_result*)s.d_results.p, e->problem_base);
        else
            poseopt_kernel<32><<<(s.C + kPoseOptWarps - 1) / kPoseOptWarps, kPoseOptWarps * 32,
                                 sizeof(double) * kPoseOptWarps * po::red_doubles(32), e->stream>>>(
                (const PoseOptMeta*)s.d_metas.p, s.C, (const float*)s.d_p3d.p, (const float*)s.d_obs.p, (const float*)s.d_isig.p,
                (uint8_t*)s.d_outlier.p, (rsac_poseopt_result*)s.d_results.p, e->problem_base);
        e->stage_end(RSAC_STAGE_SELECT);
        RSAC_CUDA(e, cudaGetLastError());
        if (s.chained) {
            e->stage_begin(RSAC_STAGE_PACK);
            poseopt_scatter_flags_kernel<<<s.C, 128, 0, e->stream>>>((const ProblemMeta*)e->pnp.d_metas.p, s.C, (const PoseOptMeta*)s.d_metas.p,
                                                                     (const uint8_t*)s.d_outlier.p, (const int32_t*)s.d_src.p, (uint8_t*)s.d_full.p);
            e->stage_end(RSAC_STAGE_PACK);
            RSAC_CUDA(e, cudaGetLastError());
        }
    }
    s.ran = true;
    return RSAC_OK;
}

int rsac_poseopt_download(rsac_engine* e, rsac_poseopt_result* results, uint8_t* outlier)
{
    if (!e) return RSAC_ERR_INVALID;
    PoseOptState& s = e->poseopt;
    if (!s.ran) { e->err = "rsac_poseopt_download before rsac_poseopt_run"; return RSAC_ERR_STATE; }
    if (results && s.C > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(results, s.d_results.p, sizeof(rsac_poseopt_result) * (size_t)s.C, cudaMemcpyDeviceToHost, e->stream));
    if (outlier && s.total > 0)
        RSAC_CUDA(e, cudaMemcpyAsync(outlier, s.chained ? s.d_full.p : s.d_outlier.p, (size_t)s.total, cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_poseopt_solve(rsac_engine* e, const rsac_poseopt_batch* b, rsac_poseopt_result* results, uint8_t* outlier)
{
    int rc = rsac_poseopt_upload(e, b);
    if (rc) return rc;
    rc = rsac_poseopt_run(e);
    if (rc) return rc;
    return rsac_poseopt_download(e, results, outlier);
}

int rsac_debug_host_poseopt(int n, const float* p3d, const float* obs, const float* inv_sigma2, const float K[5],
                            const float Tcw[12], rsac_poseopt_result* result, uint8_t* outlier)
{
    if (n < 0 || !K || !Tcw || !result) return RSAC_ERR_INVALID;
    PoseOptMeta m;
    m.off = 0;
    m.n = n;
    for (int k = 0; k < 5; ++k) m.K[k] = K[k];
    for (int k = 0; k < 9; ++k) m.Rcw[k] = Tcw[k];
    for (int k = 0; k < 3; ++k) m.tcw[k] = Tcw[9 + k];
    po::pose_optimization<1>(m, p3d, obs, inv_sigma2, outlier, 0, nullptr, 0, result);
    return RSAC_OK;
}
