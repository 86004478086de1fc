// engine_guided.cu -- C ABI for the batched ORBmatcher::SearchBySim3 (include/ransac_b200.h).
#include "engine_shared.cuh"
#include "guided.cuh"
#include "sim3opt.cuh"

// keyframe / frame views -> concatenated device arrays (shared by rsac_sim3_search_* and rsac_proj_search_*)
static int guided_upload_views(rsac_engine* e, int V, const rsac_kf_view* in, bool need_angle, std::vector<KfViewDev>& views)
{
    GuidedState& s = e->guided;
    views.assign(std::max(V, 1), KfViewDev());
    int64_t nfeat = 0, ngoff = 0, ngidx = 0;
    for (int i = 0; i < V; ++i) {
        const rsac_kf_view& v = in[i];
        const int64_t cells = (int64_t)v.grid_cols * v.grid_rows;
        if (v.n_feat < 0 || v.grid_cols <= 0 || v.grid_rows <= 0 || cells > (1 << 20) || !v.grid_off || v.n_levels < 1 || v.n_levels > kGuidedMaxLevels ||
            !v.scale_factors || (v.n_feat > 0 && (!v.kp_xy || !v.kp_octave || !v.desc || !v.mp_valid || !v.mp_xyz || !v.mp_desc || !v.mp_maxdist || !v.mp_mindist)) ||
            (need_angle && v.n_feat > 0 && !v.kp_angle) || v.n_feat >= (1 << 20)) {
            e->err = "bad keyframe view"; return RSAC_ERR_INVALID;
        }
        if (v.grid_off[0] != 0 || v.grid_off[cells] < 0 || (v.grid_off[cells] > 0 && !v.grid_idx)) { e->err = "bad keyframe grid"; return RSAC_ERR_INVALID; }
        for (int64_t k = 0; k < cells; ++k)
            if (v.grid_off[k + 1] < v.grid_off[k]) { e->err = "grid_off must ascend"; return RSAC_ERR_INVALID; }
        for (int k = 0; k < v.grid_off[cells]; ++k)
            if (v.grid_idx[k] < 0 || v.grid_idx[k] >= v.n_feat) { e->err = "grid index out of range"; return RSAC_ERR_INVALID; }
        KfViewDev& d = views[i];
        memset(&d, 0, sizeof(d));
        d.feat_off = (int32_t)nfeat; d.n_feat = v.n_feat;
        d.goff_off = (int32_t)ngoff; d.gidx_off = (int32_t)ngidx;
        d.grid_cols = v.grid_cols; d.grid_rows = v.grid_rows; d.n_levels = v.n_levels;
        d.grid_w_inv = v.grid_w_inv; d.grid_h_inv = v.grid_h_inv; d.log_scale_factor = v.log_scale_factor;
        memcpy(d.Rcw, v.Rcw, sizeof(d.Rcw)); memcpy(d.tcw, v.tcw, sizeof(d.tcw)); memcpy(d.bounds, v.bounds, sizeof(d.bounds));
        for (int k = 0; k < v.n_levels; ++k) d.scale_factors[k] = v.scale_factors[k];
        nfeat += v.n_feat; ngoff += cells + 1; ngidx += v.grid_off[cells];
        if (nfeat > INT32_MAX / 8 || ngoff > INT32_MAX || ngidx > INT32_MAX) { e->err = "batch too large"; return RSAC_ERR_INVALID; }
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t nf = (size_t)std::max<int64_t>(nfeat, 1);
    struct Seg { DevBuf* d; size_t bytes, off; };
    Seg seg[] = {{&s.d_views, sizeof(KfViewDev) * views.size(), 0}, {&s.d_kp_xy, 8 * nf, 0}, {&s.d_kp_octave, 4 * nf, 0}, {&s.d_desc, 32 * nf, 0},
                 {&s.d_mp_valid, nf, 0}, {&s.d_mp_xyz, 12 * nf, 0}, {&s.d_mp_desc, 32 * nf, 0}, {&s.d_mp_maxdist, 4 * nf, 0},
                 {&s.d_mp_mindist, 4 * nf, 0}, {&s.d_grid_off, 4 * (size_t)std::max<int64_t>(ngoff, 1), 0},
                 {&s.d_grid_idx, 4 * (size_t)std::max<int64_t>(ngidx, 1), 0}, {&s.d_kp_angle, 4 * nf, 0}};
    size_t total = 0;
    for (auto& g : seg) { g.off = total; total = al(total + g.bytes); }
    char* h = (char*)s.h_stage.ensure(total);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    memset(h, 0, total);
    memcpy(h + seg[0].off, views.data(), sizeof(KfViewDev) * views.size());
    for (int i = 0; i < V; ++i) {
        const rsac_kf_view& v = in[i];
        const KfViewDev& d = views[i];
        const size_t f = (size_t)d.feat_off, n = (size_t)v.n_feat;
        const int64_t cells = (int64_t)v.grid_cols * v.grid_rows;
        if (n > 0) {
            memcpy(h + seg[1].off + 8 * f, v.kp_xy, 8 * n); memcpy(h + seg[2].off + 4 * f, v.kp_octave, 4 * n);
            memcpy(h + seg[3].off + 32 * f, v.desc, 32 * n); memcpy(h + seg[4].off + f, v.mp_valid, n);
            memcpy(h + seg[5].off + 12 * f, v.mp_xyz, 12 * n); memcpy(h + seg[6].off + 32 * f, v.mp_desc, 32 * n);
            memcpy(h + seg[7].off + 4 * f, v.mp_maxdist, 4 * n); memcpy(h + seg[8].off + 4 * f, v.mp_mindist, 4 * n);
            if (v.kp_angle) memcpy(h + seg[11].off + 4 * f, v.kp_angle, 4 * n);
        }
        memcpy(h + seg[9].off + 4 * (size_t)d.goff_off, v.grid_off, 4 * (size_t)(cells + 1));
        if (v.grid_off[cells] > 0) memcpy(h + seg[10].off + 4 * (size_t)d.gidx_off, v.grid_idx, 4 * (size_t)v.grid_off[cells]);
    }
    for (auto& g : seg) {
        RSAC_TRY(g.d->ensure(e, g.bytes));
        RSAC_CUDA(e, cudaMemcpyAsync(g.d->p, h + g.off, g.bytes, cudaMemcpyHostToDevice, e->stream));
    }
    s.h_stage.mark(e->stream);
    s.n_views = V;
    s.views_have_angle = true;
    for (int i = 0; i < V; ++i) s.views_have_angle = s.views_have_angle && (in[i].kp_angle != nullptr || in[i].n_feat == 0);
    s.view_n_feat.assign(V, 0); s.view_feat_off.assign(V, 0);
    s.view_mp_valid.assign((size_t)nfeat, 0);
    for (int i = 0; i < V; ++i) {
        s.view_n_feat[i] = views[i].n_feat; s.view_feat_off[i] = views[i].feat_off;
        if (in[i].n_feat > 0) memcpy(s.view_mp_valid.data() + views[i].feat_off, in[i].mp_valid, (size_t)in[i].n_feat);
    }
    return RSAC_OK;
}

// the views of a batch: uploaded when given, else the resident ones (n_views == 0 && views == NULL)
static int guided_views_for_batch(rsac_engine* e, int n_views, const rsac_kf_view* in, bool need_angle, std::vector<KfViewDev>& views)
{
    GuidedState& s = e->guided;
    if (n_views == 0 && !in && s.n_views > 0) {
        if (need_angle && !s.views_have_angle) { e->err = "the resident views carry no keypoint angles"; return RSAC_ERR_STATE; }
        views.assign(s.n_views, KfViewDev());
        for (int i = 0; i < s.n_views; ++i) { views[i].n_feat = s.view_n_feat[i]; views[i].feat_off = s.view_feat_off[i]; }
        return RSAC_OK;
    }
    return guided_upload_views(e, n_views, in, need_angle, views);
}

int rsac_views_upload(rsac_engine* e, int n_views, const rsac_kf_view* views)
{
    if (!e || n_views < 0 || (n_views > 0 && !views)) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    GuidedState& s = e->guided;
    s.uploaded = false; s.ran = false; s.proj_uploaded = false; s.proj_ran = false;
    std::vector<KfViewDev> tmp;
    return guided_upload_views(e, n_views, views, false, tmp);
}

int rsac_sim3_search_upload(rsac_engine* e, const rsac_sim3_search_batch* b)
{
    if (!e || !b || b->n_views < 0 || b->C < 0 || (b->n_views > 0 && !b->views)) return RSAC_ERR_INVALID;
    if (b->C > 0 && (!b->kf1 || !b->kf2 || !b->K || !b->R12 || !b->t12)) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    GuidedState& s = e->guided;
    s.uploaded = false; s.ran = false; s.proj_uploaded = false; s.proj_ran = false;
    const int C = b->C;
    std::vector<KfViewDev> views;
    RSAC_TRY(guided_views_for_batch(e, b->n_views, b->views, false, views));
    const int V = (int)s.n_views;
    std::vector<int64_t> off1(C + 1, 0), off2(C + 1, 0);
    s.maxN1 = 0; s.maxN = 0;
    for (int c = 0; c < C; ++c) {
        const int a = b->kf1[c], q = b->kf2[c];
        if (a < 0 || a >= V || q < 0 || q >= V) { e->err = "view index out of range"; return RSAC_ERR_INVALID; }
        off1[c + 1] = off1[c] + views[a].n_feat;
        off2[c + 1] = off2[c] + views[q].n_feat;
        s.maxN1 = std::max(s.maxN1, views[a].n_feat);
        s.maxN = std::max(s.maxN, std::max(views[a].n_feat, views[q].n_feat));
    }
    s.C = C; s.total1 = off1[C]; s.total2 = off2[C]; s.th = b->th;
    s.have_matched = b->matched12_in != nullptr && s.total1 > 0;
    s.have_scale = b->s12 != nullptr;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t c1 = (size_t)std::max(C, 1);
    struct Seg { DevBuf* d; size_t bytes, off; };
    Seg seg[] = {{&s.d_kf1, 4 * c1, 0}, {&s.d_kf2, 4 * c1, 0}, {&s.d_K, 16 * c1, 0}, {&s.d_R12, 36 * c1, 0}, {&s.d_t12, 12 * c1, 0}, {&s.d_s12, 4 * c1, 0},
                 {&s.d_off1, 8 * (c1 + 1), 0}, {&s.d_off2, 8 * (c1 + 1), 0}, {&s.d_matched_in, 4 * (size_t)std::max<int64_t>(s.total1, 1), 0}};
    size_t total = 0;
    for (auto& g : seg) { g.off = total; total = al(total + g.bytes); }
    char* h = (char*)s.h_stage2.ensure(total);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    memset(h, 0, total);
    if (C > 0) {
        memcpy(h + seg[0].off, b->kf1, 4 * (size_t)C); memcpy(h + seg[1].off, b->kf2, 4 * (size_t)C);
        memcpy(h + seg[2].off, b->K, 16 * (size_t)C); memcpy(h + seg[3].off, b->R12, 36 * (size_t)C); memcpy(h + seg[4].off, b->t12, 12 * (size_t)C);
        if (b->s12) memcpy(h + seg[5].off, b->s12, 4 * (size_t)C);
    }
    memcpy(h + seg[6].off, off1.data(), 8 * (size_t)(C + 1));
    memcpy(h + seg[7].off, off2.data(), 8 * (size_t)(C + 1));
    if (s.have_matched) memcpy(h + seg[8].off, b->matched12_in, 4 * (size_t)s.total1);
    for (auto& g : seg) {
        RSAC_TRY(g.d->ensure(e, g.bytes));
        RSAC_CUDA(e, cudaMemcpyAsync(g.d->p, h + g.off, g.bytes, cudaMemcpyHostToDevice, e->stream));
    }
    s.h_stage2.mark(e->stream);
    const size_t t1 = (size_t)std::max<int64_t>(s.total1, 1), t2 = (size_t)std::max<int64_t>(s.total2, 1);
    RSAC_TRY(s.d_already1.ensure(e, t1)); RSAC_TRY(s.d_already2.ensure(e, t2));
    RSAC_TRY(s.d_m1.ensure(e, 4 * t1)); RSAC_TRY(s.d_m2.ensure(e, 4 * t2));
    RSAC_TRY(s.d_match12.ensure(e, 4 * t1)); RSAC_TRY(s.d_n_found.ensure(e, 4 * c1));
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_sim3_search_run(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    GuidedState& s = e->guided;
    if (!s.uploaded) { e->err = "rsac_sim3_search_run before rsac_sim3_search_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    s.ran = true;
    if (s.C == 0) return RSAC_OK;
    Sim3SearchArgs a;
    a.views = (const KfViewDev*)s.d_views.p; a.kp_xy = (const float*)s.d_kp_xy.p; a.kp_octave = (const int32_t*)s.d_kp_octave.p; a.kp_angle = (const float*)s.d_kp_angle.p;
    a.desc = (const uint32_t*)s.d_desc.p; a.mp_valid = (const uint8_t*)s.d_mp_valid.p; a.mp_xyz = (const float*)s.d_mp_xyz.p;
    a.mp_desc = (const uint32_t*)s.d_mp_desc.p; a.mp_maxdist = (const float*)s.d_mp_maxdist.p; a.mp_mindist = (const float*)s.d_mp_mindist.p;
    a.grid_off = (const int32_t*)s.d_grid_off.p; a.grid_idx = (const int32_t*)s.d_grid_idx.p;
    a.C = s.C; a.kf1 = (const int32_t*)s.d_kf1.p; a.kf2 = (const int32_t*)s.d_kf2.p; a.K = (const float*)s.d_K.p;
    a.R12 = (const float*)s.d_R12.p; a.t12 = (const float*)s.d_t12.p; a.s12 = s.have_scale ? (const float*)s.d_s12.p : nullptr; a.th = s.th;
    a.off1 = (const int64_t*)s.d_off1.p; a.off2 = (const int64_t*)s.d_off2.p;
    a.matched_in = s.have_matched ? (const int32_t*)s.d_matched_in.p : nullptr;
    a.already1 = (uint8_t*)s.d_already1.p; a.already2 = (uint8_t*)s.d_already2.p; a.m1 = (int32_t*)s.d_m1.p; a.m2 = (int32_t*)s.d_m2.p;
    a.match12 = (int32_t*)s.d_match12.p; a.n_found = (int32_t*)s.d_n_found.p;
    RSAC_CUDA(e, cudaMemsetAsync(s.d_already2.p, 0, (size_t)std::max<int64_t>(s.total2, 1), st));
    RSAC_CUDA(e, cudaMemsetAsync(s.d_n_found.p, 0, 4 * (size_t)s.C, st));
    const dim3 g1((unsigned)std::max(1, (s.maxN1 + 255) / 256), (unsigned)s.C);
    const dim3 gs((unsigned)std::max(1, (s.maxN + 127) / 128), (unsigned)(2 * s.C));
    e->stage_begin(RSAC_STAGE_PACK);
    sim3_search_prepare_kernel<<<g1, 256, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_PACK);
    e->stage_begin(RSAC_STAGE_SOLVE);
    sim3_search_kernel<<<gs, 128, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SOLVE);
    e->stage_begin(RSAC_STAGE_SELECT);
    sim3_search_agree_kernel<<<g1, 256, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

int rsac_sim3_search_download(rsac_engine* e, int32_t* match12, int32_t* n_found)
{
    if (!e) return RSAC_ERR_INVALID;
    GuidedState& s = e->guided;
    if (!s.ran) { e->err = "rsac_sim3_search_download before rsac_sim3_search_run"; return RSAC_ERR_STATE; }
    if (s.C > 0) {
        if (match12 && s.total1 > 0) RSAC_CUDA(e, cudaMemcpyAsync(match12, s.d_match12.p, 4 * (size_t)s.total1, cudaMemcpyDeviceToHost, e->stream));
        if (n_found) RSAC_CUDA(e, cudaMemcpyAsync(n_found, s.d_n_found.p, 4 * (size_t)s.C, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_sim3_search(rsac_engine* e, const rsac_sim3_search_batch* b, int32_t* match12, int32_t* n_found)
{
    int rc = rsac_sim3_search_upload(e, b);
    if (rc) return rc;
    rc = rsac_sim3_search_run(e);
    if (rc) return rc;
    return rsac_sim3_search_download(e, match12, n_found);
}

// ---------------------------------------------------------------- ORBmatcher::SearchByProjection(Frame, KeyFrame, ...)
int rsac_proj_search_upload(rsac_engine* e, const rsac_proj_search_batch* b)
{
    if (!e || !b || b->n_views < 0 || b->C < 0 || (b->n_views > 0 && !b->views)) return RSAC_ERR_INVALID;
    if (b->C > 0 && (!b->frame || !b->kf || !b->K || !b->Rcw || !b->tcw)) return RSAC_ERR_INVALID;
    if (b->orb_dist < 0 || b->orb_dist > 256) { e->err = "ORBdist must be in [0, 256]"; return RSAC_ERR_INVALID; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    GuidedState& s = e->guided;
    s.uploaded = false; s.ran = false; s.proj_uploaded = false; s.proj_ran = false;
    const int C = b->C;
    std::vector<KfViewDev> views;
    RSAC_TRY(guided_views_for_batch(e, b->n_views, b->views, b->check_orientation != 0, views));
    const int V = (int)s.n_views;
    std::vector<int64_t> offF(C + 1, 0), offK(C + 1, 0);
    s.maxN = 0;
    for (int c = 0; c < C; ++c) {
        const int f = b->frame[c], k = b->kf[c];
        if (f < 0 || f >= V || k < 0 || k >= V) { e->err = "view index out of range"; return RSAC_ERR_INVALID; }
        offF[c + 1] = offF[c] + views[f].n_feat;
        offK[c + 1] = offK[c] + views[k].n_feat;
        s.maxN = std::max(s.maxN, views[k].n_feat);
    }
    s.proj_C = C; s.totalF = offF[C]; s.totalK = offK[C]; s.th = b->th; s.orb_dist = b->orb_dist; s.check_orientation = b->check_orientation;
    s.have_occupied = b->occupied != nullptr && s.totalF > 0;
    s.have_found = b->already_found != nullptr && s.totalK > 0;
    s.cap = std::max(1, std::min(32, env_int("RSAC_PROJ_CAP", 16)));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t c1 = (size_t)std::max(C, 1), tF = (size_t)std::max<int64_t>(s.totalF, 1), tK = (size_t)std::max<int64_t>(s.totalK, 1);
    struct Seg { DevBuf* d; size_t bytes, off; };
    Seg seg[] = {{&s.d_kf1, 4 * c1, 0}, {&s.d_kf2, 4 * c1, 0}, {&s.d_K, 16 * c1, 0}, {&s.d_R12, 36 * c1, 0}, {&s.d_t12, 12 * c1, 0},
                 {&s.d_off1, 8 * (c1 + 1), 0}, {&s.d_off2, 8 * (c1 + 1), 0}, {&s.d_already1, tF, 0}, {&s.d_already2, tK, 0}};
    size_t total = 0;
    for (auto& g : seg) { g.off = total; total = al(total + g.bytes); }
    char* h = (char*)s.h_stage2.ensure(total);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    memset(h, 0, total);
    if (C > 0) {
        memcpy(h + seg[0].off, b->frame, 4 * (size_t)C); memcpy(h + seg[1].off, b->kf, 4 * (size_t)C);
        memcpy(h + seg[2].off, b->K, 16 * (size_t)C); memcpy(h + seg[3].off, b->Rcw, 36 * (size_t)C); memcpy(h + seg[4].off, b->tcw, 12 * (size_t)C);
    }
    memcpy(h + seg[5].off, offF.data(), 8 * (size_t)(C + 1));
    memcpy(h + seg[6].off, offK.data(), 8 * (size_t)(C + 1));
    if (s.have_occupied) memcpy(h + seg[7].off, b->occupied, (size_t)s.totalF);
    if (s.have_found) memcpy(h + seg[8].off, b->already_found, (size_t)s.totalK);
    for (auto& g : seg) {
        RSAC_TRY(g.d->ensure(e, g.bytes));
        RSAC_CUDA(e, cudaMemcpyAsync(g.d->p, h + g.off, g.bytes, cudaMemcpyHostToDevice, e->stream));
    }
    s.h_stage2.mark(e->stream);
    RSAC_TRY(s.d_cand.ensure(e, 4 * tK * (size_t)s.cap)); RSAC_TRY(s.d_cand_n.ensure(e, 4 * tK));
    RSAC_TRY(s.d_taken.ensure(e, tF)); RSAC_TRY(s.d_minidx.ensure(e, 4 * tF)); RSAC_TRY(s.d_match12.ensure(e, 4 * tF));
    RSAC_TRY(s.d_n_found.ensure(e, 4 * 3 * c1));
    s.proj_uploaded = true;
    return RSAC_OK;
}

int rsac_proj_search_run(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    GuidedState& s = e->guided;
    if (!s.proj_uploaded) { e->err = "rsac_proj_search_run before rsac_proj_search_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    s.proj_ran = true;
    if (s.proj_C == 0) return RSAC_OK;
    ProjSearchArgs a;
    a.views = (const KfViewDev*)s.d_views.p; a.kp_xy = (const float*)s.d_kp_xy.p; a.kp_octave = (const int32_t*)s.d_kp_octave.p;
    a.kp_angle = (const float*)s.d_kp_angle.p; a.desc = (const uint32_t*)s.d_desc.p; a.mp_valid = (const uint8_t*)s.d_mp_valid.p;
    a.mp_xyz = (const float*)s.d_mp_xyz.p; a.mp_desc = (const uint32_t*)s.d_mp_desc.p; a.mp_maxdist = (const float*)s.d_mp_maxdist.p;
    a.mp_mindist = (const float*)s.d_mp_mindist.p; a.grid_off = (const int32_t*)s.d_grid_off.p; a.grid_idx = (const int32_t*)s.d_grid_idx.p;
    a.C = s.proj_C; a.vframe = (const int32_t*)s.d_kf1.p; a.vkf = (const int32_t*)s.d_kf2.p; a.K = (const float*)s.d_K.p;
    a.Rcw = (const float*)s.d_R12.p; a.tcw = (const float*)s.d_t12.p; a.th = s.th; a.orb_dist = s.orb_dist; a.check_orientation = s.check_orientation;
    a.cap = s.cap;
    a.offF = (const int64_t*)s.d_off1.p; a.offK = (const int64_t*)s.d_off2.p;
    a.occupied = s.have_occupied ? (const uint8_t*)s.d_already1.p : nullptr;
    a.already_found = s.have_found ? (const uint8_t*)s.d_already2.p : nullptr;
    a.cand = (int32_t*)s.d_cand.p; a.cand_n = (int32_t*)s.d_cand_n.p; a.taken = (uint8_t*)s.d_taken.p; a.minidx = (int32_t*)s.d_minidx.p;
    a.frame_match = (int32_t*)s.d_match12.p;
    a.nmatches = (int32_t*)s.d_n_found.p; a.overflow = a.nmatches + s.proj_C; a.rounds = a.nmatches + 2 * (size_t)s.proj_C;
    RSAC_CUDA(e, cudaMemsetAsync(s.d_n_found.p, 0, 4 * 3 * (size_t)s.proj_C, st));
    const dim3 gc((unsigned)std::max(1, (s.maxN + 127) / 128), (unsigned)s.proj_C);
    e->stage_begin(RSAC_STAGE_SOLVE);
    proj_candidates_kernel<<<gc, 128, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SOLVE);
    e->stage_begin(RSAC_STAGE_SELECT);
    proj_assign_kernel<<<s.proj_C, 256, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

int rsac_proj_search_download(rsac_engine* e, int32_t* frame_match, int32_t* n_matches, int32_t* info)
{
    if (!e) return RSAC_ERR_INVALID;
    GuidedState& s = e->guided;
    if (!s.proj_ran) { e->err = "rsac_proj_search_download before rsac_proj_search_run"; return RSAC_ERR_STATE; }
    if (s.proj_C > 0) {
        if (frame_match && s.totalF > 0) RSAC_CUDA(e, cudaMemcpyAsync(frame_match, s.d_match12.p, 4 * (size_t)s.totalF, cudaMemcpyDeviceToHost, e->stream));
        if (n_matches) RSAC_CUDA(e, cudaMemcpyAsync(n_matches, s.d_n_found.p, 4 * (size_t)s.proj_C, cudaMemcpyDeviceToHost, e->stream));
        if (info) RSAC_CUDA(e, cudaMemcpyAsync(info, (int32_t*)s.d_n_found.p + s.proj_C, 8 * (size_t)s.proj_C, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_proj_search(rsac_engine* e, const rsac_proj_search_batch* b, int32_t* frame_match, int32_t* n_matches)
{
    int rc = rsac_proj_search_upload(e, b);
    if (rc) return rc;
    rc = rsac_proj_search_run(e);
    if (rc) return rc;
    return rsac_proj_search_download(e, frame_match, n_matches, nullptr);
}

// ---------------------------------------------------------------- OptimizeSim3 chained behind SearchBySim3
static __global__ void sim3opt_chain_meta_kernel(Sim3OptMeta* metas, int C, const int64_t* off1, const int32_t* n_edges, const float* K1,
                                                 const float* K2, const float* R12, const float* t12, const float* s12, float th2, int fix_scale)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    Sim3OptMeta m;
    m.off = off1[c];
    m.n = n_edges[c];
    m.fix_scale = fix_scale;
    m.th2 = th2;
    for (int k = 0; k < 4; ++k) { m.K1[k] = K1[4 * c + k]; m.K2[k] = K2[4 * c + k]; }
    for (int k = 0; k < 9; ++k) m.R12[k] = R12[9 * c + k];
    for (int k = 0; k < 3; ++k) m.t12[k] = t12[3 * c + k];
    m.s12 = s12 ? s12[c] : 1.0f;
    metas[c] = m;
}

int rsac_sim3opt_from_search(rsac_engine* e, float th2, int fix_scale, const float* K2)
{
    if (!e) return RSAC_ERR_INVALID;
    GuidedState& g = e->guided;
    if (!g.ran) { e->err = "rsac_sim3opt_from_search before rsac_sim3_search_run"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    Sim3OptState& s = e->sim3opt;
    s.uploaded = false; s.ran = false;
    const int C = g.C;
    s.C = C; s.total = g.total1; s.chained = true;
    const size_t tot = (size_t)std::max<int64_t>(g.total1, 1), c1 = (size_t)std::max(C, 1);
    RSAC_TRY(s.d_metas.ensure(e, sizeof(Sim3OptMeta) * c1));
    RSAC_TRY(s.d_x1.ensure(e, tot * 12)); RSAC_TRY(s.d_x2.ensure(e, tot * 12));
    RSAC_TRY(s.d_o1.ensure(e, tot * 8)); RSAC_TRY(s.d_o2.ensure(e, tot * 8));
    RSAC_TRY(s.d_is1.ensure(e, tot * 4)); RSAC_TRY(s.d_is2.ensure(e, tot * 4));
    RSAC_TRY(s.d_removed.ensure(e, tot)); RSAC_TRY(s.d_src.ensure(e, tot * 4)); RSAC_TRY(s.d_full.ensure(e, tot));
    RSAC_TRY(s.d_nedges.ensure(e, 4 * c1)); RSAC_TRY(s.d_K2.ensure(e, 16 * c1));
    RSAC_TRY(s.d_results.ensure(e, sizeof(rsac_sim3opt_result) * c1));
    if (C == 0) { s.uploaded = true; return RSAC_OK; }
    cudaStream_t st = e->stream;
    if (K2) {
        float* h = (float*)s.h_metas.ensure(16 * c1);
        if (!h) { e->err = "pinned allocation failed"; return RSAC_ERR_ALLOC; }
        memcpy(h, K2, 16 * (size_t)C);
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_K2.p, h, 16 * (size_t)C, cudaMemcpyHostToDevice, st));
        s.h_metas.mark(st);
    }
    Sim3OptChainArgs a;
    a.views = (const KfViewDev*)g.d_views.p; a.kp_xy = (const float*)g.d_kp_xy.p; a.kp_octave = (const int32_t*)g.d_kp_octave.p;
    a.mp_valid = (const uint8_t*)g.d_mp_valid.p; a.mp_xyz = (const float*)g.d_mp_xyz.p;
    a.C = C; a.kf1 = (const int32_t*)g.d_kf1.p; a.kf2 = (const int32_t*)g.d_kf2.p; a.off1 = (const int64_t*)g.d_off1.p;
    a.matched_in = g.have_matched ? (const int32_t*)g.d_matched_in.p : nullptr; a.match12 = (const int32_t*)g.d_match12.p;
    a.x1c = (float*)s.d_x1.p; a.x2c = (float*)s.d_x2.p; a.o1 = (float*)s.d_o1.p; a.o2 = (float*)s.d_o2.p; a.is1 = (float*)s.d_is1.p; a.is2 = (float*)s.d_is2.p;
    a.src = (int32_t*)s.d_src.p; a.n_edges = (int32_t*)s.d_nedges.p;
    e->stage_begin(RSAC_STAGE_PACK);
    sim3opt_from_search_kernel<<<C, 128, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_PACK);
    e->stage_begin(RSAC_STAGE_PACK);
    sim3opt_chain_meta_kernel<<<(C + 127) / 128, 128, 0, st>>>((Sim3OptMeta*)s.d_metas.p, C, a.off1, a.n_edges, (const float*)g.d_K.p,
                                                            K2 ? (const float*)s.d_K2.p : (const float*)g.d_K.p, (const float*)g.d_R12.p,
                                                            (const float*)g.d_t12.p, g.have_scale ? (const float*)g.d_s12.p : nullptr, th2, fix_scale);
    e->stage_end(RSAC_STAGE_PACK);
    RSAC_CUDA(e, cudaGetLastError());
    s.uploaded = true;
    return RSAC_OK;
}

// after rsac_sim3opt_run on a chained batch: results [C] and, per KF1 feature of every pair, 0 = match kept, 1 = match removed by
// the optimiser (vpMatches1[i] = nullptr, Optimizer.cpp:1196-1207), 2 = no match / not an edge; n_edges [C] (optional)
int rsac_sim3opt_download_chained(rsac_engine* e, rsac_sim3opt_result* results, uint8_t* flags, int32_t* n_edges)
{
    if (!e) return RSAC_ERR_INVALID;
    Sim3OptState& s = e->sim3opt;
    GuidedState& g = e->guided;
    if (!s.ran || !s.chained) { e->err = "rsac_sim3opt_download_chained needs rsac_sim3opt_from_search + rsac_sim3opt_run"; return RSAC_ERR_STATE; }
    if (s.C > 0) {
        e->stage_begin(RSAC_STAGE_PACK);
        sim3opt_scatter_flags_kernel<<<s.C, 128, 0, e->stream>>>((const KfViewDev*)g.d_views.p, (const int32_t*)g.d_kf1.p, (const int64_t*)g.d_off1.p,
                                                                 (const int32_t*)s.d_nedges.p, (const uint8_t*)s.d_removed.p, (const int32_t*)s.d_src.p,
                                                                 (uint8_t*)s.d_full.p);
        e->stage_end(RSAC_STAGE_PACK);
        RSAC_CUDA(e, cudaGetLastError());
        if (results) RSAC_CUDA(e, cudaMemcpyAsync(results, s.d_results.p, sizeof(rsac_sim3opt_result) * (size_t)s.C, cudaMemcpyDeviceToHost, e->stream));
        if (flags && s.total > 0) RSAC_CUDA(e, cudaMemcpyAsync(flags, s.d_full.p, (size_t)s.total, cudaMemcpyDeviceToHost, e->stream));
        if (n_edges) RSAC_CUDA(e, cudaMemcpyAsync(n_edges, s.d_nedges.p, 4 * (size_t)s.C, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}
