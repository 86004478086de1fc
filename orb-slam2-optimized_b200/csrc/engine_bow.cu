// engine_bow.cu -- C ABI for the batched ORBmatcher::SearchByBoW (include/ransac_b200.h).
#include "engine_shared.cuh"
#include "bow.cuh"

int rsac_bow_upload(rsac_engine* e, const rsac_bow_batch* b)
{
    if (!e || !b || b->n_sets < 0 || b->C < 0 || (b->n_sets > 0 && !b->sets) || (b->C > 0 && (!b->query_set || !b->target_set)))
        return RSAC_ERR_INVALID;
    if (b->mode != 0 && b->mode != 1) { e->err = "mode must be 0 (KF, Frame) or 1 (KF1, KF2)"; return RSAC_ERR_INVALID; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    BowState& s = e->bow;
    s.uploaded = false; s.ran = false;
    const int S = b->n_sets, C = b->C;
    std::vector<BowSet> sets(std::max(S, 1));
    int64_t nfeat = 0, nnodes = 0, nnf = 0;
    bool any_valid = false, any_mp = false;
    for (int i = 0; i < S; ++i) {
        const rsac_bow_features& f = b->sets[i];
        if (f.n_feat < 0 || f.n_nodes < 0 || (f.n_feat > 0 && (!f.desc || !f.angle)) || (f.n_nodes > 0 && (!f.node_ids || !f.node_off || !f.node_feat))) {
            e->err = "bad feature set"; return RSAC_ERR_INVALID;
        }
        if (f.n_nodes > 0 && f.node_off[0] != 0) { e->err = "feature vector: node_off[0] must be 0"; return RSAC_ERR_INVALID; }
        for (int k = 0; k < f.n_nodes; ++k) {
            if (f.node_off[k + 1] < f.node_off[k] || f.node_off[k + 1] - f.node_off[k] >= 32768 || (k > 0 && f.node_ids[k] <= f.node_ids[k - 1])) {
                e->err = "feature vector: node ids must ascend, node sizes must be below 32768"; return RSAC_ERR_INVALID;
            }
        }
        BowSet& d = sets[i];
        d.feat_off = (int32_t)nfeat; d.n_feat = f.n_feat;
        d.node_off = (int32_t)nnodes; d.n_nodes = f.n_nodes;
        d.noff_off = (int32_t)(nnodes + i);
        d.nf_off = (int32_t)nnf;
        nfeat += f.n_feat; nnodes += f.n_nodes; nnf += f.n_nodes > 0 ? f.node_off[f.n_nodes] : 0;
        any_valid = any_valid || f.valid != nullptr;
        any_mp = any_mp || f.mp_index != nullptr;
        if (nfeat > INT32_MAX / 8 || nnf > INT32_MAX) { e->err = "batch too large"; return RSAC_ERR_INVALID; }
    }
    // work items: (pair, query node); per-pair output offsets
    std::vector<int2> items;
    s.t2q_off.assign(C + 1, 0); s.q2t_off.assign(C + 1, 0);
    for (int p = 0; p < C; ++p) {
        const int qs = b->query_set[p], ts = b->target_set[p];
        if (qs < 0 || qs >= S || ts < 0 || ts >= S) { e->err = "set index out of range"; return RSAC_ERR_INVALID; }
        for (int k = 0; k < sets[qs].n_nodes; ++k) items.push_back(make_int2(p, k));
        s.t2q_off[p + 1] = s.t2q_off[p] + sets[ts].n_feat;
        s.q2t_off[p + 1] = s.q2t_off[p] + sets[qs].n_feat;
    }
    s.C = C; s.n_items = (int)items.size(); s.mode = b->mode; s.check_orientation = b->check_orientation; s.nn_ratio = b->nn_ratio;
    s.total_t = s.t2q_off[C]; s.total_q = s.q2t_off[C];
    s.have_valid = any_valid;
    s.have_mp_index = any_mp;
    s.one_target = true;
    s.target_n_feat.assign(C, 0);
    for (int p = 0; p < C; ++p) {
        s.one_target = s.one_target && b->target_set[p] == b->target_set[0];
        s.target_n_feat[p] = sets[b->target_set[p]].n_feat;
    }

    // one pinned staging buffer, one layout, a handful of copies
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_sets = 0, o_q = al(o_sets + sizeof(BowSet) * sets.size()), o_t = al(o_q + 4 * (size_t)std::max(C, 1));
    const size_t o_to = al(o_t + 4 * (size_t)std::max(C, 1)), o_qo = al(o_to + 8 * (size_t)(C + 1)), o_it = al(o_qo + 8 * (size_t)(C + 1));
    const size_t o_desc = al(o_it + sizeof(int2) * std::max<size_t>(items.size(), 1)), o_ang = al(o_desc + 32 * (size_t)std::max<int64_t>(nfeat, 1));
    const size_t o_val = al(o_ang + 4 * (size_t)std::max<int64_t>(nfeat, 1)), o_nid = al(o_val + (size_t)std::max<int64_t>(nfeat, 1));
    const size_t o_nst = al(o_nid + 4 * (size_t)std::max<int64_t>(nnodes, 1)), o_nf = al(o_nst + 4 * (size_t)(nnodes + S + 1));
    const size_t total = al(o_nf + 4 * (size_t)std::max<int64_t>(nnf, 1));
    char* h = (char*)s.h_stage.ensure(total);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    memcpy(h + o_sets, sets.data(), sizeof(BowSet) * sets.size());
    if (C > 0) { memcpy(h + o_q, b->query_set, 4 * (size_t)C); memcpy(h + o_t, b->target_set, 4 * (size_t)C); }
    memcpy(h + o_to, s.t2q_off.data(), 8 * (size_t)(C + 1));
    memcpy(h + o_qo, s.q2t_off.data(), 8 * (size_t)(C + 1));
    if (!items.empty()) memcpy(h + o_it, items.data(), sizeof(int2) * items.size());
    for (int i = 0; i < S; ++i) {
        const rsac_bow_features& f = b->sets[i];
        const BowSet& d = sets[i];
        if (f.n_feat > 0) {
            memcpy(h + o_desc + 32 * (size_t)d.feat_off, f.desc, 32 * (size_t)f.n_feat);
            memcpy(h + o_ang + 4 * (size_t)d.feat_off, f.angle, 4 * (size_t)f.n_feat);
            if (f.valid) memcpy(h + o_val + (size_t)d.feat_off, f.valid, (size_t)f.n_feat);
            else memset(h + o_val + (size_t)d.feat_off, 1, (size_t)f.n_feat);
        }
        if (f.n_nodes > 0) {
            memcpy(h + o_nid + 4 * (size_t)d.node_off, f.node_ids, 4 * (size_t)f.n_nodes);
            memcpy(h + o_nf + 4 * (size_t)d.nf_off, f.node_feat, 4 * (size_t)f.node_off[f.n_nodes]);
        }
        int32_t* ns = (int32_t*)(h + o_nst) + d.noff_off;
        for (int k = 0; k <= f.n_nodes; ++k) ns[k] = f.n_nodes > 0 ? f.node_off[k] - f.node_off[0] : 0;
    }
    struct { DevBuf* d; size_t off, bytes; } cp[] = {
        {&s.d_sets, o_sets, sizeof(BowSet) * sets.size()}, {&s.d_qset, o_q, 4 * (size_t)std::max(C, 1)}, {&s.d_tset, o_t, 4 * (size_t)std::max(C, 1)},
        {&s.d_t2q_off, o_to, 8 * (size_t)(C + 1)}, {&s.d_q2t_off, o_qo, 8 * (size_t)(C + 1)}, {&s.d_items, o_it, sizeof(int2) * std::max<size_t>(items.size(), 1)},
        {&s.d_desc, o_desc, 32 * (size_t)std::max<int64_t>(nfeat, 1)}, {&s.d_angle, o_ang, 4 * (size_t)std::max<int64_t>(nfeat, 1)},
        {&s.d_valid, o_val, (size_t)std::max<int64_t>(nfeat, 1)}, {&s.d_node_ids, o_nid, 4 * (size_t)std::max<int64_t>(nnodes, 1)},
        {&s.d_node_start, o_nst, 4 * (size_t)(nnodes + S + 1)}, {&s.d_node_feat, o_nf, 4 * (size_t)std::max<int64_t>(nnf, 1)}};
    for (auto& c : cp) {
        RSAC_TRY(c.d->ensure(e, c.bytes));
        RSAC_CUDA(e, cudaMemcpyAsync(c.d->p, h + c.off, c.bytes, cudaMemcpyHostToDevice, e->stream));
    }
    s.h_stage.mark(e->stream);
    if (any_mp) {
        // map-point table slots of the keyframes' features (rsac_pnp_upload_from_bow): pageable copies, made before returning
        RSAC_TRY(s.d_mp_index.ensure(e, 4 * (size_t)std::max<int64_t>(nfeat, 1)));
        RSAC_CUDA(e, cudaMemsetAsync(s.d_mp_index.p, 0xff, 4 * (size_t)std::max<int64_t>(nfeat, 1), e->stream));
        for (int i = 0; i < S; ++i)
            if (b->sets[i].mp_index && b->sets[i].n_feat > 0)
                RSAC_CUDA(e, cudaMemcpyAsync((uint32_t*)s.d_mp_index.p + sets[i].feat_off, b->sets[i].mp_index, 4 * (size_t)b->sets[i].n_feat,
                                             cudaMemcpyHostToDevice, e->stream));
        RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    }
    RSAC_TRY(s.d_t2q.ensure(e, 4 * (size_t)std::max<int64_t>(s.total_t, 1)));
    RSAC_TRY(s.d_q2t.ensure(e, 4 * (size_t)std::max<int64_t>(s.total_q, 1)));
    RSAC_TRY(s.d_bin.ensure(e, (size_t)std::max<int64_t>(s.total_t, 1)));
    RSAC_TRY(s.d_nmatches.ensure(e, 4 * (size_t)std::max(C, 1)));
    s.uploaded = true;
    return RSAC_OK;
}

int rsac_bow_run(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    BowState& s = e->bow;
    if (!s.uploaded) { e->err = "rsac_bow_run before rsac_bow_upload"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    if (s.C == 0) { s.ran = true; return RSAC_OK; }
    RSAC_CUDA(e, cudaMemsetAsync(s.d_t2q.p, 0xff, 4 * (size_t)std::max<int64_t>(s.total_t, 1), st));
    RSAC_CUDA(e, cudaMemsetAsync(s.d_q2t.p, 0xff, 4 * (size_t)std::max<int64_t>(s.total_q, 1), st));
    BowArgs a;
    a.sets = (const BowSet*)s.d_sets.p; a.qset = (const int32_t*)s.d_qset.p; a.tset = (const int32_t*)s.d_tset.p;
    a.t2q_off = (const int64_t*)s.d_t2q_off.p; a.q2t_off = (const int64_t*)s.d_q2t_off.p; a.items = (const int2*)s.d_items.p;
    a.n_items = s.n_items; a.C = s.C;
    a.desc = (const uint32_t*)s.d_desc.p; a.angle = (const float*)s.d_angle.p; a.valid = s.have_valid ? (const uint8_t*)s.d_valid.p : nullptr;
    a.node_ids = (const uint32_t*)s.d_node_ids.p; a.node_start = (const int32_t*)s.d_node_start.p; a.node_feat = (const uint32_t*)s.d_node_feat.p;
    a.t2q = (int32_t*)s.d_t2q.p; a.q2t = (int32_t*)s.d_q2t.p; a.bin_t = (int8_t*)s.d_bin.p; a.n_matches = (int32_t*)s.d_nmatches.p;
    a.nn_ratio = s.nn_ratio; a.check_orientation = s.check_orientation; a.mode = s.mode;
    if (s.n_items > 0) {
        const int blocks = std::max(1, std::min((s.n_items + 3) / 4, e->sm_count * 16));     // four warps per CTA, grid-stride
        e->stage_begin(RSAC_STAGE_SOLVE);
        bow_match_kernel<<<blocks, 128, 0, st>>>(a);
        e->stage_end(RSAC_STAGE_SOLVE);
        RSAC_CUDA(e, cudaGetLastError());
    }
    e->stage_begin(RSAC_STAGE_SELECT);
    bow_orient_kernel<<<s.C, 128, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    RSAC_CUDA(e, cudaGetLastError());
    s.ran = true;
    return RSAC_OK;
}

int rsac_bow_download(rsac_engine* e, int32_t* matches, int32_t* n_matches)
{
    if (!e) return RSAC_ERR_INVALID;
    BowState& s = e->bow;
    if (!s.ran) { e->err = "rsac_bow_download before rsac_bow_run"; return RSAC_ERR_STATE; }
    if (s.C > 0) {
        const bool by_t = s.mode == 0;
        const int64_t n = by_t ? s.total_t : s.total_q;
        if (matches && n > 0)
            RSAC_CUDA(e, cudaMemcpyAsync(matches, by_t ? s.d_t2q.p : s.d_q2t.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
        if (n_matches) RSAC_CUDA(e, cudaMemcpyAsync(n_matches, s.d_nmatches.p, 4 * (size_t)s.C, cudaMemcpyDeviceToHost, e->stream));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_bow_match(rsac_engine* e, const rsac_bow_batch* b, int32_t* matches, int32_t* n_matches)
{
    int rc = rsac_bow_upload(e, b);
    if (rc) return rc;
    rc = rsac_bow_run(e);
    if (rc) return rc;
    return rsac_bow_download(e, matches, n_matches);
}
