// engine_kfdb.cu -- C ABI for the batched keyframe-database candidate retrieval (include/ransac_b200.h):
// KeyFrameDatabase::DetectRelocalizationCandidates / DetectLoopCandidates (src/KeyFrameDatabase.cpp:174-284, 51-172).
#include "engine_shared.cuh"
#include "kfdb.cuh"

int rsac_kfdb_upload(rsac_engine* e, const rsac_kfdb* db)
{
    if (!e || !db || db->n_keyframes < 0 || !db->bow_off) return RSAC_ERR_INVALID;
    const int K = db->n_keyframes;
    const int64_t nnz = db->bow_off[K];
    if (db->bow_off[0] != 0 || nnz < 0 || (nnz > 0 && (!db->bow_word || !db->bow_val)) || (K > 0 && !db->covis)) {
        e->err = "bad keyframe database"; return RSAC_ERR_INVALID;
    }
    uint32_t max_word = 0;
    for (int k = 0; k < K; ++k) {
        if (db->bow_off[k + 1] < db->bow_off[k] || db->bow_off[k + 1] - db->bow_off[k] > INT32_MAX) { e->err = "bow_off must ascend"; return RSAC_ERR_INVALID; }
        for (int64_t i = db->bow_off[k] + 1; i < db->bow_off[k + 1]; ++i)
            if (db->bow_word[i] <= db->bow_word[i - 1]) { e->err = "BowVector word ids must ascend (std::map order)"; return RSAC_ERR_INVALID; }
        if (db->bow_off[k + 1] > db->bow_off[k]) max_word = std::max(max_word, db->bow_word[db->bow_off[k + 1] - 1]);
    }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    KfdbState& s = e->kfdb;
    s.bm_words = (int64_t)max_word / 32 + 1;
    s.db_ready = false; s.ran = false;
    cudaStream_t st = e->stream;
    const size_t k1 = (size_t)std::max(K, 1), nz = (size_t)std::max<int64_t>(nnz, 1);
    RSAC_TRY(s.d_kf_off.ensure(e, 8 * (k1 + 1)));
    RSAC_TRY(s.d_kf_word.ensure(e, 4 * nz));
    RSAC_TRY(s.d_kf_val.ensure(e, 8 * nz));
    RSAC_TRY(s.d_covis.ensure(e, 40 * k1));
    RSAC_TRY(s.d_state.ensure(e, 4 * k1));
    // the database is uploaded rarely (it grows by one keyframe at a time): plain synchronous-looking copies from pageable memory
    RSAC_CUDA(e, cudaMemcpyAsync(s.d_kf_off.p, db->bow_off, 8 * (size_t)(K + 1), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_kf_word.p, db->bow_word, 4 * (size_t)nnz, cudaMemcpyHostToDevice, st));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_kf_val.p, db->bow_val, 8 * (size_t)nnz, cudaMemcpyHostToDevice, st));
    }
    if (K > 0) {
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_covis.p, db->covis, 40 * (size_t)K, cudaMemcpyHostToDevice, st));
        if (db->score_state) RSAC_CUDA(e, cudaMemcpyAsync(s.d_state.p, db->score_state, 4 * (size_t)K, cudaMemcpyHostToDevice, st));
        else RSAC_CUDA(e, cudaMemsetAsync(s.d_state.p, 0, 4 * (size_t)K, st));
    }
    RSAC_CUDA(e, cudaStreamSynchronize(st));      // the caller's arrays may be freed on return
    s.K = K;
    s.K2 = 1;
    while (s.K2 < K) s.K2 <<= 1;
    s.db_ready = true;
    s.queries_ready = false;
    return RSAC_OK;
}

int rsac_kfdb_query_upload(rsac_engine* e, const rsac_kfdb_queries* qs)
{
    if (!e || !qs || qs->Q < 0 || qs->Q > 65535 || !qs->bow_off) return RSAC_ERR_INVALID;
    KfdbState& s = e->kfdb;
    if (!s.db_ready) { e->err = "rsac_kfdb_query_upload before rsac_kfdb_upload"; return RSAC_ERR_STATE; }
    if (qs->mode != 0 && qs->mode != 1) { e->err = "mode must be 0 (relocalisation) or 1 (loop detection)"; return RSAC_ERR_INVALID; }
    const int Q = qs->Q;
    const int64_t nnz = qs->bow_off[Q];
    if (qs->bow_off[0] != 0 || nnz < 0 || (nnz > 0 && (!qs->bow_word || !qs->bow_val))) { e->err = "bad query vectors"; return RSAC_ERR_INVALID; }
    if (qs->mode == 1 && Q > 0 && (!qs->min_score || !qs->conn_off)) { e->err = "loop detection needs min_score and conn_off"; return RSAC_ERR_INVALID; }
    for (int q = 0; q < Q; ++q) {
        if (qs->bow_off[q + 1] < qs->bow_off[q] || qs->bow_off[q + 1] - qs->bow_off[q] > INT32_MAX) { e->err = "bow_off must ascend"; return RSAC_ERR_INVALID; }
        for (int64_t i = qs->bow_off[q] + 1; i < qs->bow_off[q + 1]; ++i)
            if (qs->bow_word[i] <= qs->bow_word[i - 1]) { e->err = "BowVector word ids must ascend (std::map order)"; return RSAC_ERR_INVALID; }
    }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    s.queries_ready = false; s.ran = false;
    cudaStream_t st = e->stream;
    const int64_t nconn = (qs->mode == 1 && Q > 0) ? qs->conn_off[Q] : 0;
    if (nconn < 0 || (nconn > 0 && !qs->conn)) { e->err = "bad connected-keyframe lists"; return RSAC_ERR_INVALID; }
    // one pinned staging buffer: offsets, words, values, min scores, connected lists (sorted per query: they are sets)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t q1 = (size_t)std::max(Q, 1), nz = (size_t)std::max<int64_t>(nnz, 1), nc = (size_t)std::max<int64_t>(nconn, 1);
    const size_t o_off = 0, o_w = al(o_off + 8 * (q1 + 1)), o_v = al(o_w + 4 * nz), o_ms = al(o_v + 8 * nz), o_co = al(o_ms + 4 * q1);
    const size_t o_c = al(o_co + 8 * (q1 + 1)), total = al(o_c + 4 * nc);
    char* h = (char*)s.h_stage.ensure(total);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    memset(h, 0, total);
    memcpy(h + o_off, qs->bow_off, 8 * (size_t)(Q + 1));
    if (nnz > 0) { memcpy(h + o_w, qs->bow_word, 4 * (size_t)nnz); memcpy(h + o_v, qs->bow_val, 8 * (size_t)nnz); }
    if (qs->mode == 1 && Q > 0) {
        memcpy(h + o_ms, qs->min_score, 4 * (size_t)Q);
        memcpy(h + o_co, qs->conn_off, 8 * (size_t)(Q + 1));
        if (nconn > 0) memcpy(h + o_c, qs->conn, 4 * (size_t)nconn);
        int32_t* c = (int32_t*)(h + o_c);
        for (int q = 0; q < Q; ++q) {
            if (qs->conn_off[q + 1] < qs->conn_off[q]) { e->err = "conn_off must ascend"; return RSAC_ERR_INVALID; }
            std::sort(c + qs->conn_off[q], c + qs->conn_off[q + 1]);
        }
    }
    struct { DevBuf* d; size_t off, bytes; } cp[] = {{&s.d_q_off, o_off, 8 * (q1 + 1)}, {&s.d_q_word, o_w, 4 * nz}, {&s.d_q_val, o_v, 8 * nz},
                                                     {&s.d_min_score, o_ms, 4 * q1}, {&s.d_conn_off, o_co, 8 * (q1 + 1)}, {&s.d_conn, o_c, 4 * nc}};
    for (auto& c : cp) {
        RSAC_TRY(c.d->ensure(e, c.bytes));
        RSAC_CUDA(e, cudaMemcpyAsync(c.d->p, h + c.off, c.bytes, cudaMemcpyHostToDevice, st));
    }
    s.h_stage.mark(st);
    const size_t qk = q1 * (size_t)std::max(s.K, 1);
    RSAC_TRY(s.d_cw.ensure(e, 4 * qk)); RSAC_TRY(s.d_wstar.ensure(e, 4 * qk)); RSAC_TRY(s.d_si.ensure(e, 4 * qk));
    RSAC_TRY(s.d_eff.ensure(e, 4 * qk)); RSAC_TRY(s.d_acc.ensure(e, 4 * qk)); RSAC_TRY(s.d_best.ensure(e, 4 * qk));
    RSAC_TRY(s.d_firstpos.ensure(e, 4 * qk)); RSAC_TRY(s.d_out.ensure(e, 4 * qk));
    RSAC_TRY(s.d_keys.ensure(e, 8 * q1 * (size_t)s.K2));
    RSAC_TRY(s.d_min_common.ensure(e, 4 * q1)); RSAC_TRY(s.d_best_acc.ensure(e, 4 * q1)); RSAC_TRY(s.d_n_out.ensure(e, 4 * q1));
    // membership bitmaps: up to 2 MB per query (a 16 M-word vocabulary; ORBvoc has 10^6) and 1 GB per batch, else binary search
    s.use_bitmap = env_int("RSAC_KFDB_BITMAP", 1) != 0 && s.bm_words <= (1 << 19) && (int64_t)q1 * s.bm_words * 4 <= (1ll << 30);
    if (s.use_bitmap) RSAC_TRY(s.d_bitmap.ensure(e, 4 * q1 * (size_t)s.bm_words));
    s.max_nq = 0;
    for (int q = 0; q < Q; ++q) s.max_nq = std::max<int64_t>(s.max_nq, qs->bow_off[q + 1] - qs->bow_off[q]);
    s.Q = Q; s.mode = qs->mode;
    s.queries_ready = true;
    return RSAC_OK;
}

int rsac_kfdb_run(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    KfdbState& s = e->kfdb;
    if (!s.db_ready || !s.queries_ready) { e->err = "rsac_kfdb_run before the database and the queries are uploaded"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    s.ran = true;
    if (s.Q == 0) return RSAC_OK;
    KfdbArgs a;
    a.K = s.K; a.K2 = s.K2;
    a.kf_off = (const int64_t*)s.d_kf_off.p; a.kf_word = (const uint32_t*)s.d_kf_word.p; a.kf_val = (const double*)s.d_kf_val.p;
    a.covis = (const int32_t*)s.d_covis.p; a.state = (float*)s.d_state.p;
    a.Q = s.Q; a.mode = s.mode;
    a.q_off = (const int64_t*)s.d_q_off.p; a.q_word = (const uint32_t*)s.d_q_word.p; a.q_val = (const double*)s.d_q_val.p;
    a.min_score = (const float*)s.d_min_score.p; a.conn_off = (const int64_t*)s.d_conn_off.p; a.conn = (const int32_t*)s.d_conn.p;
    a.cw = (int32_t*)s.d_cw.p; a.wstar = (uint32_t*)s.d_wstar.p; a.si = (float*)s.d_si.p; a.eff = (float*)s.d_eff.p; a.acc = (float*)s.d_acc.p;
    a.best = (int32_t*)s.d_best.p; a.firstpos = (int32_t*)s.d_firstpos.p; a.keys = (unsigned long long*)s.d_keys.p; a.out = (int32_t*)s.d_out.p;
    a.min_common = (int32_t*)s.d_min_common.p; a.best_acc = (float*)s.d_best_acc.p; a.n_out = (int32_t*)s.d_n_out.p;
    if (s.K == 0) { RSAC_CUDA(e, cudaMemsetAsync(s.d_n_out.p, 0, 4 * (size_t)s.Q, st)); return RSAC_OK; }
    RSAC_CUDA(e, cudaMemsetAsync(s.d_firstpos.p, 0x7f, 4 * (size_t)s.Q * s.K, st));
    a.q_bitmap = nullptr; a.bm_words = (int32_t)s.bm_words;
    if (s.use_bitmap && s.max_nq > 0) {
        a.q_bitmap = (uint32_t*)s.d_bitmap.p;
        RSAC_CUDA(e, cudaMemsetAsync(s.d_bitmap.p, 0, 4 * (size_t)s.Q * (size_t)s.bm_words, st));
        e->stage_begin(RSAC_STAGE_PACK);
        kfdb_bitmap_kernel<<<dim3((unsigned)std::min<int64_t>((s.max_nq + 255) / 256, 64), (unsigned)s.Q), 256, 0, st>>>(a);
        e->stage_end(RSAC_STAGE_PACK);
    }
    const dim3 gw((unsigned)((s.K + kKfdbWarps - 1) / kKfdbWarps), (unsigned)s.Q), gt((unsigned)((s.K + 255) / 256), (unsigned)s.Q);
    e->stage_begin(RSAC_STAGE_SOLVE);
    kfdb_common_kernel<<<gw, kKfdbWarps * 32, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SOLVE);
    e->stage_begin(RSAC_STAGE_RNG);
    kfdb_threshold_kernel<<<s.Q, 256, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_RNG);
    e->stage_begin(RSAC_STAGE_SCORE);
    kfdb_score_kernel<<<gw, kKfdbWarps * 32, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SCORE);
    if (s.mode == 0) {
        e->stage_begin(RSAC_STAGE_RNG);
        kfdb_state_kernel<<<(s.K + 255) / 256, 256, 0, st>>>(a);
        e->stage_end(RSAC_STAGE_RNG);
    }
    e->stage_begin(RSAC_STAGE_SELECT);
    kfdb_acc_kernel<<<gt, 256, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    e->stage_begin(RSAC_STAGE_SELECT);
    kfdb_emit_kernel<<<s.Q, kKfdbEmitThreads, 0, st>>>(a);
    e->stage_end(RSAC_STAGE_SELECT);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

int rsac_kfdb_download(rsac_engine* e, int32_t* counts, int32_t* candidates, int32_t cap)
{
    if (!e || !counts || cap < 0 || (cap > 0 && !candidates)) return RSAC_ERR_INVALID;
    KfdbState& s = e->kfdb;
    if (!s.ran) { e->err = "rsac_kfdb_download before rsac_kfdb_run"; return RSAC_ERR_STATE; }
    if (s.Q == 0) return RSAC_OK;
    RSAC_CUDA(e, cudaMemcpyAsync(counts, s.d_n_out.p, 4 * (size_t)s.Q, cudaMemcpyDeviceToHost, e->stream));
    const int take = std::min(cap, s.K);
    if (take > 0)       // row q of the device buffer (stride K) -> row q of the caller's (stride cap)
        RSAC_CUDA(e, cudaMemcpy2DAsync(candidates, 4 * (size_t)cap, s.d_out.p, 4 * (size_t)s.K, 4 * (size_t)take, (size_t)s.Q,
                                       cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_kfdb_detect(rsac_engine* e, const rsac_kfdb_queries* qs, int32_t* counts, int32_t* candidates, int32_t cap)
{
    int rc = rsac_kfdb_query_upload(e, qs);
    if (rc) return rc;
    rc = rsac_kfdb_run(e);
    if (rc) return rc;
    return rsac_kfdb_download(e, counts, candidates, cap);
}

int rsac_kfdb_get_state(rsac_engine* e, float* score_state)
{
    if (!e || !score_state) return RSAC_ERR_INVALID;
    KfdbState& s = e->kfdb;
    if (!s.db_ready) { e->err = "no keyframe database"; return RSAC_ERR_STATE; }
    if (s.K > 0) RSAC_CUDA(e, cudaMemcpyAsync(score_state, s.d_state.p, 4 * (size_t)s.K, cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}
