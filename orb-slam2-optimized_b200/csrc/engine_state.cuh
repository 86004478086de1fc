// engine_state.cuh -- host-side state of an engine: stream, grow-only device buffers,
// per-batch bookkeeping, kernel-stage profiling with CUDA events.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/ransac_b200.h"
#include "common.cuh"
#include "score.cuh"

#define RSAC_CUDA(e, call)                                                                        \
    do {                                                                                          \
        cudaError_t _err = (call);                                                                \
        if (_err != cudaSuccess) {                                                                \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_err);                      \
            cudaGetLastError();                                                                   \
            return (_err == cudaErrorNoDevice || _err == cudaErrorInsufficientDriver) ? RSAC_ERR_NO_DEVICE : RSAC_ERR_CUDA; \
        }                                                                                         \
    } while (0)

#define RSAC_TRY(expr)                       \
    do {                                     \
        int _rc = (expr);                    \
        if (_rc != RSAC_OK) return _rc;      \
    } while (0)

struct rsac_engine;

namespace rsac {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(rsac_engine* e, size_t bytes);
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// pinned host staging for the small per-batch tables (metas, tiles, thresholds): a cudaMemcpyAsync from
// pageable memory would make the host wait for the stream's earlier work and defeat sweep pipelining
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;   // recorded after the last async copy out of this buffer
    void* ensure(size_t bytes)
    {
        if (done) cudaEventSynchronize(done);   // the previous copy out of this buffer has finished
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            if (cudaHostAlloc(&p, bytes + bytes / 4 + 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
            cap = bytes + bytes / 4 + 64;
        }
        if (!done) cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        return p;
    }
    void mark(cudaStream_t st) { if (done) cudaEventRecord(done, st); }
    void release()
    {
        if (p) cudaFreeHost(p);
        if (done) cudaEventDestroy(done);
        p = nullptr; cap = 0; done = nullptr;
    }
};

struct ScorePlanPOD {
    int threads = 32, tile_hyps = 64, chunk_cap = 32, grid = 1, hpl = 2;
    size_t smem = 0;
    int vlen = 1;                            // records per CTA in `work`
    bool by_list = false;                    // `work` is [problem][tiles]; the CTAs stride over a device-side problem list
    int tiles = 1;
    std::vector<ScoreGroup> work;            // [grid][vlen] per-CTA lists of group records (host copy)
};

struct BatchDims {
    int C = 0;
    int total = 0;             // correspondences
    int64_t sumH = 0;          // hypotheses
    int64_t total_words = 0;   // final-mask words
    int64_t total_hwords = 0;  // per-hypothesis mask words
    int64_t table_len = 0;
    int maxH = 0, maxN = 0, maxWords = 0;
};

struct PnpState {
    bool uploaded = false, ran = false, have_tables = false, have_cov = false, packed = false, tables_ready = false;
    BatchDims d;
    std::vector<ProblemMeta> metas;
    std::vector<ScoreGroup> groups;
    ScorePlanPOD plan;
    uint64_t h2d_bytes = 0;
    DevBuf d_metas, d_cP, d_th2, d_p3d, d_p2d, d_sigma2, d_cA, d_cB, d_uv, d_tables, d_poses, d_counts,
        d_results, d_masks, d_hmasks, d_sel, d_pw, d_us, d_al, d_cov, d_extra, d_visit;
    PinnedBuf h_stage, h_fromBow;
    // indexed wire format: the frame's keypoint table and the map-point table stay resident between uploads
    DevBuf d_kp_uv, d_kp_s2, d_mp_xyz, d_kp_idx, d_mp_idx, d_kpA, d_kpB, d_carry;
    int n_keypoints = 0, n_mappoints = 0;
    bool indexed_fused = false;             // the batch is packed straight from the index pairs (pnp_pack)
    bool flat_valid = true;                 // d_p3d / d_p2d / d_sigma2 hold the batch (always for flat uploads)
    bool kp_table_dirty = true;             // per-keypoint records (d_kpA / d_kpB) must be rebuilt
    float kp_th2 = -1.0f;
    double kp_K[4] = {0, 0, 0, 0};
    // early exit in phases (RSAC_FLAG_EARLY_EXIT): plans of the two hypothesis ranges, per-problem phase state
    bool ee_planned = false, ee_mode = false, ee_complete = false;
    bool run_eigen = false;                 // the current / last run used RSAC_FLAG_EPNP_EIGEN (selects the solver kernel)
    bool plans_valid = false;               // plans and device work arrays match shape_sig
    std::vector<int32_t> shape_sig;         // n, H, fx, fy of every problem + stage boundaries
    int ee_HA = 0;
    std::vector<int> ee_bounds;                      // stage boundaries b0 < ... < b(K-1) = maxH
    std::vector<ScorePlanPOD> ee_plans;              // [0,b0) static; [b(j-1),bj) list-driven; last: clean-up [b0,H)
    std::vector<std::vector<ScoreGroup>> ee_groups;
    std::vector<DevBuf> ee_visit;
    ScoreArgs ee_sa;
    DevBuf d_ee;
    PinnedBuf h_stageEE;
    // CUDA graph of the staged sweep (pnp_run_early): captured on the second run of a (shape, flags, output) key --
    // the first run is eager and sizes every buffer -- and replayed until the key or a device pointer changes
    cudaGraphExec_t graph = nullptr;
    std::vector<int64_t> graph_key;
    int64_t graph_nodes = 0;
    uint64_t alloc_epoch_seen = 0;
    int eager_runs = 0;
    int64_t plan_version = 0;               // bumped whenever the scoring plans are rebuilt
    void release()
    {
        if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; }
        h_stage.release();
        h_fromBow.release();
        h_stageEE.release();
        d_ee.release();
        for (auto& b : ee_visit) b.release();
        DevBuf* all[] = {&d_metas, &d_cP, &d_th2, &d_p3d, &d_p2d, &d_sigma2, &d_cA, &d_cB, &d_uv, &d_tables, &d_poses,
                         &d_counts, &d_results, &d_masks, &d_hmasks, &d_sel, &d_pw, &d_us, &d_al, &d_cov, &d_extra, &d_visit,
                         &d_kp_uv, &d_kp_s2, &d_mp_xyz, &d_kp_idx, &d_mp_idx, &d_kpA, &d_kpB, &d_carry};
        for (DevBuf* b : all) b->release();
    }
};

struct ScoreState {
    bool uploaded = false, ran = false, with_masks = false;
    int H = 0, n = 0;
    ScorePlanPOD plan;
    std::vector<ProblemMeta> metas;
    std::vector<ScoreGroup> groups;
    DevBuf d_metas, d_cP, d_p3d, d_p2d, d_maxerr, d_cA, d_cB, d_uv, d_poses, d_counts, d_hmasks, d_visit;
    void release()
    {
        DevBuf* all[] = {&d_metas, &d_cP, &d_p3d, &d_p2d, &d_maxerr, &d_cA, &d_cB, &d_uv, &d_poses, &d_counts, &d_hmasks, &d_visit};
        for (DevBuf* b : all) b->release();
    }
};

struct Sim3State {
    bool uploaded = false, ran = false, have_tables = false, tables_ready = false;
    DevBuf d_idx1, d_idx2;               // rsac_sim3_upload_from_views: KF1 / KF2 feature of every correspondence
    PinnedBuf h_idx;
    BatchDims d;
    std::vector<ProblemMeta> metas;
    DevBuf d_metas, d_x1, d_x2, d_s1, d_s2, d_c1, d_c2, d_tables, d_poses, d_counts, d_hmasks, d_results, d_masks, d_done;
    void release()
    {
        DevBuf* all[] = {&d_metas, &d_x1, &d_x2, &d_s1, &d_s2, &d_c1, &d_c2, &d_tables, &d_poses, &d_counts, &d_hmasks, &d_results, &d_masks, &d_done,
                         &d_idx1, &d_idx2};
        for (DevBuf* b : all) b->release();
        h_idx.release();
    }
};

struct PoseOptState {
    bool uploaded = false, ran = false;
    bool chained = false;      // edges are the PnP engine's correspondences (rsac_poseopt_from_pnp)
    int C = 0;
    int64_t total = 0;
    DevBuf d_metas, d_p3d, d_obs, d_isig, d_outlier, d_results, d_src, d_full;
    PinnedBuf h_metas;
    void release()
    {
        DevBuf* all[] = {&d_metas, &d_p3d, &d_obs, &d_isig, &d_outlier, &d_results, &d_src, &d_full};
        for (DevBuf* b : all) b->release();
        h_metas.release();
    }
};

struct Sim3OptState {
    bool uploaded = false, ran = false;
    bool chained = false;      // the batch was built on the device from the last SearchBySim3 run (rsac_sim3opt_from_search)
    int C = 0;
    int64_t total = 0;
    DevBuf d_metas, d_x1, d_x2, d_o1, d_o2, d_is1, d_is2, d_removed, d_results, d_src, d_full, d_nedges, d_K2;
    PinnedBuf h_metas;
    void release()
    {
        DevBuf* all[] = {&d_metas, &d_x1, &d_x2, &d_o1, &d_o2, &d_is1, &d_is2, &d_removed, &d_results, &d_src, &d_full, &d_nedges, &d_K2};
        for (DevBuf* b : all) b->release();
        h_metas.release();
    }
};

struct BowState {
    bool uploaded = false, ran = false;
    int C = 0, n_items = 0, mode = 0, check_orientation = 1;
    float nn_ratio = 0.6f;
    int64_t total_t = 0, total_q = 0;          // sizes of the per-pair target- / query-indexed arrays
    std::vector<int64_t> t2q_off, q2t_off;
    DevBuf d_sets, d_qset, d_tset, d_t2q_off, d_q2t_off, d_items, d_desc, d_angle, d_valid, d_node_ids, d_node_start, d_node_feat,
        d_t2q, d_q2t, d_bin, d_nmatches, d_mp_index;
    PinnedBuf h_stage;
    bool have_valid = false, have_mp_index = false, one_target = false;
    std::vector<int32_t> target_n_feat;        // [C] features of every pair's target set
    void release()
    {
        DevBuf* all[] = {&d_sets, &d_qset, &d_tset, &d_t2q_off, &d_q2t_off, &d_items, &d_desc, &d_angle, &d_valid, &d_node_ids,
                         &d_node_start, &d_node_feat, &d_t2q, &d_q2t, &d_bin, &d_nmatches, &d_mp_index};
        for (DevBuf* b : all) b->release();
        h_stage.release();
    }
};

struct KfdbState {
    bool db_ready = false, queries_ready = false, ran = false;
    int K = 0, K2 = 1, Q = 0, mode = 0;
    int64_t bm_words = 1, max_nq = 0;
    bool use_bitmap = false;
    DevBuf d_bitmap;
    DevBuf d_kf_off, d_kf_word, d_kf_val, d_covis, d_state;
    DevBuf d_q_off, d_q_word, d_q_val, d_min_score, d_conn_off, d_conn;
    DevBuf d_cw, d_wstar, d_si, d_eff, d_acc, d_best, d_firstpos, d_keys, d_out, d_min_common, d_best_acc, d_n_out;
    PinnedBuf h_stage;
    void release()
    {
        DevBuf* all[] = {&d_kf_off, &d_kf_word, &d_kf_val, &d_covis, &d_state, &d_q_off, &d_q_word, &d_q_val, &d_min_score, &d_conn_off,
                         &d_conn, &d_cw, &d_wstar, &d_si, &d_eff, &d_acc, &d_best, &d_firstpos, &d_keys, &d_out, &d_min_common,
                         &d_best_acc, &d_n_out, &d_bitmap};
        for (DevBuf* b : all) b->release();
        h_stage.release();
    }
};

struct GuidedState {
    bool uploaded = false, ran = false;
    int C = 0, n_views = 0, maxN1 = 0, maxN = 0;
    int64_t total1 = 0, total2 = 0;
    bool have_matched = false, have_scale = false;
    float th = 7.5f;
    // SearchByProjection(Frame, KeyFrame): shares the view buffers and the small per-pair arrays
    bool proj_uploaded = false, proj_ran = false, have_occupied = false, have_found = false;
    int proj_C = 0, orb_dist = 100, check_orientation = 1, cap = 16;
    int64_t totalF = 0, totalK = 0;
    DevBuf d_kp_angle, d_cand, d_cand_n, d_taken, d_minidx;
    PinnedBuf h_stage2;
    // host-side facts about the resident views (rsac_sim3_upload_from_views, batches that reference the resident views)
    bool views_have_angle = false;
    std::vector<int32_t> view_n_feat, view_feat_off;
    std::vector<uint8_t> view_mp_valid;     // concatenated
    DevBuf d_views, d_kp_xy, d_kp_octave, d_desc, d_mp_valid, d_mp_xyz, d_mp_desc, d_mp_maxdist, d_mp_mindist, d_grid_off, d_grid_idx,
        d_kf1, d_kf2, d_K, d_R12, d_t12, d_s12, d_off1, d_off2, d_matched_in, d_already1, d_already2, d_m1, d_m2, d_match12, d_n_found;
    PinnedBuf h_stage;
    void release()
    {
        DevBuf* all[] = {&d_views, &d_kp_xy, &d_kp_octave, &d_desc, &d_mp_valid, &d_mp_xyz, &d_mp_desc, &d_mp_maxdist, &d_mp_mindist,
                         &d_grid_off, &d_grid_idx, &d_kf1, &d_kf2, &d_K, &d_R12, &d_t12, &d_s12, &d_off1, &d_off2, &d_matched_in,
                         &d_already1, &d_already2, &d_m1, &d_m2, &d_match12, &d_n_found, &d_kp_angle, &d_cand, &d_cand_n, &d_taken, &d_minidx};
        for (DevBuf* b : all) b->release();
        h_stage.release();
        h_stage2.release();
    }
};

struct ProfPair {
    int stage;
    cudaEvent_t a, b;
};

}  // namespace rsac

struct rsac_engine {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t aux_stream = nullptr;           // fork/join side stream for independent preparation kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // per-stage profiling
    bool profile = false;
    std::vector<rsac::ProfPair> prof_events, prof_pool;
    rsac::ProfPair cur{};
    std::vector<std::pair<int, float>> prof_trace;   // (stage, ms) of every profiled launch since the last reset, in order
    double stage_ms[RSAC_STAGE_COUNT] = {0, 0, 0, 0, 0};
    int64_t stage_launches[RSAC_STAGE_COUNT] = {0, 0, 0, 0, 0};
    int64_t launches = 0;
    int32_t problem_base = 0;
    int32_t first_phase = 0;   // early exit: hypotheses per problem in phase A (0 = one solver wave)
    int32_t second_phase = 0;  // early exit: end of the second stage when the caller fixed it (rsac_set_phases)
    std::vector<int> stage_bounds;   // early exit: explicit stage boundaries (rsac_set_stages); empty = automatic
    rsac::PnpState pnp;
    rsac::PnpState mlpnp;
    rsac::ScoreState score;
    rsac::Sim3State sim3;
    rsac::PoseOptState poseopt;
    rsac::Sim3OptState sim3opt;
    rsac::BowState bow;
    rsac::KfdbState kfdb;
    rsac::GuidedState guided;
    rsac::DevBuf d_exact, d_scratch, d_resume, d_problem_ids;
    int32_t n_problem_ids = 0;                   // > 0: rsac_set_problem_ids is in force for batches of exactly this many problems
    uint64_t alloc_epoch = 0;                    // bumped by every device (re)allocation: captured graphs hold raw pointers
    bool graphs = true;                          // rsac_set_graphs / RSAC_GRAPH=0
    unsigned long long* last_exact = nullptr;   // diagnostic counter of the last scoring launch
    void* nccl_comm = nullptr;
    void* nccl_lib = nullptr;

    void stage_begin(int stage)
    {
        ++launches;
        if (!profile) return;
        if (!prof_pool.empty()) { cur = prof_pool.back(); prof_pool.pop_back(); }
        else { cudaEventCreate(&cur.a); cudaEventCreate(&cur.b); }
        cur.stage = stage;
        cudaEventRecord(cur.a, stream);
    }
    void stage_end(int)
    {
        if (!profile) return;
        cudaEventRecord(cur.b, stream);
        prof_events.push_back(cur);
    }
    void free_all()
    {
        pnp.release(); mlpnp.release(); score.release(); sim3.release(); poseopt.release(); sim3opt.release(); bow.release(); kfdb.release(); guided.release();
        d_exact.release(); d_scratch.release(); d_resume.release(); d_problem_ids.release();
    }
};

inline int rsac::DevBuf::ensure(rsac_engine* e, size_t bytes)
{
    if (bytes <= cap) return RSAC_OK;
    if (p) { cudaStreamSynchronize(e->stream); cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4;   // head-room so that ragged batches do not reallocate every call
    cudaError_t err = cudaMalloc(&p, want);
    if (err != cudaSuccess) {
        e->err = std::string("cudaMalloc: ") + cudaGetErrorString(err);
        cudaGetLastError();
        p = nullptr;
        return RSAC_ERR_ALLOC;
    }
    cap = want;
    ++e->alloc_epoch;
    return RSAC_OK;
}
