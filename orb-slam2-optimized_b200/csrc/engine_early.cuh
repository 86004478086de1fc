// engine_early.cuh -- host side of the staged early exit (RSAC_FLAG_EARLY_EXIT), shared by the PnP and the MLPnP engines.
// Device-side state and the flag kernel: early_exit.cuh; the stage structure is described there.
//
// A translation unit instantiates the templates below for its MODEL (0 PnPsolver, 1 MLPnPsolver) and supplies the two
// model-specific launches through EarlyHooks<MODEL>:
//     static int solve_range(e, s, list, list_count, lo, span, most)   minimal solves of hypotheses [lo, lo + span)
//     static int select(e, flags, d_resume, d_results_out, only_phase) the replay kernel
//     static int setup(e)                                              kernel attributes (outside a graph capture)
//     static int stage0_hpl()                                          hypotheses per lane of the stage-0 scoring plan (0: default)
//     static int stage_hpl(), stage_chunk_words()                      the same for the later stages; chunk size of every stage plan (0: default)
#pragma once
#include "engine_shared.cuh"
#include "early_exit.cuh"

template <int MODEL> struct EarlyHooks;

// scoring plans of the stages: [0, b0) with static work lists, [b(j-1), bj) and the clean-up range [b0, H) driven by
// device-side lists; their work arrays go H2D through pinned staging
template <int MODEL>
static int early_plan(rsac_engine* e, PnpState& s, const std::vector<int>& bounds)
{
    const BatchDims& d = s.d;
    cudaStream_t st = e->stream;
    const int K = (int)bounds.size();
    s.ee_bounds = bounds;
    s.ee_HA = bounds[0];
    s.ee_plans.assign(K + 1, ScorePlanPOD());
    s.ee_groups.assign(K + 1, std::vector<ScoreGroup>());
    if ((int)s.ee_visit.size() < K + 1) s.ee_visit.resize(K + 1);
    // stage 0: few hypotheses per problem -- one hypothesis per lane (two consumer warps per problem) measured
    // best for PnP (0.045 ms against 0.066 with two per lane at 1024 x 55)
    RSAC_TRY(plan_score<MODEL>(e, s.metas, d.maxH, s.ee_groups[0], s.ee_plans[0], 0, bounds[0],
                               env_int("RSAC_EE_HPL_A", EarlyHooks<MODEL>::stage0_hpl()), env_int("RSAC_EE_CW_A", EarlyHooks<MODEL>::stage_chunk_words())));
    const int hplB = env_int("RSAC_EE_HPL_B", EarlyHooks<MODEL>::stage_hpl()), cwB = env_int("RSAC_EE_CW_B", EarlyHooks<MODEL>::stage_chunk_words());
    for (int j = 1; j < K; ++j)
        RSAC_TRY(plan_score<MODEL>(e, s.metas, d.maxH, s.ee_groups[j], s.ee_plans[j], bounds[j - 1], bounds[j], hplB, cwB, true));
    RSAC_TRY(plan_score<MODEL>(e, s.metas, d.maxH, s.ee_groups[K], s.ee_plans[K], bounds[0], INT32_MAX, hplB, cwB, true));   // clean-up
    std::vector<size_t> off(K + 2, 0);
    for (int i = 0; i <= K; ++i) off[i + 1] = (off[i] + sizeof(ScoreGroup) * s.ee_plans[i].work.size() + 255) & ~(size_t)255;
    char* h = (char*)s.h_stageEE.ensure(off[K + 1] + 256);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    for (int i = 0; i <= K; ++i) {
        const size_t bytes = sizeof(ScoreGroup) * s.ee_plans[i].work.size();
        RSAC_TRY(s.ee_visit[i].ensure(e, std::max<size_t>(bytes, sizeof(ScoreGroup))));
        memcpy(h + off[i], s.ee_plans[i].work.data(), bytes);
        RSAC_CUDA(e, cudaMemcpyAsync(s.ee_visit[i].p, h + off[i], bytes, cudaMemcpyHostToDevice, st));
    }
    s.h_stageEE.mark(st);
    s.ee_planned = true;
    ++s.plan_version;
    return RSAC_OK;
}

static int early_flag(rsac_engine* e, PnpState& s, int stage, int lim, int mode)
{
    e->stage_begin(RSAC_STAGE_RNG);
    early_exit_flag_kernel<<<(s.d.C + 3) / 4, 128, 0, e->stream>>>((const ProblemMeta*)s.d_metas.p, s.d.C, (const int32_t*)s.d_counts.p,
                                                                  stage, lim, (int32_t*)s.d_ee.p, mode);
    e->stage_end(RSAC_STAGE_RNG);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

// minimal solves + scoring of hypotheses [lo, hi) of the problems in `list` (device-side count); plan index `pi` is
// the list-driven scoring plan of that range
template <int MODEL>
static int early_range(rsac_engine* e, PnpState& s, const int32_t* list, const int32_t* list_count, int lo, int hi, int pi)
{
    const BatchDims& d = s.d;
    const int span = std::min(hi, d.maxH) - lo;
    if (span <= 0) return RSAC_OK;
    RSAC_TRY(EarlyHooks<MODEL>::solve_range(e, s, list, list_count, lo, span, (int64_t)d.C * span));
    ScoreArgs sa = s.ee_sa;
    sa.list = list;
    sa.list_count = list_count;
    return launch_score<MODEL>(e, sa, s.ee_plans[pi], (int)s.ee_groups[pi].size(), s.ee_visit[pi]);
}

// the launches of one staged sweep (two memsets + 4 + 3 (K - 1) + 4 kernels); every buffer exists and every plan is on
// the device: nothing here allocates, copies or synchronises, so the sequence can be captured into a CUDA graph
template <int MODEL>
static int early_issue(rsac_engine* e, PnpState& s, int flags, void* d_results_out, const std::vector<int>& bounds)
{
    const BatchDims& d = s.d;
    cudaStream_t st = e->stream;
    const ProblemMeta* metas = (const ProblemMeta*)s.d_metas.p;
    const int K = (int)bounds.size();
    RSAC_CUDA(e, cudaMemsetAsync(s.d_ee.p, 0, sizeof(int32_t) * early_exit_words(d.C), st));
    const EarlyExit v = early_exit_view((int32_t*)s.d_ee.p, d.C);

    ScoreArgs sa;
    RSAC_TRY(zero_score_region(e, s.d_counts, d.sumH, 0, sa));
    sa.metas = metas;
    sa.cP = (const float4*)s.d_cP.p; sa.cC = (const float4*)s.d_uv.p;
    sa.poses = s.d_poses.p;
    sa.hmasks = nullptr;
    if (flags & RSAC_FLAG_KEEP_MASKS) {
        RSAC_TRY(s.d_hmasks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_hwords, 1)));
        sa.hmasks = (uint32_t*)s.d_hmasks.p;
    }
    s.ee_sa = sa;

    // stage 0: hypotheses [0, b0) of every problem
    {
        RSAC_TRY(EarlyHooks<MODEL>::solve_range(e, s, nullptr, nullptr, 0, bounds[0], (int64_t)d.C * bounds[0]));
        RSAC_TRY(launch_score<MODEL>(e, sa, s.ee_plans[0], (int)s.ee_groups[0].size(), s.ee_visit[0]));
    }
    // who goes on after stage j-1 (list j); stage j: [b(j-1), bj) of list j
    for (int j = 1; j < K; ++j) {
        RSAC_TRY(early_flag(e, s, j - 1, bounds[j - 1], 0));
        RSAC_TRY(early_range<MODEL>(e, s, v.list(j), v.counters + j, bounds[j - 1], bounds[j], j));
    }
    RSAC_TRY(early_flag(e, s, K - 1, bounds[K - 1], 0));     // the members of the last list have everything
    // replay; problems it cannot decide go to the clean-up
#ifdef RSAC_TIMING_EXPERIMENTS
    // developer build only (-DRSAC_TIMING_EXPERIMENTS, scripts/flight_probe.py): what the replay costs when sweeps overlap --
    // measured in round 2: 0.212 ms per sweep without it, 0.383 ms with it, six sweeps in flight (the results are wrong without it)
    if (env_int("RSAC_DBG_SKIP_SELECT", 0)) return RSAC_OK;
#endif
    RSAC_TRY(EarlyHooks<MODEL>::select(e, flags, nullptr, d_results_out, -1));
    RSAC_TRY(early_range<MODEL>(e, s, v.listC, v.counters + kCleanupCounter, bounds[0], d.maxH, K));
    RSAC_TRY(EarlyHooks<MODEL>::select(e, flags, nullptr, d_results_out, 2));
    return RSAC_OK;
}

template <int MODEL>
static int early_run(rsac_engine* e, PnpState& s, int flags, void* d_results_out, const std::vector<int>& bounds)
{
    const BatchDims& d = s.d;
    if (!s.ee_planned || s.ee_bounds != bounds) RSAC_TRY(early_plan<MODEL>(e, s, bounds));   // normally done by the upload
    s.ee_mode = true;
    s.ee_complete = false;
    // everything a sweep touches exists before the first launch (a captured sequence must not allocate)
    RSAC_TRY(s.d_ee.ensure(e, sizeof(int32_t) * early_exit_words(d.C)));
    {
        const size_t n_al = ((size_t)std::max<int64_t>(d.sumH, 1) + 1) & ~(size_t)1;
        RSAC_TRY(s.d_counts.ensure(e, (n_al + 2) * sizeof(int32_t)));
    }
    if (flags & RSAC_FLAG_KEEP_MASKS) RSAC_TRY(s.d_hmasks.ensure(e, sizeof(uint32_t) * (size_t)std::max<int64_t>(d.total_hwords, 1)));
    RSAC_TRY(EarlyHooks<MODEL>::setup(e));

    // CUDA graph: the staged sweep is 18 dependent launches (K = 4) of 5-150 us each.  The first run of a key is eager;
    // the second captures the same call sequence; later runs are one cudaGraphLaunch.  The key holds every value that
    // ends up in a kernel argument or a launch shape; device contents (metas, tables, plans) are read at run time
    std::vector<int64_t> key = {(int64_t)flags, (int64_t)(intptr_t)d_results_out, (int64_t)e->alloc_epoch, (int64_t)s.plan_version,
                                (int64_t)e->problem_base, (int64_t)e->n_problem_ids, (int64_t)d.C, d.sumH, (int64_t)d.maxH, (int64_t)d.maxWords,
                                (int64_t)s.have_cov};
    for (int b : bounds) key.push_back(b);
    const bool want_graph = e->graphs && !e->profile && d.C > 0;
    if (want_graph && s.graph && key == s.graph_key) {
        RSAC_CUDA(e, cudaGraphLaunch(s.graph, e->stream));
        e->launches += s.graph_nodes;
        const size_t n_al = ((size_t)std::max<int64_t>(d.sumH, 1) + 1) & ~(size_t)1;   // host-side state the eager path leaves behind
        e->last_exact = (unsigned long long*)((int32_t*)s.d_counts.p + n_al);
        s.ran = true;
        return RSAC_OK;
    }
    if (want_graph && key == s.graph_key && s.eager_runs >= 1) {
        if (s.graph) { cudaGraphExecDestroy(s.graph); s.graph = nullptr; }
        cudaGraph_t g = nullptr;
        const int64_t l0 = e->launches;
        RSAC_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed));
        const int rc = early_issue<MODEL>(e, s, flags, d_results_out, bounds);
        const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        const int64_t nodes = e->launches - l0;
        e->launches = l0;
        if (rc == RSAC_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&s.graph, g, 0) == cudaSuccess) {
            cudaGraphDestroy(g);
            s.graph_nodes = nodes;
            RSAC_CUDA(e, cudaGraphLaunch(s.graph, e->stream));
            e->launches += nodes;
            s.ran = true;
            return RSAC_OK;
        }
        // capture failed: stay eager for this engine
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        s.graph = nullptr;
        e->graphs = false;
    }
    if (key != s.graph_key) {
        if (s.graph) { cudaGraphExecDestroy(s.graph); s.graph = nullptr; }
        s.graph_key = key;
        s.eager_runs = 0;
    }
    RSAC_TRY(early_issue<MODEL>(e, s, flags, d_results_out, bounds));
    ++s.eager_runs;
    s.ran = true;
    return RSAC_OK;
}

// before a later iterate() call resumes a batch whose last run stopped early: problems that were decided inside the
// hypotheses they had get the rest now, so that the scan can go on wherever the caller resumes it
template <int MODEL>
static int early_complete(rsac_engine* e, PnpState& s)
{
    if (!s.ee_mode || s.ee_complete) return RSAC_OK;
    const EarlyExit v = early_exit_view((int32_t*)s.d_ee.p, s.d.C);
    RSAC_CUDA(e, cudaMemsetAsync(v.counters + kCleanupCounter, 0, sizeof(int32_t), e->stream));
    RSAC_TRY(early_flag(e, s, 0, 0, 2));
    RSAC_TRY(early_range<MODEL>(e, s, v.listC, v.counters + kCleanupCounter, s.ee_HA, s.d.maxH, (int)s.ee_bounds.size()));
    s.ee_complete = true;
    return RSAC_OK;
}

// out[0] = first stage used (0: the run was exhaustive), out[1] = |list 1|, out[2] = |clean-up list|, out[3] = hypotheses
// solved and scored in total (synchronises)
static int early_stats(rsac_engine* e, PnpState& s, int64_t out[4])
{
    out[0] = out[1] = out[2] = 0;
    out[3] = s.d.sumH;
    if (!s.ee_mode) return RSAC_OK;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    std::vector<int32_t> ee(early_exit_words(s.d.C));
    RSAC_CUDA(e, cudaMemcpyAsync(ee.data(), s.d_ee.p, sizeof(int32_t) * ee.size(), cudaMemcpyDeviceToHost, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    out[0] = s.ee_HA;
    out[1] = ee[4 * (size_t)s.d.C + 1];
    out[2] = ee[4 * (size_t)s.d.C + kCleanupCounter];
    int64_t done = 0;
    for (int c = 0; c < s.d.C; ++c) done += std::min(s.metas[c].H, std::max(ee[c], 0));   // (clean-up problems end with upto = H)
    out[3] = done;
    return RSAC_OK;
}

// stage boundaries b0 < b1 < ... < b(K-1) = maxH from the caller's choice (rsac_set_stages / rsac_set_phases / `env`), else
// b0 = first_auto and every further stage doubles what exists; sanitised: strictly increasing, inside (0, maxH], last = maxH
static std::vector<int> early_stage_bounds(rsac_engine* e, const BatchDims& d, int first_auto, const char* env_name)
{
    std::vector<int> b;
    const char* env = getenv(env_name);
    if (!e->stage_bounds.empty()) {
        b = e->stage_bounds;
    } else if (env && *env) {
        for (const char* p = env; *p;) {
            b.push_back(atoi(p));
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
    } else {
        b.push_back(first_auto);
        if (e->second_phase > 0) b.push_back(e->second_phase);
        else
            while (b.back() < d.maxH && (int)b.size() < kMaxStages - 1) b.push_back(b.back() * 2);
    }
    std::vector<int> out;
    for (int v : b) {
        v = std::min(v, d.maxH);
        if (v <= 0 || (!out.empty() && v <= out.back())) continue;
        out.push_back(v);
        if (v >= d.maxH || (int)out.size() == kMaxStages - 1) break;
    }
    if (out.empty() || out.back() < d.maxH) out.push_back(std::max(d.maxH, 1));
    return out;
}
