// engine_shared.cuh -- host-side helpers shared by the solver translation units (PnP, MLPnP, Sim3, the optimisers):
// kernel-attribute bookkeeping, the scoring planner and launcher, batch bookkeeping.  Everything is `static`:
// each translation unit gets its own copy (kernels are per translation unit too).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ransac_b200.h"
#include "common.cuh"
#include "engine_state.cuh"
#include "rng.cuh"
#include "score.cuh"

using namespace rsac;

// --------------------------------------------------------- score planning
// CTA shape for the scoring kernel: HPL hypotheses per lane (template), `warps` consumer warps per CTA
// plus one producer warp.  RSAC_SCORE_HPL / RSAC_SCORE_WARPS / RSAC_SCORE_CTAS / RSAC_SCORE_CW override the
// planner (tuning sweeps only).
// dynamic shared memory above 32 KB is always opted in (cudaFuncAttributeMaxDynamicSharedMemorySize): the 48 KB default
// limit counts the kernel's static shared memory too, so a request of exactly 48 KB fails without it
static constexpr int kChunkWordsMax = 8;     // 256 correspondences per ring slot (12 KB), 4 slots

// Kernel attributes are per device and process-wide, not per engine: engines on different host threads (Tracking and
// LoopClosing each own one) must not lower what another engine's launch relies on -- monotonic maxima under a lock
static int set_func_attr_max(rsac_engine* e, const void* kern, cudaFuncAttribute attr, int value)
{
    static std::mutex mu;
    static std::map<std::tuple<int, const void*, int>, int> cur;
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_tuple(e->device, kern, (int)attr);
    const auto it = cur.find(key);
    if (it != cur.end() && it->second >= value) return RSAC_OK;
    RSAC_CUDA(e, cudaFuncSetAttribute(kern, attr, value));
    cur[key] = value;
    return RSAC_OK;
}

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

using ScorePlan = ScorePlanPOD;

// engine_pnp.cu: materialises the flat p3d / p2d / sigma2 arrays of an indexed PnP batch (no-op for flat uploads)
int rsac_internal_pnp_ensure_flat(rsac_engine* e);

template <int MODEL>
static size_t score_smem_bytes(int chunk_cap, int tile_hyps)
{
    return (size_t)chunk_cap * 48 * kScoreStages + (size_t)tile_hyps * 12 * sizeof(typename ScoreModel<MODEL>::pose_t);
}

template <int MODEL>
static const void* score_kernel_ptr(int hpl)
{
    switch (hpl) {
        case 1: return (const void*)score_kernel<1, MODEL>;
        case 3: return (const void*)score_kernel<3, MODEL>;
        default: return (const void*)score_kernel<2, MODEL>;
    }
}

// Builds the work groups (problem x hypothesis tile), the chunking and each CTA's list of group records.
//  * many groups (a relocalisation sweep): CTAs own groups round-robin; a chunk is up to 256
//    correspondences; the producer warp prefetches the next group's chunks while the current one is scored;
//  * few groups (scoring stress: 8 tiles x 10 000 correspondences): the CTAs are dealt to the groups in
//    proportion to their work and pull small chunks from the group's counter, which balances the SMs to
//    within one chunk of work.
template <int MODEL>
static int plan_score(rsac_engine* e, const std::vector<ProblemMeta>& metas, int maxH, std::vector<ScoreGroup>& groups, ScorePlan& pl,
                      int h_lo = 0, int h_hi = INT32_MAX, int hpl_want = 0, int cw_want = 0, bool by_list = false)
{
    // [h_lo, min(H, h_hi)) of every problem: the early-exit phases score hypothesis ranges (pnp_run_early)
    groups.clear();
    pl = ScorePlan();
    pl.hpl = hpl_want > 0 ? hpl_want : env_int("RSAC_SCORE_HPL", 2);
    if (pl.hpl != 1 && pl.hpl != 2 && pl.hpl != 3) pl.hpl = 2;
    if (h_lo > 0 || h_hi != INT32_MAX) {
        maxH = 0;
        for (const auto& m : metas) maxH = std::max(maxH, std::min(m.H, h_hi) - h_lo);
    }
    int warps = (maxH + 32 * pl.hpl - 1) / (32 * pl.hpl);
    // at most 8 consumer warps per CTA (two CTAs per SM): alone it scores cfg5 like one 16-warp CTA per SM (46-47 % of
    // the FP32 peak), with independent scoring jobs in flight it is ahead (64 % against 59 %: a CTA that waits for its
    // first chunk or drains its last one has a neighbour on the SM)
    warps = std::max(1, std::min(env_int("RSAC_SCORE_WARPS", 8), std::min(16, warps)));
    pl.threads = warps * 32;
    pl.tile_hyps = warps * 32 * pl.hpl;
    std::vector<double> work;
    for (size_t p = 0; p < metas.size(); ++p) {
        const ProblemMeta& m = metas[p];
        if (m.n <= 0 || m.H <= 0) continue;
        const int h_end = std::min(m.H, h_hi);
        for (int h0 = h_lo; h0 < h_end; h0 += pl.tile_hyps) {
            ScoreGroup g;
            memset(&g, 0, sizeof(g));
            g.gid = (int32_t)groups.size();
            g.problem = (int)p; g.hyp0 = h0;
            g.corr_off = m.corr_off; g.n = m.n; g.words = m.words; g.H = h_end;
            g.hyp_off = m.hyp_off; g.word_off = m.word_off; g.hmask_off = m.hmask_off;
            if (MODEL == 0) { g.fx = (float)m.fx; g.fy = (float)m.fy; } else { g.fx = m.k1[0]; g.fy = m.k1[1]; }
            groups.push_back(g);
            work.push_back((double)m.words * std::min(pl.tile_hyps, h_end - h0));
        }
    }
    const int NG = (int)groups.size();
    ScoreGroup end_rec;
    memset(&end_rec, 0, sizeof(end_rec));
    end_rec.gid = -1;
    if (NG == 0) { pl.work.assign(1, end_rec); pl.vlen = 1; pl.grid = 1; return RSAC_OK; }
    const void* kern = score_kernel_ptr<MODEL>(pl.hpl);
    auto resident = [&](int chunk_words) -> int {
        const size_t smem = score_smem_bytes<MODEL>(chunk_words * 32, pl.tile_hyps);
        if (smem > 32 * 1024) (void)set_func_attr_max(e, kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, pl.threads + 32, smem) != cudaSuccess) { cudaGetLastError(); nb = 1; }
        nb = std::max(1, nb);
        const int cap_env = env_int("RSAC_SCORE_CTAS", 0);
        if (cap_env > 0) nb = std::min(nb, cap_env);
        return nb * e->sm_count;
    };
    int maxWords = 1;
    for (const auto& m : metas) maxWords = std::max(maxWords, m.words);
    int cw = std::min(kChunkWordsMax, maxWords);
    if (cw_want > 0) cw = std::min(cw, cw_want);
    int slots = resident(cw);
    if (by_list) {
        // which problems are scored is decided on the device (early-exit phases B, C): one record per
        // (problem, tile), whole groups per CTA, the CTAs stride over the device-side list
        const int T = std::max(1, (maxH + pl.tile_hyps - 1) / pl.tile_hyps);
        pl.by_list = true;
        pl.tiles = T;
        pl.vlen = T;
        pl.grid = slots;
        pl.work.assign(metas.size() * (size_t)T, end_rec);
        for (auto& g : groups) {
            g.chunk_words = std::min(cw, std::max(1, g.words));
            g.nchunks = (g.words + g.chunk_words - 1) / g.chunk_words;
            g.first_stride = 1 << 16;
            pl.work[(size_t)g.problem * T + (g.hyp0 - h_lo) / pl.tile_hyps] = g;
        }
        pl.chunk_cap = cw * 32;
        pl.smem = score_smem_bytes<MODEL>(pl.chunk_cap, pl.tile_hyps);
        return RSAC_OK;
    }
    std::vector<std::vector<int>> lists;
    if (NG >= slots) {
        pl.grid = slots;
        lists.assign(pl.grid, {});
        for (int g = 0; g < NG; ++g) lists[g % pl.grid].push_back(g);
    } else {
        // few groups: deal the CTAs to the groups in proportion to their work, small chunks
        cw = 1;
        slots = resident(cw);
        double total = 0;
        for (double w : work) total += w;
        const int cw_env = env_int("RSAC_SCORE_CW", 0);
        if (cw_env > 0) cw = std::min(kChunkWordsMax, cw_env);
        else cw = std::max(1, std::min(kChunkWordsMax, (int)(maxWords / (3.0 * std::max(1, slots / NG)))));   // cfg5: 2 words
        slots = resident(cw);
        pl.grid = std::max(slots, NG);
        // largest-remainder apportionment, at least one CTA per group
        std::vector<int> share(NG, 1);
        const int left = pl.grid - NG;
        std::vector<double> frac(NG);
        for (int g = 0; g < NG; ++g) {
            const double ideal = work[g] / total * left;
            share[g] += (int)ideal;
            frac[g] = ideal - (int)ideal;
        }
        int used = 0;
        for (int g = 0; g < NG; ++g) used += share[g];
        while (used < pl.grid) {
            int best = 0;
            for (int g = 1; g < NG; ++g) if (frac[g] > frac[best]) best = g;
            share[best]++; frac[best] = -1; used++;
        }
        // interleave so that the CTAs of one group spread over the SMs
        lists.assign(pl.grid, {});
        std::vector<int> rem = share;
        int b = 0;
        while (b < pl.grid)
            for (int g = 0; g < NG && b < pl.grid; ++g)
                if (rem[g] > 0) { lists[b++].push_back(g); rem[g]--; }
    }
    for (auto& g : groups) {
        g.chunk_words = std::min(cw, std::max(1, g.words));
        g.nchunks = (g.words + g.chunk_words - 1) / g.chunk_words;
    }
    size_t vlen = 1;
    for (const auto& l : lists) vlen = std::max(vlen, l.size());
    pl.vlen = (int)vlen;
    pl.work.assign((size_t)pl.grid * vlen, end_rec);
    // chunks of a group are dealt round-robin to the CTAs that visit it (one CTA per group in the many-groups case)
    std::vector<int> visitors(NG, 0), seen(NG, 0);
    for (int b = 0; b < pl.grid; ++b)
        for (int g : lists[b]) visitors[g]++;
    for (int b = 0; b < pl.grid; ++b)
        for (size_t k = 0; k < lists[b].size(); ++k) {
            const int g = lists[b][k];
            ScoreGroup rec = groups[g];
            rec.first_stride = (seen[g]++ & 0xffff) | (std::min(visitors[g], 0x7fff) << 16);
            pl.work[(size_t)b * vlen + k] = rec;
        }
    pl.chunk_cap = cw * 32;
    pl.smem = score_smem_bytes<MODEL>(pl.chunk_cap, pl.tile_hyps);
    return RSAC_OK;
}

// counts and the diagnostic counter live in ONE buffer so that a single memset node prepares a scoring
// launch: [counts: n ints][exact: 2 ints]
static int zero_score_region(rsac_engine* e, DevBuf& d_counts, int64_t n_counts, int /*ngroups*/, ScoreArgs& sa)
{
    const size_t n_al = ((size_t)std::max<int64_t>(n_counts, 1) + 1) & ~(size_t)1;     // keep the 8-byte counter aligned
    const size_t total = (n_al + 2) * sizeof(int32_t);
    RSAC_TRY(d_counts.ensure(e, total));
    RSAC_CUDA(e, cudaMemsetAsync(d_counts.p, 0, total, e->stream));
    sa.counts = (int32_t*)d_counts.p;
    sa.exact_counter = (unsigned long long*)((int32_t*)d_counts.p + n_al);
    e->last_exact = sa.exact_counter;
    return RSAC_OK;
}

template <int MODEL>
static int launch_score(rsac_engine* e, ScoreArgs& args, const ScorePlan& pl, int ngroups, DevBuf& d_visit)
{
    if (ngroups <= 0) return RSAC_OK;
    const void* kern = score_kernel_ptr<MODEL>(pl.hpl);
    if (pl.smem > 32 * 1024) RSAC_TRY(set_func_attr_max(e, kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    args.work = (const ScoreGroup*)d_visit.p;
    args.vlen = pl.vlen;
    args.tiles_per_problem = pl.tiles;
    if (pl.by_list != (args.list != nullptr)) { e->err = "scoring plan and launch disagree about list mode"; return RSAC_ERR_STATE; }
    args.chunk_cap = pl.chunk_cap;
    args.tile_hyps = pl.tile_hyps;
    void* kargs[] = {&args};
    e->stage_begin(RSAC_STAGE_SCORE);
    cudaError_t err = cudaLaunchKernel(kern, dim3(pl.grid), dim3(pl.threads + 32), kargs, pl.smem, e->stream);   // + the producer warp
    e->stage_end(RSAC_STAGE_SCORE);
    if (err != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "score launch (grid %d, threads %d, smem %zu, hpl %d, list %d): %s", pl.grid, pl.threads + 32,
                 pl.smem, pl.hpl, (int)pl.by_list, cudaGetErrorString(err));
        e->err = buf;
        cudaGetLastError();
        return RSAC_ERR_CUDA;
    }
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}

static int pnp_build_metas(rsac_engine* e, int C, const int32_t* offsets, const rsac_ransac_params* params, int n_params,
                           const uint32_t* seeds, const int64_t* table_offsets, bool have_tables,
                           std::vector<ProblemMeta>& metas, std::vector<float>& th2, BatchDims& d)
{
    if (have_tables && !table_offsets) { e->err = "tables without table_offsets"; return RSAC_ERR_INVALID; }
    metas.assign(C, ProblemMeta());
    th2.assign(C, 0.f);
    d = BatchDims();
    d.C = C;
    for (int c = 0; c < C; ++c) {
        const rsac_ransac_params& p = params[n_params == 1 ? 0 : c];
        ProblemMeta& m = metas[c];
        memset(&m, 0, sizeof(m));
        m.corr_off = offsets[c];
        m.n = offsets[c + 1] - offsets[c];
        if (m.n < 0 || p.min_set < 1 || p.min_set > 8) { e->err = "bad offsets or min_set"; return RSAC_ERR_INVALID; }
        int minInl = 0, H = 1;
        if (m.n > 0) rsac_pnp_ransac_setup(m.n, &p, &minInl, &H); else { minInl = std::max(p.min_inliers, p.min_set); H = 0; }
        if (m.n < minInl || m.n < p.min_set) H = 0;       // iterate() returns at once (PnPsolver.cpp:110-114)
        m.H = H;
        m.min_inl = minInl;
        m.min_set = p.min_set;
        m.hyp_off = (int32_t)d.sumH;
        m.words = (m.n + 31) / 32;
        m.word_off = (int32_t)d.total_words;
        m.hmask_off = d.total_hwords;
        m.seed = seeds ? seeds[c] : 0u;
        if (have_tables) {
            m.table_off = table_offsets[c];
            if (table_offsets[c + 1] - table_offsets[c] < (int64_t)H * p.min_set) { e->err = "index table too short"; return RSAC_ERR_INVALID; }
        } else {
            m.table_off = d.table_len;
        }
        th2[c] = p.th2;
        d.table_len += (int64_t)H * p.min_set;
        d.sumH += H;
        d.total_words += m.words;
        d.total_hwords += (int64_t)H * m.words;
        d.maxH = std::max(d.maxH, H);
        d.maxN = std::max(d.maxN, m.n);
        d.maxWords = std::max(d.maxWords, m.words);
    }
    d.total = offsets[C];
    if (have_tables) d.table_len = table_offsets[C];
    // hyp_off / word_off are 32-bit on the device: refuse batches whose offsets would wrap
    if (d.sumH > INT32_MAX || d.total_words > INT32_MAX || (int64_t)d.total * 16 > INT32_MAX) { e->err = "batch too large for 32-bit offsets"; return RSAC_ERR_INVALID; }
    return RSAC_OK;
}

// metas, thresholds and tiles go through one pinned staging buffer so that the H2D copies are truly
// asynchronous (the host never waits for the stream's earlier sweeps)
static int stage_small_tables(rsac_engine* e, PnpState& s, const std::vector<float>& th2, bool with_plan = true)
{
    const BatchDims& d = s.d;
    const size_t b_meta = sizeof(ProblemMeta) * (size_t)d.C, b_th = sizeof(float) * (size_t)d.C;
    const size_t b_work = sizeof(ScoreGroup) * s.plan.work.size();
    const size_t o_th = (b_meta + 255) & ~(size_t)255, o_work = (o_th + b_th + 255) & ~(size_t)255;
    char* h = (char*)s.h_stage.ensure(o_work + b_work + 256);
    if (!h) { e->err = "cudaHostAlloc failed"; return RSAC_ERR_ALLOC; }
    if (d.C > 0) {
        memcpy(h, s.metas.data(), b_meta);
        memcpy(h + o_th, th2.data(), b_th);
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_metas.p, h, b_meta, cudaMemcpyHostToDevice, e->stream));
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_th2.p, h + o_th, b_th, cudaMemcpyHostToDevice, e->stream));
    }
    if (with_plan) {
        RSAC_TRY(s.d_visit.ensure(e, std::max<size_t>(b_work, sizeof(ScoreGroup))));
        memcpy(h + o_work, s.plan.work.data(), b_work);
        RSAC_CUDA(e, cudaMemcpyAsync(s.d_visit.p, h + o_work, b_work, cudaMemcpyHostToDevice, e->stream));
    }
    s.h_stage.mark(e->stream);
    return RSAC_OK;
}

