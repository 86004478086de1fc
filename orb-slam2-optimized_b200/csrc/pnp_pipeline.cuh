// pnp_pipeline.cuh -- PnPsolver::iterate as a batched device pipeline.
//
// Reference control flow: src/PnPsolver.cpp:102-191 (iterate; the `||` at :119 makes the
// first call consume the whole budget), :193-238 (Refine).  All H hypotheses of all
// problems are solved and scored in parallel; the sequential semantics (first strict
// maximum, refine on the prefix-best set, first refine with > minInliers wins, otherwise
// the unrefined best) are then replayed per problem.
#pragma once
#include "common.cuh"
#include "epnp.cuh"
#include "rng.cuh"
#include "score.cuh"

namespace rsac {

// ---- minimal-set tables from per-problem seeds: one thread per problem ----
__global__ void rng_tables_kernel(const ProblemMeta* metas, int C, uint32_t* tables)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= C) return;
    const ProblemMeta& m = metas[p];
    if (m.n < m.min_set) return;
    GlibcRand g;
    g.seed(m.seed);
    uint32_t* out = tables + m.table_off;
    for (int h = 0; h < m.H; ++h) draw_minimal_set<8>(g, m.n, m.min_set, out + (size_t)h * m.min_set);
}

__device__ __forceinline__ int find_problem(const ProblemMeta* metas, int C, int64_t g)
{
    int lo = 0, hi = C - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)metas[mid].hyp_off <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- EPnP minimal solve: one thread per hypothesis (PnPsolver.cpp:125-141) ----
__global__ void __launch_bounds__(128) epnp_minimal_kernel(const ProblemMeta* metas, int C, int64_t sumH,
                                                           const uint32_t* tables, const float4* cA,
                                                           const float2* uv, float* poses)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sumH) return;
    const int p = find_problem(metas, C, g);
    const ProblemMeta& m = metas[p];
    const int h = (int)(g - m.hyp_off);
    const uint32_t* idx = tables + m.table_off + (size_t)h * 4;
    double pw[12], us[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t ci = (size_t)m.corr_off + idx[i];
        const float4 a = cA[ci];
        const float2 q = uv[ci];
        pw[3 * i] = (double)a.x; pw[3 * i + 1] = (double)a.y; pw[3 * i + 2] = (double)a.z;   // add_correspondence (:288-294)
        us[2 * i] = (double)q.x; us[2 * i + 1] = (double)q.y;
    }
    const Cam k = {m.fx, m.fy, m.cx, m.cy};
    float R[9], t[3];
    epnp_compute_pose_small<4>(pw, us, k, R, t);
    float* out = poses + g * 12;
#pragma unroll
    for (int i = 0; i < 9; ++i) out[i] = R[i];
    out[9] = t[0]; out[10] = t[1]; out[11] = t[2];
}

// ---- replay + refine: one CTA per problem ----
struct SelectArgs {
    const ProblemMeta* metas;
    const float4* cA;
    const float4* cB;
    const float2* uv;
    const float* poses;      // [sumH][12]
    const int32_t* counts;   // [sumH]
    // scratch, indexed like the correspondences
    uint32_t* sel;           // compacted indices of the best set
    double* pw_s;            // [total][3]
    double* us_s;            // [total][2]
    double* al_s;            // [total][4]
    double2* rec;            // [C][kMaxSweepsRec*66] recorded Jacobi rotations of the refine solve
    // outputs
    void* results;           // rsac_result[C] (layout in ransac_b200.h)
    void* results2;          // optional second copy (collective send buffer)
    uint32_t* masks;         // final masks, word_off per problem
    int32_t problem_base;    // global index of problem 0 (sharding)
};

struct ResultRec {   // mirrors rsac_result (include/ransac_b200.h)
    int32_t ok, no_more, n_inliers, best_hyp, refined, n_refines, best_count, n_hyp;
    float R[9], t[3], s;
    int32_t problem, reserved[2];
};
static_assert(sizeof(ResultRec) == 96, "rsac_result layout");

constexpr int kSelectThreads = 128;

// exact CheckInliers of one pose over all correspondences of the problem by the whole CTA:
// words -> mask (shared or global), returns the count in *s_cnt (shared)
__device__ inline void cta_score_exact_pnp(const ProblemMeta* m, const SelectArgs& a, const float* pose,
                                           uint32_t* mask_out, int* s_cnt)
{
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int local = 0;
    for (int base = (threadIdx.x >> 5) * 32; base < m->words * 32; base += blockDim.x) {
        const int i = base + lane;
        bool in = false;
        if (i < m->n) {
            const size_t g = (size_t)m->corr_off + i;
            const float4 c = a.cA[g];
            const float2 q = a.uv[g];
            in = pnp_exact_inlier(pose, c.x, c.y, c.z, q.x, q.y, a.cB[g].y, m);
        }
        const uint32_t word = __ballot_sync(0xffffffffu, in);
        if (lane == 0) { mask_out[base >> 5] = word; local += __popc(word); }
    }
    if (lane == 0 && local) atomicAdd(s_cnt, local);
    __syncthreads();
}

__global__ void __launch_bounds__(kSelectThreads) pnp_select_kernel(SelectArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ProblemMeta* m = a.metas + blockIdx.x;
    const int tid = threadIdx.x;
    const int words = m->words;
    uint32_t* bestmask = reinterpret_cast<uint32_t*>(smem_raw);   // [words]
    uint32_t* refmask = bestmask + words;                         // [words]
    int* prefix = reinterpret_cast<int*>(refmask + words);        // [words+1]

    __shared__ int s_found, s_cnt;
    __shared__ double s_C0[3], s_A[9], s_cws[12], s_CCi[9], s_MtM[78], s_U4[48], s_betas[12];
    __shared__ double s_ccs[3][12], s_sign[3], s_pc0[3][3], s_pw0[3], s_M[3][9], s_R[3][9], s_t[3][3], s_rep[3];
    __shared__ float s_pose[12], s_bestpose[12];

    ResultRec res;
    res.ok = 0; res.no_more = 0; res.n_inliers = 0; res.best_hyp = -1; res.refined = 0; res.n_refines = 0;
    res.best_count = 0; res.n_hyp = 0;
    for (int i = 0; i < 9; ++i) res.R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    res.t[0] = res.t[1] = res.t[2] = 0.0f; res.s = 1.0f;
    res.problem = a.problem_base + blockIdx.x; res.reserved[0] = res.reserved[1] = 0;

    const int N = m->n, H = m->H, minInl = m->min_inl;
    uint32_t* final_mask = a.masks + m->word_off;
    const int32_t* counts = a.counts + m->hyp_off;
    const Cam cam = {m->fx, m->fy, m->cx, m->cy};
    bool finished = false;

    if (N < minInl) {                        // PnPsolver.cpp:110-114
        res.no_more = 1;
        for (int w = tid; w < words; w += blockDim.x) final_mask[w] = 0u;
        finished = true;
    }

    int best = 0, bestH = -1, lastRefBestH = -2, lastCntR = 0, mSel = 0;
    int cursor = 0;
    while (!finished) {
        // next hypothesis with cnt >= minInliers (PnPsolver.cpp:146)
        if (tid == 0) s_found = H;
        __syncthreads();
        int h = H;
        for (int base = cursor; base < H; base += blockDim.x) {   // uniform trip count: h is read after a barrier
            const int hc = base + tid;
            if (hc < H && counts[hc] >= minInl) atomicMin(&s_found, hc);
            __syncthreads();
            h = s_found;
            __syncthreads();
            if (h < H) break;
        }
        if (h >= H) break;

        if (counts[h] > best) {              // :149 strict: first maximum wins
            best = counts[h];
            bestH = h;
            if (tid < 12) s_bestpose[tid] = a.poses[(size_t)(m->hyp_off + h) * 12 + tid];
            __syncthreads();
            cta_score_exact_pnp(m, a, s_bestpose, bestmask, &s_cnt);
            // ordered compaction of the best set -> sel (Refine, :195-204)
            if (tid == 0) {
                int acc = 0;
                for (int w = 0; w < words; ++w) { prefix[w] = acc; acc += __popc(bestmask[w]); }
                prefix[words] = acc;
            }
            __syncthreads();
            mSel = prefix[words];
            uint32_t* sel = a.sel + m->corr_off;
            for (int w = tid; w < words; w += blockDim.x) {
                uint32_t bits = bestmask[w];
                int o = prefix[w];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    sel[o++] = (uint32_t)(w * 32 + b);
                }
            }
            __syncthreads();
        }
        res.n_refines++;

        if (bestH != lastRefBestH) {
            // ---------------- Refine(): n-point EPnP on the best set (:206-217) ----------------
            const int n = mSel;
            const uint32_t* sel = a.sel + m->corr_off;
            double* pw = a.pw_s + (size_t)m->corr_off * 3;
            double* us = a.us_s + (size_t)m->corr_off * 2;
            double* al = a.al_s + (size_t)m->corr_off * 4;
            for (int i = tid; i < n; i += blockDim.x) {       // add_correspondence
                const size_t g = (size_t)m->corr_off + sel[i];
                const float4 c = a.cA[g];
                const float2 q = a.uv[g];
                pw[3 * i] = (double)c.x; pw[3 * i + 1] = (double)c.y; pw[3 * i + 2] = (double)c.z;
                us[2 * i] = (double)q.x; us[2 * i + 1] = (double)q.y;
            }
            __syncthreads();
            if (tid < 3) {                                     // centroid (:301-303)
                double s = 0.0;
#pragma unroll 8
                for (int i = 0; i < n; ++i) s += pw[3 * i + tid];
                s_C0[tid] = s / (double)n;
            }
            __syncthreads();
            if (tid < 6) {                                     // PW0^T PW0 upper triangle (:306-310)
                const int r = (tid < 3) ? 0 : (tid < 5 ? 1 : 2);
                const int c = (tid < 3) ? tid : (tid < 5 ? tid - 2 : 2);
                double s = 0.0;
#pragma unroll 8
                for (int i = 0; i < n; ++i) s += (pw[3 * i + r] - s_C0[r]) * (pw[3 * i + c] - s_C0[c]);
                s_A[r * 3 + c] = s;
            }
            __syncthreads();
            if (tid == 0) {
                double A[9];
                for (int i = 0; i < 9; ++i) A[i] = s_A[i];
                epnp_control_points(s_C0, A, n, s_cws);
                epnp_cc_inverse(s_cws, s_CCi);
            }
            __syncthreads();
            for (int i = tid; i < n; i += blockDim.x) epnp_alphas(pw + 3 * i, s_cws, s_CCi, al + 4 * i);
            __syncthreads();
            if (tid < 78) {                                    // MtM upper triangle, one entry per thread (:379)
                int ea = 0, rem = tid;
                while (rem >= 12 - ea) { rem -= 12 - ea; ++ea; }
                const int eb = ea + rem;
                double s = 0.0;
                for (int i = 0; i < n; ++i) {
                    double a0, a1, b0, b1;
                    epnp_m_entry(al + 4 * i, us[2 * i], us[2 * i + 1], cam, ea, a0, a1);
                    epnp_m_entry(al + 4 * i, us[2 * i], us[2 * i + 1], cam, eb, b0, b1);
                    s += a0 * b0;
                    s += a1 * b1;
                }
                s_MtM[tri_idx(12, ea, eb)] = s;
            }
            __syncthreads();
            if (tid == 0) {
                double MtM[78];
#pragma unroll
                for (int i = 0; i < 78; ++i) MtM[i] = s_MtM[i];
                epnp_solve_betas(MtM, s_cws, s_U4, s_betas, a.rec + (size_t)blockIdx.x * (kMaxSweepsRec * 66));
                for (int k = 0; k < 3; ++k) epnp_ccs(s_betas + 4 * k, s_U4, s_ccs[k]);
            }
            __syncthreads();
            if (tid < 3) {                                     // solve_for_sign on pcs(0,2) (:495-502)
                double pc[3];
                epnp_pc(al, s_ccs[tid], pc);
                s_sign[tid] = (pc[2] < 0.0) ? -1.0 : 1.0;
            }
            __syncthreads();
            if (tid < 9) {                                     // pc0 of the three candidates (:435,438)
                const int k = tid / 3, c = tid % 3;
                const bool neg = s_sign[k] < 0.0;
                double s = 0.0;
                for (int i = 0; i < n; ++i) {
                    double pc[3];
                    epnp_pc(al + 4 * i, s_ccs[k], pc);
                    s += neg ? -pc[c] : pc[c];
                }
                s_pc0[k][c] = s / (double)n;
            } else if (tid < 12) {                             // pw0 (:436,439)
                const int c = tid - 9;
                double s = 0.0;
#pragma unroll 8
                for (int i = 0; i < n; ++i) s += pw[3 * i + c];
                s_pw0[c] = s / (double)n;
            }
            __syncthreads();
            if (tid < 27) {                                    // M = sum (pc-pc0)^T (pw-pw0) (:443-447)
                const int k = tid / 9, r = (tid % 9) / 3, c = tid % 3;
                const bool neg = s_sign[k] < 0.0;
                double s = 0.0;
                for (int i = 0; i < n; ++i) {
                    double pc[3];
                    epnp_pc(al + 4 * i, s_ccs[k], pc);
                    const double pcr = neg ? -pc[r] : pc[r];
                    s += (pcr - s_pc0[k][r]) * (pw[3 * i + c] - s_pw0[c]);
                }
                s_M[k][r * 3 + c] = s;
            }
            __syncthreads();
            if (tid < 3) {
                epnp_horn(s_M[tid], s_pc0[tid], s_pw0, s_R[tid], s_t[tid]);
                double sum2 = 0.0;                             // reprojection_error (:417-431)
                for (int i = 0; i < n; ++i) sum2 += epnp_reproj_term(s_R[tid], s_t[tid], pw + 3 * i, us[2 * i], us[2 * i + 1], cam);
                s_rep[tid] = sum2 / (double)n;
            }
            __syncthreads();
            if (tid == 0) {
                int Nn = 0;                                    // :407-409
                if (s_rep[1] < s_rep[0]) Nn = 1;
                if (s_rep[2] < s_rep[Nn]) Nn = 2;
                for (int i = 0; i < 9; ++i) s_pose[i] = (float)s_R[Nn][i];
                for (int i = 0; i < 3; ++i) s_pose[9 + i] = (float)s_t[Nn][i];
            }
            __syncthreads();
            cta_score_exact_pnp(m, a, s_pose, refmask, &s_cnt);   // :220
            lastCntR = s_cnt;
            lastRefBestH = bestH;
        }

        if (lastCntR > minInl) {                               // :225 strict
            res.ok = 1; res.refined = 1; res.n_inliers = lastCntR; res.n_hyp = h + 1;
            for (int i = 0; i < 9; ++i) res.R[i] = s_pose[i];
            for (int i = 0; i < 3; ++i) res.t[i] = s_pose[9 + i];
            for (int w = tid; w < words; w += blockDim.x) final_mask[w] = refmask[w];
            finished = true;
            break;
        }
        cursor = h + 1;
    }

    if (!finished) {                                           // :173-188 budget exhausted
        res.no_more = 1;
        res.n_hyp = H;
        if (best >= minInl) {
            res.ok = 1;
            res.n_inliers = best;
            for (int i = 0; i < 9; ++i) res.R[i] = s_bestpose[i];
            for (int i = 0; i < 3; ++i) res.t[i] = s_bestpose[9 + i];
            for (int w = tid; w < words; w += blockDim.x) final_mask[w] = bestmask[w];
        } else {
            for (int w = tid; w < words; w += blockDim.x) final_mask[w] = 0u;
        }
    }
    res.best_hyp = bestH;
    res.best_count = best;
    if (tid == 0) {
        reinterpret_cast<ResultRec*>(a.results)[blockIdx.x] = res;
        if (a.results2) reinterpret_cast<ResultRec*>(a.results2)[blockIdx.x] = res;
    }
}

}  // namespace rsac
