// select.cuh -- sequential-semantics replay + Refine for PnPsolver and MLPnPsolver.
//
// Reference control flow: PnPsolver::iterate / Refine (src/PnPsolver.cpp:102-238) and
// MLPnPsolver::iterate / Refine (src/MLPnPsolver.cpp:56-183, 257-318) -- identical logic
// (SURVEY Appendix C): scan the hypotheses in draw order; every hypothesis with
// cnt >= minInliers (i) replaces the best set on a strict '>' and (ii) triggers Refine() =
// an n-point solve on the CURRENT BEST set followed by CheckInliers; the first refine with
// cnt > minInliers wins; if the budget is exhausted the unrefined best is returned.
//
// One CTA per problem.  The n-point solves evaluate every sum "entry-parallel": one thread per
// output entry walks the selected points in index order, which reproduces the serial summation
// order of the CPU checker exactly; the dense tails run on thread 0 with the same device
// functions as the minimal solvers.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "epnp.cuh"
#include "mlpnp.cuh"
#include "score.cuh"

namespace rsac {

struct SelectArgs {
    const ProblemMeta* metas;
    const float4* cA;
    const float4* cB;
    const float4* cC;
    const void* poses;       // [sumH][12] float (PnP) / double (MLPnP)
    const int32_t* counts;   // [sumH]
    const double* cov;       // MLPnP: optional [total][9]
    // scratch, indexed like the correspondences
    uint32_t* sel;           // compacted indices of the best set
    double* pw_s;            // EPnP [total][3]          MLPnP: [total][33] f3 p3 N6 P4 q3 J12 r2
    double* us_s;            // EPnP [total][2]
    double* al_s;            // EPnP [total][4]
    double* tm_s;            // EPnP [total][12]: per-point pcs of the three candidates (9) + reprojection terms (3)
    // outputs
    void* results;           // rsac_result[C] (layout in ransac_b200.h)
    void* results2;          // optional second copy (collective send buffer)
    uint32_t* masks;         // final masks, word_off per problem
    int32_t problem_base;    // global index of problem 0 (sharding)
    int32_t flags;
    const int32_t* resume;   // optional [C]: hypotheses already consumed by earlier iterate() calls
    // early exit in stages (pnp_pipeline.cuh): ee = [upto: C][listX: C][listY: C][listC: C][counters: 16] or nullptr
    int32_t* ee = nullptr;
    int32_t C = 0;
    int32_t first_phase = 0;   // HA: hypotheses every problem has after phase A
    int32_t only_phase = -1;   // >= 0: the phase-C replay: only problems marked undecided run; they resume where they stopped
    const int32_t* problem_ids = nullptr;   // optional [C]: global problem index of every problem (rsac_set_problem_ids)
    // The replay split around the first Refine's 12x12 eigen-solve (PnP, main replay of a sweep): split_phase 1 runs up to
    // M^T M and parks the problem in `carry`; select_eigen_kernel (one warp per problem) solves; split_phase 2 picks up
    // behind the eigen-solve and finishes the scan (any further Refine of the same problem runs its eigen-solve in place).
    // 0 = the whole replay in one launch.  The eigen-solve is 36 % of the replay's critical path and uses one warp of the
    // CTA's three: on its own it holds a sixth of the registers for that time, which is what other sweeps in flight get.
    struct SelectCarry* carry = nullptr;
    int32_t split_phase = 0;
};

struct SelectCarry {
    int32_t waiting;          // 1: parked in front of the eigen-solve
    int32_t best, bestH, lastRefH, lastCntR, mSel, cursor, h, n_refines, pad[3];
    float bestpose[12];
    double C0[3], A[9], cws[12], CCi[9], MtM[78], U4[48], w4[4], pad2;
};

struct ResultRec {   // mirrors rsac_result (include/ransac_b200.h)
    int32_t ok, no_more, n_inliers, best_hyp, refined, n_refines, best_count, n_hyp;
    float R[9], t[3], s;
    int32_t problem, reserved[2];
};
static_assert(sizeof(ResultRec) == 96, "rsac_result layout");

#ifndef RSAC_SELECT_THREADS
#define RSAC_SELECT_THREADS 96
#endif
#ifndef RSAC_SELECT_CTAS
#define RSAC_SELECT_CTAS 7
#endif
// 7 CTAs/SM: 1036 resident candidates, so a 1024-candidate sweep is one wave.  Three warps are what the phases can
// use (78 MtM entries, three beta branches, one eigen-solve); with several sweeps in flight per GPU (bench.py) what
// counts is register-file share x time, and 96 x 7 (80 registers) beats 128 x 7 (72 registers): 0.538 against 0.558 ms
// per sweep, the kernel alone 0.193 against 0.205 ms; 64 x 9 (96 registers): 0.571, 96 x 9: 0.552.  (With ONE
// exhaustive sweep at a time 128 x 7 had won because the 96-thread shapes slowed the following solve kernel down.)
constexpr int kSelectThreads = RSAC_SELECT_THREADS;
constexpr int kSelectThreadsMlpnp = 128;   // MLPnP's refine (cfg2: 64 frames, one partial wave) is faster with four warps: 0.46 against 0.54 ms
static_assert(kSelectThreadsMlpnp >= 78, "refine_mlpnp maps the 78 A^T P A entries to one thread each");
template <int MODEL> constexpr int select_threads() { return MODEL == 0 ? kSelectThreads : kSelectThreadsMlpnp; }
constexpr int kSelectCtasPerSm = RSAC_SELECT_CTAS;

// diagnostic: clock64() at phase boundaries of block 0 (rsac_debug_select_clocks)
static __device__ long long g_select_clocks[16];
#define RSAC_SEL_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_select_clocks[i] = clock64(); } while (0)
constexpr int kMlpnpScratch = 33;   // doubles per selected observation

// exact CheckInliers of one pose over all correspondences of the problem by the whole CTA
template <int MODEL>
__device__ inline void cta_score_exact(const ProblemMeta* m, const SelectArgs& a,
                                       const typename ScoreModel<MODEL>::pose_t* pose, uint32_t* mask_out, int* s_cnt)
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
            const float4 q = a.cC[g];
            in = ScoreModel<MODEL>::exact(pose, c.x, c.y, c.z, q.x, q.y, a.cB[g].y, m);
        }
        const uint32_t word = __ballot_sync(0xffffffffu, in);
        if (lane == 0) { mask_out[base >> 5] = word; local += __popc(word); }
    }
    if (lane == 0 && local) atomicAdd(s_cnt, local);
    __syncthreads();
}

// ------------------------------------------------------------------ EPnP refine
struct EpnpShared {
    double C0[3], A[9], cws[12], CCi[9], MtM[78], U4[48], betas[12], w4[4];
    double ccs[3][12], sign[3], pc0[3][3], pw0[3], M[3][9], R[3][9], t[3][3], rep[3];
};

// PnPsolver::gauss_newton (:675-691) by one warp: the same arithmetic as epnp_gauss_newton / epnp_gn_system /
// epnp_qr_solve element for element, spread over lanes.
//   * rows of the 6x4 system: lane i < 6 builds row i and b[i] (epnp_gn_system's row body);
//   * Householder QR column-wise: lane j < 4 owns column j, lane 4 owns the right-hand side.  At step k lane k
//     scales its column and forms the reflector (eta, sigma, A1, A2), broadcasts it, and the lanes j > k apply it
//     to their columns in parallel.  The reference applies the reflectors to b after the factorisation
//     (:763-773), one after another, using entries of column j that no later step touches -- applying reflector k
//     to b at step k is the same sequence of operations on b;
//   * back substitution is redundant on every lane so that the betas stay warp-uniform.
// scratch: 40 doubles of shared memory private to the warp.
__device__ inline void epnp_gauss_newton_warp(const double* L, const double* rho, double* betas, double* scratch, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    double X[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        if (lane < 6) {
            const int i = lane;
            double l[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) l[j] = L[i * 10 + j];
            const double Lt[4][4] = {{2 * l[0], l[1], l[3], l[6]},
                                     {l[1], 2 * l[2], l[4], l[7]},
                                     {l[3], l[4], 2 * l[5], l[8]},
                                     {l[6], l[7], l[8], 2 * l[9]}};
#pragma unroll
            for (int r = 0; r < 4; ++r)
                scratch[r * 6 + i] = rfma(Lt[r][3], betas[3], rfma(Lt[r][2], betas[2], rfma(Lt[r][1], betas[1], Lt[r][0] * betas[0])));
            double q = (l[0] * betas[0]) * betas[0];
            q = rfma(l[1] * betas[0], betas[1], q);
            q = rfma(l[2] * betas[1], betas[1], q);
            q = rfma(l[3] * betas[0], betas[2], q);
            q = rfma(l[4] * betas[1], betas[2], q);
            q = rfma(l[5] * betas[2], betas[2], q);
            q = rfma(l[6] * betas[0], betas[3], q);
            q = rfma(l[7] * betas[1], betas[3], q);
            q = rfma(l[8] * betas[2], betas[3], q);
            q = rfma(l[9] * betas[3], betas[3], q);
            scratch[4 * 6 + i] = rho[i] - q;
        }
        __syncwarp();
        double col[6];                    // lane j < 4: column j of A; lane 4: b
#pragma unroll
        for (int i = 0; i < 6; ++i) col[i] = scratch[(lane < 5 ? lane : 0) * 6 + i];
        double A2 = 0.0;
        bool zero_col = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double A1 = 0.0;
            bool z = false;
            if (lane == k) {
                double eta = fabs(col[k]);
#pragma unroll
                for (int i = k + 1; i < 6; ++i) {
                    const double elt = fabs(col[i]);
                    if (eta < elt) eta = elt;
                }
                if (eta == 0) {
                    z = true;
                } else {
                    const double inv_eta = 1. / eta;
                    double sum = 0.0;
#pragma unroll
                    for (int i = k; i < 6; ++i) {
                        col[i] *= inv_eta;
                        sum = rfma(col[i], col[i], sum);
                    }
                    double sigma = sqrt(sum);
                    if (col[k] < 0) sigma = -sigma;
                    col[k] += sigma;
                    A1 = sigma * col[k];
                    A2 = -eta * sigma;
                }
            }
            z = __shfl_sync(FULL, z, k);
            if (z) { zero_col = true; break; }             // :722-727: returns with X untouched
            A1 = __shfl_sync(FULL, A1, k);
            double v[6];
#pragma unroll
            for (int i = k; i < 6; ++i) v[i] = __shfl_sync(FULL, col[i], k);
            if (lane > k && lane < 5) {
                double sdot = 0;
#pragma unroll
                for (int i = k; i < 6; ++i) sdot = rfma(v[i], col[i], sdot);
                const double tau = sdot / A1;
#pragma unroll
                for (int i = k; i < 6; ++i) col[i] = rfma(-tau, v[i], col[i]);
            }
        }
        if (!zero_col) {
            // R (upper triangle, held column-wise), its diagonal A2 and the transformed b to every lane
            double Rm[4][4], bb[4], d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                d[j] = __shfl_sync(FULL, A2, j);
#pragma unroll
                for (int i = 0; i < 4; ++i) Rm[i][j] = __shfl_sync(FULL, col[i], j);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) bb[i] = __shfl_sync(FULL, col[i], 4);
            X[3] = bb[3] / d[3];
#pragma unroll
            for (int i = 2; i >= 0; --i) {
                double sum = 0;
#pragma unroll
                for (int j = i + 1; j < 4; ++j) sum = rfma(Rm[i][j], X[j], sum);
                X[i] = (bb[i] - sum) / d[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) betas[i] += X[i];
        __syncwarp();
    }
}

// PnPsolver::Refine's compute_pose on the n selected points (PnPsolver.cpp:206-217, 359-415);
// result as float R|t in pose_out[12] (shared)
constexpr int kSelTile = 64, kSelTP = kSelTile + 1;          // padded rows: entries of different rows fall into different banks
constexpr int kSelRowTile = 32, kSelRP = kSelRowTile + 1;    // MtM: 24 row entries per point
constexpr int kSelTileDoubles = 24 * kSelRP > 12 * kSelTP ? 24 * kSelRP : 12 * kSelTP;

// front half: add_correspondence .. M^T M (PnPsolver.cpp:206-217, 296-343, 364-379)
__device__ inline void refine_epnp_front(const ProblemMeta* m, const SelectArgs& a, int n, EpnpShared& S, double* s_tile)
{
    const int tid = threadIdx.x;
    const Cam cam = {m->fx, m->fy, m->cx, m->cy};
    const uint32_t* sel = a.sel + m->corr_off;
    double* pw = a.pw_s + (size_t)m->corr_off * 3;
    double* us = a.us_s + (size_t)m->corr_off * 2;
    double* al = a.al_s + (size_t)m->corr_off * 4;
    constexpr int kTile = kSelTile, kTP = kSelTP, kRowTile = kSelRowTile, kRP = kSelRP;
    RSAC_SEL_MARK(2);
    // Every n-point sum below is formed by ONE thread per output entry that adds the per-point terms in index
    // order (the checker's serial order).  The terms travel through a shared-memory tile: the whole CTA loads /
    // computes the terms of kTile points (coalesced, independent), then the few summing threads add them from
    // shared memory -- the dependent chain is the additions alone, not a global-memory load per point.
    for (int i = tid; i < n; i += blockDim.x) {       // add_correspondence
        const size_t g = (size_t)m->corr_off + sel[i];
        const float4 c = a.cA[g];
        const float4 q = a.cC[g];
        pw[3 * i] = (double)c.x; pw[3 * i + 1] = (double)c.y; pw[3 * i + 2] = (double)c.z;
        us[2 * i] = (double)q.x; us[2 * i + 1] = (double)q.y;
    }
    __syncthreads();
    {                                                  // centroid (:301-303)
        double s = 0.0;
        for (int base = 0; base < n; base += kTile) {
            const int cnt = min(kTile, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 3; e += blockDim.x) s_tile[(e % 3) * kTP + e / 3] = pw[3 * base + e];
            __syncthreads();
            if (tid < 3)
                for (int i = 0; i < cnt; ++i) s += s_tile[tid * kTP + i];
        }
        if (tid < 3) S.C0[tid] = s / (double)n;
    }
    __syncthreads();
    {                                                  // PW0^T PW0 upper triangle (:306-310)
        const int r = (tid < 3) ? 0 : (tid < 5 ? 1 : 2);
        const int c = (tid < 3) ? tid : (tid < 5 ? tid - 2 : 2);
        const double c0r = S.C0[r], c0c = S.C0[c < 3 ? c : 0];
        double s = 0.0;
        for (int base = 0; base < n; base += kTile) {
            const int cnt = min(kTile, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 3; e += blockDim.x) s_tile[(e % 3) * kTP + e / 3] = pw[3 * base + e];
            __syncthreads();
            if (tid < 6)
                for (int i = 0; i < cnt; ++i) s = rfma(s_tile[r * kTP + i] - c0r, s_tile[c * kTP + i] - c0c, s);   // fma as in epnp_compute_pose_small
        }
        if (tid < 6) S.A[r * 3 + c] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double A[9];
        for (int i = 0; i < 9; ++i) A[i] = S.A[i];
        epnp_control_points(S.C0, A, n, S.cws);
        epnp_cc_inverse(S.cws, S.CCi);
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) epnp_alphas(pw + 3 * i, S.cws, S.CCi, al + 4 * i);
    __syncthreads();
    RSAC_SEL_MARK(3);
    {                                                  // MtM upper triangle, one entry per thread (:379)
        // the two rows of M of every point of a tile are materialised in shared memory (epnp_m_rows, zeros
        // included); every entry adds its two products per point in index order
        // entries e = tid and tid + blockDim.x (78 entries over >= 64 threads)
        int ea[2] = {0, 0}, eb[2] = {0, 0};
        bool have[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int e = tid + k * blockDim.x;
            have[k] = e < 78;
            if (have[k]) {
                int rem = e;
                while (rem >= 12 - ea[k]) { rem -= 12 - ea[k]; ++ea[k]; }
                eb[k] = ea[k] + rem;
            }
        }
        double acc[2] = {0.0, 0.0};
        for (int base = 0; base < n; base += kRowTile) {
            const int cnt = min(kRowTile, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 4; e += blockDim.x) {      // (point, control point j): six row entries
                const int pt = e >> 2, j = e & 3, i = base + pt;
                const double aj = al[4 * i + j], u = us[2 * i], v = us[2 * i + 1];
                s_tile[(3 * j + 0) * kRP + pt] = aj * cam.fx;
                s_tile[(3 * j + 1) * kRP + pt] = 0.0;
                s_tile[(3 * j + 2) * kRP + pt] = aj * (cam.cx - u);
                s_tile[(12 + 3 * j + 0) * kRP + pt] = 0.0;
                s_tile[(12 + 3 * j + 1) * kRP + pt] = aj * cam.fy;
                s_tile[(12 + 3 * j + 2) * kRP + pt] = aj * (cam.cy - v);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (!have[k]) continue;
                const double* r0a = s_tile + ea[k] * kRP;
                const double* r0b = s_tile + eb[k] * kRP;
                const double* r1a = r0a + 12 * kRP;
                const double* r1b = r0b + 12 * kRP;
#pragma unroll 4
                for (int i = 0; i < cnt; ++i) {
                    acc[k] = rfma(r0a[i], r0b[i], acc[k]);
                    acc[k] = rfma(r1a[i], r1b[i], acc[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (have[k]) S.MtM[tri_idx(12, ea[k], eb[k])] = acc[k];
    }
    __syncthreads();
    RSAC_SEL_MARK(4);
}

// a 4-point best set in the default mode takes the QR null space: no eigen-solve to split off
__device__ __forceinline__ bool refine_epnp_has_eigen(const SelectArgs& a, int n) { return !(n == 4 && !(a.flags & 8)); }

// back half: null space (do_eigen: the 12x12 eigen-solve runs here; false: S.U4 / S.w4 already hold its result), betas,
// R and t of the three candidates, the best one as float R|t in pose_out[12] (shared) (PnPsolver.cpp:380-415)
__device__ inline void refine_epnp_back(const ProblemMeta* m, const SelectArgs& a, int n, EpnpShared& S, double2* s_rec, float* pose_out,
                                        double* s_tile, bool do_eigen)
{
    // do_eigen == false: select_eigen_kernel has left S.U4 and S.w4 (the split replay)
    const int tid = threadIdx.x;
    const Cam cam = {m->fx, m->fy, m->cx, m->cy};
    double* pw = a.pw_s + (size_t)m->corr_off * 3;
    double* us = a.us_s + (size_t)m->corr_off * 2;
    double* al = a.al_s + (size_t)m->corr_off * 4;
    constexpr int kTile = kSelTile, kTP = kSelTP;
    if (!refine_epnp_has_eigen(a, n)) {                // RSAC_FLAG_EPNP_EIGEN clear: QR null space for a 4-point system
        if (tid == 0) epnp_solve_betas_qr4(al, us, cam, S.cws, S.U4, S.betas);
    } else {
        // 12x12 eigen-solve (:380): warp 0, cooperative schedule of the same rotations
        if (do_eigen && tid < 32) jacobi_lowest_warp<12, 4>(S.MtM, S.w4, S.U4, s_rec, tid, blockIdx.x == 0 ? g_select_clocks + 11 : nullptr);
        __syncthreads();
        RSAC_SEL_MARK(5);
        // L (6x10, :604-637) and rho (:639-647) entry-parallel into shared memory: 72 control-point differences,
        // then 60 dot products (same expressions as epnp_L_6x10)
        __shared__ double s_dv[72], s_L[60], s_rho[6], s_gn[3][40];
        for (int e = tid; e < 72; e += blockDim.x) {
            const int i = e / 18, j = (e % 18) / 3, c = e % 3;
            const int pa = (j < 3) ? 0 : (j < 5 ? 1 : 2);
            const int pb = (j < 3) ? j + 1 : (j < 5 ? j - 1 : 3);
            s_dv[e] = S.U4[(3 * pa + c) * 4 + i] - S.U4[(3 * pb + c) * 4 + i];
        }
        if (tid == blockDim.x - 1) epnp_rho(S.cws, s_rho);
        __syncthreads();
        for (int e = tid; e < 60; e += blockDim.x) {
            const int row = e / 10, col = e % 10;
            const int xs[10] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3}, ys[10] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3};
            const double* x = s_dv + (xs[col] * 6 + row) * 3;
            const double* y = s_dv + (ys[col] * 6 + row) * 3;
            const double d = x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
            s_L[e] = (xs[col] == ys[col]) ? d : 2.0 * d;
        }
        __syncthreads();
        // the three beta approximations + Gauss-Newton (:395-405) are independent: one warp each (two warps when
        // the CTA has only two: approx_3 alone, approx_1 + approx_2 one after the other); the initial guess on
        // lane 0, the five Gauss-Newton steps on the warp (epnp_gauss_newton_warp)
        const bool three = blockDim.x >= 96;
        if (tid < (three ? 96 : 64)) {
            const int wk = tid >> 5, lane = tid & 31;
            for (int pass = 0; pass < 2; ++pass) {
                int which = -1;                          // 0: approx_1, 1: approx_2, 2: approx_3
                if (pass == 0) which = (wk == 0) ? 2 : (wk == 1 ? 0 : 1);
                else if (!three && wk == 1) which = 1;
                if (which < 0) break;
                double bt[4] = {0.0, 0.0, 0.0, 0.0};
                if (lane == 0) {
                    if (which == 0) epnp_betas_approx_1((const double*)s_L, s_rho, bt);
                    else if (which == 1) epnp_betas_approx_2((const double*)s_L, s_rho, bt);
                    else epnp_betas_approx_3((const double*)s_L, s_rho, bt);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) bt[i] = __shfl_sync(0xffffffffu, bt[i], 0);
                epnp_gauss_newton_warp(s_L, s_rho, bt, s_gn[wk], lane);
                if (lane == 0)
                    for (int i = 0; i < 4; ++i) S.betas[4 * which + i] = bt[i];
            }
        }
    }
    __syncthreads();
    RSAC_SEL_MARK(6);
    if (tid < 3) epnp_ccs(S.betas + 4 * tid, S.U4, S.ccs[tid]);
    __syncthreads();
    if (tid < 3) {                                     // solve_for_sign on pcs(0,2) (:495-502)
        double pc[3];
        epnp_pc(al, S.ccs[tid], pc);
        S.sign[tid] = (pc[2] < 0.0) ? -1.0 : 1.0;
    }
    __syncthreads();
    double* tm = a.tm_s + (size_t)m->corr_off * 12;
    {
        // pcs of the three candidates, sign applied (:354-357, :495-502), computed tile by tile into shared
        // memory (and kept in `tm` for the M sums); pc0 of the three candidates (:435,438) and pw0 (:436,439)
        double s = 0.0;
        for (int base = 0; base < n; base += kTile) {
            const int cnt = min(kTile, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 3; e += blockDim.x) {
                const int pt = e / 3, k = e - 3 * pt, i = base + pt;
                double pc[3];
                epnp_pc(al + 4 * i, S.ccs[k], pc);
                const bool neg = S.sign[k] < 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double v = neg ? -pc[c] : pc[c];
                    tm[(size_t)i * 12 + 3 * k + c] = v;
                    s_tile[(3 * k + c) * kTP + pt] = v;
                }
                s_tile[(9 + k) * kTP + pt] = pw[3 * i + k];
            }
            __syncthreads();
            if (tid < 12)
                for (int i = 0; i < cnt; ++i) s += s_tile[tid * kTP + i];
        }
        if (tid < 9) S.pc0[tid / 3][tid % 3] = s / (double)n;
        else if (tid < 12) S.pw0[tid - 9] = s / (double)n;
    }
    __syncthreads();
    {                                                  // M = sum (pc-pc0)^T (pw-pw0) (:443-447)
        const int k = (tid < 27) ? tid / 9 : 0, r = (tid % 9) / 3, c = tid % 3;
        const double p0 = S.pc0[k][r], w0 = S.pw0[c];
        double s = 0.0;
        for (int base = 0; base < n; base += kTile) {
            const int cnt = min(kTile, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 12; e += blockDim.x) {
                const int pt = e / 12, j = e - 12 * pt, i = base + pt;
                s_tile[j * kTP + pt] = (j < 9) ? tm[(size_t)i * 12 + j] : pw[3 * i + (j - 9)];
            }
            __syncthreads();
            if (tid < 27)
                for (int i = 0; i < cnt; ++i) s += (s_tile[(3 * k + r) * kTP + i] - p0) * (s_tile[(9 + c) * kTP + i] - w0);
        }
        if (tid < 27) S.M[k][r * 3 + c] = s;
    }
    __syncthreads();
    RSAC_SEL_MARK(7);
    if (tid < 3) epnp_horn(S.M[tid], S.pc0[tid], S.pw0, S.R[tid], S.t[tid]);
    __syncthreads();
    {                                                  // reprojection_error terms (:417-431), summed in index order
        double sum2 = 0.0;
        for (int base = 0; base < n; base += kTile) {
            const int cnt = min(kTile, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 3; e += blockDim.x) {
                const int pt = e / 3, k = e - 3 * pt, i = base + pt;
                s_tile[k * kTP + pt] = epnp_reproj_term(S.R[k], S.t[k], pw + 3 * i, us[2 * i], us[2 * i + 1], cam);
            }
            __syncthreads();
            if (tid < 3)
                for (int i = 0; i < cnt; ++i) sum2 += s_tile[tid * kTP + i];
        }
        if (tid < 3) S.rep[tid] = sum2 / (double)n;
    }
    __syncthreads();
    if (tid == 0) {
        int Nn = 0;                                    // :407-409
        if (S.rep[1] < S.rep[0]) Nn = 1;
        if (S.rep[2] < S.rep[Nn]) Nn = 2;
        for (int i = 0; i < 9; ++i) pose_out[i] = (float)S.R[Nn][i];
        for (int i = 0; i < 3; ++i) pose_out[9 + i] = (float)S.t[Nn][i];
    }
    __syncthreads();
}

// ----------------------------------------------------------------- MLPnP refine
struct MlpnpShared {
    double planarTest[9], eigenRot[9], AtPA[78], x[6], A[36], g[6], dx[6], ev[1], result1[12];
    unsigned long long maxdl_bits;
    int planar, dec;
};

// one design-matrix column entry of observation (N, pt): rows a0[col], a1[col] (MLPnPsolver.cpp:404-476)
__device__ inline void mlpnp_row_entry(const double* N, const double* pt, bool planar, int col, double& e0, double& e1)
{
    if (planar) {
        if (col < 6) { const int r = col >> 1, c = 1 + (col & 1); e0 = N[r * 2 + 0] * pt[c]; e1 = N[r * 2 + 1] * pt[c]; }
        else { const int r = col - 6; e0 = N[r * 2 + 0]; e1 = N[r * 2 + 1]; }
    } else {
        if (col < 9) { const int r = col / 3, c = col - 3 * r; e0 = N[r * 2 + 0] * pt[c]; e1 = N[r * 2 + 1] * pt[c]; }
        else { const int r = col - 9; e0 = N[r * 2 + 0]; e1 = N[r * 2 + 1]; }
    }
}

// MLPnPsolver::Refine's computePose on the n selected observations (MLPnPsolver.cpp:269-290,
// 321-623); result as double R|t in pose_out[12] (shared)
__device__ inline void refine_mlpnp(const ProblemMeta* m, const SelectArgs& a, int n, MlpnpShared& S, double2* s_rec, double* pose_out)
{
    const int tid = threadIdx.x;
    const uint32_t* sel = a.sel + m->corr_off;
    double* sc = a.pw_s + (size_t)m->corr_off * kMlpnpScratch;   // per observation: f3 p3 N6 P4 q3 J12 r2
    const bool use_cov = a.cov != nullptr;
    for (int i = tid; i < n; i += blockDim.x) {
        const size_t g = (size_t)m->corr_off + sel[i];
        const float4 c = a.cA[g];
        const float4 q = a.cC[g];
        double* o = sc + (size_t)i * kMlpnpScratch;
        mlpnp_bearing(q.x, q.y, m->k1, o);                                  // f
        o[3] = (double)c.x; o[4] = (double)c.y; o[5] = (double)c.z;          // p
        mlpnp_nullspace(o, o + 6);                                           // N
        if (use_cov) mlpnp_weight(o + 6, a.cov + 9 * g, o + 12);             // P
    }
    __syncthreads();
    // n-point sums: one thread per output entry adds the per-observation terms in index order (the checker's serial
    // order); the terms travel through a shared-memory tile staged by the whole CTA (see refine_epnp)
    constexpr int kTa = 24, kTPa = kTa + 1;            // A^T P A: two design-matrix rows (12 + 12) and P (4) per observation
    constexpr int kTg = 38, kTPg = kTg + 1;            // Gauss-Newton: J (12), P (4), r (2) per observation
    __shared__ double m_tile[28 * kTPa > 18 * kTPg ? 28 * kTPa : 18 * kTPg];
    {                                                  // planarTest = sum p p^T (:346)
        const int r = (tid % 9) / 3, c = tid % 3;
        double s = 0.0;
        for (int base = 0; base < n; base += kTg) {
            const int cnt = min(kTg, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * 3; e += blockDim.x) m_tile[(e % 3) * kTPg + e / 3] = sc[(size_t)(base + e / 3) * kMlpnpScratch + 3 + e % 3];
            __syncthreads();
            if (tid < 9)
                for (int i = 0; i < cnt; ++i) s = rfma(m_tile[r * kTPg + i], m_tile[c * kTPg + i], s);
        }
        if (tid < 9) S.planarTest[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < 9; ++i) S.eigenRot[i] = (i % 4 == 0) ? 1.0 : 0.0;
        S.planar = 0;
        if (rank3_fullpiv(S.planarTest) == 2) {        // :354
            S.planar = 1;
            double A[9], w[3], V[9];
            for (int i = 0; i < 9; ++i) A[i] = S.planarTest[i];
            jacobi_eig<double, 3>(A, w, V);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) S.eigenRot[r * 3 + c] = V[c * 3 + r];
        }
    }
    __syncthreads();
    const bool planar = S.planar != 0;
    for (int i = tid; i < n; i += blockDim.x) {        // points3 (:362-363)
        double* o = sc + (size_t)i * kMlpnpScratch;
        if (planar) {
            for (int r = 0; r < 3; ++r)
                o[16 + r] = S.eigenRot[r * 3 + 0] * o[3] + S.eigenRot[r * 3 + 1] * o[4] + S.eigenRot[r * 3 + 2] * o[5];
        } else {
            o[16] = o[3]; o[17] = o[4]; o[18] = o[5];
        }
    }
    __syncthreads();
    const int cols = planar ? 9 : 12;
    const int nent = cols * (cols + 1) / 2;
    {                                                  // A^T P A upper triangle, entry-parallel (:482-486)
        int ea = 0, eb = 0;
        const bool have = tid < nent;                  // nent <= 78 <= blockDim.x
        if (have) {
            int rem = tid;
            while (rem >= cols - ea) { rem -= cols - ea; ++ea; }
            eb = ea + rem;
        }
        double s = 0.0;
        for (int base = 0; base < n; base += kTa) {
            const int cnt = min(kTa, n - base);
            __syncthreads();
            for (int e = tid; e < cnt * cols; e += blockDim.x) {       // the two rows of every observation of the tile
                const int pt = e / cols, col = e - pt * cols;
                const double* o = sc + (size_t)(base + pt) * kMlpnpScratch;
                double e0, e1;
                mlpnp_row_entry(o + 6, o + 16, planar, col, e0, e1);
                m_tile[col * kTPa + pt] = e0;
                m_tile[(12 + col) * kTPa + pt] = e1;
            }
            if (use_cov)
                for (int e = tid; e < cnt * 4; e += blockDim.x)
                    m_tile[(24 + (e & 3)) * kTPa + (e >> 2)] = sc[(size_t)(base + (e >> 2)) * kMlpnpScratch + 12 + (e & 3)];
            __syncthreads();
            if (have) {
                for (int i = 0; i < cnt; ++i) {
                    const double a0 = m_tile[ea * kTPa + i], a1 = m_tile[(12 + ea) * kTPa + i];
                    const double b0 = m_tile[eb * kTPa + i], b1 = m_tile[(12 + eb) * kTPa + i];
                    double w0 = b0, w1 = b1;
                    if (use_cov) {
                        w0 = rfma(m_tile[24 * kTPa + i], b0, m_tile[25 * kTPa + i] * b1);
                        w1 = rfma(m_tile[26 * kTPa + i], b0, m_tile[27 * kTPa + i] * b1);
                    }
                    s = rfma(a0, w0, s);
                    s = rfma(a1, w1, s);
                }
            }
        }
        if (have) S.AtPA[tri_idx(cols, ea, eb)] = s;
    }
    __syncthreads();
    if (tid < 32) {                                    // smallest eigenvector of A^T P A (:488-493), warp-cooperative
        if (planar) jacobi_lowest_warp<9, 1>(S.AtPA, S.ev, S.result1, s_rec, tid);
        else jacobi_lowest_warp<12, 1>(S.AtPA, S.ev, S.result1, s_rec, tid);
    }
    __syncthreads();
    if (tid == 0) {
        double result1[12];
        for (int i = 0; i < 12; ++i) result1[i] = S.result1[i];
        // first six observations for the +-t / 4-candidate test (:547-551, :590-594)
        double p6[18], f6[18];
        for (int q = 0; q < 6; ++q)
            for (int c = 0; c < 3; ++c) {
                f6[3 * q + c] = sc[(size_t)q * kMlpnpScratch + c];
                p6[3 * q + c] = sc[(size_t)q * kMlpnpScratch + 3 + c];
            }
        double Rout[9], tout[3];
        mlpnp_recover(result1, planar, S.eigenRot, p6, f6, Rout, tout);
        rot2rodrigues(Rout, S.x);
        S.x[3] = tout[0]; S.x[4] = tout[1]; S.x[5] = tout[2];
    }
    __syncthreads();
    // Gauss-Newton (MLPnPsolver.cpp:659-723)
    for (int it_cnt = 0; it_cnt < 5;) {
        for (int i = tid; i < n; i += blockDim.x) {
            double* o = sc + (size_t)i * kMlpnpScratch;
            const double nr[3] = {o[6], o[8], o[10]}, ns[3] = {o[7], o[9], o[11]};
            mlpnp_res_jac(o + 3, nr, ns, S.x, S.x + 3, o + 31, o + 19);
        }
        if (tid == 0) S.maxdl_bits = 0ull;
        __syncthreads();
        {                                              // A = J^T P J (36 entries) and g = J^T P r (6 entries)
            const bool isg = tid >= 36;
            const int ea = isg ? (tid - 36) % 6 : tid / 6, eb = isg ? 0 : tid % 6;
            double s = 0.0;
            for (int base = 0; base < n; base += kTg) {
                const int cnt = min(kTg, n - base);
                __syncthreads();
                for (int e = tid; e < cnt * 18; e += blockDim.x) {
                    const int pt = e / 18, j = e - 18 * pt;        // j < 12: J; 12..15: P; 16, 17: r
                    const double* o = sc + (size_t)(base + pt) * kMlpnpScratch;
                    m_tile[j * kTPg + pt] = (j < 12) ? o[19 + j] : (j < 16 ? o[12 + (j - 12)] : o[31 + (j - 16)]);
                }
                __syncthreads();
                if (tid < 42) {
                    for (int i = 0; i < cnt; ++i) {
                        const double Ja = m_tile[ea * kTPg + i], Jb = m_tile[(6 + ea) * kTPg + i];
                        double W0 = Ja, W1 = Jb;
                        if (use_cov) {
                            W0 = rfma(Ja, m_tile[12 * kTPg + i], Jb * m_tile[14 * kTPg + i]);
                            W1 = rfma(Ja, m_tile[13 * kTPg + i], Jb * m_tile[15 * kTPg + i]);
                        }
                        if (isg) { s = rfma(W0, m_tile[16 * kTPg + i], s); s = rfma(W1, m_tile[17 * kTPg + i], s); }
                        else     { s = rfma(W0, m_tile[eb * kTPg + i], s); s = rfma(W1, m_tile[(6 + eb) * kTPg + i], s); }
                    }
                }
            }
            if (tid < 42) { if (isg) S.g[ea] = s; else S.A[ea * 6 + eb] = s; }
        }
        __syncthreads();
        if (tid == 0) ldlt6_solve(S.A, S.g, S.dx);
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {    // max |J dx| (:712-713)
            const double* J = sc + (size_t)i * kMlpnpScratch + 19;
            for (int k = 0; k < 2; ++k) {
                const double* Jk = J + 6 * k;
                const double dl = Jk[0] * S.dx[0] + Jk[1] * S.dx[1] + Jk[2] * S.dx[2] + Jk[3] * S.dx[3] + Jk[4] * S.dx[4] + Jk[5] * S.dx[5];
                const double v = fabs(dl);
                if (v > 0.0) atomicMax(&S.maxdl_bits, (unsigned long long)__double_as_longlong(v));
            }
        }
        __syncthreads();
        if (tid == 0) {
            const double mdl = __longlong_as_double((long long)S.maxdl_bits);
            S.dec = mlpnp_gn_decide(S.dx, mdl);
            if (S.dec != 0)
                for (int c = 0; c < 6; ++c) S.x[c] = S.x[c] - S.dx[c];
        }
        __syncthreads();
        if (S.dec != 1) break;
        ++it_cnt;
    }
    if (tid == 0) {
        rodrigues2rot(S.x, pose_out);
        pose_out[9] = S.x[3]; pose_out[10] = S.x[4]; pose_out[11] = S.x[5];
    }
    __syncthreads();
}

// ------------------------------------------------------------- the split replay's eigen-solve
// One warp per parked problem: the 12x12 eigen-solve of M^T M (PnPsolver.cpp:380), the same cooperative schedule as inside the
// replay kernel (bit-identical), with a sixth of the replay CTA's registers.  kSelEigWarps problems per CTA.
#ifndef RSAC_SEL_EIG_WARPS
#define RSAC_SEL_EIG_WARPS 4
#endif
constexpr int kSelEigWarps = RSAC_SEL_EIG_WARPS;
constexpr size_t kSelEigSmemPerWarp = sizeof(double) * (78 + 48 + 4) + sizeof(double2) * kMaxSweepsRec * 66;
// (measured and rejected: the beta branches in this kernel too -- 0.370 instead of 0.364 ms per sweep in flight, the narrow kernel
// gets longer than what it frees; the recorded rotations in global memory instead of 12.7 KB of shared memory per warp -- 0.386)
static __global__ void __launch_bounds__(kSelEigWarps * 32) select_eigen_kernel(SelectCarry* carry, int C)
{
    extern __shared__ __align__(16) unsigned char eig_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kSelEigWarps + warp;
    if (c >= C) return;
    SelectCarry* cr = carry + c;
    if (!cr->waiting) return;
    double2* rec = reinterpret_cast<double2*>(eig_smem + (size_t)warp * kSelEigSmemPerWarp);
    double* MtM = reinterpret_cast<double*>(rec + kMaxSweepsRec * 66);
    double* U4 = MtM + 78;
    double* w4 = U4 + 48;
    for (int i = lane; i < 78; i += 32) MtM[i] = cr->MtM[i];
    __syncwarp();
    jacobi_lowest_warp<12, 4>(MtM, w4, U4, rec, lane);
    __syncwarp();
    for (int i = lane; i < 48; i += 32) cr->U4[i] = U4[i];
    if (lane < 4) cr->w4[lane] = w4[lane];
}

// ------------------------------------------------------------- the replay kernel
template <int MODEL>
__global__ void __launch_bounds__(select_threads<MODEL>(), kSelectCtasPerSm) ransac_select_kernel(SelectArgs a)
{
    using PT = typename ScoreModel<MODEL>::pose_t;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const ProblemMeta* m = a.metas + blockIdx.x;
    const int tid = threadIdx.x;
    const int words = m->words;
    uint32_t* bestmask = reinterpret_cast<uint32_t*>(smem_raw);   // [words]
    uint32_t* refmask = bestmask + words;                         // [words]
    int* prefix = reinterpret_cast<int*>(refmask + words);        // [words+1]

    __shared__ int s_found, s_cnt;
    __shared__ PT s_pose[12], s_bestpose[12];
    __shared__ typename std::conditional<MODEL == 0, EpnpShared, MlpnpShared>::type S;
    __shared__ double2 s_rec[kMaxSweepsRec * 66];       // recorded rotations of the refine eigen-solve
    __shared__ double s_tile[MODEL == 0 ? kSelTileDoubles : 1];
    SelectCarry* const cr = (MODEL == 0 && a.carry && a.split_phase != 0) ? a.carry + blockIdx.x : nullptr;
    if (cr && a.split_phase == 2 && !cr->waiting) return;       // decided in the first pass
    if (cr && a.split_phase == 1 && threadIdx.x == 0) cr->waiting = 0;

    ResultRec res;
    res.ok = 0; res.no_more = 0; res.n_inliers = 0; res.best_hyp = -1; res.refined = 0; res.n_refines = 0;
    res.best_count = 0; res.n_hyp = 0;
    for (int i = 0; i < 9; ++i) res.R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    res.t[0] = res.t[1] = res.t[2] = 0.0f; res.s = 1.0f;
    res.problem = a.problem_ids ? a.problem_ids[blockIdx.x] : a.problem_base + blockIdx.x; res.reserved[0] = res.reserved[1] = 0;

    RSAC_SEL_MARK(0);
    const int N = m->n, Hfull = m->H, minInl = m->min_inl;
    // early exit: a problem may only look at the upto[p] hypotheses it has
    int H = Hfull;
    int resume_stop = 0;
    if (a.ee) {
        const int up = a.ee[blockIdx.x];
        if (a.only_phase >= 0) {
            if (up >= 0) return;                       // decided in the main replay
            resume_stop = -up - 1;                     // where the main replay stopped
            res.n_refines = reinterpret_cast<const ResultRec*>(a.results)[blockIdx.x].n_refines;   // carried over
        } else {
            H = min(Hfull, up);
        }
    }
    const bool resume_phase = a.ee && a.only_phase >= 0;
    uint32_t* final_mask = a.masks + m->word_off;
    const int32_t* counts = a.counts + m->hyp_off;
    const PT* poses = reinterpret_cast<const PT*>(a.poses) + (size_t)m->hyp_off * 12;
    const bool discard = (MODEL == 1) && (a.flags & 4) != 0;   // RSAC_FLAG_MLPNP_DISCARD_REFINE (Q6)
    bool finished = false;

    if (N < minInl || H == 0) {              // PnPsolver.cpp:110-114 / MLPnPsolver.cpp:62-66
        res.no_more = 1;
        for (int w = tid; w < words; w += blockDim.x) final_mask[w] = 0u;
        finished = true;
    }

    int best = 0, bestH = -1, lastRefBestH = -2, lastRefH = -2, lastCntR = 0, mSel = 0;
    int cursor = 0;
    __shared__ unsigned long long s_key;

    // the best set becomes hypothesis h: its mask (exact re-evaluation) and the ordered index list
    auto set_best = [&](int h) {
        best = counts[h];
        bestH = h;
        if (tid < 12) s_bestpose[tid] = poses[(size_t)h * 12 + tid];
        __syncthreads();
        cta_score_exact<MODEL>(m, a, s_bestpose, bestmask, &s_cnt);
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
                RSAC_ASSERT(w * 32 + b < N && o < N);
                sel[o++] = (uint32_t)(w * 32 + b);
            }
        }
        __syncthreads();
    };

    // second pass of the split replay: back where the first pass parked the problem, behind the eigen-solve
    bool resume_refine = false;
    int h_resume = 0;
    if constexpr (MODEL == 0) {
        if (cr && a.split_phase == 2) {
            resume_refine = true;
            best = cr->best; bestH = cr->bestH; lastRefH = cr->lastRefH; lastCntR = cr->lastCntR; mSel = cr->mSel;
            cursor = cr->cursor; h_resume = cr->h; res.n_refines = cr->n_refines;
            if (tid < 12) s_bestpose[tid] = cr->bestpose[tid];
            for (int w = tid; w < words; w += blockDim.x) bestmask[w] = final_mask[w];      // parked there by the first pass
            for (int i = tid; i < 3; i += blockDim.x) S.C0[i] = cr->C0[i];
            for (int i = tid; i < 9; i += blockDim.x) { S.A[i] = cr->A[i]; S.CCi[i] = cr->CCi[i]; }
            for (int i = tid; i < 12; i += blockDim.x) S.cws[i] = cr->cws[i];
            for (int i = tid; i < 48; i += blockDim.x) S.U4[i] = cr->U4[i];
            for (int i = tid; i < 4; i += blockDim.x) S.w4[i] = cr->w4[i];
            __syncthreads();
        }
    }
    const int resume_at = resume_phase ? resume_stop : ((a.resume && !finished) ? a.resume[blockIdx.x] : 0);
    if (!finished && resume_at > 0) {
        // a later iterate() call (or the phase-C continuation): rebuild mnBestInliers / mvbBestInliers as the
        // scan left them before `cursor` (first strict maximum among the hypotheses with cnt >= minInliers)
        cursor = min(resume_at, H);
        if (tid == 0) s_key = 0ull;
        __syncthreads();
        for (int h = tid; h < cursor; h += blockDim.x)
            if (counts[h] >= minInl)
                atomicMax(&s_key, ((unsigned long long)(unsigned)counts[h] << 32) | (unsigned long long)(0xffffffffu - (unsigned)h));
        __syncthreads();
        const unsigned long long key = s_key;
        __syncthreads();
        if (key != 0ull) set_best((int)(0xffffffffu - (unsigned)(key & 0xffffffffull)));
    }

    while (!finished) {
        int h = H;
        if (resume_refine) {
            h = h_resume;                                       // the hypothesis whose Refine is under way
        } else {
            // next hypothesis with cnt >= minInliers (PnPsolver.cpp:146)
            if (tid == 0) s_found = H;
            __syncthreads();
            for (int base = cursor; base < H; base += blockDim.x) {   // uniform trip count: h is read after a barrier
                const int hc = base + tid;
                if (hc < H && counts[hc] >= minInl) atomicMin(&s_found, hc);
                __syncthreads();
                h = s_found;
                __syncthreads();
                if (h < H) break;
            }
            if (h >= H) break;

            RSAC_SEL_MARK(1);
            if (counts[h] > best) set_best(h);   // :149 strict: first maximum wins (Refine uses this set, :195-204)
            res.n_refines++;
        }

        if (discard) {
            // MLPnPsolver::Refine as shipped never copies its result into mRi/mti
            // (MLPnPsolver.cpp:290-296): the "refined" pose and inliers are the current hypothesis'
            if (h != lastRefH) {
                if (tid < 12) s_pose[tid] = poses[(size_t)h * 12 + tid];
                __syncthreads();
                cta_score_exact<MODEL>(m, a, s_pose, refmask, &s_cnt);
                lastCntR = s_cnt;
                lastRefH = h;
            }
        } else if (resume_refine || bestH != lastRefBestH) {
            if constexpr (MODEL == 0) {
                bool do_eigen = true;
                if (!resume_refine) {
                    refine_epnp_front(m, a, mSel, S, s_tile);
                    if (cr && a.split_phase == 1 && refine_epnp_has_eigen(a, mSel)) {
                        // park the problem in front of its eigen-solve: scan state, the best set's pose and mask, the front half's results
                        if (tid == 0) {
                            cr->waiting = 1;
                            cr->best = best; cr->bestH = bestH; cr->lastRefH = lastRefH; cr->lastCntR = lastCntR; cr->mSel = mSel;
                            cr->cursor = cursor; cr->h = h; cr->n_refines = res.n_refines;
                        }
                        if (tid < 12) cr->bestpose[tid] = s_bestpose[tid];
                        for (int w = tid; w < words; w += blockDim.x) final_mask[w] = bestmask[w];
                        for (int i = tid; i < 3; i += blockDim.x) cr->C0[i] = S.C0[i];
                        for (int i = tid; i < 9; i += blockDim.x) { cr->A[i] = S.A[i]; cr->CCi[i] = S.CCi[i]; }
                        for (int i = tid; i < 12; i += blockDim.x) cr->cws[i] = S.cws[i];
                        for (int i = tid; i < 78; i += blockDim.x) cr->MtM[i] = S.MtM[i];
                        return;
                    }
                } else {
                    do_eigen = false;                           // select_eigen_kernel did it
                }
                refine_epnp_back(m, a, mSel, S, s_rec, s_pose, s_tile, do_eigen);
            } else {
                refine_mlpnp(m, a, mSel, S, s_rec, s_pose);
            }
            RSAC_SEL_MARK(8);
            cta_score_exact<MODEL>(m, a, s_pose, refmask, &s_cnt);   // :220
            RSAC_SEL_MARK(9);
            lastCntR = s_cnt;
            lastRefBestH = bestH;
            resume_refine = false;
        }

        if (lastCntR > minInl) {                               // :225 strict
            res.ok = 1; res.refined = 1; res.n_inliers = lastCntR; res.n_hyp = h + 1;
            for (int i = 0; i < 9; ++i) res.R[i] = (float)s_pose[i];
            for (int i = 0; i < 3; ++i) res.t[i] = (float)s_pose[9 + i];
            for (int w = tid; w < words; w += blockDim.x) final_mask[w] = refmask[w];
            finished = true;
            break;
        }
        cursor = h + 1;
    }

    if (!finished && H < Hfull) {
        // early exit: every refine so far failed and the rest of the hypotheses has not been computed yet:
        // hand the problem to phase C; only the refine counter survives (in the result record)
        if (tid == 0) {
            int32_t* counters = a.ee + 4 * (size_t)a.C;
            a.ee[blockIdx.x] = -(H + 1);
            const int slot = atomicAdd(counters + 15, 1);                           // kCleanupCounter
            RSAC_ASSERT(slot >= 0 && slot < a.C);
            (a.ee + 3 * (size_t)a.C)[slot] = blockIdx.x;
            res.reserved[0] = 1;   // not decided yet
            reinterpret_cast<ResultRec*>(a.results)[blockIdx.x] = res;
        }
        return;
    }
    if (!finished) {                                           // :173-188 budget exhausted
        res.no_more = 1;
        res.n_hyp = H;
        if (best >= minInl) {
            res.ok = 1;
            res.n_inliers = best;
            for (int i = 0; i < 9; ++i) res.R[i] = (float)s_bestpose[i];
            for (int i = 0; i < 3; ++i) res.t[i] = (float)s_bestpose[9 + i];
            for (int w = tid; w < words; w += blockDim.x) final_mask[w] = bestmask[w];
        } else {
            for (int w = tid; w < words; w += blockDim.x) final_mask[w] = 0u;
        }
    }
    res.best_hyp = bestH;
    res.best_count = best;
    RSAC_SEL_MARK(10);
    if (tid == 0) {
        reinterpret_cast<ResultRec*>(a.results)[blockIdx.x] = res;
        if (a.results2) reinterpret_cast<ResultRec*>(a.results2)[blockIdx.x] = res;
        if (resume_phase) a.ee[blockIdx.x] = Hfull;     // everything of this problem exists now
    }
}

}  // namespace rsac
