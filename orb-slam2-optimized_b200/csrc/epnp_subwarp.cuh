// epnp_subwarp.cuh -- 4-point EPnP minimal solve (PnPsolver::compute_pose, PnPsolver.cpp:359-415) on THREE LANES
// per hypothesis, ten hypotheses per warp.
//
// Why three lanes.  compute_pose ends in three independent branches -- find_betas_approx_{1,2,3} + gauss_newton +
// compute_R_and_t (PnPsolver.cpp:395-405) -- which are 80 % of a 4-point solve (five Gauss-Newton steps with the
// reference's qr_solve, Horn's 4x4 eigen-solve, the reprojection error).  Lane b of a hypothesis runs branch b: the
// three branches are one instruction stream on three lanes instead of three copies of it in one thread, each lane
// holds one branch's state (no 2 KB local-memory stack per hypothesis), and a hypothesis' dependent chain is a third
// as long.  The front of the solve is shared work:
//   * control points, alphas, rho            redundantly on the three lanes (2.5 % of a solve)
//   * null space of the 8 x 12 M             Householder QR of M^T, COLUMN-distributed: lane l owns columns l, l+3, l+6;
//                                            the owner of column k forms reflector k and publishes it in shared memory,
//                                            the lanes apply it to the columns they own; the four null-space vectors
//                                            (H0..H7 e_{8+i}) on lanes 0, 1, 2 and 0 again
//   * L (6 x 10)                             two rows per lane
// Every ELEMENT sees exactly the operations of the serial routines in epnp.cuh / linalg.cuh in the same order
// (nullspace_qr_8x12, qr_lstsq, epnp_gauss_newton, epnp_horn, ...), so poses are bit-identical to the one-thread
// path and to the CPU checker (arithmetic contract) -- the parity tests do not change.
//
// Shared memory per hypothesis (kSwStride doubles, odd stride => the ten hypotheses of a warp hit distinct banks):
//   U4 [48] null-space basis | alphas [16] | rho [6] | scratch [68]: the reflectors of the QR, later L (60) |
//   pw, us [20 floats]
#pragma once
#include "common.cuh"
#include "epnp.cuh"

namespace rsac {

constexpr int kSwLanes = 3;                 // lanes per hypothesis = branches of compute_pose
constexpr int kSwHyps = 10;                 // hypotheses per warp (lanes 30, 31 shadow hypothesis 9)
// One CTA per SM whose warps walk the phases of the solve TOGETHER (a __syncthreads() between phases).  The solve is
// ~7.5 k instructions of mostly straight-line FP64 code = 120 KB, the instruction caches are 6 KB (L0, per scheduler)
// and 32 KB (L1.5): with warps scattered over the whole program every fetch goes to L2 and the kernel runs at the
// instruction-fetch rate (measured: 1.4 warp-instructions per clock and SM whatever the occupancy; ncu
// stall_no_instruction 3.9 cycles per issue).  In step, the warps of an SM share one phase's code (<= 2 k instructions).
#ifndef RSAC_SW_WARPS
#define RSAC_SW_WARPS 8                     // warps per CTA
#endif
#ifndef RSAC_SW_BLOCKS
#define RSAC_SW_BLOCKS 2                    // CTAs per SM: 16 warps = 160 hypotheses resident per SM, 128 registers per thread
#endif
#ifndef RSAC_SW_PHASE_SYNC
#define RSAC_SW_PHASE_SYNC 1
#endif
#if RSAC_SW_PHASE_SYNC
#define RSAC_SW_PHASE() __syncthreads()
#else
#define RSAC_SW_PHASE() __syncwarp()
#endif
constexpr int kSwWarps = RSAC_SW_WARPS;
constexpr int kSwThreads = kSwWarps * 32;
constexpr int kSwHypsPerBlock = kSwWarps * kSwHyps;
constexpr int kSwOffAl = 48, kSwOffRho = 64, kSwOffScr = 70, kSwOffPts = 138;
constexpr int kSwStride = 149;              // doubles per hypothesis (148 used)
constexpr size_t kSwSmemBytes = sizeof(double) * kSwStride * kSwHypsPerBlock;

// offset of reflector k (entries r = k..11) in the scratch area
__host__ __device__ constexpr int sw_refl_off(int k) { return k * 12 - (k * (k - 1)) / 2; }

// qr_lstsq<6,K> (linalg.cuh) with K a run-time value per lane (4, 3, 5 for the three branches): loops unrolled to
// the largest K and predicated; the operations on the live columns are qr_lstsq's, in its order
__device__ __forceinline__ bool sw_qr_lstsq(double (&A)[6][5], double (&bb)[6], int K, double (&x)[5])
{
    double rd[5];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (c < K && ok) {
            double s = 0.0;
#pragma unroll
            for (int r = c; r < 6; ++r) s = rfma(A[r][c], A[r][c], s);
            const double norm = rsqrt_exact(s);
            if (norm == 0.0) { ok = false; }
            else {
                const double alpha = (A[c][c] > 0.0) ? -norm : norm;
                A[c][c] = A[c][c] - alpha;
                double vtv = 0.0;
#pragma unroll
                for (int r = c; r < 6; ++r) vtv = rfma(A[r][c], A[r][c], vtv);
                const double tau = rdiv(2.0, vtv);
#pragma unroll
                for (int j = c + 1; j < 5; ++j) {
                    if (j < K) {
                        double d = 0.0;
#pragma unroll
                        for (int r = c; r < 6; ++r) d = rfma(A[r][c], A[r][j], d);
                        d = d * tau;
#pragma unroll
                        for (int r = c; r < 6; ++r) A[r][j] = rfma(-d, A[r][c], A[r][j]);
                    }
                }
                double d = 0.0;
#pragma unroll
                for (int r = c; r < 6; ++r) d = rfma(A[r][c], bb[r], d);
                d = d * tau;
#pragma unroll
                for (int r = c; r < 6; ++r) bb[r] = rfma(-d, A[r][c], bb[r]);
                rd[c] = alpha;
            }
        }
    }
    if (!ok) return false;
    double rmax = 0.0, rmin = fabs(rd[0]);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (c < K) {
            const double a = fabs(rd[c]);
            if (a > rmax) rmax = a;
            if (a < rmin) rmin = a;
        }
    }
    if (!(rmin > rmax * 1e-7)) return false;
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        if (i < K) {
            double sum = 0.0;
#pragma unroll
            for (int j = i + 1; j < 5; ++j)
                if (j < K) sum = rfma(A[i][j], x[j], sum);
            x[i] = rdiv(bb[i] - sum, rd[i]);
        }
    }
    return true;
}

// rank-deficient fallback of lstsq<6,K> (never taken on sane data): the serial SVD solve, out of line
template <int K>
__device__ __noinline__ void sw_svd_fallback(const double* L, const double* rho, const int* cols, double* x)
{
    double LK[6 * K];
    for (int i = 0; i < 6; ++i)
        for (int c = 0; c < K; ++c) LK[i * K + c] = L[i * 10 + cols[c]];
    svd_lstsq<6, K>(LK, rho, x);
}

// EPnP minimal solves of hypotheses [h_lo, h_lo + span) of the listed problems (list == nullptr: all C problems).
// Persistent grid-stride form over (problem, hypothesis) slots, ten slots per warp.
__global__ void __launch_bounds__(kSwThreads, RSAC_SW_BLOCKS)
epnp_minimal_subwarp_kernel(const ProblemMeta* __restrict__ metas, int C, const int32_t* __restrict__ list,
                            const int32_t* __restrict__ list_count, int h_lo, int span, const uint32_t* __restrict__ tables,
                            const float4* __restrict__ cA, const float4* __restrict__ cC, float* __restrict__ poses)
{
    extern __shared__ double sw_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool shadow = lane >= kSwLanes * kSwHyps;           // lanes 30, 31 mirror hypothesis 9 (branches 0, 1), never write
    const int slot = shadow ? kSwHyps - 1 : lane / kSwLanes;  // hypothesis of this lane within the warp
    const int l = shadow ? lane - kSwLanes * kSwHyps : lane - slot * kSwLanes;   // branch / column group of this lane
    const int base = slot * kSwLanes;                         // lane 0 of the hypothesis
    double* S = sw_smem + (size_t)(warp * kSwHyps + slot) * kSwStride;
    double* sU = S;
    double* sAl = S + kSwOffAl;
    double* sRho = S + kSwOffRho;
    double* sScr = S + kSwOffScr;
    float* sPts = reinterpret_cast<float*>(S + kSwOffPts);    // pw [12] us [8]

    const int np = list ? *list_count : C;
    const int64_t total = (int64_t)np * span;
    // rounds are per CTA (uniform trip count: the phases are separated by CTA-wide barriers); a warp whose ten slots
    // lie past the end runs the round on dummy data
    for (int64_t r0 = (int64_t)blockIdx.x * kSwHypsPerBlock; r0 < total; r0 += (int64_t)gridDim.x * kSwHypsPerBlock) {
        const int64_t t0 = r0 + (int64_t)warp * kSwHyps;
        // `live`: this lane's hypothesis exists (its pose is stored).  Slots past the end of the work and hypotheses
        // beyond a problem's H (ragged batches) run on a fixed well-posed dummy problem instead: every lane executes the
        // same instruction stream on sane numbers, nothing of it is stored to global memory
        int64_t t = t0 + slot;
        bool live = !shadow;
        if (t >= total) { t = r0; live = false; }
        const int kq = (int)(t / span);
        const int h = h_lo + (int)(t - (int64_t)kq * span);
        const int p = list ? list[kq] : kq;
        const ProblemMeta& m = metas[p];
        const bool real = h < m.H;                            // the table and the correspondences may be read
        live = live && real;
        const bool writer = !shadow && l == 0;                // shared-memory stores of per-hypothesis values
        const Cam cam = {m.fx, m.fy, m.cx, m.cy};

        // ---- control points, alphas, rho: redundantly on the three lanes (PnPsolver.cpp:296-343, 639-647) ----
        double cws[12], CCi[9];
        {
            const uint32_t* idx = tables + m.table_off + (size_t)(real ? h : 0) * 4;
            double pw[12];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // dummy problem: four non-coplanar points in front of the camera and plausible pixels
                float4 a = make_float4((float)(i & 1), (float)(i >> 1), 5.0f + 0.75f * (float)(i * i), 0.0f);
                float4 q = make_float4(300.0f + 90.0f * (float)(i & 1), 200.0f + 80.0f * (float)(i >> 1), 0.0f, 0.0f);
                if (real) {
                    RSAC_ASSERT(p >= 0 && p < C && h >= 0 && h < m.H && idx[i] < (uint32_t)m.n);
                    const size_t ci = (size_t)m.corr_off + idx[i];
                    a = cA[ci];
                    q = cC[ci];
                }
                pw[3 * i] = (double)a.x; pw[3 * i + 1] = (double)a.y; pw[3 * i + 2] = (double)a.z;   // add_correspondence (:288-294)
                if (writer) {
                    sPts[3 * i] = a.x; sPts[3 * i + 1] = a.y; sPts[3 * i + 2] = a.z;
                    sPts[12 + 2 * i] = q.x; sPts[13 + 2 * i] = q.y;
                }
            }
            double C0[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) s += pw[i * 3 + c];
                C0[c] = rdiv(s, 4.0);
            }
            double A3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double d0 = pw[i * 3 + 0] - C0[0], d1 = pw[i * 3 + 1] - C0[1], d2 = pw[i * 3 + 2] - C0[2];
                A3[0] = rfma(d0, d0, A3[0]); A3[1] = rfma(d0, d1, A3[1]); A3[2] = rfma(d0, d2, A3[2]);
                A3[4] = rfma(d1, d1, A3[4]); A3[5] = rfma(d1, d2, A3[5]);
                A3[8] = rfma(d2, d2, A3[8]);
            }
            epnp_control_points<true>(C0, A3, 4, cws);
            epnp_cc_inverse(cws, CCi);
            if (writer) {
                double rho[6];
                epnp_rho(cws, rho);
#pragma unroll
                for (int i = 0; i < 6; ++i) sRho[i] = rho[i];
            }
            // alphas: lane l computes the points l and (lane 0) 3
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int i = l + 3 * v;
                if (i < 4 && !shadow) {
                    double al[4];
                    // static register indexing of pw: select the point by predication
                    double q0 = pw[0], q1 = pw[1], q2 = pw[2];
                    if (i == 1) { q0 = pw[3]; q1 = pw[4]; q2 = pw[5]; }
                    if (i == 2) { q0 = pw[6]; q1 = pw[7]; q2 = pw[8]; }
                    if (i == 3) { q0 = pw[9]; q1 = pw[10]; q2 = pw[11]; }
                    const double pt[3] = {q0, q1, q2};
                    epnp_alphas(pt, cws, CCi, al);
                    sAl[4 * i] = al[0]; sAl[4 * i + 1] = al[1]; sAl[4 * i + 2] = al[2]; sAl[4 * i + 3] = al[3];
                }
            }
        }
        RSAC_SW_PHASE();

        // ---- null space of M: Householder QR of A = M^T (12 x 8), column-distributed (nullspace_qr_8x12) ----
        // slot s of this lane = column l + 3 s (column 8 does not exist: lane 2, slot 2)
        double Ac[3][12];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int c = l + 3 * s;
            const int i = min(c >> 1, 3);
            const double u = (double)sPts[12 + 2 * i], v = (double)sPts[13 + 2 * i];
            const bool odd = c & 1;
            // column 2i = row 0 of point i: [a fx, 0, a (cx - u)] per control point; column 2i+1: [0, a fy, a (cy - v)] (:367-377)
            const double f = odd ? cam.fy : cam.fx;
            const double cc = odd ? (cam.cy - v) : (cam.cx - u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double a = sAl[4 * i + j];
                const double af = a * f;
                Ac[s][3 * j] = odd ? 0.0 : af;
                Ac[s][3 * j + 1] = odd ? af : 0.0;
                Ac[s][3 * j + 2] = a * cc;
            }
        }
        double tau[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            constexpr int dummy = 0; (void)dummy;
            const int owner = k % 3, ks = k / 3;
            double* V = sScr + sw_refl_off(k);
            double tk = 0.0;
            if (l == owner) {
                double s = 0.0;
#pragma unroll
                for (int r = k; r < 12; ++r) s = rfma(Ac[ks][r], Ac[ks][r], s);
                const double norm = rsqrt_exact(s);
                if (norm != 0.0) {
                    const double alpha = (Ac[ks][k] > 0.0) ? -norm : norm;
                    Ac[ks][k] = Ac[ks][k] - alpha;
                    double vtv = 0.0;
#pragma unroll
                    for (int r = k; r < 12; ++r) vtv = rfma(Ac[ks][r], Ac[ks][r], vtv);
                    tk = rdiv(2.0, vtv);
                }
                if (!shadow) {
#pragma unroll
                    for (int r = k; r < 12; ++r) V[r - k] = Ac[ks][r];
                }
            }
            __syncwarp();
            tk = __shfl_sync(FULL, tk, base + owner);
            tau[k] = tk;
            if (tk != 0.0) {
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    if (3 * s + 2 > k) {                       // some lane owns a column > k in this slot
                        const int c = l + 3 * s;
                        if (c > k && c < 8) {
                            double d = 0.0;
#pragma unroll
                            for (int r = k; r < 12; ++r) d = rfma(V[r - k], Ac[s][r], d);
                            d = d * tk;
#pragma unroll
                            for (int r = k; r < 12; ++r) Ac[s][r] = rfma(-d, V[r - k], Ac[s][r]);
                        }
                    }
                }
            }
        }
        // null-space vectors: lane l forms vector l, lane 0 also vector 3
#pragma unroll 1
        for (int v = 0; v < 2; ++v) {
            const int i = l + 3 * v;
            if (i < 4) {
                double y[12];
#pragma unroll
                for (int r = 0; r < 12; ++r) y[r] = (r == 8 + i) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 7; k >= 0; --k) {
                    if (tau[k] == 0.0) continue;
                    const double* V = sScr + sw_refl_off(k);
                    double d = 0.0;
#pragma unroll
                    for (int r = k; r < 12; ++r) d = rfma(V[r - k], y[r], d);
                    d = d * tau[k];
#pragma unroll
                    for (int r = k; r < 12; ++r) y[r] = rfma(-d, V[r - k], y[r]);
                }
                if (!shadow) {
#pragma unroll
                    for (int r = 0; r < 12; ++r) sU[r * 4 + i] = y[r];
                }
            }
        }
        RSAC_SW_PHASE();

        // ---- L (6 x 10): rows 2l, 2l+1 on lane l (compute_L_6x10, :604-637); overlays the reflectors ----
        {
#pragma unroll 1
            for (int v = 0; v < 2; ++v) {
                const int row = 2 * l + v;
                // control-point pairs in the order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3)
                const int a = row < 3 ? 0 : (row < 5 ? 1 : 2);
                const int b = row < 3 ? row + 1 : (row < 5 ? row - 1 : 3);
                double dv[4][3];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) dv[i][c] = sU[(3 * a + c) * 4 + i] - sU[(3 * b + c) * 4 + i];
#define RSAC_DOT3(x, y) ((x)[0] * (y)[0] + (x)[1] * (y)[1] + (x)[2] * (y)[2])
                double Lr[10];
                Lr[0] = RSAC_DOT3(dv[0], dv[0]);
                Lr[1] = 2.0 * RSAC_DOT3(dv[0], dv[1]);
                Lr[2] = RSAC_DOT3(dv[1], dv[1]);
                Lr[3] = 2.0 * RSAC_DOT3(dv[0], dv[2]);
                Lr[4] = 2.0 * RSAC_DOT3(dv[1], dv[2]);
                Lr[5] = RSAC_DOT3(dv[2], dv[2]);
                Lr[6] = 2.0 * RSAC_DOT3(dv[0], dv[3]);
                Lr[7] = 2.0 * RSAC_DOT3(dv[1], dv[3]);
                Lr[8] = 2.0 * RSAC_DOT3(dv[2], dv[3]);
                Lr[9] = RSAC_DOT3(dv[3], dv[3]);
#undef RSAC_DOT3
                if (!shadow) {
#pragma unroll
                    for (int j = 0; j < 10; ++j) sScr[row * 10 + j] = Lr[j];
                }
            }
        }
        __syncwarp();

        // ---- branch l: find_betas_approx_{l+1} (:520-602) + gauss_newton (:675-691) ----
        double betas[4];
        {
            const int K = l == 0 ? 4 : (l == 1 ? 3 : 5);
            // columns of L used by the approximation: {0,1,3,6} / {0,1,2} / {0,1,2,3,4}
            const int c2 = l == 0 ? 3 : 2, c3 = l == 0 ? 6 : 3;
            double A[6][5], bb[6], x[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                A[i][0] = sScr[i * 10 + 0]; A[i][1] = sScr[i * 10 + 1]; A[i][2] = sScr[i * 10 + c2];
                A[i][3] = sScr[i * 10 + c3]; A[i][4] = sScr[i * 10 + 4];
                bb[i] = sRho[i];
            }
            if (!sw_qr_lstsq(A, bb, K, x)) {
                double rho[6];
                for (int i = 0; i < 6; ++i) rho[i] = sRho[i];
                const int cols[5] = {0, 1, c2, c3, 4};
                if (l == 0) sw_svd_fallback<4>(sScr, rho, cols, x);
                else if (l == 1) sw_svd_fallback<3>(sScr, rho, cols, x);
                else sw_svd_fallback<5>(sScr, rho, cols, x);
            }
            if (l == 0) {                                      // find_betas_approx_1
                const double sg = (x[0] < 0) ? -1.0 : 1.0;        // b0 < 0: sqrt(-b0), -b_i / beta0 (exact sign flips)
                betas[0] = rsqrt_exact(sg * x[0]);
                betas[1] = rdiv(sg * x[1], betas[0]); betas[2] = rdiv(sg * x[2], betas[0]); betas[3] = rdiv(sg * x[3], betas[0]);
            } else {                                           // find_betas_approx_2 / _3
                const bool neg = x[0] < 0;
                betas[0] = rsqrt_exact(neg ? -x[0] : x[0]);
                const double x2 = neg ? -x[2] : x[2];
                betas[1] = (x2 > 0) ? rsqrt_exact(x2) : 0.0;
                if (x[1] < 0) betas[0] = -betas[0];
                betas[2] = (l == 2) ? rdiv(x[3], betas[0]) : 0.0;
                betas[3] = 0.0;
            }
            RSAC_SW_PHASE();
            epnp_gauss_newton_reg((const double*)sScr, (const double*)sRho, betas);
            RSAC_SW_PHASE();
        }

        // ---- compute_R_and_t (:504-515) of branch l ----
        double R[9], tr[3], rep;
        {
            double ccs[12];
            epnp_ccs(betas, sU, ccs);
            double pw[12], pcs[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) pw[i] = (double)sPts[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double al[4] = {sAl[4 * i], sAl[4 * i + 1], sAl[4 * i + 2], sAl[4 * i + 3]};
                epnp_pc(al, ccs, pcs + 3 * i);
            }
            if (pcs[2] < 0.0) {                                // solve_for_sign (:495-502)
#pragma unroll
                for (int i = 0; i < 12; ++i) pcs[i] = -pcs[i];
            }
            double pc0[3], pw0[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double sc = 0.0, sw = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) sc += pcs[i * 3 + c];
#pragma unroll
                for (int i = 0; i < 4; ++i) sw += pw[i * 3 + c];
                pc0[c] = rdiv(sc, 4.0);
                pw0[c] = rdiv(sw, 4.0);
            }
            double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) M[r * 3 + c] += (pcs[i * 3 + r] - pc0[r]) * (pw[i * 3 + c] - pw0[c]);
            epnp_horn<true>(M, pc0, pw0, R, tr);
            double sum2 = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sum2 += epnp_reproj_term(R, tr, pw + 3 * i, (double)sPts[12 + 2 * i], (double)sPts[13 + 2 * i], cam);
            rep = rdiv(sum2, 4.0);
        }
        // smallest reprojection error, ties to the lower branch (:407-409)
        const double rep0 = __shfl_sync(FULL, rep, base), rep1 = __shfl_sync(FULL, rep, base + 1);
        const double rep2 = __shfl_sync(FULL, rep, min(base + 2, 31));
        int N = 0;
        if (rep1 < rep0) N = 1;
        if (rep2 < (N == 1 ? rep1 : rep0)) N = 2;
        if (live && l == N) {
            float* out = poses + ((int64_t)m.hyp_off + h) * 12;
#pragma unroll
            for (int i = 0; i < 9; ++i) out[i] = (float)R[i];
            out[9] = (float)tr[0]; out[10] = (float)tr[1]; out[11] = (float)tr[2];
        }
        RSAC_SW_PHASE();                                       // the next round overwrites the shared arrays
    }
}

}  // namespace rsac
