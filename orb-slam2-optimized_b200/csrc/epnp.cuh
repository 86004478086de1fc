// epnp.cuh -- EPnP building blocks (device), following reference src/PnPsolver.cpp.
//
// The same pieces serve two schedules:
//   * minimal solve (n = 4): one thread per hypothesis, everything thread-private;
//   * n-point refine: one CTA per problem, sums evaluated "entry-parallel" -- one thread
//     per OUTPUT entry walks the points in index order -- so that every sum is formed in
//     exactly the serial order of the CPU checker, and the dense tail (12x12 eigen-solve,
//     betas, Gauss-Newton, Horn) runs on one thread with the functions below.
// FP64, -fmad=false, operation order fixed (arithmetic contract).
#pragma once
#include "linalg.cuh"

namespace rsac {

// diagnostic: clock64() at the phase boundaries of the minimal solve, thread 0 of block 0 (rsac_debug_solve_clocks)
#ifdef __CUDACC__
static __device__ long long g_solve_clocks[16];
#endif
#ifdef __CUDA_ARCH__
#define RSAC_SOLVE_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_solve_clocks[i] = clock64(); } while (0)
#else
#define RSAC_SOLVE_MARK(i) do { } while (0)
#endif

struct Cam { double fx, fy, cx, cy; };

// array view with an element stride: lets a thread keep an array in shared memory as one column of doubles
// (element i of thread t at base[i * blockDim.x + t]: conflict-free) behind the same indexing code
struct StridedD {
    double* p;
    int st;
    __host__ __device__ double& operator[](int i) const { return p[(size_t)i * st]; }
};

// PnPsolver::choose_control_points (PnPsolver.cpp:296-321) after the sums:
// C0 = centroid (already divided by n), A = upper triangle of PW0^T PW0.
template <bool STATIC_SORT = false>
__host__ __device__ inline void epnp_control_points(const double* C0, double* A, int n, double* cws /*4x3*/)
{
    for (int c = 0; c < 3; ++c) cws[c] = C0[c];
    double DC[3], UCt[9];
    jacobi_eig<double, 3, STATIC_SORT>(A, DC, UCt);
    for (int i = 0; i < 3; ++i) {
        const double k = rsqrt_exact(rdiv(DC[i], (double)n));
        for (int c = 0; c < 3; ++c) cws[(i + 1) * 3 + c] = cws[c] + k * UCt[c * 3 + i];
    }
}

// PnPsolver::compute_barycentric_coordinates (:323-331): CC and its inverse
__host__ __device__ inline void epnp_cc_inverse(const double* cws, double* CCi)
{
    double CC[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 1; j < 4; ++j) CC[i * 3 + (j - 1)] = cws[j * 3 + i] - cws[i];
    inv3(CC, CCi);
}

// (:334-342) alphas of one point
__host__ __device__ inline void epnp_alphas(const double* pw, const double* cws, const double* CCi, double* a)
{
    const double d0 = pw[0] - cws[0], d1 = pw[1] - cws[1], d2 = pw[2] - cws[2];
    for (int j = 0; j < 3; ++j) a[j + 1] = CCi[j * 3 + 0] * d0 + CCi[j * 3 + 1] * d1 + CCi[j * 3 + 2] * d2;
    a[0] = 1.0 - a[1] - a[2] - a[3];
}

// the two rows of M contributed by one correspondence (PnPsolver.cpp:367-377)
__host__ __device__ inline void epnp_m_rows(const double* a, double u, double v, const Cam& k, double* r0, double* r1)
{
    for (int j = 0; j < 4; ++j) {
        r0[3 * j] = a[j] * k.fx; r0[3 * j + 1] = 0.0;        r0[3 * j + 2] = a[j] * (k.cx - u);
        r1[3 * j] = 0.0;         r1[3 * j + 1] = a[j] * k.fy; r1[3 * j + 2] = a[j] * (k.cy - v);
    }
}

// one entry of a row of M without materialising the row (entry-parallel MtM in the refine stage)
__host__ __device__ inline void epnp_m_entry(const double* a, double u, double v, const Cam& k, int col, double& e0, double& e1)
{
    const int j = col / 3, c = col - 3 * j;
    if (c == 0)      { e0 = a[j] * k.fx;       e1 = 0.0; }
    else if (c == 1) { e0 = 0.0;               e1 = a[j] * k.fy; }
    else             { e0 = a[j] * (k.cx - u); e1 = a[j] * (k.cy - v); }
}

// PnPsolver::compute_L_6x10 (:604-637), U4[r*4+i] = i-th smallest eigenvector, row r
template <class LT>
__host__ __device__ inline void epnp_L_6x10(const double* U4, LT L /*6x10*/, int ust = 1)
{
    double dv[4][6][3];
    for (int i = 0; i < 4; ++i) {
        int a = 0, b = 1;
        for (int j = 0; j < 6; ++j) {
            for (int c = 0; c < 3; ++c) dv[i][j][c] = U4[((3 * a + c) * 4 + i) * ust] - U4[((3 * b + c) * 4 + i) * ust];
            b++;
            if (b > 3) { a++; b = a + 1; }
        }
    }
#define RSAC_DOT3(x, y) ((x)[0] * (y)[0] + (x)[1] * (y)[1] + (x)[2] * (y)[2])
    for (int i = 0; i < 6; ++i) {
        L[i * 10 + 0] = RSAC_DOT3(dv[0][i], dv[0][i]);
        L[i * 10 + 1] = 2.0 * RSAC_DOT3(dv[0][i], dv[1][i]);
        L[i * 10 + 2] = RSAC_DOT3(dv[1][i], dv[1][i]);
        L[i * 10 + 3] = 2.0 * RSAC_DOT3(dv[0][i], dv[2][i]);
        L[i * 10 + 4] = 2.0 * RSAC_DOT3(dv[1][i], dv[2][i]);
        L[i * 10 + 5] = RSAC_DOT3(dv[2][i], dv[2][i]);
        L[i * 10 + 6] = 2.0 * RSAC_DOT3(dv[0][i], dv[3][i]);
        L[i * 10 + 7] = 2.0 * RSAC_DOT3(dv[1][i], dv[3][i]);
        L[i * 10 + 8] = 2.0 * RSAC_DOT3(dv[2][i], dv[3][i]);
        L[i * 10 + 9] = RSAC_DOT3(dv[3][i], dv[3][i]);
    }
#undef RSAC_DOT3
}

__host__ __device__ inline double sqdist3(const double* a, const double* b)
{
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// PnPsolver::compute_rho (:639-647)
__host__ __device__ inline void epnp_rho(const double* cws, double* rho)
{
    rho[0] = sqdist3(cws + 0, cws + 3);
    rho[1] = sqdist3(cws + 0, cws + 6);
    rho[2] = sqdist3(cws + 0, cws + 9);
    rho[3] = sqdist3(cws + 3, cws + 6);
    rho[4] = sqdist3(cws + 3, cws + 9);
    rho[5] = sqdist3(cws + 6, cws + 9);
}

// find_betas_approx_{1,2,3} (:520-602)
template <class LT>
__host__ __device__ inline void epnp_betas_approx_1(LT L, const double* rho, double* betas)
{
    double L4[24], b4[4];
    for (int i = 0; i < 6; ++i) {
        L4[i * 4 + 0] = L[i * 10 + 0]; L4[i * 4 + 1] = L[i * 10 + 1]; L4[i * 4 + 2] = L[i * 10 + 3]; L4[i * 4 + 3] = L[i * 10 + 6];
    }
    lstsq<6, 4>(L4, rho, b4);
    if (b4[0] < 0) {
        betas[0] = sqrt(-b4[0]);
        betas[1] = -b4[1] / betas[0];
        betas[2] = -b4[2] / betas[0];
        betas[3] = -b4[3] / betas[0];
    } else {
        betas[0] = sqrt(b4[0]);
        betas[1] = b4[1] / betas[0];
        betas[2] = b4[2] / betas[0];
        betas[3] = b4[3] / betas[0];
    }
}

template <class LT>
__host__ __device__ inline void epnp_betas_approx_2(LT L, const double* rho, double* betas)
{
    double L3[18], b3[3];
    for (int i = 0; i < 6; ++i) {
        L3[i * 3 + 0] = L[i * 10 + 0]; L3[i * 3 + 1] = L[i * 10 + 1]; L3[i * 3 + 2] = L[i * 10 + 2];
    }
    lstsq<6, 3>(L3, rho, b3);
    if (b3[0] < 0) {
        betas[0] = sqrt(-b3[0]);
        betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
    } else {
        betas[0] = sqrt(b3[0]);
        betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
    }
    if (b3[1] < 0) betas[0] = -betas[0];
    betas[2] = 0.0;
    betas[3] = 0.0;
}

template <class LT>
__host__ __device__ inline void epnp_betas_approx_3(LT L, const double* rho, double* betas)
{
    double L5[30], b5[5];
    for (int i = 0; i < 6; ++i)
        for (int c = 0; c < 5; ++c) L5[i * 5 + c] = L[i * 10 + c];
    lstsq<6, 5>(L5, rho, b5);
    if (b5[0] < 0) {
        betas[0] = sqrt(-b5[0]);
        betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
    } else {
        betas[0] = sqrt(b5[0]);
        betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
    }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0];
    betas[3] = 0.0;
}

// PnPsolver::compute_A_and_b_gauss_newton (:649-673)
template <class LT>
__host__ __device__ inline void epnp_gn_system(LT L, const double* rho, const double* bt, double* A /*6x4*/, double* b)
{
    for (int i = 0; i < 6; ++i) {
        double l[10];
        for (int j = 0; j < 10; ++j) l[j] = L[i * 10 + j];
        const double Lt[4][4] = {{2 * l[0], l[1], l[3], l[6]},
                                 {l[1], 2 * l[2], l[4], l[7]},
                                 {l[3], l[4], 2 * l[5], l[8]},
                                 {l[6], l[7], l[8], 2 * l[9]}};
        for (int r = 0; r < 4; ++r)
            A[i * 4 + r] = rfma(Lt[r][3], bt[3], rfma(Lt[r][2], bt[2], rfma(Lt[r][1], bt[1], Lt[r][0] * bt[0])));   /* :659 */
        double q = (l[0] * bt[0]) * bt[0];                                                          /* :661-671 */
        q = rfma(l[1] * bt[0], bt[1], q);
        q = rfma(l[2] * bt[1], bt[1], q);
        q = rfma(l[3] * bt[0], bt[2], q);
        q = rfma(l[4] * bt[1], bt[2], q);
        q = rfma(l[5] * bt[2], bt[2], q);
        q = rfma(l[6] * bt[0], bt[3], q);
        q = rfma(l[7] * bt[1], bt[3], q);
        q = rfma(l[8] * bt[2], bt[3], q);
        q = rfma(l[9] * bt[3], bt[3], q);
        b[i] = rho[i] - q;
    }
}

// PnPsolver::qr_solve (:693-796): Householder QR of the 6x4 system with max-abs column
// scaling; a zero column returns with X untouched (:722-727).  Thread-private scratch
// replaces the reference's function-static A1/A2 (:696-697).
__host__ __device__ inline void epnp_qr_solve(double* A /*6x4*/, double* b, double* X)
{
    constexpr int nr = 6, nc = 4;
    double A1[nc], A2[nc];
    for (int k = 0; k < nc; ++k) {
        double eta = fabs(A[k * nc + k]);
        for (int i = k + 1; i < nr; ++i) {
            const double elt = fabs(A[i * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) return;
        const double inv_eta = 1. / eta;
        double sum = 0.0;
        for (int i = k; i < nr; ++i) {
            A[i * nc + k] *= inv_eta;
            sum = rfma(A[i * nc + k], A[i * nc + k], sum);
        }
        double sigma = sqrt(sum);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; ++j) {
            double s = 0;
            for (int i = k; i < nr; ++i) s = rfma(A[i * nc + k], A[i * nc + j], s);
            const double tau = s / A1[k];
            for (int i = k; i < nr; ++i) A[i * nc + j] = rfma(-tau, A[i * nc + k], A[i * nc + j]);
        }
    }
    for (int j = 0; j < nc; ++j) {
        double tau = 0;
        for (int i = j; i < nr; ++i) tau = rfma(A[i * nc + j], b[i], tau);
        tau /= A1[j];
        for (int i = j; i < nr; ++i) b[i] = rfma(-tau, A[i * nc + j], b[i]);
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; --i) {
        double sum = 0;
        for (int j = i + 1; j < nc; ++j) sum = rfma(A[i * nc + j], X[j], sum);
        X[i] = (b[i] - sum) / A2[i];
    }
}

// PnPsolver::gauss_newton (:675-691): exactly five steps
template <class LT>
__host__ __device__ inline void epnp_gauss_newton(LT L, const double* rho, double* betas)
{
    double A[24], B[6], X[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
        epnp_gn_system(L, rho, betas, A, B);
        epnp_qr_solve(A, B, X);
        for (int i = 0; i < 4; ++i) betas[i] += X[i];
    }
}

// One Gauss-Newton step (epnp_gn_system + epnp_qr_solve) with every array in registers: all loops unrolled (static
// indices only) and reflector k applied to b right after the columns -- column k is final then and b has seen
// H_0..H_{k-1}, so b goes through the reference's operation sequence (PnPsolver.cpp:766-778) unchanged while A1 and
// the lower part of the factor die early.  Element for element the arithmetic of the two routines above.
template <class LT>
__host__ __device__ __forceinline__ void epnp_gn_step_reg(LT L, const double* rho, const double (&bt)[4], double (&X)[4])
{
    double A[6][4], b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double l[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) l[j] = L[i * 10 + j];
        A[i][0] = rfma(l[6], bt[3], rfma(l[3], bt[2], rfma(l[1], bt[1], (2 * l[0]) * bt[0])));        /* :659 */
        A[i][1] = rfma(l[7], bt[3], rfma(l[4], bt[2], rfma(2 * l[2], bt[1], l[1] * bt[0])));
        A[i][2] = rfma(l[8], bt[3], rfma(2 * l[5], bt[2], rfma(l[4], bt[1], l[3] * bt[0])));
        A[i][3] = rfma(2 * l[9], bt[3], rfma(l[8], bt[2], rfma(l[7], bt[1], l[6] * bt[0])));
        double q = (l[0] * bt[0]) * bt[0];                                                           /* :661-671 */
        q = rfma(l[1] * bt[0], bt[1], q);
        q = rfma(l[2] * bt[1], bt[1], q);
        q = rfma(l[3] * bt[0], bt[2], q);
        q = rfma(l[4] * bt[1], bt[2], q);
        q = rfma(l[5] * bt[2], bt[2], q);
        q = rfma(l[6] * bt[0], bt[3], q);
        q = rfma(l[7] * bt[1], bt[3], q);
        q = rfma(l[8] * bt[2], bt[3], q);
        q = rfma(l[9] * bt[3], bt[3], q);
        b[i] = rho[i] - q;
    }
    double A2[4];
    bool singular = false;                       // eta == 0: qr_solve returns with X untouched (:722-727)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!singular) {
            double eta = fabs(A[k][k]);
#pragma unroll
            for (int i = k + 1; i < 6; ++i) {
                const double elt = fabs(A[i][k]);
                if (eta < elt) eta = elt;
            }
            if (eta == 0) {
                singular = true;
            } else {
                const double inv_eta = rdiv(1., eta);
                double sum = 0.0;
#pragma unroll
                for (int i = k; i < 6; ++i) {
                    A[i][k] *= inv_eta;
                    sum = rfma(A[i][k], A[i][k], sum);
                }
                double sigma = rsqrt_exact(sum);
                if (A[k][k] < 0) sigma = -sigma;
                A[k][k] += sigma;
                const double A1k = sigma * A[k][k];
                A2[k] = -eta * sigma;
#pragma unroll
                for (int j = k + 1; j < 4; ++j) {
                    double s = 0;
#pragma unroll
                    for (int i = k; i < 6; ++i) s = rfma(A[i][k], A[i][j], s);
                    const double tau = rdiv(s, A1k);
#pragma unroll
                    for (int i = k; i < 6; ++i) A[i][j] = rfma(-tau, A[i][k], A[i][j]);
                }
                double tau = 0;
#pragma unroll
                for (int i = k; i < 6; ++i) tau = rfma(A[i][k], b[i], tau);
                tau = rdiv(tau, A1k);
#pragma unroll
                for (int i = k; i < 6; ++i) b[i] = rfma(-tau, A[i][k], b[i]);
            }
        }
    }
    if (singular) return;
    X[3] = rdiv(b[3], A2[3]);
#pragma unroll
    for (int i = 2; i >= 0; --i) {
        double sum = 0;
#pragma unroll
        for (int j = i + 1; j < 4; ++j) sum = rfma(A[i][j], X[j], sum);
        X[i] = rdiv(b[i] - sum, A2[i]);
    }
}

// PnPsolver::gauss_newton (:675-691) on epnp_gn_step_reg; L and rho may live in shared memory (the loads are kept
// inside the loop: hoisted out of it they would be 66 live doubles)
template <class LT>
__host__ __device__ __forceinline__ void epnp_gauss_newton_reg(LT L, const double* rho, double (&betas)[4])
{
    double X[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
#ifdef __CUDA_ARCH__
        asm volatile("" ::: "memory");
#endif
        epnp_gn_step_reg(L, rho, betas, X);
#pragma unroll
        for (int i = 0; i < 4; ++i) betas[i] += X[i];
    }
}

// L, rho, approx_k + gauss_newton (PnPsolver.cpp:395-405) from the null-space basis U4 (12x4)
template <class LT>
__host__ __device__ inline void epnp_betas_from_basis_L(const double* U4, const double* cws, double* betas /*3x4*/, int ust, LT L);

__host__ __device__ inline void epnp_betas_from_basis(const double* U4, const double* cws, double* betas /*3x4*/, int ust = 1,
                                                      double* l_ext = nullptr, int lst = 1)
{
    if (l_ext) {
        epnp_betas_from_basis_L(U4, cws, betas, ust, StridedD{l_ext, lst});
    } else {
        double Lp[60];
        epnp_betas_from_basis_L(U4, cws, betas, ust, (double*)Lp);
    }
}

template <class LT>
__host__ __device__ inline void epnp_betas_from_basis_L(const double* U4, const double* cws, double* betas /*3x4*/, int ust, LT L)
{
    double rho[6];
    epnp_L_6x10(U4, L, ust);
    epnp_rho(cws, rho);
    RSAC_SOLVE_MARK(3);
    epnp_betas_approx_1(L, rho, betas + 0);
    RSAC_SOLVE_MARK(4);
    epnp_gauss_newton(L, rho, betas + 0);
    RSAC_SOLVE_MARK(5);
    epnp_betas_approx_2(L, rho, betas + 4);
    epnp_gauss_newton(L, rho, betas + 4);
    RSAC_SOLVE_MARK(6);
    epnp_betas_approx_3(L, rho, betas + 8);
    epnp_gauss_newton(L, rho, betas + 8);
    RSAC_SOLVE_MARK(7);
}

// From MtM (PACKED upper triangle, 78 entries, destroyed) to the null-space basis U4 (12x4) and
// the three refined beta vectors: 12x12 eigen-solve (:380), then the above.
// rec: scratch for the recorded rotations (kMaxSweepsRec * 66 double2).
__host__ __device__ inline void epnp_solve_betas(double* MtM, const double* cws, double* U4, double* betas /*3x4*/, double2* rec)
{
    double w4[4];
    jacobi_lowest<12, 4>(MtM, w4, U4, rec);
    epnp_betas_from_basis(U4, cws, betas);
}

// 4-point variant: the null space of the 8 x 12 M directly (Householder QR of M^T) -- for four
// correspondences M^T M is exactly rank 8 and any orthonormal null-space basis is as good as the
// eigen-solver's.  al: 4 x 4 alphas, us: 4 x 2.
__host__ __device__ inline void epnp_solve_betas_qr4(const double* al, const double* us, const Cam& k, const double* cws,
                                                     double* U4, double* betas /*3x4*/, int ust = 1,
                                                     double* l_ext = nullptr, int lst = 1)
{
    double A[96];   // M^T, row-major 12 x 8
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double r0[12], r1[12];
        epnp_m_rows(al + 4 * i, us[2 * i], us[2 * i + 1], k, r0, r1);
#pragma unroll
        for (int r = 0; r < 12; ++r) { A[r * 8 + 2 * i] = r0[r]; A[r * 8 + 2 * i + 1] = r1[r]; }
    }
    nullspace_qr_8x12(A, U4, ust);
    RSAC_SOLVE_MARK(2);
    epnp_betas_from_basis(U4, cws, betas, ust, l_ext, lst);
}

// PnPsolver::compute_ccs (:345-352)
__host__ __device__ inline void epnp_ccs(const double* betas, const double* U4, double* ccs /*4x3*/, int ust = 1)
{
    for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s = rfma(betas[j], U4[((3 * i + c) * 4 + j) * ust], s);
            ccs[i * 3 + c] = s;
        }
}

// one row of pcs = alphas * ccs (:354-357)
__host__ __device__ inline void epnp_pc(const double* a, const double* ccs, double* pc)
{
    for (int c = 0; c < 3; ++c) pc[c] = a[0] * ccs[c] + a[1] * ccs[3 + c] + a[2] * ccs[6 + c] + a[3] * ccs[9 + c];
}

// tail of PnPsolver::estimate_R_and_t (:449-492) given M = sum (pc-pc0)^T (pw-pw0):
// Horn's 4x4 with float-truncated entries, last eigenvector, q = (w,-x,-y,-z)
template <bool STATIC_SORT = false>
__host__ __device__ inline void epnp_horn(const double* M, const double* pc0, const double* pw0, double* R, double* t)
{
    const float N11 = (float)(M[0] + M[4] + M[8]);
    const float N12 = (float)(M[5] - M[7]);
    const float N13 = (float)(M[6] - M[2]);
    const float N14 = (float)(M[1] - M[3]);
    const float N22 = (float)(M[0] - M[4] - M[8]);
    const float N23 = (float)(M[1] + M[3]);
    const float N24 = (float)(M[6] + M[2]);
    const float N33 = (float)(-M[0] + M[4] - M[8]);
    const float N34 = (float)(M[5] + M[7]);
    const float N44 = (float)(-M[0] - M[4] + M[8]);
    double N[16] = {N11, N12, N13, N14, N12, N22, N23, N24, N13, N23, N33, N34, N14, N24, N34, N44};
    double w[4], V[16];
    jacobi_eig<double, 4, STATIC_SORT>(N, w, V);
    quat_to_rot<double>(V[0 * 4 + 3], -V[1 * 4 + 3], -V[2 * 4 + 3], -V[3 * 4 + 3], R);
    if (det3(R) < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int r = 0; r < 3; ++r)
        t[r] = pc0[r] - (R[r * 3 + 0] * pw0[0] + R[r * 3 + 1] * pw0[1] + R[r * 3 + 2] * pw0[2]);
}

// one term of PnPsolver::reprojection_error (:421-428)
__host__ __device__ inline double epnp_reproj_term(const double* R, const double* t, const double* pw, double u, double v, const Cam& k)
{
    const double X = R[0] * pw[0] + R[1] * pw[1] + R[2] * pw[2] + t[0];
    const double Y = R[3] * pw[0] + R[4] * pw[1] + R[5] * pw[2] + t[1];
    const double Z = R[6] * pw[0] + R[7] * pw[1] + R[8] * pw[2] + t[2];
    const double inv_Zc = rdiv(1.0, Z);
    const double ue = k.cx + k.fx * X * inv_Zc;
    const double ve = k.cy + k.fy * Y * inv_Zc;
    const double du = u - ue, dv = v - ve;
    return rsqrt_exact(du * du + dv * dv);
}

// Whole PnPsolver::compute_pose (:359-415) for NPTS thread-private correspondences.
// pw: NPTS x 3, us: NPTS x 2 (already widened to double).  Writes R (9) t (3) as float.
// QR: take the null space of a 4-point system by Householder QR instead of the 12x12 eigen-solve.
// u4_ext/ust: optional external storage for the 12x4 basis (shared memory, element stride ust) -- keeps 48
// long-lived doubles out of the register file / local memory in the one-thread-per-hypothesis kernel
template <int NPTS, bool QR = false>
__host__ __device__ inline double epnp_compute_pose_small(const double* pw, const double* us, const Cam& k, float* Rf, float* tf,
                                                          double* u4_ext = nullptr, int ust = 1, double* l_ext = nullptr, int lst = 1)
{
    RSAC_SOLVE_MARK(0);
    double cws[12], C0[3];
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int i = 0; i < NPTS; ++i) s += pw[i * 3 + c];
        C0[c] = s / (double)NPTS;
    }
    double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < NPTS; ++i) {
        const double d0 = pw[i * 3 + 0] - C0[0], d1 = pw[i * 3 + 1] - C0[1], d2 = pw[i * 3 + 2] - C0[2];
        A[0] = rfma(d0, d0, A[0]); A[1] = rfma(d0, d1, A[1]); A[2] = rfma(d0, d2, A[2]);
        A[4] = rfma(d1, d1, A[4]); A[5] = rfma(d1, d2, A[5]);
        A[8] = rfma(d2, d2, A[8]);
    }
    epnp_control_points(C0, A, NPTS, cws);
    double CCi[9];
    epnp_cc_inverse(cws, CCi);
    double alphas[NPTS * 4];
    for (int i = 0; i < NPTS; ++i) epnp_alphas(pw + 3 * i, cws, CCi, alphas + 4 * i);

    RSAC_SOLVE_MARK(1);
    double U4_priv[48], betas[12];
    double* U4 = u4_ext ? u4_ext : U4_priv;
    if (!u4_ext) ust = 1;
    if constexpr (QR && NPTS == 4) {
        epnp_solve_betas_qr4(alphas, us, k, cws, U4, betas, ust, l_ext, lst);
    } else {
        double MtM[78];   // packed upper triangle
#pragma unroll
        for (int i = 0; i < 78; ++i) MtM[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NPTS; ++i) {
            double r0[12], r1[12];
            epnp_m_rows(alphas + 4 * i, us[2 * i], us[2 * i + 1], k, r0, r1);
#pragma unroll
            for (int a = 0; a < 12; ++a)
#pragma unroll
                for (int b = a; b < 12; ++b) {
                    MtM[tri_idx(12, a, b)] = rfma(r0[a], r0[b], MtM[tri_idx(12, a, b)]);
                    MtM[tri_idx(12, a, b)] = rfma(r1[a], r1[b], MtM[tri_idx(12, a, b)]);
                }
        }
        double2 rec[kMaxSweepsRec * 66];
        epnp_solve_betas(MtM, cws, U4, betas, rec);
    }

    double rep[3], Rs[3][9], ts[3][3];
#pragma unroll 1
    for (int kk = 0; kk < 3; ++kk) {
        double ccs[12], pcs[NPTS * 3];
        epnp_ccs(betas + 4 * kk, U4, ccs, ust);
        for (int i = 0; i < NPTS; ++i) epnp_pc(alphas + 4 * i, ccs, pcs + 3 * i);
        if (pcs[2] < 0.0) {                                   // solve_for_sign (:495-502)
            for (int i = 0; i < 12; ++i) ccs[i] = -ccs[i];
            for (int i = 0; i < NPTS * 3; ++i) pcs[i] = -pcs[i];
        }
        double pc0[3], pw0[3];
        for (int c = 0; c < 3; ++c) {
            double sc = 0.0, sw = 0.0;
            for (int i = 0; i < NPTS; ++i) sc += pcs[i * 3 + c];
            for (int i = 0; i < NPTS; ++i) sw += pw[i * 3 + c];
            pc0[c] = sc / (double)NPTS;
            pw0[c] = sw / (double)NPTS;
        }
        double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < NPTS; ++i)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) M[r * 3 + c] += (pcs[i * 3 + r] - pc0[r]) * (pw[i * 3 + c] - pw0[c]);
        if (kk == 0) RSAC_SOLVE_MARK(8);
        epnp_horn(M, pc0, pw0, Rs[kk], ts[kk]);
        if (kk == 0) RSAC_SOLVE_MARK(9);
        double sum2 = 0.0;
        for (int i = 0; i < NPTS; ++i) sum2 += epnp_reproj_term(Rs[kk], ts[kk], pw + 3 * i, us[2 * i], us[2 * i + 1], k);
        rep[kk] = sum2 / (double)NPTS;
    }
    RSAC_SOLVE_MARK(10);
    int N = 0;                                                // :407-409 (index shifted by one)
    if (rep[1] < rep[0]) N = 1;
    if (rep[2] < rep[N]) N = 2;
    for (int i = 0; i < 9; ++i) Rf[i] = (float)Rs[N][i];
    for (int i = 0; i < 3; ++i) tf[i] = (float)ts[N][i];
    return rep[N];
}

}  // namespace rsac
