/*
 * oracle/orc_pnp.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Plain-C restatement of the reference's src/PnPsolver.cpp (the Eigen rewrite
 * of EPnP + RANSAC).  Function by function; every function cites the
 * reference lines it follows.  Pinned bit for bit against the reference's own source compiled with stand-in
 * Eigen headers (oracle/_ref, tests/test_cpu_reference_build.py); Eigen's own rounding stays unpinned (orc.h).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "orc.h"

extern __thread long long orc_flops;
#define FL(n) (orc_flops += (n))
long long orc_flops_take(void);

/* mirrors the solver state of include/PnPsolver.hpp:71-136 that the arithmetic touches */
typedef struct {
    double fx, fy, cx, cy;
    double cws[4][3], ccs[4][3];
    double *pws, *us, *alphas, *pcs;   /* max_n rows, grow-only (PnPsolver.cpp:271-281) */
    int max_n, n;
    int stale_rows;                    /* Q1 */
    int qr_nullspace;                  /* ORC_FLAG_EPNP_QR_NULLSPACE: n == 4 uses orc_nullspace_qr_d */
    const double *ext_basis;           /* test hook: null-space basis supplied by the caller (12 x 4, row-major) */
    double *mtm_out;                   /* test hook: receives the full symmetric 12 x 12 M^T M */
} epnp_t;

static void epnp_init(epnp_t *e, double fx, double fy, double cx, double cy, int stale_rows)
{
    memset(e, 0, sizeof(*e));
    e->fx = fx; e->fy = fy; e->cx = cx; e->cy = cy;
    e->stale_rows = stale_rows;
}

static void epnp_free(epnp_t *e)
{
    free(e->pws); free(e->us); free(e->alphas); free(e->pcs);
    e->pws = e->us = e->alphas = e->pcs = NULL;
}

/* PnPsolver::set_maximum_number_of_correspondences (PnPsolver.cpp:271-281):
 * grow-only; setZero(rows, cols) resizes AND zeroes everything. */
static void set_maximum_number_of_correspondences(epnp_t *e, int n)
{
    if (e->max_n < n) {
        e->max_n = n;
        free(e->pws); free(e->us); free(e->alphas); free(e->pcs);
        e->pws = (double *)calloc((size_t)n * 3, sizeof(double));
        e->us = (double *)calloc((size_t)n * 2, sizeof(double));
        e->alphas = (double *)calloc((size_t)n * 4, sizeof(double));
        e->pcs = (double *)calloc((size_t)n * 3, sizeof(double));
    }
}

/* PnPsolver.cpp:283-286 */
static void reset_correspondences(epnp_t *e) { e->n = 0; }

/* PnPsolver.cpp:288-294: f32 -> f64 widening */
static void add_correspondence(epnp_t *e, const float *p3d, const float *p2d)
{
    e->pws[e->n * 3 + 0] = (double)p3d[0];
    e->pws[e->n * 3 + 1] = (double)p3d[1];
    e->pws[e->n * 3 + 2] = (double)p3d[2];
    e->us[e->n * 2 + 0] = (double)p2d[0];
    e->us[e->n * 2 + 1] = (double)p2d[1];
    e->n++;
}

/* rows entering the colwise().sum() / alphas*ccs expressions: all ALLOCATED
 * rows in the as-shipped reference (Q1), number_of_correspondences in the clean
 * semantics (upstream _PnPsolver.cpp:334,423,531) */
static int sum_rows(const epnp_t *e) { return e->stale_rows ? e->max_n : e->n; }

/* PnPsolver::choose_control_points (PnPsolver.cpp:296-321) */
static void choose_control_points(epnp_t *e)
{
    const int n = e->n, ns = sum_rows(e);
    memset(e->cws, 0, sizeof(e->cws));
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int i = 0; i < ns; ++i) s += e->pws[i * 3 + c];   /* :301 colwise().sum() */
        e->cws[0][c] = s / (double)n;                          /* :303 */
    }
    /* :306-310 PW0tPW0 = PW0^T PW0 (symmetric, upper triangle computed) */
    double A[9];
    memset(A, 0, sizeof(A));
    for (int i = 0; i < n; ++i) {
        const double d0 = e->pws[i * 3 + 0] - e->cws[0][0];
        const double d1 = e->pws[i * 3 + 1] - e->cws[0][1];
        const double d2 = e->pws[i * 3 + 2] - e->cws[0][2];
        A[0] = fma(d0, d0, A[0]); A[1] = fma(d0, d1, A[1]); A[2] = fma(d0, d2, A[2]);
        A[4] = fma(d1, d1, A[4]); A[5] = fma(d1, d2, A[5]);
        A[8] = fma(d2, d2, A[8]);
    }
    FL(3 * ns + 3 + n * 15);
    double DC[3], UCt[9];
    orc_jacobi_eig_d(3, A, DC, UCt);                           /* :311 */
    FL(3 * (2 + 6));
    for (int i = 0; i < 3; ++i) {
        const double k = sqrt(DC[i] / (double)n);              /* :318 (negative => NaN) */
        for (int c = 0; c < 3; ++c) e->cws[i + 1][c] = e->cws[0][c] + k * UCt[c * 3 + i];
    }
}

/* PnPsolver::compute_barycentric_coordinates (PnPsolver.cpp:323-343) */
static void compute_barycentric_coordinates(epnp_t *e)
{
    double CC[9], CCi[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 1; j < 4; ++j) CC[i * 3 + (j - 1)] = e->cws[j][i] - e->cws[0][i];
    orc_inv3_d(CC, CCi);                                       /* :331 */
    FL(9 + e->n * (3 + 15 + 3));
    for (int i = 0; i < e->n; ++i) {
        const double d0 = e->pws[i * 3 + 0] - e->cws[0][0];
        const double d1 = e->pws[i * 3 + 1] - e->cws[0][1];
        const double d2 = e->pws[i * 3 + 2] - e->cws[0][2];
        double *a = &e->alphas[i * 4];
        for (int j = 0; j < 3; ++j)
            a[j + 1] = CCi[j * 3 + 0] * d0 + CCi[j * 3 + 1] * d1 + CCi[j * 3 + 2] * d2;   /* :338 */
        a[0] = 1.0 - a[1] - a[2] - a[3];                                                 /* :340 */
    }
}

/* PnPsolver::compute_L_6x10 (PnPsolver.cpp:604-637); U is 12x4 row-major: the 4 smallest eigenvectors (:382 uses columns 0..3 only) */
static void compute_L_6x10(const double *U, double L[6][10])
{
    double dv[4][6][3];
    for (int i = 0; i < 4; ++i) {
        int a = 0, b = 1;
        for (int j = 0; j < 6; ++j) {
            for (int c = 0; c < 3; ++c) dv[i][j][c] = U[(3 * a + c) * 4 + i] - U[(3 * b + c) * 4 + i];
            b++;
            if (b > 3) { a++; b = a + 1; }
        }
    }
#define DOT3(x, y) ((x)[0] * (y)[0] + (x)[1] * (y)[1] + (x)[2] * (y)[2])
    for (int i = 0; i < 6; ++i) {
        L[i][0] = DOT3(dv[0][i], dv[0][i]);
        L[i][1] = 2.0 * DOT3(dv[0][i], dv[1][i]);
        L[i][2] = DOT3(dv[1][i], dv[1][i]);
        L[i][3] = 2.0 * DOT3(dv[0][i], dv[2][i]);
        L[i][4] = 2.0 * DOT3(dv[1][i], dv[2][i]);
        L[i][5] = DOT3(dv[2][i], dv[2][i]);
        L[i][6] = 2.0 * DOT3(dv[0][i], dv[3][i]);
        L[i][7] = 2.0 * DOT3(dv[1][i], dv[3][i]);
        L[i][8] = 2.0 * DOT3(dv[2][i], dv[3][i]);
        L[i][9] = DOT3(dv[3][i], dv[3][i]);
    }
#undef DOT3
}

static double sqdist3(const double *a, const double *b)
{
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

/* PnPsolver::compute_rho (PnPsolver.cpp:639-647) */
static void compute_rho(const epnp_t *e, double rho[6])
{
    rho[0] = sqdist3(e->cws[0], e->cws[1]);
    rho[1] = sqdist3(e->cws[0], e->cws[2]);
    rho[2] = sqdist3(e->cws[0], e->cws[3]);
    rho[3] = sqdist3(e->cws[1], e->cws[2]);
    rho[4] = sqdist3(e->cws[1], e->cws[3]);
    rho[5] = sqdist3(e->cws[2], e->cws[3]);
}

/* PnPsolver::find_betas_approx_1 (PnPsolver.cpp:520-544): columns {0,1,3,6} */
static void find_betas_approx_1(const double L[6][10], const double rho[6], double betas[4])
{
    double L4[6 * 4], b4[4];
    for (int i = 0; i < 6; ++i) {
        L4[i * 4 + 0] = L[i][0]; L4[i * 4 + 1] = L[i][1]; L4[i * 4 + 2] = L[i][3]; L4[i * 4 + 3] = L[i][6];
    }
    orc_lstsq_d(6, 4, L4, rho, b4);
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

/* PnPsolver::find_betas_approx_2 (PnPsolver.cpp:549-573): columns {0,1,2} */
static void find_betas_approx_2(const double L[6][10], const double rho[6], double betas[4])
{
    double L3[6 * 3], b3[3];
    for (int i = 0; i < 6; ++i) {
        L3[i * 3 + 0] = L[i][0]; L3[i * 3 + 1] = L[i][1]; L3[i * 3 + 2] = L[i][2];
    }
    orc_lstsq_d(6, 3, L3, rho, b3);
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

/* PnPsolver::find_betas_approx_3 (PnPsolver.cpp:578-602): columns {0..4} */
static void find_betas_approx_3(const double L[6][10], const double rho[6], double betas[4])
{
    double L5[6 * 5], b5[5];
    for (int i = 0; i < 6; ++i)
        for (int c = 0; c < 5; ++c) L5[i * 5 + c] = L[i][c];
    orc_lstsq_d(6, 5, L5, rho, b5);
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

/* PnPsolver::compute_A_and_b_gauss_newton (PnPsolver.cpp:649-673) */
static void compute_A_and_b_gauss_newton(const double L[6][10], const double rho[6], const double bt[4],
                                         double A[6][4], double b[6])
{
    for (int i = 0; i < 6; ++i) {
        const double *l = L[i];
        const double Lt[4][4] = {{2 * l[0], l[1], l[3], l[6]},
                                 {l[1], 2 * l[2], l[4], l[7]},
                                 {l[3], l[4], 2 * l[5], l[8]},
                                 {l[6], l[7], l[8], 2 * l[9]}};
        for (int r = 0; r < 4; ++r)
            A[i][r] = fma(Lt[r][3], bt[3], fma(Lt[r][2], bt[2], fma(Lt[r][1], bt[1], Lt[r][0] * bt[0])));   /* :659 */
        double q = (l[0] * bt[0]) * bt[0];                                                          /* :661-671 */
        q = fma(l[1] * bt[0], bt[1], q);
        q = fma(l[2] * bt[1], bt[1], q);
        q = fma(l[3] * bt[0], bt[2], q);
        q = fma(l[4] * bt[1], bt[2], q);
        q = fma(l[5] * bt[2], bt[2], q);
        q = fma(l[6] * bt[0], bt[3], q);
        q = fma(l[7] * bt[1], bt[3], q);
        q = fma(l[8] * bt[2], bt[3], q);
        q = fma(l[9] * bt[3], bt[3], q);
        b[i] = rho[i] - q;
    }
}

/* PnPsolver::qr_solve (PnPsolver.cpp:693-796): in-place Householder QR of the
 * 6x4 A with max-abs column scaling, b <- Q^T b, back substitution.  The
 * reference's function-static scratch A1/A2 (:696-697, Q8) is stack scratch
 * here.  eta == 0 => return with X untouched (:722-727; the cerr message is
 * not reproduced). */
static void qr_solve(double A[6][4], double b[6], double X[4])
{
    enum { nr = 6, nc = 4 };
    double A1[nc], A2[nc];
    for (int k = 0; k < nc; ++k) {
        double eta = fabs(A[k][k]);
        for (int i = k + 1; i < nr; ++i) {
            const double elt = fabs(A[i][k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) {
            A1[k] = A2[k] = 0.0;
            return;
        }
        const double inv_eta = 1. / eta;
        double sum = 0.0;
        for (int i = k; i < nr; ++i) {
            A[i][k] *= inv_eta;
            sum = fma(A[i][k], A[i][k], sum);
        }
        double sigma = sqrt(sum);
        if (A[k][k] < 0) sigma = -sigma;
        A[k][k] += sigma;
        A1[k] = sigma * A[k][k];
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; ++j) {
            double s = 0;
            for (int i = k; i < nr; ++i) s = fma(A[i][k], A[i][j], s);
            const double tau = s / A1[k];
            for (int i = k; i < nr; ++i) A[i][j] = fma(-tau, A[i][k], A[i][j]);
        }
    }
    FL(4 * 6 + 2 * (6 + 5 + 4 + 3) + 4 * 5 + (3 * 6 + 2 * 5 + 1 * 4) * 4 + 6);   /* column scaling, norms, reflections */
    FL(2 * (6 + 5 + 4 + 3) * 2 + 4 + 4 + 12);                                      /* Qt b, back substitution */
    /* b <- Qt b (:762-780) */
    for (int j = 0; j < nc; ++j) {
        double tau = 0;
        for (int i = j; i < nr; ++i) tau = fma(A[i][j], b[i], tau);
        tau /= A1[j];
        for (int i = j; i < nr; ++i) b[i] = fma(-tau, A[i][j], b[i]);
    }
    /* X = R^-1 b (:782-795) */
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; --i) {
        double sum = 0;
        for (int j = i + 1; j < nc; ++j) sum = fma(A[i][j], X[j], sum);
        X[i] = (b[i] - sum) / A2[i];
    }
}

/* PnPsolver::gauss_newton (PnPsolver.cpp:675-691): exactly 5 iterations.  X is
 * uninitialised in the reference (:682, read only after the singular early
 * return); zero here. */
static void gauss_newton(const double L[6][10], const double rho[6], double betas[4])
{
    double A[6][4], B[6], X[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < 5; ++k) {
        compute_A_and_b_gauss_newton(L, rho, betas, A, B);
        qr_solve(A, B, X);
        for (int i = 0; i < 4; ++i) betas[i] += X[i];
    }
}

/* PnPsolver::compute_ccs (PnPsolver.cpp:345-352) */
static void compute_ccs(epnp_t *e, const double betas[4], const double *U)
{
    for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s = fma(betas[j], U[(3 * i + c) * 4 + j], s);
            e->ccs[i][c] = s;
        }
}

/* PnPsolver::compute_pcs (PnPsolver.cpp:354-357): pcs = alphas*ccs over all allocated rows (Q1) */
static void compute_pcs(epnp_t *e)
{
    const int ns = sum_rows(e);
    for (int i = 0; i < ns; ++i) {
        const double *a = &e->alphas[i * 4];
        for (int c = 0; c < 3; ++c)
            e->pcs[i * 3 + c] = a[0] * e->ccs[0][c] + a[1] * e->ccs[1][c] + a[2] * e->ccs[2][c] + a[3] * e->ccs[3][c];
    }
}

/* PnPsolver::solve_for_sign (PnPsolver.cpp:495-502) */
static void solve_for_sign(epnp_t *e)
{
    if (e->pcs[2] < 0.0) {
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 3; ++c) e->ccs[i][c] = -e->ccs[i][c];
        const int ns = sum_rows(e);
        for (int i = 0; i < ns * 3; ++i) e->pcs[i] = -e->pcs[i];
    }
}

/* Eigen::Quaternion::toRotationMatrix (PnPsolver.cpp:478), no normalisation */
static void quat_to_rot_d(double w, double x, double y, double z, double R[9])
{
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

/* PnPsolver::estimate_R_and_t (PnPsolver.cpp:433-493): Horn's quaternion method
 * with the N entries held in float temporaries (:449-462, Q5) */
static void estimate_R_and_t(epnp_t *e, double R[9], double t[3])
{
    const int n = e->n, ns = sum_rows(e);
    double pc0[3], pw0[3];
    for (int c = 0; c < 3; ++c) {
        double sc = 0.0, sw = 0.0;
        for (int i = 0; i < ns; ++i) sc += e->pcs[i * 3 + c];
        for (int i = 0; i < ns; ++i) sw += e->pws[i * 3 + c];
        pc0[c] = sc / (double)n;
        pw0[c] = sw / (double)n;
    }
    double M[9];
    memset(M, 0, sizeof(M));
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                M[r * 3 + c] += (e->pcs[i * 3 + r] - pc0[r]) * (e->pws[i * 3 + c] - pw0[c]);   /* :445 */

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
    orc_jacobi_eig_d(4, N, w, V);                                       /* :469 */
    /* :471-476 last column, q = (w, -x, -y, -z) */
    quat_to_rot_d(V[0 * 4 + 3], -V[1 * 4 + 3], -V[2 * 4 + 3], -V[3 * 4 + 3], R);
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                       R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }         /* :489-491 */
    for (int r = 0; r < 3; ++r)
        t[r] = pc0[r] - (R[r * 3 + 0] * pw0[0] + R[r * 3 + 1] * pw0[1] + R[r * 3 + 2] * pw0[2]);   /* :492 */
}

/* PnPsolver::reprojection_error (PnPsolver.cpp:417-431): mean L2 norm (not squared) */
static double reprojection_error(const epnp_t *e, const double R[9], const double t[3])
{
    double sum2 = 0.0;
    for (int i = 0; i < e->n; ++i) {
        const double *pw = &e->pws[i * 3];
        const double X = R[0] * pw[0] + R[1] * pw[1] + R[2] * pw[2] + t[0];
        const double Y = R[3] * pw[0] + R[4] * pw[1] + R[5] * pw[2] + t[1];
        const double Z = R[6] * pw[0] + R[7] * pw[1] + R[8] * pw[2] + t[2];
        const double inv_Zc = 1.0 / Z;
        const double ue = e->cx + e->fx * X * inv_Zc;
        const double ve = e->cy + e->fy * Y * inv_Zc;
        const double du = e->us[i * 2 + 0] - ue, dv = e->us[i * 2 + 1] - ve;
        sum2 += sqrt(du * du + dv * dv);
    }
    return sum2 / (double)e->n;
}

/* PnPsolver::compute_R_and_t (PnPsolver.cpp:504-515) */
static double compute_R_and_t(epnp_t *e, const double *U, const double betas[4], double R[9], double t[3])
{
    compute_ccs(e, betas, U);
    compute_pcs(e);
    solve_for_sign(e);
    estimate_R_and_t(e, R, t);
    return reprojection_error(e, R, t);
}

/* PnPsolver::compute_pose (PnPsolver.cpp:359-415) */
static double compute_pose(epnp_t *e, float Rf[9], float tf[3])
{
    choose_control_points(e);
    compute_barycentric_coordinates(e);

    /* :365-379  M (2n x 12) and MtM = M^T M, accumulated row by row, upper triangle */
    const int use_qr = e->qr_nullspace && e->n == 4;
    double MtM[144];
    memset(MtM, 0, sizeof(MtM));
    for (int i = 0; i < e->n && !use_qr; ++i) {
        double r0[12], r1[12];
        for (int j = 0; j < 4; ++j) {
            const double a = e->alphas[i * 4 + j];
            r0[3 * j] = a * e->fx; r0[3 * j + 1] = 0.0;      r0[3 * j + 2] = a * (e->cx - e->us[i * 2 + 0]);
            r1[3 * j] = 0.0;       r1[3 * j + 1] = a * e->fy; r1[3 * j + 2] = a * (e->cy - e->us[i * 2 + 1]);
        }
        for (int a = 0; a < 12; ++a)
            for (int b = a; b < 12; ++b) {
                MtM[a * 12 + b] = fma(r0[a], r0[b], MtM[a * 12 + b]);
                MtM[a * 12 + b] = fma(r1[a], r1[b], MtM[a * 12 + b]);
            }
    }
    double w[4], U[48];
    if (e->mtm_out)
        for (int a = 0; a < 12; ++a)
            for (int b = a; b < 12; ++b) e->mtm_out[a * 12 + b] = e->mtm_out[b * 12 + a] = MtM[a * 12 + b];
    if (e->ext_basis) {
        /* basis-sensitivity experiments (tests/test_cpu_basis_agreement.py): everything downstream of :380-382
         * runs on a caller-supplied orthonormal basis of the (near-)null space, e.g. LAPACK's */
        memcpy(U, e->ext_basis, sizeof(U));
    } else if (use_qr) {
        /* the same 8 x 12 M, null space taken directly (see orc_nullspace_qr_d) */
        double Mrows[8 * 12];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                const double a = e->alphas[i * 4 + j];
                double *r0 = &Mrows[(2 * i) * 12], *r1 = &Mrows[(2 * i + 1) * 12];
                r0[3 * j] = a * e->fx; r0[3 * j + 1] = 0.0;      r0[3 * j + 2] = a * (e->cx - e->us[i * 2 + 0]);
                r1[3 * j] = 0.0;       r1[3 * j + 1] = a * e->fy; r1[3 * j + 2] = a * (e->cy - e->us[i * 2 + 1]);
            }
        FL(4 * 4 * 6);
        orc_nullspace_qr_d(Mrows, U);
    } else {
        FL(e->n * (4 * 6 + 2 + 78 * 4));
        orc_jacobi_lowest_d(12, 4, MtM, w, U);                         /* :380-382: eigenvectors 0..3 */
    }
    FL(4 * 6 * 3 + 6 * (10 * 5 + 6) + 6 * 8);   /* L_6x10, rho */
    FL(3 * (8 + 5 * (6 * (16 + 4 * 7 + 20 + 1) + 4)));   /* betas post-processing, 5 x (A,b build + beta update); QR counted below */

    double L[6][10], rho[6];
    compute_L_6x10(U, L);
    compute_rho(e, rho);

    double Betas[4][4], rep[4], Rs[4][9], ts[4][3];
    memset(Betas, 0, sizeof(Betas));
    find_betas_approx_1(L, rho, Betas[1]);
    gauss_newton(L, rho, Betas[1]);
    rep[1] = compute_R_and_t(e, U, Betas[1], Rs[1], ts[1]);

    find_betas_approx_2(L, rho, Betas[2]);
    gauss_newton(L, rho, Betas[2]);
    rep[2] = compute_R_and_t(e, U, Betas[2], Rs[2], ts[2]);

    find_betas_approx_3(L, rho, Betas[3]);
    gauss_newton(L, rho, Betas[3]);
    rep[3] = compute_R_and_t(e, U, Betas[3], Rs[3], ts[3]);

    int N = 1;                                                         /* :407-409 */
    if (rep[2] < rep[1]) N = 2;
    if (rep[3] < rep[N]) N = 3;
    for (int i = 0; i < 9; ++i) Rf[i] = (float)Rs[N][i];               /* :411-412 */
    for (int i = 0; i < 3; ++i) tf[i] = (float)ts[N][i];
    return rep[N];
}

/* PnPsolver::SetRansacParameters (PnPsolver.cpp:58-94), MLPnPsolver.cpp:185-220 */
void orc_pnp_ransac_setup(int N, const orc_ransac_params *p, int *min_inl, int *max_its)
{
    float eps = p->eps;
    int nMinInliers = (int)((float)N * eps);        /* :71 int = int*float */
    if (nMinInliers < p->min_inliers) nMinInliers = p->min_inliers;
    if (nMinInliers < p->min_set) nMinInliers = p->min_set;
    if (eps < (float)nMinInliers / N) eps = (float)nMinInliers / N;
    int nIterations;
    if (nMinInliers == N)
        nIterations = 1;
    else
        nIterations = (int)ceil(log(1 - p->prob) / log(1 - pow(eps, 3)));   /* :87 exponent 3 for every minSet (Q10) */
    int its = nIterations < p->max_its ? nIterations : p->max_its;
    *max_its = its > 1 ? its : 1;
    *min_inl = nMinInliers;
}

/* PnPsolver::CheckInliers (PnPsolver.cpp:241-268).  Mixed precision (Q5): f32
 * rigid transform, f32 reciprocal, f64 projection narrowed to f32, f32 squared
 * norm.  No positive-depth test; NaN => outlier. */
int orc_pnp_check_inliers(const orc_pnp_problem *pb, const float *max_err, const float R[9],
                          const float t[3], uint8_t *mask, float *err2)
{
    int cnt = 0;
    for (int i = 0; i < pb->n; ++i) {
        const float X = pb->p3d[i * 3 + 0], Y = pb->p3d[i * 3 + 1], Z = pb->p3d[i * 3 + 2];
        const float xc = (R[0] * X + R[1] * Y + R[2] * Z) + t[0];      /* :250 */
        const float yc = (R[3] * X + R[4] * Y + R[5] * Z) + t[1];
        const float zc = (R[6] * X + R[7] * Y + R[8] * Z) + t[2];
        const float invZc = 1 / zc;                                    /* :252 */
        const float ue = (float)(pb->cx + pb->fx * (double)xc * (double)invZc);   /* :254 */
        const float ve = (float)(pb->cy + pb->fy * (double)yc * (double)invZc);
        const float du = ue - pb->p2d[i * 2 + 0], dv = ve - pb->p2d[i * 2 + 1];
        const float error2 = du * du + dv * dv;                        /* :256 */
        const int in = error2 < max_err[i];                            /* :258 */
        if (mask) mask[i] = (uint8_t)in;
        if (err2) err2[i] = error2;
        cnt += in;
    }
    return cnt;
}

void orc_pnp_score(const orc_pnp_problem *pb, const float *max_err, int H, const float *poses,
                   uint8_t *masks, int *counts)
{
    for (int h = 0; h < H; ++h)
        counts[h] = orc_pnp_check_inliers(pb, max_err, poses + (size_t)h * 12, poses + (size_t)h * 12 + 9,
                                          masks ? masks + (size_t)h * pb->n : NULL, NULL);
}

/* average algorithmic FP64 FLOP of one minimal (min_set-point) EPnP solve over the H rows of a table */
double orc_epnp_flops_mode(const orc_pnp_problem *pb, const uint32_t *table, int H, int min_set, int flags)
{
    float R[9], t[3];
    orc_flops_take();
    for (int h = 0; h < H; ++h) orc_epnp_pose_mode(pb, table + (size_t)h * min_set, min_set, flags, R, t);
    return (double)orc_flops_take() / (double)(H > 0 ? H : 1);
}

double orc_epnp_flops(const orc_pnp_problem *pb, const uint32_t *table, int H, int min_set)
{
    return orc_epnp_flops_mode(pb, table, H, min_set, 0);
}

double orc_epnp_pose(const orc_pnp_problem *pb, const uint32_t *idx, int m, float R[9], float t[3])
{
    return orc_epnp_pose_mode(pb, idx, m, 0, R, t);
}

double orc_epnp_pose_mode(const orc_pnp_problem *pb, const uint32_t *idx, int m, int flags, float R[9], float t[3])
{
    epnp_t e;
    epnp_init(&e, pb->fx, pb->fy, pb->cx, pb->cy, 0);
    e.qr_nullspace = (flags & ORC_FLAG_EPNP_QR_NULLSPACE) != 0;
    set_maximum_number_of_correspondences(&e, m);
    reset_correspondences(&e);
    for (int i = 0; i < m; ++i) add_correspondence(&e, pb->p3d + 3 * idx[i], pb->p2d + 2 * idx[i]);
    const double err = compute_pose(&e, R, t);
    epnp_free(&e);
    return err;
}

/* test hooks for the null-space-basis experiments: M^T M of the subset (12 x 12, symmetric, row-major) and the
 * solve of PnPsolver::compute_pose (:383-415) downstream of a caller-supplied basis U (12 x 4, column j = vector j) */
void orc_epnp_mtm(const orc_pnp_problem *pb, const uint32_t *idx, int m, double MtM[144])
{
    epnp_t e;
    float R[9], t[3];
    epnp_init(&e, pb->fx, pb->fy, pb->cx, pb->cy, 0);
    e.mtm_out = MtM;
    set_maximum_number_of_correspondences(&e, m);
    reset_correspondences(&e);
    for (int i = 0; i < m; ++i) add_correspondence(&e, pb->p3d + 3 * idx[i], pb->p2d + 2 * idx[i]);
    compute_pose(&e, R, t);
    epnp_free(&e);
}

double orc_epnp_pose_basis(const orc_pnp_problem *pb, const uint32_t *idx, int m, const double U[48], float R[9], float t[3])
{
    epnp_t e;
    epnp_init(&e, pb->fx, pb->fy, pb->cx, pb->cy, 0);
    e.ext_basis = U;
    set_maximum_number_of_correspondences(&e, m);
    reset_correspondences(&e);
    for (int i = 0; i < m; ++i) add_correspondence(&e, pb->p3d + 3 * idx[i], pb->p2d + 2 * idx[i]);
    const double err = compute_pose(&e, R, t);
    epnp_free(&e);
    return err;
}

static void set_T(float T[16], const float R[9], const float t[3])
{
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c];
        T[r * 4 + 3] = t[r];
    }
}

/* PnPsolver::iterate (PnPsolver.cpp:102-191) on its first call + Refine (:193-238).
 * The first call runs until mnIterations >= mRansacMaxIts because of the `||`
 * at :119 (Q2), so one call covers the whole budget. */
void orc_pnp_ransac(const orc_pnp_problem *pb, const orc_ransac_params *prm, const uint32_t *table,
                    int flags, orc_result *res, uint8_t *mask, int *hyp_counts, float *hyp_pose)
{
    const int N = pb->n;
    int minInl, H;
    orc_pnp_ransac_setup(N, prm, &minInl, &H);
    const int minSet = prm->min_set;
    const int exhaustive = (flags & ORC_FLAG_EXHAUSTIVE) != 0;

    memset(res, 0, sizeof(*res));
    res->best_hyp = -1;
    for (int i = 0; i < 16; ++i) res->T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    res->scale = 1.0f;
    if (mask) memset(mask, 0, (size_t)N);

    if (N < minInl) {            /* :110-114 */
        res->no_more = 1;
        return;
    }

    float *maxErr = (float *)malloc(sizeof(float) * (size_t)N);
    for (int i = 0; i < N; ++i) maxErr[i] = pb->sigma2[i] * prm->th2;   /* :93 */
    uint8_t *cur = (uint8_t *)malloc((size_t)N), *best = (uint8_t *)calloc((size_t)N, 1),
            *ref = (uint8_t *)malloc((size_t)N);
    int nBest = 0;
    float Rb[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tb[3] = {0, 0, 0};
    int done = 0;

    epnp_t e;
    epnp_init(&e, pb->fx, pb->fy, pb->cx, pb->cy, (flags & ORC_FLAG_STALE_ROWS) != 0);
    e.qr_nullspace = (flags & ORC_FLAG_EPNP_QR_NULLSPACE) != 0;
    set_maximum_number_of_correspondences(&e, minSet);                  /* :108 */

    for (int h = 0; h < H; ++h) {                                       /* :119 */
        float Ri[9], ti[3];
        reset_correspondences(&e);
        for (int i = 0; i < minSet; ++i) {                              /* :128-138 */
            const uint32_t idx = table[(size_t)h * minSet + i];
            add_correspondence(&e, pb->p3d + 3 * idx, pb->p2d + 2 * idx);
        }
        compute_pose(&e, Ri, ti);                                       /* :141 */
        const int cnt = orc_pnp_check_inliers(pb, maxErr, Ri, ti, cur, NULL);   /* :144 */
        res->n_hyp = h + 1;
        if (hyp_counts) hyp_counts[h] = cnt;
        if (hyp_pose) { memcpy(hyp_pose + (size_t)h * 12, Ri, sizeof(Ri)); memcpy(hyp_pose + (size_t)h * 12 + 9, ti, sizeof(ti)); }
        if (done) continue;   /* exhaustive mode: keep evaluating for the per-hypothesis outputs only */

        if (cnt >= minInl) {                                            /* :146 */
            if (cnt > nBest) {                                          /* :149 strict: first max wins */
                memcpy(best, cur, (size_t)N);
                nBest = cnt;
                memcpy(Rb, Ri, sizeof(Rb)); memcpy(tb, ti, sizeof(tb));
                res->best_hyp = h;
            }
            /* Refine() (:193-238): EPnP on all inliers of the BEST set so far */
            int m = 0;
            for (int i = 0; i < N; ++i) m += best[i];
            set_maximum_number_of_correspondences(&e, m);               /* :206 */
            reset_correspondences(&e);
            for (int i = 0; i < N; ++i)
                if (best[i]) add_correspondence(&e, pb->p3d + 3 * i, pb->p2d + 2 * i);
            float Rr[9], tr[3];
            compute_pose(&e, Rr, tr);                                   /* :217 */
            const int cr = orc_pnp_check_inliers(pb, maxErr, Rr, tr, ref, NULL);
            res->n_refines++;
            if (cr > minInl) {                                          /* :225 strict */
                res->ok = 1;
                res->refined = 1;
                res->n_inliers = cr;
                set_T(res->T, Rr, tr);
                if (mask) memcpy(mask, ref, (size_t)N);
                done = 1;
                if (!exhaustive) break;                                 /* :167 return true */
            } else {
                res->n_failed_refines++;
            }
        }
    }

    res->best_count = nBest;
    if (!done) {
        res->no_more = 1;                                               /* :173-175 */
        if (nBest >= minInl) {                                          /* :176-187 unrefined best */
            res->ok = 1;
            res->n_inliers = nBest;
            set_T(res->T, Rb, tb);
            if (mask) memcpy(mask, best, (size_t)N);
        }
    }
    epnp_free(&e);
    free(maxErr); free(cur); free(best); free(ref);
}
