/*
 * oracle/orc_mlpnp.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Plain-C restatement of the reference's src/MLPnPsolver.cpp (an ORB-SLAM3
 * transplant that the reference neither compiles nor calls, SURVEY F6: this
 * file is spec-by-source).  Pinned (round 2) against the reference's own
 * MLPnPsolver.cpp compiled with stand-in Eigen / OpenCV headers (oracle/_ref, tests/test_cpu_reference_build.py:
 * generated Jacobian 1e-12, computePose median 3e-12, whole runs exact in counts and stopping iteration with Q6
 * reproduced); Eigen's own SVD / LDLT rounding stays unpinned (orc.h).
 *
 * The 2x6 Jacobian (MLPnPsolver.cpp:773-1020) is a machine-generated symbolic
 * derivative of  r = N^T (R(w) p + t)/|R(w) p + t|  w.r.t. (w, t); it is
 * restated here by differentiating the Rodrigues formula directly (same
 * function, same singularity at w = 0), not by transcribing its temporaries.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* MLPnPsolver::rodrigues2rot (MLPnPsolver.cpp:625-640) */
void orc_rodrigues2rot(const double w[3], double R[9])
{
    const double K[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (th > DBL_EPSILON) {
        const double a = sin(th) / th;
        const double b = (1 - cos(th)) / (th * th);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const double k2 = K[r * 3 + 0] * K[0 * 3 + c] + K[r * 3 + 1] * K[1 * 3 + c] + K[r * 3 + 2] * K[2 * 3 + c];
                R[r * 3 + c] = R[r * 3 + c] + a * K[r * 3 + c] + b * k2;
            }
    }
}

/* MLPnPsolver::rot2rodrigues (MLPnPsolver.cpp:642-657) */
void orc_rot2rodrigues(const double R[9], double w[3])
{
    w[0] = w[1] = w[2] = 0.0;
    const double trace = (R[0] + R[4] + R[8]) - 1.0;
    const double wnorm = acos(trace / 2.0);
    if (wnorm > DBL_EPSILON) {
        const double sc = wnorm / (2.0 * sin(wnorm));
        w[0] = (R[7] - R[5]) * sc;
        w[1] = (R[2] - R[6]) * sc;
        w[2] = (R[3] - R[1]) * sc;
    }
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = fma(a[1], b[2], -(a[2] * b[1]));
    o[1] = fma(a[2], b[0], -(a[0] * b[2]));
    o[2] = fma(a[0], b[1], -(a[1] * b[0]));
}

/* residual pair and 2x6 Jacobian of one observation
 * (MLPnPsolver::mlpnp_residuals_and_jacs :736-742 and mlpnpJacs :773-1020) */
void orc_mlpnp_res_jac(const double p[3], const double nr[3], const double ns[3],
                       const double w[3], const double t[3], double r[2], double J[12])
{
    double R[9];
    orc_rodrigues2rot(w, R);
    double q[3];
    for (int i = 0; i < 3; ++i) q[i] = (R[i * 3 + 0] * p[0] + R[i * 3 + 1] * p[1] + R[i * 3 + 2] * p[2]) + t[i];
    const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double qh[3] = {q[0] / qn, q[1] / qn, q[2] / qn};
    r[0] = nr[0] * qh[0] + nr[1] * qh[1] + nr[2] * qh[2];
    r[1] = ns[0] * qh[0] + ns[1] * qh[1] + ns[2] * qh[2];

    /* g_k = (n_k - (n_k . qh) qh)/|q|  = d r_k / d q */
    double g[2][3];
    for (int i = 0; i < 3; ++i) {
        g[0][i] = (nr[i] - r[0] * qh[i]) / qn;
        g[1][i] = (ns[i] - r[1] * qh[i]) / qn;
    }
    /* d(R p)/d w_j from R = I + a K + b K^2, a = sin(th)/th, b = (1-cos th)/th^2 */
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    const double sn = sin(th), cs = cos(th);
    const double a = sn / th;
    const double b = (1.0 - cs) / th2;
    const double da = (th * cs - sn) / (th2 * th);                    /* (1/th) da/dth */
    const double db = (th * sn - 2.0 * (1.0 - cs)) / (th2 * th2);     /* (1/th) db/dth */
    double wxp[3], wxwxp[3];
    cross3(w, p, wxp);
    cross3(w, wxp, wxwxp);
    for (int j = 0; j < 3; ++j) {
        double e[3] = {0.0, 0.0, 0.0};
        e[j] = 1.0;
        double exp_[3], exwxp[3], wxexp[3];
        cross3(e, p, exp_);
        cross3(e, wxp, exwxp);
        cross3(w, exp_, wxexp);
        double d[3];
        for (int i = 0; i < 3; ++i)
            d[i] = (da * w[j]) * wxp[i] + a * exp_[i] + (db * w[j]) * wxwxp[i] + b * (exwxp[i] + wxexp[i]);
        J[0 * 6 + j] = g[0][0] * d[0] + g[0][1] * d[1] + g[0][2] * d[2];
        J[1 * 6 + j] = g[1][0] * d[0] + g[1][1] * d[1] + g[1][2] * d[2];
    }
    for (int i = 0; i < 3; ++i) {
        J[0 * 6 + 3 + i] = g[0][i];
        J[1 * 6 + 3 + i] = g[1][i];
    }
}

/* null space of a bearing (MLPnPsolver.cpp:336-338): JacobiSVD with a
 * Householder-QR preconditioner of the 1x3 matrix f^T, full V, columns 1..2 --
 * i.e. the last two columns of the Householder reflector that maps f onto e1.
 * Any orthonormal basis of the plane orthogonal to f is equivalent (SURVEY
 * Appendix A). N is 3x2 row-major. */
static void bearing_nullspace(const double f[3], double N[6])
{
    const double nf = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    double v[3] = {f[0], f[1], f[2]};
    v[0] = (f[0] >= 0.0) ? f[0] + nf : f[0] - nf;
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double beta = 2.0 / vv;
    for (int r = 0; r < 3; ++r)
        for (int c = 1; c < 3; ++c) N[r * 2 + (c - 1)] = ((r == c) ? 1.0 : 0.0) - beta * v[r] * v[c];
}

static double det3(const double R[9])
{
    return R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
}

static double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* MLPnPsolver::mlpnp_gn (MLPnPsolver.cpp:659-723) */
static void mlpnp_gn(double x[6], int n, const double *pts, const double *nulls, const double *P /* n*4 or NULL */)
{
    const double epsP = 1e-5;
    double *Jall = (double *)malloc(sizeof(double) * (size_t)n * 12);
    double *rall = (double *)malloc(sizeof(double) * (size_t)n * 2);
    int it_cnt = 0;
    while (it_cnt < 5) {
        double A[36], g[6], dx[6];
        memset(A, 0, sizeof(A));
        memset(g, 0, sizeof(g));
        for (int i = 0; i < n; ++i) {
            const double *N = nulls + 6 * i;
            const double nr[3] = {N[0], N[2], N[4]}, ns[3] = {N[1], N[3], N[5]};
            double *J = Jall + 12 * i, *r = rall + 2 * i;
            orc_mlpnp_res_jac(pts + 3 * i, nr, ns, x, x + 3, r, J);
            /* JacTSKll = J^T Kll (:694-697); A = JacTSKll*J (:699); g = JacTSKll*r (:702) */
            double W0[6], W1[6], wr0, wr1;
            if (P) {
                const double *p = P + 4 * i;
                for (int c = 0; c < 6; ++c) {
                    W0[c] = fma(J[c], p[0], J[6 + c] * p[2]);
                    W1[c] = fma(J[c], p[1], J[6 + c] * p[3]);
                }
            } else {
                for (int c = 0; c < 6; ++c) { W0[c] = J[c]; W1[c] = J[6 + c]; }
            }
            wr0 = r[0]; wr1 = r[1];
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) {
                    A[a * 6 + b] = fma(W0[a], J[b], A[a * 6 + b]);
                    A[a * 6 + b] = fma(W1[a], J[6 + b], A[a * 6 + b]);
                }
                g[a] = fma(W0[a], wr0, g[a]);
                g[a] = fma(W1[a], wr1, g[a]);
            }
        }
        orc_ldlt6_solve_d(A, g, dx);                                   /* :705-706 */
        double mx = 0.0, mn = INFINITY;
        for (int c = 0; c < 6; ++c) {
            const double v = fabs(dx[c]);
            if (v > mx) mx = v;
            if (v < mn) mn = v;
        }
        if (mx > 5.0 || mn > 1.0) break;                               /* :709-710 */
        double mdl = 0.0;
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 2; ++k) {
                const double *J = Jall + 12 * i + 6 * k;
                const double dl = J[0] * dx[0] + J[1] * dx[1] + J[2] * dx[2] + J[3] * dx[3] + J[4] * dx[4] + J[5] * dx[5];
                if (fabs(dl) > mdl) mdl = fabs(dl);
            }
        for (int c = 0; c < 6; ++c) x[c] = x[c] - dx[c];               /* :715 / :718 */
        if (mdl < epsP) break;                                         /* :713-716 */
        ++it_cnt;
    }
    free(Jall);
    free(rall);
}

/* MLPnPsolver::computePose (MLPnPsolver.cpp:321-623).
 * f: n bearings, p: n points (both double), cov: n 3x3 or NULL. */
static void compute_pose(int n, const double *f, const double *p, const double *cov, double Rres[9], double tres[3])
{
    double *nulls = (double *)malloc(sizeof(double) * (size_t)n * 6);
    double *pts3 = (double *)malloc(sizeof(double) * (size_t)n * 3);
    double *P = cov ? (double *)malloc(sizeof(double) * (size_t)n * 4) : NULL;
    for (int i = 0; i < n; ++i) bearing_nullspace(f + 3 * i, nulls + 6 * i);    /* :332-340 */
    memcpy(pts3, p, sizeof(double) * (size_t)n * 3);

    /* 1. planarity test (:346-364) */
    double planarTest[9];
    memset(planarTest, 0, sizeof(planarTest));
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) planarTest[r * 3 + c] = fma(p[3 * i + r], p[3 * i + c], planarTest[r * 3 + c]);
    double eigenRot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int planar = 0;
    if (orc_rank3_fullpiv_d(planarTest) == 2) {
        planar = 1;
        double A[9], w[3], V[9];
        memcpy(A, planarTest, sizeof(A));
        orc_jacobi_eig_d(3, A, w, V);                                  /* :359 */
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) eigenRot[r * 3 + c] = V[c * 3 + r];   /* transposeInPlace :361 */
        for (int i = 0; i < n; ++i) {
            const double *q = p + 3 * i;
            for (int r = 0; r < 3; ++r)
                pts3[3 * i + r] = eigenRot[r * 3 + 0] * q[0] + eigenRot[r * 3 + 1] * q[1] + eigenRot[r * 3 + 2] * q[2];
        }
    }

    /* 2. stochastic model (:368-388): P_i = (N_i^T Sigma_i N_i)^-1 */
    if (cov) {
        for (int i = 0; i < n; ++i) {
            const double *N = nulls + 6 * i, *S = cov + 9 * i;
            double SN[6];   /* Sigma * N, 3x2 */
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 2; ++c)
                    SN[r * 2 + c] = S[r * 3 + 0] * N[0 * 2 + c] + S[r * 3 + 1] * N[1 * 2 + c] + S[r * 3 + 2] * N[2 * 2 + c];
            double T[4];
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 2; ++c)
                    T[r * 2 + c] = N[0 * 2 + r] * SN[0 * 2 + c] + N[1 * 2 + r] * SN[1 * 2 + c] + N[2 * 2 + r] * SN[2 * 2 + c];
            const double det = fma(T[0], T[3], -(T[1] * T[2]));
            const double id = 1.0 / det;                               /* Matrix2d::inverse() :381 */
            P[4 * i + 0] = T[3] * id;
            P[4 * i + 1] = -T[1] * id;
            P[4 * i + 2] = -T[2] * id;
            P[4 * i + 3] = T[0] * id;
        }
    }

    /* 3. design matrix rows (:393-477) and 4. AtPA (:482-486), accumulated point by point */
    const int cols = planar ? 9 : 12;
    double AtPA[144];
    memset(AtPA, 0, sizeof(AtPA));
    for (int i = 0; i < n; ++i) {
        const double *N = nulls + 6 * i, *pt = pts3 + 3 * i;
        double a0[12], a1[12];
        if (planar) {
            for (int r = 0; r < 3; ++r) {
                a0[2 * r + 0] = N[r * 2 + 0] * pt[1]; a1[2 * r + 0] = N[r * 2 + 1] * pt[1];
                a0[2 * r + 1] = N[r * 2 + 0] * pt[2]; a1[2 * r + 1] = N[r * 2 + 1] * pt[2];
                a0[6 + r] = N[r * 2 + 0];             a1[6 + r] = N[r * 2 + 1];
            }
        } else {
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) {
                    a0[3 * r + c] = N[r * 2 + 0] * pt[c];
                    a1[3 * r + c] = N[r * 2 + 1] * pt[c];
                }
                a0[9 + r] = N[r * 2 + 0];
                a1[9 + r] = N[r * 2 + 1];
            }
        }
        double w0[12], w1[12];   /* rows of P_i * A_i */
        if (P) {
            const double *pp = P + 4 * i;
            for (int c = 0; c < cols; ++c) {
                w0[c] = fma(pp[0], a0[c], pp[1] * a1[c]);
                w1[c] = fma(pp[2], a0[c], pp[3] * a1[c]);
            }
        } else {
            memcpy(w0, a0, sizeof(double) * (size_t)cols);
            memcpy(w1, a1, sizeof(double) * (size_t)cols);
        }
        for (int a = 0; a < cols; ++a)
            for (int b = a; b < cols; ++b) {
                AtPA[a * cols + b] = fma(a0[a], w0[b], AtPA[a * cols + b]);
                AtPA[a * cols + b] = fma(a1[a], w1[b], AtPA[a * cols + b]);
            }
    }
    double ev[1], result1[12];
    orc_jacobi_lowest_d(cols, 1, AtPA, ev, result1);   /* :488-489 last right-singular vector = eigenvector of the smallest eigenvalue */

    double Rout[9], tout[3];
    if (planar) {                                                       /* :497-558 */
        double tmp[9] = {0.0, result1[0], result1[1], 0.0, result1[2], result1[3], 0.0, result1[4], result1[5]};
        const double c1[3] = {tmp[1], tmp[4], tmp[7]}, c2[3] = {tmp[2], tmp[5], tmp[8]};
        double c0[3];
        cross3(c1, c2, c0);
        tmp[0] = c0[0]; tmp[3] = c0[1]; tmp[6] = c0[2];
        double tt[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) tt[r * 3 + c] = tmp[c * 3 + r];                  /* transposeInPlace :507 */
        const double tc1[3] = {tt[1], tt[4], tt[7]}, tc2[3] = {tt[2], tt[5], tt[8]};
        const double scale = 1.0 / sqrt(fabs(norm3(tc1) * norm3(tc2)));                   /* :509 */
        double Rout1[9];
        orc_polar3_d(tt, Rout1);                                                          /* :511-512 */
        if (det3(Rout1) < 0)
            for (int i = 0; i < 9; ++i) Rout1[i] *= -1.0;
        double Rb[9];
        for (int r = 0; r < 3; ++r)                                                       /* eigenRot^T * Rout1 :517 */
            for (int c = 0; c < 3; ++c)
                Rb[r * 3 + c] = eigenRot[0 * 3 + r] * Rout1[0 * 3 + c] + eigenRot[1 * 3 + r] * Rout1[1 * 3 + c] + eigenRot[2 * 3 + r] * Rout1[2 * 3 + c];
        const double t[3] = {scale * result1[6], scale * result1[7], scale * result1[8]};
        double Rc[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rc[r * 3 + c] = -Rb[c * 3 + r];                   /* transpose, *= -1 :520-521 */
        if (det3(Rc) < 0.0) { Rc[2] *= -1; Rc[5] *= -1; Rc[8] *= -1; }                   /* col(2) *= -1 :523 */
        double Rs[2][9];
        for (int r = 0; r < 3; ++r) {
            Rs[0][r * 3 + 0] = Rc[r * 3 + 0]; Rs[0][r * 3 + 1] = Rc[r * 3 + 1]; Rs[0][r * 3 + 2] = Rc[r * 3 + 2];
            Rs[1][r * 3 + 0] = -Rc[r * 3 + 0]; Rs[1][r * 3 + 1] = -Rc[r * 3 + 1]; Rs[1][r * 3 + 2] = Rc[r * 3 + 2];
        }
        double best = 0.0;
        int bi = -1;
        for (int k = 0; k < 4; ++k) {                                                     /* :533-556 */
            const double *Rk = Rs[k / 2];
            const double sg = (k % 2 == 0) ? 1.0 : -1.0;
            double norms = 0.0;
            for (int q = 0; q < 6; ++q) {
                const double *pp = p + 3 * q;
                double v[3];
                for (int r = 0; r < 3; ++r) v[r] = (Rk[r * 3 + 0] * pp[0] + Rk[r * 3 + 1] * pp[1] + Rk[r * 3 + 2] * pp[2]) + sg * t[r];
                const double vn = norm3(v);
                norms += (1.0 - ((v[0] / vn) * f[3 * q + 0] + (v[1] / vn) * f[3 * q + 1] + (v[2] / vn) * f[3 * q + 2]));
            }
            if (bi < 0 || norms < best) { best = norms; bi = k; }                         /* min_element: first minimum */
        }
        memcpy(Rout, Rs[bi / 2], sizeof(Rout));
        for (int r = 0; r < 3; ++r) tout[r] = (bi % 2 == 0) ? t[r] : -t[r];
    } else {                                                            /* :559-602 */
        double tmp[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) tmp[r * 3 + c] = result1[3 * c + r];              /* :562-564 */
        const double c0[3] = {tmp[0], tmp[3], tmp[6]}, c1[3] = {tmp[1], tmp[4], tmp[7]}, c2[3] = {tmp[2], tmp[5], tmp[8]};
        const double scale = 1.0 / pow(fabs(norm3(c0) * norm3(c1) * norm3(c2)), 1.0 / 3.0);   /* :566-567 */
        orc_polar3_d(tmp, Rout);                                                          /* :570-571 */
        if (det3(Rout) < 0)
            for (int i = 0; i < 9; ++i) Rout[i] *= -1.0;
        const double ts[3] = {scale * result1[9], scale * result1[10], scale * result1[11]};
        double t0[3];
        for (int r = 0; r < 3; ++r) t0[r] = Rout[r * 3 + 0] * ts[0] + Rout[r * 3 + 1] * ts[1] + Rout[r * 3 + 2] * ts[2];   /* :576 */
        /* :579-600  Ts[s] = [Rout, +-tout]^-1 (general inverse of an affine 4x4:
         * inverse of the 3x3 block by cofactors, translation -Rinv*t) */
        double Rinv[9];
        orc_inv3_d(Rout, Rinv);
        double err[2], tinv[2][3];
        for (int s = 0; s < 2; ++s) {
            const double sg = (s == 0) ? 1.0 : -1.0;
            for (int r = 0; r < 3; ++r)
                tinv[s][r] = -(Rinv[r * 3 + 0] * (sg * t0[0]) + Rinv[r * 3 + 1] * (sg * t0[1]) + Rinv[r * 3 + 2] * (sg * t0[2]));
            err[s] = 0.0;
            for (int q = 0; q < 6; ++q) {
                const double *pp = p + 3 * q;
                double v[3];
                for (int r = 0; r < 3; ++r) v[r] = (Rinv[r * 3 + 0] * pp[0] + Rinv[r * 3 + 1] * pp[1] + Rinv[r * 3 + 2] * pp[2]) + tinv[s][r];
                const double vn = norm3(v);
                err[s] += (1.0 - ((v[0] / vn) * f[3 * q + 0] + (v[1] / vn) * f[3 * q + 1] + (v[2] / vn) * f[3 * q + 2]));
            }
        }
        const int pick = (err[0] < err[1]) ? 0 : 1;                                       /* :596-599 */
        for (int r = 0; r < 3; ++r) tout[r] = tinv[pick][r];
        memcpy(Rout, Rinv, sizeof(Rout));                                                 /* :600 */
    }

    /* 5. Gauss-Newton (:607-622) */
    double x[6];
    orc_rot2rodrigues(Rout, x);
    x[3] = tout[0]; x[4] = tout[1]; x[5] = tout[2];
    mlpnp_gn(x, n, p, nulls, P);
    orc_rodrigues2rot(x, Rres);
    tres[0] = x[3]; tres[1] = x[4]; tres[2] = x[5];
    free(nulls); free(pts3); free(P);
}

/* bearing of a keypoint as the constructor builds it (MLPnPsolver.cpp:33-37):
 * f32 arithmetic, widened, NOT normalised */
static void make_bearing(const orc_mlpnp_problem *pb, int i, double f[3])
{
    const float x = (pb->p2d[2 * i + 0] - pb->cx) / pb->fx;
    const float y = (pb->p2d[2 * i + 1] - pb->cy) / pb->fy;
    f[0] = (double)x; f[1] = (double)y; f[2] = (double)1.f;
}

void orc_mlpnp_pose(const orc_mlpnp_problem *pb, const uint32_t *idx, int m, double R[9], double t[3])
{
    double *f = (double *)malloc(sizeof(double) * (size_t)m * 3);
    double *p = (double *)malloc(sizeof(double) * (size_t)m * 3);
    double *cov = pb->cov ? (double *)malloc(sizeof(double) * (size_t)m * 9) : NULL;
    for (int i = 0; i < m; ++i) {
        make_bearing(pb, (int)idx[i], f + 3 * i);
        for (int c = 0; c < 3; ++c) p[3 * i + c] = (double)pb->p3d[3 * idx[i] + c];     /* :40-42 */
        if (cov) memcpy(cov + 9 * i, pb->cov + 9 * (size_t)idx[i], 9 * sizeof(double));
    }
    compute_pose(m, f, p, cov, R, t);
    free(f); free(p); free(cov);
}

/* MLPnPsolver::CheckInliers (MLPnPsolver.cpp:222-255): f64 rigid transform of
 * the f32-narrowed point narrowed to f32, two true f32 divisions */
int orc_mlpnp_check_inliers(const orc_mlpnp_problem *pb, const float *max_err, const double R[9],
                            const double t[3], uint8_t *mask, float *err2)
{
    int cnt = 0;
    for (int i = 0; i < pb->n; ++i) {
        const float X = pb->p3d[3 * i + 0], Y = pb->p3d[3 * i + 1], Z = pb->p3d[3 * i + 2];   /* cv::Point3f :228 */
        const float xc = (float)(R[0] * X + R[1] * Y + R[2] * Z + t[0]);                      /* :231-233 */
        const float yc = (float)(R[3] * X + R[4] * Y + R[5] * Z + t[1]);
        const float zc = (float)(R[6] * X + R[7] * Y + R[8] * Z + t[2]);
        const float u = pb->fx * xc / zc + pb->cx;                                            /* :236-237 */
        const float v = pb->fy * yc / zc + pb->cy;
        const float distX = pb->p2d[2 * i + 0] - u, distY = pb->p2d[2 * i + 1] - v;
        const float error2 = distX * distX + distY * distY;
        const int in = error2 < max_err[i];
        if (mask) mask[i] = (uint8_t)in;
        if (err2) err2[i] = error2;
        cnt += in;
    }
    return cnt;
}

static void set_T_d(float T[16], const double R[9], const double t[3])
{
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = (float)R[r * 3 + c];   /* convertTo(CV_32F) :135-136 */
        T[r * 4 + 3] = (float)t[r];
    }
}

/* MLPnPsolver::iterate (MLPnPsolver.cpp:56-183) first call + Refine (:257-318) */
void orc_mlpnp_ransac(const orc_mlpnp_problem *pb, const orc_ransac_params *prm, const uint32_t *table,
                      int flags, orc_result *res, uint8_t *mask, int *hyp_counts, double *hyp_pose)
{
    const int N = pb->n;
    int minInl, H;
    orc_pnp_ransac_setup(N, prm, &minInl, &H);                          /* :185-220 (same formula) */
    const int minSet = prm->min_set;
    const int exhaustive = (flags & ORC_FLAG_EXHAUSTIVE) != 0;
    const int discard = (flags & ORC_FLAG_MLPNP_DISCARD_REFINE) != 0;

    memset(res, 0, sizeof(*res));
    res->best_hyp = -1;
    for (int i = 0; i < 16; ++i) res->T[i] = (i % 5 == 0) ? 1.0f : 0.0f;   /* Tout.setIdentity() :57 */
    res->scale = 1.0f;
    if (mask) memset(mask, 0, (size_t)N);
    if (N < minInl) { res->no_more = 1; return; }                       /* :62-66 */

    float *maxErr = (float *)malloc(sizeof(float) * (size_t)N);
    for (int i = 0; i < N; ++i) maxErr[i] = pb->sigma2[i] * prm->th2;   /* :219 */
    uint8_t *cur = (uint8_t *)malloc((size_t)N), *best = (uint8_t *)calloc((size_t)N, 1), *ref = (uint8_t *)malloc((size_t)N);
    uint32_t *sel = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)N);
    int nBest = 0, done = 0;
    double Rb[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tb[3] = {0, 0, 0};

    for (int h = 0; h < H; ++h) {                                       /* :71 */
        double Ri[9], ti[3];
        orc_mlpnp_pose(pb, table + (size_t)h * minSet, minSet, Ri, ti); /* :84-120 */
        const int cnt = orc_mlpnp_check_inliers(pb, maxErr, Ri, ti, cur, NULL);   /* :123 */
        res->n_hyp = h + 1;
        if (hyp_counts) hyp_counts[h] = cnt;
        if (hyp_pose) { memcpy(hyp_pose + (size_t)h * 12, Ri, sizeof(Ri)); memcpy(hyp_pose + (size_t)h * 12 + 9, ti, sizeof(ti)); }
        if (done) continue;
        if (cnt >= minInl) {                                            /* :125 */
            if (cnt > nBest) {                                          /* :128 */
                memcpy(best, cur, (size_t)N);
                nBest = cnt;
                memcpy(Rb, Ri, sizeof(Rb)); memcpy(tb, ti, sizeof(tb));
                res->best_hyp = h;
            }
            /* Refine (:257-318) */
            int m = 0;
            for (int i = 0; i < N; ++i)
                if (best[i]) sel[m++] = (uint32_t)i;
            double Rr[9], tr[3];
            orc_mlpnp_pose(pb, sel, m, Rr, tr);                         /* :290 */
            if (discard) { memcpy(Rr, Ri, sizeof(Rr)); memcpy(tr, ti, sizeof(tr)); }   /* Q6: result never copied to mRi/mti */
            const int cr = orc_mlpnp_check_inliers(pb, maxErr, Rr, tr, ref, NULL);     /* :293 */
            res->n_refines++;
            if (cr > minInl) {                                          /* :298 */
                res->ok = 1; res->refined = 1; res->n_inliers = cr;
                set_T_d(res->T, Rr, tr);
                if (mask) memcpy(mask, ref, (size_t)N);
                done = 1;
                if (!exhaustive) break;
            } else {
                res->n_failed_refines++;
            }
        }
    }
    res->best_count = nBest;
    if (!done) {
        res->no_more = 1;                                               /* :165-167 */
        if (nBest >= minInl) {                                          /* :168-179 */
            res->ok = 1; res->n_inliers = nBest;
            set_T_d(res->T, Rb, tb);
            if (mask) memcpy(mask, best, (size_t)N);
        }
    }
    free(maxErr); free(cur); free(best); free(ref); free(sel);
}
