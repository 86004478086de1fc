/*
 * oracle/orc_poseopt.c -- CPU oracle for Optimizer::PoseOptimization.
 *
 * TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Restates, function by function, what the reference executes for one frame:
 *   src/Optimizer.cpp:205-424            PoseOptimization (4 rounds x 10 LM iterations, outlier re-classification)
 *   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:59-179   LM step control
 *   Thirdparty/g2o/g2o/core/sparse_optimizer.cpp  optimize()              iteration loop / termination
 *   Thirdparty/g2o/g2o/core/base_unary_edge.hpp   constructQuadraticForm  H += J^T (rho' Omega) J, b -= rho' J^T Omega e
 *   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91                  Huber
 *   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:266-364             Jacobians, cam_project (mono, stereo)
 *   Thirdparty/g2o/g2o/types/se3quat.h:104-110,217-260                    SE3Quat product, map, exp
 *   Thirdparty/g2o/g2o/solvers/linear_solver_dense.h                      dense 6x6 Cholesky (Eigen LDLT)
 *
 * PARITY UNPINNED: g2o's vector arithmetic is Eigen (Quaterniond(R), quaternion * vector, LDLT with pivoting,
 * A^T * Omega * A evaluation order); it is restated here by the published formulas, in a fixed scalar order that
 * the device core (csrc/poseopt.cuh) shares.  (2 rho - 1)^3 is evaluated as d*d*d instead of pow(d, 3).
 *
 * Reference behaviours that are reproduced on purpose:
 *   - every round restarts from the frame's initial pose (Optimizer.cpp:318: setEstimate(pFrame->mTcw));
 *   - an edge that was an inlier going into the classification keeps the error of the LAST computeActiveErrors(),
 *     i.e. of the last LM trial even when that trial was rejected and the estimate rolled back
 *     (optimization_algorithm_levenberg.cpp:125-126,148 and Optimizer.cpp:332-337 recompute only flagged edges);
 *   - chi2 is narrowed to float before the comparison with the float thresholds 5.991f / 7.815f (Optimizer.cpp:339);
 *   - stereo projection uses a float 1/z (types_six_dof_expmap.cpp:300);
 *   - the robust kernel is dropped after the third round (it == 2), `edges().size() < 10` stops after one round.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orc.h"

typedef struct { double q[4]; /* w x y z */ double t[3]; } se3_t;

/* Eigen::Quaterniond(Matrix3d) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>) */
static void quat_from_rot(const double m[9], double q[4])
{
    double t = m[0] + m[4] + m[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (m[7] - m[5]) * t;
        q[2] = (m[2] - m[6]) * t;
        q[3] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

/* SE3Quat::normalizeRotation (se3quat.h:60-66) */
static void quat_normalize(double q[4])
{
    if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

static void quat_mul(const double a[4], const double b[4], double o[4])
{
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

/* Quaternion * Vector3 (Eigen _transformVector): v + w*uv + qv x uv with uv = 2 (qv x v) */
static void quat_rotate(const double q[4], const double v[3], double o[3])
{
    double uv[3];
    uv[0] = q[2] * v[2] - q[3] * v[1];
    uv[1] = q[3] * v[0] - q[1] * v[2];
    uv[2] = q[1] * v[1] - q[2] * v[0];
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    o[0] = v[0] + q[0] * uv[0] + (q[2] * uv[2] - q[3] * uv[1]);
    o[1] = v[1] + q[0] * uv[1] + (q[3] * uv[0] - q[1] * uv[2]);
    o[2] = v[2] + q[0] * uv[2] + (q[1] * uv[1] - q[2] * uv[0]);
}

/* Quaternion::toRotationMatrix */
static void quat_to_rot(const double q[4], double R[9])
{
    const double tx = 2.0 * q[1], ty = 2.0 * q[2], tz = 2.0 * q[3];
    const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
    const double txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
    const double tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

/* Converter::toSE3Quat (Converter.cpp:16-22) + SE3Quat(R, t) constructor */
static void se3_from_float(const float R[9], const float t[3], se3_t *T)
{
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = (double)R[i];
    quat_from_rot(m, T->q);
    quat_normalize(T->q);
    for (int i = 0; i < 3; ++i) T->t[i] = (double)t[i];
}

/* SE3Quat::exp (se3quat.h:223-260): update = (omega, upsilon) */
static void se3_exp(const double x[6], se3_t *E)
{
    const double w0 = x[0], w1 = x[1], w2 = x[2];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double O2[9], R[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta;
        const double b = (1.0 - cos(theta)) / (theta * theta);
        const double c = (theta - sin(theta)) / (theta * theta * theta);
        for (int i = 0; i < 9; ++i) {
            const double id = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = id + a * O[i] + b * O2[i];
            V[i] = id + b * O[i] + c * O2[i];
        }
    }
    quat_from_rot(R, E->q);
    quat_normalize(E->q);
    for (int i = 0; i < 3; ++i) E->t[i] = V[3 * i] * x[3] + V[3 * i + 1] * x[4] + V[3 * i + 2] * x[5];
}

/* SE3Quat::operator* (se3quat.h:104-110): result = a * b */
static void se3_mul(const se3_t *a, const se3_t *b, se3_t *o)
{
    double rt[3], q[4];
    quat_rotate(a->q, b->t, rt);
    for (int i = 0; i < 3; ++i) o->t[i] = a->t[i] + rt[i];
    quat_mul(a->q, b->q, q);
    memcpy(o->q, q, sizeof q);
    quat_normalize(o->q);
}

/* computeError of both edge types (types_six_dof_expmap.h:186-187 and the mono twin, cam_project :290-306) */
static int edge_error(const se3_t *T, const orc_poseopt_problem *pb, int i, double e[3], double pc[3])
{
    const double Xw[3] = {(double)pb->p3d[3 * i], (double)pb->p3d[3 * i + 1], (double)pb->p3d[3 * i + 2]};
    double r[3];
    quat_rotate(T->q, Xw, r);
    pc[0] = r[0] + T->t[0]; pc[1] = r[1] + T->t[1]; pc[2] = r[2] + T->t[2];
    const double fx = (double)pb->K[0], fy = (double)pb->K[1], cx = (double)pb->K[2], cy = (double)pb->K[3];
    const double ou = (double)pb->obs[3 * i], ov = (double)pb->obs[3 * i + 1];
    const float ur = pb->obs[3 * i + 2];
    if (ur < 0.0f) {
        e[0] = ou - ((pc[0] / pc[2]) * fx + cx);
        e[1] = ov - ((pc[1] / pc[2]) * fy + cy);
        e[2] = 0.0;
        return 2;
    }
    const float invz = (float)(1.0 / pc[2]);
    const double p0 = pc[0] * (double)invz * fx + cx;
    const double p1 = pc[1] * (double)invz * fy + cy;
    const double p2 = p0 - (double)pb->K[4] * (double)invz;
    e[0] = ou - p0; e[1] = ov - p1; e[2] = (double)ur - p2;
    return 3;
}

static double edge_chi2(const double e[3], int dim, double s)
{
    double c = e[0] * (s * e[0]) + e[1] * (s * e[1]);
    if (dim == 3) c += e[2] * (s * e[2]);
    return c;
}

/* no-pivot LDL^T of a symmetric positive definite 6x6; returns 0 when a pivot is not positive (isPositive() false) */
static int ldlt6(const double Hin[36], const double b[6], double x[6])
{
    double L[36], d[6], r[6], y[6];
    memcpy(L, Hin, sizeof L);
    for (int j = 0; j < 6; ++j) {
        double dj = L[7 * j];
        for (int k = 0; k < j; ++k) dj -= L[6 * j + k] * L[6 * j + k] * d[k];
        if (!(dj > 0.0)) return 0;
        d[j] = dj;
        r[j] = 1.0 / dj;               /* one reciprocal per pivot (shared with the device core) */
        for (int i = j + 1; i < 6; ++i) {
            double v = L[6 * i + j];
            for (int k = 0; k < j; ++k) v -= L[6 * i + k] * L[6 * j + k] * d[k];
            L[6 * i + j] = v * r[j];
        }
    }
    for (int i = 0; i < 6; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= L[6 * i + k] * y[k];
        y[i] = v;
    }
    for (int i = 0; i < 6; ++i) y[i] *= r[i];
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < 6; ++k) v -= L[6 * k + i] * x[k];
        x[i] = v;
    }
    return 1;
}

typedef struct {
    const orc_poseopt_problem *pb;
    const uint8_t *level;    /* 1 = excluded from the optimisation (setLevel(1)) */
    int robust;
    double *err;             /* [n][3] the edges' stored _error */
} graph_t;

/* computeActiveErrors + activeRobustChi2 */
static double active_chi2(const graph_t *g, const se3_t *T)
{
    const orc_poseopt_problem *pb = g->pb;
    double sum = 0.0;
    for (int i = 0; i < pb->n; ++i) {
        if (g->level[i]) continue;
        double pc[3];
        const int dim = edge_error(T, pb, i, g->err + 3 * i, pc);
        const double c = edge_chi2(g->err + 3 * i, dim, (double)pb->inv_sigma2[i]);
        if (g->robust) {
            const double delta = (double)(dim == 2 ? sqrtf(5.991f) : sqrtf(7.815f));   /* const float deltaMono = sqrt(5.991) */
            const double dsqr = delta * delta;
            sum += (c <= dsqr) ? c : 2.0 * sqrt(c) * delta - dsqr;
        } else {
            sum += c;
        }
    }
    return sum;
}

/* buildSystem: linearizeOplus + constructQuadraticForm over the active edges (errors already stored) */
static void build_system(const graph_t *g, const se3_t *T, double H[36], double b[6])
{
    const orc_poseopt_problem *pb = g->pb;
    const double fx = (double)pb->K[0], fy = (double)pb->K[1], bf = (double)pb->K[4];
    memset(H, 0, 36 * sizeof(double));
    memset(b, 0, 6 * sizeof(double));
    for (int i = 0; i < pb->n; ++i) {
        if (g->level[i]) continue;
        const double Xw[3] = {(double)pb->p3d[3 * i], (double)pb->p3d[3 * i + 1], (double)pb->p3d[3 * i + 2]};
        double r[3];
        quat_rotate(T->q, Xw, r);
        const double x = r[0] + T->t[0], y = r[1] + T->t[1], z = r[2] + T->t[2];
        const int dim = pb->obs[3 * i + 2] < 0.0f ? 2 : 3;
        const double invz = 1.0 / z, invz_2 = invz * invz;
        double J[18];
        J[0] = x * y * invz_2 * fx;
        J[1] = -(1.0 + (x * x * invz_2)) * fx;
        J[2] = y * invz * fx;
        J[3] = -invz * fx;
        J[4] = 0.0;
        J[5] = x * invz_2 * fx;
        J[6] = (1.0 + y * y * invz_2) * fy;
        J[7] = -x * y * invz_2 * fy;
        J[8] = -x * invz * fy;
        J[9] = 0.0;
        J[10] = -invz * fy;
        J[11] = y * invz_2 * fy;
        if (dim == 3) {
            J[12] = J[0] - bf * y * invz_2;
            J[13] = J[1] + bf * x * invz_2;
            J[14] = J[2];
            J[15] = J[3];
            J[16] = 0.0;
            J[17] = J[5] - bf * invz_2;
        }
        const double s = (double)pb->inv_sigma2[i];
        const double *e = g->err + 3 * i;
        double rho1 = 1.0;
        if (g->robust) {
            const double c = edge_chi2(e, dim, s);
            const double delta = (double)(dim == 2 ? sqrtf(5.991f) : sqrtf(7.815f));
            if (c > delta * delta) rho1 = delta / sqrt(c);
        }
        /* explicit fused multiply-adds (arithmetic contract shared with csrc/poseopt.cuh); the exact zeros of the
         * Jacobian (column 4 of rows 0 and 2, column 3 of row 1) are skipped */
        const double wo = rho1 * s;
        const int st = dim == 3;
        double w0[6], w1[6], w2[6];
        for (int c2 = 0; c2 < 6; ++c2) { w0[c2] = wo * J[c2]; w1[c2] = wo * J[6 + c2]; w2[c2] = st ? wo * J[12 + c2] : 0.0; }
        const double we0 = s * e[0], we1 = s * e[1], we2 = s * e[2];
        for (int a = 0; a < 6; ++a) {
            double be = 0.0;
            if (a != 4) be = J[a] * we0;
            if (a != 3) be = fma(J[6 + a], we1, be);
            if (st && a != 4) be = fma(J[12 + a], we2, be);
            b[a] -= rho1 * be;
            for (int c2 = a; c2 < 6; ++c2) {
                double h = H[6 * a + c2];
                if (a != 4 && c2 != 4) h = fma(J[a], w0[c2], h);
                if (a != 3 && c2 != 3) h = fma(J[6 + a], w1[c2], h);
                if (st && a != 4 && c2 != 4) h = fma(J[12 + a], w2[c2], h);
                H[6 * a + c2] = h;
            }
        }
    }
    for (int a = 0; a < 6; ++a)
        for (int c2 = 0; c2 < a; ++c2) H[6 * a + c2] = H[6 * c2 + a];
}

/* SparseOptimizer::optimize(10) with OptimizationAlgorithmLevenberg; T is the estimate, *Terr the pose the stored errors belong to */
static void optimize(const graph_t *g, se3_t *T, se3_t *Terr, int iterations, int *n_iter, int *n_trials)
{
    double lambda = 0.0, ni = 2.0;
    int nBad = 0;
    double x[6] = {0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iterations; ++it) {
        double currentChi = active_chi2(g, T);
        *Terr = *T;
        const double iniChi = currentChi;
        double tempChi;
        double H[36], b[6];
        build_system(g, T, H, b);
        ++*n_iter;
        if (it == 0) {
            double maxDiag = 0.0;
            for (int j = 0; j < 6; ++j) maxDiag = fmax(fabs(H[7 * j]), maxDiag);
            lambda = 1e-5 * maxDiag;
            ni = 2.0;
            nBad = 0;
        }
        double rho = 0.0;
        int qmax = 0;
        do {
            const se3_t backup = *T;
            double Hl[36];
            memcpy(Hl, H, sizeof Hl);
            for (int j = 0; j < 6; ++j) Hl[7 * j] += lambda;
            const int ok2 = ldlt6(Hl, b, x);
            ++*n_trials;
            se3_t E, Tn;
            se3_exp(x, &E);
            se3_mul(&E, T, &Tn);
            *T = Tn;
            tempChi = active_chi2(g, T);
            *Terr = *T;
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0.0;
            for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0.0 && isfinite(tempChi)) {
                const double d = 2.0 * rho - 1.0;
                double alpha = 1.0 - d * d * d;
                alpha = fmin(alpha, 2.0 / 3.0);
                const double scaleFactor = fmax(1.0 / 3.0, alpha);
                lambda *= scaleFactor;
                ni = 2.0;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2.0;
                *T = backup;
            }
            ++qmax;
        } while (rho < 0.0 && qmax < 10);
        if (qmax == 10 || rho == 0.0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) break;
    }
}

void orc_pose_optimization(const orc_poseopt_problem *pb, orc_poseopt_result *res, uint8_t *outlier)
{
    memset(res, 0, sizeof *res);
    const int n = pb->n;
    se3_t T;
    se3_from_float(pb->Rcw, pb->tcw, &T);
    for (int i = 0; i < n; ++i) outlier[i] = 0;
    int nBad = 0;
    if (n >= 3) {
        uint8_t *level = (uint8_t *)calloc((size_t)n, 1);
        double *err = (double *)calloc((size_t)n * 3, sizeof(double));
        graph_t g = {pb, level, 1, err};
        const float chi2Mono = 5.991f, chi2Stereo = 7.815f;
        for (int it = 0; it < 4; ++it) {
            se3_from_float(pb->Rcw, pb->tcw, &T);
            se3_t Terr = T;
            int active = 0;
            for (int i = 0; i < n; ++i) active += !level[i];
            if (active > 0) optimize(&g, &T, &Terr, 10, &res->iterations, &res->trials);
            nBad = 0;
            for (int i = 0; i < n; ++i) {
                double pc[3];
                int dim = pb->obs[3 * i + 2] < 0.0f ? 2 : 3;
                if (outlier[i]) dim = edge_error(&T, pb, i, err + 3 * i, pc);
                const float chi2 = (float)edge_chi2(err + 3 * i, dim, (double)pb->inv_sigma2[i]);
                if (chi2 > (dim == 2 ? chi2Mono : chi2Stereo)) { outlier[i] = 1; level[i] = 1; ++nBad; }
                else { outlier[i] = 0; level[i] = 0; }
            }
            if (it == 2) g.robust = 0;
            ++res->rounds;
            if (n < 10) break;
        }
        free(level);
        free(err);
        res->n_inliers = n - nBad;
    }
    res->n_bad = nBad;
    quat_to_rot(T.q, res->R);
    for (int i = 0; i < 3; ++i) res->t[i] = T.t[i];
    for (int i = 0; i < 9; ++i) res->Rf[i] = (float)res->R[i];
    for (int i = 0; i < 3; ++i) res->tf[i] = (float)res->t[i];
}

/* one call per problem, sequentially (the reference's shape: Tracking.cpp:1284,1300,1315 call it per candidate) */
void orc_pose_optimization_batch(int C, const orc_poseopt_problem *pbs, orc_poseopt_result *res, uint8_t **outliers)
{
    for (int c = 0; c < C; ++c) orc_pose_optimization(&pbs[c], &res[c], outliers[c]);
}

/* ===================================================================================================
 * Optimizer::OptimizeSim3 (src/Optimizer.cpp:1054-1249) -- SURVEY 8(f) N1, second half.
 *   Thirdparty/g2o/g2o/types/sim3.h:69-143 (exp), :145-147 (map), :232-235 (inverse), :263-269 (product)
 *   Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h:60-69 (oplus, _fix_scale), :74-88, :138-167 (edges)
 *   Thirdparty/g2o/g2o/core/base_binary_edge.hpp:131-205 (numeric Jacobians, delta = 1e-9), :55-115 (quadratic form)
 * The map points are fixed vertices: one free 7-dof vertex.  g2o::Sim3 never normalises its quaternion.
 * Both classifications read the edges' stored errors (those of the last LM trial, accepted or not).
 * PARITY UNPINNED (Eigen arithmetic restated by formula; accumulation with explicit fma as in PoseOptimization).
 * =================================================================================================== */
typedef struct { double q[4]; double t[3]; double s; } sim3_t;

static void sim3_exp(const double x[7], sim3_t *E)
{
    const double w0 = x[0], w1 = x[1], w2 = x[2];
    const double sigma = x[6];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double O2[9], R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    const double s = exp(sigma);
    const double eps = 0.00001;
    double A, B, C;
    if (fabs(sigma) < eps) {
        C = 1.0;
        if (theta < eps) {
            A = 1.0 / 2.0;
            B = 1.0 / 6.0;
            for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i];
        } else {
            const double theta2 = theta * theta;
            A = (1.0 - cos(theta)) / theta2;
            B = (theta - sin(theta)) / (theta2 * theta);
            const double a = sin(theta) / theta, b = (1.0 - cos(theta)) / (theta * theta);
            for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
        }
    } else {
        C = (s - 1.0) / sigma;
        if (theta < eps) {
            const double sigma2 = sigma * sigma;
            A = ((sigma - 1.0) * s + 1.0) / sigma2;
            B = ((0.5 * sigma2 - sigma + 1.0) * s) / (sigma2 * sigma);
            for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i];
        } else {
            const double ra = sin(theta) / theta, rb = (1.0 - cos(theta)) / (theta * theta);
            for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + ra * O[i] + rb * O2[i];
            const double a = s * sin(theta);
            const double b = s * cos(theta);
            const double theta2 = theta * theta;
            const double sigma2 = sigma * sigma;
            const double c = theta2 + sigma2;
            A = (a * sigma + (1.0 - b) * theta) / (theta * c);
            B = (C - ((b - 1.0) * sigma + a * theta) / c) * 1.0 / theta2;
        }
    }
    quat_from_rot(R, E->q);
    for (int i = 0; i < 3; ++i) {
        const double W0 = A * O[3 * i] + B * O2[3 * i] + ((i == 0) ? C : 0.0);
        const double W1 = A * O[3 * i + 1] + B * O2[3 * i + 1] + ((i == 1) ? C : 0.0);
        const double W2 = A * O[3 * i + 2] + B * O2[3 * i + 2] + ((i == 2) ? C : 0.0);
        E->t[i] = W0 * x[3] + W1 * x[4] + W2 * x[5];
    }
    E->s = s;
}

static void sim3_mul(const sim3_t *a, const sim3_t *b, sim3_t *o)
{
    double r[3], q[4];
    quat_rotate(a->q, b->t, r);
    quat_mul(a->q, b->q, q);
    sim3_t out;
    memcpy(out.q, q, sizeof q);
    for (int i = 0; i < 3; ++i) out.t[i] = a->s * r[i] + a->t[i];
    out.s = a->s * b->s;
    *o = out;
}

static void sim3_inverse(const sim3_t *a, sim3_t *o)
{
    o->q[0] = a->q[0]; o->q[1] = -a->q[1]; o->q[2] = -a->q[2]; o->q[3] = -a->q[3];
    const double m = -1.0 / a->s;
    const double v[3] = {m * a->t[0], m * a->t[1], m * a->t[2]};
    quat_rotate(o->q, v, o->t);
    o->s = 1.0 / a->s;
}

static void sim3_edge_err(const sim3_t *S, const float *X, const float *obs, const double K[4], double e[2])
{
    const double v[3] = {(double)X[0], (double)X[1], (double)X[2]};
    double r[3];
    quat_rotate(S->q, v, r);
    const double p0 = S->s * r[0] + S->t[0], p1 = S->s * r[1] + S->t[1], p2 = S->s * r[2] + S->t[2];
    e[0] = (double)obs[0] - ((p0 / p2) * K[0] + K[2]);
    e[1] = (double)obs[1] - ((p1 / p2) * K[1] + K[3]);
}

typedef struct {
    const orc_sim3opt_problem *pb;
    const uint8_t *removed;
    double K1[4], K2[4], delta, dsqr;
    double *err;             /* [n][4]: e12 (2), e21 (2) -- the edges' stored _error */
} sgraph_t;

static double s_huber(double c, double delta, double dsqr) { return (c <= dsqr) ? c : 2.0 * sqrt(c) * delta - dsqr; }

static double s_active_chi2(const sgraph_t *g, const sim3_t *S)
{
    const orc_sim3opt_problem *pb = g->pb;
    sim3_t Si;
    sim3_inverse(S, &Si);
    double sum = 0.0;
    for (int i = 0; i < pb->n; ++i) {
        if (g->removed[i]) continue;
        double *e = g->err + 4 * i;
        sim3_edge_err(S, pb->x2c + 3 * i, pb->obs1 + 2 * i, g->K1, e);
        const double s1 = (double)pb->inv_sigma2_1[i];
        sum += s_huber(e[0] * (s1 * e[0]) + e[1] * (s1 * e[1]), g->delta, g->dsqr);
        sim3_edge_err(&Si, pb->x1c + 3 * i, pb->obs2 + 2 * i, g->K2, e + 2);
        const double s2 = (double)pb->inv_sigma2_2[i];
        sum += s_huber(e[2] * (s2 * e[2]) + e[3] * (s2 * e[3]), g->delta, g->dsqr);
    }
    return sum;
}

static void s_build_system(const sgraph_t *g, const sim3_t *S, double H[49], double b[7])
{
    const orc_sim3opt_problem *pb = g->pb;
    sim3_t fwd[15], inv[15];
    fwd[0] = *S;
    for (int p = 1; p < 15; ++p) {
        double x[7] = {0, 0, 0, 0, 0, 0, 0};
        const int d = (p - 1) >> 1;
        x[d] = ((p - 1) & 1) ? -1e-9 : 1e-9;
        if (pb->fix_scale) x[6] = 0.0;
        sim3_t E;
        sim3_exp(x, &E);
        sim3_mul(&E, S, &fwd[p]);
    }
    for (int p = 0; p < 15; ++p) sim3_inverse(&fwd[p], &inv[p]);
    memset(H, 0, 49 * sizeof(double));
    memset(b, 0, 7 * sizeof(double));
    const double scalar = 1.0 / (2.0 * 1e-9);
    for (int i = 0; i < pb->n; ++i) {
        if (g->removed[i]) continue;
        for (int side = 0; side < 2; ++side) {
            const float *X = side ? pb->x1c + 3 * i : pb->x2c + 3 * i;
            const float *ob = side ? pb->obs2 + 2 * i : pb->obs1 + 2 * i;
            const double *K = side ? g->K2 : g->K1;
            const sim3_t *P = side ? inv : fwd;
            const double s = (double)(side ? pb->inv_sigma2_2[i] : pb->inv_sigma2_1[i]);
            const double *e = g->err + 4 * i + 2 * side;
            double J0[7], J1[7];
            for (int d = 0; d < 7; ++d) {
                double ea[2], eb[2];
                sim3_edge_err(&P[1 + 2 * d], X, ob, K, ea);
                sim3_edge_err(&P[2 + 2 * d], X, ob, K, eb);
                J0[d] = scalar * (ea[0] - eb[0]);
                J1[d] = scalar * (ea[1] - eb[1]);
            }
            const double c = e[0] * (s * e[0]) + e[1] * (s * e[1]);
            double rho1 = 1.0;
            if (c > g->dsqr) rho1 = g->delta / sqrt(c);
            const double wo = rho1 * s;
            const double we0 = s * e[0], we1 = s * e[1];
            for (int a = 0; a < 7; ++a) {
                const double be = fma(J1[a], we1, J0[a] * we0);
                b[a] -= rho1 * be;
                for (int c2 = a; c2 < 7; ++c2) {
                    double h = H[7 * a + c2];
                    h = fma(J0[a], wo * J0[c2], h);
                    h = fma(J1[a], wo * J1[c2], h);
                    H[7 * a + c2] = h;
                }
            }
        }
    }
    for (int a = 0; a < 7; ++a)
        for (int c2 = 0; c2 < a; ++c2) H[7 * a + c2] = H[7 * c2 + a];
}

static int ldlt7(const double Hin[49], const double b[7], double x[7])
{
    double L[49], d[7], r[7], y[7];
    memcpy(L, Hin, sizeof L);
    for (int j = 0; j < 7; ++j) {
        double dj = L[8 * j];
        for (int k = 0; k < j; ++k) dj -= L[7 * j + k] * L[7 * j + k] * d[k];
        if (!(dj > 0.0)) return 0;
        d[j] = dj;
        r[j] = 1.0 / dj;
        for (int i = j + 1; i < 7; ++i) {
            double v = L[7 * i + j];
            for (int k = 0; k < j; ++k) v -= L[7 * i + k] * L[7 * j + k] * d[k];
            L[7 * i + j] = v * r[j];
        }
    }
    for (int i = 0; i < 7; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= L[7 * i + k] * y[k];
        y[i] = v;
    }
    for (int i = 0; i < 7; ++i) y[i] *= r[i];
    for (int i = 6; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < 7; ++k) v -= L[7 * k + i] * x[k];
        x[i] = v;
    }
    return 1;
}

static void s_optimize(const sgraph_t *g, sim3_t *S, int iterations, int *n_iter, int *n_trials)
{
    double lambda = 0.0, ni = 2.0;
    int nBad = 0;
    double x[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iterations; ++it) {
        double currentChi = s_active_chi2(g, S);
        const double iniChi = currentChi;
        double H[49], b[7];
        s_build_system(g, S, H, b);
        ++*n_iter;
        if (it == 0) {
            double maxDiag = 0.0;
            for (int j = 0; j < 7; ++j) maxDiag = fmax(fabs(H[8 * j]), maxDiag);
            lambda = 1e-5 * maxDiag;
            ni = 2.0;
            nBad = 0;
        }
        double rho = 0.0;
        int qmax = 0;
        do {
            const sim3_t backup = *S;
            double Hl[49];
            memcpy(Hl, H, sizeof Hl);
            for (int j = 0; j < 7; ++j) Hl[8 * j] += lambda;
            const int ok2 = ldlt7(Hl, b, x);
            ++*n_trials;
            if (g->pb->fix_scale) x[6] = 0.0;
            sim3_t E, Sn;
            sim3_exp(x, &E);
            sim3_mul(&E, S, &Sn);
            *S = Sn;
            double tempChi = s_active_chi2(g, S);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0.0;
            for (int j = 0; j < 7; ++j) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0.0 && isfinite(tempChi)) {
                const double d = 2.0 * rho - 1.0;
                double alpha = 1.0 - d * d * d;
                alpha = fmin(alpha, 2.0 / 3.0);
                lambda *= fmax(1.0 / 3.0, alpha);
                ni = 2.0;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2.0;
                *S = backup;
            }
            ++qmax;
        } while (rho < 0.0 && qmax < 10);
        if (qmax == 10 || rho == 0.0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) break;
    }
}

static int s_classify(const sgraph_t *g, uint8_t *removed, int *kept)
{
    const orc_sim3opt_problem *pb = g->pb;
    int bad = 0, good = 0;
    const double th2 = (double)pb->th2;
    for (int i = 0; i < pb->n; ++i) {
        if (removed[i]) continue;
        const double *e = g->err + 4 * i;
        const double s1 = (double)pb->inv_sigma2_1[i], s2 = (double)pb->inv_sigma2_2[i];
        const double c12 = e[0] * (s1 * e[0]) + e[1] * (s1 * e[1]);
        const double c21 = e[2] * (s2 * e[2]) + e[3] * (s2 * e[3]);
        if (c12 > th2 || c21 > th2) { removed[i] = 1; ++bad; }
        else ++good;
    }
    *kept = good;
    return bad;
}

void orc_optimize_sim3(const orc_sim3opt_problem *pb, orc_sim3opt_result *res, uint8_t *removed)
{
    memset(res, 0, sizeof *res);
    const int n = pb->n;
    for (int i = 0; i < n; ++i) removed[i] = 0;
    sim3_t S;
    {
        double R[9];
        for (int i = 0; i < 9; ++i) R[i] = (double)pb->S12[i];
        quat_from_rot(R, S.q);
        for (int i = 0; i < 3; ++i) S.t[i] = (double)pb->S12[9 + i];
        S.s = (double)pb->S12[12];
    }
    const sim3_t S0 = S;
    int second = 0, nBad = 0, nIn = 0;
    if (n > 0) {
        sgraph_t g;
        g.pb = pb; g.removed = removed;
        for (int k = 0; k < 4; ++k) { g.K1[k] = (double)pb->K1[k]; g.K2[k] = (double)pb->K2[k]; }
        g.delta = (double)sqrtf(pb->th2);
        g.dsqr = g.delta * g.delta;
        g.err = (double *)calloc((size_t)n * 4, sizeof(double));
        s_optimize(&g, &S, 5, &res->iterations, &res->trials);
        int kept = 0;
        nBad = s_classify(&g, removed, &kept);
        if (n - nBad >= 10) {
            second = 1;
            s_optimize(&g, &S, nBad > 0 ? 10 : 5, &res->iterations, &res->trials);
            s_classify(&g, removed, &kept);
            nIn = kept;
        }
        free(g.err);
    }
    const sim3_t *F = second ? &S : &S0;
    res->n_inliers = second ? nIn : 0;
    res->n_bad = nBad;
    res->optimized = second;
    quat_to_rot(F->q, res->R);
    for (int i = 0; i < 3; ++i) res->t[i] = F->t[i];
    res->s = F->s;
    for (int i = 0; i < 4; ++i) res->q[i] = F->q[i];
}
