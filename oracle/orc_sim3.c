/*
 * oracle/orc_sim3.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Plain-C restatement of the reference's src/Sim3Solver.cpp (Horn 1987 on three
 * pairs + two-way reprojection scoring, all f32).  The reference has no scale
 * step (Sim3Solver.cpp:250); fix_scale=0 adds Horn's scale as
 * upstream ORB-SLAM2 does -- PARITY UNPINNED for that variant.  The fixed-scale
 * path (the reference's) equals the reference's own Sim3Solver.cpp, compiled
 * with stand-in Eigen headers, bit for bit (oracle/_ref,
 * tests/test_cpu_reference_build.py); the 4x4 eigen-solve is orc_linalg.c's on
 * both sides (Eigen's own rounding unpinned, orc.h).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* Sim3Solver::SetRansacParameters (Sim3Solver.cpp:87-111) */
void orc_sim3_ransac_setup(int N, double prob, int min_inliers, int max_its_in, int *max_its)
{
    const float epsilon = (float)min_inliers / N;                       /* :98 */
    int nIterations;
    if (min_inliers == N)
        nIterations = 1;
    else
        nIterations = (int)ceil(log(1 - prob) / log(1 - pow(epsilon, 3)));   /* :106 */
    int its = nIterations < max_its_in ? nIterations : max_its_in;
    *max_its = its > 1 ? its : 1;
}

/* Sim3Solver::ComputeCentroid (Sim3Solver.cpp:186-194); P holds the 3 points as
 * rows here (P[k*3+c] = component c of point k), i.e. the transpose of the
 * reference's column layout. */
static void compute_centroid(const float P[9], float Pr[9], float C[3])
{
    for (int c = 0; c < 3; ++c) {
        C[c] = P[0 * 3 + c] + P[1 * 3 + c] + P[2 * 3 + c];   /* rowwise().sum() */
        C[c] = C[c] / 3.f;
    }
    for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) Pr[k * 3 + c] = P[k * 3 + c] - C[c];
}

/* Eigen::Quaternionf::toRotationMatrix (Sim3Solver.cpp:248) */
static void quat_to_rot_f(float w, float x, float y, float z, float R[9])
{
    const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0f - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
    R[3] = txy + twz;          R[4] = 1.0f - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = 1.0f - (txx + tyy);
}

/* Sim3Solver::ComputeSim3 (Sim3Solver.cpp:196-266) */
void orc_sim3_compute(const float P1[9], const float P2[9], int fix_scale,
                      float R12[9], float t12[3], float *s12)
{
    float Pr1[9], Pr2[9], O1[3], O2[3];
    compute_centroid(P1, Pr1, O1);                                      /* :207 */
    compute_centroid(P2, Pr2, O2);
    /* :212 M = Pr2 * Pr1^T  (reference layout: columns are points) => M(r,c) = sum_k Pr2(r,k) Pr1(c,k) */
    float M[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            M[r * 3 + c] = Pr2[0 * 3 + r] * Pr1[0 * 3 + c] + Pr2[1 * 3 + r] * Pr1[1 * 3 + c] + Pr2[2 * 3 + r] * Pr1[2 * 3 + c];
    /* :216-234 */
    const float N11 = M[0] + M[4] + M[8];
    const float N12 = M[5] - M[7];
    const float N13 = M[6] - M[2];
    const float N14 = M[1] - M[3];
    const float N22 = M[0] - M[4] - M[8];
    const float N23 = M[1] + M[3];
    const float N24 = M[6] + M[2];
    const float N33 = -M[0] + M[4] - M[8];
    const float N34 = M[5] + M[7];
    const float N44 = -M[0] - M[4] + M[8];
    float N[16] = {N11, N12, N13, N14, N12, N22, N23, N24, N13, N23, N33, N34, N14, N24, N34, N44};
    float w[4], V[16];
    orc_jacobi_eig_f(4, N, w, V);                                       /* :238-239 */
    /* :241-246 last eigenvector used directly as (w,x,y,z) */
    quat_to_rot_f(V[0 * 4 + 3], V[1 * 4 + 3], V[2 * 4 + 3], V[3 * 4 + 3], R12);   /* :248 */

    float s = 1.0f;
    if (!fix_scale) {
        /* Horn 1987 / upstream ORB-SLAM2 Sim3Solver::ComputeSim3 step 5-6 (absent
         * from the reference, Q7): s = <Pr1, R Pr2> / |R Pr2|^2, accumulated in double */
        double nom = 0.0, den = 0.0;
        for (int k = 0; k < 3; ++k)
            for (int r = 0; r < 3; ++r) {
                const float p3 = R12[r * 3 + 0] * Pr2[k * 3 + 0] + R12[r * 3 + 1] * Pr2[k * 3 + 1] + R12[r * 3 + 2] * Pr2[k * 3 + 2];
                nom += (double)Pr1[k * 3 + r] * (double)p3;
                den += (double)(p3 * p3);
            }
        s = (float)(nom / den);
    }
    *s12 = s;
    /* :253 t12 = O1 - s*R12*O2 */
    for (int r = 0; r < 3; ++r) {
        const float sr0 = s * R12[r * 3 + 0], sr1 = s * R12[r * 3 + 1], sr2 = s * R12[r * 3 + 2];
        const float ro = fix_scale ? (R12[r * 3 + 0] * O2[0] + R12[r * 3 + 1] * O2[1] + R12[r * 3 + 2] * O2[2])
                                   : (sr0 * O2[0] + sr1 * O2[1] + sr2 * O2[2]);
        t12[r] = O1[r] - ro;
    }
}

/* Sim3Solver::Project (Sim3Solver.cpp:306-327), linear part A (= s*R), all f32 */
static void project(const float A[9], const float t[3], const float K[4], const float X[3], float uv[2])
{
    const float x = (A[0] * X[0] + A[1] * X[1] + A[2] * X[2]) + t[0];
    const float y = (A[3] * X[0] + A[4] * X[1] + A[5] * X[2]) + t[1];
    const float z = (A[6] * X[0] + A[7] * X[1] + A[8] * X[2]) + t[2];
    const float invz = 1 / z;
    const float xn = x * invz, yn = y * invz;
    uv[0] = K[0] * xn + K[2];
    uv[1] = K[1] * yn + K[3];
}

/* Sim3Solver::FromCameraToImage (Sim3Solver.cpp:329-347) */
static void from_camera_to_image(const float K[4], const float X[3], float uv[2])
{
    const float invz = 1 / X[2];
    const float xn = X[0] * invz, yn = X[1] * invz;
    uv[0] = K[0] * xn + K[2];
    uv[1] = K[1] * yn + K[3];
}

/* T12 = [sR t]; T21 = T12^-1 = [R^T/s, -(R^T/s) t]  (Sim3Solver.cpp:259-265; for
 * s = 1 this is Isometry3f::inverse(): linear^T, -(linear^T * t)) */
static void make_T(const float R12[9], const float t12[3], float s, float A12[9], float A21[9], float t21[3])
{
    const float inv_s = 1.0f / s;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            A12[r * 3 + c] = (s == 1.0f) ? R12[r * 3 + c] : s * R12[r * 3 + c];
            A21[r * 3 + c] = (s == 1.0f) ? R12[c * 3 + r] : inv_s * R12[c * 3 + r];
        }
    for (int r = 0; r < 3; ++r)
        t21[r] = -(A21[r * 3 + 0] * t12[0] + A21[r * 3 + 1] * t12[1] + A21[r * 3 + 2] * t12[2]);
}

/* Sim3Solver::CheckInliers (Sim3Solver.cpp:269-293).  Thresholds are
 * size_t(9.210*sigma2) (Sim3Solver.cpp:51-52, Sim3Solver.hpp:55-56, Q4) and
 * are converted to float by the comparison. */
int orc_sim3_check_inliers(const orc_sim3_problem *pb, const float R12[9], const float t12[3],
                           float s12, uint8_t *mask, float *err)
{
    float A12[9], A21[9], t21[3];
    make_T(R12, t12, s12, A12, A21, t21);
    int cnt = 0;
    for (int i = 0; i < pb->n; ++i) {
        float p1im1[2], p2im2[2], p2im1[2], p1im2[2];
        from_camera_to_image(pb->K1, pb->x1c + 3 * i, p1im1);           /* ctor :81-82 */
        from_camera_to_image(pb->K2, pb->x2c + 3 * i, p2im2);
        project(A12, t12, pb->K1, pb->x2c + 3 * i, p2im1);              /* :272 */
        project(A21, t21, pb->K2, pb->x1c + 3 * i, p1im2);              /* :273 */
        const float d1x = p1im1[0] - p2im1[0], d1y = p1im1[1] - p2im1[1];
        const float d2x = p1im2[0] - p2im2[0], d2y = p1im2[1] - p2im2[1];
        const float err1 = d1x * d1x + d1y * d1y;
        const float err2 = d2x * d2x + d2y * d2y;
        const float thr1 = (float)(size_t)(9.210 * pb->sigma2_1[i]);
        const float thr2 = (float)(size_t)(9.210 * pb->sigma2_2[i]);
        const int in = (err1 < thr1) && (err2 < thr2);                   /* :285 */
        if (mask) mask[i] = (uint8_t)in;
        if (err) { err[2 * i] = err1; err[2 * i + 1] = err2; }
        cnt += in;
    }
    return cnt;
}

/* Sim3Solver::iterate (Sim3Solver.cpp:113-178) called until true or bNoMore.
 * `&&` loop: outcome = first hypothesis with cnt > minInliers (strict, :163);
 * best = last arg-max (`>=`, :155). */
void orc_sim3_ransac(const orc_sim3_problem *pb, double prob, int min_inliers, int max_its_in,
                     const uint32_t *table, int flags, orc_result *res, uint8_t *mask,
                     int *hyp_counts, float *hyp_pose)
{
    const int N = pb->n;
    int H;
    orc_sim3_ransac_setup(N, prob, min_inliers, max_its_in, &H);
    const int exhaustive = (flags & ORC_FLAG_EXHAUSTIVE) != 0;

    memset(res, 0, sizeof(*res));
    res->best_hyp = -1;
    for (int i = 0; i < 16; ++i) res->T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    res->scale = 1.0f;
    if (mask) memset(mask, 0, (size_t)N);
    if (N < min_inliers) {                                              /* :119-123 */
        res->no_more = 1;
        return;
    }
    uint8_t *cur = (uint8_t *)malloc((size_t)(N > 0 ? N : 1));
    int nBest = 0, done = 0;
    for (int h = 0; h < H; ++h) {                                       /* :131 */
        float P1[9], P2[9], R[9], t[3], s;
        for (int i = 0; i < 3; ++i) {                                   /* :139-149 */
            const uint32_t idx = table[(size_t)h * 3 + i];
            memcpy(P1 + 3 * i, pb->x1c + 3 * idx, 3 * sizeof(float));
            memcpy(P2 + 3 * i, pb->x2c + 3 * idx, 3 * sizeof(float));
        }
        orc_sim3_compute(P1, P2, pb->fix_scale, R, t, &s);              /* :151 */
        const int cnt = orc_sim3_check_inliers(pb, R, t, s, cur, NULL); /* :153 */
        res->n_hyp = h + 1;
        if (hyp_counts) hyp_counts[h] = cnt;
        if (hyp_pose) { memcpy(hyp_pose + (size_t)h * 13, R, sizeof(R)); memcpy(hyp_pose + (size_t)h * 13 + 9, t, sizeof(t)); hyp_pose[(size_t)h * 13 + 12] = s; }
        if (done) continue;
        if (cnt >= nBest) {                                             /* :155 */
            nBest = cnt;
            res->best_hyp = h;
            res->scale = s;
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) res->T[r * 4 + c] = R[r * 3 + c];   /* mBestRotation (:160), scale kept apart */
                res->T[r * 4 + 3] = t[r];
            }
            if (cnt > min_inliers) {                                    /* :163 */
                res->ok = 1;
                res->n_inliers = cnt;
                if (mask) memcpy(mask, cur, (size_t)N);
                done = 1;
                if (!exhaustive) break;
            }
        }
    }
    res->best_count = nBest;
    if (!done) res->no_more = 1;                                        /* :174-175 */
    free(cur);
}
