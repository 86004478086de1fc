// mlpnp.cuh -- MLPnP building blocks (device), following reference src/MLPnPsolver.cpp
// (an ORB-SLAM3 transplant the reference neither compiles nor calls: spec-by-source).
//
// Like epnp.cuh the pieces serve the thread-per-hypothesis minimal solve (n = 6) and the
// CTA-cooperative n-point refine, whose sums are evaluated entry-parallel in index order.
// FP64, -fmad=false.  sin/cos/acos/pow come from the CUDA math library (<= 2 ulp from glibc):
// MLPnP hypotheses therefore agree with the CPU checker to ~1e-13, not bit for bit; n = 6
// MLPnP is well conditioned, the parity tests state the tolerance they use.
#pragma once
#include "linalg.cuh"

namespace rsac {

__host__ __device__ inline void cross3(const double* a, const double* b, double* o)
{
    o[0] = rfma(a[1], b[2], -(a[2] * b[1]));
    o[1] = rfma(a[2], b[0], -(a[0] * b[2]));
    o[2] = rfma(a[0], b[1], -(a[1] * b[0]));
}

__host__ __device__ inline double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// MLPnPsolver::rodrigues2rot (MLPnPsolver.cpp:625-640)
__host__ __device__ inline void rodrigues2rot(const double* w, double* R)
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

// MLPnPsolver::rot2rodrigues (MLPnPsolver.cpp:642-657)
__host__ __device__ inline void rot2rodrigues(const double* R, double* w)
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

// residual pair and 2x6 Jacobian of one observation: r = N^T q/|q|, q = R(w) p + t
// (mlpnp_residuals_and_jacs :736-742; mlpnpJacs :773-1020 restated as the direct derivative
// of the Rodrigues formula, singular at w = 0 like the reference's generated code)
__host__ __device__ inline void mlpnp_res_jac(const double* p, const double* nr, const double* ns,
                                              const double* w, const double* t, double* r, double* J)
{
    double R[9];
    rodrigues2rot(w, R);
    double q[3];
    for (int i = 0; i < 3; ++i) q[i] = (R[i * 3 + 0] * p[0] + R[i * 3 + 1] * p[1] + R[i * 3 + 2] * p[2]) + t[i];
    const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double qh[3] = {q[0] / qn, q[1] / qn, q[2] / qn};
    r[0] = nr[0] * qh[0] + nr[1] * qh[1] + nr[2] * qh[2];
    r[1] = ns[0] * qh[0] + ns[1] * qh[1] + ns[2] * qh[2];
    double g[2][3];
    for (int i = 0; i < 3; ++i) {
        g[0][i] = (nr[i] - r[0] * qh[i]) / qn;
        g[1][i] = (ns[i] - r[1] * qh[i]) / qn;
    }
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    const double sn = sin(th), cs = cos(th);
    const double a = sn / th;
    const double b = (1.0 - cs) / th2;
    const double da = (th * cs - sn) / (th2 * th);
    const double db = (th * sn - 2.0 * (1.0 - cs)) / (th2 * th2);
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

// null space of a bearing: last two columns of the Householder reflector mapping f onto e1
// (JacobiSVD with Householder-QR preconditioner of f^T, MLPnPsolver.cpp:336-338); N is 3x2
__host__ __device__ inline void mlpnp_nullspace(const double* f, double* N)
{
    const double nf = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    double v[3] = {f[0], f[1], f[2]};
    v[0] = (f[0] >= 0.0) ? f[0] + nf : f[0] - nf;
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double beta = 2.0 / vv;
    for (int r = 0; r < 3; ++r)
        for (int c = 1; c < 3; ++c) N[r * 2 + (c - 1)] = ((r == c) ? 1.0 : 0.0) - beta * v[r] * v[c];
}

// bearing of a keypoint as the constructor builds it (MLPnPsolver.cpp:33-37): f32, widened, not normalised
__host__ __device__ inline void mlpnp_bearing(float u, float v, const float* K, double* f)
{
    const float x = (u - K[2]) / K[0];
    const float y = (v - K[3]) / K[1];
    f[0] = (double)x; f[1] = (double)y; f[2] = (double)1.f;
}

// P_i = (N^T Sigma N)^-1 (MLPnPsolver.cpp:375-388), 2x2 row-major
__host__ __device__ inline void mlpnp_weight(const double* N, const double* S, double* P)
{
    double SN[6];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 2; ++c)
            SN[r * 2 + c] = S[r * 3 + 0] * N[0 * 2 + c] + S[r * 3 + 1] * N[1 * 2 + c] + S[r * 3 + 2] * N[2 * 2 + c];
    double T[4];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c)
            T[r * 2 + c] = N[0 * 2 + r] * SN[0 * 2 + c] + N[1 * 2 + r] * SN[1 * 2 + c] + N[2 * 2 + r] * SN[2 * 2 + c];
    const double det = rfma(T[0], T[3], -(T[1] * T[2]));
    const double id = 1.0 / det;
    P[0] = T[3] * id;
    P[1] = -T[1] * id;
    P[2] = -T[2] * id;
    P[3] = T[0] * id;
}

// the two design-matrix rows of one observation (MLPnPsolver.cpp:404-476) and their weighted copies
__host__ __device__ inline void mlpnp_rows(const double* N, const double* pt, const double* P /* or nullptr */, bool planar,
                                           double* a0, double* a1, double* w0, double* w1)
{
    const int cols = planar ? 9 : 12;
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
    if (P) {
        for (int c = 0; c < cols; ++c) {
            w0[c] = rfma(P[0], a0[c], P[1] * a1[c]);
            w1[c] = rfma(P[2], a0[c], P[3] * a1[c]);
        }
    } else {
        for (int c = 0; c < cols; ++c) { w0[c] = a0[c]; w1[c] = a1[c]; }
    }
}

// From the smallest eigenvector of A^T P A to the initial (R, t) (MLPnPsolver.cpp:495-602).
// p6 / f6: the first six points / bearings (used by the +-t and 4-candidate tests).
__host__ __device__ inline void mlpnp_recover(const double* result1, bool planar, const double* eigenRot,
                                              const double* p6, const double* f6, double* Rout, double* tout)
{
    if (planar) {
        double tmp[9] = {0.0, result1[0], result1[1], 0.0, result1[2], result1[3], 0.0, result1[4], result1[5]};
        const double c1[3] = {tmp[1], tmp[4], tmp[7]}, c2[3] = {tmp[2], tmp[5], tmp[8]};
        double c0[3];
        cross3(c1, c2, c0);
        tmp[0] = c0[0]; tmp[3] = c0[1]; tmp[6] = c0[2];
        double tt[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) tt[r * 3 + c] = tmp[c * 3 + r];
        const double tc1[3] = {tt[1], tt[4], tt[7]}, tc2[3] = {tt[2], tt[5], tt[8]};
        const double scale = 1.0 / sqrt(fabs(norm3(tc1) * norm3(tc2)));
        double Rout1[9];
        polar3(tt, Rout1);
        if (det3(Rout1) < 0)
            for (int i = 0; i < 9; ++i) Rout1[i] *= -1.0;
        double Rb[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                Rb[r * 3 + c] = eigenRot[0 * 3 + r] * Rout1[0 * 3 + c] + eigenRot[1 * 3 + r] * Rout1[1 * 3 + c] + eigenRot[2 * 3 + r] * Rout1[2 * 3 + c];
        const double t[3] = {scale * result1[6], scale * result1[7], scale * result1[8]};
        double Rc[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rc[r * 3 + c] = -Rb[c * 3 + r];
        if (det3(Rc) < 0.0) { Rc[2] *= -1; Rc[5] *= -1; Rc[8] *= -1; }
        double Rs[2][9];
        for (int r = 0; r < 3; ++r) {
            Rs[0][r * 3 + 0] = Rc[r * 3 + 0]; Rs[0][r * 3 + 1] = Rc[r * 3 + 1]; Rs[0][r * 3 + 2] = Rc[r * 3 + 2];
            Rs[1][r * 3 + 0] = -Rc[r * 3 + 0]; Rs[1][r * 3 + 1] = -Rc[r * 3 + 1]; Rs[1][r * 3 + 2] = Rc[r * 3 + 2];
        }
        double best = 0.0;
        int bi = -1;
        for (int k = 0; k < 4; ++k) {
            const double* Rk = Rs[k / 2];
            const double sg = (k % 2 == 0) ? 1.0 : -1.0;
            double norms = 0.0;
            for (int q = 0; q < 6; ++q) {
                const double* pp = p6 + 3 * q;
                double v[3];
                for (int r = 0; r < 3; ++r) v[r] = (Rk[r * 3 + 0] * pp[0] + Rk[r * 3 + 1] * pp[1] + Rk[r * 3 + 2] * pp[2]) + sg * t[r];
                const double vn = norm3(v);
                norms += (1.0 - ((v[0] / vn) * f6[3 * q + 0] + (v[1] / vn) * f6[3 * q + 1] + (v[2] / vn) * f6[3 * q + 2]));
            }
            if (bi < 0 || norms < best) { best = norms; bi = k; }
        }
        for (int i = 0; i < 9; ++i) Rout[i] = Rs[bi / 2][i];
        for (int r = 0; r < 3; ++r) tout[r] = (bi % 2 == 0) ? t[r] : -t[r];
    } else {
        double tmp[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) tmp[r * 3 + c] = result1[3 * c + r];
        const double c0[3] = {tmp[0], tmp[3], tmp[6]}, c1[3] = {tmp[1], tmp[4], tmp[7]}, c2[3] = {tmp[2], tmp[5], tmp[8]};
        const double scale = 1.0 / pow(fabs(norm3(c0) * norm3(c1) * norm3(c2)), 1.0 / 3.0);
        double Rp[9];
        polar3(tmp, Rp);
        if (det3(Rp) < 0)
            for (int i = 0; i < 9; ++i) Rp[i] *= -1.0;
        const double ts[3] = {scale * result1[9], scale * result1[10], scale * result1[11]};
        double t0[3];
        for (int r = 0; r < 3; ++r) t0[r] = Rp[r * 3 + 0] * ts[0] + Rp[r * 3 + 1] * ts[1] + Rp[r * 3 + 2] * ts[2];
        double Rinv[9];
        inv3(Rp, Rinv);
        double err[2], tinv[2][3];
        for (int s = 0; s < 2; ++s) {
            const double sg = (s == 0) ? 1.0 : -1.0;
            for (int r = 0; r < 3; ++r)
                tinv[s][r] = -(Rinv[r * 3 + 0] * (sg * t0[0]) + Rinv[r * 3 + 1] * (sg * t0[1]) + Rinv[r * 3 + 2] * (sg * t0[2]));
            err[s] = 0.0;
            for (int q = 0; q < 6; ++q) {
                const double* pp = p6 + 3 * q;
                double v[3];
                for (int r = 0; r < 3; ++r) v[r] = (Rinv[r * 3 + 0] * pp[0] + Rinv[r * 3 + 1] * pp[1] + Rinv[r * 3 + 2] * pp[2]) + tinv[s][r];
                const double vn = norm3(v);
                err[s] += (1.0 - ((v[0] / vn) * f6[3 * q + 0] + (v[1] / vn) * f6[3 * q + 1] + (v[2] / vn) * f6[3 * q + 2]));
            }
        }
        const int pick = (err[0] < err[1]) ? 0 : 1;
        for (int r = 0; r < 3; ++r) tout[r] = tinv[pick][r];
        for (int i = 0; i < 9; ++i) Rout[i] = Rinv[i];
    }
}

// one Gauss-Newton decision after the 6x6 solve (MLPnPsolver.cpp:709-718).
// returns 0 = abort without update, 1 = update and continue, 2 = update and stop
__host__ __device__ inline int mlpnp_gn_decide(const double* dx, double max_dl)
{
    double mx = 0.0, mn = INFINITY;
    for (int c = 0; c < 6; ++c) {
        const double v = fabs(dx[c]);
        if (v > mx) mx = v;
        if (v < mn) mn = v;
    }
    if (mx > 5.0 || mn > 1.0) return 0;
    return (max_dl < 1e-5) ? 2 : 1;
}

// Whole MLPnPsolver::computePose (:321-623) for NPTS thread-private observations.
// f, p: NPTS x 3 (double); cov: NPTS x 9 or nullptr.  Writes R (9) and t (3).
#if defined(__CUDA_ARCH__)
#define RSAC_MLPNP_MARK(i) do { if (clk) clk[i] = clock64(); } while (0)
#else
#define RSAC_MLPNP_MARK(i) ((void)0)
#endif
// clk: optional clock64() stamps of the phases (rsac_debug_mlpnp_clocks).  Measured on B200 (cfg2, hypothesis 0): null spaces +
// weights 11 k cycles, A^T P A 7 k, the 12 x 12 eigen-solve 669 k (72 %: 78 doubles of matrix beside ~100 live doubles of the
// caller under a 255-register cap, 260 KB of unrolled code per sweep), pose recovery 34 k, Gauss-Newton 211 k.  Tried and
// measured slower: the matrix in shared memory, one column per thread, with the step loop rolled (LSU-bound: 998 k), and the
// eigen-solve out of line (1.3 M); the next step is the eigen-solve on six lanes per hypothesis.
template <int NPTS>
__host__ __device__ inline void mlpnp_compute_pose_small(const double* f, const double* p, const double* cov, double* Rres, double* tres, double2* rec,
                                                         long long* clk = nullptr)
{
    RSAC_MLPNP_MARK(0);
    double nulls[NPTS * 6], pts3[NPTS * 3], P[NPTS * 4];
    for (int i = 0; i < NPTS; ++i) mlpnp_nullspace(f + 3 * i, nulls + 6 * i);
    for (int i = 0; i < NPTS * 3; ++i) pts3[i] = p[i];
    double planarTest[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < NPTS; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) planarTest[r * 3 + c] = rfma(p[3 * i + r], p[3 * i + c], planarTest[r * 3 + c]);
    double eigenRot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    bool planar = false;
    if (rank3_fullpiv(planarTest) == 2) {
        planar = true;
        double A[9], w[3], V[9];
        for (int i = 0; i < 9; ++i) A[i] = planarTest[i];
        jacobi_eig<double, 3>(A, w, V);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) eigenRot[r * 3 + c] = V[c * 3 + r];
        for (int i = 0; i < NPTS; ++i) {
            const double* q = p + 3 * i;
            for (int r = 0; r < 3; ++r)
                pts3[3 * i + r] = eigenRot[r * 3 + 0] * q[0] + eigenRot[r * 3 + 1] * q[1] + eigenRot[r * 3 + 2] * q[2];
        }
    }
    if (cov)
        for (int i = 0; i < NPTS; ++i) mlpnp_weight(nulls + 6 * i, cov + 9 * i, P + 4 * i);
    RSAC_MLPNP_MARK(1);

    double result1[12], ev[1];
    if (planar) {
        double AtPA[45];
        for (int i = 0; i < 45; ++i) AtPA[i] = 0.0;
        for (int i = 0; i < NPTS; ++i) {
            double a0[12], a1[12], w0[12], w1[12];
            mlpnp_rows(nulls + 6 * i, pts3 + 3 * i, cov ? P + 4 * i : nullptr, true, a0, a1, w0, w1);
#pragma unroll
            for (int a = 0; a < 9; ++a)
#pragma unroll
                for (int b = a; b < 9; ++b) {
                    AtPA[tri_idx(9, a, b)] = rfma(a0[a], w0[b], AtPA[tri_idx(9, a, b)]);
                    AtPA[tri_idx(9, a, b)] = rfma(a1[a], w1[b], AtPA[tri_idx(9, a, b)]);
                }
        }
        jacobi_lowest<9, 1>(AtPA, ev, result1, rec);
    } else {
        double AtPA[78];
        for (int i = 0; i < 78; ++i) AtPA[i] = 0.0;
        for (int i = 0; i < NPTS; ++i) {
            double a0[12], a1[12], w0[12], w1[12];
            mlpnp_rows(nulls + 6 * i, pts3 + 3 * i, cov ? P + 4 * i : nullptr, false, a0, a1, w0, w1);
#pragma unroll
            for (int a = 0; a < 12; ++a)
#pragma unroll
                for (int b = a; b < 12; ++b) {
                    AtPA[tri_idx(12, a, b)] = rfma(a0[a], w0[b], AtPA[tri_idx(12, a, b)]);
                    AtPA[tri_idx(12, a, b)] = rfma(a1[a], w1[b], AtPA[tri_idx(12, a, b)]);
                }
        }
        RSAC_MLPNP_MARK(2);
        jacobi_lowest<12, 1>(AtPA, ev, result1, rec);
    }
    RSAC_MLPNP_MARK(3);
    double Rout[9], tout[3];
    mlpnp_recover(result1, planar, eigenRot, p, f, Rout, tout);
    RSAC_MLPNP_MARK(4);

    // Gauss-Newton (MLPnPsolver.cpp:607-622, 659-723)
    double x[6];
    rot2rodrigues(Rout, x);
    x[3] = tout[0]; x[4] = tout[1]; x[5] = tout[2];
    double Jall[NPTS * 12];
    int it_cnt = 0;
    while (it_cnt < 5) {
        double A[36], g[6], dx[6];
        for (int i = 0; i < 36; ++i) A[i] = 0.0;
        for (int i = 0; i < 6; ++i) g[i] = 0.0;
        for (int i = 0; i < NPTS; ++i) {
            const double* N = nulls + 6 * i;
            const double nr[3] = {N[0], N[2], N[4]}, ns[3] = {N[1], N[3], N[5]};
            double* J = Jall + 12 * i;
            double r[2];
            mlpnp_res_jac(p + 3 * i, nr, ns, x, x + 3, r, J);
            double W0[6], W1[6];
            if (cov) {
                const double* pp = P + 4 * i;
                for (int c = 0; c < 6; ++c) {
                    W0[c] = rfma(J[c], pp[0], J[6 + c] * pp[2]);
                    W1[c] = rfma(J[c], pp[1], J[6 + c] * pp[3]);
                }
            } else {
                for (int c = 0; c < 6; ++c) { W0[c] = J[c]; W1[c] = J[6 + c]; }
            }
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) {
                    A[a * 6 + b] = rfma(W0[a], J[b], A[a * 6 + b]);
                    A[a * 6 + b] = rfma(W1[a], J[6 + b], A[a * 6 + b]);
                }
                g[a] = rfma(W0[a], r[0], g[a]);
                g[a] = rfma(W1[a], r[1], g[a]);
            }
        }
        ldlt6_solve(A, g, dx);
        double mdl = 0.0;
        for (int i = 0; i < NPTS; ++i)
            for (int k = 0; k < 2; ++k) {
                const double* J = Jall + 12 * i + 6 * k;
                const double dl = J[0] * dx[0] + J[1] * dx[1] + J[2] * dx[2] + J[3] * dx[3] + J[4] * dx[4] + J[5] * dx[5];
                if (fabs(dl) > mdl) mdl = fabs(dl);
            }
        const int dec = mlpnp_gn_decide(dx, mdl);
        if (dec == 0) break;
        for (int c = 0; c < 6; ++c) x[c] = x[c] - dx[c];
        if (dec == 2) break;
        ++it_cnt;
    }
    RSAC_MLPNP_MARK(5);
    if (clk) clk[7] = it_cnt;
    rodrigues2rot(x, Rres);
    tres[0] = x[3]; tres[1] = x[4]; tres[2] = x[5];
    RSAC_MLPNP_MARK(6);
}

}  // namespace rsac
