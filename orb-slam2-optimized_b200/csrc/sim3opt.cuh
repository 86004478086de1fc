// sim3opt.cuh -- batched Optimizer::OptimizeSim3 (second half): refinement of the Sim3 between two
// keyframes over their matched map points, forward (x1 = S12 X2) and inverse (x2 = S21 X1) reprojection edges, Huber
// kernel, 5 Levenberg-Marquardt iterations, outlier removal, 5 or 10 more iterations on the inliers.
//
// Reference: src/Optimizer.cpp:1054-1249 (called from LoopClosing.cpp:311 with th2 = 10) on top of the vendored g2o:
//   types/sim3.h:69-143 (exponential map), :145-147 (map), :232-235 (inverse), :263-269 (product);
//   types/types_seven_dof_expmap.h:60-69 (oplus, _fix_scale), :74-88 (cam_map1/2), :138-167 (the two edges' computeError);
//   core/base_binary_edge.hpp:131-205 (NUMERIC Jacobians, central differences with delta = 1e-9: the edges do not
//   override linearizeOplus) and :55-115 (quadratic form); core/optimization_algorithm_levenberg.cpp:59-179.
// The map points are fixed vertices, so the system has one 7-dof vertex: same shape as poseopt.cuh.
//
// Mapping: one warp per keyframe pair.  A build pass needs the estimate and its 14 perturbations exp(+-delta e_d) * S
// (g2o recomputes them for every edge; they do not depend on the edge): lanes 0..14 compute one of the 15 transforms and
// its inverse each and park them in shared memory, then the lanes stride over the matches and evaluate both edges
// under all 15 transforms (30 projections per match).  Reductions as in poseopt.cuh (36 accumulators through shared memory).
// Stale errors as in the reference: both classifications use the edges' stored errors, i.e. those of the last LM
// trial of the preceding optimize() even when that trial was rejected (Optimizer.cpp:1181,1216 call chi2() directly).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include "poseopt.cuh"

namespace rsac {

struct Sim3OptMeta {
    int64_t off;       // first match of the pair in the flat arrays
    int32_t n;         // matches (nCorrespondences)
    int32_t fix_scale; // VertexSim3Expmap::_fix_scale (the reference hard-codes true, Optimizer.cpp:1076)
    float th2;
    float K1[4], K2[4];   // fx, fy, cx, cy of both keyframes
    float R12[9], t12[3], s12;   // g2oS12 on entry
};

struct Sim3T { double q[4]; double t[3]; double s; };

namespace so {

using po::po_fma;
using po::quat_from_rot;
using po::quat_rotate;

// Sim3(const Vector7d& update) (sim3.h:69-143)
__host__ __device__ inline void sim3_exp(const double* x, Sim3T& E)
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
    quat_from_rot(R, E.q);                 // Quaterniond(R): not normalised here
    for (int i = 0; i < 3; ++i) {
        const double W0 = A * O[3 * i] + B * O2[3 * i] + ((i == 0) ? C : 0.0);
        const double W1 = A * O[3 * i + 1] + B * O2[3 * i + 1] + ((i == 1) ? C : 0.0);
        const double W2 = A * O[3 * i + 2] + B * O2[3 * i + 2] + ((i == 2) ? C : 0.0);
        E.t[i] = W0 * x[3] + W1 * x[4] + W2 * x[5];
    }
    E.s = s;
}

// Sim3::operator* (sim3.h:263-269): ret.r = r * o.r; ret.t = s (r * o.t) + t; ret.s = s * o.s   (no normalisation)
__host__ __device__ inline void sim3_mul(const Sim3T& a, const Sim3T& b, Sim3T& o)
{
    double r0, r1, r2;
    quat_rotate(a.q, b.t[0], b.t[1], b.t[2], r0, r1, r2);
    const double t0 = a.s * r0 + a.t[0], t1 = a.s * r1 + a.t[1], t2 = a.s * r2 + a.t[2];
    const double q0 = a.q[0] * b.q[0] - a.q[1] * b.q[1] - a.q[2] * b.q[2] - a.q[3] * b.q[3];
    const double q1 = a.q[0] * b.q[1] + a.q[1] * b.q[0] + a.q[2] * b.q[3] - a.q[3] * b.q[2];
    const double q2 = a.q[0] * b.q[2] + a.q[2] * b.q[0] + a.q[3] * b.q[1] - a.q[1] * b.q[3];
    const double q3 = a.q[0] * b.q[3] + a.q[3] * b.q[0] + a.q[1] * b.q[2] - a.q[2] * b.q[1];
    o.q[0] = q0; o.q[1] = q1; o.q[2] = q2; o.q[3] = q3;
    o.t[0] = t0; o.t[1] = t1; o.t[2] = t2;
    o.s = a.s * b.s;
}

// Sim3::inverse (sim3.h:232-235): Sim3(r.conjugate(), r.conjugate() * ((-1/s) t), 1/s)
__host__ __device__ inline void sim3_inverse(const Sim3T& a, Sim3T& o)
{
    o.q[0] = a.q[0]; o.q[1] = -a.q[1]; o.q[2] = -a.q[2]; o.q[3] = -a.q[3];
    const double m = -1.0 / a.s;
    quat_rotate(o.q, m * a.t[0], m * a.t[1], m * a.t[2], o.t[0], o.t[1], o.t[2]);
    o.s = 1.0 / a.s;
}

// obs - cam_map(project(S.map(X)))  (types_seven_dof_expmap.h:74-88,138-167; project = (x/z, y/z))
__host__ __device__ inline void edge_err(const Sim3T& S, double X, double Y, double Z, double u, double v, const double* K,
                                         double& e0, double& e1)
{
    double r0, r1, r2;
    quat_rotate(S.q, X, Y, Z, r0, r1, r2);
    const double p0 = S.s * r0 + S.t[0], p1 = S.s * r1 + S.t[1], p2 = S.s * r2 + S.t[2];
    e0 = u - ((p0 / p2) * K[0] + K[2]);
    e1 = v - ((p1 / p2) * K[1] + K[3]);
}

struct PairView {
    const float* x1;   // [n][3] P3D1c (map point of keyframe 1 in camera 1)
    const float* x2;   // [n][3] P3D2c
    const float* o1;   // [n][2] kpUn1.pt
    const float* o2;   // [n][2] kpUn2.pt
    const float* is1;  // [n] invSigmaSquare1
    const float* is2;  // [n]
    double K1[4], K2[4];
    double delta, dsqr, th2;
    int n;
};

constexpr int kSimPoses = 15;                      // estimate + 14 perturbations
constexpr int kSimAcc = 36;                        // H upper triangle 28 | b 7 | chi2
constexpr int kSimRedStride = 33;
// per pair: W warps x (partials [36][33] + warp totals [36]) + W single-value slots + 2 x 15 transforms
__host__ __device__ constexpr int sim_smem_doubles(int lanes)
{
    return (lanes / 32) * (kSimAcc * kSimRedStride + kSimAcc) + 8 + 2 * kSimPoses * 8;
}
constexpr int kSimSmemDoubles = sim_smem_doubles(32);

// LANES = 32: one warp per pair (__syncwarp only); LANES = 128: four warps per pair, one pair per CTA (see poseopt.cuh)
template <int LANES>
struct SimShared {
    double* red;        // W x ([36][33] + [36]), then 8 single-value slots
    Sim3T* fwd;         // [15]
    Sim3T* inv;         // [15]
    static constexpr int W = LANES >= 32 ? LANES / 32 : 1;
    static constexpr int kWarpDoubles = kSimAcc * kSimRedStride + kSimAcc;
    __host__ __device__ inline void sync() const
    {
#ifdef __CUDA_ARCH__
        if (LANES > 32) __syncthreads(); else __syncwarp();
#endif
    }
    __host__ __device__ inline void sum1(double& v) const
    {
#ifdef __CUDA_ARCH__
        if (LANES >= 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        if (LANES > 32) {
            double* s1 = red + W * kWarpDoubles;
            const int w = (threadIdx.x % LANES) >> 5;
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s1[w] = v;
            __syncthreads();
            double t = s1[0];
#pragma unroll
            for (int k = 1; k < W; ++k) t += s1[k];
            v = t;
        }
#else
        (void)v;
#endif
    }
    __host__ __device__ inline void sum36(double* v) const
    {
#ifdef __CUDA_ARCH__
        if (LANES >= 32) {
            const int lane = threadIdx.x & 31;
            const int w = (threadIdx.x % LANES) >> 5;
            double* part = red + w * kWarpDoubles;
            sync();
#pragma unroll
            for (int k = 0; k < kSimAcc; ++k) part[k * kSimRedStride + lane] = v[k];
            __syncwarp();
            for (int k = lane; k < kSimAcc; k += 32) {
                const double* row = part + k * kSimRedStride;
                double t0 = row[0], t1 = row[1];
#pragma unroll
                for (int l = 2; l < 32; l += 2) { t0 += row[l]; t1 += row[l + 1]; }
                part[kSimAcc * kSimRedStride + k] = t0 + t1;
            }
            sync();
#pragma unroll
            for (int k = 0; k < kSimAcc; ++k) {
                double t = red[kSimAcc * kSimRedStride + k];
#pragma unroll
                for (int x = 1; x < W; ++x) t += red[x * kWarpDoubles + kSimAcc * kSimRedStride + k];
                v[k] = t;
            }
        }
#else
        (void)v;
#endif
    }
};

__host__ __device__ inline double huber_rho(double c, double delta, double dsqr)
{
    return (c <= dsqr) ? c : 2.0 * sqrt(c) * delta - dsqr;
}

// computeActiveErrors + activeRobustChi2 at S (edges of matches with flag == 0)
template <int LANES>
__host__ __device__ inline double active_chi2(const PairView& f, const uint8_t* flag, const Sim3T& S, int lane, const SimShared<LANES>& sh)
{
    Sim3T Si;
    sim3_inverse(S, Si);
    double sum = 0.0;
    for (int i = lane; i < f.n; i += LANES) {
        if (flag[i]) continue;
        double e0, e1;
        edge_err(S, (double)f.x2[3 * i], (double)f.x2[3 * i + 1], (double)f.x2[3 * i + 2], (double)f.o1[2 * i], (double)f.o1[2 * i + 1], f.K1, e0, e1);
        const double s1 = (double)f.is1[i];
        sum += huber_rho(e0 * (s1 * e0) + e1 * (s1 * e1), f.delta, f.dsqr);
        edge_err(Si, (double)f.x1[3 * i], (double)f.x1[3 * i + 1], (double)f.x1[3 * i + 2], (double)f.o2[2 * i], (double)f.o2[2 * i + 1], f.K2, e0, e1);
        const double s2 = (double)f.is2[i];
        sum += huber_rho(e0 * (s2 * e0) + e1 * (s2 * e1), f.delta, f.dsqr);
    }
    sh.sum1(sum);
    return sum;
}

// errors at S, numeric Jacobians (central differences, delta = 1e-9) and the quadratic form of every active edge
template <int LANES>
__host__ __device__ inline double build_system(const PairView& f, const uint8_t* flag, int fix_scale, const Sim3T& S, int lane,
                                               const SimShared<LANES>& sh, double* H /*49*/, double* b /*7*/)
{
    // the 15 transforms: [0] = S, [1 + 2d] = exp(+delta e_d) S, [2 + 2d] = exp(-delta e_d) S, and their inverses
    sh.sync();
    for (int p = lane; p < kSimPoses; p += LANES) {
        Sim3T T = S;
        if (p > 0) {
            double x[7] = {0, 0, 0, 0, 0, 0, 0};
            const int d = (p - 1) >> 1;
            x[d] = ((p - 1) & 1) ? -1e-9 : 1e-9;
            if (fix_scale) x[6] = 0.0;                 // VertexSim3Expmap::oplusImpl
            Sim3T E;
            sim3_exp(x, E);
            sim3_mul(E, S, T);
        }
        sh.fwd[p] = T;
        sim3_inverse(T, sh.inv[p]);
    }
    sh.sync();
    double all[kSimAcc];
#pragma unroll
    for (int k = 0; k < kSimAcc; ++k) all[k] = 0.0;
    const double scalar = 1.0 / (2.0 * 1e-9);
    for (int i = lane; i < f.n; i += LANES) {
        if (flag[i]) continue;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const float* xp = side ? f.x1 : f.x2;
            const float* op = side ? f.o2 : f.o1;
            const double* K = side ? f.K2 : f.K1;
            const Sim3T* P = side ? sh.inv : sh.fwd;
            const double X = (double)xp[3 * i], Y = (double)xp[3 * i + 1], Z = (double)xp[3 * i + 2];
            const double u = (double)op[2 * i], v = (double)op[2 * i + 1];
            const double s = (double)(side ? f.is2[i] : f.is1[i]);
            double e0, e1;
            edge_err(P[0], X, Y, Z, u, v, K, e0, e1);
            double J0[7], J1[7];
#pragma unroll
            for (int d = 0; d < 7; ++d) {
                double a0, a1, b0, b1;
                edge_err(P[1 + 2 * d], X, Y, Z, u, v, K, a0, a1);
                edge_err(P[2 + 2 * d], X, Y, Z, u, v, K, b0, b1);
                J0[d] = scalar * (a0 - b0);
                J1[d] = scalar * (a1 - b1);
            }
            const double c = e0 * (s * e0) + e1 * (s * e1);
            double rho1 = 1.0;
            if (c <= f.dsqr) all[35] += c;
            else { const double sq = sqrt(c); all[35] += 2.0 * sq * f.delta - f.dsqr; rho1 = f.delta / sq; }
            const double wo = rho1 * s;
            const double we0 = s * e0, we1 = s * e1;
            int k = 0;
#pragma unroll
            for (int a = 0; a < 7; ++a) {
                const double be = po_fma(J1[a], we1, J0[a] * we0);
                all[28 + a] -= rho1 * be;
#pragma unroll
                for (int c2 = a; c2 < 7; ++c2) {
                    double h = all[k];
                    h = po_fma(J0[a], wo * J0[c2], h);
                    h = po_fma(J1[a], wo * J1[c2], h);
                    all[k++] = h;
                }
            }
        }
    }
    sh.sum36(all);
    int k = 0;
    for (int a = 0; a < 7; ++a) {
        b[a] = all[28 + a];
        for (int c2 = a; c2 < 7; ++c2) {
            H[7 * a + c2] = all[k];
            H[7 * c2 + a] = all[k];
            ++k;
        }
    }
    return all[35];
}

// LDL^T without pivoting of (H + lambda I), 7x7, one reciprocal per pivot
__host__ __device__ inline bool ldlt7(const double* Hin, double lambda, const double* b, double* x)
{
    double L[49], d[7], r[7], y[7];
    for (int i = 0; i < 49; ++i) L[i] = Hin[i];
    for (int j = 0; j < 7; ++j) L[8 * j] += lambda;
    for (int j = 0; j < 7; ++j) {
        double dj = L[8 * j];
        for (int k = 0; k < j; ++k) dj -= L[7 * j + k] * L[7 * j + k] * d[k];
        if (!(dj > 0.0)) return false;
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
    return true;
}

template <int LANES>
__host__ __device__ inline void optimize(const PairView& f, const uint8_t* flag, int fix_scale, Sim3T& S, Sim3T& Serr, int iterations,
                                         int lane, const SimShared<LANES>& sh, po::Stats& st)
{
    double lambda = 0.0, ni = 2.0;
    int nBad = 0;
    double x[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iterations; ++it) {
        double H[49], b[7];
        double currentChi = build_system<LANES>(f, flag, fix_scale, S, lane, sh, H, b);
        Serr = S;
        const double iniChi = currentChi;
        ++st.iterations;
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
            const Sim3T backup = S;
            const bool ok2 = ldlt7(H, lambda, b, x);
            ++st.trials;
            if (fix_scale) x[6] = 0.0;                  // oplusImpl writes through the solver's x (types_seven_dof_expmap.h:62-65)
            Sim3T E, Sn;
            sim3_exp(x, E);
            sim3_mul(E, S, Sn);
            S = Sn;
            double tempChi = active_chi2<LANES>(f, flag, S, lane, sh);
            Serr = S;
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
                S = backup;
            }
            ++qmax;
        } while (rho < 0.0 && qmax < 10);
        if (qmax == 10 || rho == 0.0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) break;
    }
}

// match i is bad when either edge's stored chi2 exceeds th2 (Optimizer.cpp:1181,1216)
template <int LANES>
__host__ __device__ inline int classify(const PairView& f, uint8_t* flag, const Sim3T& Serr, int lane, const SimShared<LANES>& sh, int& kept)
{
    Sim3T Si;
    sim3_inverse(Serr, Si);
    int bad = 0, good = 0;
    for (int i = lane; i < f.n; i += LANES) {
        if (flag[i]) continue;
        double e0, e1;
        edge_err(Serr, (double)f.x2[3 * i], (double)f.x2[3 * i + 1], (double)f.x2[3 * i + 2], (double)f.o1[2 * i], (double)f.o1[2 * i + 1], f.K1, e0, e1);
        const double s1 = (double)f.is1[i];
        const double c12 = e0 * (s1 * e0) + e1 * (s1 * e1);
        edge_err(Si, (double)f.x1[3 * i], (double)f.x1[3 * i + 1], (double)f.x1[3 * i + 2], (double)f.o2[2 * i], (double)f.o2[2 * i + 1], f.K2, e0, e1);
        const double s2 = (double)f.is2[i];
        const double c21 = e0 * (s2 * e0) + e1 * (s2 * e1);
        if (c12 > f.th2 || c21 > f.th2) { flag[i] = 1; ++bad; }
        else ++good;
    }
    double bd = (double)bad, gd = (double)good;
    sh.sum1(bd);
    sh.sum1(gd);
    sh.sync();
    kept = (int)gd;
    return (int)bd;
}

template <int LANES>
__host__ __device__ inline void optimize_sim3(const Sim3OptMeta& m, const float* x1, const float* x2, const float* o1, const float* o2,
                                              const float* is1, const float* is2, uint8_t* removed, int lane, double* smem,
                                              int problem, rsac_sim3opt_result* out)
{
    PairView f;
    f.x1 = x1 + 3 * m.off; f.x2 = x2 + 3 * m.off; f.o1 = o1 + 2 * m.off; f.o2 = o2 + 2 * m.off; f.is1 = is1 + m.off; f.is2 = is2 + m.off;
    for (int k = 0; k < 4; ++k) { f.K1[k] = (double)m.K1[k]; f.K2[k] = (double)m.K2[k]; }
    f.delta = (double)sqrtf(m.th2);           // const float deltaHuber = sqrt(th2)
    f.dsqr = f.delta * f.delta;
    f.th2 = (double)m.th2;
    f.n = m.n;
    SimShared<LANES> sh;
    sh.red = smem;
    sh.fwd = reinterpret_cast<Sim3T*>(smem + SimShared<LANES>::W * SimShared<LANES>::kWarpDoubles + 8);
    sh.inv = sh.fwd + kSimPoses;
    uint8_t* flag = removed + m.off;
    for (int i = lane; i < f.n; i += LANES) flag[i] = 0;
    Sim3T S;
    {
        double R[9];
        for (int i = 0; i < 9; ++i) R[i] = (double)m.R12[i];
        quat_from_rot(R, S.q);                 // g2o::Sim3(R, t, s): Quaterniond(R)
        for (int i = 0; i < 3; ++i) S.t[i] = (double)m.t12[i];
        S.s = (double)m.s12;
    }
    const Sim3T S0 = S;
    po::Stats st = {0, 0};
    int nBad = 0, nIn = 0;
    bool second = false;
    if (f.n > 0) {
        Sim3T Serr = S;
        optimize<LANES>(f, flag, m.fix_scale, S, Serr, 5, lane, sh, st);
        int kept = 0;
        nBad = classify<LANES>(f, flag, Serr, lane, sh, kept);
        if (f.n - nBad >= 10) {
            second = true;
            const int more = nBad > 0 ? 10 : 5;
            Serr = S;
            optimize<LANES>(f, flag, m.fix_scale, S, Serr, more, lane, sh, st);
            classify<LANES>(f, flag, Serr, lane, sh, kept);
            nIn = kept;
        }
    }
    if (lane == 0) {
        const Sim3T& F = second ? S : S0;      // `return 0` leaves g2oS12 untouched (Optimizer.cpp:1203-1204)
        out->n_inliers = second ? nIn : 0;
        out->n_bad = nBad;
        out->optimized = second ? 1 : 0;
        out->iterations = st.iterations;
        out->trials = st.trials;
        out->problem = problem;
        po::quat_to_rot_d(F.q, out->R);
        for (int i = 0; i < 3; ++i) out->t[i] = F.t[i];
        out->s = F.s;
        for (int i = 0; i < 4; ++i) out->q[i] = F.q[i];
    }
}

}  // namespace so

constexpr int kSim3OptWarps = 4;      // warps per CTA: four pairs (LANES = 32) or one pair (LANES = 128)

template <int LANES>
__global__ void __launch_bounds__(kSim3OptWarps * 32) sim3opt_kernel(const Sim3OptMeta* __restrict__ metas, int C,
                                                                     const float* __restrict__ x1, const float* __restrict__ x2,
                                                                     const float* __restrict__ o1, const float* __restrict__ o2,
                                                                     const float* __restrict__ is1, const float* __restrict__ is2,
                                                                     uint8_t* __restrict__ removed, rsac_sim3opt_result* __restrict__ results, int problem_base)
{
    extern __shared__ double sim3opt_smem[];
    const int w = threadIdx.x / LANES;
    const int c = blockIdx.x * (blockDim.x / LANES) + w;
    if (c >= C) return;
    const Sim3OptMeta m = metas[c];
    so::optimize_sim3<LANES>(m, x1, x2, o1, o2, is1, is2, removed, threadIdx.x % LANES, sim3opt_smem + w * so::sim_smem_doubles(LANES),
                             problem_base + c, results + c);
}

}  // namespace rsac
