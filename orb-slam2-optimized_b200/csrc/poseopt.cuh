// poseopt.cuh -- batched Optimizer::PoseOptimization: motion-only bundle adjustment of one
// frame pose against its matched map points, Huber kernel, 4 rounds x <= 10 Levenberg-Marquardt iterations with
// outlier re-classification between rounds.
//
// Reference: src/Optimizer.cpp:205-424 on top of the vendored g2o
//   (core/optimization_algorithm_levenberg.cpp:59-179, core/sparse_optimizer.cpp optimize(),
//    core/base_unary_edge.hpp constructQuadraticForm, core/robust_kernel_impl.cpp:78-91,
//    types/types_six_dof_expmap.cpp:266-364, types/se3quat.h:60-66,104-110,217-260, solvers/linear_solver_dense.h).
//
// Mapping: ONE WARP PER FRAME.  The graph has a single 6-dof vertex, so an LM iteration is two streaming passes
// over the frame's edges (errors + Jacobians -> 6x6 normal equations; errors -> robust chi2 of the trial pose)
// and a 6x6 solve.  Lanes stride over the edges and keep partial sums of the 21 + 6 + 1 accumulators in
// registers; a butterfly of warp shuffles leaves the identical total in every lane, so the 6x6 LDL^T, the SE3
// exponential and the step control run redundantly in all lanes without any broadcast or barrier.
// The edges' stored errors (g2o keeps `_error` per edge) are never materialised: an edge that is active in a round
// is classified with the pose of the round's last error pass (`Terr`, which is the rejected trial pose when the last
// LM trial failed -- the reference classifies with those stale errors, Optimizer.cpp:332-339), a flagged edge with
// the round's final estimate.
//
// The numerical core is __host__ __device__ and templated on the lane count: LANES = 1 compiled for the host
// (rsac_debug_host_poseopt) sums the edges in the reference's order and is compared bit-for-bit with the oracle
// in the CPU suite; LANES = 32 is the kernel.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include "common.cuh"

namespace rsac {

struct PoseOptMeta {
    int64_t off;       // first edge of the frame in the flat arrays
    int32_t n;         // edges (keypoints with a MapPoint)
    float K[5];        // fx, fy, cx, cy, bf
    float Rcw[9];      // Frame::mTcw
    float tcw[3];
};

struct Se3 { double q[4]; double t[3]; };   // quaternion (w, x, y, z) + translation, as g2o::SE3Quat

namespace po {

#ifndef RSAC_PO_UNROLL
#define RSAC_PO_UNROLL 1
#endif
constexpr int kEdgeUnroll = RSAC_PO_UNROLL;     // edges of a lane unrolled together; measured on B200 (1024 x 250): 1: 0.53 ms, 2: 0.70, 3: 0.81
                                                // -- the code outgrows the instruction cache faster than the ILP pays

__host__ __device__ inline double po_fma(double a, double b, double c) { return fma(a, b, c); }   // twin: fma() in oracle/orc_poseopt.c

__host__ __device__ inline void quat_from_rot(const double* m, double* q)
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
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
        v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
        q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
    }
}

__host__ __device__ inline void quat_normalize(double* q)
{
    if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// Quaternion * Vector3 as Eigen evaluates it: v + w uv + qv x uv, uv = 2 (qv x v)
__host__ __device__ inline void quat_rotate(const double* q, double v0, double v1, double v2, double& o0, double& o1, double& o2)
{
    double u0 = q[2] * v2 - q[3] * v1;
    double u1 = q[3] * v0 - q[1] * v2;
    double u2 = q[1] * v1 - q[2] * v0;
    u0 += u0; u1 += u1; u2 += u2;
    o0 = v0 + q[0] * u0 + (q[2] * u2 - q[3] * u1);
    o1 = v1 + q[0] * u1 + (q[3] * u0 - q[1] * u2);
    o2 = v2 + q[0] * u2 + (q[1] * u1 - q[2] * u0);
}

__host__ __device__ inline void se3_from_float(const float* R, const float* t, Se3& T)
{
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = (double)R[i];
    quat_from_rot(m, T.q);
    quat_normalize(T.q);
    for (int i = 0; i < 3; ++i) T.t[i] = (double)t[i];
}

// T <- exp(x) * T  (VertexSE3Expmap::oplusImpl, types_six_dof_expmap.h:73-76)
__host__ __device__ inline void se3_oplus(const double* x, Se3& T)
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
    double eq[4], et[3];
    quat_from_rot(R, eq);
    quat_normalize(eq);
    for (int i = 0; i < 3; ++i) et[i] = V[3 * i] * x[3] + V[3 * i + 1] * x[4] + V[3 * i + 2] * x[5];
    // SE3Quat::operator*: t = et + eq * T.t ; q = eq * T.q ; normalise
    double r0, r1, r2;
    quat_rotate(eq, T.t[0], T.t[1], T.t[2], r0, r1, r2);
    T.t[0] = et[0] + r0; T.t[1] = et[1] + r1; T.t[2] = et[2] + r2;
    const double b0 = T.q[0], b1 = T.q[1], b2 = T.q[2], b3 = T.q[3];
    T.q[0] = eq[0] * b0 - eq[1] * b1 - eq[2] * b2 - eq[3] * b3;
    T.q[1] = eq[0] * b1 + eq[1] * b0 + eq[2] * b3 - eq[3] * b2;
    T.q[2] = eq[0] * b2 + eq[2] * b0 + eq[3] * b1 - eq[1] * b3;
    T.q[3] = eq[0] * b3 + eq[3] * b0 + eq[1] * b2 - eq[2] * b1;
    quat_normalize(T.q);
}

__host__ __device__ inline void quat_to_rot_d(const double* q, double* R)
{
    const double tx = 2.0 * q[1], ty = 2.0 * q[2], tz = 2.0 * q[3];
    const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
    const double txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
    const double tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

struct Edge {
    double X, Y, Z, u, v, s;
    float ur;
};

struct FrameView {
    const float* p3d;      // [n][3]
    const float* obs;      // [n][3]
    const float* isig;     // [n]
    double fx, fy, cx, cy, bf;
    int n;
};

__host__ __device__ inline Edge load_edge(const FrameView& f, int i)
{
    Edge e;
    e.X = (double)f.p3d[3 * i]; e.Y = (double)f.p3d[3 * i + 1]; e.Z = (double)f.p3d[3 * i + 2];
    e.u = (double)f.obs[3 * i]; e.v = (double)f.obs[3 * i + 1]; e.ur = f.obs[3 * i + 2];
    e.s = (double)f.isig[i];
    return e;
}

// computeError (mono: project2d then *f + c; stereo: float 1/z, types_six_dof_expmap.cpp:290-306); returns chi2
__host__ __device__ inline double edge_error(const FrameView& f, const Edge& e, const Se3& T, double& x, double& y, double& z,
                                             double& e0, double& e1, double& e2)
{
    double r0, r1, r2;
    quat_rotate(T.q, e.X, e.Y, e.Z, r0, r1, r2);
    x = r0 + T.t[0]; y = r1 + T.t[1]; z = r2 + T.t[2];
    if (e.ur < 0.0f) {
        e0 = e.u - ((x / z) * f.fx + f.cx);
        e1 = e.v - ((y / z) * f.fy + f.cy);
        e2 = 0.0;
        return e0 * (e.s * e0) + e1 * (e.s * e1);
    }
    const float invz = (float)(1.0 / z);
    const double p0 = x * (double)invz * f.fx + f.cx;
    const double p1 = y * (double)invz * f.fy + f.cy;
    const double p2 = p0 - f.bf * (double)invz;
    e0 = e.u - p0; e1 = e.v - p1; e2 = (double)e.ur - p2;
    double c = e0 * (e.s * e0) + e1 * (e.s * e1);
    c += e2 * (e.s * e2);
    return c;
}

__host__ __device__ inline double huber_delta(bool stereo)
{
    // const float deltaMono = sqrt(5.991), deltaStereo = sqrt(7.815) (Optimizer.cpp:241-242), widened by setDelta(double)
    return stereo ? 0x1.65d4p+1 : 0x1.394ca8p+1;   // (double)(float)sqrt(7.815) = 2.7955322265625, (double)(float)sqrt(5.991) = 2.4476518630981445
}

// Sum over the 32 lanes that share a frame.  Host (LANES = 1): identity.
// One value: shuffle butterfly (every lane ends with the identical total).  The 28 accumulators of the normal
// equations: a butterfly would cost 280 SHFL (each with its WARPSYNC / ENDCOLLECTIVE pair, the loops around it
// have data-dependent exits) -- instead every lane parks its partials in shared memory ([28][33] doubles, padded
// against bank conflicts), lane k adds the 32 partials of accumulator k in lane order (two interleaved chains), and
// all lanes read the 28 totals back: ~130 instructions, and every lane holds bit-identical totals, which the
// redundant 6x6 solve and step control need to take the same branches.
constexpr int kRedStride = 33;
// per problem: W warps x (partials [28][33] + warp totals [28]) + W single-value slots
__host__ __device__ constexpr int red_doubles(int lanes) { return (lanes / 32) * (28 * kRedStride + 28) + (lanes / 32); }
constexpr int kRedDoubles = red_doubles(32);

// LANES = 32: one warp per problem (throughput shape, __syncwarp only).  LANES = 128: four warps per problem, one
// problem per CTA (latency shape for small batches): every warp reduces its own rows, the warp totals meet in shared
// memory across __syncthreads() and every thread adds them in warp order -- all 128 threads hold bit-identical totals
// and follow identical control flow, which is what makes the block-wide barriers legal.
template <int LANES>
struct Reducer {
    double* sm;     // this problem's red_doubles(LANES)
    static constexpr int W = LANES >= 32 ? LANES / 32 : 1;
    __host__ __device__ inline void barrier() const
    {
#ifdef __CUDA_ARCH__
        if (LANES > 32) __syncthreads(); else __syncwarp();
#endif
    }
    __host__ __device__ inline void sum1(double& v)
    {
#ifdef __CUDA_ARCH__
        if (LANES >= 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        if (LANES > 32) {
            double* s1 = sm + W * (28 * kRedStride + 28);
            const int w = (threadIdx.x % LANES) >> 5;
            __syncthreads();                                // the previous totals have been read
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
    __host__ __device__ inline void sum28(double* v)
    {
#ifdef __CUDA_ARCH__
        if (LANES >= 32) {
            const int lane = threadIdx.x & 31;
            const int w = (threadIdx.x % LANES) >> 5;
            double* part = sm + w * (28 * kRedStride + 28);
            barrier();                                      // the previous totals have been read by every lane
#pragma unroll
            for (int k = 0; k < 28; ++k) part[k * kRedStride + lane] = v[k];
            __syncwarp();
            if (lane < 28) {
                const double* row = part + lane * kRedStride;
                double t0 = row[0], t1 = row[1];
#pragma unroll
                for (int l = 2; l < 32; l += 2) { t0 += row[l]; t1 += row[l + 1]; }
                part[28 * kRedStride + lane] = t0 + t1;
            }
            barrier();
#pragma unroll
            for (int k = 0; k < 28; ++k) {
                double t = sm[28 * kRedStride + k];
#pragma unroll
                for (int x = 1; x < W; ++x) t += sm[x * (28 * kRedStride + 28) + 28 * kRedStride + k];
                v[k] = t;
            }
        }
#else
        (void)v;
#endif
    }
};

// computeActiveErrors + activeRobustChi2 at pose T
template <int LANES>
__host__ __device__ inline double active_chi2(const FrameView& f, const uint8_t* level, bool robust, const Se3& T, int lane,
                                              Reducer<LANES>& red)
{
    double sum = 0.0;
#pragma unroll(kEdgeUnroll)
    for (int i = lane; i < f.n; i += LANES) {
        if (level[i]) continue;
        const Edge e = load_edge(f, i);
        double x, y, z, e0, e1, e2;
        const double c = edge_error(f, e, T, x, y, z, e0, e1, e2);
        if (robust) {
            const double delta = huber_delta(!(e.ur < 0.0f));
            const double dsqr = delta * delta;
            sum += (c <= dsqr) ? c : 2.0 * sqrt(c) * delta - dsqr;
        } else {
            sum += c;
        }
    }
    red.sum1(sum);
    return sum;
}

// errors at T + linearizeOplus + constructQuadraticForm: upper triangle of H (21), b (6), robust chi2
template <int LANES>
__host__ __device__ inline double build_system(const FrameView& f, const uint8_t* level, bool robust, const Se3& T, int lane,
                                               Reducer<LANES>& red, double* H /*36, symmetric on return*/, double* b)
{
    double all[28];                       // acc[21] | bb[6] | chi : reduced together
    double* acc = all;
    double* bb = all + 21;
    double& chi = all[27];
#pragma unroll
    for (int k = 0; k < 28; ++k) all[k] = 0.0;
#pragma unroll(kEdgeUnroll)
    for (int i = lane; i < f.n; i += LANES) {
        if (level[i]) continue;
        const Edge e = load_edge(f, i);
        double x, y, z, er[3];
        const double c = edge_error(f, e, T, x, y, z, er[0], er[1], er[2]);
        const bool stereo = !(e.ur < 0.0f);
        double rho1 = 1.0;
        if (robust) {
            const double delta = huber_delta(stereo);
            const double dsqr = delta * delta;
            if (c <= dsqr) chi += c;
            else { const double sq = sqrt(c); chi += 2.0 * sq * delta - dsqr; rho1 = delta / sq; }
        } else {
            chi += c;
        }
        const double invz = 1.0 / z, invz_2 = invz * invz;
        double J[18];
        J[0] = x * y * invz_2 * f.fx;
        J[1] = -(1.0 + (x * x * invz_2)) * f.fx;
        J[2] = y * invz * f.fx;
        J[3] = -invz * f.fx;
        J[4] = 0.0;
        J[5] = x * invz_2 * f.fx;
        J[6] = (1.0 + y * y * invz_2) * f.fy;
        J[7] = -x * y * invz_2 * f.fy;
        J[8] = -x * invz * f.fy;
        J[9] = 0.0;
        J[10] = -invz * f.fy;
        J[11] = y * invz_2 * f.fy;
        if (stereo) {
            J[12] = J[0] - f.bf * y * invz_2;
            J[13] = J[1] + f.bf * x * invz_2;
            J[14] = J[2];
            J[15] = J[3];
            J[16] = 0.0;
            J[17] = J[5] - f.bf * invz_2;
        } else {
#pragma unroll
            for (int k = 12; k < 18; ++k) J[k] = 0.0;
        }
        // constructQuadraticForm: H += J^T (rho' Omega) J, b -= rho' J^T Omega e, accumulated with explicit fused
        // multiply-adds (arithmetic contract).  Column 4 of rows 0 and 2 and column 3 of row 1 are exact
        // zeros (J[4] = J[9] = J[16] = 0): their products are skipped -- 30 of 42 (45 of 63 with the stereo row) remain.
        const double wo = rho1 * e.s;
        double w0[6], w1[6], w2[6];
#pragma unroll
        for (int c2 = 0; c2 < 6; ++c2) { w0[c2] = wo * J[c2]; w1[c2] = wo * J[6 + c2]; w2[c2] = wo * J[12 + c2]; }
        const double we0 = e.s * er[0], we1 = e.s * er[1], we2 = e.s * er[2];
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double be = 0.0;
            if (a != 4) be = J[a] * we0;
            if (a != 3) be = po_fma(J[6 + a], we1, be);
            if (stereo && a != 4) be = po_fma(J[12 + a], we2, be);
            bb[a] -= rho1 * be;
#pragma unroll
            for (int c2 = a; c2 < 6; ++c2) {
                double h = acc[k];
                if (a != 4 && c2 != 4) h = po_fma(J[a], w0[c2], h);
                if (a != 3 && c2 != 3) h = po_fma(J[6 + a], w1[c2], h);
                if (stereo && a != 4 && c2 != 4) h = po_fma(J[12 + a], w2[c2], h);
                acc[k++] = h;
            }
        }
    }
    red.sum28(all);
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        b[a] = bb[a];
#pragma unroll
        for (int c2 = a; c2 < 6; ++c2) {
            const double h = acc[k++];
            H[6 * a + c2] = h;
            H[6 * c2 + a] = h;
        }
    }
    return chi;
}

// LDL^T without pivoting of (H + lambda I); false when a pivot is not positive (Eigen LDLT::isPositive())
__host__ __device__ inline bool ldlt6(const double* Hin, double lambda, const double* b, double* x)
{
    double L[36], d[6], r[6], y[6];
#pragma unroll
    for (int i = 0; i < 36; ++i) L[i] = Hin[i];
#pragma unroll
    for (int j = 0; j < 6; ++j) L[7 * j] += lambda;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double dj = L[7 * j];
#pragma unroll
        for (int k = 0; k < j; ++k) dj -= L[6 * j + k] * L[6 * j + k] * d[k];
        if (!(dj > 0.0)) return false;
        d[j] = dj;
        const double rj = 1.0 / dj;          // one reciprocal per pivot; the column and the diagonal solve multiply by it
        r[j] = rj;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = L[6 * i + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[6 * i + k] * L[6 * j + k] * d[k];
            L[6 * i + j] = v * rj;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= L[6 * i + k] * y[k];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] *= r[i];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) v -= L[6 * k + i] * x[k];
        x[i] = v;
    }
    return true;
}

struct Stats { int iterations, trials; };

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg.  T: estimate; Terr: pose of the last error pass.
template <int LANES>
__host__ __device__ inline void optimize(const FrameView& f, const uint8_t* level, bool robust, Se3& T, Se3& Terr, int iterations,
                                         int lane, Reducer<LANES>& red, Stats& st)
{
    double lambda = 0.0, ni = 2.0;
    int nBad = 0;
    double x[6] = {0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iterations; ++it) {
        double H[36], b[6];
        double currentChi = build_system<LANES>(f, level, robust, T, lane, red, H, b);
        Terr = T;
        const double iniChi = currentChi;
        ++st.iterations;
        if (it == 0) {
            double maxDiag = 0.0;
            for (int j = 0; j < 6; ++j) maxDiag = fmax(fabs(H[7 * j]), maxDiag);
            lambda = 1e-5 * maxDiag;     // _tau * maxDiagonal (computeLambdaInit)
            ni = 2.0;
            nBad = 0;
        }
        double rho = 0.0;
        int qmax = 0;
        do {
            const Se3 backup = T;
            const bool ok2 = ldlt6(H, lambda, b, x);
            ++st.trials;
            se3_oplus(x, T);
            double tempChi = active_chi2<LANES>(f, level, robust, T, lane, red);
            Terr = T;
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
                T = backup;
            }
            ++qmax;
        } while (rho < 0.0 && qmax < 10);
        if (qmax == 10 || rho == 0.0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) break;
    }
}

// the whole of Optimizer::PoseOptimization for one frame.  outlier[n] doubles as the edges' level (level == outlier
// after every classification, Optimizer.cpp:341-352).  Returns nInitialCorrespondences - nBad.
template <int LANES>
__host__ __device__ inline void pose_optimization(const PoseOptMeta& m, const float* p3d, const float* obs, const float* isig,
                                                  uint8_t* outlier, int lane, double* red_smem, int problem, rsac_poseopt_result* out)
{
    Reducer<LANES> red;
    red.sm = red_smem;
    FrameView f;
    f.p3d = p3d + 3 * m.off; f.obs = obs + 3 * m.off; f.isig = isig + m.off;
    f.fx = (double)m.K[0]; f.fy = (double)m.K[1]; f.cx = (double)m.K[2]; f.cy = (double)m.K[3]; f.bf = (double)m.K[4];
    f.n = m.n;
    uint8_t* flag = outlier + m.off;
    Se3 T;
    se3_from_float(m.Rcw, m.tcw, T);
    const int nEdges = f.n;
    for (int i = lane; i < f.n; i += LANES) flag[i] = 0;
    Stats st = {0, 0};
    int nBad = 0, rounds = 0;
    bool robust = true;
    if (nEdges >= 3) {
        int active = nEdges;
        for (int it = 0; it < 4; ++it) {
            se3_from_float(m.Rcw, m.tcw, T);
            Se3 Terr = T;
            if (active > 0) optimize<LANES>(f, flag, robust, T, Terr, 10, lane, red, st);
            int bad = 0;
            for (int i = lane; i < f.n; i += LANES) {
                const Edge e = load_edge(f, i);
                double x, y, z, e0, e1, e2;
                const float chi2 = (float)edge_error(f, e, flag[i] ? T : Terr, x, y, z, e0, e1, e2);
                const bool is_out = chi2 > (e.ur < 0.0f ? 5.991f : 7.815f);
                flag[i] = is_out ? 1 : 0;
                bad += is_out ? 1 : 0;
            }
            double badd = (double)bad;     // exact: counts are far below 2^53
            red.sum1(badd);
            bad = (int)badd;
            nBad = bad;
            active = nEdges - bad;
            if (it == 2) robust = false;
            ++rounds;
            if (nEdges < 10) break;
        }
    }
    if (lane == 0) {
        out->n_inliers = nEdges >= 3 ? nEdges - nBad : 0;
        out->n_bad = nBad;
        out->rounds = rounds;
        out->iterations = st.iterations;
        out->trials = st.trials;
        out->problem = problem;
        quat_to_rot_d(T.q, out->R);
        for (int i = 0; i < 3; ++i) out->t[i] = T.t[i];
        for (int i = 0; i < 9; ++i) out->Rf[i] = (float)out->R[i];
        for (int i = 0; i < 3; ++i) out->tf[i] = (float)out->t[i];
    }
}

}  // namespace po

// Chains PoseOptimization behind a PnP sweep without leaving the device (Tracking.cpp:1258-1284: the inliers of the
// accepted RANSAC pose become the frame's map points, the pose becomes mTcw, then Optimizer::PoseOptimization): one CTA
// per candidate compacts the inliers of the final mask, in correspondence order, into the optimiser's own arrays --
// point, observation (u, v, -1: monocular), 1/sigma^2 (float division, as ORBextractor fills mvInvLevelSigma2) -- at the
// problem's correspondence offset, remembers where each came from (`src`) and writes the frame record.  Candidates
// without a pose get an empty frame.
static __global__ void __launch_bounds__(128) poseopt_from_pnp_kernel(const ProblemMeta* __restrict__ metas, int C,
                                                              const rsac_result* __restrict__ results, const uint32_t* __restrict__ masks,
                                                              const float* __restrict__ p3d_in, const float* __restrict__ p2d,
                                                              const float* __restrict__ sigma2, float bf, PoseOptMeta* __restrict__ out_metas,
                                                              float* __restrict__ p3d, float* __restrict__ obs, float* __restrict__ isig,
                                                              int32_t* __restrict__ src)
{
    __shared__ int s_pre[2049];                 // exclusive popcount prefix over the mask words (n <= 65536)
    const int c = blockIdx.x;
    if (c >= C) return;
    const ProblemMeta m = metas[c];
    const rsac_result r = results[c];
    const int words = r.ok ? min(m.words, 2048) : 0;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < words; ++w) { s_pre[w] = acc; acc += __popc(masks[m.word_off + w]); }
        s_pre[words] = acc;
        PoseOptMeta o;
        o.off = m.corr_off;
        o.n = acc;
        o.K[0] = (float)m.fx; o.K[1] = (float)m.fy; o.K[2] = (float)m.cx; o.K[3] = (float)m.cy; o.K[4] = bf;
        for (int k = 0; k < 9; ++k) o.Rcw[k] = r.R[k];
        for (int k = 0; k < 3; ++k) o.tcw[k] = r.t[k];
        out_metas[c] = o;
    }
    __syncthreads();
    const int n = min(m.n, words * 32);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t wv = masks[m.word_off + (i >> 5)];
        if (!((wv >> (i & 31)) & 1u)) continue;
        const int k = s_pre[i >> 5] + __popc(wv & ((1u << (i & 31)) - 1u));
        const size_t g = (size_t)m.corr_off + i, d = (size_t)m.corr_off + k;
        p3d[3 * d] = p3d_in[3 * g]; p3d[3 * d + 1] = p3d_in[3 * g + 1]; p3d[3 * d + 2] = p3d_in[3 * g + 2];
        obs[3 * d] = p2d[2 * g]; obs[3 * d + 1] = p2d[2 * g + 1]; obs[3 * d + 2] = -1.0f;
        isig[d] = 1.0f / sigma2[g];
        src[d] = i;
    }
}

// flags of a chained run back in the PnP correspondence index space: 2 = not an edge, else the optimiser's flag
static __global__ void __launch_bounds__(128) poseopt_scatter_flags_kernel(const ProblemMeta* __restrict__ metas, int C,
                                                                   const PoseOptMeta* __restrict__ frames, const uint8_t* __restrict__ flag,
                                                                   const int32_t* __restrict__ src, uint8_t* __restrict__ full)
{
    const int c = blockIdx.x;
    if (c >= C) return;
    const ProblemMeta m = metas[c];
    const int n_edges = frames[c].n;
    for (int i = threadIdx.x; i < m.n; i += blockDim.x) full[(size_t)m.corr_off + i] = 2;
    __syncthreads();
    for (int k = threadIdx.x; k < n_edges; k += blockDim.x) full[(size_t)m.corr_off + src[(size_t)m.corr_off + k]] = flag[(size_t)m.corr_off + k];
}

constexpr int kPoseOptWarps = 4;      // warps per CTA: four frames (LANES = 32) or one frame (LANES = 128)

// LANES = 32: one warp per frame, blockDim/32 frames per CTA (large batches).  LANES = 128: one frame per CTA, four warps
// on its edges (batches too small to fill the machine: a relocalisation with a handful of accepted candidates).
template <int LANES>
__global__ void __launch_bounds__(kPoseOptWarps * 32) poseopt_kernel(const PoseOptMeta* __restrict__ metas, int C,
                                                                     const float* __restrict__ p3d, const float* __restrict__ obs,
                                                                     const float* __restrict__ isig, uint8_t* __restrict__ outlier,
                                                                     rsac_poseopt_result* __restrict__ results, int problem_base)
{
    extern __shared__ double poseopt_smem[];
    const int w = threadIdx.x / LANES;
    const int c = blockIdx.x * (blockDim.x / LANES) + w;
    if (c >= C) return;                                    // uniform over the LANES threads of a frame (and over the CTA when LANES = 128)
    const PoseOptMeta m = metas[c];
    po::pose_optimization<LANES>(m, p3d, obs, isig, outlier, threadIdx.x % LANES, poseopt_smem + w * po::red_doubles(LANES),
                                 problem_base + c, results + c);
}

}  // namespace rsac
