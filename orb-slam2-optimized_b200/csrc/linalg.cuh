// linalg.cuh -- fixed-size dense solves for the RANSAC minimal solvers (device side).
//
// These stand in for the Eigen calls of the reference (SelfAdjointEigenSolver at
// src/PnPsolver.cpp:311,380,469 / src/Sim3Solver.cpp:238-241 / src/MLPnPsolver.cpp:359,
// bdcSvd().solve at PnPsolver.cpp:531,559,590, Matrix3d::inverse at :331, JacobiSVD /
// FullPivHouseholderQR / LDLT in MLPnPsolver.cpp).  Arithmetic contract:
// only + - * / sqrt, IEEE round-to-nearest, NO FMA contraction (this TU is compiled
// with -fmad=false), fixed operation order -- the CPU checker evaluates the same
// sequences, which is what makes 4-point EPnP hypotheses comparable at all.
#pragma once
#include <cfloat>
#include <cmath>

namespace rsac {

template <typename T> struct JacobiTol;
template <> struct JacobiTol<double> { __host__ __device__ static double scale() { return 0x1p-56; } };
template <> struct JacobiTol<float>  { __host__ __device__ static float scale() { return 0x1p-27f; } };

// IEEE division and square root behind one name each.  RSAC_NOINLINE_DIVSQRT compiles the FP64 ones as out-of-line device
// functions: an inlined double division / square root is ~25 instructions plus a slow-path subroutine, and the solvers
// contain ~200 of them -- out of line the 4-point solver's hot code shrinks by a third (instruction-cache footprint)
#if defined(__CUDA_ARCH__) && defined(RSAC_NOINLINE_DIVSQRT)
__device__ __noinline__ double rsqrt_exact(double x) { return sqrt(x); }
__device__ __noinline__ double rdiv(double a, double b) { return a / b; }
#else
__host__ __device__ inline double rsqrt_exact(double x) { return sqrt(x); }
__host__ __device__ inline double rdiv(double a, double b) { return a / b; }
#endif
__host__ __device__ inline float rsqrt_exact(float x) { return sqrtf(x); }
__host__ __device__ inline float rdiv(float a, float b) { return a / b; }
// explicit fused multiply-add: the only contraction allowed by the arithmetic contract (the build uses
// -fmad=false); every rfma below has a twin in oracle/orc_linalg.c (ORC_FMA)
__host__ __device__ inline double rfma(double a, double b, double c) { return fma(a, b, c); }
__host__ __device__ inline float rfma(float a, float b, float c) { return fmaf(a, b, c); }
__host__ __device__ inline double rabs(double x) { return fabs(x); }
__host__ __device__ inline float rabs(float x) { return fabsf(x); }

constexpr int kMaxSweeps = 30;
constexpr int kMaxSweepsRec = 12;   // recorded-rotation variant: rotations of at most 12 sweeps are kept

// One Jacobi rotation for the symmetric 2x2 block [app apq; apq aqq] (arithmetic contract):
//   h = aqq - app,  r = sqrt(h*h + 4*(apq*apq)),  cos(2t) = |h|/r,
//   c = sqrt(0.5 + 0.5*cos(2t)),  s = apq/(r*c) with the sign of h,
//   new diagonal = (app+aqq)/2 -/+ r/2 (the smaller entry stays the smaller).
// 2 sqrt + 2 div; the angle satisfies |t| <= pi/4.
template <typename T>
__host__ __device__ inline void jacobi_angle(T app, T aqq, T apq, T& c, T& s, T& napp, T& naqq)
{
    const T half = T(0.5);
    const T h = aqq - app;
    const T b2 = apq + apq;
    const T r = rsqrt_exact(rfma(h, h, b2 * b2));
    const T ah = rabs(h);
    const T uu = (r + r) * (r + ah);
    const T ww = rdiv(half + half, rsqrt_exact(uu));
    c = (r + ah) * ww;
    const T s0 = b2 * ww;
    const T m = half * (app + aqq);
    const T hr = half * r;
    if (h < T(0)) { s = -s0; napp = m + hr; naqq = m - hr; }
    else          { s = s0;  napp = m - hr; naqq = m + hr; }
}

// Cyclic Jacobi on the upper triangle of the symmetric N x N matrix a (row-major,
// destroyed).  Eigenvalues ascending in w, eigenvectors in the columns of v (accumulated
// forward).  Rotations with |a_pq| <= ||a||_F * 2^-56 (2^-27 in float) are skipped; the
// solve ends after a sweep without rotations.  Used for the small solves (N = 3, 4).
// STATIC_SORT: the final ordering with compile-time indices only, so that w and v stay in registers (the sub-warp
// solver); the default keeps the run-time-indexed selection sort (w, v in local memory).  Same permutation.
template <typename T, int N, bool STATIC_SORT = false>
__host__ __device__ inline void jacobi_eig(T* a, T* w, T* v)
{
    const T one = T(1), zero = T(0);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) v[i * N + j] = (i == j) ? one : zero;
    T fro2 = zero;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) fro2 = rfma(a[i * N + j], a[i * N + j], fro2);
    const T tol = rsqrt_exact(fro2) * JacobiTol<T>::scale();
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const T apq = a[p * N + q];
                if (!(rabs(apq) > tol)) continue;
                rotated = true;
                T c, s, napp, naqq;
                jacobi_angle<T>(a[p * N + p], a[q * N + q], apq, c, s, napp, naqq);
                a[p * N + p] = napp;
                a[q * N + q] = naqq;
                a[p * N + q] = zero;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if (j == p || j == q) continue;
                    // upper-triangle storage of the symmetric entries (j,p) and (j,q)
                    const int ip = (j < p) ? j * N + p : p * N + j;
                    const int iq = (j < q) ? j * N + q : q * N + j;
                    const T g = a[ip], k = a[iq];
                    a[ip] = rfma(c, g, -(s * k));
                    a[iq] = rfma(s, g, c * k);
                }
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const T g = v[j * N + p], k = v[j * N + q];
                    v[j * N + p] = rfma(c, g, -(s * k));
                    v[j * N + q] = rfma(s, g, c * k);
                }
            }
        }
        if (!rotated) break;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = a[i * N + i];
    if constexpr (!STATIC_SORT) {
    for (int i = 0; i < N - 1; ++i) {
        int k = i;
        for (int j = i + 1; j < N; ++j)
            if (w[j] < w[k]) k = j;
        if (k != i) {
            const T tw = w[i]; w[i] = w[k]; w[k] = tw;
            for (int r = 0; r < N; ++r) {
                const T tv = v[r * N + i]; v[r * N + i] = v[r * N + k]; v[r * N + k] = tv;
            }
        }
    }
    return;
    }
    // selection sort, ascending, ties to the lower index.  Static indices only (the arrays stay in registers): the
    // minimum of w[i..N) is tracked by value, the swap partner is found by comparing the (run-time) position with
    // every (compile-time) j
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        int k = i;
        T wk = w[i];
#pragma unroll
        for (int j = i + 1; j < N; ++j)
            if (w[j] < wk) { k = j; wk = w[j]; }
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
            if (k == j) {
                const T tw = w[i]; w[i] = w[j]; w[j] = tw;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const T tv = v[r * N + i]; v[r * N + i] = v[r * N + j]; v[r * N + j] = tv;
                }
            }
        }
    }
}

// Same eigen-solve for the large matrices (N = 12, 9) when only the NV eigenvectors of the
// smallest eigenvalues are wanted (EPnP: the 4-dimensional null-space basis, PnPsolver.cpp:379-382;
// MLPnP: the last right-singular vector, MLPnPsolver.cpp:488-489).  The matrix is kept as a
// packed upper triangle with compile-time indices (registers on the device), the (c, s) of
// every rotation is recorded, and the wanted eigenvectors are obtained by applying the recorded
// rotations in reverse order to unit vectors: V e_j = J_1 (J_2 ( ... (J_K e_j))).
// a: packed upper triangle, row-major, N(N+1)/2 entries (destroyed).  w: the NV smallest
// eigenvalues ascending; v: N x NV (row-major), column j = eigenvector of w[j].
__host__ __device__ constexpr int tri_idx(int N, int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

// Round-robin (tournament) ordering of the Jacobi pairs (mirrors oracle/orc_linalg.c tour_pair): M = N rounded
// up to even players, M-1 steps per sweep, M/2 disjoint pairs per step; for odd N the pair with the dummy
// player N is a bye.  p < q.
__host__ __device__ constexpr int tour_lo(int M, int t, int i)
{
    const int r = M - 1;
    const int a = (i == 0) ? r : (t + i) % r;
    const int b = (i == 0) ? (t % r) : (t - i + r) % r;
    return a < b ? a : b;
}
__host__ __device__ constexpr int tour_hi(int M, int t, int i)
{
    const int r = M - 1;
    const int a = (i == 0) ? r : (t + i) % r;
    const int b = (i == 0) ? (t % r) : (t - i + r) % r;
    return a < b ? b : a;
}
__host__ __device__ constexpr int tri_sym(int N, int x, int y) { return x < y ? tri_idx(N, x, y) : tri_idx(N, y, x); }

// One step = the rotations of M/2 disjoint pairs: (1) every pair's parameters from the current matrix and
// its diagonal block, (2) for every two pairs i < j the 2x2 block with one index in each -- pair i's
// rotation first, then pair j's (what applying rotation i and then rotation j to the whole matrix does to
// those four elements).  rec holds (c, s) per (sweep, step, pair); identity for skipped pairs and byes.
template <int N, int NV>
__host__ __device__ inline void jacobi_lowest(double* a, double* w, double* v, double2* rec /* kMaxSweepsRec*66 */)
{
    constexpr int M = N + (N & 1), Hh = M / 2, STEPS = M - 1, SLOTS = STEPS * Hh;
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) fro2 = rfma(a[tri_idx(N, i, j)], a[tri_idx(N, i, j)], fro2);
    const double tol = sqrt(fro2) * 0x1p-56;
    int sweeps = 0;
    for (int sweep = 0; sweep < kMaxSweepsRec; ++sweep) {
        bool rotated = false;
        double2* rs = rec + sweep * SLOTS;
#pragma unroll
        for (int t = 0; t < STEPS; ++t) {
            double C[Hh], S[Hh];
            bool rot[Hh];
#pragma unroll
            for (int i = 0; i < Hh; ++i) {
                constexpr int dummy = 0;
                (void)dummy;
                const int p = tour_lo(M, t, i), q = tour_hi(M, t, i);
                rot[i] = false; C[i] = 1.0; S[i] = 0.0;
                if (q < N) {
                    const double apq = a[tri_idx(N, p, q)];
                    if (fabs(apq) > tol) {
                        rot[i] = true;
                        rotated = true;
                        double napp, naqq;
                        jacobi_angle<double>(a[tri_idx(N, p, p)], a[tri_idx(N, q, q)], apq, C[i], S[i], napp, naqq);
                        a[tri_idx(N, p, p)] = napp;
                        a[tri_idx(N, q, q)] = naqq;
                        a[tri_idx(N, p, q)] = 0.0;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < Hh; ++i) {
                const int pi = tour_lo(M, t, i), qi = tour_hi(M, t, i);
#pragma unroll
                for (int j = i + 1; j < Hh; ++j) {
                    const int pj = tour_lo(M, t, j), qj = tour_hi(M, t, j);
                    if (rot[i]) {                              // rot[i] implies a real pair
                        {
                            const double g = a[tri_sym(N, pi, pj)], k = a[tri_sym(N, qi < N ? qi : pi, pj)];
                            a[tri_sym(N, pi, pj)] = rfma(C[i], g, -(S[i] * k));
                            a[tri_sym(N, qi < N ? qi : pi, pj)] = rfma(S[i], g, C[i] * k);
                        }
                        {
                            const double g = a[tri_sym(N, pi, qj)], k = a[tri_sym(N, qi < N ? qi : pi, qj)];
                            a[tri_sym(N, pi, qj)] = rfma(C[i], g, -(S[i] * k));
                            a[tri_sym(N, qi < N ? qi : pi, qj)] = rfma(S[i], g, C[i] * k);
                        }
                    }
                    if (rot[j]) {
                        {
                            const double g = a[tri_sym(N, pi, pj)], k = a[tri_sym(N, pi, qj)];
                            a[tri_sym(N, pi, pj)] = rfma(C[j], g, -(S[j] * k));
                            a[tri_sym(N, pi, qj)] = rfma(S[j], g, C[j] * k);
                        }
                        if (qi < N) {                          // a bye has one real member only
                            const double g = a[tri_sym(N, qi < N ? qi : pi, pj)], k = a[tri_sym(N, qi < N ? qi : pi, qj)];
                            a[tri_sym(N, qi < N ? qi : pi, pj)] = rfma(C[j], g, -(S[j] * k));
                            a[tri_sym(N, qi < N ? qi : pi, qj)] = rfma(S[j], g, C[j] * k);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < Hh; ++i) rs[t * Hh + i] = make_double2(C[i], S[i]);
        }
        if (!rotated) break;
        sweeps = sweep + 1;
    }
    // the NV smallest diagonal entries, ascending, ties to the lower index (static indexing only)
    double d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = a[tri_idx(N, i, i)];
    unsigned usedmask = 0u;
    int sel[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        int best = -1;
        double bv = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if ((usedmask >> i) & 1u) continue;
            if (best < 0 || d[i] < bv) { best = i; bv = d[i]; }
        }
        usedmask |= 1u << best;
        sel[k] = best;
        w[k] = bv;
    }
    double x[NV][N];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int i = 0; i < N; ++i) x[k][i] = (i == sel[k]) ? 1.0 : 0.0;
    for (int sweep = sweeps - 1; sweep >= 0; --sweep) {
        const double2* rs = rec + sweep * SLOTS;
#pragma unroll
        for (int t = STEPS - 1; t >= 0; --t) {
#pragma unroll
            for (int i = Hh - 1; i >= 0; --i) {
                const int p = tour_lo(M, t, i), q = tour_hi(M, t, i);
                const double2 cs = rs[t * Hh + i];
                if (q < N && cs.y != 0.0) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        const double xp = x[k][p], xq = x[k][q < N ? q : p];
                        x[k][p] = rfma(cs.x, xp, cs.y * xq);
                        x[k][q < N ? q : p] = rfma(cs.x, xq, -(cs.y * xp));
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < NV; ++k) v[i * NV + k] = x[k][i];
}

#ifdef __CUDACC__
// Warp-cooperative schedule of jacobi_lowest<N,NV> for the refine stage (one problem per CTA, warp 0): the
// M/2 rotation parameters of a step -- each a sqrt -> sqrt -> div chain -- are computed on M/2 lanes at once,
// the (M/2 choose 2) 2x2 blocks on as many lanes, the NV back-substitutions on NV lanes.  Every element sees
// the operations of the serial routine in the same order, so the result is bit-identical.
// a: packed upper triangle in SHARED memory (destroyed); w: NV smallest eigenvalues; v: N x NV (row-major)
// in shared memory; rec: kMaxSweepsRec*66 double2 in shared memory.
template <int N, int NV>
__device__ inline void jacobi_lowest_warp(double* a, double* w, double* v, double2* rec, int lane, long long* clk = nullptr)
{
    constexpr int M = N + (N & 1), Hh = M / 2, STEPS = M - 1, SLOTS = STEPS * Hh, NB = Hh * (Hh - 1) / 2;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NB <= 32 && Hh <= 32, "one lane per pair / per block");
    double fro2 = 0.0;
    if (lane == 0) {
        for (int i = 0; i < N; ++i)
            for (int j = i; j < N; ++j) fro2 = rfma(a[tri_idx(N, i, j)], a[tri_idx(N, i, j)], fro2);
    }
    fro2 = __shfl_sync(FULL, fro2, 0);
    const double tol = sqrt(fro2) * 0x1p-56;
    // this lane's block (bi < bj), fixed for the whole solve
    int bi = 0, bj = 1;
    if (lane < NB) {
        int rem = lane;
        while (rem >= Hh - 1 - bi) { rem -= Hh - 1 - bi; ++bi; }
        bj = bi + 1 + rem;
    }
    int sweeps = 0;
    for (int sweep = 0; sweep < kMaxSweepsRec; ++sweep) {
        bool rotated = false;
        double2* rs = rec + sweep * SLOTS;
        for (int t = 0; t < STEPS; ++t) {
            // (1) lane i < Hh: pair i
            int p = 0, q = N;
            double c = 1.0, s = 0.0;
            bool rot = false;
            if (lane < Hh) {
                p = tour_lo(M, t, lane); q = tour_hi(M, t, lane);
                if (q < N) {
                    const double apq = a[tri_idx(N, p, q)];
                    if (fabs(apq) > tol) {
                        rot = true;
                        double napp, naqq;
                        jacobi_angle<double>(a[tri_idx(N, p, p)], a[tri_idx(N, q, q)], apq, c, s, napp, naqq);
                        a[tri_idx(N, p, p)] = napp;          // diagonal blocks are private to their pair
                        a[tri_idx(N, q, q)] = naqq;
                        a[tri_idx(N, p, q)] = 0.0;
                    }
                }
                rs[t * Hh + lane] = make_double2(c, s);
            }
            const unsigned anyrot = __ballot_sync(FULL, rot);
            if (anyrot) {
                rotated = true;
                // (2) lane b < NB: block (bi, bj)
                const int pi = __shfl_sync(FULL, p, bi), qi = __shfl_sync(FULL, q, bi);
                const int pj = __shfl_sync(FULL, p, bj), qj = __shfl_sync(FULL, q, bj);
                const double ci = __shfl_sync(FULL, c, bi), si = __shfl_sync(FULL, s, bi);
                const double cj = __shfl_sync(FULL, c, bj), sj = __shfl_sync(FULL, s, bj);
                const bool roti = (anyrot >> bi) & 1u, rotj = (anyrot >> bj) & 1u;
                if (lane < NB && (roti || rotj)) {
                    const bool real_i = qi < N;                   // pair i is a bye when qi == N (then !roti)
                    const int qir = real_i ? qi : pi;
                    const int e00 = tri_sym(N, pi, pj), e01 = tri_sym(N, pi, qj);
                    const int e10 = tri_sym(N, qir, pj), e11 = tri_sym(N, qir, qj);
                    double x00 = a[e00], x01 = a[e01], x10 = a[e10], x11 = a[e11];
                    if (roti) {
                        const double g0 = x00, k0 = x10, g1 = x01, k1 = x11;
                        x00 = rfma(ci, g0, -(si * k0)); x10 = rfma(si, g0, ci * k0);
                        x01 = rfma(ci, g1, -(si * k1)); x11 = rfma(si, g1, ci * k1);
                    }
                    if (rotj) {
                        const double g0 = x00, k0 = x01;
                        x00 = rfma(cj, g0, -(sj * k0)); x01 = rfma(sj, g0, cj * k0);
                        if (real_i) {
                            const double g1 = x10, k1 = x11;
                            x10 = rfma(cj, g1, -(sj * k1)); x11 = rfma(sj, g1, cj * k1);
                        }
                    }
                    a[e00] = x00; a[e01] = x01;
                    if (real_i) { a[e10] = x10; a[e11] = x11; }
                }
            }
            __syncwarp();
        }
        if (!rotated) break;
        sweeps = sweep + 1;
    }
    __syncwarp();
    if (clk && lane == 0) { clk[0] = clock64(); clk[1] = sweeps; }     // diagnostic: end of the forward sweeps
    // the NV smallest diagonal entries, ascending, ties to the lower index
    int sel = -1;
    {
        unsigned usedmask = 0u;
        for (int k = 0; k < NV; ++k) {
            int best = -1;
            double bv = 0.0;
            for (int i = 0; i < N; ++i) {
                if ((usedmask >> i) & 1u) continue;
                const double di = a[tri_idx(N, i, i)];
                if (best < 0 || di < bv) { best = i; bv = di; }
            }
            usedmask |= 1u << best;
            if (lane == k) sel = best;
            if (lane == 0) w[k] = bv;
        }
    }
    if (lane < NV) {
        // back-application of the recorded rotations to e_sel: the vector lives in registers (the pairs of a step are
        // compile-time constants once the step loop is unrolled; only the sweep loop is dynamic) -- through shared memory
        // this phase was a third of the eigen-solve (40 k of 125 k cycles: two dependent LDS/STS round trips per rotation)
        double x[N];
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = (i == sel) ? 1.0 : 0.0;
        for (int sweep = sweeps - 1; sweep >= 0; --sweep) {
            const double2* rs = rec + sweep * SLOTS;
#pragma unroll
            for (int t = STEPS - 1; t >= 0; --t) {
                double2 cs[Hh];
#pragma unroll
                for (int i = 0; i < Hh; ++i) cs[i] = rs[t * Hh + i];
#pragma unroll
                for (int i = Hh - 1; i >= 0; --i) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int p = tour_lo(M, t, i), q = tour_hi(M, t, i);
                    if (q < N && cs[i].y != 0.0) {
                        const double xp = x[p], xq = x[q < N ? q : p];
                        x[p] = rfma(cs[i].x, xp, cs[i].y * xq);
                        x[q < N ? q : p] = rfma(cs[i].x, xq, -(cs[i].y * xp));
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) v[i * NV + lane] = x[i];
    }
    __syncwarp();
}
#endif  // __CUDACC__

// Orthonormal basis of the null space of an 8 x 12 matrix M (4-point EPnP) by Householder QR of
// A = M^T (12 x 8, row-major A[r*8+c], destroyed): null(M) = last four columns of Q = H0..H7 e_{8..11}.
// U4[r*4+i] = component r of basis vector i.  Operation order mirrors oracle/orc_linalg.c
// orc_nullspace_qr_d exactly (arithmetic contract).  2.4 kFLOP, fully unrolled: A stays in registers.
// ust: element stride of U4 (1 for a private array; blockDim.x when the basis lives in shared memory, one
// column of doubles per thread)
__host__ __device__ inline void nullspace_qr_8x12(double* A, double* U4, int ust = 1)
{
    double tau[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double s = 0.0;
#pragma unroll
        for (int r = k; r < 12; ++r) s = rfma(A[r * 8 + k], A[r * 8 + k], s);
        const double norm = sqrt(s);
        if (norm == 0.0) { tau[k] = 0.0; continue; }
        const double alpha = (A[k * 8 + k] > 0.0) ? -norm : norm;
        A[k * 8 + k] = A[k * 8 + k] - alpha;
        double vtv = 0.0;
#pragma unroll
        for (int r = k; r < 12; ++r) vtv = rfma(A[r * 8 + k], A[r * 8 + k], vtv);
        tau[k] = 2.0 / vtv;
#pragma unroll
        for (int j = k + 1; j < 8; ++j) {
            double d = 0.0;
#pragma unroll
            for (int r = k; r < 12; ++r) d = rfma(A[r * 8 + k], A[r * 8 + j], d);
            d = d * tau[k];
#pragma unroll
            for (int r = k; r < 12; ++r) A[r * 8 + j] = rfma(-d, A[r * 8 + k], A[r * 8 + j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double y[12];
#pragma unroll
        for (int r = 0; r < 12; ++r) y[r] = (r == 8 + i) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 7; k >= 0; --k) {
            if (tau[k] == 0.0) continue;
            double d = 0.0;
#pragma unroll
            for (int r = k; r < 12; ++r) d = rfma(A[r * 8 + k], y[r], d);
            d = d * tau[k];
#pragma unroll
            for (int r = k; r < 12; ++r) y[r] = rfma(-d, A[r * 8 + k], y[r]);
        }
#pragma unroll
        for (int r = 0; r < 12; ++r) U4[(r * 4 + i) * ust] = y[r];
    }
}

// One-sided (Hestenes) Jacobi: orthogonalises the columns of U (M x K), accumulates V (K x K).
template <int M, int K>
__host__ __device__ inline void onesided_jacobi(double* U, double* V)
{
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) V[i * K + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int i = 0; i < K - 1; ++i) {
#pragma unroll
            for (int j = i + 1; j < K; ++j) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int r = 0; r < M; ++r) {
                    const double ui = U[r * K + i], uj = U[r * K + j];
                    alpha = rfma(ui, ui, alpha);
                    beta = rfma(uj, uj, beta);
                    gamma = rfma(ui, uj, gamma);
                }
                if (!(gamma * gamma > (DBL_EPSILON * DBL_EPSILON) * (alpha * beta))) continue;   // |gamma| > eps sqrt(alpha beta)
                rotated = true;
                // rotation that zeroes gamma: tan(2t) = 2 gamma / (beta - alpha); two square roots, one division
                const double h = beta - alpha;
                const double b2 = gamma + gamma;
                const double rr = sqrt(rfma(h, h, b2 * b2));
                const double ah = fabs(h);
                const double uu = (rr + rr) * (rr + ah);
                const double ww = 1.0 / sqrt(uu);
                const double c = (rr + ah) * ww;
                double s = b2 * ww;
                if (h < 0.0) s = -s;
#pragma unroll
                for (int r = 0; r < M; ++r) {
                    const double ui = U[r * K + i], uj = U[r * K + j];
                    U[r * K + i] = rfma(c, ui, -(s * uj));
                    U[r * K + j] = rfma(s, ui, c * uj);
                }
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const double vi = V[r * K + i], vj = V[r * K + j];
                    V[r * K + i] = rfma(c, vi, -(s * vj));
                    V[r * K + j] = rfma(s, vi, c * vj);
                }
            }
        }
        if (!rotated) break;
    }
}

// Minimum-norm least squares through the one-sided Jacobi SVD; singular values not larger
// than sigma_max * K * eps count as zero (Eigen's SVDBase threshold, diagSize = K).
// Stands in for L.bdcSvd(ThinU|ThinV).solve(b) (PnPsolver.cpp:531,559,590).
template <int M, int K>
__host__ __device__ inline void svd_lstsq(const double* L, const double* b, double* x)
{
    double U[M * K], V[K * K], sig2[K], sig[K];
    for (int i = 0; i < M * K; ++i) U[i] = L[i];
    onesided_jacobi<M, K>(U, V);
    double smax = 0.0;
    for (int j = 0; j < K; ++j) {
        double s2 = 0.0;
        for (int r = 0; r < M; ++r) s2 = rfma(U[r * K + j], U[r * K + j], s2);
        sig2[j] = s2;
        sig[j] = sqrt(s2);
        if (sig[j] > smax) smax = sig[j];
    }
    const double thresh = smax * ((double)K * DBL_EPSILON);
    for (int r = 0; r < K; ++r) x[r] = 0.0;
    for (int j = 0; j < K; ++j) {
        if (!(sig[j] > thresh)) continue;
        double ub = 0.0;
        for (int r = 0; r < M; ++r) ub = rfma(U[r * K + j], b[r], ub);
        const double coef = ub / sig2[j];
        for (int r = 0; r < K; ++r) x[r] = rfma(coef, V[r * K + j], x[r]);
    }
}

// Least squares of an M x K system by Householder QR without pivoting (mirrors oracle/orc_linalg.c qr_lstsq):
// for full column rank the least-squares solution is unique, so this is what the SVD solve returns up to
// rounding at ~1/30 of the cost.  Returns false, x untouched, when min |R_kk| <= 1e-7 max |R_kk|.
template <int M, int K>
__host__ __device__ inline bool qr_lstsq(const double* L, const double* b, double* x)
{
    double A[M * K], bb[M], rd[K];
#pragma unroll
    for (int i = 0; i < M * K; ++i) A[i] = L[i];
#pragma unroll
    for (int i = 0; i < M; ++i) bb[i] = b[i];
#pragma unroll
    for (int c = 0; c < K; ++c) {
        double s = 0.0;
#pragma unroll
        for (int r = c; r < M; ++r) s = rfma(A[r * K + c], A[r * K + c], s);
        const double norm = sqrt(s);
        if (norm == 0.0) return false;
        const double alpha = (A[c * K + c] > 0.0) ? -norm : norm;
        A[c * K + c] = A[c * K + c] - alpha;
        double vtv = 0.0;
#pragma unroll
        for (int r = c; r < M; ++r) vtv = rfma(A[r * K + c], A[r * K + c], vtv);
        const double tau = 2.0 / vtv;
#pragma unroll
        for (int j = c + 1; j < K; ++j) {
            double d = 0.0;
#pragma unroll
            for (int r = c; r < M; ++r) d = rfma(A[r * K + c], A[r * K + j], d);
            d = d * tau;
#pragma unroll
            for (int r = c; r < M; ++r) A[r * K + j] = rfma(-d, A[r * K + c], A[r * K + j]);
        }
        double d = 0.0;
#pragma unroll
        for (int r = c; r < M; ++r) d = rfma(A[r * K + c], bb[r], d);
        d = d * tau;
#pragma unroll
        for (int r = c; r < M; ++r) bb[r] = rfma(-d, A[r * K + c], bb[r]);
        rd[c] = alpha;
    }
    double rmax = 0.0, rmin = fabs(rd[0]);
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const double a = fabs(rd[c]);
        if (a > rmax) rmax = a;
        if (a < rmin) rmin = a;
    }
    if (!(rmin > rmax * 1e-7)) return false;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        double sum = 0.0;
#pragma unroll
        for (int j = i + 1; j < K; ++j) sum = rfma(A[i * K + j], x[j], sum);
        x[i] = (bb[i] - sum) / rd[i];
    }
    return true;
}

// rank-deficient fallback kept out of line: it is (almost) never taken and its unrolled Jacobi sweeps would
// otherwise sit in the middle of the hot instruction stream
template <int M, int K>
__host__ __device__ __noinline__ void svd_lstsq_cold(const double* L, const double* b, double* x)
{
    svd_lstsq<M, K>(L, b, x);
}

// the least-squares solve of find_betas_approx_{1,2,3} (PnPsolver.cpp:531,559,590)
template <int M, int K>
__host__ __device__ inline void lstsq(const double* L, const double* b, double* x)
{
    if (!qr_lstsq<M, K>(L, b, x)) svd_lstsq_cold<M, K>(L, b, x);
}

// closed-form cofactor inverse (Matrix3d::inverse(), PnPsolver.cpp:331); singular => inf/NaN
__host__ __device__ inline void inv3(const double* m, double* out)
{
    const double c00 = rfma(m[4], m[8], -(m[5] * m[7]));
    const double c01 = rfma(m[5], m[6], -(m[3] * m[8]));
    const double c02 = rfma(m[3], m[7], -(m[4] * m[6]));
    const double c10 = rfma(m[2], m[7], -(m[1] * m[8]));
    const double c11 = rfma(m[0], m[8], -(m[2] * m[6]));
    const double c12 = rfma(m[1], m[6], -(m[0] * m[7]));
    const double c20 = rfma(m[1], m[5], -(m[2] * m[4]));
    const double c21 = rfma(m[2], m[3], -(m[0] * m[5]));
    const double c22 = rfma(m[0], m[4], -(m[1] * m[3]));
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = rdiv(1.0, det);
    out[0] = c00 * id; out[1] = c10 * id; out[2] = c20 * id;
    out[3] = c01 * id; out[4] = c11 * id; out[5] = c21 * id;
    out[6] = c02 * id; out[7] = c12 * id; out[8] = c22 * id;
}

__host__ __device__ inline double det3(const double* R)
{
    return R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
           R[2] * (R[3] * R[7] - R[4] * R[6]);
}

// Eigen::Quaternion::toRotationMatrix (PnPsolver.cpp:478, Sim3Solver.cpp:248), no normalisation
template <typename T>
__host__ __device__ inline void quat_to_rot(T w, T x, T y, T z, T* R)
{
    const T two = T(2), one = T(1);
    const T tx = two * x, ty = two * y, tz = two * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x;
    const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = one - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = one - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = one - (txx + tyy);
}

// orthogonal polar factor U V^T of a 3x3 (JacobiSVD, MLPnPsolver.cpp:511-512,570-571)
__host__ __device__ inline void polar3(const double* a, double* r)
{
    double U[9], V[9], inv[3];
    for (int i = 0; i < 9; ++i) U[i] = a[i];
    onesided_jacobi<3, 3>(U, V);
    for (int j = 0; j < 3; ++j) {
        double s2 = 0.0;
        for (int i = 0; i < 3; ++i) s2 = rfma(U[i * 3 + j], U[i * 3 + j], s2);
        inv[j] = 1.0 / sqrt(s2);
    }
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int j = 0; j < 3; ++j) acc += (U[i * 3 + j] * inv[j]) * V[c * 3 + j];
            r[i * 3 + c] = acc;
        }
}

// rank of a 3x3 by full-pivot Householder QR with Eigen's thresholds
// (FullPivHouseholderQR<Matrix3d>::rank(), MLPnPsolver.cpp:347,354)
__host__ __device__ inline int rank3_fullpiv(const double* a_in)
{
    double a[9];
    for (int i = 0; i < 9; ++i) a[i] = a_in[i];
    double diag[3] = {0.0, 0.0, 0.0};
    int nonzero = 3;
    double maxpivot = 0.0, biggest = 0.0;
    const double precision = DBL_EPSILON * 3.0;
    for (int k = 0; k < 3; ++k) {
        int pr = k, pc = k;
        double big = -1.0;
        for (int c = k; c < 3; ++c)
            for (int r = k; r < 3; ++r)
                if (fabs(a[r * 3 + c]) > big) { big = fabs(a[r * 3 + c]); pr = r; pc = c; }
        if (k == 0) biggest = big;
        if (!(big > biggest * precision)) { nonzero = k; break; }
        if (pr != k)
            for (int c = 0; c < 3; ++c) { const double t = a[k * 3 + c]; a[k * 3 + c] = a[pr * 3 + c]; a[pr * 3 + c] = t; }
        if (pc != k)
            for (int r = 0; r < 3; ++r) { const double t = a[r * 3 + k]; a[r * 3 + k] = a[r * 3 + pc]; a[r * 3 + pc] = t; }
        double tail2 = 0.0;
        for (int r = k + 1; r < 3; ++r) tail2 = rfma(a[r * 3 + k], a[r * 3 + k], tail2);
        const double c0 = a[k * 3 + k];
        double beta, tau;
        double vv[3] = {0.0, 0.0, 0.0};
        if (tail2 <= DBL_MIN) {
            beta = c0; tau = 0.0;
        } else {
            beta = sqrt(c0 * c0 + tail2);
            if (c0 >= 0.0) beta = -beta;
            for (int r = k + 1; r < 3; ++r) vv[r] = a[r * 3 + k] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        vv[k] = 1.0;
        diag[k] = beta;
        if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
        for (int c = k + 1; c < 3; ++c) {
            double dot = 0.0;
            for (int r = k; r < 3; ++r) dot = rfma(vv[r], a[r * 3 + c], dot);
            for (int r = k; r < 3; ++r) a[r * 3 + c] -= tau * vv[r] * dot;
        }
    }
    int rank = 0;
    const double thr = maxpivot * (DBL_EPSILON * 3.0);
    for (int i = 0; i < nonzero; ++i)
        if (fabs(diag[i]) > thr) ++rank;
    return rank;
}

// LDL^T with diagonal pivoting, 6x6 (Eigen::LDLT::solve, MLPnPsolver.cpp:705-706)
__host__ __device__ inline void ldlt6_solve(const double* a_in, const double* g, double* x)
{
    constexpr int N = 6;
    double a[36];
    int perm[N];
    for (int i = 0; i < 36; ++i) a[i] = a_in[i];
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double big = fabs(a[k * N + k]);
        for (int i = k + 1; i < N; ++i)
            if (fabs(a[i * N + i]) > big) { big = fabs(a[i * N + i]); piv = i; }
        if (piv != k) {
            for (int c = 0; c < N; ++c) { const double t = a[k * N + c]; a[k * N + c] = a[piv * N + c]; a[piv * N + c] = t; }
            for (int r = 0; r < N; ++r) { const double t = a[r * N + k]; a[r * N + k] = a[r * N + piv]; a[r * N + piv] = t; }
            const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double d = a[k * N + k];
        if (!(fabs(d) > DBL_MIN)) continue;
        for (int i = k + 1; i < N; ++i) a[i * N + k] = a[i * N + k] / d;
        for (int i = k + 1; i < N; ++i)
            for (int j = k + 1; j <= i; ++j) {
                a[i * N + j] -= a[i * N + k] * d * a[j * N + k];
                a[j * N + i] = a[i * N + j];
            }
    }
    double y[N];
    for (int i = 0; i < N; ++i) y[i] = g[perm[i]];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j) y[i] = rfma(-a[i * N + j], y[j], y[i]);
    for (int i = 0; i < N; ++i) {
        const double d = a[i * N + i];
        y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0;
    }
    for (int i = N - 1; i >= 0; --i)
        for (int j = i + 1; j < N; ++j) y[i] = rfma(-a[j * N + i], y[j], y[i]);
    for (int i = 0; i < N; ++i) x[perm[i]] = y[i];
}

}  // namespace rsac
