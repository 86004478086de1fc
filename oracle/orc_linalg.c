/*
 * oracle/orc_linalg.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Small dense solves that the reference delegates to Eigen (un-vendored, no
 * pinned version: CMakeLists.txt:49).  Each routine restates the published
 * algorithm class of the Eigen call it stands in for; PARITY UNPINNED against
 * an Eigen-built binary.  Only + - * / sqrt, fixed operation
 * order: the CUDA kernels implement the same sequences (arithmetic
 * contract) so that 4-point EPnP hypotheses are bit-identical on both sides.
 */
#include <math.h>
#include <float.h>
#include <string.h>
#include "orc.h"

/* Algorithmic FLOP counter (bench.py's roofline numerator for the solver kernels).
 * Convention of SURVEY 8(d): + - * / sqrt count 1 each; comparisons, negations and copies 0. */
__thread long long orc_flops = 0;
#define FL(n) (orc_flops += (n))
long long orc_flops_take(void) { long long v = orc_flops; orc_flops = 0; return v; }

/* Arithmetic contract: only + - * / sqrt and EXPLICIT fused multiply-adds, in a fixed order.  Every fma below
 * has a twin in csrc/linalg.cuh (rfma); compilers are forbidden to contract anything else (-ffp-contract=off,
 * nvcc -fmad=false).  fma()/fmaf() are correctly rounded whether they compile to a vfmadd or call libm. */
#define ORC_FMA(a, b, c) _Generic((a), float: fmaf, default: fma)((a), (b), (c))

#define ORC_MAX_SWEEPS 30

#define ORC_MAX_SWEEPS_REC 12

/* ---- cyclic Jacobi on the upper triangle, rotations below ||a||_F * 2^-56 (2^-27) skipped ----
 * Rotation of the 2x2 block [app apq; apq aqq] (arithmetic contract):
 *   h = aqq - app, b = 2 apq, r = sqrt(h*h + b*b), u = 2r (r + |h|), w = 1/sqrt(u),
 *   c = (r + |h|) w  (= sqrt((1 + |h|/r)/2)),  s = b w with the sign of h,
 *   new diagonal = (app+aqq)/2 -/+ r/2.
 * Two square roots and ONE division, the division last: the dependent chain is
 * sqrt -> sqrt -> div instead of the textbook sqrt -> div -> sqrt -> div. */
#define ORC_ANGLE(T, SQRT, FABS, HALF, FOUR, ZERO)                                              \
    const T h = aqq - app;                                                                      \
    const T b2 = apq + apq;                                                                     \
    const T r = SQRT(ORC_FMA(h, h, b2 * b2));                                                   \
    const T ah = FABS(h);                                                                       \
    const T uu = (r + r) * (r + ah);                                                            \
    const T ww = (HALF + HALF) / SQRT(uu);                                                      \
    const T c = (r + ah) * ww;                                                                  \
    const T s0 = b2 * ww;                                                                       \
    const T m = HALF * (app + aqq);                                                             \
    const T hr = HALF * r;                                                                      \
    T s, napp, naqq;                                                                            \
    if (h < ZERO) { s = -s0; napp = m + hr; naqq = m - hr; }                                    \
    else          { s = s0;  napp = m - hr; naqq = m + hr; }

#define ORC_JACOBI_IMPL(NAME, T, SQRT, FABS, TOLSCALE, ONE, HALF, FOUR, ZERO)                   \
void NAME(int n, T *a, T *w, T *v)                                                              \
{                                                                                               \
    for (int i = 0; i < n; ++i)                                                                 \
        for (int j = 0; j < n; ++j) v[i * n + j] = (i == j) ? ONE : ZERO;                       \
    T fro2 = ZERO;                                                                              \
    for (int i = 0; i < n; ++i)                                                                 \
        for (int j = i; j < n; ++j) fro2 = ORC_FMA(a[i * n + j], a[i * n + j], fro2);                        \
    FL(n * (n + 1) + 2);                                                                        \
    const T tol = SQRT(fro2) * TOLSCALE;                                                        \
    for (int sweep = 0; sweep < ORC_MAX_SWEEPS; ++sweep) {                                      \
        int rotated = 0;                                                                        \
        for (int p = 0; p < n - 1; ++p) {                                                       \
            for (int q = p + 1; q < n; ++q) {                                                   \
                const T apq = a[p * n + q];                                                     \
                if (!(FABS(apq) > tol)) continue;                                               \
                rotated = 1;                                                                    \
                FL(19 + 6 * (n - 2) + 6 * n);                                                   \
                const T app = a[p * n + p], aqq = a[q * n + q];                                 \
                ORC_ANGLE(T, SQRT, FABS, HALF, FOUR, ZERO)                                      \
                a[p * n + p] = napp;                                                            \
                a[q * n + q] = naqq;                                                            \
                a[p * n + q] = ZERO;                                                            \
                for (int j = 0; j < n; ++j) {                                                   \
                    if (j == p || j == q) continue;                                             \
                    const int ip = (j < p) ? j * n + p : p * n + j;                             \
                    const int iq = (j < q) ? j * n + q : q * n + j;                             \
                    const T g = a[ip], k = a[iq];                                               \
                    a[ip] = ORC_FMA(c, g, -(s * k));                                                      \
                    a[iq] = ORC_FMA(s, g, c * k);                                                      \
                }                                                                               \
                for (int j = 0; j < n; ++j) {                                                   \
                    const T g = v[j * n + p], k = v[j * n + q];                                 \
                    v[j * n + p] = ORC_FMA(c, g, -(s * k));                                               \
                    v[j * n + q] = ORC_FMA(s, g, c * k);                                               \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
        if (!rotated) break;                                                                    \
    }                                                                                           \
    for (int i = 0; i < n; ++i) w[i] = a[i * n + i];                                            \
    /* ascending order, stable selection sort (SelfAdjointEigenSolver sorts ascending) */       \
    for (int i = 0; i < n - 1; ++i) {                                                           \
        int k = i;                                                                              \
        for (int j = i + 1; j < n; ++j)                                                         \
            if (w[j] < w[k]) k = j;                                                             \
        if (k != i) {                                                                           \
            T tw = w[i]; w[i] = w[k]; w[k] = tw;                                                \
            for (int r = 0; r < n; ++r) {                                                       \
                T tv = v[r * n + i]; v[r * n + i] = v[r * n + k]; v[r * n + k] = tv;            \
            }                                                                                   \
        }                                                                                       \
    }                                                                                           \
}

ORC_JACOBI_IMPL(orc_jacobi_eig_d, double, sqrt, fabs, 0x1p-56, 1.0, 0.5, 4.0, 0.0)
ORC_JACOBI_IMPL(orc_jacobi_eig_f, float, sqrtf, fabsf, 0x1p-27f, 1.0f, 0.5f, 4.0f, 0.0f)

/* The same eigen-solve when only the nv eigenvectors of the smallest eigenvalues are wanted
 * (EPnP null-space basis, PnPsolver.cpp:379-382; MLPnP last singular vector, MLPnPsolver.cpp:488-489).
 * The (c, s) of every rotation of at most ORC_MAX_SWEEPS_REC sweeps is recorded and the wanted
 * eigenvectors are formed by applying the rotations in reverse order to unit vectors.
 * a: n x n row-major, upper triangle read, destroyed.  w: nv eigenvalues ascending.
 * v: n x nv row-major. */
/* Round-robin (tournament) ordering of the Jacobi pairs: m = n rounded up to even players, m-1 steps per
 * sweep, m/2 disjoint pairs per step (circle method: player m-1 is fixed, the others rotate).  Every
 * unordered pair appears exactly once per sweep.  For odd n the pair that contains the dummy player n is a
 * bye.  Disjoint pairs are what lets the refine kernel compute the m/2 rotation parameters of a step on
 * m/2 lanes at once; the arithmetic of every element is specified below and is the same in all three
 * implementations (this one, csrc/linalg.cuh jacobi_lowest and jacobi_lowest_warp). */
static void tour_pair(int m, int t, int i, int *p, int *q)
{
    const int r = m - 1;
    int a, b;
    if (i == 0) { a = r; b = t % r; }
    else { a = (t + i) % r; b = (t - i + r) % r; }
    if (a < b) { *p = a; *q = b; } else { *p = b; *q = a; }
}

#define ORC_E(x, y) a[((x) < (y) ? (x) : (y)) * n + ((x) < (y) ? (y) : (x))]

/* One step: (1) the rotation of every pair from the current matrix (pairs are disjoint, so no pair's
 * (app, aqq, apq) is touched by another pair's rotation) and its diagonal block; (2) for every two pairs
 * i < j the 2x2 block of elements with one index in each: first pair i's rotation mixes the elements that
 * share the other index, then pair j's -- which is exactly what applying rotation i and then rotation j to
 * the whole matrix does to those four elements. */
void orc_jacobi_lowest_d(int n, int nv, double *a, double *w, double *v)
{
    const int m = n + (n & 1), h = m / 2, steps = m - 1, slots = steps * h;
    static __thread double rc[ORC_MAX_SWEEPS_REC * 66], rs[ORC_MAX_SWEEPS_REC * 66];
    double fro2 = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) fro2 = ORC_FMA(a[i * n + j], a[i * n + j], fro2);
    FL(n * (n + 1) + 2);
    const double tol = sqrt(fro2) * 0x1p-56;
    int sweeps = 0;
    for (int sweep = 0; sweep < ORC_MAX_SWEEPS_REC; ++sweep) {
        int rotated = 0;
        for (int t = 0; t < steps; ++t) {
            int P[6], Q[6], rot[6];
            double C[6], S[6];
            for (int i = 0; i < h; ++i) {
                tour_pair(m, t, i, &P[i], &Q[i]);
                rot[i] = 0; C[i] = 1.0; S[i] = 0.0;
                if (Q[i] >= n) continue;                       /* bye */
                const int p = P[i], q = Q[i];
                const double apq = a[p * n + q];
                if (!(fabs(apq) > tol)) continue;
                rot[i] = 1;
                rotated = 1;
                FL(19 + 6 * (n - 2));
                const double app = a[p * n + p], aqq = a[q * n + q];
                ORC_ANGLE(double, sqrt, fabs, 0.5, 4.0, 0.0)
                a[p * n + p] = napp;
                a[q * n + q] = naqq;
                a[p * n + q] = 0.0;
                C[i] = c; S[i] = s;
            }
            /* blocks; a bye (odd n: pair 0 = {player, dummy}) still has one real member whose elements are
             * mixed by the other pairs' rotations */
            for (int i = 0; i < h; ++i) {
                const int ni = (Q[i] >= n) ? 1 : 2;            /* real members of pair i: P[i] (and Q[i]) */
                for (int j = i + 1; j < h; ++j) {
                    if (rot[i]) {                              /* rot[i] implies a real pair */
                        const int ys[2] = {P[j], Q[j]};
                        for (int e = 0; e < 2; ++e) {
                            const double g = ORC_E(P[i], ys[e]), k = ORC_E(Q[i], ys[e]);
                            ORC_E(P[i], ys[e]) = ORC_FMA(C[i], g, -(S[i] * k));
                            ORC_E(Q[i], ys[e]) = ORC_FMA(S[i], g, C[i] * k);
                        }
                    }
                    if (rot[j]) {
                        const int xs[2] = {P[i], Q[i]};
                        for (int e = 0; e < ni; ++e) {
                            const double g = ORC_E(xs[e], P[j]), k = ORC_E(xs[e], Q[j]);
                            ORC_E(xs[e], P[j]) = ORC_FMA(C[j], g, -(S[j] * k));
                            ORC_E(xs[e], Q[j]) = ORC_FMA(S[j], g, C[j] * k);
                        }
                    }
                }
            }
            for (int i = 0; i < h; ++i) {
                rc[sweep * slots + t * h + i] = C[i];
                rs[sweep * slots + t * h + i] = S[i];
            }
        }
        if (!rotated) break;
        sweeps = sweep + 1;
    }
    /* the nv smallest diagonal entries, ascending, ties to the lower index */
    unsigned usedmask = 0u;
    int sel[12];
    for (int k = 0; k < nv; ++k) {
        int best = -1;
        double bv = 0.0;
        for (int i = 0; i < n; ++i) {
            if ((usedmask >> i) & 1u) continue;
            if (best < 0 || a[i * n + i] < bv) { best = i; bv = a[i * n + i]; }
        }
        usedmask |= 1u << best;
        sel[k] = best;
        w[k] = bv;
    }
    /* eigenvectors: the recorded rotations applied in reverse to unit vectors (rotations of one step act on
     * disjoint index pairs, so their order within the step is immaterial) */
    double x[12][12];
    for (int k = 0; k < nv; ++k)
        for (int i = 0; i < n; ++i) x[k][i] = (i == sel[k]) ? 1.0 : 0.0;
    for (int sweep = sweeps - 1; sweep >= 0; --sweep)
        for (int t = steps - 1; t >= 0; --t)
            for (int i = h - 1; i >= 0; --i) {
                const double c = rc[sweep * slots + t * h + i], s = rs[sweep * slots + t * h + i];
                if (s != 0.0) {
                    int p, q;
                    tour_pair(m, t, i, &p, &q);
                    FL(6 * nv);
                    for (int k = 0; k < nv; ++k) {
                        const double xp = x[k][p], xq = x[k][q];
                        x[k][p] = ORC_FMA(c, xp, s * xq);
                        x[k][q] = ORC_FMA(c, xq, -(s * xp));
                    }
                }
            }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < nv; ++k) v[i * nv + k] = x[k][i];
}
#undef ORC_E

/* ---- one-sided (Hestenes) Jacobi SVD: columns of U orthogonalised, V accumulated ---- */
static void onesided_jacobi(int m, int k, double *U /* m*k */, double *V /* k*k */)
{
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) V[i * k + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < ORC_MAX_SWEEPS; ++sweep) {
        int rotated = 0;
        for (int i = 0; i < k - 1; ++i) {
            for (int j = i + 1; j < k; ++j) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int r = 0; r < m; ++r) {
                    const double ui = U[r * k + i], uj = U[r * k + j];
                    alpha = ORC_FMA(ui, ui, alpha);
                    beta = ORC_FMA(uj, uj, beta);
                    gamma = ORC_FMA(ui, uj, gamma);
                }
                FL(6 * m + 3);
                /* |gamma| > eps sqrt(alpha beta), tested on the squares (no square root) */
                if (!(gamma * gamma > (DBL_EPSILON * DBL_EPSILON) * (alpha * beta))) continue;
                rotated = 1;
                FL(13 + 6 * m + 6 * k);
                /* rotation that zeroes gamma: tan(2t) = 2 gamma / (beta - alpha); same chain as ORC_ANGLE:
                 * two square roots and one division */
                const double h = beta - alpha;
                const double b2 = gamma + gamma;
                const double rr = sqrt(ORC_FMA(h, h, b2 * b2));
                const double ah = fabs(h);
                const double uu = (rr + rr) * (rr + ah);
                const double ww = 1.0 / sqrt(uu);
                const double c = (rr + ah) * ww;
                double s = b2 * ww;
                if (h < 0.0) s = -s;
                for (int r = 0; r < m; ++r) {
                    const double ui = U[r * k + i], uj = U[r * k + j];
                    U[r * k + i] = ORC_FMA(c, ui, -(s * uj));
                    U[r * k + j] = ORC_FMA(s, ui, c * uj);
                }
                for (int r = 0; r < k; ++r) {
                    const double vi = V[r * k + i], vj = V[r * k + j];
                    V[r * k + i] = ORC_FMA(c, vi, -(s * vj));
                    V[r * k + j] = ORC_FMA(s, vi, c * vj);
                }
            }
        }
        if (!rotated) break;
    }
}

/* PnPsolver.cpp:531,559,590: x = L.bdcSvd(ThinU|ThinV).solve(b).  For < 16
 * columns Eigen's BDCSVD delegates to JacobiSVD; solve() is the minimum-norm
 * least-squares solution with singular values <= sigma_max * diagSize * eps
 * treated as zero. */
long long orc_cond_hist[8][16];
void orc_lstsq_cond_hist(long long *out /* 8*16 */, int reset)
{
    memcpy(out, orc_cond_hist, sizeof(orc_cond_hist));
    if (reset) memset(orc_cond_hist, 0, sizeof(orc_cond_hist));
}

void orc_svd_lstsq_d(int m, int k, const double *L, const double *b, double *x)
{
    double U[8 * 6], V[6 * 6], sig2[6], sig[6];
    memcpy(U, L, sizeof(double) * (size_t)(m * k));
    onesided_jacobi(m, k, U, V);
    double smax = 0.0;
    for (int j = 0; j < k; ++j) {
        double s2 = 0.0;
        for (int r = 0; r < m; ++r) s2 = ORC_FMA(U[r * k + j], U[r * k + j], s2);
        sig2[j] = s2;
        sig[j] = sqrt(s2);
        if (sig[j] > smax) smax = sig[j];
    }
    FL(k * (2 * m + 1) + 2);
    {   /* diagnostic histogram of log10(sigma_min / sigma_max) per column count (orc_lstsq_cond_hist) */
        double smin = smax;
        for (int j = 0; j < k; ++j) if (sig[j] < smin) smin = sig[j];
        int bin = 0;
        if (smax > 0.0 && smin > 0.0) { bin = (int)(-log10(smin / smax)); if (bin < 0) bin = 0; if (bin > 15) bin = 15; } else bin = 15;
        __atomic_fetch_add(&orc_cond_hist[k][bin], 1, __ATOMIC_RELAXED);
    }
    const double thresh = smax * ((double)k * DBL_EPSILON);
    for (int r = 0; r < k; ++r) x[r] = 0.0;
    for (int j = 0; j < k; ++j) {
        if (!(sig[j] > thresh)) continue;
        double ub = 0.0;
        for (int r = 0; r < m; ++r) ub = ORC_FMA(U[r * k + j], b[r], ub);
        const double coef = ub / sig2[j];
        FL(2 * m + 1 + 2 * k);
        for (int r = 0; r < k; ++r) x[r] = ORC_FMA(coef, V[r * k + j], x[r]);
    }
}

/* Least squares of an m x k system (m <= 8, k <= 6) by Householder QR without pivoting.  For a system of
 * full column rank the least-squares solution is unique, so this is the value L.bdcSvd().solve(b) returns
 * (PnPsolver.cpp:531,559,590) up to rounding -- at ~1/30 of the cost of a Jacobi SVD.  Returns 0, x
 * untouched, when the triangular factor is numerically rank deficient (min |R_kk| <= 1e-7 max |R_kk|; the
 * cfg4 systems have sigma_min/sigma_max >= 1e-5): the caller then takes the minimum-norm SVD solve, which
 * is what the reference's call means for a rank-deficient system.  Mirrored by csrc/linalg.cuh qr_lstsq. */
static int qr_lstsq(int m, int k, const double *L, const double *b, double *x)
{
    double A[8 * 6], bb[8], rd[6] = {0, 0, 0, 0, 0, 0};   /* k >= 1 always; the initialiser only quiets -Wmaybe-uninitialized */
    memcpy(A, L, sizeof(double) * (size_t)(m * k));
    memcpy(bb, b, sizeof(double) * (size_t)m);
    for (int c = 0; c < k; ++c) {
        double s = 0.0;
        for (int r = c; r < m; ++r) s = ORC_FMA(A[r * k + c], A[r * k + c], s);
        const double norm = sqrt(s);
        if (norm == 0.0) return 0;
        const double alpha = (A[c * k + c] > 0.0) ? -norm : norm;
        A[c * k + c] = A[c * k + c] - alpha;
        double vtv = 0.0;
        for (int r = c; r < m; ++r) vtv = ORC_FMA(A[r * k + c], A[r * k + c], vtv);
        const double tau = 2.0 / vtv;
        for (int j = c + 1; j < k; ++j) {
            double d = 0.0;
            for (int r = c; r < m; ++r) d = ORC_FMA(A[r * k + c], A[r * k + j], d);
            d = d * tau;
            for (int r = c; r < m; ++r) A[r * k + j] = ORC_FMA(-d, A[r * k + c], A[r * k + j]);
        }
        double d = 0.0;
        for (int r = c; r < m; ++r) d = ORC_FMA(A[r * k + c], bb[r], d);
        d = d * tau;
        for (int r = c; r < m; ++r) bb[r] = ORC_FMA(-d, A[r * k + c], bb[r]);
        rd[c] = alpha;
        FL(4 * (m - c) + 3 + (k - c) * (4 * (m - c) + 1));
    }
    double rmax = 0.0, rmin = fabs(rd[0]);
    for (int c = 0; c < k; ++c) {
        const double a = fabs(rd[c]);
        if (a > rmax) rmax = a;
        if (a < rmin) rmin = a;
    }
    if (!(rmin > rmax * 1e-7)) return 0;
    for (int i = k - 1; i >= 0; --i) {
        double sum = 0.0;
        for (int j = i + 1; j < k; ++j) sum = ORC_FMA(A[i * k + j], x[j], sum);
        x[i] = (bb[i] - sum) / rd[i];
    }
    FL(k * k + k);
    return 1;
}

/* the least-squares solve of find_betas_approx_{1,2,3}: QR, minimum-norm SVD only if rank deficient */
void orc_lstsq_d(int m, int k, const double *L, const double *b, double *x)
{
    if (!qr_lstsq(m, k, L, b, x)) orc_svd_lstsq_d(m, k, L, b, x);
}

/* PnPsolver.cpp:331: CC.inverse() -- closed-form cofactor inverse, no pivoting,
 * singular => inf/NaN propagate. */
void orc_inv3_d(const double m[9], double out[9])
{
    const double c00 = ORC_FMA(m[4], m[8], -(m[5] * m[7]));
    const double c01 = ORC_FMA(m[5], m[6], -(m[3] * m[8]));
    const double c02 = ORC_FMA(m[3], m[7], -(m[4] * m[6]));
    const double c10 = ORC_FMA(m[2], m[7], -(m[1] * m[8]));
    const double c11 = ORC_FMA(m[0], m[8], -(m[2] * m[6]));
    const double c12 = ORC_FMA(m[1], m[6], -(m[0] * m[7]));
    const double c20 = ORC_FMA(m[1], m[5], -(m[2] * m[4]));
    const double c21 = ORC_FMA(m[2], m[3], -(m[0] * m[5]));
    const double c22 = ORC_FMA(m[0], m[4], -(m[1] * m[3]));
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    FL(27 + 5 + 1 + 9);
    out[0] = c00 * id; out[1] = c10 * id; out[2] = c20 * id;
    out[3] = c01 * id; out[4] = c11 * id; out[5] = c21 * id;
    out[6] = c02 * id; out[7] = c12 * id; out[8] = c22 * id;
}

/* MLPnPsolver.cpp:511-512,570-571: JacobiSVD(tmp).matrixU() * matrixV()^T, the
 * orthogonal polar factor.  A V = U_work (orthogonal columns of norm sigma_j),
 * so U V^T = sum_j (U_work_j / sigma_j) V_j^T. */
void orc_polar3_d(const double a[9], double r[9])
{
    double U[9], V[9], inv[3];
    memcpy(U, a, sizeof(U));
    onesided_jacobi(3, 3, U, V);
    for (int j = 0; j < 3; ++j) {
        double s2 = 0.0;
        for (int i = 0; i < 3; ++i) s2 = ORC_FMA(U[i * 3 + j], U[i * 3 + j], s2);
        inv[j] = 1.0 / sqrt(s2);
    }
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int j = 0; j < 3; ++j) acc += (U[i * 3 + j] * inv[j]) * V[c * 3 + j];
            r[i * 3 + c] = acc;
        }
}

/* MLPnPsolver.cpp:347,354: FullPivHouseholderQR<Matrix3d>(planarTest).rank().
 * Eigen: at step k pick the largest |entry| of the trailing corner, stop early
 * when it is much smaller than the first one (eps*size), Householder the
 * column; rank = #{ |r_ii| > |maxpivot| * eps * 3 }. */
int orc_rank3_fullpiv_d(const double a_in[9])
{
    double a[9];
    memcpy(a, a_in, sizeof(a));
    double diag[3] = {0.0, 0.0, 0.0};
    int nonzero = 3;
    double maxpivot = 0.0, biggest = 0.0;
    const double precision = DBL_EPSILON * 3.0;
    for (int k = 0; k < 3; ++k) {
        int pr = k, pc = k;
        double big = -1.0;
        for (int c = k; c < 3; ++c)      /* Eigen visits column-major */
            for (int r = k; r < 3; ++r)
                if (fabs(a[r * 3 + c]) > big) { big = fabs(a[r * 3 + c]); pr = r; pc = c; }
        if (k == 0) biggest = big;
        if (!(big > biggest * precision)) { nonzero = k; break; }   /* isMuchSmallerThan */
        if (pr != k)
            for (int c = 0; c < 3; ++c) { double t = a[k * 3 + c]; a[k * 3 + c] = a[pr * 3 + c]; a[pr * 3 + c] = t; }
        if (pc != k)
            for (int r = 0; r < 3; ++r) { double t = a[r * 3 + k]; a[r * 3 + k] = a[r * 3 + pc]; a[r * 3 + pc] = t; }
        /* Householder on column k, rows k..2 */
        double tail2 = 0.0;
        for (int r = k + 1; r < 3; ++r) tail2 = ORC_FMA(a[r * 3 + k], a[r * 3 + k], tail2);
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
            for (int r = k; r < 3; ++r) dot = ORC_FMA(vv[r], a[r * 3 + c], dot);
            for (int r = k; r < 3; ++r) a[r * 3 + c] -= tau * vv[r] * dot;
        }
    }
    int rank = 0;
    const double thr = maxpivot * (DBL_EPSILON * 3.0);
    for (int i = 0; i < nonzero; ++i)
        if (fabs(diag[i]) > thr) ++rank;
    return rank;
}

/* MLPnPsolver.cpp:705-706: Eigen::LDLT<MatrixXd>(A).solve(g) -- LDL^T with
 * diagonal pivoting (largest |a_ii| first); D entries that are not larger than
 * the smallest normal number are treated as zero in the solve. */
void orc_ldlt6_solve_d(const double a_in[36], const double g[6], double x[6])
{
    enum { N = 6 };
    double a[36];
    int perm[N];
    memcpy(a, a_in, sizeof(a));
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double big = fabs(a[k * N + k]);
        for (int i = k + 1; i < N; ++i)
            if (fabs(a[i * N + i]) > big) { big = fabs(a[i * N + i]); piv = i; }
        if (piv != k) {   /* symmetric row+column swap */
            for (int c = 0; c < N; ++c) { double t = a[k * N + c]; a[k * N + c] = a[piv * N + c]; a[piv * N + c] = t; }
            for (int r = 0; r < N; ++r) { double t = a[r * N + k]; a[r * N + k] = a[r * N + piv]; a[r * N + piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double d = a[k * N + k];
        if (!(fabs(d) > DBL_MIN)) continue;   /* zero pivot: leave column, handled in solve */
        for (int i = k + 1; i < N; ++i) a[i * N + k] = a[i * N + k] / d;          /* L column */
        for (int i = k + 1; i < N; ++i)
            for (int j = k + 1; j <= i; ++j) {
                a[i * N + j] -= a[i * N + k] * d * a[j * N + k];
                a[j * N + i] = a[i * N + j];
            }
    }
    double y[N];
    for (int i = 0; i < N; ++i) y[i] = g[perm[i]];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j) y[i] = ORC_FMA(-a[i * N + j], y[j], y[i]);
    for (int i = 0; i < N; ++i) {
        const double d = a[i * N + i];
        y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0;
    }
    for (int i = N - 1; i >= 0; --i)
        for (int j = i + 1; j < N; ++j) y[i] = ORC_FMA(-a[j * N + i], y[j], y[i]);
    for (int i = 0; i < N; ++i) x[perm[i]] = y[i];
}

/* Null space of an 8 x 12 matrix (the M of a 4-point EPnP system, PnPsolver.cpp:365-379) by Householder QR
 * of its transpose: M^T = Q [R; 0], null(M) = span of the last four columns of Q = H0 H1 ... H7 e_{8..11}.
 * For four correspondences M^T M has an exact four-dimensional null space, so "the four eigenvectors of
 * the smallest eigenvalues" (PnPsolver.cpp:380-382) are ANY orthonormal basis of that space -- which one
 * the reference's eigen-solver returns is decided by rounding noise.  This routine returns a
 * deterministic one for 2.4 kFLOP instead of a 37 kFLOP 12 x 12 eigen-solve.  Operation order is part of
 * the arithmetic contract (mirrored by csrc/linalg.cuh nullspace_qr_8x12).
 * Mrows: 8 x 12 row-major.  U4: 12 x 4 row-major, column i = basis vector i. */
void orc_nullspace_qr_d(const double *Mrows, double *U4)
{
    double A[12][8], tau[8];
    for (int r = 0; r < 12; ++r)
        for (int c = 0; c < 8; ++c) A[r][c] = Mrows[c * 12 + r];
    for (int k = 0; k < 8; ++k) {
        double s = 0.0;
        for (int r = k; r < 12; ++r) s = ORC_FMA(A[r][k], A[r][k], s);
        const double norm = sqrt(s);
        if (norm == 0.0) { tau[k] = 0.0; continue; }
        const double alpha = (A[k][k] > 0.0) ? -norm : norm;
        A[k][k] = A[k][k] - alpha;
        double vtv = 0.0;
        for (int r = k; r < 12; ++r) vtv = ORC_FMA(A[r][k], A[r][k], vtv);
        tau[k] = 2.0 / vtv;
        for (int j = k + 1; j < 8; ++j) {
            double d = 0.0;
            for (int r = k; r < 12; ++r) d = ORC_FMA(A[r][k], A[r][j], d);
            d = d * tau[k];
            for (int r = k; r < 12; ++r) A[r][j] = ORC_FMA(-d, A[r][k], A[r][j]);
        }
        FL(2 * (12 - k) + 1 + 1 + 2 * (12 - k) + 1 + (7 - k) * (4 * (12 - k) + 1));
    }
    for (int i = 0; i < 4; ++i) {
        double y[12];
        for (int r = 0; r < 12; ++r) y[r] = (r == 8 + i) ? 1.0 : 0.0;
        for (int k = 7; k >= 0; --k) {
            if (tau[k] == 0.0) continue;
            double d = 0.0;
            for (int r = k; r < 12; ++r) d = ORC_FMA(A[r][k], y[r], d);
            d = d * tau[k];
            for (int r = k; r < 12; ++r) y[r] = ORC_FMA(-d, A[r][k], y[r]);
            FL(4 * (12 - k) + 1);
        }
        for (int r = 0; r < 12; ++r) U4[r * 4 + i] = y[r];
    }
}
