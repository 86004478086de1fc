// sim3.cuh -- Sim3Solver on the device: Horn's closed form on three 3D-3D pairs, two-way
// reprojection scoring, and the reference's selection rule, following src/Sim3Solver.cpp.
//
// All arithmetic is FP32 with the reference's operation order (this TU is compiled with
// -fmad=false), so hypotheses, errors and inlier bits are bit-identical to the CPU checker.
// A loop-closure batch is small (a handful of candidates x 300 hypotheses x ~200 matches), so
// the whole of Sim3Solver::iterate for one candidate runs in ONE CTA: thread <-> hypothesis for
// Horn (4x4 Jacobi in registers) and for scoring (correspondences broadcast from shared memory),
// then the first hypothesis with more than minInliers inliers wins, otherwise the last arg-max
// (Sim3Solver.cpp:155,163).
#pragma once
#include "common.cuh"
#include "linalg.cuh"
#include "rng.cuh"            // rng_tables_kernel
#include "select.cuh"         // ResultRec

namespace rsac {

// Sim3Solver::ComputeCentroid (Sim3Solver.cpp:186-194); P[k*3+c] = component c of point k
__host__ __device__ inline void sim3_centroid(const float* P, float* Pr, float* C)
{
    for (int c = 0; c < 3; ++c) {
        C[c] = P[0 * 3 + c] + P[1 * 3 + c] + P[2 * 3 + c];
        C[c] = C[c] / 3.f;
    }
    for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) Pr[k * 3 + c] = P[k * 3 + c] - C[c];
}

// Sim3Solver::ComputeSim3 (Sim3Solver.cpp:196-266); fix_scale = 0 adds Horn's scale step as
// upstream ORB-SLAM2 does (absent from the reference)
__host__ __device__ inline void sim3_compute(const float* P1, const float* P2, int fix_scale, float* R12, float* t12, float* s12)
{
    float Pr1[9], Pr2[9], O1[3], O2[3];
    sim3_centroid(P1, Pr1, O1);
    sim3_centroid(P2, Pr2, O2);
    float M[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            M[r * 3 + c] = Pr2[0 * 3 + r] * Pr1[0 * 3 + c] + Pr2[1 * 3 + r] * Pr1[1 * 3 + c] + Pr2[2 * 3 + r] * Pr1[2 * 3 + c];
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
    jacobi_eig<float, 4>(N, w, V);
    quat_to_rot<float>(V[0 * 4 + 3], V[1 * 4 + 3], V[2 * 4 + 3], V[3 * 4 + 3], R12);
    float s = 1.0f;
    if (!fix_scale) {
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
    for (int r = 0; r < 3; ++r) {
        const float sr0 = s * R12[r * 3 + 0], sr1 = s * R12[r * 3 + 1], sr2 = s * R12[r * 3 + 2];
        const float ro = fix_scale ? (R12[r * 3 + 0] * O2[0] + R12[r * 3 + 1] * O2[1] + R12[r * 3 + 2] * O2[2])
                                   : (sr0 * O2[0] + sr1 * O2[1] + sr2 * O2[2]);
        t12[r] = O1[r] - ro;
    }
}

// T12 = [sR t], T21 = T12^-1 (Sim3Solver.cpp:259-265)
__host__ __device__ inline void sim3_make_T(const float* R12, const float* t12, float s, float* A12, float* A21, float* t21)
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

// Sim3Solver::Project (Sim3Solver.cpp:306-327)
__host__ __device__ inline void sim3_project(const float* A, const float* t, const float* K, float X0, float X1, float X2, float& u, float& v)
{
    const float x = (A[0] * X0 + A[1] * X1 + A[2] * X2) + t[0];
    const float y = (A[3] * X0 + A[4] * X1 + A[5] * X2) + t[1];
    const float z = (A[6] * X0 + A[7] * X1 + A[8] * X2) + t[2];
    const float invz = 1 / z;
    const float xn = x * invz, yn = y * invz;
    u = K[0] * xn + K[2];
    v = K[1] * yn + K[3];
}

// Sim3Solver::FromCameraToImage (Sim3Solver.cpp:329-347)
__host__ __device__ inline void sim3_cam2img(const float* K, float X0, float X1, float X2, float& u, float& v)
{
    const float invz = 1 / X2;
    const float xn = X0 * invz, yn = X1 * invz;
    u = K[0] * xn + K[2];
    v = K[1] * yn + K[3];
}

// packed correspondence records (built once per problem, Sim3Solver.cpp:51-52,57-63,81-82):
//   c1 = (X1c.xyz, thr1), c2 = (X2c.xyz, thr2), c3 = (p1im1.uv, p2im2.uv)
static __global__ void sim3_pack_kernel(const ProblemMeta* metas, const float* x1c, const float* x2c, const float* s1, const float* s2,
                                 float4* c1, float4* c2, float4* c3, int C)
{
  for (int pr = blockIdx.y; pr < C; pr += gridDim.y) {       // grid.y is capped at 65535 problems
    const ProblemMeta& m = metas[pr];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.n; i += gridDim.x * blockDim.x) {
        const size_t g = (size_t)m.corr_off + i;
        const float a0 = x1c[3 * g], a1 = x1c[3 * g + 1], a2 = x1c[3 * g + 2];
        const float b0 = x2c[3 * g], b1 = x2c[3 * g + 1], b2 = x2c[3 * g + 2];
        // thresholds: size_t(9.210*sigma2) converted to float by the comparison (Q4)
        const float thr1 = (float)(unsigned long long)(9.210 * (double)s1[g]);
        const float thr2 = (float)(unsigned long long)(9.210 * (double)s2[g]);
        float u1, v1, u2, v2;
        sim3_cam2img(m.k1, a0, a1, a2, u1, v1);
        sim3_cam2img(m.k2, b0, b1, b2, u2, v2);
        c1[g] = make_float4(a0, a1, a2, thr1);
        c2[g] = make_float4(b0, b1, b2, thr2);
        c3[g] = make_float4(u1, v1, u2, v2);
    }
  }
}

struct Sim3Args {
    const ProblemMeta* metas;
    const uint32_t* tables;
    const float4* c1;
    const float4* c2;
    const float4* c3;
    float* poses;        // [sumH][13]  R 9, t 3, s
    int32_t* counts;     // [sumH]
    uint32_t* hmasks;    // [sum H*words]
    void* results;
    void* results2;
    uint32_t* masks;
    int32_t* done;       // [C] CTAs of a candidate that have finished (zero on entry, zero again on exit)
    int32_t problem_base;
    int32_t tile;        // correspondences per shared-memory tile (a multiple of 32)
    int32_t tiles_h;     // CTAs per candidate in the grid = ceil(maxH / hypotheses per CTA)
};

constexpr int kSim3Threads = 256;
// LANES per hypothesis (each scores every LANES-th mask word): 8 for a loop-closure-sized batch (latency: cfg3, one
// candidate, 0.023 ms), 2 once the batch fills the machine anyway (Horn is computed redundantly on the lanes of a hypothesis)
__host__ __device__ constexpr int sim3_hyps_per_cta(int lanes) { return kSim3Threads / lanes; }

// Sim3Solver::iterate for a batch of loop candidates.  Grid = candidates x hypothesis tiles: a CTA takes 32 hypotheses
// of one candidate, EIGHT LANES per hypothesis -- Horn's closed form redundantly on the eight lanes (2 kFLOP), then the
// two-way reprojection test of every 8th mask word on each lane (the correspondences are staged in shared memory), the
// inlier count reduced over the eight lanes by shuffles.  A loop-closure attempt has a handful of candidates
// (LoopClosing.cpp:238-265), so one CTA per candidate (round 1) left 147 SMs idle and ran 200 evaluations + Horn as ONE
// thread's dependent chain: cfg3 (1 candidate x 300 hypotheses x 200 matches) 0.068 ms.  The CTA that finishes last for
// a candidate (a per-candidate arrival counter) replays the selection rule (Sim3Solver.cpp:155,163) -- one launch.
template <int kSim3Lanes>
__global__ void __launch_bounds__(kSim3Threads) sim3_kernel(Sim3Args a)
{
    constexpr int kSim3HypsPerCta = sim3_hyps_per_cta(kSim3Lanes);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* s1 = reinterpret_cast<float4*>(smem_raw);
    float4* s2 = s1 + a.tile;
    float4* s3 = s2 + a.tile;
    __shared__ int s_first, s_bestcnt, s_besth, s_last;

    const int prob = blockIdx.x / a.tiles_h, hb = blockIdx.x - prob * a.tiles_h;
    const ProblemMeta* m = a.metas + prob;
    const int N = m->n, H = m->H, words = m->words, minInl = m->min_inl;
    const int tid = threadIdx.x;
    const int my_tiles = (N < minInl || H == 0) ? 1 : (H + kSim3HypsPerCta - 1) / kSim3HypsPerCta;
    if (hb >= my_tiles) return;                          // ragged batches: this candidate has fewer hypothesis tiles
    ResultRec res;
    res.ok = 0; res.no_more = 0; res.n_inliers = 0; res.best_hyp = -1; res.refined = 0; res.n_refines = 0;
    res.best_count = 0; res.n_hyp = 0;
    for (int i = 0; i < 9; ++i) res.R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    res.t[0] = res.t[1] = res.t[2] = 0.0f; res.s = 1.0f;
    res.problem = a.problem_base + prob; res.reserved[0] = res.reserved[1] = 0;
    uint32_t* final_mask = a.masks + m->word_off;

    if (N < minInl || H == 0) {                 // Sim3Solver.cpp:119-123
        res.no_more = 1;
        for (int w = tid; w < words; w += blockDim.x) final_mask[w] = 0u;
        if (tid == 0) {
            reinterpret_cast<ResultRec*>(a.results)[prob] = res;
            if (a.results2) reinterpret_cast<ResultRec*>(a.results2)[prob] = res;
        }
        return;
    }

    {
        const int sub = tid & (kSim3Lanes - 1);
        const int h = hb * kSim3HypsPerCta + tid / kSim3Lanes;
        const bool live = h < H;
        float R[9], t[3], s = 1.0f, A12[9], A21[9], t21[3];
        if (live) {
            const uint32_t* idx = a.tables + m->table_off + (size_t)h * 3;
            float P1[9], P2[9];
            for (int i = 0; i < 3; ++i) {           // Sim3Solver.cpp:139-149
                const float4 p = a.c1[(size_t)m->corr_off + idx[i]];
                const float4 q = a.c2[(size_t)m->corr_off + idx[i]];
                P1[3 * i] = p.x; P1[3 * i + 1] = p.y; P1[3 * i + 2] = p.z;
                P2[3 * i] = q.x; P2[3 * i + 1] = q.y; P2[3 * i + 2] = q.z;
            }
            sim3_compute(P1, P2, m->fix_scale, R, t, &s);
            sim3_make_T(R, t, s, A12, A21, t21);
            if (sub == 0) {
                float* out = a.poses + (size_t)(m->hyp_off + h) * 13;
                for (int i = 0; i < 9; ++i) out[i] = R[i];
                out[9] = t[0]; out[10] = t[1]; out[11] = t[2]; out[12] = s;
            }
        }
        // CheckInliers (Sim3Solver.cpp:269-293): correspondences staged tile by tile, mask words dealt to the eight lanes
        int cnt = 0;
        uint32_t* hm = a.hmasks + m->hmask_off + (int64_t)h * words;
        for (int c0 = 0; c0 < N; c0 += a.tile) {
            const int nc = min(a.tile, N - c0);
            __syncthreads();
            for (int i = tid; i < nc; i += blockDim.x) {
                const size_t g = (size_t)m->corr_off + c0 + i;
                s1[i] = a.c1[g]; s2[i] = a.c2[g]; s3[i] = a.c3[g];
            }
            __syncthreads();
            if (live) {
                for (int w0 = sub * 32; w0 < nc; w0 += 32 * kSim3Lanes) {
                    uint32_t bits = 0u;
                    const int ne = min(32, nc - w0);
                    for (int i = 0; i < ne; ++i) {
                        const float4 p = s1[w0 + i], q = s2[w0 + i], im = s3[w0 + i];
                        float u21, v21, u12, v12;
                        sim3_project(A12, t, m->k1, q.x, q.y, q.z, u21, v21);     // set 2 into image 1
                        sim3_project(A21, t21, m->k2, p.x, p.y, p.z, u12, v12);   // set 1 into image 2
                        const float d1x = im.x - u21, d1y = im.y - v21;
                        const float d2x = u12 - im.z, d2y = v12 - im.w;
                        const float err1 = d1x * d1x + d1y * d1y;
                        const float err2 = d2x * d2x + d2y * d2y;
                        if (err1 < p.w && err2 < q.w) bits |= 1u << i;
                    }
                    cnt += __popc(bits);
                    hm[(c0 + w0) >> 5] = bits;
                }
            }
        }
#pragma unroll
        for (int d = kSim3Lanes / 2; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (live && sub == 0) a.counts[m->hyp_off + h] = cnt;
    }
    // arrival: the last CTA of this candidate replays the selection
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int prev = atomicAdd(a.done + prob, 1);
        s_last = (prev == my_tiles - 1);
        s_first = H; s_bestcnt = -1; s_besth = -1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile int32_t* counts = a.counts + m->hyp_off;
    for (int h = tid; h < H; h += blockDim.x)
        if (counts[h] > minInl) atomicMin(&s_first, h);    // first hypothesis that returns true (:163)
    __syncthreads();
    // outcome: first h with cnt > minInl; otherwise best = LAST arg-max over all H (>=, :155)
    const int first = s_first;
    const int limit = (first < H) ? first + 1 : H;         // hypotheses the reference evaluates
    for (int h = tid; h < limit; h += blockDim.x) atomicMax(&s_bestcnt, counts[h]);
    __syncthreads();
    const int bestcnt = s_bestcnt;
    for (int h = tid; h < limit; h += blockDim.x)
        if (counts[h] == bestcnt) atomicMax(&s_besth, h);
    __syncthreads();
    const int besth = s_besth;
    const volatile float* bp = a.poses + (size_t)(m->hyp_off + besth) * 13;
    res.best_hyp = besth;
    res.best_count = bestcnt;
    res.n_hyp = limit;
    for (int i = 0; i < 9; ++i) res.R[i] = bp[i];      // mBestRotation / mBestTranslation (:160-161)
    res.t[0] = bp[9]; res.t[1] = bp[10]; res.t[2] = bp[11]; res.s = bp[12];
    if (first < H) {
        res.ok = 1;
        res.n_inliers = counts[first];
        const volatile uint32_t* hm = a.hmasks + m->hmask_off + (int64_t)first * words;
        for (int w = tid; w < words; w += blockDim.x) final_mask[w] = hm[w];
    } else {
        res.no_more = 1;                                // :174-175
        for (int w = tid; w < words; w += blockDim.x) final_mask[w] = 0u;
    }
    if (tid == 0) {
        reinterpret_cast<ResultRec*>(a.results)[prob] = res;
        if (a.results2) reinterpret_cast<ResultRec*>(a.results2)[prob] = res;
        a.done[prob] = 0;                               // ready for the next run
    }
}

}  // namespace rsac
