// kfdb.cuh -- KeyFrameDatabase::DetectRelocalizationCandidates / DetectLoopCandidates for a batch of queries
// (where the candidate keyframes of every relocalisation / loop closure come from).
//
// Reference: src/KeyFrameDatabase.cpp:174-284 (relocalisation, Tracking.cpp:1199), :51-172 (loop detection,
// LoopClosing.cpp:135), DBoW2 L1Scoring::score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66),
// KeyFrame::GetBestCovisibilityKeyFrames(10) (src/KeyFrame.cpp:161-169).
//
// The reference walks an inverted file (word -> list of keyframes) with one pointer chase per posting and keeps its
// per-query state in the KeyFrame objects.  Here the database is the keyframes' BowVectors as CSR arrays, resident on the
// device, and a query is compared with EVERY keyframe -- K x nnz binary searches are cheaper on this machine than
// maintaining posting lists, and every step is independent:
//   1. common words        one warp per (query, keyframe): lanes stride over the keyframe's words and binary-search the
//                          query's (sorted, in shared memory); count, and the SMALLEST shared word w* (step 6)
//   2. threshold           per query: maxCommonWords -> minCommonWords = int(max * 0.8f)
//   3. L1 score            one warp per (query, keyframe above the threshold); the reference's double sum runs over the
//                          shared words in ascending order -- reproduced term by term in that order (ballot + shuffle),
//                          so the float score is bit-identical
//   4. score state         DetectRelocalizationCandidates reads mRelocScore of covisible keyframes that merely share a
//                          word, i.e. possibly the score of an EARLIER query (quirk Q11, oracle/orc_kfdb.c): one thread
//                          per keyframe carries the state through the batch's queries in order
//   5. covisibility sums   one thread per (query, scored keyframe): float adds over its ten neighbours in their order
//   6. emit                the reference returns candidates in the order of lKFsSharingWords = first encounter while
//                          walking the query's words in ascending order and each word's list in insertion order, i.e.
//                          sorted by (w*, keyframe index): one CTA per query sorts the scored keyframes by that key,
//                          applies the 0.75 * best rule and the "each keyframe once" rule (first occurrence wins)
// Integer / index output: bit-identical to the oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

namespace rsac {

struct KfdbArgs {
    // database
    int32_t K;
    const int64_t* kf_off;       // [K+1]
    const uint32_t* kf_word;
    const double* kf_val;
    const int32_t* covis;        // [K][10]
    float* state;                // [K] mRelocScore carried between queries (mode 0)
    // queries
    int32_t Q, mode;
    const int64_t* q_off;        // [Q+1]
    const uint32_t* q_word;
    const double* q_val;
    const float* min_score;      // [Q] (mode 1)
    const int64_t* conn_off;     // [Q+1] (mode 1), lists sorted ascending
    const int32_t* conn;
    // per (query, keyframe) scratch, [Q][K]
    int32_t* cw;                 // shared words (0: none, or connected in mode 1)
    uint32_t* wstar;             // smallest shared word
    float* si;                   // L1 score (valid where cw > minCommon)
    float* eff;                  // mRelocScore as the accumulation loop of query q reads it
    float* acc;                  // accumulated score
    int32_t* best;               // pBestKF
    int32_t* firstpos;           // first qualifying list position whose pBestKF is this keyframe (INT_MAX-filled)
    unsigned long long* keys;    // [Q][K2] sort keys (w* << 32 | keyframe)
    int32_t K2;                  // K rounded up to a power of two
    int32_t* out;                // [Q][K] candidates
    // membership bitmaps of the queries' words over [0, bm_words * 32) (nullptr: binary search in the sorted word list)
    uint32_t* q_bitmap;          // [Q][bm_words]
    int32_t bm_words;
    // per query
    int32_t* min_common;         // [Q]
    float* best_acc;             // [Q] (initialised by the threshold kernel)
    int32_t* n_out;              // [Q]
};

constexpr int kKfdbWarps = 8;
constexpr int kKfdbQMax = 4096;     // query words staged in shared memory (an ORB frame has at most ~2000 features)

// position of w in the ascending array a[0..n), or -1
__device__ __forceinline__ int kfdb_find(const uint32_t* a, int n, uint32_t w)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < w) lo = mid + 1; else hi = mid;
    }
    return (lo < n && a[lo] == w) ? lo : -1;
}

__device__ __forceinline__ bool kfdb_in_sorted(const int32_t* a, int n, int32_t v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < n && a[lo] == v;
}

// 0. membership bitmap of every query's words (the vocabulary has ~10^6 words: 128 KB per query, L2-resident): step 1 then
// costs one load and a bit test per keyframe word instead of a ten-step binary search.  grid (ceil(max nq / 256), Q); the
// bitmaps are zeroed by the host
static __global__ void __launch_bounds__(256) kfdb_bitmap_kernel(KfdbArgs a)
{
    const int q = blockIdx.y;
    const int64_t q0 = a.q_off[q];
    const int nq = (int)(a.q_off[q + 1] - q0);
    uint32_t* bm = a.q_bitmap + (size_t)q * a.bm_words;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
        const uint32_t w = a.q_word[q0 + i];
        if ((w >> 5) < (uint32_t)a.bm_words) atomicOr(bm + (w >> 5), 1u << (w & 31));     // words beyond the database's largest cannot match
    }
}

// 1. shared words of every (query, keyframe) pair.  grid (ceil(K / kKfdbWarps), Q), kKfdbWarps warps
static __global__ void __launch_bounds__(kKfdbWarps * 32) kfdb_common_kernel(KfdbArgs a)
{
    __shared__ uint32_t s_q[kKfdbQMax];
    const int q = blockIdx.y;
    const int64_t q0 = a.q_off[q];
    const int nq = (int)(a.q_off[q + 1] - q0);
    const uint32_t* bm = a.q_bitmap ? a.q_bitmap + (size_t)q * a.bm_words : nullptr;
    const bool staged = !bm && nq <= kKfdbQMax;
    if (staged)
        for (int i = threadIdx.x; i < nq; i += blockDim.x) s_q[i] = a.q_word[q0 + i];
    __syncthreads();
    const uint32_t* qw = staged ? s_q : a.q_word + q0;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * kKfdbWarps + (threadIdx.x >> 5);
    if (k >= a.K) return;
    int cnt = 0;
    uint32_t first = 0xffffffffu;
    bool excluded = false;
    if (a.mode == 1) {
        const int64_t c0 = a.conn_off[q];
        excluded = kfdb_in_sorted(a.conn + c0, (int)(a.conn_off[q + 1] - c0), k);      // spConnectedKeyFrames.count(pKFi)
    }
    if (!excluded) {
        const int64_t k0 = a.kf_off[k];
        const int nk = (int)(a.kf_off[k + 1] - k0);
        for (int base = 0; base < nk; base += 32) {
            const int i = base + lane;
            bool found = false;
            uint32_t w = 0xffffffffu;
            if (i < nk) {
                w = a.kf_word[k0 + i];
                found = bm ? ((bm[w >> 5] >> (w & 31)) & 1u) != 0 : kfdb_find(qw, nq, w) >= 0;      // w <= the database's largest word: inside the bitmap
            }
            cnt += __popc(__ballot_sync(0xffffffffu, found));
            first = min(first, __reduce_min_sync(0xffffffffu, found ? w : 0xffffffffu));
        }
    }
    if (lane == 0) {
        a.cw[(size_t)q * a.K + k] = cnt;
        a.wstar[(size_t)q * a.K + k] = first;
    }
}

// 2. per query: maxCommonWords -> minCommonWords (KeyFrameDatabase.cpp:204-211); initial bestAccScore.  grid Q
static __global__ void __launch_bounds__(256) kfdb_threshold_kernel(KfdbArgs a)
{
    __shared__ int s_max;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    int m = 0;
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) m = max(m, a.cw[(size_t)q * a.K + k]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, m);
    __syncthreads();
    if (threadIdx.x == 0) {
        a.min_common[q] = (int)((float)s_max * 0.8f);
        a.best_acc[q] = a.mode == 0 ? 0.0f : a.min_score[q];
        a.n_out[q] = 0;
    }
}

// 3. L1 score of the keyframes above the threshold (:217-224 / :104-116).  grid as kernel 1
static __global__ void __launch_bounds__(kKfdbWarps * 32) kfdb_score_kernel(KfdbArgs a)
{
    __shared__ uint32_t s_q[kKfdbQMax];
    const int q = blockIdx.y;
    const int64_t q0 = a.q_off[q];
    const int nq = (int)(a.q_off[q + 1] - q0);
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * kKfdbWarps + (threadIdx.x >> 5);
    const int minc = a.min_common[q];
    // does any warp of this CTA have work?  (most keyframes are below the threshold: skip the staging)
    const bool mine = k < a.K && a.cw[(size_t)q * a.K + k] > minc && a.cw[(size_t)q * a.K + k] > 0;
    if (!__syncthreads_or(mine)) return;
    const bool staged = nq <= kKfdbQMax;
    if (staged)
        for (int i = threadIdx.x; i < nq; i += blockDim.x) s_q[i] = a.q_word[q0 + i];
    __syncthreads();
    if (!mine) return;
    const uint32_t* qw = staged ? s_q : a.q_word + q0;
    const int64_t k0 = a.kf_off[k];
    const int nk = (int)(a.kf_off[k + 1] - k0);
    double score = 0;
    for (int base = 0; base < nk; base += 32) {
        const int i = base + lane;
        double term = 0.0;
        bool found = false;
        if (i < nk) {
            const int j = kfdb_find(qw, nq, a.kf_word[k0 + i]);
            if (j >= 0) {
                found = true;
                const double vi = a.q_val[q0 + j], wi = a.kf_val[k0 + i];      // score(F->mBowVec, pKFi->mBowVec)
                term = fabs(vi - wi) - fabs(vi) - fabs(wi);
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, found);
        while (m) {                                                            // ascending word order, one add at a time
            const int src = __ffs(m) - 1;
            m &= m - 1;
            score += __shfl_sync(0xffffffffu, term, src);
        }
    }
    score = -score / 2.0;
    if (lane == 0) a.si[(size_t)q * a.K + k] = (float)score;
}

// 4. mRelocScore as each query's accumulation loop sees it: a keyframe scored by query q holds that score, the others keep
// what an earlier query (or the state carried in) left.  One thread per keyframe, queries in order.
static __global__ void __launch_bounds__(256) kfdb_state_kernel(KfdbArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    float st = a.state[k];
    for (int q = 0; q < a.Q; ++q) {
        const size_t i = (size_t)q * a.K + k;
        const int c = a.cw[i];
        if (c > 0 && c > a.min_common[q]) st = a.si[i];
        a.eff[i] = st;
    }
    a.state[k] = st;
}

__device__ __forceinline__ void kfdb_atomic_max_float(float* addr, float v)
{
    int* ia = reinterpret_cast<int*>(addr);
    int old = *ia;
    while (__int_as_float(old) < v) {
        const int seen = atomicCAS(ia, old, __float_as_int(v));
        if (seen == old) break;
        old = seen;
    }
}

// is keyframe k of query q in lScoreAndMatch?
__device__ __forceinline__ bool kfdb_listed(const KfdbArgs& a, int q, int k)
{
    const size_t i = (size_t)q * a.K + k;
    const int c = a.cw[i];
    if (!(c > 0 && c > a.min_common[q])) return false;
    return a.mode == 0 || a.si[i] >= a.min_score[q];
}

// 5. accumulate by covisibility (:233-259 / :125-148).  grid (ceil(K / 256), Q)
static __global__ void __launch_bounds__(256) kfdb_acc_kernel(KfdbArgs a)
{
    const int q = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K || !kfdb_listed(a, q, k)) return;
    const size_t row = (size_t)q * a.K;
    const int minc = a.min_common[q];
    const float s0 = a.si[row + k];
    float bestScore = s0, accScore = s0;
    int best = k;
    for (int j = 0; j < 10; ++j) {
        const int k2 = a.covis[(size_t)k * 10 + j];
        if (k2 < 0) break;
        if (k2 >= a.K) continue;
        const int c2 = a.cw[row + k2];
        float s2;
        if (a.mode == 0) {
            if (c2 <= 0) continue;                     // mnRelocQuery != F->mnId
            s2 = a.eff[row + k2];
        } else {
            if (!(c2 > 0 && c2 > minc)) continue;
            s2 = a.si[row + k2];
        }
        accScore += s2;
        if (s2 > bestScore) { best = k2; bestScore = s2; }
    }
    a.acc[row + k] = accScore;
    a.best[row + k] = best;
    kfdb_atomic_max_float(a.best_acc + q, accScore);
}

// 6. candidates in the reference's order.  One CTA per query; shared memory holds the sort keys when they fit.
constexpr int kKfdbEmitThreads = 256;
constexpr int kKfdbSortSmem = 4096;
static __global__ void __launch_bounds__(kKfdbEmitThreads) kfdb_emit_kernel(KfdbArgs a)
{
    __shared__ unsigned long long s_keys[kKfdbSortSmem];
    __shared__ int s_n, s_scan[kKfdbEmitThreads], s_base;
    const int q = blockIdx.x, tid = threadIdx.x;
    const size_t row = (size_t)q * a.K;
    if (tid == 0) { s_n = 0; s_base = 0; }
    __syncthreads();
    unsigned long long* gkeys = a.keys + (size_t)q * a.K2;      // K2 = K rounded up to a power of two slots per query
    for (int k = tid; k < a.K; k += blockDim.x)
        if (kfdb_listed(a, q, k)) {
            const int slot = atomicAdd(&s_n, 1);
            RSAC_ASSERT(slot < a.K && a.wstar[row + k] != 0xffffffffu);
            gkeys[slot] = ((unsigned long long)a.wstar[row + k] << 32) | (unsigned)k;
        }
    __syncthreads();
    const int n = s_n;
    if (n == 0) { if (tid == 0) a.n_out[q] = 0; return; }
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    // bitonic sort of the keys (unique: the keyframe index is part of the key), padded with +inf to a power of two
    unsigned long long* keys = np2 <= kKfdbSortSmem ? s_keys : gkeys;
    for (int i = tid; i < np2; i += blockDim.x) {
        if (keys == s_keys) s_keys[i] = i < n ? gkeys[i] : ~0ull;
        else if (i >= n) gkeys[i] = ~0ull;
    }
    __syncthreads();
    for (int size = 2; size <= np2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < np2; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const unsigned long long x = keys[i], y = keys[j];
                    const bool up = (i & size) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[j] = x; }
                }
            }
            __syncthreads();
        }
    const float retain = 0.75f * a.best_acc[q];
    for (int j = tid; j < n; j += blockDim.x) {
        const int k = (int)(keys[j] & 0xffffffffull);
        if (a.acc[row + k] > retain) atomicMin(a.firstpos + row + a.best[row + k], j);
    }
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int j = base + tid;
        int flag = 0, b = -1;
        if (j < n) {
            const int k = (int)(keys[j] & 0xffffffffull);
            b = a.best[row + k];
            flag = (a.acc[row + k] > retain && a.firstpos[row + b] == j) ? 1 : 0;
        }
        s_scan[tid] = flag;
        __syncthreads();
        for (int off = 1; off < blockDim.x; off <<= 1) {                      // inclusive Hillis-Steele scan
            const int v = tid >= off ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        RSAC_ASSERT(!flag || (s_base + s_scan[tid] - 1 >= 0 && s_base + s_scan[tid] - 1 < a.K && b >= 0 && b < a.K));
        if (flag) a.out[row + s_base + s_scan[tid] - 1] = b;
        __syncthreads();
        if (tid == blockDim.x - 1) s_base += s_scan[tid];
        __syncthreads();
    }
    if (tid == 0) a.n_out[q] = s_base;
}

}  // namespace rsac
