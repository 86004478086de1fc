// bow.cuh -- ORBmatcher::SearchByBoW (both overloads) for a batch of (keyframe, frame) / (keyframe, keyframe) pairs
// (the producer of every correspondence set the RANSAC engine verifies).
//
// Reference: src/ORBmatcher.cpp:110-239 (SearchByBoW(KF, Frame): Tracking::Relocalization, Tracking.cpp:1214),
// :354-487 (SearchByBoW(KF1, KF2): LoopClosing::ComputeSim3, LoopClosing.cpp:251), DescriptorDistance :1492-1508,
// ComputeThreeMaxima :1445-1488.
//
// The reference's loop is greedy and sequential -- a target feature that an earlier query feature took is skipped
// (:162, :410) -- but only WITHIN a vocabulary node: a feature sits in exactly one node of the FeatureVector, so the
// nodes of a pair are independent.  One warp per (pair, query node): the query features of the node are taken in
// order; for each, the lanes compute the 256-bit Hamming distances to the node's target features (XOR + POPC, 8
// words), keep (smallest, second smallest, first position of the smallest) per lane and merge them with shuffles --
// the reference's scan returns exactly the two smallest values of the multiset and the FIRST position of the minimum
// -- then the TH_LOW / ratio test, the rotation bin, and the `taken` mark that the next query feature sees.  A second
// kernel (one CTA per pair) builds the 30-bin rotation histogram, finds the three maxima and drops the other bins.
// Integer work throughout: bit-identical to the oracle (oracle/orc_bow.c).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rsac {

constexpr int kBowHisto = 30;      // ORBmatcher::HISTO_LENGTH
constexpr int kBowThLow = 50;      // ORBmatcher::TH_LOW

struct BowSet {                    // one Frame / KeyFrame in the concatenated device arrays
    int32_t feat_off, n_feat;      // features: desc [.][8], angle, valid
    int32_t node_off, n_nodes;     // node_ids [n_nodes]
    int32_t noff_off;              // node_start [n_nodes + 1] (values relative to nf_off)
    int32_t nf_off;                // node_feat
};

struct BowArgs {
    const BowSet* sets;
    const int32_t* qset;           // [C] query set (outer loop: pKF / pKF1) of every pair
    const int32_t* tset;           // [C] target set (F / pKF2)
    const int64_t* t2q_off;        // [C + 1] offsets of the per-pair target-indexed arrays
    const int64_t* q2t_off;        // [C + 1] offsets of the per-pair query-indexed arrays
    const int2* items;             // [n_items] (pair, query node)
    int32_t n_items, C;
    const uint32_t* desc;
    const float* angle;
    const uint8_t* valid;          // nullptr = every feature is usable
    const uint32_t* node_ids;
    const int32_t* node_start;
    const uint32_t* node_feat;
    int32_t* t2q;                  // target feature -> query feature (-1: free); doubles as the `taken` mark
    int32_t* q2t;                  // query feature -> target feature (mode 1 output)
    int8_t* bin_t;                 // rotation bin of the match stored at t2q[.]
    int32_t* n_matches;            // [C]
    float nn_ratio;
    int32_t check_orientation, mode;
};

__device__ __forceinline__ int bow_distance(const uint4 a0, const uint4 a1, const uint32_t* __restrict__ b)
{
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 4);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

static __global__ void __launch_bounds__(128) bow_match_kernel(BowArgs a)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int w = warp; w < a.n_items; w += nwarps) {
        const int2 it = a.items[w];
        const int p = it.x;
        const BowSet Q = a.sets[a.qset[p]], T = a.sets[a.tset[p]];
        const uint32_t node = a.node_ids[Q.node_off + it.y];
        // the target's node with the same id (the reference walks both maps with lower_bound: a sorted intersection)
        int lo = 0, hi = T.n_nodes;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a.node_ids[T.node_off + mid] < node) lo = mid + 1; else hi = mid;
        }
        if (lo >= T.n_nodes || a.node_ids[T.node_off + lo] != node) continue;
        const int qb = a.node_start[Q.noff_off + it.y], qe = a.node_start[Q.noff_off + it.y + 1];
        const int tb = a.node_start[T.noff_off + lo], te = a.node_start[T.noff_off + lo + 1];
        const uint32_t* qf = a.node_feat + Q.nf_off;
        const uint32_t* tf = a.node_feat + T.nf_off;
        int32_t* t2q = a.t2q + a.t2q_off[p];
        int32_t* q2t = a.q2t + a.q2t_off[p];
        int8_t* bin_t = a.bin_t + a.t2q_off[p];
        for (int qi = qb; qi < qe; ++qi) {
            const int idx_q = (int)qf[qi];
            if (a.valid && !a.valid[Q.feat_off + idx_q]) continue;          // !pMP || pMP->isBad()
            const uint32_t* dq = a.desc + 8 * (size_t)(Q.feat_off + idx_q);
            const uint4 q0 = *reinterpret_cast<const uint4*>(dq), q1 = *reinterpret_cast<const uint4*>(dq + 4);
            // per lane: the two smallest distances of its targets and the first position of the smallest
            int b1 = 256, b2 = 256, bpos = 0x7fff;
            for (int j = tb + lane; j < te; j += 32) {
                const int idx_t = (int)tf[j];
                if (t2q[idx_t] >= 0) continue;                               // already taken (:162 / vbMatched2 :410)
                if (a.mode == 1 && a.valid && !a.valid[T.feat_off + idx_t]) continue;
                const int dist = bow_distance(q0, q1, a.desc + 8 * (size_t)(T.feat_off + idx_t));
                if (dist < b1) { b2 = b1; b1 = dist; bpos = j - tb; }
                else if (dist < b2) { b2 = dist; }
            }
            // warp: minimum of (distance, position) -- the reference's strict `<` keeps the first minimum
            unsigned key = ((unsigned)b1 << 16) | (unsigned)bpos;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) key = min(key, __shfl_xor_sync(FULL, key, d));
            const int best1 = (int)(key >> 16), bestpos = (int)(key & 0xffffu);
            // second smallest of the multiset: every lane offers its smallest, the winner its second smallest
            const bool winner = (b1 == best1 && bpos == bestpos && best1 < 256);
            int sec = winner ? b2 : b1;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sec = min(sec, __shfl_xor_sync(FULL, sec, d));
            const bool pass = a.mode == 0 ? (best1 <= kBowThLow) : (best1 < kBowThLow);      // :179 vs :431
            if (pass && (float)best1 < a.nn_ratio * (float)sec) {
                if (lane == 0) {
                    const int idx_t = (int)tf[tb + bestpos];
                    t2q[idx_t] = idx_q;
                    q2t[idx_q] = idx_t;
                    if (a.check_orientation) {
                        float rot = a.angle[Q.feat_off + idx_q] - a.angle[T.feat_off + idx_t];
                        if (rot < 0.0f) rot += 360.0f;
                        int bin = (int)roundf(rot * (1.0f / kBowHisto));
                        if (bin == kBowHisto) bin = 0;
                        bin_t[idx_t] = (int8_t)bin;
                    }
                }
            }
            __syncwarp();                                                     // the mark is visible to the next query feature
        }
    }
}

// rotation consistency (:214-234) and the final count, one CTA per pair
static __global__ void __launch_bounds__(128) bow_orient_kernel(BowArgs a)
{
    __shared__ int histo[kBowHisto];
    __shared__ int s_ind[3], s_count;
    const int p = blockIdx.x;
    const BowSet T = a.sets[a.tset[p]];
    int32_t* t2q = a.t2q + a.t2q_off[p];
    int32_t* q2t = a.q2t + a.q2t_off[p];
    const int8_t* bin_t = a.bin_t + a.t2q_off[p];
    if (threadIdx.x < kBowHisto) histo[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (a.check_orientation) {
        for (int t = threadIdx.x; t < T.n_feat; t += blockDim.x)
            if (t2q[t] >= 0) {
                const int b = bin_t[t];
                if (b >= 0 && b < kBowHisto) atomicAdd(&histo[b], 1);
            }
        __syncthreads();
        if (threadIdx.x == 0) {                           // ComputeThreeMaxima (:1445-1488)
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int i = 0; i < kBowHisto; i++) {
                const int s = histo[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
            s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        }
        __syncthreads();
    }
    int cnt = 0;
    for (int t = threadIdx.x; t < T.n_feat; t += blockDim.x) {
        const int q = t2q[t];
        if (q < 0) continue;
        if (a.check_orientation) {
            const int b = bin_t[t];
            if (b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) {
                t2q[t] = -1;
                q2t[q] = -1;
                continue;
            }
        }
        ++cnt;
    }
    atomicAdd(&s_count, cnt);
    __syncthreads();
    if (threadIdx.x == 0) a.n_matches[p] = s_count;
}

}  // namespace rsac
