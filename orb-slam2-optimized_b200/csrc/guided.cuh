// guided.cuh -- ORBmatcher::SearchBySim3 for a batch of keyframe pairs (the guided matching between
// Sim3Solver and Optimizer::OptimizeSim3 in LoopClosing::ComputeSim3, LoopClosing.cpp:286-311).
//
// Reference: src/ORBmatcher.cpp:948-1171 with KeyFrame::GetFeaturesInArea (src/KeyFrame.cpp:560-599), IsInImage
// (:601-604), MapPoint::PredictScale (src/MapPoint.cpp:367-382), Get{Min,Max}DistanceInvariance (:355-365),
// DescriptorDistance (ORBmatcher.cpp:1492-1508), TH_HIGH = 100.
//
// Unlike SearchByBoW there is no greedy dependence between features: every map point of KF1 is projected into KF2 with
// the Sim3 and takes the most similar keypoint of the right octave inside a window, independently of the others (the same
// from KF2 into KF1), and a match survives when the two directions agree.  So: one THREAD per (pair, direction, feature)
// for the search -- projection, window walk over the keyframe's grid cells in the reference's (ix, iy, insertion) order so
// that ties in the Hamming distance resolve to the same keypoint, 8 x (XOR + POPC) per candidate -- and one thread per KF1
// feature for the agreement.  Float arithmetic in the reference's order (this translation unit is built with -fmad=false);
// index output bit-identical to the oracle (oracle/orc_guided.c).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

namespace rsac {

constexpr int kGuidedThHigh = 100;     // ORBmatcher::TH_HIGH
constexpr int kGuidedMaxLevels = 16;

struct KfViewDev {
    int32_t feat_off, n_feat;          // features in the concatenated per-feature arrays
    int32_t goff_off, gidx_off;        // grid: cell offsets (cols*rows + 1 entries) and feature indices
    int32_t grid_cols, grid_rows, n_levels, pad;
    float grid_w_inv, grid_h_inv, log_scale_factor, pad2;
    float Rcw[9], tcw[3], bounds[4];   // bounds = mnMinX, mnMaxX, mnMinY, mnMaxY
    float scale_factors[kGuidedMaxLevels];
};

struct Sim3SearchArgs {
    const KfViewDev* views;
    const float* kp_xy;                // [.][2]
    const int32_t* kp_octave;
    const float* kp_angle;
    const uint32_t* desc;              // [.][8]
    const uint8_t* mp_valid;
    const float* mp_xyz;               // [.][3]
    const uint32_t* mp_desc;           // [.][8]
    const float* mp_maxdist;
    const float* mp_mindist;
    const int32_t* grid_off;
    const int32_t* grid_idx;
    int32_t C;
    const int32_t* kf1;                // [C]
    const int32_t* kf2;
    const float* K;                    // [C][4] pKF1's fx, fy, cx, cy (used for both directions, ORBmatcher.cpp:951-954)
    const float* R12;                  // [C][9]
    const float* t12;                  // [C][3]
    const float* s12;                  // [C] or nullptr (= 1)
    float th;
    const int64_t* off1;               // [C+1] offsets of the per-pair KF1-indexed arrays
    const int64_t* off2;               // [C+1] KF2-indexed
    const int32_t* matched_in;         // KF1-indexed or nullptr
    uint8_t* already1;                 // vbAlreadyMatched1 / 2
    uint8_t* already2;
    int32_t* m1;                       // vnMatch1 / vnMatch2
    int32_t* m2;
    int32_t* match12;                  // output, KF1-indexed
    int32_t* n_found;                  // [C]
};

__device__ __forceinline__ int guided_descriptor_distance(const uint32_t* a, const uint32_t* b)
{
    const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 4);
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 4);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// MapPoint::PredictScale (MapPoint.cpp:367-382): ::log(double) on the float ratio (oracle/orc_guided.c, Q12)
__device__ __forceinline__ int guided_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels)
{
    const float ratio = max_distance / current_dist;
    int nScale = (int)ceil(log((double)ratio) / (double)log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= n_levels) nScale = n_levels - 1;
    return nScale;
}

__device__ __forceinline__ void guided_mat3_vec(const float* R, const float* p, const float* t, float* o)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = ((R[3 * i] * p[0] + R[3 * i + 1] * p[1]) + R[3 * i + 2] * p[2]) + t[i];
}

// vbAlreadyMatched1 / 2 from vpMatches12 on entry (:975-987).  grid (ceil(maxN1 / 256), C); already2 zeroed by the host
static __global__ void __launch_bounds__(256) sim3_search_prepare_kernel(Sim3SearchArgs a)
{
    const int c = blockIdx.y;
    const int n1 = a.views[a.kf1[c]].n_feat, n2 = a.views[a.kf2[c]].n_feat;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
        const int idx2 = a.matched_in ? a.matched_in[a.off1[c] + i] : -1;
        a.already1[a.off1[c] + i] = idx2 != -1;
        if (idx2 >= 0 && idx2 < n2) a.already2[a.off2[c] + idx2] = 1;
    }
}

// one direction per blockIdx.y parity: the map points of `src` searched in `dst` (:994-1067 / :1070-1150)
static __global__ void __launch_bounds__(128) sim3_search_kernel(Sim3SearchArgs a)
{
    const int c = blockIdx.y >> 1, dir = blockIdx.y & 1;
    const KfViewDev& src = a.views[dir == 0 ? a.kf1[c] : a.kf2[c]];
    const KfViewDev& dst = a.views[dir == 0 ? a.kf2[c] : a.kf1[c]];
    const int64_t obase = dir == 0 ? a.off1[c] : a.off2[c];
    const uint8_t* already = (dir == 0 ? a.already1 : a.already2) + obase;
    int32_t* match = (dir == 0 ? a.m1 : a.m2) + obase;
    // transformation applied to a point in src's camera frame: dir 0: sR21, t21; dir 1: sR12, t12 (:966-968, upstream scale)
    const float* R12 = a.R12 + 9 * (size_t)c;
    const float* t12 = a.t12 + 3 * (size_t)c;
    const float s12 = a.s12 ? a.s12[c] : 1.0f;
    float Rds[9], tds[3];
    if (dir == 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Rds[k] = s12 * R12[k];
        tds[0] = t12[0]; tds[1] = t12[1]; tds[2] = t12[2];
    } else {
        const float inv_s = (float)(1.0 / (double)s12);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) Rds[3 * r + q] = inv_s * R12[3 * q + r];
#pragma unroll
        for (int i = 0; i < 3; ++i) tds[i] = -((Rds[3 * i] * t12[0] + Rds[3 * i + 1] * t12[1]) + Rds[3 * i + 2] * t12[2]);
    }
    const float fx = a.K[4 * c], fy = a.K[4 * c + 1], cx = a.K[4 * c + 2], cy = a.K[4 * c + 3];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < src.n_feat; i += gridDim.x * blockDim.x) {
        const size_t g = (size_t)src.feat_off + i;
        int bestIdx = -1;
        if (a.mp_valid[g] && !already[i]) {
            float pw[3] = {a.mp_xyz[3 * g], a.mp_xyz[3 * g + 1], a.mp_xyz[3 * g + 2]}, pcs[3], pc[3];
            guided_mat3_vec(src.Rcw, pw, src.tcw, pcs);
            guided_mat3_vec(Rds, pcs, tds, pc);
            if (!(pc[2] < 0.0f)) {
                const float invz = (float)(1.0 / (double)pc[2]);
                const float x = pc[0] * invz, y = pc[1] * invz;
                const float u = fx * x + cx, v = fy * y + cy;
                const float maxD = 1.2f * a.mp_maxdist[g], minD = 0.8f * a.mp_mindist[g];
                const float dist3D = sqrtf((pc[0] * pc[0] + pc[1] * pc[1]) + pc[2] * pc[2]);
                if (u >= dst.bounds[0] && u < dst.bounds[1] && v >= dst.bounds[2] && v < dst.bounds[3] && !(dist3D < minD || dist3D > maxD)) {
                    const int lvl = guided_predict_scale(a.mp_maxdist[g], dist3D, dst.log_scale_factor, dst.n_levels);
                    const float r = a.th * dst.scale_factors[lvl];
                    // KeyFrame::GetFeaturesInArea (KeyFrame.cpp:560-599), candidates visited in its order
                    const float mnMinX = dst.bounds[0], mnMinY = dst.bounds[2];
                    const int nMinCellX = max(0, (int)floorf((u - mnMinX - r) * dst.grid_w_inv));
                    const int nMaxCellX = min(dst.grid_cols - 1, (int)ceilf((u - mnMinX + r) * dst.grid_w_inv));
                    const int nMinCellY = max(0, (int)floorf((v - mnMinY - r) * dst.grid_h_inv));
                    const int nMaxCellY = min(dst.grid_rows - 1, (int)ceilf((v - mnMinY + r) * dst.grid_h_inv));
                    if (nMinCellX < dst.grid_cols && nMaxCellX >= 0 && nMinCellY < dst.grid_rows && nMaxCellY >= 0) {
                        const uint32_t* dMP = a.mp_desc + 8 * g;
                        int bestDist = INT_MAX;
                        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
                            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
                                const int cell = dst.goff_off + ix * dst.grid_rows + iy;
                                for (int j = a.grid_off[cell]; j < a.grid_off[cell + 1]; ++j) {
                                    const int idx = a.grid_idx[dst.gidx_off + j];
                                    const size_t gd = (size_t)dst.feat_off + idx;
                                    const float distx = a.kp_xy[2 * gd] - u, disty = a.kp_xy[2 * gd + 1] - v;
                                    if (!(fabsf(distx) < r && fabsf(disty) < r)) continue;
                                    const int oct = a.kp_octave[gd];
                                    if (oct < lvl - 1 || oct > lvl) continue;
                                    const int dist = guided_descriptor_distance(dMP, a.desc + 8 * gd);
                                    if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
                                }
                            }
                        if (bestDist > kGuidedThHigh) bestIdx = -1;
                    }
                }
            }
        }
        match[i] = bestIdx;
    }
}

// agreement of the two directions (:1153-1168).  grid (ceil(maxN1 / 256), C); n_found zeroed by the host
static __global__ void __launch_bounds__(256) sim3_search_agree_kernel(Sim3SearchArgs a)
{
    const int c = blockIdx.y;
    const int n1 = a.views[a.kf1[c]].n_feat;
    int found = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
        const int idx2 = a.m1[a.off1[c] + i];
        int out = -1;
        if (idx2 >= 0 && a.m2[a.off2[c] + idx2] == i) { out = idx2; ++found; }
        a.match12[a.off1[c] + i] = out;
    }
    found = __reduce_add_sync(0xffffffffu, found);
    if ((threadIdx.x & 31) == 0 && found) atomicAdd(a.n_found + c, found);
}

// ------------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(Frame&, KeyFrame, sAlreadyFound, th, ORBdist)   (src/ORBmatcher.cpp:1317-1444;
// Tracking::Relocalization, Tracking.cpp:1296,1310).  GREEDY in the reference: map points are taken in keyframe-feature
// order and a frame keypoint that an earlier map point took is skipped by the later ones (:1389, :1403).  Reproduced
// exactly, in parallel:
//   1. every map point (thread) projects itself, walks its window in Frame::GetFeaturesInArea's order and keeps the
//      candidates that could ever be accepted -- free on entry and within ORBdist -- as a PREFERENCE LIST sorted by
//      (distance, position in the walk): the reference's scan picks the first minimum among the keypoints still free;
//   2. rounds (one CTA per pair): every undecided map point i marks each free keypoint of its list with the smallest
//      interested index (atomicMin) and looks at its first free entry b; i is FINAL iff no undecided j < i has b anywhere in
//      its list -- no earlier point can ever take b, and losing other entries does not change i's first choice.  The
//      smallest undecided index is always final, so the rounds terminate; dense frames need a handful;
//   3. rotation histogram over the assigned keypoints, ComputeThreeMaxima, the other bins dropped (:1420-1441).
// A map point with more than `cap` acceptable candidates makes the pair fall back to the literal sequential scan on one
// thread (never seen with ORBdist <= 100: unrelated descriptors are 128 +- 8 bits apart).
struct ProjSearchArgs {
    const KfViewDev* views;
    const float* kp_xy;
    const int32_t* kp_octave;
    const float* kp_angle;
    const uint32_t* desc;
    const uint8_t* mp_valid;
    const float* mp_xyz;
    const uint32_t* mp_desc;
    const float* mp_maxdist;
    const float* mp_mindist;
    const int32_t* grid_off;
    const int32_t* grid_idx;
    int32_t C;
    const int32_t* vframe;             // [C] view index of the frame
    const int32_t* vkf;                // [C] view index of the keyframe
    const float* K;                    // [C][4]
    const float* Rcw;                  // [C][9] CurrentFrame.mTcw
    const float* tcw;                  // [C][3]
    float th;
    int32_t orb_dist, check_orientation, cap;
    const int64_t* offF;               // [C+1] frame-indexed arrays
    const int64_t* offK;               // [C+1] keyframe-indexed arrays
    const uint8_t* occupied;           // frame-indexed (or nullptr)
    const uint8_t* already_found;      // keyframe-indexed (or nullptr)
    int32_t* cand;                     // [sumK][cap] preference lists (frame keypoint indices)
    int32_t* cand_n;                   // [sumK]
    uint8_t* taken;                    // frame-indexed
    int32_t* minidx;                   // frame-indexed
    int32_t* frame_match;              // frame-indexed output
    int32_t* nmatches;                 // [C]
    int32_t* overflow;                 // [C] (zeroed by the host)
    int32_t* rounds;                   // [C] diagnostic
};

// the projection of keyframe feature i of pair c into the frame: false if the reference `continue`s before the window search
__device__ __forceinline__ bool proj_project(const ProjSearchArgs& a, int c, const KfViewDev& fr, size_t g, float& u, float& v, int& lvl, float& radius)
{
    const float* Rcw = a.Rcw + 9 * (size_t)c;
    const float* tcw = a.tcw + 3 * (size_t)c;
    const float fx = a.K[4 * c], fy = a.K[4 * c + 1], cx = a.K[4 * c + 2], cy = a.K[4 * c + 3];
    const float pw[3] = {a.mp_xyz[3 * g], a.mp_xyz[3 * g + 1], a.mp_xyz[3 * g + 2]};
    float pc[3];
    guided_mat3_vec(Rcw, pw, tcw, pc);
    const float invzc = (float)(1.0 / (double)pc[2]);
    u = fx * pc[0] * invzc + cx;
    v = fy * pc[1] * invzc + cy;
    if (u < fr.bounds[0] || u > fr.bounds[1]) return false;
    if (v < fr.bounds[2] || v > fr.bounds[3]) return false;
    float Ow[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Ow[i] = (-Rcw[i] * tcw[0] + -Rcw[3 + i] * tcw[1]) + -Rcw[6 + i] * tcw[2];
    const float P0 = pw[0] - Ow[0], P1 = pw[1] - Ow[1], P2 = pw[2] - Ow[2];
    const float dist3D = sqrtf((P0 * P0 + P1 * P1) + P2 * P2);
    const float maxD = 1.2f * a.mp_maxdist[g], minD = 0.8f * a.mp_mindist[g];
    if (dist3D < minD || dist3D > maxD) return false;
    lvl = guided_predict_scale(a.mp_maxdist[g], dist3D, fr.log_scale_factor, fr.n_levels);
    radius = a.th * fr.scale_factors[lvl];
    return true;
}

// Frame::GetFeaturesInArea(u, v, r, lvl - 1, lvl + 1) (Frame.cpp:393-446): calls f(idx) for every keypoint, in its order
template <typename F>
__device__ __forceinline__ void proj_walk_window(const ProjSearchArgs& a, const KfViewDev& fr, float u, float v, float r, int lvl, F f)
{
    const float mnMinX = fr.bounds[0], mnMinY = fr.bounds[2];
    const int nMinCellX = max(0, (int)floorf((u - mnMinX - r) * fr.grid_w_inv));
    const int nMaxCellX = min(fr.grid_cols - 1, (int)ceilf((u - mnMinX + r) * fr.grid_w_inv));
    const int nMinCellY = max(0, (int)floorf((v - mnMinY - r) * fr.grid_h_inv));
    const int nMaxCellY = min(fr.grid_rows - 1, (int)ceilf((v - mnMinY + r) * fr.grid_h_inv));
    if (nMinCellX >= fr.grid_cols || nMaxCellX < 0 || nMinCellY >= fr.grid_rows || nMaxCellY < 0) return;
    const int minLevel = lvl - 1, maxLevel = lvl + 1;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
            const int cell = fr.goff_off + ix * fr.grid_rows + iy;
            for (int j = a.grid_off[cell]; j < a.grid_off[cell + 1]; ++j) {
                const int idx = a.grid_idx[fr.gidx_off + j];
                const size_t gd = (size_t)fr.feat_off + idx;
                if (bCheckLevels) {
                    const int oct = a.kp_octave[gd];
                    if (oct < minLevel) continue;
                    if (maxLevel >= 0 && oct > maxLevel) continue;
                }
                const float distx = a.kp_xy[2 * gd] - u, disty = a.kp_xy[2 * gd + 1] - v;
                if (fabsf(distx) < r && fabsf(disty) < r) f(idx, gd);
            }
        }
}

// 1. preference lists.  grid (ceil(maxK / 128), C)
static __global__ void __launch_bounds__(128) proj_candidates_kernel(ProjSearchArgs a)
{
    const int c = blockIdx.y;
    const KfViewDev& fr = a.views[a.vframe[c]];
    const KfViewDev& kf = a.views[a.vkf[c]];
    const uint8_t* occ = a.occupied ? a.occupied + a.offF[c] : nullptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kf.n_feat; i += gridDim.x * blockDim.x) {
        const size_t g = (size_t)kf.feat_off + i;
        int n = 0;
        int32_t* list = a.cand + ((size_t)a.offK[c] + i) * a.cap;
        float u, v, r;
        int lvl;
        if (a.mp_valid[g] && !(a.already_found && a.already_found[a.offK[c] + i]) && proj_project(a, c, fr, g, u, v, lvl, r)) {
            const uint32_t* dMP = a.mp_desc + 8 * g;
            // sorted insertion by (distance, walk position): keys of the kept entries live beside the list in registers' stead
            // -- the distances are recomputed from the list when needed is avoided by packing (dist << 20 | idx)
            bool over = false;
            proj_walk_window(a, fr, u, v, r, lvl, [&](int idx, size_t gd) {
                if (occ && occ[idx]) return;
                const int dist = guided_descriptor_distance(dMP, a.desc + 8 * gd);
                if (dist > a.orb_dist) return;
                // insert behind every entry with distance <= dist (equal distances keep walk order)
                int pos = n;
                while (pos > 0 && (list[pos - 1] >> 20) > dist) --pos;
                if (n == a.cap) {
                    over = true;
                    if (pos == n) return;            // would be last: dropped (the pair falls back anyway)
                    for (int k = n - 1; k > pos; --k) list[k] = list[k - 1];
                } else {
                    for (int k = n; k > pos; --k) list[k] = list[k - 1];
                    ++n;
                }
                list[pos] = (dist << 20) | idx;
            });
            if (over) a.overflow[c] = 1;
        }
        a.cand_n[a.offK[c] + i] = n;
    }
}

// ORBmatcher::ComputeThreeMaxima (:1445-1488) on the bin sizes
__device__ inline void guided_three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; ++i) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

__device__ __forceinline__ int proj_rot_bin(float angle_kf, float angle_f)
{
    float rot = angle_kf - angle_f;
    if (rot < 0.0f) rot += 360.0f;
    int bin = (int)roundf(rot * (1.0f / 30));
    if (bin == 30) bin = 0;
    return bin;
}

// 2 + 3. one CTA per pair
static __global__ void __launch_bounds__(256) proj_assign_kernel(ProjSearchArgs a)
{
    __shared__ int s_undecided, s_histo[30], s_ind[3], s_n;
    const int c = blockIdx.x, tid = threadIdx.x;
    const KfViewDev& fr = a.views[a.vframe[c]];
    const KfViewDev& kf = a.views[a.vkf[c]];
    const int NF = fr.n_feat, NK = kf.n_feat;
    uint8_t* taken = a.taken + a.offF[c];
    int32_t* minidx = a.minidx + a.offF[c];
    int32_t* fmatch = a.frame_match + a.offF[c];
    int32_t* cn = a.cand_n + a.offK[c];                // becomes -1 once a map point is decided
    const int32_t* cand = a.cand + (size_t)a.offK[c] * a.cap;
    const uint8_t* occ = a.occupied ? a.occupied + a.offF[c] : nullptr;
    for (int f = tid; f < NF; f += blockDim.x) { taken[f] = occ ? occ[f] : 0; fmatch[f] = -1; }
    if (tid < 30) s_histo[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
    int rounds = 0;
    if (a.overflow[c]) {
        // literal sequential scan (one thread): :1335-1417
        if (tid == 0) {
            for (int i = 0; i < NK; ++i) {
                const size_t g = (size_t)kf.feat_off + i;
                float u, v, r;
                int lvl;
                if (!a.mp_valid[g] || (a.already_found && a.already_found[a.offK[c] + i]) || !proj_project(a, c, fr, g, u, v, lvl, r)) continue;
                const uint32_t* dMP = a.mp_desc + 8 * g;
                int bestDist = 256, bestIdx2 = -1;
                proj_walk_window(a, fr, u, v, r, lvl, [&](int idx, size_t gd) {
                    if (taken[idx]) return;
                    const int dist = guided_descriptor_distance(dMP, a.desc + 8 * gd);
                    if (dist < bestDist) { bestDist = dist; bestIdx2 = idx; }
                });
                if (bestDist <= a.orb_dist) { taken[bestIdx2] = 1; fmatch[bestIdx2] = i; }
            }
        }
        __syncthreads();
    } else {
        for (;;) {
            for (int f = tid; f < NF; f += blockDim.x) minidx[f] = INT_MAX;
            if (tid == 0) s_undecided = 0;
            __syncthreads();
            for (int i = tid; i < NK; i += blockDim.x) {
                const int n = cn[i];
                if (n <= 0) continue;
                bool any = false;
                for (int k = 0; k < n; ++k) {
                    const int idx = cand[(size_t)i * a.cap + k] & 0xfffff;
                    if (!taken[idx]) { atomicMin(minidx + idx, i); any = true; }
                }
                if (!any) cn[i] = -1;                         // everything it could accept is gone: no match
            }
            __syncthreads();
            for (int i = tid; i < NK; i += blockDim.x) {
                const int n = cn[i];
                if (n <= 0) continue;
                int b = -1;
                for (int k = 0; k < n && b < 0; ++k) {
                    const int idx = cand[(size_t)i * a.cap + k] & 0xfffff;
                    if (!taken[idx]) b = idx;
                }
                RSAC_ASSERT(b >= 0 && b < NF && minidx[b] <= i);
                if (b < 0) { cn[i] = -1; continue; }                     // (cannot happen: step b left only points with a free entry)
                if (minidx[b] == i) { fmatch[b] = i; cn[i] = -1; }      // final (taken[] is written after the barrier below)
                else atomicAdd(&s_undecided, 1);
            }
            __syncthreads();
            for (int f = tid; f < NF; f += blockDim.x)
                if (fmatch[f] >= 0) taken[f] = 1;
            ++rounds;
            const int left = s_undecided;
            __syncthreads();
            RSAC_ASSERT(rounds <= NK);
            if (left == 0 || rounds > NK) break;                         // every round decides the smallest undecided index at least
        }
    }
    // 3. rotation consistency (:1405-1441)
    int mine = 0;
    if (a.check_orientation) {
        for (int f = tid; f < NF; f += blockDim.x)
            if (fmatch[f] >= 0) atomicAdd(&s_histo[proj_rot_bin(a.kp_angle[(size_t)kf.feat_off + fmatch[f]], a.kp_angle[(size_t)fr.feat_off + f])], 1);
        __syncthreads();
        if (tid == 0) guided_three_maxima(s_histo, 30, s_ind[0], s_ind[1], s_ind[2]);
        __syncthreads();
        for (int f = tid; f < NF; f += blockDim.x)
            if (fmatch[f] >= 0) {
                const int b = proj_rot_bin(a.kp_angle[(size_t)kf.feat_off + fmatch[f]], a.kp_angle[(size_t)fr.feat_off + f]);
                if (b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) fmatch[f] = -1;
                else ++mine;
            }
    } else {
        for (int f = tid; f < NF; f += blockDim.x) mine += fmatch[f] >= 0;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_n, mine);
    __syncthreads();
    if (tid == 0) { a.nmatches[c] = s_n; a.rounds[c] = rounds; }
}

// ------------------------------------------------------------------------------------------------------------------
// Optimizer::OptimizeSim3 chained behind SearchBySim3 on the device (LoopClosing::ComputeSim3, LoopClosing.cpp:309-311:
// SearchBySim3 extends vpMapPointMatches, OptimizeSim3 then uses every non-null entry).  One CTA per pair walks the KF1
// features in order and compacts the matches -- vpMatches12 on entry (matched_in >= 0) or the one SearchBySim3 just added
// (match12) -- that pass Optimizer.cpp:1107-1121 (pMP1 and pMP2 exist, neither is bad, pMP2 is observed by KF2) into the
// optimiser's flat arrays, exactly as the reference builds its vertices and edges (:1122-1176): camera-frame points
// R_iw * P_w + t_iw in float, keypoint observations, mvInvLevelSigma2[octave] = 1 / (scale_factor[octave]^2).
struct Sim3OptChainArgs {
    const KfViewDev* views;
    const float* kp_xy;
    const int32_t* kp_octave;
    const uint8_t* mp_valid;
    const float* mp_xyz;
    int32_t C;
    const int32_t* kf1;
    const int32_t* kf2;
    const int64_t* off1;
    const int32_t* matched_in;         // or nullptr
    const int32_t* match12;
    float* x1c; float* x2c; float* o1; float* o2; float* is1; float* is2;   // optimiser arrays, pair c at off1[c]
    int32_t* src;                      // KF1 feature of every compacted match
    int32_t* n_edges;                  // [C]
};

static __global__ void __launch_bounds__(128) sim3opt_from_search_kernel(Sim3OptChainArgs a)
{
    __shared__ int s_base, s_warp[4];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KfViewDev& v1 = a.views[a.kf1[c]];
    const KfViewDev& v2 = a.views[a.kf2[c]];
    const int64_t base = a.off1[c];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < v1.n_feat; i0 += blockDim.x) {
        const int i = i0 + tid;
        int i2 = -1;
        if (i < v1.n_feat) {
            const int mi = a.matched_in ? a.matched_in[base + i] : -1;
            i2 = mi != -1 ? mi : a.match12[base + i];                  // -2: a MapPoint KF2 does not observe (i2 < 0: skipped, :1110,1114)
            if (i2 >= v2.n_feat) i2 = -1;
            if (i2 >= 0 && !(a.mp_valid[(size_t)v1.feat_off + i] && a.mp_valid[(size_t)v2.feat_off + i2])) i2 = -1;
        }
        const unsigned m = __ballot_sync(0xffffffffu, i2 >= 0);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int k = before + __popc(m & ((1u << lane) - 1u));
        if (i2 >= 0) {
            const size_t g1 = (size_t)v1.feat_off + i, g2 = (size_t)v2.feat_off + i2, o = (size_t)base + k;
            const float p1[3] = {a.mp_xyz[3 * g1], a.mp_xyz[3 * g1 + 1], a.mp_xyz[3 * g1 + 2]};
            const float p2[3] = {a.mp_xyz[3 * g2], a.mp_xyz[3 * g2 + 1], a.mp_xyz[3 * g2 + 2]};
            float c1[3], c2[3];
            guided_mat3_vec(v1.Rcw, p1, v1.tcw, c1);
            guided_mat3_vec(v2.Rcw, p2, v2.tcw, c2);
            a.x1c[3 * o] = c1[0]; a.x1c[3 * o + 1] = c1[1]; a.x1c[3 * o + 2] = c1[2];
            a.x2c[3 * o] = c2[0]; a.x2c[3 * o + 1] = c2[1]; a.x2c[3 * o + 2] = c2[2];
            a.o1[2 * o] = a.kp_xy[2 * g1]; a.o1[2 * o + 1] = a.kp_xy[2 * g1 + 1];
            a.o2[2 * o] = a.kp_xy[2 * g2]; a.o2[2 * o + 1] = a.kp_xy[2 * g2 + 1];
            const float sf1 = v1.scale_factors[a.kp_octave[g1]], sf2 = v2.scale_factors[a.kp_octave[g2]];
            a.is1[o] = 1.0f / (sf1 * sf1);                              // ORBextractor: mvLevelSigma2 = scale^2, mvInvLevelSigma2 = 1.0f / it
            a.is2[o] = 1.0f / (sf2 * sf2);
            a.src[o] = i;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 4; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) a.n_edges[c] = s_base;
}

// flags of a chained run back in the KF1 feature index space: 2 = not an edge, 1 = removed by the optimiser, 0 = kept
static __global__ void __launch_bounds__(128) sim3opt_scatter_flags_kernel(const KfViewDev* views, const int32_t* kf1, const int64_t* off1,
                                                                   const int32_t* n_edges, const uint8_t* removed, const int32_t* src,
                                                                   uint8_t* full)
{
    const int c = blockIdx.x;
    const int n1 = views[kf1[c]].n_feat;
    const int64_t base = off1[c];
    for (int i = threadIdx.x; i < n1; i += blockDim.x) full[base + i] = 2;
    __syncthreads();
    for (int k = threadIdx.x; k < n_edges[c]; k += blockDim.x) full[base + src[base + k]] = removed[base + k];
}

}  // namespace rsac
