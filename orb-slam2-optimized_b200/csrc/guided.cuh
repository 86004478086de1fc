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

}  // namespace rsac
