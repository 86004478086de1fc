/*
 * oracle/orc_guided.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * CPU restatement of the guided matching of the reference:
 *   ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, R12, t12, th)     src/ORBmatcher.cpp:948-1171
 *           (LoopClosing::ComputeSim3, LoopClosing.cpp:309: th = 7.5, after Sim3Solver accepted a candidate)
 * over KeyFrame::GetFeaturesInArea (src/KeyFrame.cpp:560-599), KeyFrame::IsInImage (:601-604),
 * MapPoint::PredictScale (src/MapPoint.cpp:367-382), Get{Min,Max}DistanceInvariance (:355-365) and
 * ORBmatcher::DescriptorDistance (:1492-1508), TH_HIGH = 100 (:8).
 *
 * A "keyframe view" is what these functions read from a KeyFrame and its MapPoints.
 * Float arithmetic in the reference's order; Eigen's 3x3 * 3x1 product is taken as ((a0*x + a1*y) + a2*z) per row,
 * then + t (no FMA contraction: the reference builds without -march flags).
 * Unpinned detail (Q12): MapPoint::PredictScale calls log(ratio) unqualified on a float in a file without
 * `using namespace std`; with libstdc++ that is ::log(double), which is what is restated here (the float overload would
 * differ only when log(ratio)/logScaleFactor is within 1e-7 of an integer).
 * An upstream-only scale s12 is supported (sR12 = s12*R12, sR21 = (1/s12)*R12^T, t21 = -sR21*t12, upstream
 * ORBmatcher.cc:1105-1122); s12 = 1 is the reference's fixed-scale path bit for bit.
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

#define ORC_TH_HIGH 100

/* KeyFrame::GetFeaturesInArea (KeyFrame.cpp:560-599): indices appended to out (capacity n_feat), returns the count */
int orc_features_in_area(const orc_kf_view *kf, float x, float y, float r, int32_t *out)
{
    int n = 0;
    const float mnMinX = kf->bounds[0], mnMinY = kf->bounds[2];
    int nMinCellX = (int)floorf((x - mnMinX - r) * kf->grid_w_inv);
    if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= kf->grid_cols) return 0;
    int nMaxCellX = (int)ceilf((x - mnMinX + r) * kf->grid_w_inv);
    if (nMaxCellX > kf->grid_cols - 1) nMaxCellX = kf->grid_cols - 1;
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - mnMinY - r) * kf->grid_h_inv);
    if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= kf->grid_rows) return 0;
    int nMaxCellY = (int)ceilf((y - mnMinY + r) * kf->grid_h_inv);
    if (nMaxCellY > kf->grid_rows - 1) nMaxCellY = kf->grid_rows - 1;
    if (nMaxCellY < 0) return 0;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * kf->grid_rows + iy;
            for (int j = kf->grid_off[c]; j < kf->grid_off[c + 1]; j++) {
                const int idx = kf->grid_idx[j];
                const float distx = kf->kp_xy[2 * idx] - x;
                const float disty = kf->kp_xy[2 * idx + 1] - y;
                if (fabsf(distx) < r && fabsf(disty) < r) out[n++] = idx;
            }
        }
    return n;
}

/* MapPoint::PredictScale (MapPoint.cpp:367-382) */
int orc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels)
{
    const float ratio = max_distance / current_dist;
    int nScale = (int)ceil(log((double)ratio) / (double)log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= n_levels) nScale = n_levels - 1;
    return nScale;
}

static void mat3_vec(const float *R, const float *p, const float *t, float *o)
{
    for (int i = 0; i < 3; i++) o[i] = ((R[3 * i] * p[0] + R[3 * i + 1] * p[1]) + R[3 * i + 2] * p[2]) + t[i];
}

/* one direction of the search (:994-1067 / :1070-1150): map points of `src` projected into `dst` with (Rds, tds) applied
 * to the point in src's camera frame; intrinsics are pKF1's for both directions (:951-954) */
static void search_one_way(const orc_kf_view *src, const orc_kf_view *dst, const float *K, const float *Rds, const float *tds,
                           float th, const uint8_t *already, int32_t *match, int32_t *scratch)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    for (int i = 0; i < src->n_feat; i++) {
        match[i] = -1;
        if (!src->mp_valid[i] || already[i]) continue;           /* !pMP || vbAlreadyMatched || pMP->isBad() */
        float pc_src[3], pc[3];
        mat3_vec(src->Rcw, src->mp_xyz + 3 * i, src->tcw, pc_src);
        mat3_vec(Rds, pc_src, tds, pc);
        if (pc[2] < 0.0f) continue;                              /* compared against the double 0.0: same for floats */
        const float invz = (float)(1.0 / (double)pc[2]);         /* const float invz = 1.0/p3Dc2.z(); */
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!(u >= dst->bounds[0] && u < dst->bounds[1] && v >= dst->bounds[2] && v < dst->bounds[3])) continue;   /* IsInImage */
        const float maxDistance = 1.2f * src->mp_maxdist[i], minDistance = 0.8f * src->mp_mindist[i];
        const float dist3D = sqrtf((pc[0] * pc[0] + pc[1] * pc[1]) + pc[2] * pc[2]);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = orc_predict_scale(src->mp_maxdist[i], dist3D, dst->log_scale_factor, dst->n_levels);
        const float radius = th * dst->scale_factors[nPredictedLevel];
        const int nc = orc_features_in_area(dst, u, v, radius, scratch);
        if (nc == 0) continue;
        int bestDist = INT_MAX, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = scratch[c];
            const int oct = dst->kp_octave[idx];
            if (oct < nPredictedLevel - 1 || oct > nPredictedLevel) continue;
            const int dist = orc_descriptor_distance(src->mp_desc + 8 * (size_t)i, dst->desc + 8 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= ORC_TH_HIGH) match[i] = bestIdx;
    }
}

/* matched12_in[i1]: >= 0 the KF2 feature of vpMatches12[i1] (pMP->GetIndexInKeyFrame(pKF2)), -2 a match whose MapPoint is
 * not observed by KF2, -1 no match.  match12_out[i1]: the KF2 feature newly matched to KF1 feature i1 by this call
 * (vpMatches12[i1] = vpMapPoints2[idx2], :1164), else -1.  Returns nFound. */
int orc_search_by_sim3(const orc_kf_view *kf1, const orc_kf_view *kf2, const float K[4], const float R12[9], const float t12[3],
                       float s12, float th, const int32_t *matched12_in, int32_t *match12_out)
{
    const int N1 = kf1->n_feat, N2 = kf2->n_feat;
    uint8_t *am1 = (uint8_t *)calloc((size_t)(N1 > 0 ? N1 : 1), 1), *am2 = (uint8_t *)calloc((size_t)(N2 > 0 ? N2 : 1), 1);
    int32_t *m1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N1 > 0 ? N1 : 1)), *m2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N2 > 0 ? N2 : 1));
    int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * (size_t)((N1 > N2 ? N1 : N2) + 1));
    int nFound = -1;
    if (!am1 || !am2 || !m1 || !m2 || !scratch) goto done;
    for (int i = 0; i < N1; i++) {
        const int idx2 = matched12_in ? matched12_in[i] : -1;
        if (idx2 != -1) {
            am1[i] = 1;
            if (idx2 >= 0 && idx2 < N2) am2[idx2] = 1;
        }
    }
    float sR12[9], sR21[9], t21[3];
    const float inv_s = (float)(1.0 / (double)s12);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) { sR12[3 * r + c] = s12 * R12[3 * r + c]; sR21[3 * r + c] = inv_s * R12[3 * c + r]; }
    for (int i = 0; i < 3; i++) t21[i] = -((sR21[3 * i] * t12[0] + sR21[3 * i + 1] * t12[1]) + sR21[3 * i + 2] * t12[2]);
    search_one_way(kf1, kf2, K, sR21, t21, th, am1, m1, scratch);
    search_one_way(kf2, kf1, K, sR12, t12, th, am2, m2, scratch);
    nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        match12_out[i1] = -1;
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { match12_out[i1] = idx2; nFound++; }
    }
done:
    free(am1); free(am2); free(m1); free(m2); free(scratch);
    return nFound;
}

/* Frame::GetFeaturesInArea (src/Frame.cpp:393-446) with the level window */
static int frame_features_in_area(const orc_kf_view *f, float x, float y, float r, int minLevel, int maxLevel, int32_t *out)
{
    int n = 0;
    const float mnMinX = f->bounds[0], mnMinY = f->bounds[2];
    int nMinCellX = (int)floorf((x - mnMinX - r) * f->grid_w_inv);
    if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= f->grid_cols) return 0;
    int nMaxCellX = (int)ceilf((x - mnMinX + r) * f->grid_w_inv);
    if (nMaxCellX > f->grid_cols - 1) nMaxCellX = f->grid_cols - 1;
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - mnMinY - r) * f->grid_h_inv);
    if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= f->grid_rows) return 0;
    int nMaxCellY = (int)ceilf((y - mnMinY + r) * f->grid_h_inv);
    if (nMaxCellY > f->grid_rows - 1) nMaxCellY = f->grid_rows - 1;
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * f->grid_rows + iy;
            for (int j = f->grid_off[c]; j < f->grid_off[c + 1]; j++) {
                const int idx = f->grid_idx[j];
                if (bCheckLevels) {
                    if (f->kp_octave[idx] < minLevel) continue;
                    if (maxLevel >= 0 && f->kp_octave[idx] > maxLevel) continue;
                }
                const float distx = f->kp_xy[2 * idx] - x, disty = f->kp_xy[2 * idx + 1] - y;
                if (fabsf(distx) < r && fabsf(disty) < r) out[n++] = idx;
            }
        }
    return n;
}

/*
 * ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame, sAlreadyFound, th, ORBdist)   src/ORBmatcher.cpp:1317-1444
 * (Tracking::Relocalization, Tracking.cpp:1296 with th = 10, ORBdist = 100 and :1310 with th = 3, ORBdist = 64: more
 * matches for a candidate whose pose PoseOptimization left with 10 <= nGood < 50 inliers).
 *   frame          the current frame: keypoints, descriptors, grid, scale pyramid (view fields of the Frame)
 *   kf             the candidate keyframe: its MapPoints (mp_*), the angles of its keypoints
 *   Rcw, tcw, K    CurrentFrame.mTcw and intrinsics
 *   occupied       [frame->n_feat] CurrentFrame.mvpMapPoints[i2] != nullptr on entry
 *   already_found  [kf->n_feat] sAlreadyFound.count(pMP)
 *   frame_match    [frame->n_feat] out: the keyframe feature whose MapPoint this call assigned to the frame keypoint, else -1
 * Returns nmatches.  GREEDY and sequential: a frame keypoint taken by an earlier map point is skipped by the later ones (:1389).
 */
int orc_search_by_projection(const orc_kf_view *frame, const orc_kf_view *kf, const float K[4], const float Rcw[9], const float tcw[3],
                             float th, int orb_dist, int check_orientation, const uint8_t *occupied, const uint8_t *already_found,
                             int32_t *frame_match)
{
    const int NF = frame->n_feat;
    int nmatches = 0;
    uint8_t *taken = (uint8_t *)calloc((size_t)(NF > 0 ? NF : 1), 1);
    int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * (size_t)(NF + 1));
    int8_t *bin_of = (int8_t *)malloc((size_t)(NF > 0 ? NF : 1));
    if (!taken || !scratch || !bin_of) { free(taken); free(scratch); free(bin_of); return -1; }
    for (int i = 0; i < NF; i++) { taken[i] = occupied ? occupied[i] : 0; frame_match[i] = -1; bin_of[i] = -1; }
    /* Ow = -Rcw^T * tcw (:1323) */
    float Ow[3];
    for (int i = 0; i < 3; i++) Ow[i] = (-Rcw[i] * tcw[0] + -Rcw[3 + i] * tcw[1]) + -Rcw[6 + i] * tcw[2];
    int histo[30];
    memset(histo, 0, sizeof(histo));
    const float factor = 1.0f / 30;
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    for (int i = 0; i < kf->n_feat; i++) {
        if (!kf->mp_valid[i] || (already_found && already_found[i])) continue;
        const float *x3Dw = kf->mp_xyz + 3 * (size_t)i;
        float x3Dc[3];
        mat3_vec(Rcw, x3Dw, tcw, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / (double)x3Dc[2]);
        const float u = fx * xc * invzc + cx;
        const float v = fy * yc * invzc + cy;
        if (u < frame->bounds[0] || u > frame->bounds[1]) continue;
        if (v < frame->bounds[2] || v > frame->bounds[3]) continue;
        const float PO[3] = {x3Dw[0] - Ow[0], x3Dw[1] - Ow[1], x3Dw[2] - Ow[2]};
        const float dist3D = sqrtf((PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2]);
        const float maxDistance = 1.2f * kf->mp_maxdist[i], minDistance = 0.8f * kf->mp_mindist[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        const int nPredictedLevel = orc_predict_scale(kf->mp_maxdist[i], dist3D, frame->log_scale_factor, frame->n_levels);
        const float radius = th * frame->scale_factors[nPredictedLevel];
        const int nc = frame_features_in_area(frame, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1, scratch);
        if (nc == 0) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = scratch[c];
            if (taken[i2]) continue;                                     /* CurrentFrame.mvpMapPoints[i2] */
            const int dist = orc_descriptor_distance(kf->mp_desc + 8 * (size_t)i, frame->desc + 8 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            taken[bestIdx2] = 1;
            frame_match[bestIdx2] = i;
            nmatches++;
            if (check_orientation) {
                float rot = kf->kp_angle[i] - frame->kp_angle[bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == 30) bin = 0;
                bin_of[bestIdx2] = (int8_t)bin;
                histo[bin]++;
            }
        }
    }
    if (check_orientation) {
        int ind1, ind2, ind3;
        orc_three_maxima(histo, 30, &ind1, &ind2, &ind3);
        for (int i2 = 0; i2 < NF; i2++) {
            const int b = bin_of[i2];
            if (b >= 0 && b != ind1 && b != ind2 && b != ind3) { frame_match[i2] = -1; nmatches--; }
        }
    }
    free(taken); free(scratch); free(bin_of);
    return nmatches;
}
