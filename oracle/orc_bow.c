/*
 * oracle/orc_bow.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * CPU restatement of ORBmatcher::SearchByBoW, both overloads, of the reference
 * (SURVEY 8(f) N2):
 *   mode 0  SearchByBoW(KeyFrame, Frame&, vpMapPointMatches)      src/ORBmatcher.cpp:110-239
 *           (Tracking::Relocalization, Tracking.cpp:1214; TrackReferenceKeyFrame, :611)
 *   mode 1  SearchByBoW(KeyFrame1, KeyFrame2, vpMatches12)        src/ORBmatcher.cpp:354-487
 *           (LoopClosing::ComputeSim3, LoopClosing.cpp:251)
 * with DescriptorDistance (:1492-1508) and ComputeThreeMaxima (:1445-1488).
 * Integer work: the device path must match bit for bit.
 *
 * A "feature set" is what the function reads from a Frame / KeyFrame: mDescriptors
 * (256-bit ORB, 8 x int32), the keypoint angles, which features have a usable MapPoint
 * (non-null and !isBad()), and mFeatVec (DBoW2::FeatureVector = std::map<NodeId,
 * vector<unsigned>>: node ids ascending, feature indices in insertion order) as CSR arrays.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* ORBmatcher::DescriptorDistance (:1492-1508): the SWAR popcount, literally */
int orc_descriptor_distance(const uint32_t *a, const uint32_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = a[i] ^ b[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* ORBmatcher::ComputeThreeMaxima (:1445-1488) on the bin sizes */
void orc_three_maxima(const int *histo, int L, int *ind1, int *ind2, int *ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    *ind1 = *ind2 = *ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) {
            max3 = max2; max2 = max1; max1 = s;
            *ind3 = *ind2; *ind2 = *ind1; *ind1 = i;
        } else if (s > max2) {
            max3 = max2; max2 = s;
            *ind3 = *ind2; *ind2 = i;
        } else if (s > max3) {
            max3 = s;
            *ind3 = i;
        }
    }
    if (max2 < 0.1f * (float)max1) {
        *ind2 = -1; *ind3 = -1;
    } else if (max3 < 0.1f * (float)max1) {
        *ind3 = -1;
    }
}

static int lower_bound_u32(const uint32_t *a, int n, uint32_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

#define ORC_HISTO_LENGTH 30
#define ORC_TH_LOW 50

/* q = the keyframe whose features drive the outer loop (pKF / pKF1), t = the other side (F / pKF2).
 * match_out: mode 0: [t->n_feat] query feature matched to each target feature, -1 = none
 *                    (vpMapPointMatches[iF] = MapPoint of that keyframe feature)
 *            mode 1: [q->n_feat] target feature matched to each query feature, -1 = none
 *                    (vpMatches12[idx1] = MapPoint of that pKF2 feature)
 * returns nmatches */
int orc_search_by_bow(const orc_bow_features *q, const orc_bow_features *t, float nn_ratio, int check_orientation,
                      int mode, int32_t *match_out)
{
    const int n_out = mode == 0 ? t->n_feat : q->n_feat;
    for (int i = 0; i < n_out; i++) match_out[i] = -1;
    uint8_t *matched_t = (uint8_t *)calloc((size_t)(t->n_feat > 0 ? t->n_feat : 1), 1);   /* mode 0: vpMapPointMatches[iF] != null; mode 1: vbMatched2 */
    int *bin_of = (int *)malloc(sizeof(int) * (size_t)(n_out > 0 ? n_out : 1));
    int histo[ORC_HISTO_LENGTH];
    memset(histo, 0, sizeof(histo));
    const float factor = 1.0f / ORC_HISTO_LENGTH;
    int nmatches = 0;

    int qi = 0, ti = 0;
    while (qi < q->n_nodes && ti < t->n_nodes) {                         /* :131 / :383 */
        if (q->node_ids[qi] == t->node_ids[ti]) {
            for (int a = q->node_off[qi]; a < q->node_off[qi + 1]; a++) {
                const int idx_q = (int)q->node_feat[a];
                if (q->valid && !q->valid[idx_q]) continue;             /* !pMP || pMP->isBad() (:144-150 / :392-396) */
                const uint32_t *dq = q->desc + 8 * (size_t)idx_q;
                int bestDist1 = 256, bestIdx = -1, bestDist2 = 256;
                for (int b = t->node_off[ti]; b < t->node_off[ti + 1]; b++) {
                    const int idx_t = (int)t->node_feat[b];
                    if (matched_t[idx_t]) continue;                      /* :162 / :410 */
                    if (mode == 1 && t->valid && !t->valid[idx_t]) continue;   /* :410-414 */
                    const int dist = orc_descriptor_distance(dq, t->desc + 8 * (size_t)idx_t);
                    if (dist < bestDist1) {
                        bestDist2 = bestDist1; bestDist1 = dist; bestIdx = idx_t;
                    } else if (dist < bestDist2) {
                        bestDist2 = dist;
                    }
                }
                const int pass = mode == 0 ? (bestDist1 <= ORC_TH_LOW) : (bestDist1 < ORC_TH_LOW);   /* :179 vs :431 */
                if (pass && (float)bestDist1 < nn_ratio * (float)bestDist2) {
                    matched_t[bestIdx] = 1;
                    const int out_idx = mode == 0 ? bestIdx : idx_q;
                    match_out[out_idx] = mode == 0 ? idx_q : bestIdx;
                    if (check_orientation) {
                        float rot = q->angle[idx_q] - t->angle[bestIdx];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)roundf(rot * factor);
                        if (bin == ORC_HISTO_LENGTH) bin = 0;
                        bin_of[out_idx] = bin;
                        if (bin >= 0 && bin < ORC_HISTO_LENGTH) histo[bin]++;
                    }
                    nmatches++;
                }
            }
            qi++; ti++;
        } else if (q->node_ids[qi] < t->node_ids[ti]) {
            qi = lower_bound_u32(q->node_ids, q->n_nodes, t->node_ids[ti]);     /* :225 */
        } else {
            ti = lower_bound_u32(t->node_ids, t->n_nodes, q->node_ids[qi]);
        }
    }
    if (check_orientation) {                                            /* :214-234 */
        int ind1, ind2, ind3;
        orc_three_maxima(histo, ORC_HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int i = 0; i < n_out; i++) {
            if (match_out[i] < 0) continue;
            const int bin = bin_of[i];
            if (bin == ind1 || bin == ind2 || bin == ind3) continue;
            match_out[i] = -1;
            nmatches--;
        }
    }
    free(matched_t);
    free(bin_of);
    return nmatches;
}
