/*
 * oracle/orc_kfdb.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * CPU restatement of the candidate retrieval of the reference:
 *   KeyFrameDatabase::DetectRelocalizationCandidates(Frame*)            src/KeyFrameDatabase.cpp:174-284
 *           (Tracking::Relocalization, Tracking.cpp:1199)
 *   KeyFrameDatabase::DetectLoopCandidates(KeyFrame, minScore)          src/KeyFrameDatabase.cpp:51-172
 *           (LoopClosing::DetectLoop, LoopClosing.cpp:135)
 * over DBoW2's L1 score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66; ORBvoc.txt is an L1_NORM / TF_IDF
 * vocabulary) and KeyFrame::GetBestCovisibilityKeyFrames(10) (src/KeyFrame.cpp:161-169).
 *
 * The database is what the two functions read: per keyframe its BowVector (std::map<WordId, double>: word ids
 * ascending) and its ten best covisible keyframes; the inverted file (mvInvertedFile[word] = list of keyframes in
 * insertion order, KeyFrameDatabase.cpp:19-26) is rebuilt here from the vectors, keyframe index = insertion order.
 * The walk below is the reference's, loop for loop: the ORDER of lKFsSharingWords (first encounter while walking the
 * query's words in ascending order and each word's list in insertion order) decides the order of the returned
 * candidates, and candidate order is what Tracking::Relocalization's "first candidate that verifies" depends on.
 *
 * Quirk kept (Q11): DetectRelocalizationCandidates adds pKF2->mRelocScore of every covisible keyframe that merely SHARES
 * a word with the frame (mnRelocQuery == F->mnId, :243-244) although only keyframes with more than minCommonWords
 * were scored in this query (:217-224): the others contribute the score of the last query that scored them.  The
 * member is not initialised by the KeyFrame constructor (src/KeyFrame.cpp:15); the restatement carries it as explicit
 * state (score_state, in/out), zero before the first query.  DetectLoopCandidates has no such read (:131).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "orc.h"

/* seconds the last orc_detect_candidates call spent AFTER rebuilding the inverted file (the reference maintains its
 * inverted file incrementally, so only the walk and the scoring are its per-query cost); bench.py's CPU figure */
static double g_last_query_seconds = 0.0;
double orc_kfdb_last_query_seconds(void) { return g_last_query_seconds; }
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* L1Scoring::score (ScoringObject.cpp:23-66) on two ascending (word, value) arrays */
double orc_bow_l1_score(int n1, const uint32_t *w1, const double *v1, int n2, const uint32_t *w2, const double *v2)
{
    int i = 0, j = 0;
    double score = 0;
    while (i < n1 && j < n2) {
        if (w1[i] == w2[j]) {
            const double vi = v1[i], wi = v2[j];
            score += fabs(vi - wi) - fabs(vi) - fabs(wi);
            ++i; ++j;
        } else if (w1[i] < w2[j]) {
            while (i < n1 && w1[i] < w2[j]) ++i;      /* lower_bound(v2_it->first) */
        } else {
            while (j < n2 && w2[j] < w1[i]) ++j;
        }
    }
    score = -score / 2.0;
    return score;
}

/* inverted file as CSR: inv_off[word] .. inv_off[word+1] = keyframes holding `word`, in insertion (= index) order */
static int build_inverted(const orc_kfdb *db, uint32_t *max_word_out, int64_t **off_out, int32_t **kf_out)
{
    uint32_t mw = 0;
    const int64_t nnz = db->bow_off[db->K];
    for (int64_t i = 0; i < nnz; i++)
        if (db->bow_word[i] > mw) mw = db->bow_word[i];
    int64_t *off = (int64_t *)calloc((size_t)mw + 2, sizeof(int64_t));
    int32_t *kf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    if (!off || !kf) { free(off); free(kf); return -1; }
    for (int64_t i = 0; i < nnz; i++) off[db->bow_word[i] + 1]++;
    for (uint32_t w = 0; w <= mw; w++) off[w + 1] += off[w];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * ((size_t)mw + 1));
    if (!cur) { free(off); free(kf); return -1; }
    memcpy(cur, off, sizeof(int64_t) * ((size_t)mw + 1));
    for (int k = 0; k < db->K; k++)                                  /* KeyFrameDatabase::add in insertion order */
        for (int64_t i = db->bow_off[k]; i < db->bow_off[k + 1]; i++) kf[cur[db->bow_word[i]]++] = k;
    free(cur);
    *max_word_out = mw; *off_out = off; *kf_out = kf;
    return 0;
}

/* mode 0: DetectRelocalizationCandidates; mode 1: DetectLoopCandidates (conn = spConnectedKeyFrames, min_score).
 * out[cap] receives the candidates in the reference's order; returns their number (or -1).
 * score_state: mode 0 only, [K] in/out (NULL: no state is kept, unscored neighbours contribute 0). */
int orc_detect_candidates(const orc_kfdb *db, int mode, int nq, const uint32_t *qword, const double *qval, int n_conn,
                          const int32_t *conn, float min_score, float *score_state, int32_t *out, int cap)
{
    const int K = db->K;
    uint32_t mw = 0;
    int64_t *inv_off = NULL;
    int32_t *inv_kf = NULL;
    if (build_inverted(db, &mw, &inv_off, &inv_kf)) return -1;
    const double t_query0 = now_s();
    int *query = (int *)calloc((size_t)(K > 0 ? K : 1), sizeof(int));        /* mnRelocQuery / mnLoopQuery == this query */
    int *words = (int *)calloc((size_t)(K > 0 ? K : 1), sizeof(int));        /* mnRelocWords / mnLoopWords */
    char *connected = (char *)calloc((size_t)(K > 0 ? K : 1), 1);
    int32_t *share = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));   /* lKFsSharingWords */
    float *score = (float *)calloc((size_t)(K > 0 ? K : 1), sizeof(float));  /* mRelocScore / mLoopScore */
    int32_t *sm_kf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));   /* lScoreAndMatch */
    float *sm_s = (float *)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
    int32_t *acc_kf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));  /* lAccScoreAndMatch */
    float *acc_s = (float *)malloc(sizeof(float) * (size_t)(K > 0 ? K : 1));
    char *added = (char *)calloc((size_t)(K > 0 ? K : 1), 1);
    int n_out = -1;
    if (!query || !words || !connected || !share || !score || !sm_kf || !sm_s || !acc_kf || !acc_s || !added) goto done;
    if (mode == 0 && score_state) memcpy(score, score_state, sizeof(float) * (size_t)K);
    for (int i = 0; i < n_conn; i++)
        if (conn[i] >= 0 && conn[i] < K) connected[conn[i]] = 1;

    /* keyframes sharing a word with the query (:181-200 / :58-82) */
    int n_share = 0;
    for (int i = 0; i < nq; i++) {
        const uint32_t w = qword[i];
        if (w > mw) continue;
        for (int64_t p = inv_off[w]; p < inv_off[w + 1]; p++) {
            const int k = inv_kf[p];
            if (!query[k]) {
                words[k] = 0;
                if (mode == 0 || !connected[k]) {
                    query[k] = 1;
                    share[n_share++] = k;
                }
            }
            words[k]++;
        }
    }
    n_out = 0;
    if (n_share == 0) goto done;

    int maxCommonWords = 0;
    for (int i = 0; i < n_share; i++)
        if (words[share[i]] > maxCommonWords) maxCommonWords = words[share[i]];
    const int minCommonWords = (int)((float)maxCommonWords * 0.8f);

    /* similarity scores (:213-225 / :100-117) */
    int n_sm = 0;
    for (int i = 0; i < n_share; i++) {
        const int k = share[i];
        if (words[k] > minCommonWords) {
            const float si = (float)orc_bow_l1_score(nq, qword, qval, (int)(db->bow_off[k + 1] - db->bow_off[k]),
                                                     db->bow_word + db->bow_off[k], db->bow_val + db->bow_off[k]);
            score[k] = si;
            if (mode == 0 || si >= min_score) { sm_kf[n_sm] = k; sm_s[n_sm] = si; n_sm++; }
        }
    }
    if (n_sm == 0) goto done;

    /* accumulate by covisibility (:233-259 / :125-148) */
    float bestAccScore = mode == 0 ? 0.0f : min_score;
    for (int i = 0; i < n_sm; i++) {
        const int k = sm_kf[i];
        float bestScore = sm_s[i];
        float accScore = sm_s[i];
        int best = k;
        for (int j = 0; j < 10; j++) {
            const int k2 = db->covis[(size_t)k * 10 + j];
            if (k2 < 0) break;                                   /* fewer than ten connected keyframes */
            if (k2 >= K) continue;
            if (mode == 0) {
                if (!query[k2]) continue;                        /* mnRelocQuery != F->mnId */
            } else {
                if (!(query[k2] && words[k2] > minCommonWords)) continue;
            }
            accScore += score[k2];
            if (score[k2] > bestScore) { best = k2; bestScore = score[k2]; }
        }
        acc_kf[i] = best;
        acc_s[i] = accScore;
        if (accScore > bestAccScore) bestAccScore = accScore;
    }

    /* everything above 0.75 of the best accumulated score, each keyframe once (:262-281 / :151-169) */
    const float minScoreToRetain = 0.75f * bestAccScore;
    for (int i = 0; i < n_sm; i++) {
        if (acc_s[i] > minScoreToRetain) {
            const int k = acc_kf[i];
            if (!added[k]) {
                if (n_out < cap) out[n_out] = k;
                n_out++;
                added[k] = 1;
            }
        }
    }
    if (mode == 0 && score_state) memcpy(score_state, score, sizeof(float) * (size_t)K);
done:
    g_last_query_seconds = now_s() - t_query0;
    free(inv_off); free(inv_kf); free(query); free(words); free(connected); free(share); free(score);
    free(sm_kf); free(sm_s); free(acc_kf); free(acc_s); free(added);
    return n_out;
}
