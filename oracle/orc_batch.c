/*
 * oracle/orc_batch.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * CPU-baseline drivers for bench.py: the oracle solvers run over C independent
 * problems, single-threaded as the reference ships (mode A, nthreads = 1) or one
 * solver call per core across candidates (mode B), BASELINE.md section 2.  The
 * timed region is the iterate() work only, the boundary the reference itself
 * instruments (LoopClosing.cpp:285-288, Tracking.cpp:313-316).  Per-problem
 * index tables replace the shared rand() stream (Q8) so threads do not race.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "orc.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int kind;   /* 0 pnp, 1 sim3, 2 mlpnp, 3 score */
    int C;
    const void *pbs;
    const orc_ransac_params *prm;
    double prob; int min_inliers; int max_its;
    const uint32_t *const *tables;
    int flags;
    orc_result *res;
    uint8_t **masks;    /* optional: per-problem inlier masks (n bytes each) */
    long long *evals;   /* per task */
    int next;           /* atomic task counter */
    /* score */
    const float *max_err; const float *poses; int H; int *counts;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        const int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (j->kind == 3) {
            /* scoring: tasks are blocks of 64 hypotheses */
            const int h0 = i * 64;
            if (h0 >= j->H) break;
            const int h1 = h0 + 64 < j->H ? h0 + 64 : j->H;
            const orc_pnp_problem *pb = (const orc_pnp_problem *)j->pbs;
            orc_pnp_score(pb, j->max_err, h1 - h0, j->poses + (size_t)h0 * 12, NULL, j->counts + h0);
            continue;
        }
        if (i >= j->C) break;
        if (j->kind == 0) {
            const orc_pnp_problem *pb = (const orc_pnp_problem *)j->pbs + i;
            orc_pnp_ransac(pb, j->prm, j->tables[i], j->flags, &j->res[i], j->masks ? j->masks[i] : NULL, NULL, NULL);
            j->evals[i] = (long long)(j->res[i].n_hyp + j->res[i].n_refines) * pb->n;
        } else if (j->kind == 1) {
            const orc_sim3_problem *pb = (const orc_sim3_problem *)j->pbs + i;
            orc_sim3_ransac(pb, j->prob, j->min_inliers, j->max_its, j->tables[i], j->flags, &j->res[i], j->masks ? j->masks[i] : NULL, NULL, NULL);
            j->evals[i] = (long long)j->res[i].n_hyp * pb->n;
        } else {
            const orc_mlpnp_problem *pb = (const orc_mlpnp_problem *)j->pbs + i;
            orc_mlpnp_ransac(pb, j->prm, j->tables[i], j->flags, &j->res[i], j->masks ? j->masks[i] : NULL, NULL, NULL);
            j->evals[i] = (long long)(j->res[i].n_hyp + j->res[i].n_refines) * pb->n;
        }
    }
    return NULL;
}

static double run_job(job_t *j, int nthreads, long long *evals_done)
{
    if (nthreads < 1) nthreads = 1;
    j->next = 0;
    if (j->kind != 3) j->evals = (long long *)calloc((size_t)(j->C > 0 ? j->C : 1), sizeof(long long));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    const double t0 = now_s();
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, j);
    worker(j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    const double dt = now_s() - t0;
    free(th);
    if (j->kind != 3) {
        long long tot = 0;
        for (int i = 0; i < j->C; ++i) tot += j->evals[i];
        if (evals_done) *evals_done = tot;
        free(j->evals);
    }
    return dt;
}

double orc_pnp_batch(int C, const orc_pnp_problem *pbs, const orc_ransac_params *prm,
                     const uint32_t *const *tables, int flags, int nthreads,
                     orc_result *res, long long *evals_done)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 0; j.C = C; j.pbs = pbs; j.prm = prm; j.tables = tables; j.flags = flags; j.res = res;
    return run_job(&j, nthreads, evals_done);
}

/* same with the returned inlier masks (masks[i]: n_i bytes), for full-size parity tests */
double orc_pnp_batch_masks(int C, const orc_pnp_problem *pbs, const orc_ransac_params *prm,
                           const uint32_t *const *tables, int flags, int nthreads,
                           orc_result *res, uint8_t **masks, long long *evals_done)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 0; j.C = C; j.pbs = pbs; j.prm = prm; j.tables = tables; j.flags = flags; j.res = res; j.masks = masks;
    return run_job(&j, nthreads, evals_done);
}

double orc_mlpnp_batch_masks(int C, const orc_mlpnp_problem *pbs, const orc_ransac_params *prm,
                             const uint32_t *const *tables, int flags, int nthreads,
                             orc_result *res, uint8_t **masks, long long *evals_done)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 2; j.C = C; j.pbs = pbs; j.prm = prm; j.tables = tables; j.flags = flags; j.res = res; j.masks = masks;
    return run_job(&j, nthreads, evals_done);
}

double orc_sim3_batch(int C, const orc_sim3_problem *pbs, double prob, int min_inliers, int max_its,
                      const uint32_t *const *tables, int flags, int nthreads,
                      orc_result *res, long long *evals_done)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 1; j.C = C; j.pbs = pbs; j.prob = prob; j.min_inliers = min_inliers; j.max_its = max_its;
    j.tables = tables; j.flags = flags; j.res = res;
    return run_job(&j, nthreads, evals_done);
}

double orc_mlpnp_batch(int C, const orc_mlpnp_problem *pbs, const orc_ransac_params *prm,
                       const uint32_t *const *tables, int flags, int nthreads,
                       orc_result *res, long long *evals_done)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 2; j.C = C; j.pbs = pbs; j.prm = prm; j.tables = tables; j.flags = flags; j.res = res;
    return run_job(&j, nthreads, evals_done);
}

double orc_pnp_score_timed(const orc_pnp_problem *pb, const float *max_err, int H, const float *poses,
                           int nthreads, int *counts)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.kind = 3; j.pbs = pb; j.max_err = max_err; j.poses = poses; j.H = H; j.counts = counts;
    return run_job(&j, nthreads, NULL);
}
