/*
 * ransac_b200.h -- C ABI of the B200-native batched RANSAC pose-estimation engine.
 *
 * This is the drop-in boundary for ORB-SLAM2's geometric-verification hot path
 * (reference: Luigi940260/orb-slam2-optimized).  The reference has no FFI layer: the
 * path is three C++ classes -- PnPsolver (include/PnPsolver.hpp:21-31), MLPnPsolver
 * (include/MLPnPsolver.hpp:10-21) and Sim3Solver (include/Sim3Solver.hpp:16-30) --
 * constructed per candidate keyframe by Tracking::Relocalization
 * (src/Tracking.cpp:1225-1255) and LoopClosing::ComputeSim3 (src/LoopClosing.cpp:260-308).
 * The C++ classes of the same names in include/ransac_b200/solvers.hpp keep that API and are
 * thin wrappers over the entry points below; the *_batch entry points are what the two
 * callers use to verify all candidates in one device pass.  The two consumers of the accepted poses,
 * Optimizer::PoseOptimization (src/Optimizer.cpp:205-424) and Optimizer::OptimizeSim3 (src/Optimizer.cpp:1054-1249),
 * are batched behind rsac_poseopt_* / rsac_sim3opt_* (SURVEY 8(f) N1).
 *
 * Conventions: extern "C", POD structs, plain pointers and sizes, caller-allocated
 * outputs, integer status codes (never throws across the ABI).  Host pointers unless a
 * parameter is named d_* (device pointer).  All matrices row-major.  There is no CPU
 * fallback: without a CUDA device every compute entry point returns RSAC_ERR_NO_DEVICE.
 */
#ifndef RANSAC_B200_H
#define RANSAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSAC_VERSION 100

/* status codes */
#define RSAC_OK 0
#define RSAC_ERR_INVALID 1    /* bad argument */
#define RSAC_ERR_NO_DEVICE 2  /* no usable CUDA device */
#define RSAC_ERR_CUDA 3       /* CUDA runtime error (rsac_last_error has the text) */
#define RSAC_ERR_STATE 4      /* call order (e.g. run before upload) */
#define RSAC_ERR_ALLOC 5

/* run flags */
#define RSAC_FLAG_KEEP_MASKS 1      /* also keep per-hypothesis inlier bitmasks (debug / rsac_*_get_hypotheses) */
#define RSAC_FLAG_MLPNP_DISCARD_REFINE 4 /* reproduce MLPnPsolver::Refine not storing its pose (MLPnPsolver.cpp:290-296) */
#define RSAC_FLAG_EPNP_EIGEN 8      /* 4-point EPnP: null-space basis from the 12x12 M^T M eigen-solve (PnPsolver.cpp:380)
                                       instead of the default Householder QR of M^T (same subspace, different basis) */
#define RSAC_FLAG_EARLY_EXIT 16     /* PnP and MLPnP: stop where the sequential reference stops (PnPsolver.cpp:225-236 and
                                       MLPnPsolver.cpp:144-160 return at the first successful Refine): hypotheses are solved
                                       and scored in stages -- the first
                                       `first_phase` of every problem, the next ones, then the rest, only for the problems
                                       that still need them.
                                       Results are identical to the exhaustive run; hypotheses behind the stopping point
                                       are simply never computed (rsac_pnp_get_hypotheses returns zeros/stale data there;
                                       rsac_pnp_rerun computes them on demand).  Works in both null-space modes. */

typedef struct rsac_engine rsac_engine;

/* RANSAC parameters exactly as PnPsolver::SetRansacParameters takes them
 * (PnPsolver.hpp:26-27; MLPnPsolver.hpp:16-17 with min_set = 6) */
typedef struct {
    double prob;
    int32_t min_inliers;
    int32_t max_its;
    int32_t min_set;
    float eps;
    float th2;
} rsac_ransac_params;

/* Sim3Solver::SetRansacParameters(probability, minInliers, maxIterations) (Sim3Solver.hpp:23)
 * + the bFixScale constructor flag of upstream ORB-SLAM2 (absent from the reference,
 * which is fixed-scale only: Sim3Solver.cpp:250) */
typedef struct {
    double prob;
    int32_t min_inliers;
    int32_t max_its;
    int32_t fix_scale;
} rsac_sim3_params;

/* Per-problem outcome: what iterate()/find() return plus what the getters expose.
 * 96-byte POD; this is the record gathered across GPUs. */
typedef struct {
    int32_t ok;          /* return value of iterate()/find() */
    int32_t no_more;     /* bNoMore */
    int32_t n_inliers;   /* nInliers (0 when !ok) */
    int32_t best_hyp;    /* hypothesis that set the best pose (-1 if none) */
    int32_t refined;     /* pose comes from Refine() */
    int32_t n_refines;   /* Refine() calls the sequential reference would have made */
    int32_t best_count;  /* mnBestInliers */
    int32_t n_hyp;       /* hypotheses the sequential reference would have evaluated */
    float R[9];          /* rotation (PnP/MLPnP: Tcw top-left; Sim3: GetEstimatedRotation) */
    float t[3];          /* translation */
    float s;             /* Sim3 scale (1 for PnP/MLPnP and fix_scale) */
    int32_t problem;     /* global problem index (set by the sharding layer) */
    int32_t reserved[2];
} rsac_result;

/* ------------------------------------------------------------------ engine */
int rsac_version(void);
int rsac_device_count(void);
/* creates an engine on `device` with its own non-blocking stream */
int rsac_create(int device, rsac_engine** out);
void rsac_destroy(rsac_engine* e);
const char* rsac_last_error(rsac_engine* e);
/* use an external cudaStream_t (passed as void*) for all work; NULL restores the own stream */
int rsac_set_stream(rsac_engine* e, void* cuda_stream);
int rsac_sync(rsac_engine* e);
/* global index of this engine's problem 0 (written to rsac_result.problem; used when candidates are sharded) */
int rsac_set_problem_base(rsac_engine* e, int base);
/* per-problem global indices written to rsac_result.problem instead of base + local index: ids[C] for the NEXT batches of
 * exactly C problems (a rank whose batch concatenates its shards of several sweeps); C = 0 switches back to the base */
int rsac_set_problem_ids(rsac_engine* e, const int32_t* ids, int C);
/* CUDA graphs for the staged PnP sweep (default on; RSAC_GRAPH=0 in the environment turns them off): the second run of
 * a (batch shape, flags, output buffer) combination captures its 18 launches, later runs replay them with one call */
int rsac_set_graphs(rsac_engine* e, int on);
/* RSAC_FLAG_EARLY_EXIT: hypotheses per problem in the first stage (later stages double); 0 (default) = two full waves
 * of the minimal-solver kernel over the batch (a batch that fits one wave runs all hypotheses at once) */
int rsac_set_first_phase(rsac_engine* e, int hypotheses);
/* two explicit boundaries: hypotheses [0, first) for every problem, [first, second) for the problems still without an
 * acceptable hypothesis, [second, H) for those still without one after that (second >= H: two stages only, the lower
 * latency for a single sweep); 0, 0 = automatic */
int rsac_set_phases(rsac_engine* e, int first, int second);
/* any number of stages (n <= 7 boundaries, strictly increasing; [b(n-1), H) is the last stage); n = 0: automatic =
 * first stage two full waves of the minimal-solver kernel, every further stage doubles the hypotheses a problem has (cfg4: 46,
 * 92, 184, 300); fewer stages = lower latency of a single sweep, more stages = less work when sweeps overlap */
int rsac_set_stages(rsac_engine* e, int n, const int32_t* bounds);
/* diagnostic (synchronises): after an early-exit run: out[0] = first_phase used (0: the run was exhaustive),
 * out[1] = problems that went on to the second stage, out[2] = problems handed to the clean-up phase by the
 * replay, out[3] = hypotheses solved and scored in total */
int rsac_pnp_phase_stats(rsac_engine* e, int64_t out[4]);
/* pinned host memory for callers that want asynchronous H2D/D2H */
int rsac_host_alloc(void** ptr, uint64_t bytes);
int rsac_host_free(void* ptr);
/* device facts for roofline arithmetic */
typedef struct {
    int32_t sm_count;
    int32_t sm_clock_khz;     /* cudaDevAttrClockRate */
    int32_t mem_clock_khz;
    int32_t cc_major, cc_minor;
    uint64_t total_mem;
    char name[64];
} rsac_device_info;
int rsac_get_device_info(rsac_engine* e, rsac_device_info* info);
/* CUDA-event stopwatch on the engine stream */
int rsac_timer_begin(rsac_engine* e);
int rsac_timer_end(rsac_engine* e, float* ms);   /* synchronises */
/* per-stage kernel timing with CUDA events around each launch (off by default) */
#define RSAC_STAGE_PACK 0
#define RSAC_STAGE_RNG 1
#define RSAC_STAGE_SOLVE 2     /* minimal solver kernel */
#define RSAC_STAGE_SCORE 3     /* CheckInliers kernel */
#define RSAC_STAGE_SELECT 4    /* replay + refine kernel */
#define RSAC_STAGE_COUNT 5
int rsac_profile_enable(rsac_engine* e, int on);
int rsac_profile_reset(rsac_engine* e);
int rsac_profile_get(rsac_engine* e, int stage, double* total_ms, int64_t* launches);
/* the profiled launches since the last reset in launch order (first 4096): stage id and milliseconds of each;
 * returns the number of entries written, -1 on error */
int rsac_profile_trace(rsac_engine* e, int max_entries, int32_t* stages, float* ms);
/* number of this library's kernels launched since creation / last reset */
int64_t rsac_launch_count(rsac_engine* e);
/* FFMA / DFMA saturating micro-kernels: measured FP32 / FP64 CUDA-core peaks (TFLOP/s) */
int rsac_measure_peaks(rsac_engine* e, double* fp32_tflops, double* fp64_tflops);

/* ------------------------------------------------ host helpers (no device needed) */
/* PnPsolver::SetRansacParameters arithmetic (PnPsolver.cpp:58-94, MLPnPsolver.cpp:185-220) */
int rsac_pnp_ransac_setup(int n, const rsac_ransac_params* p, int* min_inliers, int* max_its);
/* Sim3Solver::SetRansacParameters arithmetic (Sim3Solver.cpp:87-111) */
int rsac_sim3_ransac_setup(int n, const rsac_sim3_params* p, int* max_its);
/* H minimal sets of k distinct indices in [0,n): DUtils::Random::RandomInt
 * (Thirdparty/DBoW2/DUtils/Random.cpp:47-50) over a private glibc-TYPE_3 stream seeded
 * like srand(seed), with the reference's swap-remove draw (PnPsolver.cpp:125-138) */
int rsac_index_table(uint32_t seed, int n, int k, int H, uint32_t* out);
/* the raw stream: count outputs of rand() after srand(seed) */
int rsac_rand_stream(uint32_t seed, int count, int32_t* out);

/* ----------------------------------------------------------------- PnPsolver */
/* A batch of C independent 2D-3D problems, ragged, concatenated.  Fields are the ones the
 * PnPsolver constructor snapshots from Frame/MapPoint (PnPsolver.cpp:11-55). */
typedef struct {
    int32_t C;
    const int32_t* offsets;      /* [C+1] first correspondence of each problem */
    const float* p3d;            /* [total][3] mvP3Dw */
    const float* p2d;            /* [total][2] mvP2D  */
    const float* sigma2;         /* [total]    mvSigma2 */
    const double* K;             /* [C][4] fx, fy, cx, cy (double, PnPsolver.hpp:71) */
    const rsac_ransac_params* params;   /* [n_params]: one per problem, or one shared */
    int32_t n_params;
    const uint32_t* seeds;       /* [C] per-problem srand() seed; used when tables == NULL */
    const uint32_t* tables;      /* optional explicit minimal-set tables, concatenated H_c*min_set each */
    const int64_t* table_offsets;/* [C+1] offsets into tables (in uint32 units) */
} rsac_pnp_batch;

/* stage 1: host -> device (async on the engine stream) and device-side packing */
int rsac_pnp_upload(rsac_engine* e, const rsac_pnp_batch* b);
/* The same batch in the INDEXED wire format (first half): in a relocalisation all candidates match the SAME
 * frame against ONE map (Tracking.cpp:1196-1232: vvpMapPointMatches[i][j] pairs keypoint j of mCurrentFrame with a MapPoint),
 * so a correspondence is a (keypoint index, map-point index) pair -- 6 bytes instead of 24 -- over two tables that are
 * uploaded once and stay resident: kp_uv / kp_sigma2 = Frame::mvKeysUn[j].pt and mvLevelSigma2[mvKeysUn[j].octave]
 * (PnPsolver.cpp:33-36), mp_xyz = MapPoint::GetWorldPos() by map-point index (:38-39).  kp_uv == NULL / mp_xyz == NULL keep
 * the tables of the previous upload.  The flat arrays are gathered on the device; everything behind is rsac_pnp_upload's. */
typedef struct {
    int32_t n_keypoints;         /* <= 65536 */
    const float* kp_uv;          /* [n_keypoints][2] or NULL */
    const float* kp_sigma2;      /* [n_keypoints] */
    int32_t n_mappoints;
    const float* mp_xyz;         /* [n_mappoints][3] or NULL */
    int32_t C;
    const int32_t* offsets;      /* [C+1] */
    const uint16_t* kp_idx;      /* [total] */
    const uint32_t* mp_idx;      /* [total] */
    const double* K;             /* [4] fx, fy, cx, cy of the frame */
    const rsac_ransac_params* params;
    int32_t n_params;
    const uint32_t* seeds;
    const uint32_t* tables;
    const int64_t* table_offsets;
} rsac_pnp_indexed_batch;
int rsac_pnp_upload_indexed(rsac_engine* e, const rsac_pnp_indexed_batch* b);
/* The same batch built ON THE DEVICE from the engine's last rsac_bow_run (mode 0, every pair against the same frame): candidate
 * c = pair c, its correspondences = the frame keypoints SearchByBoW matched, in keypoint order (the order PnPsolver's constructor
 * walks vpMapPointMatches, PnPsolver.cpp:25-50), each paired with mp_index[] of the matched keyframe feature -- the matches never
 * leave the device; the host only needs their COUNTS (n_matches from rsac_bow_download) to shape the batch (Tracking.cpp:1207-1232:
 * SearchByBoW, then a PnPsolver per candidate with at least 15 matches -- candidates below min_matches become empty problems).
 * One parameter set (n_params = 1).  kp_uv / mp_xyz == NULL keep the resident tables. */
typedef struct {
    const int32_t* n_matches;    /* [C] as downloaded */
    int32_t min_matches;         /* nmatches < min_matches: vbDiscarded (15 in Relocalization) */
    int32_t n_keypoints;
    const float* kp_uv;
    const float* kp_sigma2;
    int32_t n_mappoints;
    const float* mp_xyz;
    const double* K;             /* [4] */
    const rsac_ransac_params* params;
    const uint32_t* seeds;       /* [C] */
} rsac_pnp_from_bow;
int rsac_pnp_upload_from_bow(rsac_engine* e, const rsac_pnp_from_bow* b);
/* stage 2: all hypotheses of all problems: EPnP minimal solves, CheckInliers scoring,
 * sequential-semantics replay with Refine (PnPsolver::iterate, PnPsolver.cpp:102-238).
 * Device-resident, asynchronous.  d_results_out: optional device buffer of C rsac_result
 * that also receives the records (for a collective); may be NULL. */
int rsac_pnp_run(rsac_engine* e, int flags, void* d_results_out);
/* later iterate() calls on solvers that already ran (Tracking.cpp:1239-1334 keeps calling iterate(5)
 * on a candidate whose pose failed PoseOptimization): replays only the sequential stage, starting at
 * resume_from[c] = hypotheses already consumed by problem c (host array of C ints) */
int rsac_pnp_rerun(rsac_engine* e, int flags, const int32_t* resume_from, void* d_results_out);
/* stage 3: device -> host, synchronises.  masks: optional, concatenated inlier bitmasks,
 * ceil(n_c/32) words per problem in problem order (bit i of word w = correspondence 32w+i,
 * COMPACT index; the C++ wrapper scatters to keypoint indices like PnPsolver.cpp:160-165) */
int rsac_pnp_download(rsac_engine* e, rsac_result* results, uint32_t* masks);
/* same without the final synchronisation (pinned destinations; pipelined sweeps) */
int rsac_pnp_download_async(rsac_engine* e, rsac_result* results, uint32_t* masks);
/* the reference-facing call: 1+2+3 with host buffers */
int rsac_pnp_solve(rsac_engine* e, const rsac_pnp_batch* b, int flags, rsac_result* results, uint32_t* masks);
/* parity/debug: per-hypothesis poses ([sumH][12]: R 9, t 3) and inlier counts ([sumH]) of the last run */
int rsac_pnp_get_hypotheses(rsac_engine* e, float* poses, int32_t* counts);
int64_t rsac_pnp_total_hypotheses(rsac_engine* e);

/* CheckInliers for H given poses x n correspondences of ONE problem (PnPsolver.cpp:241-268):
 * masks [H][ceil(n/32)] and counts [H].  max_err[i] = mvMaxError[i]. */
int rsac_score_pnp_upload(rsac_engine* e, int H, const float* poses, int n, const float* p3d,
                          const float* p2d, const float* max_err, const double K[4]);
int rsac_score_pnp_run(rsac_engine* e, int want_masks);
int rsac_score_pnp_download(rsac_engine* e, uint32_t* masks, int32_t* counts);
int rsac_score_pnp(rsac_engine* e, int H, const float* poses, int n, const float* p3d, const float* p2d,
                   const float* max_err, const double K[4], uint32_t* masks, int32_t* counts);
/* diagnostic: evaluations that took the exact (reference-arithmetic) path in the last scoring run */
int64_t rsac_score_exact_evals(rsac_engine* e);

/* ---------------------------------------------------------------- Sim3Solver */
/* Fields the Sim3Solver constructor snapshots from the two keyframes (Sim3Solver.cpp:6-85):
 * camera-frame points of both keyframes, level sigma^2 of both keypoints, calibrations. */
typedef struct {
    int32_t C;
    const int32_t* offsets;      /* [C+1] */
    const float* x1c;            /* [total][3] mvX3Dc1 */
    const float* x2c;            /* [total][3] mvX3Dc2 */
    const float* sigma2_1;       /* [total] mvLevelSigma2[kp1.octave]; threshold = size_t(9.210*sigma2) */
    const float* sigma2_2;       /* [total] */
    const float* K1;             /* [C][4] fx, fy, cx, cy of keyframe 1 (float, mK1) */
    const float* K2;             /* [C][4] */
    const rsac_sim3_params* params; /* [n_params] */
    int32_t n_params;
    const uint32_t* seeds;       /* [C] */
    const uint32_t* tables;      /* optional, H_c*3 each */
    const int64_t* table_offsets;
} rsac_sim3_batch;

int rsac_sim3_upload(rsac_engine* e, const rsac_sim3_batch* b);
int rsac_sim3_run(rsac_engine* e, int flags, void* d_results_out);
int rsac_sim3_download(rsac_engine* e, rsac_result* results, uint32_t* masks);
int rsac_sim3_solve(rsac_engine* e, const rsac_sim3_batch* b, int flags, rsac_result* results, uint32_t* masks);
/* per-hypothesis ([sumH][13]: R 9, t 3, s) poses, counts, and masks ([sumH][words_c]) of the last run;
 * Sim3Solver::iterate(5,...) round-robin is replayed on the host from these (Sim3Solver.cpp:113-178) */
int rsac_sim3_get_hypotheses(rsac_engine* e, float* poses, int32_t* counts, uint32_t* masks);
int64_t rsac_sim3_total_hypotheses(rsac_engine* e);

/* --------------------------------------------------------------- MLPnPsolver */
typedef struct {
    int32_t C;
    const int32_t* offsets;
    const float* p3d;            /* [total][3] */
    const float* p2d;            /* [total][2] */
    const float* sigma2;         /* [total] */
    const float* K;              /* [C][4] fx, fy, cx, cy (float, MLPnPsolver.hpp:198) */
    const double* cov;           /* optional [total][9] bearing covariances (use_cov branch, MLPnPsolver.cpp:375-388) */
    const rsac_ransac_params* params;
    int32_t n_params;
    const uint32_t* seeds;
    const uint32_t* tables;      /* optional, H_c*min_set each */
    const int64_t* table_offsets;
} rsac_mlpnp_batch;

int rsac_mlpnp_upload(rsac_engine* e, const rsac_mlpnp_batch* b);
int rsac_mlpnp_run(rsac_engine* e, int flags, void* d_results_out);
int rsac_mlpnp_rerun(rsac_engine* e, int flags, const int32_t* resume_from, void* d_results_out);
/* RSAC_FLAG_EARLY_EXIT for MLPnP (MLPnPsolver::iterate returns at the first successful Refine, MLPnPsolver.cpp:144-160): the
 * same staged scheme as rsac_pnp_run, same records and masks as the exhaustive run.  Stage boundaries: rsac_set_stages /
 * rsac_set_phases / rsac_set_first_phase; automatic: a batch that fits one wave of the solver kernel runs all its
 * hypotheses at once.  out[] as rsac_pnp_phase_stats. */
int rsac_mlpnp_phase_stats(rsac_engine* e, int64_t out[4]);
int rsac_mlpnp_download(rsac_engine* e, rsac_result* results, uint32_t* masks);
int rsac_mlpnp_solve(rsac_engine* e, const rsac_mlpnp_batch* b, int flags, rsac_result* results, uint32_t* masks);
/* per-hypothesis poses in double ([sumH][12]) and counts */
int rsac_mlpnp_get_hypotheses(rsac_engine* e, double* poses, int32_t* counts);
int64_t rsac_mlpnp_total_hypotheses(rsac_engine* e);

/* -------------------------------------------- Optimizer::PoseOptimization (batched) */
/* the consumer of every pose the RANSAC engine accepts (Tracking.cpp:1284,1300,1315 call
 * Optimizer::PoseOptimization(&mCurrentFrame) per candidate; src/Optimizer.cpp:205-424).  One problem = one frame:
 * the keypoints that have a MapPoint (Optimizer.cpp:250-323), in keypoint order.  4 rounds of <= 10
 * Levenberg-Marquardt iterations with the Huber kernel (dropped for the last round), every round restarted from
 * the initial pose on the edges the previous round classified as inliers -- g2o's step control reproduced
 * (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:59-179).  One warp per frame. */
typedef struct {
    int32_t C;
    const int32_t* offsets;      /* [C+1] */
    const float* p3d;            /* [total][3] MapPoint::GetWorldPos() */
    const float* obs;            /* [total][3] kpUn.pt.x, kpUn.pt.y, mvuRight[i] (< 0: monocular edge, else stereo) */
    const float* inv_sigma2;     /* [total] mvInvLevelSigma2[kpUn.octave] */
    const float* K;              /* [C][5] fx, fy, cx, cy, mbf */
    const float* Tcw;            /* [C][12] initial pose pFrame->mTcw: R row-major (9), t (3) */
} rsac_poseopt_batch;

typedef struct {
    int32_t n_inliers;           /* return value of PoseOptimization: nInitialCorrespondences - nBad */
    int32_t n_bad;
    int32_t rounds;              /* outer rounds executed (4; 1 with fewer than 10 edges; 0 with fewer than 3) */
    int32_t iterations;          /* LM iterations over all rounds */
    int32_t trials;              /* LM trials (6x6 solves) over all rounds */
    int32_t problem;             /* global frame index (rsac_set_problem_base + local index): the sharding layer's key */
    double R[9], t[3];           /* SE3quat_recov.to_homogeneous_matrix() */
    float Rf[9], tf[3];          /* Converter::toIso(...) as handed to Frame::SetPose (Optimizer.cpp:418-421) */
} rsac_poseopt_result;

int rsac_poseopt_upload(rsac_engine* e, const rsac_poseopt_batch* b);
/* Device-side chaining behind a PnP sweep (Tracking.cpp:1258-1284: the inliers of the accepted RANSAC pose become the
 * frame's map points, the pose becomes mTcw, then PoseOptimization): frame c = PnP problem c of the engine's last run,
 * edges = the correspondences of its final inlier mask (monocular), initial pose = its result; problems without a pose
 * get an empty frame.  Nothing crosses PCIe.  After rsac_poseopt_run, rsac_poseopt_download returns one flag per PnP
 * correspondence: 0 inlier, 1 outlier, 2 not an edge.  bf is only recorded (no stereo edges arise). */
int rsac_poseopt_from_pnp(rsac_engine* e, float bf);
int rsac_poseopt_run(rsac_engine* e);
/* outlier: [total] pFrame->mvbOutlier of the matched keypoints (1 = outlier) */
int rsac_poseopt_download(rsac_engine* e, rsac_poseopt_result* results, uint8_t* outlier);
int rsac_poseopt_solve(rsac_engine* e, const rsac_poseopt_batch* b, rsac_poseopt_result* results, uint8_t* outlier);

/* ------------------------------------------------ Optimizer::OptimizeSim3 (batched) */
/* SURVEY 8(f) N1, second half: LoopClosing::ComputeSim3 refines every Sim3 that RANSAC accepts
 * (LoopClosing.cpp:311: Optimizer::OptimizeSim3(mpCurrentKF, pKF, vpMapPointMatches, gScm, 10); src/Optimizer.cpp:1054-1249).
 * One problem = one keyframe pair: the matches that pass the validity tests of Optimizer.cpp:1107-1143, in order.
 * 5 LM iterations with Huber(sqrt(th2)), matches with a chi2 > th2 on either edge removed, 5 (no outlier) or 10 more
 * iterations, final inlier count; g2o's numeric Jacobians (central differences, delta 1e-9) and step control
 * reproduced.  One warp per pair. */
typedef struct {
    int32_t C;
    const int32_t* offsets;      /* [C+1] */
    const float* x1c;            /* [total][3] R1w*P3D1w + t1w (map point of keyframe 1 in camera 1, float as in the reference) */
    const float* x2c;            /* [total][3] R2w*P3D2w + t2w */
    const float* obs1;           /* [total][2] pKF1->mvKeysUn[i].pt */
    const float* obs2;           /* [total][2] pKF2->mvKeysUn[i2].pt */
    const float* inv_sigma2_1;   /* [total] pKF1->mvInvLevelSigma2[kpUn1.octave] */
    const float* inv_sigma2_2;   /* [total] */
    const float* K1;             /* [C][4] fx, fy, cx, cy of keyframe 1 */
    const float* K2;             /* [C][4] */
    const float* S12;            /* [C][13] g2oS12 on entry: R row-major (9), t (3), s */
    const float* th2;            /* [C] (LoopClosing passes 10) */
    const int32_t* fix_scale;    /* [C] or NULL = 1 (the reference hard-codes _fix_scale = true, Optimizer.cpp:1076) */
} rsac_sim3opt_batch;

typedef struct {
    int32_t n_inliers;           /* return value nIn (0 when fewer than 10 matches survive the first pass) */
    int32_t n_bad;               /* matches removed after the first optimisation */
    int32_t optimized;           /* 1 when the second optimisation ran and g2oS12 was written back */
    int32_t iterations;          /* LM iterations */
    int32_t trials;              /* LM trials */
    int32_t problem;             /* global pair index (rsac_set_problem_base + local index) */
    double R[9], t[3], s;        /* g2oS12 on return (rotation().toRotationMatrix(), translation(), scale()) */
    double q[4];                 /* rotation() as (w, x, y, z), not normalised (g2o::Sim3 never normalises) */
} rsac_sim3opt_result;

int rsac_sim3opt_upload(rsac_engine* e, const rsac_sim3opt_batch* b);
int rsac_sim3opt_run(rsac_engine* e);
/* removed: [total] 1 where the reference sets vpMatches1[idx] = nullptr */
int rsac_sim3opt_download(rsac_engine* e, rsac_sim3opt_result* results, uint8_t* removed);
int rsac_sim3opt_solve(rsac_engine* e, const rsac_sim3opt_batch* b, rsac_sim3opt_result* results, uint8_t* removed);
/* Device-side chaining behind SearchBySim3 (LoopClosing::ComputeSim3, LoopClosing.cpp:309-311: the guided matching extends
 * vpMapPointMatches, OptimizeSim3 then runs on every non-null entry): pair c = pair c of the engine's last rsac_sim3_search_run,
 * matches = vpMatches12 on entry (matched12_in >= 0) plus the ones SearchBySim3 added, filtered as Optimizer.cpp:1107-1121 does,
 * camera-frame points / observations / inverse level sigmas gathered from the resident keyframe views, g2oS12 = the (R12, t12, s12)
 * the search ran with.  K2: [C][4] intrinsics of pKF2 (NULL: the search's K).  Nothing crosses PCIe.  Then rsac_sim3opt_run and
 * rsac_sim3opt_download_chained: flags per KF1 feature of every pair (concatenated like match12): 0 = match kept, 1 = removed by
 * the optimiser, 2 = no match; n_edges[C] (optional) = nCorrespondences. */
int rsac_sim3opt_from_search(rsac_engine* e, float th2, int fix_scale, const float* K2);
int rsac_sim3opt_download_chained(rsac_engine* e, rsac_sim3opt_result* results, uint8_t* flags, int32_t* n_edges);

/* ------------------------------------------------ ORBmatcher::SearchByBoW (batched) */
/* the producer of every correspondence set the RANSAC engine verifies.  Both overloads of the reference:
 *   mode 0  SearchByBoW(KeyFrame, Frame&, vpMapPointMatches)   src/ORBmatcher.cpp:110-239  (Tracking.cpp:611,1214)
 *   mode 1  SearchByBoW(KeyFrame1, KeyFrame2, vpMatches12)     src/ORBmatcher.cpp:354-487  (LoopClosing.cpp:251)
 * TH_LOW = 50, HISTO_LENGTH = 30 (ORBmatcher.cpp:9-10), DescriptorDistance :1492-1508, ComputeThreeMaxima :1445-1488.
 * A feature set is what the function reads from a Frame / KeyFrame; sets are uploaded once and referenced by index, so
 * the current frame of a relocalisation (or the current keyframe of a loop closure) is shared by all its candidates. */
typedef struct {
    int32_t n_feat;
    const uint32_t* desc;        /* [n_feat][8] mDescriptors rows (256-bit ORB, 32 bytes, 16-byte aligned) */
    const float* angle;          /* [n_feat] mvKeys[i].angle / mvKeysUn[i].angle */
    const uint8_t* valid;        /* [n_feat] 1 = the feature has a MapPoint that is not bad; NULL = all (a Frame) */
    int32_t n_nodes;             /* mFeatVec.size() */
    const uint32_t* node_ids;    /* [n_nodes] NodeIds, ascending (std::map order) */
    const int32_t* node_off;     /* [n_nodes + 1] */
    const uint32_t* node_feat;   /* [node_off[n_nodes]] feature indices, per node in insertion order */
    const uint32_t* mp_index;    /* optional [n_feat]: slot of every feature's MapPoint in the map-point table of
                                    rsac_pnp_upload_indexed (keyframes only; used by rsac_pnp_upload_from_bow) */
} rsac_bow_features;

typedef struct {
    int32_t n_sets;
    const rsac_bow_features* sets;
    int32_t C;                   /* pairs */
    const int32_t* query_set;    /* [C] the keyframe whose features drive the outer loop (pKF / pKF1) */
    const int32_t* target_set;   /* [C] the other side (F / pKF2) */
    float nn_ratio;              /* ORBmatcher(nnratio, ...): 0.75 in Relocalization and ComputeSim3 */
    int32_t check_orientation;   /* mbCheckOrientation */
    int32_t mode;                /* 0 / 1, see above */
} rsac_bow_batch;

int rsac_bow_upload(rsac_engine* e, const rsac_bow_batch* b);
int rsac_bow_run(rsac_engine* e);
/* matches: concatenated per pair.  mode 0: sets[target].n_feat entries, the keyframe feature matched to each frame feature
 * (vpMapPointMatches[iF] = that feature's MapPoint), -1 = none; mode 1: sets[query].n_feat entries, the pKF2 feature matched
 * to each pKF1 feature (vpMatches12[idx1] = its MapPoint).  n_matches: [C] the return values. */
int rsac_bow_download(rsac_engine* e, int32_t* matches, int32_t* n_matches);
int rsac_bow_match(rsac_engine* e, const rsac_bow_batch* b, int32_t* matches, int32_t* n_matches);

/* ------------------------------------------------ KeyFrameDatabase candidate retrieval (batched) */
/* where the candidate keyframes of every relocalisation / loop closure come from.
 *   mode 0  KeyFrameDatabase::DetectRelocalizationCandidates(Frame*)          src/KeyFrameDatabase.cpp:174-284  (Tracking.cpp:1199)
 *   mode 1  KeyFrameDatabase::DetectLoopCandidates(KeyFrame, minScore)        src/KeyFrameDatabase.cpp:51-172   (LoopClosing.cpp:135)
 * with DBoW2's L1 score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-66) and GetBestCovisibilityKeyFrames(10).
 * The database -- what the two functions read from the keyframes -- is uploaded once and stays resident; keyframe index =
 * insertion order into the inverted file (KeyFrameDatabase::add), which decides the ORDER of the returned candidates.
 * score_state: KeyFrame::mRelocScore carried in (NULL = zeros): DetectRelocalizationCandidates reads it for covisible
 * keyframes that share a word but were not scored by the current query (KeyFrameDatabase.cpp:243-246), i.e. it may hold an
 * earlier query's score; the engine keeps it on the device from batch to batch and applies the queries of a batch in order. */
typedef struct {
    int32_t n_keyframes;
    const int64_t* bow_off;      /* [n_keyframes + 1] */
    const uint32_t* bow_word;    /* KeyFrame::mBowVec word ids, ascending per keyframe (std::map order) */
    const double* bow_val;       /* word values (DBoW2::WordValue = double) */
    const int32_t* covis;        /* [n_keyframes][10] GetBestCovisibilityKeyFrames(10), -1 padded */
    const float* score_state;    /* optional [n_keyframes] */
} rsac_kfdb;

typedef struct {
    int32_t Q;                   /* queries (<= 65535) */
    const int64_t* bow_off;      /* [Q + 1] */
    const uint32_t* bow_word;    /* Frame::mBowVec / KeyFrame::mBowVec of the queries, ascending per query */
    const double* bow_val;
    int32_t mode;                /* 0 relocalisation, 1 loop detection */
    const float* min_score;      /* mode 1: [Q] minScore (LoopClosing.cpp:121-133) */
    const int64_t* conn_off;     /* mode 1: [Q + 1] */
    const int32_t* conn;         /* mode 1: pKF->GetConnectedKeyFrames() as database indices (any order) */
} rsac_kfdb_queries;

int rsac_kfdb_upload(rsac_engine* e, const rsac_kfdb* db);
int rsac_kfdb_query_upload(rsac_engine* e, const rsac_kfdb_queries* q);
int rsac_kfdb_run(rsac_engine* e);
/* counts[Q]: candidates of every query (may exceed cap); candidates[Q][cap]: the first min(count, cap) in the reference's order */
int rsac_kfdb_download(rsac_engine* e, int32_t* counts, int32_t* candidates, int32_t cap);
int rsac_kfdb_detect(rsac_engine* e, const rsac_kfdb_queries* q, int32_t* counts, int32_t* candidates, int32_t cap);
/* mRelocScore of every keyframe after the queries run so far */
int rsac_kfdb_get_state(rsac_engine* e, float* score_state);

/* ------------------------------------------------ ORBmatcher::SearchBySim3 (batched) */
/* the guided matching between Sim3Solver and Optimizer::OptimizeSim3 in LoopClosing::ComputeSim3
 * (LoopClosing.cpp:286-311: SearchBySim3(mpCurrentKF, pKF, vpMapPointMatches, R, t, 7.5)); src/ORBmatcher.cpp:948-1171.
 * A keyframe view is what the function reads from a KeyFrame and its MapPoints (KeyFrame::GetFeaturesInArea
 * src/KeyFrame.cpp:560-599, IsInImage :601-604, MapPoint::PredictScale src/MapPoint.cpp:367-382). */
typedef struct {
    int32_t n_feat;
    const float* kp_xy;          /* [n_feat][2] mvKeysUn[i].pt */
    const int32_t* kp_octave;    /* [n_feat] mvKeysUn[i].octave */
    const float* kp_angle;       /* [n_feat] mvKeysUn[i].angle (rsac_proj_search_* only; may be NULL for rsac_sim3_search_*) */
    const uint32_t* desc;        /* [n_feat][8] mDescriptors rows */
    const uint8_t* mp_valid;     /* [n_feat] the feature has a MapPoint that is not bad */
    const float* mp_xyz;         /* [n_feat][3] MapPoint::GetWorldPos() */
    const uint32_t* mp_desc;     /* [n_feat][8] MapPoint::GetDescriptor() */
    const float* mp_maxdist;     /* [n_feat] MapPoint::mfMaxDistance */
    const float* mp_mindist;     /* [n_feat] MapPoint::mfMinDistance */
    float Rcw[9], tcw[3];        /* GetRotation(), GetTranslation() */
    float bounds[4];             /* mnMinX, mnMaxX, mnMinY, mnMaxY */
    int32_t grid_cols, grid_rows;/* mnGridCols, mnGridRows */
    float grid_w_inv, grid_h_inv;/* mfGridElementWidthInv, mfGridElementHeightInv */
    const int32_t* grid_off;     /* [grid_cols*grid_rows + 1] mGrid[ix][iy] at ix*grid_rows + iy */
    const int32_t* grid_idx;     /* feature indices per cell in insertion order */
    int32_t n_levels;            /* mnScaleLevels (<= 16) */
    const float* scale_factors;  /* [n_levels] mvScaleFactors */
    float log_scale_factor;      /* mfLogScaleFactor */
} rsac_kf_view;

typedef struct {
    int32_t n_views;
    const rsac_kf_view* views;
    int32_t C;                   /* keyframe pairs */
    const int32_t* kf1;          /* [C] view index of pKF1 */
    const int32_t* kf2;          /* [C] */
    const float* K;              /* [C][4] pKF1's fx, fy, cx, cy (the reference projects with them in both directions) */
    const float* R12;            /* [C][9] */
    const float* t12;            /* [C][3] */
    const float* s12;            /* [C] or NULL = 1 (the reference is fixed-scale; upstream's s12 is supported) */
    float th;                    /* 7.5 in ComputeSim3 */
    const int32_t* matched12_in; /* optional, concatenated per pair [views[kf1[c]].n_feat]: vpMatches12 on entry as the KF2 feature
                                    index of the matched MapPoint (GetIndexInKeyFrame(pKF2)), -2 = a MapPoint KF2 does not observe,
                                    -1 = no match */
} rsac_sim3_search_batch;

/* keyframe views resident on the device: batches with n_views = 0 and views = NULL (rsac_sim3_search_batch, rsac_proj_search_batch)
 * and rsac_sim3_upload_from_views reference them by index */
int rsac_views_upload(rsac_engine* e, int n_views, const rsac_kf_view* views);
/* The Sim3Solver constructor (Sim3Solver.cpp:6-85) over the resident views: pair c = (views[kf1[c]], views[kf2[c]]) with
 * vpMatched12 given as matches12 (concatenated per pair [n_feat(kf1)]: the KF2 feature of the matched MapPoint, -1 = none).  The
 * host walks the matches (which ones survive :31-44 decides the batch's shape) and sends 12 B per correspondence; camera-frame
 * points and level sigmas are gathered on the device.  offsets_out [C+1], idx1_out / idx2_out (optional, [offsets_out[C]]):
 * mvnIndices1 and the KF2 features, i.e. what bit k of a result mask refers to.  Then rsac_sim3_run / rsac_sim3_download. */
typedef struct {
    int32_t C;
    const int32_t* kf1;
    const int32_t* kf2;
    const int32_t* matches12;
    const float* K1;             /* [C][4] */
    const float* K2;
    const rsac_sim3_params* params;
    int32_t n_params;
    const uint32_t* seeds;       /* [C] */
} rsac_sim3_from_views;
int rsac_sim3_upload_from_views(rsac_engine* e, const rsac_sim3_from_views* b, int32_t* offsets_out, int32_t* idx1_out, int32_t* idx2_out);
int rsac_sim3_search_upload(rsac_engine* e, const rsac_sim3_search_batch* b);
int rsac_sim3_search_run(rsac_engine* e);
/* match12: concatenated per pair [views[kf1[c]].n_feat]: the KF2 feature newly matched to each KF1 feature
 * (vpMatches12[i1] = vpMapPoints2[idx2], ORBmatcher.cpp:1164) or -1; n_found[C]: the return values */
int rsac_sim3_search_download(rsac_engine* e, int32_t* match12, int32_t* n_found);
int rsac_sim3_search(rsac_engine* e, const rsac_sim3_search_batch* b, int32_t* match12, int32_t* n_found);

/* ------------------------------------------------ ORBmatcher::SearchByProjection(Frame, KeyFrame, ...) (batched) */
/* SURVEY 8(f) N3, second half: src/ORBmatcher.cpp:1317-1444, called by Tracking::Relocalization for a candidate whose pose
 * PoseOptimization left with 10 <= nGood < 50 inliers (Tracking.cpp:1296: th = 10, ORBdist = 100; :1310: th = 3, ORBdist = 64).
 * The frame is a view too (its keypoints, descriptors, grid, scale pyramid; Frame::GetFeaturesInArea src/Frame.cpp:393-446);
 * the keyframe view supplies the MapPoints.  The reference's assignment is greedy in keyframe-feature order; the device
 * reproduces it exactly (csrc/guided.cuh). */
typedef struct {
    int32_t n_views;
    const rsac_kf_view* views;
    int32_t C;                     /* (frame, keyframe) pairs */
    const int32_t* frame;          /* [C] view index of CurrentFrame */
    const int32_t* kf;             /* [C] view index of the candidate keyframe */
    const float* K;                /* [C][4] CurrentFrame.fx, fy, cx, cy */
    const float* Rcw;              /* [C][9] CurrentFrame.mTcw */
    const float* tcw;              /* [C][3] */
    float th;
    int32_t orb_dist;              /* ORBdist */
    int32_t check_orientation;     /* mbCheckOrientation */
    const uint8_t* occupied;       /* optional, concatenated per pair [views[frame[c]].n_feat]: CurrentFrame.mvpMapPoints[i] != nullptr on entry */
    const uint8_t* already_found;  /* optional, concatenated per pair [views[kf[c]].n_feat]: sAlreadyFound.count(pMP) */
} rsac_proj_search_batch;

int rsac_proj_search_upload(rsac_engine* e, const rsac_proj_search_batch* b);
int rsac_proj_search_run(rsac_engine* e);
/* frame_match: concatenated per pair [views[frame[c]].n_feat]: the keyframe feature whose MapPoint this call assigned to the
 * frame keypoint (CurrentFrame.mvpMapPoints[i2] = pMP), else -1; n_matches[C]: the return values; info (optional) [2][C]:
 * pairs that took the sequential fall-back, assignment rounds */
int rsac_proj_search_download(rsac_engine* e, int32_t* frame_match, int32_t* n_matches, int32_t* info);
int rsac_proj_search(rsac_engine* e, const rsac_proj_search_batch* b, int32_t* frame_match, int32_t* n_matches);

/* ------------------------------------------------ multi-GPU (candidates shard) */
/* contiguous block partition of C problems over `world` ranks: rank r owns [*first, *first + *count) */
int rsac_shard_range(int C, int rank, int world, int* first, int* count);
/* NCCL all-gather of per-problem records (the one exchange step of the path).
 * rsac_nccl_get_unique_id on rank 0 -> broadcast the 128 bytes by any means -> rsac_nccl_init on every rank
 * (persistent communicator owned by the engine; released by rsac_nccl_destroy or rsac_destroy).
 * rsac_nccl_allgather_results: enqueued on the engine's stream, so it orders after the replay kernel without a
 * host round trip; d_send = this rank's `count_per_rank` rsac_result records on the device (e.g. the
 * d_results_out buffer of rsac_*_run; pad with problem = -1), d_gathered receives world*count_per_rank records
 * in rank order. */
int rsac_nccl_get_unique_id(void* id128);
int rsac_nccl_init(rsac_engine* e, const void* id128, int rank, int world);
int rsac_nccl_allgather_results(rsac_engine* e, const void* d_send, int count_per_rank, void* d_gathered);
int rsac_nccl_destroy(rsac_engine* e);

/* ------------------------------------------------------- host-side debug hooks */
/* The device numerical core compiled for the host (same templates, same arithmetic
 * contract).  Test-only: lets the CPU test-suite compare the solver source with the
 * oracle bit-for-bit without a GPU.  Not a fallback: nothing in the engine calls these. */
/* diagnostic: clock64() stamps of the replay kernel's phases (block 0 of the last launch) */
int rsac_debug_select_clocks(rsac_engine* e, long long out[16]);
/* diagnostic: clock64() stamps of the MLPnP 6-point solve's phases (hypothesis 0 of the last exhaustive launch): entry, null spaces +
 * weights, A^T P A built, eigen-solve, pose recovery, Gauss-Newton, exit; out[7] = Gauss-Newton iterations */
int rsac_debug_mlpnp_clocks(rsac_engine* e, long long out[8]);
/* diagnostic: clock64() stamps of the EPnP minimal solve's phases (thread 0 of block 0 of the last launch) */
int rsac_debug_solve_clocks(rsac_engine* e, long long out[16]);
/* diagnostic: globaltimer stamps (ns) of the first and last four CTAs of the last scoring launch:
 * [cta][0] entry, [1] first chunk landed, [2] poses folded, [3] last chunk done, [4] exit, [5] chunks processed */
int rsac_debug_score_clocks(rsac_engine* e, unsigned long long out[64]);
/* per CTA (first 1024) of the last scoring launch: entry ns, exit ns, chunks processed, SM id */
int rsac_debug_score_all(rsac_engine* e, unsigned long long out[4096]);
int rsac_debug_host_epnp4(const double K[4], const float p3d[12], const float p2d[8], float R[9], float t[3]);
/* the same with the default QR null space (RSAC_FLAG_EPNP_EIGEN clear) */
int rsac_debug_host_epnp4_qr(const double K[4], const float p3d[12], const float p2d[8], float R[9], float t[3]);
/* the 4 smallest eigenpairs of a symmetric 12x12 (upper triangle read): w[4], v[12][4] */
int rsac_debug_host_jacobi12(const double a[144], double w[4], double v[48]);
int rsac_debug_host_sim3(const float P1[9], const float P2[9], int fix_scale, float R[9], float t[3], float* s);
/* PoseOptimization of one frame with the device source compiled for the host (one lane, edges summed in order) */
int rsac_debug_host_poseopt(int n, const float* p3d, const float* obs, const float* inv_sigma2, const float K[5],
                            const float Tcw[12], rsac_poseopt_result* result, uint8_t* outlier);
/* OptimizeSim3 of one keyframe pair with the device source compiled for the host (one lane) */
int rsac_debug_host_sim3opt(int n, const float* x1c, const float* x2c, const float* obs1, const float* obs2,
                            const float* inv_sigma2_1, const float* inv_sigma2_2, const float K1[4], const float K2[4],
                            const float S12[13], float th2, int fix_scale, rsac_sim3opt_result* result, uint8_t* removed);
int rsac_debug_host_mlpnp6(const float K[4], const float p3d[18], const float p2d[12], const double* cov54,
                           double R[9], double t[3]);

#ifdef __cplusplus
}
#endif
#endif /* RANSAC_B200_H */
