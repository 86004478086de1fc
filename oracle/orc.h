/*
 * oracle/orc.h -- CPU oracle for the RANSAC pose-estimation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / CPU baseline.
 *
 * What it is: a dependency-free plain-C restatement of the reference's
 *   src/PnPsolver.cpp, src/Sim3Solver.cpp, src/MLPnPsolver.cpp and
 *   Thirdparty/DBoW2/DUtils/Random.cpp:47-50
 * and, for SURVEY 8(f) N1, of src/Optimizer.cpp:205-424 (PoseOptimization) and :1054-1249 (OptimizeSim3) over the
 * vendored g2o's Levenberg-Marquardt, SE3/Sim3 exponential maps and numeric Jacobians (oracle/orc_poseopt.c),
 * following the reference function by function (each function cites file:line).
 *
 * PARITY PIN (round 2): `make -C oracle ref` compiles the reference's OWN sources, unmodified and from where they
 * lie under /root/reference -- src/PnPsolver.cpp, src/Sim3Solver.cpp, src/MLPnPsolver.cpp, src/KeyFrameDatabase.cpp,
 * src/ORBmatcher.cpp, Thirdparty/DBoW2/DUtils/Random.cpp, Thirdparty/DBoW2/DBoW2/{BowVector,FeatureVector,
 * ScoringObject}.cpp -- into
 * oracle/_ref/libref_solvers.so, against stand-in headers for what this image lacks (oracle/shim/: Eigen, OpenCV,
 * Frame / KeyFrame / MapPoint / ORBVocabulary).  The oracle equals that library BIT FOR BIT: per call (compute_pose on
 * 4..250 points, CheckInliers, ComputeSim3, SetRansacParameters, the L1 score, DescriptorDistance) and for whole RANSAC /
 * retrieval / matching runs (SearchByBoW, SearchBySim3, SearchByProjection(Frame, KeyFrame): match arrays equal)
 * (tests/test_cpu_reference_build.py); golden vectors generated from it are committed
 * (tests/golden/reference_build.npz) and the CUDA engine is compared with them directly
 * (tests/test_gpu_reference_golden.py).  MLPnPsolver.cpp (which the reference's own CMakeLists.txt:75 leaves out) is
 * compared with tolerances -- it calls libm and takes singular vectors from Eigen SVDs, which are defined only up to
 * sign / rotation: generated Jacobian vs the re-derived one 1e-12, computePose median 3e-12 (6-point sets, 99 % within
 * 4e-9), whole runs exact in return value / counts / stopping iteration once quirk Q6 is reproduced.
 * A second build (`make ref-lapack`) sends every dense solve to LAPACK (dsyev = Eigen's eigen-solver algorithm, dgelsd =
 * bdcSvd's class) and shares no kernel with this oracle: outcome-level agreement is asserted against it (Sim3 runs
 * identical; PnP: same candidates accepted, poses at the pixel-noise level -- DESIGN.md section 2 explains why an
 * Eigen-built reference cannot be matched more closely: EPnP depends on the signs of the 3 x 3 principal axes).
 * STILL UNPINNED: Eigen's own rounding.  Eigen is not installed and cannot be fetched, so the stand-in forwards the
 * dense solves the sources call (SelfAdjointEigenSolver, bdcSvd().solve, inverse()) to the kernels in orc_linalg.c
 * (cyclic / tournament Jacobi, Householder-QR least squares with a Jacobi-SVD fallback, cofactor inverse) and
 * evaluates products in index order; an Eigen-built binary will differ from both in rounding -- for 4-point EPnP
 * that means a different (equally valid) null-space basis per hypothesis (outcome-level
 * agreement across bases).  Optimizer.cpp (the vendored g2o is templated on Eigen throughout) is not compiled: its
 * restatement is pinned by independent checks only; the Frame / KeyFrame / MapPoint helpers ORBmatcher.cpp calls
 * (GetFeaturesInArea, PredictScale: other translation units of the reference) are the oracle's restatements (numpy / scipy / cv2, literal transcriptions, ground truth), as listed in DESIGN.md.
 *
 * Arithmetic contract shared with the CUDA kernels (DESIGN.md "arithmetic
 * contract"): only + - * / sqrt in IEEE double/float, no FMA contraction
 * (-ffp-contract=off here, -fmad=false there), identical operation order.  For
 * 4-point EPnP this is a hard requirement.
 */
#ifndef ORC_H
#define ORC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ RNG (R01) */
void orc_rng_seed(unsigned seed);             /* DUtils::Random::SeedRand(int), Random.cpp:33-36 */
int orc_random_int(int min, int max);         /* DUtils::Random::RandomInt, Random.cpp:47-50 */
/* H draws of k distinct indices out of [0,n) with the reference's
 * vAvailableIndices idiom (PnPsolver.cpp:125-138, Sim3Solver.cpp:136-149,
 * MLPnPsolver.cpp:76-96).  srand(seed) first.  out is H*k uint32. */
void orc_index_table(unsigned seed, int n, int k, int H, uint32_t *out);

/* ------------------------------------------------------------ small dense solves */
/* symmetric eigen-solve, eigenvalues ascending, eigenvectors = columns of v
 * (restates Eigen::SelfAdjointEigenSolver as used at PnPsolver.cpp:311,380,469,
 * Sim3Solver.cpp:238-241, MLPnPsolver.cpp:359).  a is n*n row-major, upper
 * triangle read, destroyed. */
void orc_jacobi_eig_d(int n, double *a, double *w, double *v);
void orc_jacobi_eig_f(int n, float *a, float *w, float *v);
/* the nv smallest eigenpairs of a symmetric n x n (n <= 12), recorded-rotation variant; v is n x nv */
void orc_jacobi_lowest_d(int n, int nv, double *a, double *w, double *v);
/* orthonormal basis of the null space of an 8 x 12 matrix (4-point EPnP) by Householder QR of its transpose */
void orc_nullspace_qr_d(const double *Mrows /*8x12*/, double *U4 /*12x4*/);
/* minimum-norm least squares of an m x k system through a one-sided Jacobi SVD
 * with Eigen's rank threshold (restates A.bdcSvd(ThinU|ThinV).solve(b),
 * PnPsolver.cpp:531,559,590).  L row-major m*k, m<=8, k<=6. */
void orc_svd_lstsq_d(int m, int k, const double *L, const double *b, double *x);
/* Householder-QR least squares with the SVD solve above as the rank-deficient fallback */
void orc_lstsq_d(int m, int k, const double *L, const double *b, double *x);
/* closed-form cofactor inverse (restates Matrix3d::inverse(), PnPsolver.cpp:331) */
void orc_inv3_d(const double m[9], double out[9]);
/* orthogonal polar factor U*V^T of a 3x3 (restates JacobiSVD U*V^T,
 * MLPnPsolver.cpp:511-512,570-571) */
void orc_polar3_d(const double a[9], double r[9]);
/* rank of a 3x3 by full-pivot Householder QR with Eigen's threshold
 * (restates FullPivHouseholderQR<Matrix3d>::rank(), MLPnPsolver.cpp:347,354) */
int orc_rank3_fullpiv_d(const double a[9]);
/* LDL^T solve of a symmetric 6x6 (restates Eigen::LDLT::solve, MLPnPsolver.cpp:705-706) */
void orc_ldlt6_solve_d(const double a[36], const double g[6], double x[6]);

/* --------------------------------------------------------------- RANSAC set-up */
typedef struct {
    double prob;
    int min_inliers;
    int max_its;
    int min_set;
    float eps;
    float th2;
} orc_ransac_params;

/* PnPsolver::SetRansacParameters (PnPsolver.cpp:58-94) / MLPnPsolver.cpp:185-220.
 * Returns adjusted minInliers and iteration count. */
void orc_pnp_ransac_setup(int n, const orc_ransac_params *p, int *min_inl, int *max_its);
/* Sim3Solver::SetRansacParameters (Sim3Solver.cpp:87-111) */
void orc_sim3_ransac_setup(int n, double prob, int min_inliers, int max_its_in, int *max_its);

/* ------------------------------------------------------------------- PnPsolver */
typedef struct {
    int n;                 /* valid correspondences (mvP2D.size()) */
    const float *p3d;      /* [n][3] mvP3Dw */
    const float *p2d;      /* [n][2] mvP2D */
    const float *sigma2;   /* [n]    mvSigma2 */
    double fx, fy, cx, cy; /* PnPsolver.hpp:71 keeps them double */
} orc_pnp_problem;

#define ORC_FLAG_STALE_ROWS 1   /* Q1: reproduce colwise().sum() over all allocated rows */
#define ORC_FLAG_EXHAUSTIVE 2   /* evaluate all H hypotheses (no early return), for throughput + per-hyp parity */
#define ORC_FLAG_MLPNP_DISCARD_REFINE 4 /* Q6: reproduce MLPnP Refine() not storing its pose */
#define ORC_FLAG_EPNP_QR_NULLSPACE 8    /* 4-point EPnP: null-space basis by Householder QR instead of the 12x12 eigen-solve */

typedef struct {
    int ok;           /* return value of iterate() */
    int no_more;      /* bNoMore */
    int n_inliers;    /* nInliers */
    int best_hyp;     /* index of the hypothesis that set mBestTcw (strict >, first max), -1 if none */
    int refined;      /* 1 if T is the refined pose */
    int n_hyp;        /* hypotheses actually evaluated */
    int n_refines;    /* Refine() calls made */
    int n_failed_refines;
    int best_count;   /* mnBestInliers when the call ended */
    float T[16];      /* row-major 4x4 [R t]; Sim3: mBestRotation/mBestTranslation, scale kept in `scale` */
    float scale;      /* Sim3 only */
} orc_result;

/* PnPsolver::iterate as called the first time (PnPsolver.cpp:102-191; the `||`
 * at :119 makes the first call consume the whole budget).  table is H*min_set
 * indices.  mask (n bytes, compact index, NOT scattered to keypoint indices)
 * receives the returned inlier set.  Optional per-hypothesis outputs (may be
 * NULL): hyp_counts[H], hyp_pose[H*12] (R row-major 9 + t 3, float). */
void orc_pnp_ransac(const orc_pnp_problem *pb, const orc_ransac_params *prm, const uint32_t *table,
                    int flags, orc_result *res, uint8_t *mask, int *hyp_counts, float *hyp_pose);

/* One EPnP solve on the subset idx[0..m) (PnPsolver::compute_pose, :359-415).
 * Returns the reprojection error of the chosen solution. */
double orc_epnp_pose(const orc_pnp_problem *pb, const uint32_t *idx, int m, float R[9], float t[3]);
double orc_epnp_pose_mode(const orc_pnp_problem *pb, const uint32_t *idx, int m, int flags, float R[9], float t[3]);
/* average algorithmic FP64 FLOP (+,-,*,/,sqrt = 1 each) of one minimal EPnP solve over a table of H sets */
double orc_epnp_flops(const orc_pnp_problem *pb, const uint32_t *table, int H, int min_set);
double orc_epnp_flops_mode(const orc_pnp_problem *pb, const uint32_t *table, int H, int min_set, int flags);
long long orc_flops_take(void);
/* PnPsolver::CheckInliers (:241-268) for one pose; max_err[i] = sigma2[i]*th2 (f32*f32).
 * err2 (optional) receives the f32 squared errors. */
int orc_pnp_check_inliers(const orc_pnp_problem *pb, const float *max_err, const float R[9],
                          const float t[3], uint8_t *mask, float *err2);
/* scoring stress (cfg5): H poses x n correspondences, masks as bytes [H][n], counts[H] */
void orc_pnp_score(const orc_pnp_problem *pb, const float *max_err, int H, const float *poses,
                   uint8_t *masks, int *counts);

/* test hooks for the null-space-basis experiments ("The 4-point null space") */
void orc_epnp_mtm(const orc_pnp_problem *pb, const uint32_t *idx, int m, double MtM[144]);
double orc_epnp_pose_basis(const orc_pnp_problem *pb, const uint32_t *idx, int m, const double U[48], float R[9], float t[3]);

/* ------------------------------------------------------------------ Sim3Solver */
typedef struct {
    int n;
    const float *x1c;   /* [n][3] mvX3Dc1 */
    const float *x2c;   /* [n][3] mvX3Dc2 */
    const float *sigma2_1; /* [n] level sigma^2 of kp1 (threshold = size_t(9.210*sigma2), Q4) */
    const float *sigma2_2;
    float K1[4];        /* fx, fy, cx, cy of KF1 (mK1 is Matrix3f) */
    float K2[4];
    int fix_scale;      /* reference: always 1 (Sim3Solver.cpp:250); 0 = Horn scale step (Q7, parity unpinned) */
} orc_sim3_problem;

/* Sim3Solver::iterate called repeatedly with n_its_per_call until it returns
 * true or bNoMore (Sim3Solver.cpp:113-178); because the loop condition is `&&`
 * the outcome does not depend on n_its_per_call, only calls_made does. */
void orc_sim3_ransac(const orc_sim3_problem *pb, double prob, int min_inliers, int max_its,
                     const uint32_t *table, int flags, orc_result *res, uint8_t *mask,
                     int *hyp_counts, float *hyp_pose /* H*13: R9,t3,s */);
/* Sim3Solver::ComputeSim3 (:196-266) on three pairs */
void orc_sim3_compute(const float P1[9] /*3 pts x xyz*/, const float P2[9], int fix_scale,
                      float R12[9], float t12[3], float *s12);
/* Sim3Solver::CheckInliers (:269-293) for a given (s,R,t) */
int orc_sim3_check_inliers(const orc_sim3_problem *pb, const float R12[9], const float t12[3],
                           float s12, uint8_t *mask, float *err /* optional [n][2] */);

/* ----------------------------------------------------------------- MLPnPsolver */
typedef struct {
    int n;
    const float *p3d;     /* [n][3] world points (stored as double in the solver, MLPnPsolver.cpp:40-42) */
    const float *p2d;     /* [n][2] */
    const float *sigma2;  /* [n] */
    float fx, fy, cx, cy; /* MLPnPsolver.hpp:198 keeps them float */
    const double *cov;    /* optional [n][9] bearing covariances (use_cov branch, MLPnPsolver.cpp:375-388); NULL = off */
} orc_mlpnp_problem;

void orc_mlpnp_ransac(const orc_mlpnp_problem *pb, const orc_ransac_params *prm, const uint32_t *table,
                      int flags, orc_result *res, uint8_t *mask, int *hyp_counts,
                      double *hyp_pose /* H*12 double: R9,t3 */);
/* MLPnPsolver::computePose (:321-623) on subset idx[0..m); result R(9) t(3) double */
void orc_mlpnp_pose(const orc_mlpnp_problem *pb, const uint32_t *idx, int m, double R[9], double t[3]);
/* MLPnPsolver::CheckInliers (:222-255) */
int orc_mlpnp_check_inliers(const orc_mlpnp_problem *pb, const float *max_err, const double R[9],
                            const double t[3], uint8_t *mask, float *err2);
void orc_rodrigues2rot(const double w[3], double R[9]);   /* MLPnPsolver.cpp:625-640 */
void orc_rot2rodrigues(const double R[9], double w[3]);   /* MLPnPsolver.cpp:642-657 */
/* residuals + analytic Jacobian of one point (MLPnPsolver.cpp:725-771, 773-1020) */
void orc_mlpnp_res_jac(const double pt[3], const double nr[3], const double ns[3],
                       const double w[3], const double t[3], double r[2], double J[12]);

/* -------------------------------------------------- CPU baseline drivers (bench) */
/* Runs C independent PnP problems on nthreads host threads (one solver call per
 * task, as BASELINE.md mode B; nthreads=1 is mode A).  Returns wall seconds of
 * the iterate() calls only (solver set-up excluded, like the reference's
 * timers at Tracking.cpp:313-316 / LoopClosing.cpp:285-288). */
double orc_pnp_batch(int C, const orc_pnp_problem *pbs, const orc_ransac_params *prm,
                     const uint32_t *const *tables, int flags, int nthreads,
                     orc_result *res, long long *evals_done);
double orc_pnp_batch_masks(int C, const orc_pnp_problem *pbs, const orc_ransac_params *prm,
                           const uint32_t *const *tables, int flags, int nthreads,
                           orc_result *res, uint8_t **masks, long long *evals_done);
double orc_mlpnp_batch_masks(int C, const orc_mlpnp_problem *pbs, const orc_ransac_params *prm,
                             const uint32_t *const *tables, int flags, int nthreads,
                             orc_result *res, uint8_t **masks, long long *evals_done);
double orc_sim3_batch(int C, const orc_sim3_problem *pbs, double prob, int min_inliers, int max_its,
                      const uint32_t *const *tables, int flags, int nthreads,
                      orc_result *res, long long *evals_done);
double orc_mlpnp_batch(int C, const orc_mlpnp_problem *pbs, const orc_ransac_params *prm,
                       const uint32_t *const *tables, int flags, int nthreads,
                       orc_result *res, long long *evals_done);
double orc_pnp_score_timed(const orc_pnp_problem *pb, const float *max_err, int H, const float *poses,
                           int nthreads, int *counts);

/* ------------------------------------------------ ORBmatcher::SearchByBoW */
/* what SearchByBoW reads from a Frame / KeyFrame */
typedef struct {
    int n_feat;
    const uint32_t *desc;        /* [n_feat][8] mDescriptors rows (256-bit ORB) */
    const float *angle;          /* [n_feat] mvKeys / mvKeysUn [i].angle */
    const uint8_t *valid;        /* [n_feat] feature has a MapPoint that is not bad; NULL = all (a Frame in mode 0) */
    int n_nodes;                 /* mFeatVec.size() */
    const uint32_t *node_ids;    /* [n_nodes] ascending NodeIds */
    const int32_t *node_off;     /* [n_nodes + 1] */
    const uint32_t *node_feat;   /* [node_off[n_nodes]] feature indices, per node in insertion order */
} orc_bow_features;

int orc_descriptor_distance(const uint32_t *a, const uint32_t *b);                    /* ORBmatcher.cpp:1492-1508 */
void orc_three_maxima(const int *histo, int L, int *ind1, int *ind2, int *ind3);      /* :1445-1488 */
/* mode 0: SearchByBoW(KF, Frame) (:110-239); mode 1: SearchByBoW(KF1, KF2) (:354-487) */
int orc_search_by_bow(const orc_bow_features *q, const orc_bow_features *t, float nn_ratio, int check_orientation,
                      int mode, int32_t *match_out);

/* ------------------------------------------------ KeyFrameDatabase candidate retrieval */
/* what DetectRelocalizationCandidates / DetectLoopCandidates read from the keyframe database: keyframe index =
 * insertion order into the inverted file */
typedef struct {
    int K;
    const int64_t *bow_off;      /* [K+1] */
    const uint32_t *bow_word;    /* word ids, ascending per keyframe (DBoW2::BowVector is a std::map) */
    const double *bow_val;       /* WordValue = double */
    const int32_t *covis;        /* [K][10] KeyFrame::GetBestCovisibilityKeyFrames(10), -1 padded */
} orc_kfdb;
double orc_bow_l1_score(int n1, const uint32_t *w1, const double *v1, int n2, const uint32_t *w2, const double *v2);
double orc_kfdb_last_query_seconds(void);
int orc_detect_candidates(const orc_kfdb *db, int mode, int nq, const uint32_t *qword, const double *qval, int n_conn,
                          const int32_t *conn, float min_score, float *score_state, int32_t *out, int cap);

/* ------------------------------------------------ guided matching */
/* what SearchBySim3 / SearchByProjection read from a KeyFrame (or Frame) and its MapPoints */
typedef struct {
    int n_feat;
    const float *kp_xy;          /* [n][2] mvKeysUn[i].pt */
    const int32_t *kp_octave;    /* [n] mvKeysUn[i].octave */
    const float *kp_angle;       /* [n] mvKeysUn[i].angle (SearchByProjection's rotation histogram; may be NULL for SearchBySim3) */
    const uint32_t *desc;        /* [n][8] mDescriptors rows */
    const uint8_t *mp_valid;     /* [n] feature has a MapPoint that is not bad */
    const float *mp_xyz;         /* [n][3] MapPoint::GetWorldPos() */
    const uint32_t *mp_desc;     /* [n][8] MapPoint::GetDescriptor() */
    const float *mp_maxdist;     /* [n] mfMaxDistance */
    const float *mp_mindist;     /* [n] mfMinDistance */
    float Rcw[9], tcw[3];        /* GetRotation(), GetTranslation() */
    float bounds[4];             /* mnMinX, mnMaxX, mnMinY, mnMaxY */
    int grid_cols, grid_rows;    /* mnGridCols, mnGridRows */
    float grid_w_inv, grid_h_inv;/* mfGridElementWidthInv, mfGridElementHeightInv */
    const int32_t *grid_off;     /* [cols*rows + 1], cell (ix, iy) at ix*rows + iy */
    const int32_t *grid_idx;     /* feature indices per cell, ascending (AssignFeaturesToGrid order) */
    int n_levels;                /* mnScaleLevels */
    const float *scale_factors;  /* [n_levels] mvScaleFactors */
    float log_scale_factor;      /* mfLogScaleFactor */
} orc_kf_view;
int orc_features_in_area(const orc_kf_view *kf, float x, float y, float r, int32_t *out);
int orc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int n_levels);
/* ORBmatcher::SearchByProjection(Frame&, KeyFrame, sAlreadyFound, th, ORBdist) (ORBmatcher.cpp:1317-1444) */
int orc_search_by_projection(const orc_kf_view *frame, const orc_kf_view *kf, const float K[4], const float Rcw[9], const float tcw[3],
                             float th, int orb_dist, int check_orientation, const uint8_t *occupied, const uint8_t *already_found,
                             int32_t *frame_match);
int orc_search_by_sim3(const orc_kf_view *kf1, const orc_kf_view *kf2, const float K[4], const float R12[9], const float t12[3],
                       float s12, float th, const int32_t *matched12_in, int32_t *match12_out);

/* ------------------------------------------------ Optimizer::PoseOptimization */
/* one frame: what Optimizer.cpp:244-323 reads from the Frame and its MapPoints */
typedef struct {
    int n;                    /* keypoints with a MapPoint (nInitialCorrespondences) */
    const float *p3d;         /* [n][3] MapPoint::GetWorldPos() */
    const float *obs;         /* [n][3] kpUn.pt.x, kpUn.pt.y, mvuRight (< 0: monocular edge) */
    const float *inv_sigma2;  /* [n] mvInvLevelSigma2[kpUn.octave] */
    float K[5];               /* fx, fy, cx, cy, mbf */
    float Rcw[9], tcw[3];     /* pFrame->mTcw */
} orc_poseopt_problem;

typedef struct {
    int32_t n_inliers;        /* return value: nInitialCorrespondences - nBad */
    int32_t n_bad;
    int32_t rounds;           /* outer rounds executed */
    int32_t iterations;       /* LM iterations (solve() calls) */
    int32_t trials;           /* LM trials (linear solves + chi2 passes) */
    int32_t reserved;
    double R[9], t[3];        /* SE3quat_recov.to_homogeneous_matrix() */
    float Rf[9], tf[3];       /* Converter::toIso -> Frame::SetPose */
} orc_poseopt_result;

/* outlier: [n] pFrame->mvbOutlier of the matched keypoints */
void orc_pose_optimization(const orc_poseopt_problem *pb, orc_poseopt_result *res, uint8_t *outlier);
void orc_pose_optimization_batch(int C, const orc_poseopt_problem *pbs, orc_poseopt_result *res, uint8_t **outliers);

/* ------------------------------------------------ Optimizer::OptimizeSim3 */
typedef struct {
    int n;                        /* matches that pass Optimizer.cpp:1107-1143 (nCorrespondences) */
    const float *x1c, *x2c;       /* [n][3] P3D1c, P3D2c */
    const float *obs1, *obs2;     /* [n][2] kpUn1.pt, kpUn2.pt */
    const float *inv_sigma2_1, *inv_sigma2_2;   /* [n] */
    float K1[4], K2[4];           /* fx, fy, cx, cy */
    float S12[13];                /* g2oS12 on entry: R (9), t (3), s */
    float th2;
    int fix_scale;
} orc_sim3opt_problem;

typedef struct {
    int32_t n_inliers, n_bad, optimized, iterations, trials, reserved;
    double R[9], t[3], s;
    double q[4];
} orc_sim3opt_result;

void orc_optimize_sim3(const orc_sim3opt_problem *pb, orc_sim3opt_result *res, uint8_t *removed);

#ifdef __cplusplus
}
#endif
#endif
