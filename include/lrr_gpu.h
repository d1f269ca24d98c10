/* lrr_gpu.h — C ABI of the B200-native CDCL robust-ICP backend (liblrr_b200.so).
 *
 * Drop-in boundary for the LRR hot path.  Every entry point cites the reference
 * interface (file:line under the upstream msofka/LRR tree) it replaces.  Plain C: opaque handles,
 * plain pointers and sizes, `int` status returns (0 = ok), no exceptions, no mutable global state (per-device facts such as the
 * SM count are cached per device; contexts on different devices may be driven from different threads of one process).
 * One lrr_ctx per CUDA device; calls on one ctx are serialised by the caller — the reference
 * estimators are single-threaded too (rrl/rrl_estimation_symmetric_ICP_matching_all.txx:760).
 *
 * All feature arrays are HOST pointers in structure-of-arrays order (component-major):
 *   xyz      : 3*n doubles  x[0..n) y[0..n) z[0..n)                       cdcl_feature::location_
 *   P6       : 6*n doubles  Pxx Pxy Pxz Pyy Pyz Pzz (symmetric projector)  ::error_projector_
 *   C6       : 6*n doubles  Cxx Cxy Cxz Cyy Cyz Czz, NULL for rrl          cdcl_feature::covariance_
 *   strength : n doubles                                                   ::strength_
 *   shape    : n bytes  0=CORNER 1=TUBE 2=SHEET                            ::shape_
 *   dirs     : 9*n doubles d0x d0y d0z d1x d1y d1z d2x d2y d2z, NULL ok    ::directions_
 * (cdcl/cdcl_feature.h:17-64, cdcl/cdcl_feature_with_shape.h:23-89, cdcl/cdcl_feature_ICP.h:23-96).
 *
 * The CPU oracle (oracle/liblrr_oracle.so, test infrastructure only) exports the same functions
 * with the prefix `lrro_`, so parity tests run one script against both libraries.
 */
#ifndef LRR_GPU_H_
#define LRR_GPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lrr_ctx  lrr_ctx;   /* one CUDA device + stream + scratch                          */
typedef struct lrr_fset lrr_fset;  /* one uploaded feature set: SoA stream copy + NN grid records  */
typedef struct lrr_est  lrr_est;   /* one estimator object (forward or symmetric)                  */

/* status codes */
enum {
  LRR_OK = 0,
  LRR_ERR_INVALID = 1,      /* bad argument                                                       */
  LRR_ERR_UNSUPPORTED = 2,  /* (dim,dof,kind) the reference does not instantiate, see Q6           */
  LRR_ERR_CUDA = 3,         /* CUDA runtime failure; lrr_ctx_last_error has the text               */
  LRR_ERR_NO_DEVICE = 4,    /* no CUDA device: the product path has NO CPU fallback                */
  LRR_ERR_STATE = 5         /* call order violated (e.g. metric before any match exists)           */
};

/* estimator kinds = the four reference classes on the path + the 2-D demo estimator of config 1 */
enum {
  LRR_CDCL_FWD = 0, /* cdcl_estimation_ICP_matching_all<dim,dof>            cdcl/cdcl_estimation_ICP_matching_all.h:34  */
  LRR_CDCL_SYM = 1, /* cdcl_estimation_symmetric_ICP_matching_all<dim,dof>  cdcl/cdcl_estimation_symmetric_ICP_matching_all.h */
  LRR_RRL_FWD  = 2, /* rrl_estimation_ICP_matching_all<dim,dof>             rrl/rrl_estimation_ICP_matching_all.h       */
  LRR_RRL_SYM  = 3, /* rrl_estimation_symmetric_ICP_matching_all<dim,dof>   rrl/rrl_estimation_symmetric_ICP_matching_all.h */
  LRR_CDCL_ICP = 4  /* cdcl_estimation_ICP<2,4> (ctor #1, one resolution): normal-distance ICP over plain cdcl_feature<2> with
                       cdcl_trans_similarity2d, cdcl/cdcl_estimation_ICP.txx:352-746; selected by `-method 4` in
                       cdcl/examples/estimate.cxx:99-101 (BASELINE.json configs[0]).  All moving points in every iteration (no
                       point ramp), scale = 2 x (MUSE, then RMS of the fresh matches) (:567-584), convergence |dRMS| < 1e-4 (:424) */
};

/* transform kinds (cdcl/cdcl_trans_affine.h, cdcl/cdcl_trans_rigid3d.h, cdcl/cdcl_trans_similarity2d.h).
 * LRR_TRANS_SIMILARITY2D: y = s R x + t in the plane; at this boundary it travels as the 3x3 / 3-vector pair (A = s R in the upper
 * left 2x2 block, A22 = 1, t2 = 0, centre 0) and its covariance is 4x4 over (s cos a, s sin a, tx, ty)
 * (cdcl_trans_similarity2d.cxx:76-110). */
enum { LRR_TRANS_AFFINE = 0, LRR_TRANS_RIGID3D = 1, LRR_TRANS_SIMILARITY2D = 2 };

/* direction selector for symmetric estimators (forward_ / backward_ members) */
enum { LRR_DIR_FORWARD = 0, LRR_DIR_BACKWARD = 1 };

/* moving-set order policy: the reference calls vcl_random_shuffle(moving_) in every
 * find_closest_euclidean (cdcl/...ICP_matching_all.txx:557, rrl/...ICP_matching_all.txx:508; Q4),
 * unseeded libstdc++ rand().  The permutation is therefore an injectable input here. */
enum {
  LRR_ORDER_FIXED = 0,     /* keep the order given by lrr_est_set_order (identity by default)      */
  LRR_ORDER_RESHUFFLE = 1  /* a new keyed 4-round Feistel permutation (cycle walking) of the moving vector per find_closest call,
                              keys from SplitMix64(seed, direction, call number): every element computable independently */
};

/* scalar metrics (lrr_est_metric) */
enum {
  LRR_METRIC_RMS = 0,            /* RMS_error()      cdcl :196-230 ; rrl :211-241 (returns sum r^2, Q1) */
  LRR_METRIC_WEIGHTED = 1,       /* weighted_error() cdcl :236-268 ; rrl :247-276                     */
  LRR_METRIC_MEDIAN = 2,         /* median_error()   cdcl :159-190 ; rrl :177-205                     */
  LRR_METRIC_SHEET_ANGLES = 3,   /* sheet_angles()   cdcl :273-323 ; rrl :281-323                     */
  LRR_METRIC_TUBE_ANGLES = 4     /* tube_angles()    cdcl :328-382 ; rrl :328-375                     */
};

/* iteration-control state of a symmetric estimator (cdcl sym .h / rrl sym .h:203-224) */
typedef struct lrr_est_state {
  int64_t number_points[2];   /* number_points_ of forward_/backward_                             */
  int64_t number_matches;     /* number_matches_ (ramp 1000, x2 ...)                               */
  int32_t finest_level;       /* finest_level_                                                     */
  int32_t oscillation_count;  /* oscillation_count_                                                */
  double  error_difference;   /* error_difference_                                                 */
  double  weighted_error;     /* weighted_error_                                                   */
  double  scale[2];           /* last robust scale returned by estimate_scale_and_assign_weight    */
  int32_t scale_estimated[2]; /* scale_estimated_                                                  */
  int32_t iterations_done;    /* one_iteration calls since set_transform                           */
  double  norm_center_moving[3], norm_center_fixed[3];  /* last normalisation (A9)                 */
  double  norm_radius_moving, norm_radius_fixed;
} lrr_est_state;

/* ---- context ------------------------------------------------------------------------------- */
int  lrr_ctx_create(int device, lrr_ctx** out);
void lrr_ctx_destroy(lrr_ctx* ctx);
/* optional cudaStream_t; NULL = the context's own non-blocking stream */
int  lrr_ctx_set_stream(lrr_ctx* ctx, void* cuda_stream);
int  lrr_ctx_synchronize(lrr_ctx* ctx);
const char* lrr_ctx_last_error(lrr_ctx* ctx);
/* number of kernel launches issued by this context so far (bench `gpu_launches`) */
int64_t lrr_ctx_launch_count(lrr_ctx* ctx);

/* instrumentation (no reference counterpart; the reference prints vul_timer wall times per phase,
 * cdcl/cdcl_estimation_ICP_matching_all.txx:407-408,442-444): CUDA-event time per phase kernel.
 * phase index: 0 match, 1 norm1, 2 norm2, 3 weight, 4 collect, 5 accumulate+solve, 6 resid, 7 muse, 8 metric,
 * 9 objcov, 10 batch (one block per registration), 11 restamp, 12 match completion: staged search / sparse list, 13 match
 * completion: queued queries.  Phase 0 is a whole find_closest outside the pipelined finest level; there the certificate pass is
 * the residual pass (6).  enable = 1 resets the counters and starts, 2 resumes without a reset, 0 pauses (counters kept). */
int  lrr_ctx_set_profiling(lrr_ctx* ctx, int enable);
int  lrr_ctx_phase_times(lrr_ctx* ctx, int cap, double* ms, int64_t* count);

/* ---- feature sets ----------------------------------------------------------------------------
 * Replaces the vector<feature_sptr> arguments of the estimator constructors and build_kd_trees()
 * (cdcl/cdcl_estimation_ICP_matching_all.txx:113-135) / the Voronoi map of
 * rrl/examples/compute_distance_map.cxx:55-151: the upload also builds the exact-NN uniform grid. */
int  lrr_features_upload(lrr_ctx* ctx, int64_t n, const double* xyz, const double* P6, const double* C6,
                         const double* strength, const uint8_t* shape, const double* dirs, lrr_fset** out);
/* Behaviour switches of a context (measurement A/B and fall-backs; every default is the product path).  No reference counterpart.
 * keys: "certificates" 0/1 (neighbour certificates), "tile" 0/1 (staged search, certificate pass and sparse list of the finest level),
 * "speculative" 0 never / 1 after an iteration that kept every query / 2 in every finest-level iteration (the device refuses a lost bet),
 * "pipeline" 0/1 (trailing residual pass = certificate pass of the next find_closest), "norm_skip" 0/1, "flat_scan" 0/1,
 * "tile_sparse" -1 (by the host hint) or 0..32, "first_sparse" 0/1 (a level's first find_closest on the per-query list); batched mode: "batch_sort" 0/1 (jobs handed out longest-expected-first),
 * "batch_quantum" 0..250 (iterations a job runs before it is parked and the block takes the next unit; 0 = run to completion; default 8:
 * the workspace then holds a slot per job in flight, up to "batch_slots" (default 1024) within 1/16 of the device memory, and is kept by
 * the context).
 * The environment variables listed in README.md only set the initial values.
 * Estimators created before the call see the new value from their next call on. */
int  lrr_ctx_set_option(lrr_ctx* ctx, const char* key, int value);
/* Asynchronous upload (off by default; no reference counterpart — the reference builds its k-d trees / reads its Voronoi volumes before
 * the first estimator exists, rrl/itkLocationRegistration.cxx:2035-2047).  on = 1: lrr_features_upload returns as soon as the search grid and
 * the cell-ordered records exist; the `dirs` array — 47 % of a feature set's bytes, read by the sheet / tube angle metrics only — may still
 * be on its way over PCIe while the registration already runs.  The caller keeps the `dirs` host buffer (pinned, for a truly asynchronous
 * copy) valid until lrr_features_destroy or the next lrr_ctx_synchronize; every consumer of the directions waits for the copy on the device. */
int  lrr_ctx_set_async_upload(lrr_ctx* ctx, int on);
/* Plain cdcl_feature<dim> sets (location, strength, covariance; cdcl/cdcl_feature.h:17-64) for LRR_CDCL_ICP.  loc = dim x n SoA,
 * cov = dim(dim+1)/2 x n SoA (xx, xy, yy).  The error projector the reference re-derives at every use from the fixed feature's
 * covariance — outer product of the eigenvector of the smallest eigenvalue, cdcl/cdcl_estimation_ICP.txx:470-477,686-696 — is
 * computed once here, on the device.  dim must be 2 (the 3-D form of that estimator needs the multi-resolution subsampling of
 * cdcl/cdcl_utils.txx:640-960, which is outside the path). */
int  lrr_features_upload_cov(lrr_ctx* ctx, int dim, int64_t n, const double* loc, const double* cov, const double* strength, lrr_fset** out);
void lrr_features_destroy(lrr_fset* fs);
int64_t lrr_features_size(const lrr_fset* fs);
/* exact Euclidean 1-NN of n_query host points (SoA xyz) among the set; lowest index wins ties.
 * Replaces rsdl_kd_tree::n_nearest(n=1) (cdcl :572) and the Voronoi lookup (rrl :530-543, Q5). */
int  lrr_features_nearest(lrr_fset* fs, int64_t n_query, const double* q_xyz, int32_t* idx_out);

/* FeaturesInRegions (rrl/itkLocationRegistration.cxx:1363-1522) for one volume, on the device.  labels[i] = region label of feature i:
 * the caller's watershed over-segmentation looked up at the feature's location (:1407).  From the features inside the cube
 * [x0, x1], sorted by distance from reference_point (:1297-1311; ties: lowest index), regions are taken in the order of their
 * nearest feature and all features of a region are added, region after region, until at least min_count (1500, :1487) are in
 * the set.  idx_out (capacity cap) receives the indices in the reference's order (region after region, nearest first inside a
 * region), n_out their number.  The complement of FeaturesInROIs, which lrr_est_set_subsets / lrr_batch_run already cover. */
int  lrr_features_set_labels(lrr_fset* fs, const int32_t* labels);
int  lrr_features_in_regions(lrr_fset* fs, const double* x0, const double* x1, const double* reference_point, int64_t min_count,
                             int64_t cap, int32_t* idx_out, int64_t* n_out);

/* ---- estimators ------------------------------------------------------------------------------ */
/* ctor #3 semantics (cdcl sym .txx:144-164, rrl sym .txx:198-220): forward_ = (moving_inside -> fixed_all),
 * backward_ = (fixed_inside -> moving_all).  Until lrr_est_set_subsets is called the inside sets are
 * the whole sets.  Forward kinds ignore the backward half.  dim must be 3 with dof 12 (affine) or 6
 * (rigid3d, cdcl kinds only — Q6), or — LRR_CDCL_ICP only — dim 2 with dof 4 over sets from lrr_features_upload_cov. */
int  lrr_est_create(lrr_ctx* ctx, int kind, int dim, int dof, lrr_fset* moving_all, lrr_fset* fixed_all, lrr_est** out);
void lrr_est_destroy(lrr_est* est);
/* set_features(moving_inside, fixed_inside)  (cdcl sym :206-226, rrl sym :269-289; forward: set_moving_data).
 * Index lists into the *_all sets (the driver's deep clones, rrl/itkLocationRegistration.cxx:1346,1357). */
int  lrr_est_set_subsets(lrr_est* est, const int32_t* moving_inside, int64_t nm, const int32_t* fixed_inside, int64_t nf);
/* set_transform(trans): forward <- trans, backward <- trans->inverse(), scale_estimated_=false,
 * number_points_=1000, ramp/oscillation state reset (cdcl :141-153, sym :231-242).  cov may be NULL (zeros). */
int  lrr_est_set_transform(lrr_est* est, int trans_kind, const double* A9, const double* t3, const double* c3, const double* cov);
/* normalize_matches() / weight_by_strength(bool)  (cdcl .h:92,117) */
int  lrr_est_set_flags(lrr_est* est, int normalize_matches, int weight_by_strength);
/* injection points for the unpinned third-party pieces */
int  lrr_est_set_order(lrr_est* est, int dir, const int32_t* perm, int64_t n);      /* Q4 */
int  lrr_est_set_shuffle(lrr_est* est, int mode, uint64_t seed);                    /* Q4 */
int  lrr_est_set_initial_scale(lrr_est* est, int dir, double scale /* NaN => MUSE */);
/* set_number_points (cdcl :95-108; protected in the reference, reached through the symmetric friend) */
int  lrr_est_set_number_points(lrr_est* est, int dir, int64_t number_points, int* at_max);
/* find_closest_euclidean()  (cdcl :541-654, rrl :494-571, sym wrappers) */
int  lrr_est_find_closest(lrr_est* est);
/* one_iteration(caller=0, display=0, iteration)  (cdcl :434-509, cdcl sym :268-363, rrl :427-462, rrl sym :341-438) */
int  lrr_est_one_iteration(lrr_est* est, unsigned iteration, int* converged);
/* run(): while(!converged && it<max_iter) one_iteration  (cdcl :413-428; driver loop rrl/itkLocationRegistration.cxx:2418) */
int  lrr_est_run(lrr_est* est, int max_iter, int* converged, int* iterations);
/* get_transform() / get_transform_backward(): A, t, center_moving_, covariance (dof x dof row-major) */
int  lrr_est_get_transform(lrr_est* est, int dir, int* trans_kind, double* A9, double* t3, double* c3, double* cov);
/* current matches_ of a direction: from = index into the inside set's parent (*_all), to = index into the
 * other *_all set, w = weight.  n receives the match count; arrays may be NULL to query n only. */
int  lrr_est_get_matches(lrr_est* est, int dir, int64_t cap, int64_t* n, int32_t* from_idx, int32_t* to_idx, double* w);
int  lrr_est_metric(lrr_est* est, int which, int dir, double* out);
/* estimate_LS_backward(estimate_parameters)  (cdcl sym :705-863, rrl sym :922-934) */
int  lrr_est_estimate_ls_backward(lrr_est* est, int estimate_parameters);
/* compute_transfer_error_covariance(): covarJ = mean_i J_i Sigma J_i^T  (cdcl :388-404, rrl :381-397) */
int  lrr_est_transfer_cov(lrr_est* est, int dir, double* S9);
int  lrr_est_get_state(lrr_est* est, lrr_est_state* out);
/* instrumentation of the neighbour search (no reference counterpart): per direction, the queries the last certificate pass could
 * not keep and the queries finished by the one-warp-per-query pass in the last find_closest; order-statistic windows missed and
 * steady-state bets lost on this estimator's context so far. */
typedef struct lrr_est_stats {
  int64_t cert_failed[2], passb_queries[2], work_list_entries[2];
  int64_t window_misses, spec_misses;
} lrr_est_stats;
int  lrr_est_get_stats(lrr_est* est, lrr_est_stats* out);
/* oscillated()  (rrl sym .h:156) */
int  lrr_est_oscillated(lrr_est* est, int* out);

/* ---- covariance objective (cdcl_obj_fun_cov<3,dof>::compute, cdcl/cdcl_obj_fun_cov.txx:44-134) --------
 * x = vec(upper Cholesky factor of Sigma_theta), length dof(dof+1)/2; f and g (same length) may be NULL.
 * Evaluated over the current matches/weights of `dir` in the current (unnormalised) coordinates. */
int  lrr_objcov_eval(lrr_est* est, int dir, const double* x, double* f, double* g);
/* estimate_covariance(): normalise, x0 = chol(Sigma), L-BFGS (<= max_evals evaluations, stop once f dropped
 * by more than 5 %: cdcl/cdcl_lbfgs.h:30-36), Sigma <- U^T U, unnormalise (cdcl/cdcl_estimation.txx:264-337). */
int  lrr_objcov_refine(lrr_est* est, int dir, int max_evals, double* f_start, double* f_end, int* n_evals);

/* ---- parameter objective of the CDC M-step (cdcl_obj_fun_par<3,12>::compute, cdcl/cdcl_obj_fun_par.txx:46-144) ------------------------
 * theta = (A row-major, t) (cdcl_trans_affine.txx:65-106); f and g (12 entries) may be NULL.  Evaluated over the current matches /
 * weights of `dir`; the transfer-error covariances J Sigma J^T are frozen at the direction's current transform and covariance
 * (the objective's constructor, :27-38); the transform itself is left unchanged. */
int  lrr_objpar_eval(lrr_est* est, int dir, const double* theta, double* f, double* g);
/* estimate_parameters(): normalise, L-BFGS over theta (<= max_evals evaluations, 5 % early stop), set_parameterization, unnormalise;
 * *converged = (f_start - f_end) < 1e-4 (cdcl/cdcl_estimation.txx:207-262). */
int  lrr_objpar_refine(lrr_est* est, int dir, int max_evals, double* f_start, double* f_end, int* n_evals, int* converged);

/* ---- E-step of the full CDC estimator: covariance-driven radius matching
 * (cdcl_estimation<3,12>::compute_weights, cdcl/cdcl_estimation.txx:342-545) -------------------------------------------------------------
 * For every from-feature p of `dir` under the direction's current transform and covariance: Cij_p = Jth Sigma Jth^T + Jp Cp Jp^T,
 * search radius eval_multiplier (5.0) * sqrt(lambda_max(Cij_p)), every to-feature within the radius (rsdl points_in_radius) but at most
 * max_matches (100) of them — the reference takes the first 100 of a random shuffle (:396-398; Q4, unpinned: the backend takes the
 * lowest original indices) —, w0 = weight(e^T (Cij_p + Cq)^-1 e) kept when > 0, finally w = w0^2 / (row sum * to-feature sum).
 * The many-to-many set then REPLACES the 1-NN matches as the set lrr_objcov_* and lrr_objpar_* run over (f over entries,
 * cdcl_obj_fun_cov.txx:124; over rows, cdcl_obj_fun_par.txx:136; normalisation as cdcl_estimation_ICP.txx:751-832) until the next
 * find_closest / one_iteration.  entries == 0: the covariance is too small for any match, the caller falls back to
 * find_closest_euclidean as :544 does. */
typedef struct lrr_cdc_summary {
  int64_t rows;              /* matches_.size(): from-features with at least one kept to-feature                     */
  int64_t entries;           /* sum of to_.size()                                                                    */
  int32_t min_matches, max_matches;   /* smallest / largest row (:465-466)                                          */
  double  avg_largest_eval;  /* sum over ALL from-features of sqrt(lambda_max(Jth Sigma Jth^T)) / rows (:375, :480)  */
} lrr_cdc_summary;
int  lrr_cdc_compute_weights(lrr_est* est, int dir, int max_matches, double eval_multiplier, lrr_cdc_summary* out);
/* CSR export for display / tests: rows ascending by from position (row_from = index in the parent set), entries of a row ascending by
 * to index.  Sizes through n_rows / n_entries (the arrays may be NULL to query them). */
int  lrr_cdc_get_matches(lrr_est* est, int dir, int64_t cap_rows, int64_t cap_entries, int64_t* n_rows, int64_t* n_entries,
                         int32_t* row_from, int64_t* row_start, int32_t* to_idx, double* w);
/* the resolution switch compute_weights ends with (:494-508): one level finer when avg_largest_eval < 2 spacing[res], else one level
 * coarser when it exceeds 2 spacing[res + 1]; the caller re-invokes set_subsets with that level's feature sets and matches again. */
int  lrr_cdc_resolution_switch(double avg_largest_eval, const double* fixed_spacing, int res, int max_res, int allowed, int* new_res);
/* cdcl_estimation<3,12>::one_iteration (:88-186) at one resolution: weights; for iteration > 5 estimate_parameters and weights again;
 * estimate_covariance.  *converged = parameters && covariance && |Sigma_theta|_F < 1e-3. */
int  lrr_cdc_one_iteration(lrr_est* est, int dir, unsigned iteration, int max_matches, double eval_multiplier, int* converged, lrr_cdc_summary* out);

/* ---- batched location registration (outer loops of LocationRegistration::Run,
 * rrl/itkLocationRegistration.cxx:2138,2230,2332) ------------------------------------- */
typedef struct lrr_job {
  double  A[9], t[3], c[3];      /* initial transform (affine) of this candidate                    */
  double  roi_center_moving[3];  /* ROI cube centre in the moving volume                            */
  double  roi_center_fixed[3];   /* ROI cube centre in the fixed volume                             */
  double  roi_half_size;         /* half edge of the cube (25 mm = the driver's 50 mm ROI, :2275)   */
} lrr_job;

typedef struct lrr_result {
  int32_t status;        /* 0 ok, 1 rejected (<1000 features in an ROI, :1632), 2 invalid transform  */
  int32_t converged, oscillated, iterations;
  int32_t n_moving_inside, n_fixed_inside;
  double  A[9], t[3], c[3];
  double  cov[144];
  double  rms, weighted_error;
} lrr_result;

/* Run njobs independent rrl symmetric <3,12> registrations (set_features + set_transform + <=max_iter
 * one_iteration calls each) over one resident volume pair. */
int  lrr_batch_run(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t njobs,
                   int max_iter, uint64_t shuffle_seed, lrr_result* out);
/* The same over a SHARD of a larger list: job_ids[i] is the index of jobs[i] in the whole list.  Job i is seeded with
 * shuffle_seed + job_ids[i], so its result does not depend on how the list was sharded (one process per GPU: block-cyclic shards,
 * the 32 candidates of a location — rrl/itkLocationRegistration.cxx:2230 — spread over all devices). */
int  lrr_batch_run_ids(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, const int64_t* job_ids,
                       int64_t njobs, int max_iter, uint64_t shuffle_seed, lrr_result* out);
/* The whole list over ndev devices of ONE process (the reference's outer loops, :2138 / :2230 / :2332, are one sequential program):
 * ctx[d], moving_all[d], fixed_all[d] = replicas of the same volume pair on device d.  One host thread per device, one launch per device
 * over its share of the list (dealt from the jobs sorted by expected cost, so every device gets the same mix; the kernel schedules
 * inside a launch), results written in place; job j is seeded with shuffle_seed + j whichever device runs it. */
int  lrr_batch_run_multi(lrr_ctx** ctx, int ndev, lrr_fset** moving_all, lrr_fset** fixed_all, const lrr_job* jobs, int64_t njobs,
                         int max_iter, uint64_t shuffle_seed, lrr_result* out);

/* Expected cost of every job before it runs: the features covered by its two ROI cubes (whole grid cells), 0 when a cube cannot
 * reach the 1000-feature gate (:1632).  The batched entry points above hand their jobs out longest-expected-first by this figure
 * (a job's time is iterations x ROI points; the makespan is set by the jobs still running when the queue is empty; switch:
 * lrr_ctx_set_option "batch_sort"); a host that shards the list over processes balances the shares with it (lrr_b200/sharding.py). */
int  lrr_batch_expected_cost(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t njobs, int32_t* cost);

/* ---- the driver's candidate loop (rrl/itkLocationRegistration.cxx:2230-2332, 1816-1871) -----------------------------------------
 * jobs[l * n_candidates + c] = candidate initialisation c of nodule location l.  For every candidate: SetupEstimation,
 * find_closest_euclidean and the four angle measures -> ranking key = their mean (:2303-2326).  Then per location, best angle
 * first: SetupEstimation, the <= max_iter one_iteration loop (:2418), and — when it converged and did not oscillate (:2495-2507) —
 * ComputeDecisionMeasurements (:1816-1871: angles, weighted errors, estimate_LS_backward(false), the maximum trace of the transfer-
 * error covariance J Sigma J^T over a 5 mm grid of the ROI, forward and backward, :2793-2858).  until_verified != 0: a location
 * stops at its first verified candidate (:2332) — verified = converged, not oscillated, validTransform; the reference's libsvm
 * model file is not in its repository — else every candidate is run.  All on the device, one launch for the ranking and one for
 * the registrations; out / decisions are indexed like jobs. */
typedef struct lrr_decision {
  double  rank_angles[4];       /* sheet, tube, sheet backward, tube backward (degrees) at the initial transform   :2305-2311 */
  double  rank_angle_average;   /* the ranking key                                                                 :2313      */
  double  measurements[8];      /* max / min of {weighted error, maxTrace, sheet angles, tube angles}              :1853-1860 */
  double  weighted_error[2], max_trace[2], sheet_angles[2], tube_angles[2];   /* [0] forward, [1] backward, before max / min */
  int32_t rank;                 /* position among the location's candidates, best angle first; -1: ROI rejected    :2326      */
  int32_t ran;                  /* the registration was run (until_verified: candidates after the verified one are not)       */
  int32_t verified;             /* converged, not oscillated, validTransform                                                  */
  int32_t measured;             /* measurements[] are valid (the driver computes them only for converged, non-oscillating runs) */
} lrr_decision;
int  lrr_batch_run_locations(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t n_locations,
                             int32_t n_candidates, int max_iter, uint64_t shuffle_seed, int until_verified, double voxel_spacing,
                             lrr_result* out, lrr_decision* decisions);

#ifdef __cplusplus
}
#endif
#endif /* LRR_GPU_H_ */
