// Host-side objects behind the opaque handles of include/lrr_gpu.h.
#pragma once
#include <string>
#include <vector>
#include <cstdint>
#include "../../include/lrr_gpu.h"
#include "lrr_internal.h"
#include "lrr_linalg.cuh"

#define LRR_NPHASE 14
enum { PH_MATCH = 0, PH_NORM1, PH_NORM2, PH_WEIGHT, PH_COLLECT, PH_ACCUM, PH_RESID, PH_MUSE, PH_METRIC, PH_OBJCOV, PH_BATCH, PH_RESTAMP, PH_SEARCH, PH_QUEUE };
struct lrr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  // uploads: everything but the locations is copied on a second stream while the grid is built from the locations
  cudaStream_t copy_stream = nullptr; cudaEvent_t copy_ev = nullptr;
  std::string err;
  long long launches = 0;
  double* partials = nullptr;
  unsigned int* ticket = nullptr;
  unsigned int* sel_hist = nullptr; double* sel_cand = nullptr; unsigned long long* sel_misc = nullptr;
  SelDev sel[4];
  // behaviour switches (lrr_ctx_set_option; the environment variables of README.md only set their initial values)
  struct Options { int certificates = 1, tile = 1, spec = 1 /* 0 never, 1 when the last iteration kept every query, 2 always */, pipe = 1, norm_skip = 1,
                       tile_sparse = -1 /* -1: by the host hint */, flat = 1, first_sparse = 1 /* a level's first find_closest also goes to the per-query list */,
                       batch_sort = 1 /* batched mode: jobs handed out longest-expected-first */,
                       batch_quantum = 8 /* batched mode: iterations a job runs before it is parked for the next unit (0: run to completion) */,
                       batch_slots = 1024 /* ... and the most workspace slots (jobs in flight) it may use */; } opt;
  bool async_upload = false; cudaStream_t dirs_stream = nullptr;
  unsigned long long sel_owner = 0; long long window_misses = 0;   // signature of the last estimate that left valid search windows in the selectors
  SortScratch sort; double* sort_keys = nullptr;
  double* muse_table = nullptr;
  double* scratch_d = nullptr;
  void* pinned = nullptr;
  double* mailbox = nullptr; double* mailbox_dev = nullptr; unsigned long long mailbox_seq = 0; long long spec_misses = 0;   // result mailbox (mapped pinned, lrr_icp.cu)
  // workspaces of the batched mode kept between calls (up to two sizes: a warm-up and a full-size call alternate in practice): with the
  // slot pool of the time-sliced kernel a workspace is gigabytes, and a fresh allocation of that size from the driver costs tens of ms
  struct Ws { void* p; size_t bytes; }; std::vector<Ws> batch_ws; size_t batch_budget = 0;
  // instrumentation: CUDA-event timing per phase kernel (lrr_ctx_set_profiling)
  bool profiling = false;
  struct Span { int phase; cudaEvent_t a, b; };
  std::vector<Span> spans;
  double phase_ms[LRR_NPHASE] = {0}; long long phase_count[LRR_NPHASE] = {0};
};

struct lrr_fset {
  lrr_ctx* ctx = nullptr;
  int64_t n = 0;
  int dim = 3;   // 2: plain planar cdcl_feature<2> set (lrr_features_upload_cov), stored with z = 0
  double *xyz = nullptr, *s = nullptr, *P6 = nullptr, *C6 = nullptr, *dirs = nullptr;
  uint8_t* shape = nullptr;
  double4* pos4 = nullptr; double* rec = nullptr; double* crec = nullptr; int* sorted_of_orig = nullptr; int* cell_start = nullptr;
  FsetDev dev;
  // asynchronous upload (lrr_ctx_set_async_upload): the direction arrays — half the bytes, read by the angle metrics only — may still be on
  // their way when lrr_features_upload returns; their consumers wait for this event on their stream (fset_wait_dirs)
  cudaEvent_t dirs_ev = nullptr; bool dirs_pending = false;
  int* labels = nullptr;          // n region labels (watershed over-segmentation at the feature locations) for lrr_features_in_regions, or null
  std::vector<double> host_xyz;   // 3 x n copy for host-side ROI selection (FeaturesInROIs); fetched from the device on first use
};

// one direction (forward_ / backward_) of an estimator
struct lrr_dir {
  lrr_fset *from_all = nullptr, *to_all = nullptr;
  size_t cap = 0;
  double *fx = nullptr, *fy = nullptr, *fz = nullptr, *fs = nullptr, *fP6 = nullptr;
  int *forig = nullptr, *order = nullptr, *sidx = nullptr;
  double *w = nullptr, *r2 = nullptr, *gkey = nullptr, *skey = nullptr;
  unsigned long long* stamp = nullptr;
  int *hq_k = nullptr, *hq_s = nullptr, *hq_o = nullptr, *hq_count = nullptr, *prev_s = nullptr, *proc = nullptr, *proc_mark = nullptr; double* hq_best = nullptr; double4* proof = nullptr; int* pos_of = nullptr; int* cprev = nullptr; double* tf = nullptr; int2* todo = nullptr; int* todo_count = nullptr; FqRec* fq = nullptr;
  int *tile_cnt = nullptr, *tile_start = nullptr, *chunk_start = nullptr, *tile_of = nullptr; int2* chunk_tab = nullptr; int ntiles = 0; TileGeom tgeom;   // lrr_tile.cu
  Scal* sc = nullptr;
  int64_t n_in = 0, number_points = 0, matches = 0;
  unsigned int epoch = 0;
  bool subset = false, has_order = false, scale_estimated = false, weights_fresh = false;
  double initial_scale = 0;
  uint64_t shuffle_calls = 0;        // RESHUFFLE: find_closest calls so far (selects the permutation key)
  unsigned int perm_key[4] = {0, 0, 0, 0}; int perm_half = 0;
  // many-to-many match set of the CDC E-step (lrr_cdc_compute_weights, lrr_cdc.cu): flat entries (compact from index, neighbour slot,
  // weight), per-from-feature rows, per-to-feature weight sums.  While cdc_active the objective entry points (lrr_objcov_*,
  // lrr_objpar_*) run over this set instead of the 1-NN matches; any new 1-NN match drops it.
  int *cdc_efrom = nullptr, *cdc_eslot = nullptr, *cdc_rowcnt = nullptr, *cdc_rowstart = nullptr, *cdc_istat = nullptr;
  double *cdc_ew = nullptr, *cdc_roww = nullptr, *cdc_tow = nullptr, *cdc_dsum = nullptr, *cdc_candw = nullptr;
  int2* cdc_cand = nullptr; unsigned long long* cdc_cursor = nullptr;
  int64_t cdc_ecap = 0, cdc_entries = 0, cdc_rows = 0; size_t cdc_scratch = 0; bool cdc_active = false;
};

struct lrr_est {
  lrr_ctx* ctx = nullptr;
  int kind = 0, dof = 12, dim = 3;
  lrr_fset *moving_all = nullptr, *fixed_all = nullptr;
  lrr_dir d[2];
  Scal h[2];                 // host mirrors (head part is kept current)
  bool normalize = false, strength = false;
  int shuffle_mode = 0; uint64_t seed = 0;
  // symmetric iteration control (cdcl sym .h / rrl sym .h:203-224)
  unsigned oscillation_count = 0; double error_difference = 0, weighted_error = 0;
  unsigned number_matches = 1000; int finest_level = 0; int iterations_done = 0;
  bool werr_cache_valid = false; double werr_cache = 0;
  int window_holdoff = 0, window_backoff = 0;   // order-statistic window: estimates to skip after a miss
  // the normalisation scalars on the device are those of this estimator's last forward-source estimate_LS (nothing else wrote
  // them since): k_norm1 / k_norm2 may leave at entry while the device says no match changed (Scal::norm_dirty)
  bool norm_token = false;
  // every query of every direction was kept by its certificate in the last iteration (Scal::match_left == 0): the next iteration is
  // launched without the staged search, pass B and the normalisation (speculative plan; verified on the device, see estimate_and_pull)
  bool steady_hint = false;
  // under a quarter of the queries lost their certificate in the last certificate pass: the next one routes every failed query to the
  // per-query sparse list instead of staging whole chunks (est_args, lrr_capi.cu)
  bool sparse_hint = false;
  // the certificate pass of the NEXT find_closest has already run, as the trailing (residual) pass of the last one_iteration (lrr_tile.cu)
  bool pre_matched = false;
  bool sym() const { return kind == LRR_CDCL_SYM || kind == LRR_RRL_SYM; }
  bool icp() const { return kind == LRR_CDCL_ICP; }   // cdcl_estimation_ICP<2,4>: no point ramp, scale constant 2, strength by sum
  bool rrl() const { return kind == LRR_RRL_FWD || kind == LRR_RRL_SYM; }
};

// kernels with host-visible extras (lrr_icp.cu / lrr_metrics.cu)
void lrr_k_weight_ex(const EstArgs& a, const int* scale_mode, const double* scale_in, cudaStream_t st, long long* launches, int spec = 0);

// extra arguments of the block-per-registration launches issued by lrr_batch_run_locations (lrr_batch.cu)
struct BatchLocArgs { int mode, ncand, until_verified; const int* order; lrr_decision* dec; double spacing; };
extern thread_local BatchLocArgs lrr_batch_loc_args;
