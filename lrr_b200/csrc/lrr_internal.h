// Internal declarations shared by the CUDA translation units of liblrr_b200.so (sm_100a only).
//
// Data layout in HBM:
//   * every uploaded feature set keeps (a) its caller-order structure-of-arrays (fp64 x,y,z,strength,P6,C6,
//     dirs + u8 shape) and (b) a copy sorted by uniform-grid cell: `pos4` = one 32-byte record per feature
//     {x,y,z,original index} — the only array the exact 1-NN search touches, one sector per candidate —
//     and `rec` = one 64-byte record {Pxx,Pxy,Pxz,Pyy,Pyz,Pzz,strength,strength as fp32} gathered once per
//     match.  x-adjacent cells are contiguous in the sorted order.
//   * the "from" side of a direction (the streamed side) is a compact SoA in subset order
//     (x,y,z,strength,P6): coalesced fp64 loads, gathered once per set_features.
//   * per-match arrays (sorted slot of the neighbour, weight, residual^2, selection keys) are plain SoA.
//   * all scalars of a direction (transform, normalisation, reduction results) live in one `Scal` struct
//     in device memory; the host reads it back once per one_iteration call.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LRR_MAX_BLOCKS 1184   // 148 SMs x 8
#define LRR_NRED 128          // max doubles reduced per block in one pass (108 moments + 12 sums of the fused estimate pass)
#define LRR_SEL_FINE 65536    // fine bins of the order-statistic histogram
#define LRR_SEL_COARSE 1024   // coarse bins (fine >> 6)
#define LRR_SEL_CAP 2048      // candidate capacity of the final exact selection
#define LRR_STRENGTH_CAP 90000.0

// spatial tiles of the from-features (lrr_tile.cu): cubes of LRR_TILE_T cells of the to-grid; the finest-level match stages
// the cells of one tile's queries plus a halo of LRR_TILE_H cells in shared memory
#define LRR_TILE_T 8
#define LRR_TILE_H 1
struct TileGeom { double ox, oy, oz, inv_edge; int ntx, nty, ntz; };

struct GridDev {
  double ox, oy, oz, h, inv_h;
  int nx, ny, nz, n;
  const int* cell_start;    // ncell+1
  const double4* pos4;      // n, cell order: x,y,z,(original index as int64 bits)
};

// device view of one uploaded feature set
struct FsetDev {
  int n;
  const double *x, *y, *z, *s;   // caller order
  const double* P6;              // 6 x n
  const double* C6;              // 6 x n or null
  const double* dirs;            // 9 x n or null
  const uint8_t* shape;          // n
  GridDev grid;
  const double* rec;             // 8 x n interleaved, cell order
  const double* crec;            // 8 x n interleaved, cell order: Cxx Cxy Cxz Cyy Cyz Czz 0 0 (null without covariances): one 64-byte gather per match for cdcl_obj_fun_cov
  const int* sorted_of_orig;     // n
};

// one exact order-statistic selector (MAD scale of Q12) = histogram + candidate list
struct SelDev {
  unsigned int* fine;      // LRR_SEL_FINE
  unsigned int* coarse;    // LRR_SEL_COARSE
  double* cand;            // LRR_SEL_CAP
  unsigned int* cand_count;
  unsigned long long* cand_minmax;   // [0]=min bits, [1]=max bits of the candidates
  // written by the finalizer of the weight phase
  int* bin;                // fine bin holding the k-th value
  unsigned int* rank;      // rank of the k-th value inside that bin
  double* count;           // number of keys
  double* result;          // the MAD scale (0 when count <= 1)
  double inv_width;        // key -> fine bin = min(FINE-1, (int)(key * inv_width))
  double* bracket;         // [0] lo, [1] hi, [2] relative half width, [3] last order statistic: search window of the next call
};

// Device-resident scalars of one direction (forward_ or backward_ of an estimator).
struct Scal {
  // current transform of this direction, unnormalised: y = A (x - c) + t
  double A[9], t[3], c[3];
  double cov[144];
  // previous global transform while the matches' from_ are mapped clones (rigid increment, cdcl :455-484)
  double Ag[9], tg[3], cg[3];
  int from_mapped, status;
  // normalisation (A9) of this direction's from / to features
  double cfrom[3], rfrom, cto[3], rto;
  // reduction results
  double sum_loc[3], sum_r2;                 // match phase
  double sum_q[3], n_unique, sum_rad_from;   // norm1
  double sum_rad_to;                         // norm2
  double scale; int scale_estimated;
  // 1 = some match of this direction was (re)written since its normalisation was last computed.  Set by the finalizers of the
  // match kernels (the certificate pass only when it left work over), cleared by k_norm2: while every query of the source
  // direction is kept by its certificate the match set — all the normalisation depends on for an affine estimate — is unchanged,
  // and k_norm1 / k_norm2 leave at entry (see do_estimate, lrr_capi.cu).
  int norm_dirty;
  // queries of this direction's last find_closest that the certificate pass could NOT keep (work list + queue; any non-zero value
  // when the search did not go through the certificate pass).  0 in every direction and norm_dirty == 0 = the steady state: the
  // staged search, pass B and the normalisation have nothing to do — the host may leave their launches out (speculative plan,
  // lrr_capi.cu), and the kernels after them verify exactly this before they commit anything.
  int match_left, spec_pad;
  double n_nonzero, sel_below[2];
  double G, S;
  double res_sum_r2, res_sum_wr2, res_sum_w, res_sum_wr, res_count;   // resid phase
  // PRE-MATCH (lrr_tile.cu, k_match_cert in its trailing role): the residual pass that ends an iteration at the finest level is at
  // the same time the certificate pass of the NEXT find_closest (same x, q, P, same e^T P e under the same transform).  When it
  // keeps every query it writes the next match's sums into sum_loc / sum_r2 at once; otherwise those stay what they are (the
  // matches of the finished iteration remain intact for any call in between) and the kept part waits here until the staged search
  // completes the match (k_match_tile's finalizer folds it in).
  double pm_sum_loc[3], pm_sum_r2;
  double used_sum_r2;   // sum_r2 of the match the last estimate was made from (RMS_error "before", cdcl :440): copied by the weight phase
  // from_mapped == 2: the matches' from_ are clones of clones — mapped through (Ag, tg, cg) and then through this second transform
  // (cdcl symmetric rigid estimate_LS_backward maps the already mapped clones once more, cdcl sym :708-731)
  double Ag2[9], tg2[3], cg2[3];
  double cert_failed, passb_queries;   // instrumentation: queries the last certificate pass could not keep / queries finished by pass B in the last find_closest
  double wfac;          // fused estimate pass: the weights it stored are final up to this factor 1 / (G S), applied by the trailing pass
  double mom[LRR_NRED];
  double AtA[144], Atb[12];
  double ang_num, ang_den;
  double covJ[10];   // scratch: second moments of the Jacobian locations
  double obj_f, obj_kl[144];
};

enum {
  LRR_ST_OK = 0,
  LRR_ST_SELECT_OVERFLOW = 1,   // candidate list overflow in the exact selection (host falls back to sort)
  LRR_ST_SINGULAR = 2,
  LRR_ST_SELECT_MISS = 4,       // the order statistic fell outside the predicted window: the host reruns the estimate with histograms
  LRR_ST_NOT_STEADY = 8         // speculative plan refused: some query was not kept / the normalisation is stale; the host completes the match and reruns
};

// one query the certificate pass could not keep, with everything it had already loaded for it (one hop instead of four)
struct __align__(16) FqRec { int k, cand, cand_o, tq; double x, y, z, best; };

struct DirArgs {
  // from side (compact, subset order)
  const double *fx, *fy, *fz, *fs;
  const double* fP6;     // 6 x n_in
  const int* forig;      // n_in: index in the parent *_all set
  int n_in;
  const int* order;      // n_in or null (identity): position k -> compact index
  unsigned int perm_key[4]; int perm_half;   // RESHUFFLE: keyed Feistel permutation of [0,n_in) (perm_half = 0: off)
  const int* proc;       // n_in or null: match positions in spatially coherent (tile) processing order (only when M == n_in)
  const int2* chunk_tab; // or null: chunk c = positions [chunk_tab[c].x, chunk_tab[c].y) of proc, at most 32 queries of one row of tiles (lrr_tile.cu)
  const int* nchunks;    // device scalar: number of chunks
  int M;                 // number_points_: matches in use
  int ball_rows;         // pass A keeps a query whose candidate ball touches at most this many grid rows (0: LRR_BALL_ROWS), the rest goes to pass B
  // to side
  FsetDev to;
  FsetDev from_all;      // parent set of the from side (rarely used attributes are gathered through forig)
  // per-match arrays (length n_in)
  int* sidx;
  double *w, *r2, *gkey, *skey;
  unsigned long long* stamp;   // n_to: (epoch, match position) of the representative match of a to-feature
  unsigned long long tag;      // epoch tag of the current matches (high bits)
  // 1: every finished match stamps its to-feature at once (ramp levels, batched mode).  0: the certified finest level — kept matches
  // are not touched at all there, so the stamps are laid down by k_restamp right before the normalisation reads them (and only when
  // the match set changed).
  int stamp_in_match;
  // queue of the queries the 27-cell pass could not finish (pass B, one warp per query)
  int* hq_k; double* hq_best; int* hq_s; int* hq_o; int* hq_count;
  int* prev_s;                 // n_in: neighbour slot of each from-feature in its last match (-1: none) = search candidate
  // n_in or null: NEIGHBOUR CERTIFICATE of each from-feature = {mapped position y0 at which prev_s was proven to be the unique
  // nearest neighbour, lower bound B on the distance from y0 to every OTHER to-feature} (w <= 0: none).  For a new mapped position
  // y the neighbour is unchanged, without any search, when |y - q_prev| + |y - y0| < B (triangle inequality; lrr_tile.cu).
  double4* proof;
  // The certificate pass reads everything of a query by its TILE-ORDER position tq (coalesced, no hop through proc): `proof` is
  // indexed by tq = pos_of[from-feature], tfx/tfy/tfz are the from-locations in tile order and cprev[tq] mirrors prev_s of the
  // feature at tq (both written by lrr_match_finish only).  null: certificates off.
  const int* pos_of; const double *tfx, *tfy, *tfz; int* cprev;
  // work list of the staged search when certificates are in use: (chunk, lane mask of the queries the certificate pass could not
  // keep), appended by k_match_cert, drained and re-armed by k_match_tile
  int2* todo; int* todo_count;
  // ... and, for chunks with only a few such queries, the queries themselves (tile-order positions): searched by groups of eight lanes
  struct FqRec* fq; int* fq_count;
  Scal* sc;
};

enum {
  LRR_F_RRL = 1,          // rrl flavour: fp32 strength product, rrl RMS (Q1), w<1e-4 skip in accumulation
  LRR_F_RIGID = 2,        // rigid3d increment: from = mapped clone, J = [-[x]x | I], b = q - p
  LRR_F_NORMALIZE = 4,    // normalize_matches_
  LRR_F_SYM = 8,          // symmetric estimator (scale before normalisation, spatial weighting)
  LRR_F_STRENGTH = 16,    // weight_by_strength_
  LRR_F_MERGED = 32,      // cdcl symmetric: backward matches join the forward list reversed (one G, one S)
  LRR_F_USE_OLD_W = 64,   // scale from the current weights (estimate_LS_backward after estimate_LS)
  LRR_F_NOSPATIAL = 1024, // cdcl estimate_LS_backward: merged list without weight_spatially (cdcl sym :744-763)
  LRR_F_NO_OTHER_RENORM = 2048,   // ... and only the estimated direction's transform goes through normalize / unnormalize
  LRR_F_ICP = 4096,       // cdcl_estimation_ICP: scale constant 2.0 (cdcl_estimation_ICP.txx:580), strength weights normalised by their SUM (:607-637)
  LRR_F_SIM2D = 8192,     // cdcl_trans_similarity2d<2,4>: J = [x -y 1 0; y x 0 1], parameters (s cos a, s sin a, tx, ty)
  LRR_F_INV_OF_INCREMENT = 32768,   // cdcl symmetric rigid estimate_LS_backward(true): the other direction becomes the inverse of the INCREMENT (cdcl sym :849-851 runs before the recomposition :856-862)
  LRR_F_NOFLAT = 16384,   // A/B switch: lane-per-row ball scans instead of the flattened warp-cooperative one (lrr_dev.cuh)
  LRR_F_BRACKET = 512     // order statistics from the window around the previous iteration's values (no histogram atomics, no collect pass)
};

struct EstArgs {
  DirArgs d[2];
  int ndir;
  int flags;
  int dof;
  int src;             // direction that defines the normalisation (0; 1 for rrl estimate_LS_backward)
  int estimate_parameters;
  int tile_sparse;     // certificate pass: chunks with at most this many queries left put them on the sparse list (DirArgs::fq) instead of the staged work list
  SelDev selG[2], selS[2];
  double* partials;    // LRR_MAX_BLOCKS * LRR_NRED per direction slot (2 slots)
  unsigned int* ticket;   // per direction slot
};

// launch with programmatic stream serialization (see lrr_pdl_enter, lrr_dev.cuh); LRR_NO_PDL=1 falls back to plain launches (A/B)
#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
template <class... KArgs, class... Args>
static inline cudaError_t lrr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool off = getenv("LRR_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
  // (a failed launch — e.g. a function attribute missing on this device — stays in the runtime's sticky last error, which every
  //  entry point checks when it reads the scalars back; the status is returned for callers that want it at once)
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---- grid build / NN  (lrr_grid.cu) ----------------------------------------------------------------
struct FsetBuild {
  int64_t n;
  const double *x, *y, *z;          // device SoA, caller order
  int* cell_start; double4* pos4; int* sorted_of_orig;
  GridDev grid;
};
cudaError_t lrr_grid_build(FsetBuild* fb, cudaStream_t st, long long* launches);
cudaError_t lrr_nn_query(const GridDev& g, int64_t nq, const double* qx, const double* qy, const double* qz,
                         int* out_orig, cudaStream_t st, long long* launches);
#define LRR_MAX_DEVICES 64
int lrr_num_sms();          // of the CURRENT device (cached per device)
int lrr_current_device();   // clamped to [0, LRR_MAX_DEVICES)

// ---- ICP kernels (lrr_icp.cu) ----------------------------------------------------------------------
void lrr_k_build_proc(const FsetDev& from_all, const int* forig, int n_in, int* mark /* from_all.n ints scratch */, int* proc, cudaStream_t st, long long* launches);
void lrr_k_fill_records(const FsetDev& f, double* rec, double* crec, cudaStream_t st, long long* launches);
// planar cdcl_feature<2> sets: P6 / C6 (3-D layout, z parts zero) from the 2x2 covariances (xx, xy, yy) x n
void lrr_k_projector2_from_cov(const double* cov3, int n, double* P6, double* C6, cudaStream_t st, long long* launches);
void lrr_k_gather_from(const FsetDev& src, const int* idx, int n_in, double* fx, double* fy, double* fz, double* fs, double* fP6,
                       cudaStream_t st, long long* launches);
void lrr_k_match(const EstArgs& a, cudaStream_t st, long long* launches, bool cert_only = false);   // certificate pass / pass A + pass B
void lrr_k_match_rest(const EstArgs& a, cudaStream_t st, long long* launches);   // what a cert_only call left out: staged search + pass B
void lrr_k_match_search(const EstArgs& a, cudaStream_t st, long long* launches);  // ... its first half: staged search / sparse list / ramp-level pass A
void lrr_k_match_queue(const EstArgs& a, cudaStream_t st, long long* launches);   // ... its second half: pass B over the queue
// lrr_tile.cu: tile construction and the shared-memory-staged pass A of the finest level
void lrr_k_build_tiles(const double* fx, const double* fy, const double* fz, int n_in, const TileGeom& g, int* tile_cnt, int* tile_start,
                       int* chunk_start, int* tile_of, int* tmp, int2* chunk_tab, int* proc, cudaStream_t st, long long* launches);
// tile-order mirrors for the certificate pass: pos_of[proc[tq]] = tq, tf*[tq] = f*[proc[tq]]
void lrr_k_tile_mirror(const int* proc, const double* fx, const double* fy, const double* fz, int n_in, int* pos_of, double* tf, cudaStream_t st, long long* launches);
bool lrr_match_tile_usable(const EstArgs& a);
void lrr_k_match_tile(const EstArgs& a, cudaStream_t st, long long* launches);
bool lrr_match_cert_usable(const EstArgs& a);
// mode: LRR_CERT_RESID (1) = also the residual pass of the finished iteration (+ result mailbox), LRR_CERT_RESCALE (2) = the stored
// weights are final up to Scal::wfac (fused estimate pass)
void lrr_k_match_cert(const EstArgs& a, cudaStream_t st, long long* launches, int mode = 0, double* mbox = nullptr, unsigned long long seq = 0,
                      int pull_mask = 0, int nhead = 0);
void lrr_k_restamp(const EstArgs& a, cudaStream_t st, long long* launches, int skip_clean = 0);
void lrr_k_map_clones_again(Scal* sc, cudaStream_t st, long long* launches);   // from_mapped 1 -> 2: the clones are mapped through the current transform once more   // stamps of the source direction's matches (DirArgs::stamp_in_match == 0)
void lrr_tile_set_timeline(unsigned long long* p);
void lrr_k_norm1(const EstArgs& a, cudaStream_t st, long long* launches, int skip_clean = 0);
void lrr_k_norm2(const EstArgs& a, cudaStream_t st, long long* launches, int skip_clean = 0);
void lrr_k_norm_merged(const EstArgs& a, cudaStream_t st, long long* launches);   // cdcl estimate_LS_backward: normalisation over the merged list
void lrr_k_weight(const EstArgs& a, cudaStream_t st, long long* launches);
void lrr_k_collect(const EstArgs& a, cudaStream_t st, long long* launches, int spec = 0);
void lrr_k_accumulate(const EstArgs& a, cudaStream_t st, long long* launches, int spec = 0);
void lrr_k_resid(const EstArgs& a, int dir_mask, cudaStream_t st, long long* launches, double* mbox = nullptr, unsigned long long seq = 0,
                 int pull_mask = 0, int nhead = 0, int spec = 0);   // mbox: result mailbox in mapped host memory (lrr_icp.cu)
void lrr_k_angles(const DirArgs& a, int flags, int which, double* partials, unsigned int* ticket, cudaStream_t st, long long* launches);
void lrr_k_transfer_cov(const DirArgs& a, int dof, double* partials, unsigned int* ticket, cudaStream_t st, long long* launches);
void lrr_k_resid_keys(const DirArgs& a, double* keys, cudaStream_t st, long long* launches);
void lrr_k_export_matches(const DirArgs& a, int* from_idx, int* to_idx, cudaStream_t st, long long* launches);
void lrr_k_objcov(const DirArgs& a, int dof, const double* U /*dof x dof upper, device*/, int want_g, double* partials,
                  unsigned int* ticket, cudaStream_t st, long long* launches);

// ---- CDC E-step: covariance-driven radius matching (lrr_cdc.cu) -----------------------------------------
struct CdcLaunch {
  const double *A, *t, *c, *Sigma_dev; double eval_mult; int max_matches, M;
  const double *fx, *fy, *fz; const int* forig; FsetDev to, from_all;
  int *efrom, *eslot; double* ew; long long ecap; unsigned long long* cursor; int *rowcnt, *rowstart; double *roww, *tow, *dsum; int* istat;
  int2* cand; double* candw; int blocks;
};
int lrr_cdc_blocks(int M);
size_t lrr_cdc_scratch_entries(int blocks);
void lrr_k_cdc_match(const CdcLaunch& L, cudaStream_t st, long long* launches);
void lrr_k_cdc_normalise(int n, const int* efrom, const int* eslot, double* ew, const double* roww, const double* tow, cudaStream_t st, long long* launches);
void lrr_k_cdc_norm(int M, int nto, const double* fx, const double* fy, const double* fz, const int* rowcnt, const double* tow, const double4* pos4,
                    double* out10, cudaStream_t st, long long* launches);

// ---- sort / MUSE (lrr_select.cu) -------------------------------------------------------------------
struct SortScratch { double* keys2; unsigned int* hist; size_t cap; };
cudaError_t lrr_sort_alloc(SortScratch* s, size_t cap);
void lrr_sort_free(SortScratch* s);
// ascending sort of n non-negative doubles in place (single-block LSD radix sort; once per registration)
void lrr_sort_doubles(SortScratch* s, double* keys, int n, cudaStream_t st, long long* launches);
// MUSE objective over sorted |r| (table: per fraction {k, expected_kth, stddev_kth, divisor}); writes *out_scale
void lrr_muse_from_sorted(const double* sorted, int n, const double* table, int nfrac, double* out_scale, cudaStream_t st,
                          long long* launches);
// k-th smallest of sorted keys -> MAD scale (fallback of the histogram selector)
void lrr_mad_from_sorted(const double* sorted, int n, double* out, cudaStream_t st, long long* launches);
