// Internal declarations shared by the CUDA translation units of liblrr_b200.so (sm_100a only).
// Data layout in HBM:
//   * "from" side of a direction (the streamed side): compact structure-of-arrays in subset order,
//     fp64 x,y,z,strength (+ P6, C6, dirs, u8 shape) — coalesced loads.
//   * "to" side (the gathered side): the feature set sorted by uniform-grid cell, one 96-byte record
//     per feature (x,y,z,Pxx,Pxy,Pxz,Pyy,Pyz,Pzz,strength,orig-index bits,pad): one match gathers 3
//     32-byte sectors instead of 10 scattered SoA sectors.  A compact interleaved xyz copy (24 B/feature)
//     in the same order feeds the exact 1-NN search; x-adjacent cells are contiguous in that order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LRR_REC_STRIDE 12   // doubles per gather record (96 B)
#define LRR_REC2_STRIDE 16  // doubles per auxiliary record: C6(6) dirs(9) shape(1)
#define LRR_MAX_BLOCKS 1184 // 148 SMs x 8
#define LRR_NRED 96         // max doubles reduced per block in one pass

struct GridDev {
  double ox, oy, oz, h, inv_h;
  int nx, ny, nz, n;
  const int* cell_start;   // ncell+1
  const double* gpos;      // 3*n interleaved, cell order
  const int* gorig;        // n: original index of sorted slot
};

// Device-resident scalars of one direction (forward_ or backward_ of an estimator).
struct Scal {
  // current transform of this direction, unnormalised: y = A (x - c) + t
  double A[9], t[3], c[3];
  double cov[144];
  int trans_kind, dof;
  // previous global transform (rigid increment bookkeeping: the clones of cdcl :455-484)
  double Ag[9], tg[3], cg[3];
  int from_mapped;         // 1 while matches' from_ are clones mapped through (Ag,tg,cg)
  // normalisation (A9) and the normalised transform
  double cm[3], rm, cf[3], rf;
  double An[9], tn[3], cn[3];
  // reductions
  double sum_x[3], sum_q[3], n_unique, sum_rad_m, sum_rad_f;
  double sum_r2, sum_wr2, sum_w, n_nonzero;
  double err_num, err_den;     // weighted_error
  double rms_sum;
  double scale; int scale_estimated; int pad0;
  double G, S;                 // strength / spatial MAD scales
  double AtA[144], Atb[12];
  int solve_status; int pad1;
  double ang_num, ang_den;
  double covJ[9];
  double obj_f; double obj_kl[144];
};

struct DirArgs {
  // from side
  const double *fx, *fy, *fz, *fs;
  const double *fP6, *fC6, *fdirs;   // SoA with leading dimension n_in
  const uint8_t* fshape;
  const int* forig;
  int n_in;
  const int* order;
  int M;
  // to side
  const double* rec;
  const double* rec2;
  GridDev grid;
  int n_to;
  // per-match arrays (length n_in)
  int* sidx;
  double *w, *r2, *g, *sw;
  unsigned char* rep;
  int* stamp;
  int stamp_id;
  // scalars + reduction scratch
  Scal* sc;
  double* partials;       // LRR_MAX_BLOCKS * LRR_NRED
  unsigned int* ticket;
};

enum {
  LRR_F_RRL = 1,          // rrl flavour: fp32 strength product, w<1e-4 skip in accumulation
  LRR_F_RIGID = 2,        // rigid3d increment: from = mapped clone, J = [-[x]x | I], b = q - p
  LRR_F_NORMALIZED = 4,   // evaluate in normalised coordinates
  LRR_F_REVERSED = 8,     // accumulate with roles reversed (J at to, P and b of from)
  LRR_F_STRENGTH = 16,
  LRR_F_SPATIAL = 32,
  LRR_F_SKIP_SMALL = 64,
  LRR_F_USE_OLD_W = 128   // scale from current weights (scale_estimated_)
};

struct LaunchCounter { long long n; };

// grid build / NN  (lrr_grid.cu)
struct FsetBuild {
  int64_t n;
  const double *x, *y, *z;          // device SoA, caller order
  int* cell_start; double* gpos; int* gorig; int* sorted_of_orig;
  GridDev grid;
};
cudaError_t lrr_grid_build(FsetBuild* fb, cudaStream_t st, long long* launches);
cudaError_t lrr_nn_query(const GridDev& g, int64_t nq, const double* qx, const double* qy, const double* qz,
                         int* out_orig, cudaStream_t st, long long* launches);
int lrr_num_sms();

// ICP kernels (lrr_icp.cu)
void lrr_k_gather_from(const double* sx, const double* sy, const double* sz, const double* ss, const double* sP6, const double* sC6,
                       const double* sdirs, const uint8_t* sshape, int64_t n_src, const int* idx, int n_in,
                       double* fx, double* fy, double* fz, double* fs, double* fP6, double* fC6, double* fdirs, uint8_t* fshape,
                       cudaStream_t st, long long* launches);
void lrr_k_fill_records(const double* sx, const double* sy, const double* sz, const double* ss, const double* sP6, const double* sC6,
                        const double* sdirs, const uint8_t* sshape, int64_t n, const int* gorig, double* rec, double* rec2,
                        cudaStream_t st, long long* launches);
void lrr_k_iota(int* a, int n, cudaStream_t st, long long* launches);
void lrr_k_match(const DirArgs& a, cudaStream_t st, long long* launches);
void lrr_k_radii(const DirArgs& a, int known /*0: compute, 1: known (cm,rm,cf,rf) already in sc*/, cudaStream_t st, long long* launches);
void lrr_k_resid(const DirArgs& a, int flags, cudaStream_t st, long long* launches);
void lrr_k_weight(const DirArgs& a, int flags, double* gsel_out, cudaStream_t st, long long* launches);
void lrr_k_normal(const DirArgs& a, int flags, const Scal* trans_owner, int accumulate_into_existing, const double* GS /*device: G,S*/, cudaStream_t st, long long* launches);
void lrr_k_error(const DirArgs& a, int which, int flags, cudaStream_t st, long long* launches);
void lrr_k_transfer_cov(const DirArgs& a, cudaStream_t st, long long* launches);
// scalar epilogues (single-block kernels)
void lrr_k_prepare_norm(Scal* f, Scal* b_or_null, int normalize, int flags, cudaStream_t st, long long* launches);
void lrr_k_solve(Scal* f, Scal* b_or_null, int estimate_parameters, int normalize, int flags, cudaStream_t st, long long* launches);
void lrr_k_scale_from_sums(Scal* s, cudaStream_t st, long long* launches);

// selection / sort (lrr_select.cu)
struct SelectScratch { unsigned int* hist; unsigned long long* state; double* keys; double* keys2; size_t cap; };
cudaError_t lrr_select_alloc(SelectScratch* s, size_t cap);
void lrr_select_free(SelectScratch* s);
// MAD scale (Q12) over keys[0..n) where n is read from device (*n_dev as double); result -> *out_dev
void lrr_mad_scale(SelectScratch* s, const double* keys, const double* n_dev, int n_max, double* out_dev, cudaStream_t st, long long* launches);
// sort ascending (for MUSE) n host-known
void lrr_sort_doubles(SelectScratch* s, double* keys, int n, cudaStream_t st, long long* launches);
void lrr_muse_from_sorted(const double* sorted, int n, const double* table /*device: 3 doubles per fraction + k*/, int nfrac, double* out_scale, cudaStream_t st, long long* launches);
