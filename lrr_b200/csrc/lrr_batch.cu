// Batched location registration on the device: ONE THREAD BLOCK PER REGISTRATION (sm_100a).
// The outer loops of LocationRegistration::Run (rrl/itkLocationRegistration.cxx:2138,2230,2332) hand out
// independent (location, initialisation) jobs of 1000-16000 features; launching the eight per-iteration kernels of
// the single-registration path for each of them would be pure launch latency.  Here a persistent grid of blocks
// pulls jobs from an atomic counter; each block runs the whole registration — ROI selection (FeaturesInROIs,
// :1324-1360), set_transform, the ramp / oscillation state machine of
// rrl_estimation_symmetric_ICP_matching_all<3,12>::one_iteration (rrl sym :341-438), validTransform (:2914-2926) —
// with the SAME phase bodies as the grid-wide kernels, instantiated for BlockScope (block-level reductions,
// __syncthreads instead of kernel boundaries).  No host round trip per iteration, no launches per iteration.
#include "lrr_phases.cuh"
#include "../../include/lrr_gpu.h"
#include "lrr_host.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <algorithm>

#define BATCH_CAP 16384            // max features of one ROI (larger ROIs are returned to the host path)
#define BATCH_SEL_BINS 2048
#define BATCH_SEL_CAND 1024
#ifndef BATCH_THREADS
#define BATCH_THREADS 512           // one block of 16 warps per SM: half the per-registration latency of 2 x 256
#endif
#define BATCH_MIN_BLOCKS (512 / BATCH_THREADS)

struct BatchWork {                 // per-block workspace in global memory
  // per direction
  double* f[2];                    // 4 x CAP: x y z strength
  double* fP6[2];                  // 6 x CAP
  int* forig[2];                   // CAP
  int* sidx[2]; double* pm[2];     // per match: sidx, 4 x CAP doubles (w r2 gkey skey)
  int* prev_s[2];
  int* hq[2]; double* hq_best[2];  // 2 queues x (3 x CAP + 4) ints, 2 x CAP doubles
  unsigned long long* stamp[2];    // n_to
  Scal* sc[2];
};

struct BatchArgs {
  FsetDev all[2];                  // 0 = moving_all, 1 = fixed_all
  const lrr_job* jobs; lrr_result* out; long long njobs;
  int max_iter; unsigned long long shuffle_seed;
  const long long* job_ids;          // global index of job j (null: j): the shuffle seed of a job is shuffle_seed + its GLOBAL index, however the list was sharded
  int* next_job;
  // workspace bases (strided by block)
  double* w_f; double* w_fP6; int* w_forig; int* w_sidx; double* w_pm; int* w_prev; int* w_hq; double* w_hqb;
  unsigned long long* w_stamp0; unsigned long long* w_stamp1; Scal* w_sc;
  double muse_table[4 * 32]; int muse_rows;   // table for n = 1000
  // the driver's candidate loop (lrr_batch_run_locations): mode 0 = registration only (lrr_batch_run), 1 = ranking angles only,
  // 2 = registration + decision measurements; order != null: a unit of work is a LOCATION whose candidates (order[l * ncand + k], -1
  // terminated) are run one after the other, until_verified stops at the first verified one
  int mode, ncand, until_verified; const int* order; lrr_decision* dec; double spacing; long long nunits;
  int ball_rows, g8;                            // pass A keeps balls of at most ball_rows rows; the rest: eight-lane groups (g8) / one warp per query
  long long* dbg;                               // optional: 16 cycle counters (block 0), phase profile
  // TIME SLICING (plain job lists only).  A job's time is iterations x ROI points and the iteration count (1 ... 70) is not known before
  // it runs: with run-to-completion the jobs that hit the cap late in the queue are the tail of the launch (0.74 of ideal on an
  // eighth of the C4 list).  Here the workspace is a pool of SLOTS (one per job in flight, more slots than blocks), a block runs a job
  // for at most `quantum` iterations, parks it — control state into the slot's BatchCtl, everything else already lives in the slot —
  // and takes the next unit: fresh jobs first while slots are free (every job is started early), then parked jobs in FIFO order.
  // quantum == 0: one slot per block, run to completion (the candidate loop of lrr_batch_run_locations).
  int quantum, nslots;
  int* sched;                                   // [0] lock, [1] next fresh unit, [2] free slots, [3] queue head, [4] queue length, then free[nslots], queue[nslots]
  struct BatchCtl* ctl;                         // per slot
};
struct BatchCtl {
  int job, n_in[2];
  int converged, status, it, finest, osc;
  unsigned int number_matches, seq;             // seq: registrations this SLOT has started — the stamp epochs must grow over the life of its workspace
  double err_diff, werr_before;
  unsigned long long calls[2];
  int number_points[2], scale_estimated[2];
  // instrumentation (LRR_BATCH_DEBUG=2 prints the latest finishers): first start / last end of the slot's current job (globaltimer ns), its
  // slices, the time it spent parked
  unsigned long long t_first, t_last, t_parked_at, t_waited; int slices, done_job, done_it, pad;
};
__device__ __forceinline__ unsigned long long batch_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ unsigned long long bt_splitmix(unsigned long long& s) {
  unsigned long long z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// ---- block-level helpers -------------------------------------------------------------------------------------
// ascending bitonic sort of n ints (n <= BATCH_CAP) in shared memory `buf` (size = next power of two >= n)
__device__ void block_sort_ints(int* buf, int n) {
  int np = 1; while (np < n) np <<= 1;
  for (int i = n + threadIdx.x; i < np; i += blockDim.x) buf[i] = 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= np; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int a = buf[i], b = buf[ixj];
          if ((a > b) == ((i & k) == 0)) { buf[i] = b; buf[ixj] = a; }
        }
      }
      __syncthreads();
    }
}
// ascending bitonic sort of 1024 doubles in shared memory
__device__ void block_sort_1024(double* buf) {
  for (int k = 2; k <= 1024; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const double a = buf[i], b = buf[ixj];
          if ((a > b) == ((i & k) == 0)) { buf[i] = b; buf[ixj] = a; }
        }
      }
      __syncthreads();
    }
}

// FeaturesInROIs: indices (ascending, like the reference's scan over the feature vector) of the features of `f`
// inside the cube; returns the count in every thread (BATCH_CAP + 1 on overflow).  list = shared memory ints.
__device__ int block_roi(const FsetDev& f, const double* c, double half, int* list, int* s_cnt) {
  const GridDev& g = f.grid;
  if (threadIdx.x == 0) *s_cnt = 0;
  __syncthreads();
  const double lo0 = c[0] - half, hi0 = c[0] + half, lo1 = c[1] - half, hi1 = c[1] + half, lo2 = c[2] - half, hi2 = c[2] + half;
  const int x0 = lrr_cell_clamp(lo0, g.ox, g.inv_h, g.nx), x1 = lrr_cell_clamp(hi0, g.ox, g.inv_h, g.nx);
  const int y0 = lrr_cell_clamp(lo1, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(hi1, g.oy, g.inv_h, g.ny);
  const int z0 = lrr_cell_clamp(lo2, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(hi2, g.oz, g.inv_h, g.nz);
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  // one warp per row, lanes stride over the row's points
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = wid; r < nrows; r += nw) {
    const int z = z0 + r / ny, y = y0 + r % ny, row = (z * g.ny + y) * g.nx;
    const int lo = __ldg(g.cell_start + row + x0), hi = __ldg(g.cell_start + row + x1 + 1);
    for (int j = lo + lane; j < hi; j += 32) {
      const double4 p = lrr_ld_pos4(g.pos4 + j);
      if (!(lo0 > p.x || p.x > hi0 || lo1 > p.y || p.y > hi1 || lo2 > p.z || p.z > hi2)) {
        const int slot = atomicAdd(s_cnt, 1);
        if (slot < BATCH_CAP) list[slot] = lrr_pos4_orig(p);
      }
    }
  }
  __syncthreads();
  const int n = *s_cnt;
  __syncthreads();
  if (n > BATCH_CAP) return BATCH_CAP + 1;
  block_sort_ints(list, n);
  return n;
}

// exact MAD scale (cdcl :520-536, Q12) of the keys >= 0 among keys[0..M): linear histogram in shared memory,
// candidates of the target bin, rank selection.  Returns the scale in every thread; *ok = 0 on candidate overflow.
__device__ double block_mad(const double* __restrict__ keys, int M, double inv_width, unsigned int* hist /* BINS */, double* cand /* CAND */,
                            int* s_i /* >= 8 ints */, double* s_d /* >= 2 */, int* ok) {
  for (int i = threadIdx.x; i < BATCH_SEL_BINS; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) { s_i[0] = 0; s_i[1] = 0; s_i[2] = -1; s_i[3] = 0; }
  __syncthreads();
  int local = 0;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    const double key = keys[k];
    if (key >= 0.0) { int b = (int)(key * inv_width); b = b < 0 ? 0 : (b >= BATCH_SEL_BINS ? BATCH_SEL_BINS - 1 : b); atomicAdd(&hist[b], 1u); ++local; }
  }
  atomicAdd(&s_i[0], local);
  __syncthreads();
  const int n = s_i[0];
  if (n <= 1) { __syncthreads(); return 0.0; }
  const unsigned int target = (unsigned int)((n - 1) / 2 + 1);
  if (threadIdx.x == 0) {   // 2048 bins: a serial scan is ~2 us; the bins are in shared memory
    unsigned int base = 0; int b = 0;
    for (; b < BATCH_SEL_BINS; ++b) { if (target >= base + hist[b]) base += hist[b]; else break; }
    s_i[2] = b; s_i[3] = (int)(target - base);
  }
  __syncthreads();
  const int bin = s_i[2], rank = s_i[3];
  unsigned long long* mm = reinterpret_cast<unsigned long long*>(s_d);
  if (threadIdx.x == 0) { mm[0] = ~0ull; mm[1] = 0ull; }
  __syncthreads();
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    const double key = keys[k];
    if (key >= 0.0) {
      int b = (int)(key * inv_width); b = b < 0 ? 0 : (b >= BATCH_SEL_BINS ? BATCH_SEL_BINS - 1 : b);
      if (b == bin) {
        const int slot = atomicAdd(&s_i[1], 1);
        if (slot < BATCH_SEL_CAND) cand[slot] = key;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(key);
        atomicMin(&mm[0], bits); atomicMax(&mm[1], bits);
      }
    }
  }
  __syncthreads();
  const int cnt = s_i[1];
  __shared__ double s_val;
  if (cnt > BATCH_SEL_CAND) {
    if (threadIdx.x == 0) { if (mm[0] == mm[1]) s_val = __longlong_as_double((long long)mm[0]); else { s_val = 0.0; *ok = 0; } }
  } else {
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const double v = cand[i]; int r = 0;
      for (int j = 0; j < cnt; ++j) { const double u = cand[j]; r += (u < v) || (u == v && j < i); }
      if (r == rank) s_val = v;
    }
  }
  __syncthreads();
  const double res = 1.4826 * (1.0 + 5.0 / (double)(n - 1)) * s_val;
  __syncthreads();
  return res;
}

// MUSE scale of the residuals of one direction (M <= 1024)
__device__ double block_muse(const DirArgs& a, const BatchArgs& B, double* buf /* 1024 doubles shared */, double* s_seg /* 33 */) {
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i < a.M) ? sqrt(a.r2[i]) : 1e308;
  __syncthreads();
  block_sort_1024(buf);
  // segment sums between consecutive table rows (k ascending), one warp per segment
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int f = wid; f < B.muse_rows; f += nw) {
    const int k0 = f ? (int)B.muse_table[4 * (f - 1)] : 0, k1 = (int)B.muse_table[4 * f];
    double s = 0;
    for (int i = k0 + lane; i < k1 && i < 1024; i += 32) s += buf[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_seg[f] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0, best_sk = 0, best_obj = 0; bool first = true;
    for (int f = 0; f < B.muse_rows; ++f) {
      sum += s_seg[f];
      const double sk = sum / B.muse_table[4 * f + 3], obj = sk * B.muse_table[4 * f + 2] / B.muse_table[4 * f + 1];
      if (first || obj < best_obj) { first = false; best_sk = sk; best_obj = obj; }
    }
    s_seg[32] = best_sk;
  }
  __syncthreads();
  const double r = s_seg[32];
  __syncthreads();
  return r;
}

// ComputeTransferErrorCovariance (rrl/itkLocationRegistration.cxx:2793-2858): maximum over a 5 mm grid of the ROI of
// trace( J(x) Sigma J(x)^T ), J the affine Jacobian at x (x~ = (x - centre, 1)): sum_d sum_ab x~_a x~_b Sigma[(d,a),(d,b)].
// start = ROI centre - half, end = start + 2 half - voxel spacing; the grid counts follow the reference's floating loops.
__device__ double block_max_trace(const Scal* sc, const double* centre, double half, double spacing, double* s_d /* >= 2 doubles */) {
  __shared__ int s_n[3];
  unsigned long long* mx = reinterpret_cast<unsigned long long*>(s_d);
  if (threadIdx.x < 3) {
    const double start = centre[threadIdx.x] - half, end = start + 2.0 * half - spacing;
    int n = 0; for (double v = start; v <= end; v += 5.0) ++n;
    s_n[threadIdx.x] = n;
  }
  if (threadIdx.x == 0) mx[0] = 0ull;
  __syncthreads();
  const int nx = s_n[0], ny = s_n[1], nz = s_n[2], total = nx * ny * nz;
  double best = 0.0;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    const int ix = p % nx, iy = (p / nx) % ny, iz = p / (nx * ny);
    // (coordinates by repeated addition, as the reference's loops form them)
    double x = centre[0] - half; for (int q = 0; q < ix; ++q) x += 5.0;
    double y = centre[1] - half; for (int q = 0; q < iy; ++q) y += 5.0;
    double z = centre[2] - half; for (int q = 0; q < iz; ++q) z += 5.0;
    const double xt[4] = {x - sc->c[0], y - sc->c[1], z - sc->c[2], 1.0};
    double tr = 0.0;
    for (int d = 0; d < 3; ++d) for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
      const int ia = a < 3 ? 3 * d + a : 9 + d, ib = b < 3 ? 3 * d + b : 9 + d;
      tr += xt[a] * xt[b] * sc->cov[ia * 12 + ib];
    }
    best = fmax(best, tr);
  }
  if (best > 0.0) atomicMax(mx, (unsigned long long)__double_as_longlong(best));
  __syncthreads();
  const double r = __longlong_as_double((long long)mx[0]);
  __syncthreads();
  return r;
}

struct BatchShared {
  PhaseSharedT<BATCH_THREADS / 32> ph;
  DirArgs d[2];
  unsigned int hist[BATCH_SEL_BINS];
  double cand[BATCH_SEL_CAND];          // also the 1024-double MUSE sort buffer
  double seg[40];
  int si[16]; double sd[4];
  int cnt;
  DirArgs d2[2];                        // the same directions with the second queue (what the eight-lane groups pass on)
  int rows[(BATCH_THREADS / 8) * 2 * LRR_G8_ROWS];   // row tables of the eight-lane groups
  // control state (thread 0 writes, everybody reads after a barrier)
  int converged, status, it, finest, osc, job;
  int slot, resume;   // workspace slot of the unit in hand; 1: a parked job is taken up again
  unsigned int seq;   // registrations the slot has started (BatchCtl::seq): the stamp epochs must GROW over the life of its workspace (jobs may come in any order)
  unsigned int number_matches;
  double err_diff, werr_before;
  unsigned long long calls[2];
  int number_points[2], scale_estimated[2];
  long long tph[16];
};

// ---- slot pool and FIFO of parked jobs (one thread per block calls these; a spin lock in global memory: a few thousand operations
//      per launch) ----
__device__ __forceinline__ void sched_lock(int* sched) { while (atomicCAS(sched, 0, 1) != 0) { __nanosleep(64); } __threadfence(); }
__device__ __forceinline__ void sched_unlock(int* sched) { __threadfence(); atomicExch(sched, 0); }
// next unit of work: job >= 0 fresh (slot reserved), job == -2 a parked job (slot = its slot), job == -1 nothing left for this block
__device__ void sched_take(const BatchArgs& B, int& job, int& slot) {
  int* sc = B.sched; int* fr = sc + 8; int* q = fr + B.nslots;
  volatile int* v = sc;
  for (;;) {
    sched_lock(sc);
    if (v[1] < (int)B.nunits && v[2] > 0) { job = v[1]; sc[1] = job + 1; slot = fr[v[2] - 1]; sc[2] = v[2] - 1; sched_unlock(sc); return; }
    if (v[4] > 0) { slot = q[v[3]]; sc[3] = (v[3] + 1) % B.nslots; sc[4] = v[4] - 1; job = -2; sched_unlock(sc); return; }
    const bool fresh_left = v[1] < (int)B.nunits;
    sched_unlock(sc);
    if (!fresh_left) { job = -1; slot = -1; return; }   // parked jobs that appear later are taken up by the block that parks them
    __nanosleep(2000);                                  // every slot is held by a RUNNING block (nslots >= blocks keeps this from happening)
  }
}
__device__ void sched_release(const BatchArgs& B, int slot) {
  int* sc = B.sched; int* fr = sc + 8; volatile int* v = sc;
  sched_lock(sc); fr[v[2]] = slot; sc[2] = v[2] + 1; sched_unlock(sc);
}
__device__ void sched_park(const BatchArgs& B, int slot) {
  int* sc = B.sched; int* q = sc + 8 + B.nslots; volatile int* v = sc;
  sched_lock(sc); q[(v[3] + v[4]) % B.nslots] = slot; sc[4] = v[4] + 1; sched_unlock(sc);
}

extern __shared__ __align__(16) unsigned char batch_smem[];
#define TPH(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); S.tph[i] += now_ - tmark; tmark = now_; } } while (0)

__global__ void __launch_bounds__(BATCH_THREADS, BATCH_MIN_BLOCKS) k_batch(const __grid_constant__ BatchArgs B) {
  BatchShared& S = *reinterpret_cast<BatchShared*>(batch_smem);
  int* list = reinterpret_cast<int*>(batch_smem + ((sizeof(BatchShared) + 15) & ~size_t(15)));   // BATCH_CAP ints
  const BlockScope scope;
  const int flags = LRR_F_RRL | LRR_F_NORMALIZE | LRR_F_SYM | LRR_F_STRENGTH;
  SelDev nosel; nosel.fine = nullptr;
  long long tmark = clock64();
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) S.tph[i] = 0; S.seq = 0; }

  for (;;) {
    if (threadIdx.x == 0) {
      if (B.quantum > 0) { int jb, sl; sched_take(B, jb, sl); S.resume = jb == -2; S.slot = sl; S.job = jb == -2 ? B.ctl[sl].job : jb; }
      else { const int jb = atomicAdd(B.next_job, 1); S.job = jb < B.nunits ? jb : -1; S.slot = blockIdx.x; S.resume = 0; }
    }
    __syncthreads();
    const int unit = S.job; const bool resume = S.resume != 0;
    const size_t blk = (size_t)(S.slot < 0 ? 0 : S.slot);   // the workspace slot (quantum == 0: the block's own)
    __syncthreads();
    if (unit < 0) break;
    // a slot's contents may have been written by another SM (the block that parked the job, or an earlier tenant of the slot): acquire
    // at device scope in every thread before anything of the slot is read (the fence also drops this SM's L1 lines)
    if (B.quantum > 0) __threadfence();
   int cpos = 0; bool unit_verified = false;
   do {   // the candidates of a location, best angle first (rrl/itkLocationRegistration.cxx:2332); a plain job list has one "candidate" per unit
    int j = unit;
    if (B.order) { const int oc = B.order[(size_t)unit * B.ncand + cpos]; if (oc < 0) break; j = unit * B.ncand + oc; }
    ++cpos;
    const lrr_job job = B.jobs[j];
    lrr_result* R = B.out + j;
    BatchCtl* C = B.ctl + blk;
    bool parked = false;
    // ---- set_features: ROI selection for both directions (forward from = moving, backward from = fixed) ----
    int n_in[2]; bool reject = false, overflow = false;
    if (resume) { n_in[0] = C->n_in[0]; n_in[1] = C->n_in[1]; }
    else for (int d = 0; d < 2; ++d) {
      const FsetDev& fa = B.all[d];
      const int n = block_roi(fa, d == 0 ? job.roi_center_moving : job.roi_center_fixed, job.roi_half_size, list, &S.cnt);
      n_in[d] = n;
      if (n > BATCH_CAP) { overflow = true; break; }
      if (n < 1000) reject = true;
      double* f = B.w_f + (blk * 2 + d) * 4 * BATCH_CAP; double* fP6 = B.w_fP6 + (blk * 2 + d) * 6 * BATCH_CAP;
      int* forig = B.w_forig + (blk * 2 + d) * BATCH_CAP; int* prev = B.w_prev + (blk * 2 + d) * BATCH_CAP;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int o = list[i];
        forig[i] = o; prev[i] = -1;
        f[i] = fa.x[o]; f[BATCH_CAP + i] = fa.y[o]; f[2 * BATCH_CAP + i] = fa.z[o]; f[3 * BATCH_CAP + i] = fa.s[o];
#pragma unroll
        for (int p = 0; p < 6; ++p) fP6[(size_t)p * n + i] = fa.P6[(size_t)p * fa.n + o];
      }
      __syncthreads();
    }
    lrr_decision* D = B.dec ? B.dec + j : nullptr;
    if (threadIdx.x == 0 && !resume) {
      for (size_t i = 0; i < sizeof(lrr_result) / 8; ++i) reinterpret_cast<double*>(R)[i] = 0.0;
      R->n_moving_inside = overflow ? -1 : n_in[0]; R->n_fixed_inside = overflow ? -1 : n_in[1];
      R->status = overflow ? -1 : (reject ? 1 : 0);     // -1: ROI larger than BATCH_CAP, the host path takes the job
      if (D && B.mode == 1) { for (size_t i = 0; i < sizeof(lrr_decision) / 8; ++i) reinterpret_cast<double*>(D)[i] = 0.0; D->rank = -1; }
    }
    __syncthreads();
    TPH(0);
    if (!(overflow || reject)) {

    // ---- direction descriptors ----
    if (threadIdx.x < 2) {
      const int d = threadIdx.x;
      DirArgs a; memset(&a, 0, sizeof(a));
      double* f = B.w_f + (blk * 2 + d) * 4 * BATCH_CAP;
      a.fx = f; a.fy = f + BATCH_CAP; a.fz = f + 2 * BATCH_CAP; a.fs = f + 3 * BATCH_CAP;
      a.fP6 = B.w_fP6 + (blk * 2 + d) * 6 * BATCH_CAP; a.forig = B.w_forig + (blk * 2 + d) * BATCH_CAP; a.n_in = n_in[d];
      a.order = nullptr; a.proc = nullptr; a.M = min(1000, n_in[d]);
      a.to = B.all[1 - d]; a.from_all = B.all[d];
      a.sidx = B.w_sidx + (blk * 2 + d) * BATCH_CAP;
      double* pm = B.w_pm + (blk * 2 + d) * 4 * BATCH_CAP; a.w = pm; a.r2 = pm + BATCH_CAP; a.gkey = pm + 2 * BATCH_CAP; a.skey = pm + 3 * BATCH_CAP;
      a.stamp = (d == 0 ? B.w_stamp0 + blk * (size_t)B.all[1].n : B.w_stamp1 + blk * (size_t)B.all[0].n);
      int* hq = B.w_hq + (blk * 2 + d) * 2 * (3 * BATCH_CAP + 4);
      a.hq_k = hq; a.hq_s = hq + BATCH_CAP; a.hq_o = hq + 2 * BATCH_CAP; a.hq_count = hq + 3 * BATCH_CAP;
      a.hq_best = B.w_hqb + (blk * 2 + d) * 2 * BATCH_CAP;
      a.ball_rows = B.ball_rows;
      a.prev_s = B.w_prev + (blk * 2 + d) * BATCH_CAP;
      a.sc = B.w_sc + blk * 2 + d;
      a.stamp_in_match = 1;   // (no certificates in batched mode: every finished match stamps its to-feature)
      if (!resume) *a.hq_count = 0;
      S.d[d] = a;
      {   // second queue
        int* hq2 = hq + (3 * BATCH_CAP + 4);
        a.hq_k = hq2; a.hq_s = hq2 + BATCH_CAP; a.hq_o = hq2 + 2 * BATCH_CAP; a.hq_count = hq2 + 3 * BATCH_CAP;
        a.hq_best = B.w_hqb + (blk * 2 + d) * 2 * BATCH_CAP + BATCH_CAP;
        if (!resume) *a.hq_count = 0;
        S.d2[d] = a;
      }
      if (!resume) { S.calls[d] = 0; S.number_points[d] = a.M; S.scale_estimated[d] = 0; }
      else { S.calls[d] = C->calls[d]; S.number_points[d] = C->number_points[d]; S.scale_estimated[d] = C->scale_estimated[d]; }
    }
    __syncthreads();
    Scal* F = S.d[0].sc; Scal* Bk = S.d[1].sc;
    // ---- set_transform (cdcl sym :231-242): forward <- job transform, backward <- inverse ----
    if (resume) {
      if (threadIdx.x == 0) {   // the parked job's control state; transforms, scales and matches never left the slot
        S.converged = C->converged; S.status = C->status; S.it = C->it; S.finest = C->finest; S.osc = C->osc; S.number_matches = C->number_matches;
        S.err_diff = C->err_diff; S.werr_before = C->werr_before; S.seq = C->seq;
        C->t_waited += batch_now() - C->t_parked_at;
      }
    } else {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 9; ++i) F->A[i] = job.A[i];
      for (int i = 0; i < 3; ++i) { F->t[i] = job.t[i]; F->c[i] = job.c[i]; }
      double invA[9]; lrr_inv3(F->A, invA);
      for (int i = 0; i < 9; ++i) Bk->A[i] = invA[i];
      for (int i = 0; i < 3; ++i) {
        double acc = (-invA[3 * i]) * F->t[0]; acc += (-invA[3 * i + 1]) * F->t[1]; acc += (-invA[3 * i + 2]) * F->t[2];
        Bk->t[i] = acc + F->c[i]; Bk->c[i] = 0.0;
      }
      F->from_mapped = 0; Bk->from_mapped = 0; F->status = 0; Bk->status = 0;
      S.converged = 0; S.status = 0; S.it = 0; S.finest = 0; S.osc = 0; S.number_matches = 1000; S.err_diff = 0.0; S.werr_before = 0.0;
      C->seq += 1; S.seq = C->seq;
      C->t_first = batch_now(); C->slices = 0; C->t_waited = 0;
    }
    for (int i = threadIdx.x; i < 144; i += blockDim.x) { F->cov[i] = 0.0; Bk->cov[i] = 0.0; }
    }
    __syncthreads();

    if (B.mode == 1) {
      // ---- ranking (rrl/itkLocationRegistration.cxx:2303-2313): find_closest_euclidean at the initial transform (1000 points, its own
      //      permutation stream), then sheet / tube angles forward and backward, in degrees ----
      if (threadIdx.x < 2) {
        const int d = threadIdx.x;
        const unsigned long long gid = B.job_ids ? (unsigned long long)B.job_ids[j] : (unsigned long long)j;
        unsigned long long st = (B.shuffle_seed + gid) + (unsigned long long)d * 0x632BE59BD9B4E019ULL + (1ULL << 40) * 0xD1342543DE82EF95ULL;
        for (int r = 0; r < 4; ++r) { S.d[d].perm_key[r] = (unsigned int)(bt_splitmix(st) >> 32); S.d2[d].perm_key[r] = S.d[d].perm_key[r]; }
        int bits = 2; while ((1ull << bits) < (unsigned long long)S.d[d].n_in) ++bits;
        if (bits & 1) ++bits;
        S.d[d].perm_half = bits / 2; S.d2[d].perm_half = bits / 2;
        S.d[d].M = min(1000, S.d[d].n_in); S.d2[d].M = S.d[d].M;
        const unsigned long long epoch = ((unsigned long long)S.seq << 8) | 255ull;
        S.d[d].tag = (~epoch) << 32; S.d2[d].tag = S.d[d].tag;
      }
      __syncthreads();
      for (int d = 0; d < 2; ++d) {
        lrr_phase_match(scope, S.d[d], flags, S.ph);
        if (B.g8) lrr_phase_match_queue_g8(scope, S.d[d], S.d2[d], flags, S.ph, S.rows);
        lrr_phase_match_hard(scope, B.g8 ? S.d2[d] : S.d[d], flags, S.ph);
      }
      for (int d = 0; d < 2; ++d) for (int which = 0; which < 2; ++which) {
        lrr_phase_angles_rrl(scope, S.d[d], which, S.ph);
        if (threadIdx.x == 0 && D) D->rank_angles[2 * d + which] = (S.d[d].sc->ang_num / S.d[d].sc->ang_den) * (180.0 / 3.14159265358979323846);
        __syncthreads();
      }
      if (threadIdx.x == 0 && D) D->rank_angle_average = (D->rank_angles[0] + D->rank_angles[1] + D->rank_angles[2] + D->rank_angles[3]) / 4;
      __syncthreads();
    } else {
    // ---- the driver loop (rrl/itkLocationRegistration.cxx:2418) ----
    int slice = B.quantum > 0 ? B.quantum : 0x7fffffff;
    while (!S.converged && S.it < B.max_iter && S.status == 0 && slice-- > 0) {
      // find_closest_euclidean, both directions: new permutation key, new stamp epoch
      if (threadIdx.x < 2) {
        const int d = threadIdx.x;
        const unsigned long long gid = B.job_ids ? (unsigned long long)B.job_ids[j] : (unsigned long long)j;
        unsigned long long st = (B.shuffle_seed + gid) + (unsigned long long)d * 0x632BE59BD9B4E019ULL + S.calls[d] * 0xD1342543DE82EF95ULL;
        S.calls[d] += 1;
        for (int r = 0; r < 4; ++r) S.d[d].perm_key[r] = (unsigned int)(bt_splitmix(st) >> 32);
        int bits = 2; while ((1ull << bits) < (unsigned long long)S.d[d].n_in) ++bits;
        if (bits & 1) ++bits;
        S.d[d].perm_half = bits / 2;
        S.d[d].M = min(S.number_points[d], S.d[d].n_in);
        for (int r = 0; r < 4; ++r) S.d2[d].perm_key[r] = S.d[d].perm_key[r];
        S.d2[d].perm_half = S.d[d].perm_half; S.d2[d].M = S.d[d].M;
        // epoch: (registration number of this block, iteration) — unique and growing over the life of the workspace (stamps are never cleared)
        const unsigned long long epoch = ((unsigned long long)S.seq << 8) | (unsigned long long)(S.it + 1);
        S.d[d].tag = (~epoch) << 32; S.d2[d].tag = S.d[d].tag;
      }
      __syncthreads();
      TPH(1);
      for (int d = 0; d < 2; ++d) {
        lrr_phase_match(scope, S.d[d], flags, S.ph); TPH(2);
        if (B.g8) lrr_phase_match_queue_g8(scope, S.d[d], S.d2[d], flags, S.ph, S.rows);
        lrr_phase_match_hard(scope, B.g8 ? S.d2[d] : S.d[d], flags, S.ph); TPH(3);
      }
      // estimate_LS: scales (MUSE on the first call), normalisation, weights, order statistics, normal equations, solve
      int mode[2] = {0, 0};
      for (int d = 0; d < 2; ++d) if (!S.scale_estimated[d]) {
        const double sc = block_muse(S.d[d], B, S.cand, S.seg);
        if (threadIdx.x == 0) { S.d[d].sc->scale = sc; S.scale_estimated[d] = 1; }
        mode[d] = 1;
      }
      __syncthreads();
      TPH(4);
      lrr_phase_norm1(scope, S.d[0], S.ph);
      lrr_phase_norm2(scope, S.d[0], Bk, S.ph);
      TPH(5);
      for (int d = 0; d < 2; ++d) { lrr_phase_weight(scope, S.d[d], flags, 0, nosel, nosel, mode[d], 0.0, S.ph); scope.phase_end(); }
      TPH(6);
      for (int d = 0; d < 2; ++d) {
        int ok = 1;
        const double G = block_mad(S.d[d].gkey, S.d[d].M, (double)BATCH_SEL_BINS / LRR_STRENGTH_CAP, S.hist, S.cand, S.si, S.sd, &ok);
        const double Sp = block_mad(S.d[d].skey, S.d[d].M, (double)BATCH_SEL_BINS, S.hist, S.cand, S.si, S.sd, &ok);
        if (threadIdx.x == 0) { S.d[d].sc->G = G; S.d[d].sc->S = Sp; if (!ok) S.status = 3; }
      }
      __syncthreads();
      TPH(7);
      const double jc[3] = {0, 0, 0};
      for (int d = 0; d < 2; ++d) lrr_phase_accumulate(scope, S.d[d], flags, d == 1, 0, S.d[d].sc->G, S.d[d].sc->S, jc, S.ph);
      __syncthreads();
      TPH(8);
      lrr_solve(F, Bk, flags, 12, 1);
      scope.phase_end();
      TPH(9);
      lrr_phase_resid(scope, S.d[0], S.ph);
      TPH(10);
      // iteration control (rrl sym :374-437) and validTransform (:2914-2926)
      if (threadIdx.x == 0) {
        const double after = F->res_sum_w != 0 ? F->res_sum_wr / F->res_sum_w : F->res_sum_wr;
        const double before = S.werr_before;
        const double curr = after - before, diff = (after - before) / after;
        bool converged = fabs(diff) < 1e-4;
        if (S.it > 0) {
          if (!S.finest && (S.osc > 1 || fabs(diff) < 0.01)) converged = true;
          else if (!converged) {
            if (after > 1.5) {
              if (curr * S.err_diff < 0.0) ++S.osc; else if (S.osc > 0) --S.osc;
              if (S.osc > 10) converged = true;
            } else {
              if (curr > 0.0) ++S.osc;
              if (S.osc > 3) converged = true;
            }
          }
        }
        S.err_diff = curr;
        if (converged && !S.finest) {
          converged = false;
          S.number_matches *= 2;
          bool fmax, bmax;
          if ((int)S.number_matches >= S.d[0].n_in) { S.number_points[0] = S.d[0].n_in; fmax = true; } else { S.number_points[0] = (int)S.number_matches; fmax = false; }
          if ((int)S.number_matches >= S.d[1].n_in) { S.number_points[1] = S.d[1].n_in; bmax = true; } else { S.number_points[1] = (int)S.number_matches; bmax = false; }
          S.finest = fmax && bmax;
          S.osc = 0; S.err_diff = 0.0;
        }
        S.werr_before = after;
        S.it += 1;
        S.converged = converged && S.finest;
        double U[9], V[9], w3[3];
        lrr_jacobi_svd<3>(3, F->A, U, w3, V);
        for (int d = 0; d < 3; ++d) if (!(w3[d] <= 10.0 && w3[d] >= 0.1)) S.status = 2;
      }
      __syncthreads();
    }
    // ---- out of iterations for this slice and not finished: park the job ----
    parked = B.quantum > 0 && !S.converged && S.it < B.max_iter && S.status == 0;
    if (parked) {
      __threadfence();    // every thread's writes into the slot are visible device-wide before the slot goes on the queue
      __syncthreads();
      if (threadIdx.x == 0) {
        C->job = j; C->n_in[0] = n_in[0]; C->n_in[1] = n_in[1];
        C->converged = S.converged; C->status = S.status; C->it = S.it; C->finest = S.finest; C->osc = S.osc; C->number_matches = S.number_matches;
        C->err_diff = S.err_diff; C->werr_before = S.werr_before;
        for (int d = 0; d < 2; ++d) { C->calls[d] = S.calls[d]; C->number_points[d] = S.number_points[d]; C->scale_estimated[d] = S.scale_estimated[d]; }
        C->slices += 1; C->t_parked_at = batch_now();
        __threadfence();
        sched_park(B, (int)blk);
      }
      __syncthreads();
    } else {
    // ---- result ----
    for (int i = threadIdx.x; i < 144; i += blockDim.x) R->cov[i] = F->cov[i];
    if (threadIdx.x == 0) {
      R->status = S.status == 3 ? -2 : S.status;      // -2: selector overflow, the host path takes the job
      R->converged = S.converged; R->oscillated = S.osc > 10; R->iterations = S.it;
      C->slices += 1; C->t_last = batch_now(); C->done_job = j; C->done_it = S.it;
      for (int i = 0; i < 9; ++i) R->A[i] = F->A[i];
      for (int i = 0; i < 3; ++i) { R->t[i] = F->t[i]; R->c[i] = F->c[i]; }
      R->rms = F->res_sum_r2;                          // rrl RMS_error: the plain sum (Q1)
      R->weighted_error = F->res_sum_w != 0 ? F->res_sum_wr / F->res_sum_w : F->res_sum_wr;
      if (D) { D->ran = 1; D->verified = (S.status == 0 && S.converged && !(S.osc > 10)) ? 1 : 0; }
    }
    __syncthreads();
    if (S.status == 0 && S.converged && !(S.osc > 10)) unit_verified = true;
    // ---- ComputeDecisionMeasurements (rrl/itkLocationRegistration.cxx:1816-1871), only for converged, non-oscillating runs (:2495-2507) ----
    if (B.mode == 2 && D && S.status == 0 && S.converged && !(S.osc > 10)) {
      const double deg = 180.0 / 3.14159265358979323846;
      // forward: angles and weighted error of the final matches, then the sampled transfer-error covariance
      for (int which = 0; which < 2; ++which) {
        lrr_phase_angles_rrl(scope, S.d[0], which, S.ph);
        if (threadIdx.x == 0) (which == 0 ? D->sheet_angles : D->tube_angles)[0] = (F->ang_num / F->ang_den) * deg;
        __syncthreads();
      }
      if (threadIdx.x == 0) D->weighted_error[0] = F->res_sum_w != 0 ? F->res_sum_wr / F->res_sum_w : F->res_sum_wr;
      { const double mt = block_max_trace(F, job.roi_center_moving, job.roi_half_size, B.spacing, S.sd); if (threadIdx.x == 0) D->max_trace[0] = mt; }
      // estimate_LS_backward( false ) = estimate_LS with forward_ / backward_ swapped (rrl sym :922-934): stale weights -> residuals
      // under the current transforms first, scale from the current weights (cdcl :698-701), the backward direction defines the
      // normalisation, its list is the "forward" one of the accumulation
      for (int d = 0; d < 2; ++d) lrr_phase_resid(scope, S.d[d], S.ph);
      lrr_phase_norm1(scope, S.d[1], S.ph);
      lrr_phase_norm2(scope, S.d[1], F, S.ph);
      for (int d = 0; d < 2; ++d) { lrr_phase_weight(scope, S.d[d], flags | LRR_F_USE_OLD_W, 0, nosel, nosel, 0, 0.0, S.ph); scope.phase_end(); }
      for (int d = 0; d < 2; ++d) {
        int ok = 1;
        const double G = block_mad(S.d[d].gkey, S.d[d].M, (double)BATCH_SEL_BINS / LRR_STRENGTH_CAP, S.hist, S.cand, S.si, S.sd, &ok);
        const double Sp = block_mad(S.d[d].skey, S.d[d].M, (double)BATCH_SEL_BINS, S.hist, S.cand, S.si, S.sd, &ok);
        if (threadIdx.x == 0) { S.d[d].sc->G = G; S.d[d].sc->S = Sp; if (!ok) S.status = 3; }
      }
      __syncthreads();
      const double jc0[3] = {0, 0, 0};
      for (int d = 0; d < 2; ++d) lrr_phase_accumulate(scope, S.d[d], flags, d == 0, 0, S.d[d].sc->G, S.d[d].sc->S, jc0, S.ph);
      __syncthreads();
      lrr_solve(Bk, F, flags, 12, 0);
      scope.phase_end();
      { const double mt = block_max_trace(Bk, job.roi_center_fixed, job.roi_half_size, B.spacing, S.sd); if (threadIdx.x == 0) D->max_trace[1] = mt; }
      for (int which = 0; which < 2; ++which) {
        lrr_phase_angles_rrl(scope, S.d[1], which, S.ph);
        if (threadIdx.x == 0) (which == 0 ? D->sheet_angles : D->tube_angles)[1] = (Bk->ang_num / Bk->ang_den) * deg;
        __syncthreads();
      }
      lrr_phase_resid(scope, S.d[1], S.ph);
      if (threadIdx.x == 0) {
        D->weighted_error[1] = Bk->res_sum_w != 0 ? Bk->res_sum_wr / Bk->res_sum_w : Bk->res_sum_wr;
        const double* q[4] = {D->weighted_error, D->max_trace, D->sheet_angles, D->tube_angles};
        for (int k = 0; k < 4; ++k) { D->measurements[k] = fmax(q[k][0], q[k][1]); D->measurements[4 + k] = fmin(q[k][0], q[k][1]); }
        D->measured = S.status == 0 ? 1 : 0;
        if (S.status == 3) R->status = -2;
      }
      __syncthreads();
    }
    }   // not parked
    }   // mode
    }   // ROI accepted
    __syncthreads();
    if (B.quantum > 0 && !parked) { __threadfence(); if (threadIdx.x == 0) sched_release(B, (int)blk); }   // the job is done (or was rejected): its slot is free again
    __syncthreads();
    TPH(11);
   } while (B.order && cpos < B.ncand && !(B.until_verified && unit_verified));
  }
  if (B.dbg && blockIdx.x == 0 && threadIdx.x == 0) for (int i = 0; i < 16; ++i) B.dbg[i] = S.tph[i];
}

// ---- host side -------------------------------------------------------------------------------------------------
size_t lrr_batch_smem_bytes() { return ((sizeof(BatchShared) + 15) & ~size_t(15)) + (size_t)BATCH_CAP * sizeof(int); }

// extra arguments of the candidate-loop launches (lrr_batch_run_locations); mode == 0: a plain job list.  thread_local: one host
// thread drives one context
thread_local BatchLocArgs lrr_batch_loc_args = {0, 1, 0, nullptr, nullptr, 1.0};
// Expected cost of a job before it runs = the features its two ROI cubes cover, counted by whole grid cells (run bounds of the
// cell-sorted arrays: one pair of loads per grid row, no point is touched); a cube that cannot reach the 1000-feature gate costs
// nothing.  One warp per job.  Used to hand the jobs out longest-expected-first (lrr_capi.cu): the makespan of a run is set by the
// jobs that are still running when the queue is empty, and a job's time is (iterations) x (ROI points).
__device__ __forceinline__ int warp_roi_cells(const GridDev& g, const double* c, double half) {
  const int x0 = lrr_cell_clamp(c[0] - half, g.ox, g.inv_h, g.nx), x1 = lrr_cell_clamp(c[0] + half, g.ox, g.inv_h, g.nx);
  const int y0 = lrr_cell_clamp(c[1] - half, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(c[1] + half, g.oy, g.inv_h, g.ny);
  const int z0 = lrr_cell_clamp(c[2] - half, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(c[2] + half, g.oz, g.inv_h, g.nz);
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  int cnt = 0;
  for (int r = threadIdx.x & 31; r < nrows; r += 32) {
    const int row = ((z0 + r / ny) * g.ny + (y0 + r % ny)) * g.nx;
    cnt += __ldg(g.cell_start + row + x1 + 1) - __ldg(g.cell_start + row + x0);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  return cnt;
}
__global__ void k_roi_cost(const FsetDev moving_all, const FsetDev fixed_all, const lrr_job* __restrict__ jobs, long long njobs, int* __restrict__ cost) {
  const long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= njobs) return;
  const lrr_job& job = jobs[j];
  const int cm = warp_roi_cells(moving_all.grid, job.roi_center_moving, job.roi_half_size);
  const int cf = warp_roi_cells(fixed_all.grid, job.roi_center_fixed, job.roi_half_size);
  if ((threadIdx.x & 31) == 0) cost[j] = (cm < 1000 || cf < 1000) ? 0 : cm + cf;
}
void lrr_batch_roi_cost(const FsetDev& moving_all, const FsetDev& fixed_all, const lrr_job* jobs_dev, long long njobs, int* cost_dev, cudaStream_t st, long long* launches) {
  if (njobs <= 0) return;
  k_roi_cost<<<(unsigned)((njobs + 3) / 4), 128, 0, st>>>(moving_all, fixed_all, jobs_dev, njobs, cost_dev); ++*launches;
}

cudaError_t lrr_batch_launch(const FsetDev& moving_all, const FsetDev& fixed_all, const lrr_job* jobs_dev, lrr_result* out_dev, long long njobs,
                             const long long* job_ids_dev, int max_iter, unsigned long long shuffle_seed, const double* muse_table, int muse_rows, int nblocks, void* workspace,
                             size_t workspace_bytes, cudaStream_t st, long long* launches, int nslots, int quantum) {
  if (nslots < nblocks) nslots = nblocks;
  BatchArgs B; memset(&B, 0, sizeof(B));
  B.all[0] = moving_all; B.all[1] = fixed_all; B.jobs = jobs_dev; B.out = out_dev; B.njobs = njobs; B.max_iter = max_iter; B.shuffle_seed = shuffle_seed;
  B.job_ids = job_ids_dev;
  B.mode = 0; B.ncand = 1; B.until_verified = 0; B.order = nullptr; B.dec = nullptr; B.spacing = 1.0; B.nunits = njobs;
  if (lrr_batch_loc_args.mode) {
    B.mode = lrr_batch_loc_args.mode; B.ncand = lrr_batch_loc_args.ncand; B.until_verified = lrr_batch_loc_args.until_verified;
    B.order = lrr_batch_loc_args.order; B.dec = lrr_batch_loc_args.dec; B.spacing = lrr_batch_loc_args.spacing;
    if (B.order) B.nunits = njobs / B.ncand;
  }
  { const char* e = getenv("LRR_BATCH_G8"); B.g8 = e ? atoi(e) : 1; const char* r = getenv("LRR_BATCH_BALL_ROWS"); B.ball_rows = r ? atoi(r) : (B.g8 ? 16 : 0); }
  for (int i = 0; i < 4 * 32; ++i) B.muse_table[i] = muse_table[i];
  B.muse_rows = muse_rows;
  // carve the workspace
  char* p = (char*)workspace; const size_t nb = (size_t)nslots;   // workspace slots (>= blocks)
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~size_t(255); return r; };
  B.next_job = (int*)take(256);
  B.dbg = getenv("LRR_BATCH_DEBUG") ? (long long*)((char*)B.next_job + 64) : nullptr;
  B.w_sc = (Scal*)take(nb * 2 * sizeof(Scal));
  B.w_f = (double*)take(nb * 2 * 4 * BATCH_CAP * sizeof(double));
  B.w_fP6 = (double*)take(nb * 2 * 6 * BATCH_CAP * sizeof(double));
  B.w_pm = (double*)take(nb * 2 * 4 * BATCH_CAP * sizeof(double));
  B.w_hqb = (double*)take(nb * 2 * 2 * BATCH_CAP * sizeof(double));
  B.w_forig = (int*)take(nb * 2 * BATCH_CAP * sizeof(int));
  B.w_sidx = (int*)take(nb * 2 * BATCH_CAP * sizeof(int));
  B.w_prev = (int*)take(nb * 2 * BATCH_CAP * sizeof(int));
  B.w_hq = (int*)take(nb * 2 * 2 * (3 * BATCH_CAP + 4) * sizeof(int));
  B.w_stamp0 = (unsigned long long*)take(nb * (size_t)fixed_all.n * 8);
  B.w_stamp1 = (unsigned long long*)take(nb * (size_t)moving_all.n * 8);
  B.ctl = (BatchCtl*)take(nb * sizeof(BatchCtl));
  B.sched = (int*)take((8 + 2 * nb) * sizeof(int));
  if ((size_t)(p - (char*)workspace) > workspace_bytes) return cudaErrorMemoryAllocation;
  // time slicing: plain job lists with more slots than blocks (nslots = the workspace slots the caller sized the workspace for)
  B.nslots = nslots; B.quantum = (B.mode == 0 && !B.order && nslots > nblocks) ? quantum : 0;
  cudaError_t e;
  if ((e = cudaMemsetAsync(B.ctl, 0, nb * sizeof(BatchCtl), st)) != cudaSuccess) return e;
  {   // scheduler state: unlocked, first fresh unit 0, every slot free, empty queue
    std::vector<int> h(8 + 2 * nb, 0);
    h[2] = nslots; for (int i = 0; i < nslots; ++i) h[8 + i] = nslots - 1 - i;
    if ((e = cudaMemcpyAsync(B.sched, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;   // (h is a local buffer)
  }
  if ((e = cudaMemsetAsync(B.next_job, 0, 256, st)) != cudaSuccess) return e;
  // stamps: all ones = "older than any epoch"
  if ((e = cudaMemsetAsync(B.w_stamp0, 0xFF, nb * (size_t)fixed_all.n * 8, st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(B.w_stamp1, 0xFF, nb * (size_t)moving_all.n * 8, st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(B.w_sc, 0, nb * 2 * sizeof(Scal), st)) != cudaSuccess) return e;
  const size_t smem = lrr_batch_smem_bytes();
  if ((e = cudaFuncSetAttribute(k_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  const bool dbg_time = getenv("LRR_BATCH_DEBUG") != nullptr;
  if (dbg_time) {
    const auto c0 = std::chrono::steady_clock::now(); cudaStreamSynchronize(st);
    fprintf(stderr, "[batch] clears of the workspace drained in %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
    cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, st);
  }
  k_batch<<<nblocks, BATCH_THREADS, smem, st>>>(B); ++*launches;
  if (dbg_time) { cudaEventRecord(ev1, st); cudaEventSynchronize(ev1); float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1); fprintf(stderr, "[batch] k_batch %.2f ms (events), %d blocks, %d slots, quantum %d\n", ms, nblocks, nslots, B.quantum); cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
  if (getenv("LRR_BATCH_DEBUG") && atoi(getenv("LRR_BATCH_DEBUG")) == 2) {   // the jobs that finished last: when they started, how long they waited parked
    cudaStreamSynchronize(st);
    std::vector<BatchCtl> hc((size_t)nslots); cudaMemcpy(hc.data(), B.ctl, hc.size() * sizeof(BatchCtl), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t1 = 0; for (auto& c : hc) if (c.t_first) { t0 = std::min(t0, c.t_first); t1 = std::max(t1, c.t_last); }
    std::vector<int> idx; for (int i = 0; i < nslots; ++i) if (hc[(size_t)i].t_last) idx.push_back(i);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return hc[(size_t)a].t_last > hc[(size_t)b].t_last; });
    fprintf(stderr, "[batch] quantum %d slots %d makespan %.1f ms; last finishers (last tenant of the slot):\n", B.quantum, nslots, (t1 - t0) * 1e-6);
    for (size_t k = 0; k < idx.size() && k < 6; ++k) { const BatchCtl& c = hc[(size_t)idx[k]];
      fprintf(stderr, "  job %d its %d start %.1f end %.1f ms slices %d parked %.1f ms\n", c.done_job, c.done_it, (c.t_first - t0) * 1e-6, (c.t_last - t0) * 1e-6, c.slices, c.t_waited * 1e-6); }
  }
  if (B.dbg) {
    long long h[16]; cudaStreamSynchronize(st); cudaMemcpy(h, B.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[12] = {"roi", "ctl", "matchA", "matchB", "muse", "norm", "weight", "mad", "accum", "solve", "resid", "result"};
    for (int i = 0; i < 12; ++i) fprintf(stderr, "[batch block 0] %-7s %10.1f us\n", nm[i], h[i] / 1.9e3);
  }
  return cudaGetLastError();
}

size_t lrr_batch_workspace_bytes(int nblocks, int n_moving, int n_fixed) {
  const size_t nb = (size_t)nblocks;
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  size_t t = up(256) + up(nb * 2 * sizeof(Scal)) + 2 * up(nb * 2 * 4 * BATCH_CAP * 8) + up(nb * 2 * 6 * BATCH_CAP * 8) + up(nb * 2 * 2 * BATCH_CAP * 8) +
             3 * up(nb * 2 * BATCH_CAP * 4) + up(nb * 2 * 2 * (3 * BATCH_CAP + 4) * 4) + up(nb * (size_t)n_fixed * 8) + up(nb * (size_t)n_moving * 8) +
             up(nb * sizeof(BatchCtl)) + up((8 + 2 * nb) * sizeof(int));
  return t + 4096;
}
int lrr_batch_blocks_per_sm() {
  int n = 0;
  cudaFuncSetAttribute(k_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lrr_batch_smem_bytes());
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_batch, BATCH_THREADS, lrr_batch_smem_bytes());
  return n < 1 ? 1 : n;
}
