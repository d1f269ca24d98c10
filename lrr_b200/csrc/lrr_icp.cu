// K1/K3/K4/K5: one kernel per phase of the robust ICP iteration (sm_100a).  The phase bodies live in
// lrr_phases.cuh; this file instantiates them over the whole GPU (GridScope), one grid row per direction of
// a symmetric estimator, and adds the cross-direction finalisation (merged order statistics, joint solve).
#include "lrr_phases.cuh"
#include <algorithm>
#include <unordered_map>
#include <cstdlib>

static inline int nblk_for(long long items) {
  long long b = (items + LRR_THREADS - 1) / LRR_THREADS;
  long long cap = (long long)lrr_num_sms() * 4;
  if (cap > LRR_MAX_BLOCKS) cap = LRR_MAX_BLOCKS;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
// Grid width such that ALL blocks of the launch (gridDim.y = directions) are resident at once.  Measured with per-block
// %globaltimer marks (scripts/timeline.py): with more blocks than fit, the blocks of direction 1 (scheduled after those of
// direction 0) only start when direction 0 has finished, and every phase kernel ends in a last-block reduction — the two
// directions ran back to back, each with its own tail.
template <class K>
static inline int resident_width(K kernel, int threads, int ny, long long items) {
  // (occupancy per (device, kernel); thread_local: contexts on different devices may be driven from different host threads)
  static thread_local std::unordered_map<unsigned long long, int> cache;
  const unsigned long long key = (unsigned long long)(uintptr_t)(const void*)kernel ^ ((unsigned long long)lrr_current_device() << 56);
  int per_sm;
  auto it = cache.find(key);
  if (it != cache.end()) per_sm = it->second;
  else {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
    cache[key] = per_sm = n;
  }
  long long b = (items + threads - 1) / threads;
  long long cap = (long long)lrr_num_sms() * per_sm / (ny < 1 ? 1 : ny);
  if (cap > LRR_MAX_BLOCKS) cap = LRR_MAX_BLOCKS;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static inline int maxM(const EstArgs& a) { int m = a.d[0].M; if (a.ndir == 2 && a.d[1].M > m) m = a.d[1].M; return m; }

__device__ __forceinline__ GridScope lrr_scope(const EstArgs& a, int dir) {
  GridScope s; s.partials = a.partials + (size_t)dir * LRR_MAX_BLOCKS * LRR_NRED; s.ticket = a.ticket + dir; return s;
}
// true in every thread of the block that completes the pair of directions (or at once when ndir == 1)
__device__ __forceinline__ bool lrr_pair_done(const EstArgs& a, int needed) {
  if (needed <= 1) return true;
  __shared__ bool s_second;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(a.ticket + 2, 1u);
    s_second = (t == (unsigned int)needed - 1);
    if (s_second) a.ticket[2] = 0;
  }
  __syncthreads();
  if (s_second) __threadfence();
  return s_second;
}

// Speculative plan (lrr_capi.cu): the host left out the staged search, pass B and the normalisation because the previous iteration
// kept every query.  True when that bet is lost for this iteration: nothing may be committed, the host completes the match and
// reruns.  Every kernel of the plan evaluates the same device scalars, which none of them writes.
__device__ __forceinline__ bool lrr_spec_bad(const EstArgs& a) {
  bool bad = a.d[a.src].sc->norm_dirty != 0;
  for (int d = 0; d < a.ndir; ++d) bad = bad || a.d[d].sc->match_left != 0;
  return bad;
}

// ---- upload-time helpers ---------------------------------------------------------------------------------
__global__ void k_fill_records(FsetDev f, double* __restrict__ rec, double* __restrict__ crec) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f.n) return;
  const int o = lrr_pos4_orig(f.grid.pos4[j]);
  double* r = rec + 8 * (size_t)j;
#pragma unroll
  for (int p = 0; p < 6; ++p) r[p] = f.P6[(size_t)p * f.n + o];
  const double s = f.s[o];
  r[6] = s; r[7] = (double)(float)s;
  if (crec && f.C6) {
    double* c = crec + 8 * (size_t)j;
#pragma unroll
    for (int p = 0; p < 6; ++p) c[p] = f.C6[(size_t)p * f.n + o];
    c[6] = 0.0; c[7] = 0.0;
  }
}
void lrr_k_fill_records(const FsetDev& f, double* rec, double* crec, cudaStream_t st, long long* launches) {
  if (f.n == 0) return;
  k_fill_records<<<(f.n + 255) / 256, 256, 0, st>>>(f, rec, crec); ++*launches;
}

// Error projector of a planar cdcl_feature: outer product of the eigenvector of the smallest eigenvalue of its 2x2 covariance = the
// point normal (cdcl/cdcl_estimation_ICP.txx:470-477, re-derived there at every use; constant per feature, so computed once here).
// Closed form of the symmetric 2x2 eigenproblem; an isotropic covariance (no defined normal) gives the x axis.
__global__ void k_projector2_from_cov(const double* __restrict__ cov3, int n, double* __restrict__ P6, double* __restrict__ C6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = cov3[i], b = cov3[(size_t)n + i], c = cov3[2 * (size_t)n + i];
  const double half = __dmul_rn(0.5, a - c), rad = sqrt(__dadd_rn(__dmul_rn(half, half), __dmul_rn(b, b)));
  const double lam = __dmul_rn(0.5, a + c) - rad;
  double v0, v1;
  if (a >= c) { v0 = b; v1 = lam - a; } else { v0 = lam - c; v1 = b; }
  const double n2 = __dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1));
  double pxx = 1.0, pxy = 0.0, pyy = 0.0;
  if (n2 > 0.0) {
    const double inv = 1.0 / sqrt(n2);
    v0 = __dmul_rn(v0, inv); v1 = __dmul_rn(v1, inv);
    pxx = __dmul_rn(v0, v0); pxy = __dmul_rn(v0, v1); pyy = __dmul_rn(v1, v1);
  }
  const size_t N = (size_t)n;
  P6[i] = pxx; P6[N + i] = pxy; P6[2 * N + i] = 0.0; P6[3 * N + i] = pyy; P6[4 * N + i] = 0.0; P6[5 * N + i] = 0.0;
  C6[i] = a; C6[N + i] = b; C6[2 * N + i] = 0.0; C6[3 * N + i] = c; C6[4 * N + i] = 0.0; C6[5 * N + i] = 0.0;
}
void lrr_k_projector2_from_cov(const double* cov3, int n, double* P6, double* C6, cudaStream_t st, long long* launches) {
  if (n == 0) return;
  k_projector2_from_cov<<<(n + 255) / 256, 256, 0, st>>>(cov3, n, P6, C6); ++*launches;
}

__global__ void k_gather_from(FsetDev src, const int* __restrict__ idx, int n_in, double* __restrict__ fx, double* __restrict__ fy,
                              double* __restrict__ fz, double* __restrict__ fs, double* __restrict__ fP6) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  const int o = idx ? idx[i] : i;
  fx[i] = src.x[o]; fy[i] = src.y[o]; fz[i] = src.z[o]; fs[i] = src.s[o];
#pragma unroll
  for (int p = 0; p < 6; ++p) fP6[(size_t)p * n_in + i] = src.P6[(size_t)p * src.n + o];
}
void lrr_k_gather_from(const FsetDev& src, const int* idx, int n_in, double* fx, double* fy, double* fz, double* fs, double* fP6,
                       cudaStream_t st, long long* launches) {
  if (n_in == 0) return;
  k_gather_from<<<(n_in + 255) / 256, 256, 0, st>>>(src, idx, n_in, fx, fy, fz, fs, fP6); ++*launches;
}

// processing order of the match phase: compact indices sorted by the from-set's own cell order, so that the
// 32 queries of a warp are spatial neighbours (same rows of the to-grid, shared cache lines, converged loops)
__global__ void k_proc_mark(FsetDev f, const int* __restrict__ forig, int n_in, int* __restrict__ mark) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  const int o = forig ? forig[i] : i;
  mark[f.sorted_of_orig[o]] = i + 1;   // duplicates in the subset: the last writer wins, the others are appended below
}
__global__ void __launch_bounds__(1024) k_proc_compact(const int* __restrict__ mark, int n_all, int n_in, int* __restrict__ proc, int* __restrict__ seen) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (n_all + T - 1) / T, lo = min(t * per, n_all), hi = min(lo + per, n_all);
  int cnt = 0;
  for (int j = lo; j < hi; ++j) cnt += mark[j] != 0;
  s_part[t] = cnt;
  __syncthreads();
  for (int o = 1; o < T; o <<= 1) { int v = (t >= o) ? s_part[t - o] : 0; __syncthreads(); s_part[t] += v; __syncthreads(); }
  int pos = s_part[t] - cnt;
  for (int j = lo; j < hi; ++j) { const int m = mark[j]; if (m) { proc[pos++] = m - 1; seen[m - 1] = 1; } }
  __syncthreads();
  // subset entries that lost a duplicate slot (same feature listed twice) keep their turn at the end
  if (t == 0) { int p = s_part[T - 1]; if (p < n_in) for (int i = 0; i < n_in; ++i) if (!seen[i]) proc[p++] = i; }
}
void lrr_k_build_proc(const FsetDev& from_all, const int* forig, int n_in, int* mark, int* proc, cudaStream_t st, long long* launches) {
  if (n_in == 0) return;
  cudaMemsetAsync(mark, 0, sizeof(int) * ((size_t)from_all.n + n_in), st);
  k_proc_mark<<<(n_in + 255) / 256, 256, 0, st>>>(from_all, forig, n_in, mark); ++*launches;
  k_proc_compact<<<1, 1024, 0, st>>>(mark, from_all.n, n_in, proc, mark + from_all.n); ++*launches;
}

// ---- phase kernels -----------------------------------------------------------------------------------------
#ifndef LRR_MATCH_MINBLK
#define LRR_MATCH_MINBLK 2
#endif
__global__ void __launch_bounds__(LRR_THREADS, LRR_MATCH_MINBLK) k_match(const __grid_constant__ EstArgs a) {
  lrr_pdl_enter();
  __shared__ PhaseShared sh;
  const int dir = blockIdx.y;
  lrr_tl(0, 0);
  lrr_phase_match(lrr_scope(a, dir), a.d[dir], a.flags, sh, a.ticket + 4 + dir);
  lrr_tl(0, 3);
}
#ifndef LRR_HARD_MINBLK
#define LRR_HARD_MINBLK 2
#endif
// Pass B of every direction in ONE 1-D grid that is resident as a whole: each warp drains the queue of "its" direction
// (blockIdx.x parity) and then helps with the other one, so both queues have the whole GPU and neither waits for a second wave
// (measured with per-block timeline marks: as gridDim.y = 2 the blocks of direction 1 only started when direction 0 had finished,
// two serial tails of ~15 us each when the queues hold a few hundred far queries).  One reduction of 2 x 4 sums, one finalizer.
__global__ void __launch_bounds__(LRR_THREADS, LRR_HARD_MINBLK) k_match_hard(const __grid_constant__ EstArgs a) {
  lrr_pdl_enter();
  __shared__ PhaseShared sh;
  __shared__ double s_T[2][15];
  __shared__ int s_rows[LRR_NWARPS][2 * (LRR_FLAT_ROWS + 1)];   // row tables of the flattened ball scan, one per warp
  int* const rowtab = (a.flags & LRR_F_NOFLAT) ? nullptr : s_rows[threadIdx.x >> 5];
  lrr_tl(1, 0);
  {
    // nothing queued in any direction (the steady state: every query kept by its certificate or finished by the staged search):
    // one thread per direction does the bookkeeping, every other block leaves at once
    bool any = false;
    for (int d = 0; d < a.ndir; ++d) any = any || *reinterpret_cast<volatile int*>(a.d[d].hq_count) > 0;
    if (!any) {
      if (blockIdx.x == 0 && threadIdx.x < a.ndir) {
        const Scal* sc = a.d[threadIdx.x].sc;
        const double tot[4] = {sc->sum_loc[0], sc->sum_loc[1], sc->sum_loc[2], sc->sum_r2};
        lrr_match_finalize(a.d[threadIdx.x], a.flags, tot);
      }
      return;
    }
  }
  if (threadIdx.x < 15 * a.ndir) {
    const int d = threadIdx.x / 15, j = threadIdx.x % 15;
    const Scal* sc = a.d[d].sc;
    s_T[d][j] = j < 9 ? sc->A[j] : (j < 12 ? sc->t[j - 9] : sc->c[j - 12]);
  }
  __syncthreads();
  double acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
  const int wg = blockIdx.x * LRR_NWARPS + (threadIdx.x >> 5), nwg = gridDim.x * LRR_NWARPS;
  if (a.ndir == 2) {
    if (blockIdx.x & 1) { lrr_match_b_body_dynamic(a.d[1], a.flags, s_T[1], a.ticket + 7, acc1, wg, nwg, rowtab); lrr_match_b_body_dynamic(a.d[0], a.flags, s_T[0], a.ticket + 6, acc0, wg, nwg, rowtab); }
    else { lrr_match_b_body_dynamic(a.d[0], a.flags, s_T[0], a.ticket + 6, acc0, wg, nwg, rowtab); lrr_match_b_body_dynamic(a.d[1], a.flags, s_T[1], a.ticket + 7, acc1, wg, nwg, rowtab); }
  } else lrr_match_b_body_dynamic(a.d[0], a.flags, s_T[0], a.ticket + 6, acc0, wg, nwg, rowtab);
  lrr_tl(1, 1);
  double v[8] = {acc0[0], acc0[1], acc0[2], acc0[3], acc1[0], acc1[1], acc1[2], acc1[3]};
  if (lrr_scope(a, 0).template reduce<8>(v, sh.out, sh.tmp)) {
    lrr_tl(1, 2);
    if (threadIdx.x < a.ndir) {
      const int d = threadIdx.x;
      const Scal* sc = a.d[d].sc;
      const double tot[4] = {sc->sum_loc[0] + sh.out[4 * d], sc->sum_loc[1] + sh.out[4 * d + 1], sc->sum_loc[2] + sh.out[4 * d + 2], sc->sum_r2 + sh.out[4 * d + 3]};
      lrr_match_finalize(a.d[d], a.flags, tot);
      a.ticket[6 + d] = 0;
    }
  }
  lrr_tl(1, 3);
}
// (the "tile" switch lives in lrr_ctx::Options: with it off dir_args withholds the chunk table and the tile path is not usable)
void lrr_k_match(const EstArgs& a, cudaStream_t st, long long* launches, bool cert_only) {
  if (cert_only) { lrr_k_match_cert(a, st, launches); return; }   // the caller has checked that every direction is certified
  if (lrr_match_tile_usable(a) && lrr_match_cert_usable(a)) lrr_k_match_cert(a, st, launches);
  lrr_k_match_rest(a, st, launches);
}
void lrr_k_match_rest(const EstArgs& a, cudaStream_t st, long long* launches) {
  lrr_k_match_search(a, st, launches);
  lrr_k_match_queue(a, st, launches);
}
// pass A: the staged search / sparse list of the finest level (lrr_tile.cu) or the thread-per-query ball scan of the ramp levels
void lrr_k_match_search(const EstArgs& a, cudaStream_t st, long long* launches) {
  // persistent-style grid: every resident warp pulls 32-query chunks until the direction's queries are exhausted
  if (lrr_match_tile_usable(a)) {   // finest level (lrr_tile.cu): shared-memory-staged cells for what the certificates did not keep
    lrr_k_match_tile(a, st, launches);
  }
  else {
    dim3 grid(std::min(nblk_for(maxM(a)), 2 * lrr_num_sms()), a.ndir);
    lrr_launch_pdl(k_match, grid, dim3(LRR_THREADS), 0, st, a); ++*launches;
  }
}
// pass B: one warp per queued query
void lrr_k_match_queue(const EstArgs& a, cudaStream_t st, long long* launches) {
  // queue length is only known on the device: a fixed grid of warps strides over it
  dim3 gridb(resident_width(k_match_hard, LRR_THREADS, 1, 32LL * maxM(a)));
  lrr_launch_pdl(k_match_hard, gridb, dim3(LRR_THREADS), 0, st, a); ++*launches;
}

// Certified finest level (DirArgs::stamp_in_match == 0): no match kernel stamps there — a kept match is not touched at all — so the
// "unique matched to-features" stamps (the reference's std::set of smart pointers, cdcl :1173-1189) are laid down here, in a fresh
// epoch, right before the normalisation reads them; like the normalisation itself only when some match changed.
__global__ void __launch_bounds__(LRR_THREADS) k_restamp(const __grid_constant__ EstArgs a, int skip_clean) {
  lrr_pdl_enter();
  const DirArgs& d = a.d[a.src];
  if (skip_clean && !d.sc->norm_dirty) return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < d.M; k += gridDim.x * blockDim.x) {
    const int s = d.sidx[k];
    if (s >= 0) atomicMin(d.stamp + s, d.tag | (unsigned long long)(unsigned int)k);
  }
}
// cdcl symmetric rigid estimate_LS_backward (cdcl sym :708-731): the matches' from_ — clones mapped through the transform of the
// iteration's find_closest — are cloned and mapped AGAIN through the current global transform, which becomes the transform the
// increment is recomposed onto.
__global__ void k_map_clones_again(Scal* sc) {
  const int t = threadIdx.x;
  if (t < 9) sc->Ag2[t] = sc->A[t];
  else if (t < 12) sc->tg2[t - 9] = sc->t[t - 9];
  else if (t < 15) sc->cg2[t - 12] = sc->c[t - 12];
  else if (t == 15) sc->from_mapped = 2;
}
void lrr_k_map_clones_again(Scal* sc, cudaStream_t st, long long* launches) { k_map_clones_again<<<1, 32, 0, st>>>(sc); ++*launches; }
__global__ void k_fill_ones(double* __restrict__ w, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = 1.0;
}
void lrr_k_fill_ones(double* w, int n, cudaStream_t st, long long* launches) {
  if (n <= 0) return;
  k_fill_ones<<<(n + 255) / 256, 256, 0, st>>>(w, n); ++*launches;
}
void lrr_k_restamp(const EstArgs& a, cudaStream_t st, long long* launches, int skip_clean) {
  lrr_launch_pdl(k_restamp, dim3(nblk_for(a.d[a.src].M)), dim3(LRR_THREADS), 0, st, a, skip_clean); ++*launches;
}

// skip_clean: the host knows that the normalisation scalars on the device are those of this estimator's last estimate_LS in the
// same configuration (lrr_est::norm_token); if in addition no match of the source direction changed since (Scal::norm_dirty),
// they are what this launch would recompute bit for bit, and every block leaves at entry.
__global__ void __launch_bounds__(LRR_THREADS) k_norm1(const __grid_constant__ EstArgs a, int skip_clean) {
  lrr_pdl_enter();
  if (skip_clean && !a.d[a.src].sc->norm_dirty) return;
  __shared__ PhaseShared sh;
  lrr_tl(6, 0);
  lrr_phase_norm1(lrr_scope(a, a.src), a.d[a.src], sh);
  lrr_tl(6, 3);
}
void lrr_k_norm1(const EstArgs& a, cudaStream_t st, long long* launches, int skip_clean) {
  lrr_launch_pdl(k_norm1, dim3(nblk_for(a.d[a.src].M)), dim3(LRR_THREADS), 0, st, a, skip_clean); ++*launches;
}
__global__ void __launch_bounds__(LRR_THREADS) k_norm2(const __grid_constant__ EstArgs a, int skip_clean) {
  lrr_pdl_enter();
  if (skip_clean && !a.d[a.src].sc->norm_dirty) return;
  __shared__ PhaseShared sh;
  lrr_tl(7, 0);
  lrr_phase_norm2(lrr_scope(a, a.src), a.d[a.src], a.ndir == 2 ? a.d[1 - a.src].sc : nullptr, sh);
  lrr_tl(7, 3);
}
void lrr_k_norm2(const EstArgs& a, cudaStream_t st, long long* launches, int skip_clean) {
  lrr_launch_pdl(k_norm2, dim3(nblk_for(a.d[a.src].M)), dim3(LRR_THREADS), 0, st, a, skip_clean); ++*launches;
}

// ---- cdcl estimate_LS_backward (cdcl sym :705-863): cdcl_normalize_matches over the MERGED list = the estimated
// direction's own matches (src) plus the other direction's matches reversed (add_constraints, :744).  The reversed
// clones are distinct objects, so each of them counts in the "unique to" set; the own to-features count once (stamps).
// pass 0: centres, pass 1: mean radii.  Scratch: sc->mom[0..8).
__global__ void __launch_bounds__(LRR_THREADS) k_norm_merged(const __grid_constant__ EstArgs a, int pass) {
  __shared__ PhaseShared sh;
  const int dir = blockIdx.y;
  const DirArgs& d = a.d[dir];
  Scal* sc = d.sc;
  const bool own = dir == a.src;
  const GridScope scope = lrr_scope(a, dir);
  const int mapped = sc->from_mapped;
  // centres of the merged from / to sides in THIS direction's roles (own: from/to, reversed: to/from)
  const Scal* S = a.d[a.src].sc;
  const double cm0 = S->cfrom[0], cm1 = S->cfrom[1], cm2 = S->cfrom[2], cf0 = S->cto[0], cf1 = S->cto[1], cf2 = S->cto[2];
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = scope.first(); k < d.M; k += scope.stride()) {
    const int s = d.sidx[k];
    if (s < 0) continue;
    const int i = lrr_from_index(d, k);
    double px, py, pz; lrr_from_loc(d, sc, mapped, i, px, py, pz);
    const double4 q = lrr_ld_pos4(d.to.grid.pos4 + s);
    const bool rep = own ? (d.stamp[s] == (d.tag | (unsigned long long)(unsigned int)k)) : true;
    // merged "from" = own from / reversed to ; merged "to" = own to (unique) / reversed from (every clone)
    const double fx = own ? px : q.x, fy = own ? py : q.y, fz = own ? pz : q.z;
    const double tx = own ? q.x : px, ty = own ? q.y : py, tz = own ? q.z : pz;
    if (pass == 0) {
      acc[0] += fx; acc[1] += fy; acc[2] += fz; acc[6] += 1.0;
      if (rep) { acc[3] += tx; acc[4] += ty; acc[5] += tz; acc[7] += 1.0; }
    } else {
      const double ax = fx - cm0, ay = fy - cm1, az = fz - cm2;
      acc[0] += sqrt(ax * ax + ay * ay + az * az);
      if (rep) { const double bx = tx - cf0, by = ty - cf1, bz = tz - cf2; acc[1] += sqrt(bx * bx + by * by + bz * bz); }
    }
  }
  if (!scope.reduce<8>(acc, sh.out, sh.tmp)) return;
  if (threadIdx.x < 8) sc->mom[threadIdx.x] = sh.out[threadIdx.x];
  if (!lrr_pair_done(a, 2)) return;
  if (threadIdx.x == 0) {
    Scal* own_sc = a.d[a.src].sc; Scal* oth = a.d[1 - a.src].sc;
    double m[8];
    for (int j = 0; j < 8; ++j) m[j] = __ldcg(&own_sc->mom[j]) + __ldcg(&oth->mom[j]);
    if (pass == 0) {
      for (int j = 0; j < 3; ++j) { own_sc->cfrom[j] = m[j] / m[6]; own_sc->cto[j] = m[3 + j] / m[7]; }
      own_sc->n_unique = m[7]; own_sc->res_count = m[6];   // res_count: scratch for the merged list length until pass 1
    } else {
      own_sc->rfrom = m[0] / own_sc->res_count; own_sc->rto = m[1] / own_sc->n_unique;
      // the reversed list sees the same normalisation with the roles swapped
      for (int j = 0; j < 3; ++j) { oth->cfrom[j] = own_sc->cto[j]; oth->cto[j] = own_sc->cfrom[j]; }
      oth->rfrom = own_sc->rto; oth->rto = own_sc->rfrom;
    }
  }
}
void lrr_k_norm_merged(const EstArgs& a, cudaStream_t st, long long* launches) {
  dim3 grid(nblk_for(maxM(a)), 2);
  k_norm_merged<<<grid, LRR_THREADS, 0, st>>>(a, 0); ++*launches;
  k_norm_merged<<<grid, LRR_THREADS, 0, st>>>(a, 1); ++*launches;
}

struct WeightExtra { int scale_mode[2]; double scale_in[2]; };

__global__ void __launch_bounds__(LRR_THREADS) k_weight(const __grid_constant__ EstArgs a, const WeightExtra ex, int spec) {
  lrr_pdl_enter();
  if (spec && lrr_spec_bad(a)) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) for (int d = 0; d < a.ndir; ++d) atomicOr(&a.d[d].sc->status, LRR_ST_NOT_STEADY);
    return;
  }
  __shared__ PhaseShared sh;
  __shared__ double sh_cand[LRR_SEL_CAP], sh_cand2[LRR_SEL_CAP];
  const int dir = blockIdx.y;
  const bool merged = (a.flags & LRR_F_MERGED) != 0, window = (a.flags & LRR_F_BRACKET) != 0;
  const int group = merged ? 0 : dir;
  const int reversed = merged && dir != a.src;
  lrr_tl(4, 0);
  const bool fin = lrr_phase_weight(lrr_scope(a, dir), a.d[dir], a.flags, reversed, a.selG[group], a.selS[group], ex.scale_mode[dir],
                                    ex.scale_in[dir], sh);
  lrr_tl(4, 1);
  if (!fin) return;
  lrr_tl(4, 2);
  if (a.flags & LRR_F_ICP) {   // cdcl_estimation_ICP: the strength weights are normalised by their sum, no order statistic
    if (threadIdx.x == 0) { Scal* sc = a.d[dir].sc; sc->G = (a.flags & LRR_F_STRENGTH) ? sc->n_nonzero : 1.0; sc->S = 1.0; }
    return;
  }
  // group-level: the two order statistics (window mode: finished here; histogram mode: bins located for k_collect)
  const bool pair = merged && a.ndir == 2;
  if (pair && !lrr_pair_done(a, 2)) return;
  double nG, nS, bG, bS;
  if (pair) {
    nG = __ldcg(&a.d[0].sc->n_nonzero) + __ldcg(&a.d[1].sc->n_nonzero); nS = (double)(a.d[0].M + a.d[1].M);
    bG = __ldcg(&a.d[0].sc->sel_below[0]) + __ldcg(&a.d[1].sc->sel_below[0]); bS = __ldcg(&a.d[0].sc->sel_below[1]) + __ldcg(&a.d[1].sc->sel_below[1]);
  } else {
    nG = __ldcg(&a.d[dir].sc->n_nonzero); nS = (double)a.d[dir].M;
    bG = __ldcg(&a.d[dir].sc->sel_below[0]); bS = __ldcg(&a.d[dir].sc->sel_below[1]);
  }
  if (!window) {
    if (a.flags & LRR_F_STRENGTH) lrr_sel_locate(a.selG[group], nG, sh.tmp);
    if (LRR_SPATIAL(a.flags)) lrr_sel_locate(a.selS[group], nS, sh.tmp);
    return;
  }
  bool hit = true;
  {
    // both candidate lists into shared memory in one round trip (counts, then keys), then the two exact selections
    const bool useG = (a.flags & LRR_F_STRENGTH) != 0, useS = LRR_SPATIAL(a.flags);
    const unsigned int cG = useG ? __ldcg(a.selG[group].cand_count) : 0u, cS = useS ? __ldcg(a.selS[group].cand_count) : 0u;
    const unsigned int lG = cG <= LRR_SEL_CAP ? cG : 0u, lS = cS <= LRR_SEL_CAP ? cS : 0u;
    for (unsigned int i = threadIdx.x; i < lG; i += blockDim.x) sh_cand[i] = __ldcg(a.selG[group].cand + i);
    for (unsigned int i = threadIdx.x; i < lS; i += blockDim.x) sh_cand2[i] = __ldcg(a.selS[group].cand + i);
    if (useG) hit = lrr_sel_window_finish(a.selG[group], nG, bG, sh_cand, (int)cG) && hit;
    if (useS) hit = lrr_sel_window_finish(a.selS[group], nS, bS, sh_cand2, (int)cS) && hit;
  }
  if (threadIdx.x == 0) {
    const double G = (a.flags & LRR_F_STRENGTH) ? *a.selG[group].result : 1.0, S = LRR_SPATIAL(a.flags) ? *a.selS[group].result : 1.0;
    for (int d = 0; d < a.ndir; ++d) {
      if (merged || d == dir) { a.d[d].sc->G = G; a.d[d].sc->S = S; }
      if (!hit) atomicOr(&a.d[d].sc->status, LRR_ST_SELECT_MISS);   // every direction: k_accumulate must skip as a whole
    }
  }
  lrr_tl(4, 3);
}
void lrr_k_weight_ex(const EstArgs& a, const int* scale_mode, const double* scale_in, cudaStream_t st, long long* launches, int spec) {
  WeightExtra ex;
  for (int d = 0; d < 2; ++d) { ex.scale_mode[d] = scale_mode[d]; ex.scale_in[d] = scale_in[d]; }
  dim3 grid(resident_width(k_weight, LRR_THREADS, a.ndir, maxM(a)), a.ndir);
  lrr_launch_pdl(k_weight, grid, dim3(LRR_THREADS), 0, st, a, ex, spec); ++*launches;
}
void lrr_k_weight(const EstArgs& a, cudaStream_t st, long long* launches) {
  const int m[2] = {0, 0}; const double s[2] = {0, 0};
  lrr_k_weight_ex(a, m, s, st, launches, 0);
}

__global__ void __launch_bounds__(LRR_THREADS) k_collect(const __grid_constant__ EstArgs a, int spec) {
  lrr_pdl_enter();
  if (spec && lrr_spec_bad(a)) return;
  __shared__ double sh_cand[LRR_SEL_CAP];
  const int dir = blockIdx.y;
  const bool merged = (a.flags & LRR_F_MERGED) != 0;
  const int group = merged ? 0 : dir;
  const GridScope scope = lrr_scope(a, dir);
  // the histograms are not read after the weight phase located the bins: re-zero them for the next use
  {
    const int nthreads = gridDim.x * gridDim.y * blockDim.x, t0 = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    for (int g = 0; g < 2; ++g) {
      for (int i = t0; i < LRR_SEL_FINE; i += nthreads) { a.selG[g].fine[i] = 0; a.selS[g].fine[i] = 0; }
      for (int i = t0; i < LRR_SEL_COARSE; i += nthreads) { a.selG[g].coarse[i] = 0; a.selS[g].coarse[i] = 0; }
    }
  }
  lrr_phase_collect(scope, a.d[dir], a.flags, a.selG[group], a.selS[group]);
  if (!scope.last_block()) return;
  if (merged && a.ndir == 2 && !lrr_pair_done(a, 2)) return;
  bool ok = true;
  if (a.flags & LRR_F_STRENGTH) ok = lrr_sel_finish(a.selG[group], sh_cand) && ok;
  if (LRR_SPATIAL(a.flags)) ok = lrr_sel_finish(a.selS[group], sh_cand) && ok;
  if (threadIdx.x == 0) {
    const double G = (a.flags & LRR_F_STRENGTH) ? *a.selG[group].result : 1.0, S = LRR_SPATIAL(a.flags) ? *a.selS[group].result : 1.0;
    for (int d = 0; d < a.ndir; ++d) if (merged || d == dir) {
      a.d[d].sc->G = G; a.d[d].sc->S = S;
      if (!ok) a.d[d].sc->status |= LRR_ST_SELECT_OVERFLOW;
    }
  }
}
void lrr_k_collect(const EstArgs& a, cudaStream_t st, long long* launches, int spec) {
  dim3 grid(nblk_for(maxM(a)), a.ndir);
  lrr_launch_pdl(k_collect, grid, dim3(LRR_THREADS), 0, st, a, spec); ++*launches;
}

// 512 threads, one block per SM (128 registers): half as many partial rows of 108 moments for the last block to add as with two
// blocks of 256 per SM — its row sum is one batch of loads in flight instead of three
#ifndef LRR_ACC_THREADS
#define LRR_ACC_THREADS 512
#endif
__global__ void __launch_bounds__(LRR_ACC_THREADS) k_accumulate(const __grid_constant__ EstArgs a, int spec) {
  lrr_pdl_enter();
  if (spec && lrr_spec_bad(a)) return;
  __shared__ PhaseSharedT<LRR_ACC_THREADS / 32> sh;
  const int dir = blockIdx.y;
  const bool merged = (a.flags & LRR_F_MERGED) != 0, sym = (a.flags & LRR_F_SYM) != 0;
  const int reversed = sym && dir != a.src;
  const int keep_w = merged && dir != a.src;
  const Scal* scd = a.d[dir].sc;
  const double G = (a.flags & LRR_F_STRENGTH) ? scd->G : 1.0, S = LRR_SPATIAL(a.flags) ? scd->S : 1.0;
  double jc[3] = {0, 0, 0};
  if (!(a.flags & LRR_F_NORMALIZE) && !(a.flags & LRR_F_RIGID)) { const Scal* f = a.d[a.src].sc; jc[0] = f->c[0]; jc[1] = f->c[1]; jc[2] = f->c[2]; }
  if (scd->status & LRR_ST_SELECT_MISS) return;   // window miss (set on every direction by k_weight): nothing is committed, the host reruns
  lrr_tl(2, 0);
  const bool fin = lrr_phase_accumulate(lrr_scope(a, dir), a.d[dir], a.flags, reversed, keep_w, G, S, jc, sh);
  if (!fin) return;
  lrr_tl(9, 0);
  if (!lrr_pair_done(a, a.ndir)) return;
  lrr_tl(2, 2);
  lrr_solve(a.d[a.src].sc, a.ndir == 2 ? a.d[1 - a.src].sc : nullptr, a.flags, a.dof, a.estimate_parameters);
  lrr_tl(2, 3);
}
void lrr_k_accumulate(const EstArgs& a, cudaStream_t st, long long* launches, int spec) {
  // two blocks per SM at most: the last block sums one row of 108 moments per block before it can solve
  dim3 grid(resident_width(k_accumulate, LRR_ACC_THREADS, a.ndir, maxM(a)), a.ndir);   // one match per lane and trip
  lrr_launch_pdl(k_accumulate, grid, dim3(LRR_ACC_THREADS), 0, st, a, spec); ++*launches;
}

// Result MAILBOX: one_iteration returns `converged`, so the host reads ~1.6 kB of scalars per direction after every iteration.
// Instead of two device-to-host copies behind the kernel and a stream synchronisation, the block that finishes the residuals
// writes the head of each Scal into mapped pinned host memory, fences at system scope and publishes a sequence number; the host
// polls that number (pull_scal_mailbox, lrr_capi.cu).  Slot d = mbox + 512 d doubles, sequence number at [500].
#define LRR_MBOX_STRIDE 512
#define LRR_MBOX_FLAG 500
__device__ __forceinline__ void lrr_mailbox_post(const EstArgs& a, int dir_mask, double* mbox, unsigned long long seq, int nhead) {
  __syncthreads();   // the finalizer's scalar writes are visible to the block
  for (int d = 0; d < a.ndir; ++d) if (dir_mask & (1 << d)) {
    const double* src = reinterpret_cast<const double*>(a.d[d].sc);
    for (int i = threadIdx.x; i < nhead; i += blockDim.x) mbox[d * LRR_MBOX_STRIDE + i] = __ldcg(src + i);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int d = 0; d < a.ndir; ++d) if (dir_mask & (1 << d)) *reinterpret_cast<volatile unsigned long long*>(mbox + d * LRR_MBOX_STRIDE + LRR_MBOX_FLAG) = seq;
    __threadfence_system();
  }
}
__global__ void __launch_bounds__(LRR_THREADS) k_resid(const __grid_constant__ EstArgs a, int dir0, double* mbox, unsigned long long seq, int pull_mask, int nhead, int spec) {
  lrr_pdl_enter();
  __shared__ PhaseShared sh;
  const int dir = dir0 + blockIdx.y;
  // one row of blocks: its finalizer posts every direction the host wants; two rows: each finalizer posts its own direction
  const int post = gridDim.y == 2 ? (pull_mask & (1 << dir)) : pull_mask;
  if ((a.d[dir].sc->status & LRR_ST_SELECT_MISS) || (spec && lrr_spec_bad(a))) {   // the estimate was not committed: keep the residuals of the match for the rerun
    if (mbox && blockIdx.x == 0) lrr_mailbox_post(a, post, mbox, seq, nhead);
    return;
  }
  lrr_tl(8, 0);
  const bool fin = lrr_phase_resid(lrr_scope(a, dir), a.d[dir], sh);
  if (fin && mbox) lrr_mailbox_post(a, post, mbox, seq, nhead);
  lrr_tl(8, 3);
}
void lrr_k_resid(const EstArgs& a, int dir_mask, cudaStream_t st, long long* launches, double* mbox, unsigned long long seq, int pull_mask, int nhead, int spec) {
  if (dir_mask == 3) {
    dim3 grid(resident_width(k_resid, LRR_THREADS, 2, maxM(a)), 2);
    lrr_launch_pdl(k_resid, grid, dim3(LRR_THREADS), 0, st, a, 0, mbox, seq, pull_mask, nhead, spec); ++*launches;
  } else {
    const int d = (dir_mask == 2) ? 1 : 0;
    lrr_launch_pdl(k_resid, dim3(nblk_for(a.d[d].M)), dim3(LRR_THREADS), 0, st, a, d, mbox, seq, pull_mask, nhead, spec); ++*launches;
  }
}

// development aid: arm / read the per-block timeline marks (lrr_dev.cuh).  out: 12 kernels x 2048 blocks x 4 marks.
extern "C" int lrr_debug_timeline(int arm, unsigned long long* out) {
  static unsigned long long* buf = nullptr;
  const size_t n = 12 * 2048 * 4;
  if (arm) {
    if (!buf) { if (cudaMalloc(&buf, n * sizeof(unsigned long long)) != cudaSuccess) return 1; }
    cudaMemset(buf, 0, n * sizeof(unsigned long long));
    cudaMemcpyToSymbol(g_lrr_tl, &buf, sizeof(buf));
    lrr_tile_set_timeline(buf);
  } else {
    unsigned long long* z = nullptr;
    cudaDeviceSynchronize();
    cudaMemcpyToSymbol(g_lrr_tl, &z, sizeof(z));
    lrr_tile_set_timeline(nullptr);
    if (buf && out) cudaMemcpy(out, buf, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  }
  return 0;
}
