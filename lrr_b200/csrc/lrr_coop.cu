// One ICP iteration as ONE persistent cooperative kernel (sm_100a): the phases of lrr_icp.cu separated by grid-wide
// barriers instead of kernel boundaries.  At 100k features the per-iteration working set (~30 MB) is L2-resident and
// every phase is a single wave of dependent loads followed by a global reduction, so eight launches per iteration are
// mostly launch / drain / finalizer latency.  Here every block keeps running: a phase ends with one row of partials
// per block, a grid barrier, and then EVERY block sums the rows itself (fixed order: deterministic, and identical in
// every block), so no block has to wait for a designated finalizer.  Scalars written by the redundant finalizers are
// bit-identical across blocks.  The joint solve is done by block 0 between two barriers.
// Used for the steady state (scales estimated); the first iteration after set_transform (MUSE sort) takes the
// multi-kernel path.
#include <cooperative_groups.h>
#include <cstdlib>
#include "lrr_phases.cuh"
namespace cg = cooperative_groups;

struct CoopScope {
  double* partials;          // two row buffers (parity), LRR_MAX_BLOCKS * LRR_NRED each
  int* phase;                // per-thread-block phase counter in shared memory (buffer parity)
  __device__ __forceinline__ int first() const { return blockIdx.x * blockDim.x + threadIdx.x; }
  __device__ __forceinline__ int stride() const { return gridDim.x * blockDim.x; }
  __device__ __forceinline__ int block_first() const { return blockIdx.x * blockDim.x; }
  template <int K, int G = 1>
  __device__ __forceinline__ bool reduce(double (&v)[K], double* sh_out, double* sh_tmp) const {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int KG = K * G;
    __syncthreads();
    double* rows = partials + (size_t)((*phase) & 1) * LRR_MAX_BLOCKS * LRR_NRED;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = v[k];
#pragma unroll
      for (int o = 16; o >= G; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane < G) sh_tmp[wid * KG + lane * K + k] = x;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < KG; k += blockDim.x) {
      double s = 0;
      for (int w = 0; w < nw; ++w) s += sh_tmp[w * KG + k];
      rows[(size_t)blockIdx.x * LRR_NRED + k] = s;
    }
    cg::this_grid().sync();
    if (KG <= 32) {
      for (int k = wid; k < KG; k += nw) {
        double s = 0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(&rows[(size_t)b * LRR_NRED + k]);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh_out[k] = s;
      }
    } else {
      constexpr int KL = (KG + 31) / 32;
      double s[KL];
#pragma unroll
      for (int j = 0; j < KL; ++j) s[j] = 0;
#pragma unroll 4
      for (unsigned int b = wid; b < gridDim.x; b += nw) {
#pragma unroll
        for (int j = 0; j < KL; ++j) { const int k = lane + 32 * j; if (k < KG) s[j] += __ldcg(&rows[(size_t)b * LRR_NRED + k]); }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < KL; ++j) { const int k = lane + 32 * j; if (k < KG) sh_tmp[wid * KG + k] = s[j]; }
      __syncthreads();
      for (int k = threadIdx.x; k < KG; k += blockDim.x) {
        double t = 0;
        for (int w = 0; w < nw; ++w) t += sh_tmp[w * KG + k];
        sh_out[k] = t;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) *phase += 1;
    __syncthreads();
    return true;
  }
  __device__ __forceinline__ bool last_block() const { cg::this_grid().sync(); return true; }
  __device__ __forceinline__ void phase_end() const { __syncthreads(); }   // the block's own Scal writes become visible to its threads
};

struct CoopExtra { int scale_mode[2]; double scale_in[2]; long long* dbg; };
#define CPH(i) do { if (ex.dbg && blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); ex.dbg[i] += now_ - tmark; tmark = now_; } } while (0)

__global__ void __launch_bounds__(LRR_THREADS, 2) k_iteration(const __grid_constant__ EstArgs a, const CoopExtra ex) {
  __shared__ PhaseShared sh;
  __shared__ double s_T[2][15];
  __shared__ double sh_cand[LRR_SEL_CAP];
  __shared__ int s_phase;
  cg::grid_group grid = cg::this_grid();
  if (threadIdx.x == 0) s_phase = 0;
  __syncthreads();
  CoopScope scope; scope.partials = a.partials; scope.phase = &s_phase;
  const int nd = a.ndir, flags = a.flags;
  const bool merged = (flags & LRR_F_MERGED) != 0, sym = LRR_SPATIAL(flags), strength = (flags & LRR_F_STRENGTH) != 0;   // sym: spatial weighting on
  const bool use_sel = strength || sym;
  long long tmark = clock64();

  // ---- housekeeping: candidate lists empty (invariant shared with the multi-kernel path: the histograms are zero on entry) ----
  if (use_sel) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 < 4) {
      const SelDev& s = (t0 < 2) ? a.selG[t0] : a.selS[t0 - 2];
      *s.cand_count = 0; s.cand_minmax[0] = ~0ull; s.cand_minmax[1] = 0ull;
    }
  }
  // ---- MATCH: pass A, barrier (queues complete), pass B, one reduction for both directions ----
  for (int d = 0; d < nd; ++d) lrr_stage_transform(a.d[d].sc, s_T[d]);
  double m0[4] = {0, 0, 0, 0}, m1[4] = {0, 0, 0, 0};
  lrr_match_a_body_dynamic(a.d[0], flags, s_T[0], a.ticket + 4, m0);
  if (nd == 2) lrr_match_a_body_dynamic(a.d[1], flags, s_T[1], a.ticket + 5, m1);
  CPH(0);
  grid.sync();
  CPH(1);
  lrr_match_b_body_dynamic(a.d[0], flags, s_T[0], a.ticket + 6, m0);
  if (nd == 2) lrr_match_b_body_dynamic(a.d[1], flags, s_T[1], a.ticket + 7, m1);
  {
    CPH(2);
    double acc[8] = {m0[0], m0[1], m0[2], m0[3], m1[0], m1[1], m1[2], m1[3]};
    scope.reduce<8>(acc, sh.out, sh.tmp);
    CPH(3);
    if (threadIdx.x < nd) lrr_match_finalize(a.d[threadIdx.x], flags, sh.out + 4 * threadIdx.x);
    scope.phase_end();
  }
  // ---- NORM ----
  if (flags & LRR_F_NORMALIZE) {
    lrr_phase_norm1(scope, a.d[a.src], sh);
    lrr_phase_norm2(scope, a.d[a.src], nd == 2 ? a.d[1 - a.src].sc : nullptr, sh);
  }
  CPH(4);
  // ---- WEIGHT ----
  for (int d = 0; d < nd; ++d) {
    const int group = merged ? 0 : d;
    lrr_phase_weight(scope, a.d[d], flags, merged && d != a.src, a.selG[group], a.selS[group], ex.scale_mode[d], ex.scale_in[d], sh);
    scope.phase_end();
  }
  CPH(5);
  if (use_sel) {
    // every block locates the bins of the order statistics itself (identical results)
    if (merged && nd == 2) {
      if (strength) lrr_sel_locate(a.selG[0], a.d[0].sc->n_nonzero + a.d[1].sc->n_nonzero, sh.tmp, false);
      if (sym) lrr_sel_locate(a.selS[0], (double)(a.d[0].M + a.d[1].M), sh.tmp, false);
    } else {
      for (int d = 0; d < nd; ++d) {
        if (strength) lrr_sel_locate(a.selG[d], a.d[d].sc->n_nonzero, sh.tmp, false);
        if (sym) lrr_sel_locate(a.selS[d], (double)a.d[d].M, sh.tmp, false);
      }
    }
    __syncthreads();
    CPH(6);
    // ---- COLLECT ----
    for (int d = 0; d < nd; ++d) { const int group = merged ? 0 : d; lrr_phase_collect(scope, a.d[d], flags, a.selG[group], a.selS[group]); }
    grid.sync();
    const int ngroups = (merged || nd == 1) ? 1 : 2;
    for (int g = 0; g < ngroups; ++g) {
      bool ok = true;
      if (strength) ok = lrr_sel_finish(a.selG[g], sh_cand, false) && ok;   // every block reads the list: no reset here
      if (sym) ok = lrr_sel_finish(a.selS[g], sh_cand, false) && ok;
      if (threadIdx.x == 0) {
        const double G = strength ? *a.selG[g].result : 1.0, S = sym ? *a.selS[g].result : 1.0;
        for (int d = 0; d < nd; ++d) if (merged || d == g) { a.d[d].sc->G = G; a.d[d].sc->S = S; if (!ok) a.d[d].sc->status |= LRR_ST_SELECT_OVERFLOW; }
      }
    }
    __syncthreads();
  }
  CPH(7);
  // ---- ACCUMULATE + SOLVE ----
  {
    double jc[3] = {0, 0, 0};
    if (!(flags & LRR_F_NORMALIZE) && !(flags & LRR_F_RIGID)) { const Scal* f = a.d[a.src].sc; jc[0] = f->c[0]; jc[1] = f->c[1]; jc[2] = f->c[2]; }
    for (int d = 0; d < nd; ++d) {
      const Scal* scd = a.d[d].sc;
      lrr_phase_accumulate(scope, a.d[d], flags, (flags & LRR_F_SYM) && d != a.src, merged && d != a.src, strength ? scd->G : 1.0, sym ? scd->S : 1.0, jc, sh);
    }
    __syncthreads();
    CPH(8);
    if (blockIdx.x == 0) lrr_solve(a.d[a.src].sc, nd == 2 ? a.d[1 - a.src].sc : nullptr, flags, a.dof, a.estimate_parameters);
    __threadfence();
    CPH(9);
    grid.sync();
    CPH(10);
  }
  // the histograms are not read after the bins were located: zero them again for the next weight phase
  if (use_sel) {
    const int nthreads = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int g = 0; g < 2; ++g) {
      for (int i = t0; i < LRR_SEL_FINE; i += nthreads) { a.selG[g].fine[i] = 0; a.selS[g].fine[i] = 0; }
      for (int i = t0; i < LRR_SEL_COARSE; i += nthreads) { a.selG[g].coarse[i] = 0; a.selS[g].coarse[i] = 0; }
    }
  }
  // ---- RESID under the new transform ----
  lrr_phase_resid(scope, a.d[0], sh);
  // work counters of the two search passes: zero again for the next launch (every block is past the last barrier)
  if (blockIdx.x == 0 && threadIdx.x < 4) a.ticket[4 + threadIdx.x] = 0;
  CPH(11);
}

int lrr_coop_blocks() {
  static int nb = 0;
  if (!nb) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_iteration, LRR_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    nb = per_sm * lrr_num_sms();
    if (nb > LRR_MAX_BLOCKS) nb = LRR_MAX_BLOCKS;
  }
  return nb;
}

cudaError_t lrr_k_iteration(const EstArgs& a, const int* scale_mode, const double* scale_in, long long* dbg, cudaStream_t st, long long* launches) {
  CoopExtra ex; ex.dbg = dbg;
  for (int d = 0; d < 2; ++d) { ex.scale_mode[d] = scale_mode[d]; ex.scale_in[d] = scale_in[d]; }
  void* args[2] = {(void*)&a, (void*)&ex};
  // grid sized to the problem: with few points most of a full grid only adds barrier participants and partial rows
  int maxM = a.d[0].M; if (a.ndir == 2 && a.d[1].M > maxM) maxM = a.d[1].M;
  int nb = (maxM + LRR_THREADS - 1) / LRR_THREADS;
  if (const char* env = getenv("LRR_FUSED_BLOCKS")) nb = atoi(env);
  nb = nb < 8 ? 8 : (nb > lrr_coop_blocks() ? lrr_coop_blocks() : nb);
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_iteration, dim3(nb), dim3(LRR_THREADS), args, 0, st);
  ++*launches;
  return e;
}

// ---- micro-benchmark of the barrier / reduction primitives (diagnostics, not part of the product path) ----
__global__ void __launch_bounds__(LRR_THREADS, 2) k_sync_probe(double* partials, int reps, int mode, double* out) {
  __shared__ PhaseShared sh;
  __shared__ int s_phase;
  if (threadIdx.x == 0) s_phase = 0;
  __syncthreads();
  CoopScope scope; scope.partials = partials; scope.phase = &s_phase;
  double acc[5] = {1, 2, 3, 4, 5};
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) cg::this_grid().sync();
    else { scope.reduce<5>(acc, sh.out, sh.tmp); acc[0] = sh.out[0] * 1e-9; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = acc[0];
}
extern "C" int lrr_debug_sync_probe(double* partials, double* out, int reps, int mode, float* ms) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  void* args[4] = {(void*)&partials, (void*)&reps, (void*)&mode, (void*)&out};
  const int nb = lrr_coop_blocks();
  cudaLaunchCooperativeKernel((const void*)k_sync_probe, dim3(nb), dim3(LRR_THREADS), args, 0, 0);
  cudaDeviceSynchronize();
  cudaEventRecord(a, 0);
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_sync_probe, dim3(nb), dim3(LRR_THREADS), args, 0, 0);
  cudaEventRecord(b, 0); cudaEventSynchronize(b);
  cudaEventElapsedTime(ms, a, b);
  return (int)e;
}
