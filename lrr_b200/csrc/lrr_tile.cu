// K2, finest level: exact 1-NN with SHARED-MEMORY-STAGED CELLS (sm_100a).
//
// Why: the per-thread ball scan of lrr_dev.cuh walks a chain of dependent L2 loads (cell_start -> points -> next row)
// under heavy divergence; a 32-query chunk occupies its warp for 10-19 us and every resident warp has to work through
// several chunks one after the other (profiles/r1_summary.md, per-block timeline).  Here the from-features are put into
// spatial TILES once per set_features (cubes of LRR_TILE_T cells of the to-grid); a warp takes the queries of one tile
// 32 at a time, maps them, takes the bounding box of their cells plus a halo of LRR_TILE_H cells, copies the x-runs of
// that box (two coalesced rounds of independent loads: the run bounds, then the points) into its slice of shared memory
// and every lane scans from there: the dependent chain runs at shared-memory latency.
//
// Exactness is the ball-scan argument of lrr_dev.cuh unchanged: a query is finished here only when every cell its
// candidate ball touches lies inside the staged box (checked per query with the same inflated radius / slack), rows are
// pruned only when strictly farther than the current best, distances use lrr_dist2_exact, ties go to the lowest
// original index.  Everything else (no candidate and an empty neighbourhood, ball leaving the box, box too big for
// shared memory) goes to the same queue as before and is finished by k_match_hard.
// Replaces rsdl_kd_tree::n_nearest (cdcl/cdcl_estimation_ICP_matching_all.txx:572) / the Voronoi lookup
// (rrl/rrl_estimation_ICP_matching_all.txx:530-543) for the iterations where all points are in use.
#include "lrr_phases.cuh"
#include <algorithm>
#include <atomic>

#define LRR_TILE_WARPS 4
#define LRR_TILE_THREADS (32 * LRR_TILE_WARPS)
#define LRR_TW_CAP 512      // staged points per warp
#define LRR_TW_ROWS 128     // staged (y,z) rows per warp

struct TileWarp {
  double px[LRR_TW_CAP], py[LRR_TW_CAP], pz[LRR_TW_CAP];
  int rowlo[LRR_TW_ROWS];        // sorted slot of the first staged point of the row
  int rowbase[LRR_TW_ROWS + 1];  // offset of the row inside px/py/pz
};
struct TileBox { int rx0, rx1, ry0, ry1, rz0, rz1, RY; };

__device__ __forceinline__ GridScope lrr_scope_of(const EstArgs& a, int dir) {
  GridScope s; s.partials = a.partials + (size_t)dir * LRR_MAX_BLOCKS * LRR_NRED; s.ticket = a.ticket + dir; return s;
}
// true when this direction's matches go through the certificate pass first (finest level, tile order, certificates allocated)
__device__ __forceinline__ bool lrr_dir_certified(const DirArgs& d) { return d.proof && d.todo && d.chunk_tab && d.proc && d.M == d.n_in; }
void lrr_tile_set_timeline(unsigned long long* p) { cudaMemcpyToSymbol(g_lrr_tl, &p, sizeof(p)); }

// ---- tile construction (once per set_features) --------------------------------------------------------------
// tile_of[i] = tile * 512 + fine, fine = Morton code of the sub-cell (edge / 8) of feature i inside its tile: queries of one tile
// are processed in (fine, index) order, so the 32 queries of a chunk of a DENSE tile cover only part of the tile and
// their staged box stays small.
__global__ void k_tile_count(const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz, int n,
                             TileGeom g, int* __restrict__ cnt, int* __restrict__ tile_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ux = (fx[i] - g.ox) * g.inv_edge, uy = (fy[i] - g.oy) * g.inv_edge, uz = (fz[i] - g.oz) * g.inv_edge;
  int tx = (int)floor(ux), ty = (int)floor(uy), tz = (int)floor(uz);
  tx = tx < 0 ? 0 : (tx >= g.ntx ? g.ntx - 1 : tx); ty = ty < 0 ? 0 : (ty >= g.nty ? g.nty - 1 : ty); tz = tz < 0 ? 0 : (tz >= g.ntz ? g.ntz - 1 : tz);
  int sx = (int)((ux - tx) * 8.0), sy = (int)((uy - ty) * 8.0), sz = (int)((uz - tz) * 8.0);
  sx = sx < 0 ? 0 : (sx > 7 ? 7 : sx); sy = sy < 0 ? 0 : (sy > 7 ? 7 : sy); sz = sz < 0 ? 0 : (sz > 7 ? 7 : sz);
  int fine = 0;   // Morton code of the 8 x 8 x 8 sub-cells: 32 consecutive queries of a tile form a compact block
#pragma unroll
  for (int b = 0; b < 3; ++b) fine |= (((sx >> b) & 1) << (3 * b)) | (((sy >> b) & 1) << (3 * b + 1)) | (((sz >> b) & 1) << (3 * b + 2));
  const int t = (tz * g.nty + ty) * g.ntx + tx;
  tile_of[i] = t * 512 + fine;
  atomicAdd(cnt + t, 1);
}
// exclusive scan of the tile counts (single block); the counts are zeroed for their second life as scatter cursors
__global__ void __launch_bounds__(1024) k_tile_scan(int* __restrict__ cnt, int ntiles, int* __restrict__ tile_start) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (ntiles + T - 1) / T, lo = min(t * per, ntiles), hi = min(lo + per, ntiles);
  int sum = 0;
  for (int j = lo; j < hi; ++j) sum += cnt[j];
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < T; o <<= 1) { const int v = (t >= o) ? s_part[t - o] : 0; __syncthreads(); s_part[t] += v; __syncthreads(); }
  int run = s_part[t] - sum;
  for (int j = lo; j < hi; ++j) { const int c = cnt[j]; tile_start[j] = run; run += c; cnt[j] = 0; }
  if (t == T - 1) tile_start[ntiles] = s_part[T - 1];
}
// Chunks = runs of at most 32 consecutive entries of proc that stay inside one group of `ntx` consecutive tiles
// (ntx = 1: one tile; ntx = tiles per row: sparse tiles share chunks with their x-neighbours).
// Single block: scan of the rows' chunk counts, then the table entries (begin, end positions in proc).
__global__ void __launch_bounds__(1024) k_tile_chunks(const int* __restrict__ tile_start, int ntx, int nrows_t, int* __restrict__ chunk_start,
                                                     int2* __restrict__ chunk_tab) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (nrows_t + T - 1) / T, lo = min(t * per, nrows_t), hi = min(lo + per, nrows_t);
  int sum = 0;
  for (int r = lo; r < hi; ++r) sum += (tile_start[(r + 1) * ntx] - tile_start[r * ntx] + 31) >> 5;
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < T; o <<= 1) { const int v = (t >= o) ? s_part[t - o] : 0; __syncthreads(); s_part[t] += v; __syncthreads(); }
  int run = s_part[t] - sum;
  for (int r = lo; r < hi; ++r) {
    const int q0 = tile_start[r * ntx], q1 = tile_start[(r + 1) * ntx];
    for (int q = q0; q < q1; q += 32) chunk_tab[run++] = make_int2(q, min(q + 32, q1));
  }
  if (t == T - 1) chunk_start[0] = s_part[T - 1];   // total number of chunks
}
__global__ void k_tile_scatter(const int* __restrict__ tile_of, int n, const int* __restrict__ tile_start, int* __restrict__ cursor,
                               int* __restrict__ proc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tile_of[i] >> 9;
  proc[tile_start[t] + atomicAdd(cursor + t, 1)] = i;
}
// The scatter order inside a tile depends on the atomics: sort each tile's entries by (fine cell, index) so that the
// processing order (and with it the summation order of the match phase) is the same from run to run.  One block per
// tile, bitonic sort of 64-bit keys in shared memory (a tile with more than 2048 entries keeps its scatter order: still a
// valid processing order).
#define LRR_TILE_SORT_CAP 2048
__global__ void __launch_bounds__(256) k_tile_sort(int* __restrict__ proc, const int* __restrict__ tile_of, const int* __restrict__ tile_start) {
  __shared__ unsigned long long keys[LRR_TILE_SORT_CAP];
  const int t = blockIdx.x;
  const int q0 = tile_start[t], n = tile_start[t + 1] - q0;
  if (n < 2 || n > LRR_TILE_SORT_CAP) return;
  unsigned int np = 2; while ((int)np < n) np <<= 1;
  for (unsigned int i = threadIdx.x; i < np; i += blockDim.x) {
    unsigned long long kv = ~0ull;
    if ((int)i < n) { const int v = proc[q0 + i]; kv = ((unsigned long long)(tile_of[v] & 511) << 32) | (unsigned int)v; }
    keys[i] = kv;
  }
  __syncthreads();
  for (unsigned int k = 2; k <= np; k <<= 1)
    for (unsigned int j = k >> 1; j > 0; j >>= 1) {
      for (unsigned int i = threadIdx.x; i < np; i += blockDim.x) {
        const unsigned int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = keys[i], y = keys[ixj];
          if ((x > y) == ((i & k) == 0)) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) proc[q0 + i] = (int)(unsigned int)(keys[i] & 0xffffffffull);
}

void lrr_k_build_tiles(const double* fx, const double* fy, const double* fz, int n_in, const TileGeom& g, int* tile_cnt, int* tile_start,
                       int* chunk_start, int* tile_of, int* tmp, int2* chunk_tab, int* proc, cudaStream_t st, long long* launches) {
  if (n_in == 0) return;
  const int ntiles = g.ntx * g.nty * g.ntz;
  cudaMemsetAsync(tile_cnt, 0, sizeof(int) * (size_t)ntiles, st);
  k_tile_count<<<(n_in + 255) / 256, 256, 0, st>>>(fx, fy, fz, n_in, g, tile_cnt, tile_of); ++*launches;
  k_tile_scan<<<1, 1024, 0, st>>>(tile_cnt, ntiles, tile_start); ++*launches;
  k_tile_chunks<<<1, 1024, 0, st>>>(tile_start, 1, ntiles, chunk_start, chunk_tab); ++*launches;
  k_tile_scatter<<<(n_in + 255) / 256, 256, 0, st>>>(tile_of, n_in, tile_start, tile_cnt, proc); ++*launches;
  k_tile_sort<<<ntiles, 256, 0, st>>>(proc, tile_of, tile_start); ++*launches;
  (void)tmp;
}

__global__ void k_tile_mirror(const int* __restrict__ proc, const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz, int n,
                              int* __restrict__ pos_of, double* __restrict__ tf) {
  const int tq = blockIdx.x * blockDim.x + threadIdx.x;
  if (tq >= n) return;
  const int k = proc[tq];
  pos_of[k] = tq;
  tf[tq] = fx[k]; tf[(size_t)n + tq] = fy[k]; tf[2 * (size_t)n + tq] = fz[k];
}
void lrr_k_tile_mirror(const int* proc, const double* fx, const double* fy, const double* fz, int n_in, int* pos_of, double* tf, cudaStream_t st, long long* launches) {
  if (n_in <= 0) return;
  k_tile_mirror<<<(n_in + 255) / 256, 256, 0, st>>>(proc, fx, fy, fz, n_in, pos_of, tf); ++*launches;
}

// ---- search over the staged box ---------------------------------------------------------------------------------
// Every lane scans ALL staged points of the chunk's box (broadcast shared-memory reads, no per-lane control flow, four
// independent distance chains in flight).  The result is the exact neighbour iff the ball of the found distance lies
// inside the box — checked once at the end with the inflated radius / slack of lrr_dev.cuh; otherwise the query goes to
// pass B with what was found as its candidate.  Exact ties (equal fp64 distances of different points) are resolved
// towards the lowest original index in a second look at the tied points.
__device__ __forceinline__ void lrr_match_tile_body(const DirArgs& a, int flags, const double* s_T, TileWarp& W, unsigned int* counter, double (&acc)[4]) {
  const double* A = s_T; const double* tt = s_T + 9; const double* cc = s_T + 12;
  const bool rigid = (flags & LRR_F_RIGID) != 0;
  const GridDev& g = a.to.grid;
  const int lane = threadIdx.x & 31;
  const unsigned int full = 0xffffffffu;
  const int nchunks = __ldg(a.nchunks);
  const double slack = 1e-9 * g.h;
  unsigned long long* dbg = g_lrr_tl ? g_lrr_tl + 3 * 2048 * 4 : nullptr;   // development counters (scripts/timeline.py)
  // work list: every chunk of the direction, or — certificates in use — the (chunk, lane mask) entries k_match_cert left over
  const bool use_todo = a.todo != nullptr && a.proof != nullptr;
  const int nwork = use_todo ? *reinterpret_cast<volatile int*>(a.todo_count) : nchunks;
  // (the first item of a warp is assigned statically, the following ones are pulled from the counter: a work list shorter than the
  //  number of warps — the steady state with certificates — costs no same-address atomics)
  const int nstatic = gridDim.x * LRR_TILE_WARPS;
  unsigned int wi = blockIdx.x * LRR_TILE_WARPS + (threadIdx.x >> 5);
  bool pull = false;
  for (;;) {
    if (pull) {
      if (lane == 0) wi = atomicAdd(counter, 1u) + (unsigned int)nstatic;
      wi = __shfl_sync(full, wi, 0);
    }
    pull = true;
    if ((int)wi >= nwork) break;
    int ch = (int)wi; unsigned int lanes = full;
    if (use_todo) { const int2 e = __ldcg(a.todo + wi); ch = e.x; lanes = (unsigned int)e.y; }
    const int2 qq = __ldg(a.chunk_tab + ch);
    const int tq = qq.x + lane;
    const bool valid = tq < qq.y && ((lanes >> lane) & 1u);
    int k = 0; double x = 0, y = 0, z = 0, yx = 0, yy = 0, yz = 0;
    int cx = 0, cy = 0, cz = 0;
    if (valid) {
      k = __ldg(a.proc + tq);   // identity order: position k == compact index
      x = __ldg(a.fx + k); y = __ldg(a.fy + k); z = __ldg(a.fz + k);
      lrr_map_exact(A, tt, cc, x, y, z, yx, yy, yz);
      cx = lrr_cell_clamp(yx, g.ox, g.inv_h, g.nx); cy = lrr_cell_clamp(yy, g.oy, g.inv_h, g.ny); cz = lrr_cell_clamp(yz, g.oz, g.inv_h, g.nz);
    }
    // box of the chunk's cells plus the halo
    TileBox B;
    B.rx0 = max(__reduce_min_sync(full, valid ? cx : 0x7fffffff) - LRR_TILE_H, 0); B.rx1 = min(__reduce_max_sync(full, valid ? cx : -1) + LRR_TILE_H, g.nx - 1);
    B.ry0 = max(__reduce_min_sync(full, valid ? cy : 0x7fffffff) - LRR_TILE_H, 0); B.ry1 = min(__reduce_max_sync(full, valid ? cy : -1) + LRR_TILE_H, g.ny - 1);
    B.rz0 = max(__reduce_min_sync(full, valid ? cz : 0x7fffffff) - LRR_TILE_H, 0); B.rz1 = min(__reduce_max_sync(full, valid ? cz : -1) + LRR_TILE_H, g.nz - 1);
    B.RY = B.ry1 - B.ry0 + 1;
    const int nrows = B.RY * (B.rz1 - B.rz0 + 1);
    bool staged = nrows <= LRR_TW_ROWS;
    int total = 0;
    __syncwarp();   // the previous chunk's scans are done with W
    if (staged) {
      // bounds of the box's x-runs: all loads of a lane (up to LRR_TW_ROWS / 32 rows) in flight together
      int lo4[LRR_TW_ROWS / 32], sz4[LRR_TW_ROWS / 32];
#pragma unroll
      for (int u = 0; u < LRR_TW_ROWS / 32; ++u) {
        const int r = 32 * u + lane;
        lo4[u] = 0; sz4[u] = 0;
        if (r < nrows) {
          const int row = ((B.rz0 + r / B.RY) * g.ny + (B.ry0 + r % B.RY)) * g.nx;
          lo4[u] = __ldg(g.cell_start + row + B.rx0); sz4[u] = __ldg(g.cell_start + row + B.rx1 + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < LRR_TW_ROWS / 32; ++u) {
        if (32 * u < nrows) {   // warp-uniform
          const int r = 32 * u + lane, sz = (r < nrows) ? sz4[u] - lo4[u] : 0;
          int inc = sz;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(full, inc, o); if (lane >= o) inc += v; }
          if (r < nrows) { W.rowlo[r] = lo4[u]; W.rowbase[r] = total + inc - sz; }
          total += __shfl_sync(full, inc, 31);
        }
      }
      if (lane == 0) W.rowbase[nrows] = total;
      staged = total <= LRR_TW_CAP;
    }
    __syncwarp();
    if (dbg && lane == 0) { atomicAdd(dbg + 0, 1ull); if (!staged) atomicAdd(dbg + (nrows > LRR_TW_ROWS ? 1 : 2), 1ull); atomicAdd(dbg + 4, (unsigned long long)total); atomicAdd(dbg + 5, (unsigned long long)nrows); atomicAdd(dbg + 6, (unsigned long long)(qq.y - qq.x)); }
    if (!staged) {
      if (valid) lrr_match_a_one(a, flags, s_T, a.proc, tq, acc);   // box too large for shared memory: the global-memory scan
      continue;
    }
    // copy the points: four independent loads per lane in flight
    for (int p0 = 0; p0 < total; p0 += 128) {
      double4 P[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + 32 * u + lane;
        if (p < total) {
          int ra = 0, rb = nrows;   // last row with rowbase <= p
          while (rb - ra > 1) { const int mid = (ra + rb) >> 1; if (W.rowbase[mid] <= p) ra = mid; else rb = mid; }
          P[u] = lrr_ld_pos4(g.pos4 + W.rowlo[ra] + (p - W.rowbase[ra]));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + 32 * u + lane;
        if (p < total) { W.px[p] = P[u].x; W.py[p] = P[u].y; W.pz[p] = P[u].z; }
      }
    }
    __syncwarp();
    // uniform scan: every lane, every staged point
    // (second = smallest squared distance of any staged point other than the winner, for the certificate)
    double best = 1e300, second = 1e300; int bi = -1; bool tie = false;
    {
      int p = 0;
      for (; p + 4 <= total; p += 4) {
        const double d0 = lrr_dist2_exact(yx, yy, yz, W.px[p], W.py[p], W.pz[p]);
        const double d1 = lrr_dist2_exact(yx, yy, yz, W.px[p + 1], W.py[p + 1], W.pz[p + 1]);
        const double d2 = lrr_dist2_exact(yx, yy, yz, W.px[p + 2], W.py[p + 2], W.pz[p + 2]);
        const double d3 = lrr_dist2_exact(yx, yy, yz, W.px[p + 3], W.py[p + 3], W.pz[p + 3]);
        if (d0 <= best) { tie = (d0 == best); second = best; best = d0; bi = p; } else second = fmin(second, d0);
        if (d1 <= best) { tie = (d1 == best); second = best; best = d1; bi = p + 1; } else second = fmin(second, d1);
        if (d2 <= best) { tie = (d2 == best); second = best; best = d2; bi = p + 2; } else second = fmin(second, d2);
        if (d3 <= best) { tie = (d3 == best); second = best; best = d3; bi = p + 3; } else second = fmin(second, d3);
      }
      for (; p < total; ++p) {
        const double d0 = lrr_dist2_exact(yx, yy, yz, W.px[p], W.py[p], W.pz[p]);
        if (d0 <= best) { tie = (d0 == best); second = best; best = d0; bi = p; } else second = fmin(second, d0);
      }
    }
    // (tie is true iff the final minimum was reached by at least two points: a smaller distance found later clears it)
    int best_s = -1, best_o = 0x7fffffff;
    bool done = false;
    if (valid && bi >= 0) {
      int ra = 0, rb = nrows;
      while (rb - ra > 1) { const int mid = (ra + rb) >> 1; if (W.rowbase[mid] <= bi) ra = mid; else rb = mid; }
      best_s = W.rowlo[ra] + (bi - W.rowbase[ra]);
      if (tie) {   // rare: several points at exactly the minimum distance -> lowest original index
        best_o = lrr_pos4_orig(lrr_ld_pos4(g.pos4 + best_s));
        for (int p = 0; p < total; ++p) {
          if (p != bi && lrr_dist2_exact(yx, yy, yz, W.px[p], W.py[p], W.pz[p]) == best) {
            int r2a = 0, r2b = nrows;
            while (r2b - r2a > 1) { const int mid = (r2a + r2b) >> 1; if (W.rowbase[mid] <= p) r2a = mid; else r2b = mid; }
            const int s2 = W.rowlo[r2a] + (p - W.rowbase[r2a]);
            const int o2 = lrr_pos4_orig(lrr_ld_pos4(g.pos4 + s2));
            if (o2 < best_o) { best_o = o2; best_s = s2; }
          }
        }
      }
      // proven iff every cell the ball touches is inside the staged box (grid borders clamp on both sides alike)
      const double R = sqrt(best) * (1.0 + 1e-9) + slack;
      const int z0 = lrr_cell_clamp(yz - R, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(yz + R, g.oz, g.inv_h, g.nz);
      const int y0 = lrr_cell_clamp(yy - R, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(yy + R, g.oy, g.inv_h, g.ny);
      const int xa = lrr_cell_clamp(yx - R, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(yx + R, g.ox, g.inv_h, g.nx);
      done = !(z0 < B.rz0 || z1 > B.rz1 || y0 < B.ry0 || y1 > B.ry1 || xa < B.rx0 || xb > B.rx1);
    }
    // unfinished queries -> pass B queue, one atomic per warp
    const unsigned int qm = __ballot_sync(full, valid && !done);
    if (qm) {
      int base = 0;
      if (lane == __ffs(qm) - 1) base = atomicAdd(a.hq_count, __popc(qm));
      base = __shfl_sync(full, base, __ffs(qm) - 1);
      if (valid && !done) {
        if (best_s >= 0 && !tie) best_o = lrr_pos4_orig(lrr_ld_pos4(g.pos4 + best_s));
        if (best_s < 0) best = 1e300;
        const int cand = a.prev_s[k];   // the neighbour of the previous iteration is usually a much tighter bound than the best point of the box
        if (cand >= 0) {
          const double4 pc = lrr_ld_pos4(g.pos4 + cand);
          const double dc = lrr_dist2_exact(yx, yy, yz, pc.x, pc.y, pc.z);
          const int oc = lrr_pos4_orig(pc);
          if (dc < best || (dc == best && oc < best_o)) { best = dc; best_s = cand; best_o = oc; }
        }
        const int slot = base + __popc(qm & ((1u << lane) - 1u));
        a.hq_k[slot] = k; a.hq_best[slot] = best_s >= 0 ? best : 1e300; a.hq_s[slot] = best_s; a.hq_o[slot] = best_s >= 0 ? best_o : 0x7fffffff;
      }
      if (dbg && lane == 0) atomicAdd(dbg + 3, (unsigned long long)__popc(qm));
    }
    if (valid && done) {
      // certificate: every other staged point is at least sqrt(second) away, every point that was not staged lies outside the
      // box, i.e. at least as far as the nearest box face that has cells behind it (faces on the grid border have none: border
      // cells hold everything beyond them).  Rounded down.
      double bound = sqrt(second);
      if (B.rx0 > 0) bound = fmin(bound, yx - (g.ox + B.rx0 * g.h));
      if (B.rx1 < g.nx - 1) bound = fmin(bound, (g.ox + (B.rx1 + 1) * g.h) - yx);
      if (B.ry0 > 0) bound = fmin(bound, yy - (g.oy + B.ry0 * g.h));
      if (B.ry1 < g.ny - 1) bound = fmin(bound, (g.oy + (B.ry1 + 1) * g.h) - yy);
      if (B.rz0 > 0) bound = fmin(bound, yz - (g.oz + B.rz0 * g.h));
      if (B.rz1 < g.nz - 1) bound = fmin(bound, (g.oz + (B.rz1 + 1) * g.h) - yz);
      bound = bound * (1.0 - 1e-9) - slack;
      lrr_match_finish(a, rigid, k, k, x, y, z, yx, yy, yz, best_s, acc, (a.proof && !tie && bound > 0.0) ? 2 : 0, bound);
    }
  }
}


// ---- sparse list: the queries of chunks in which the certificate pass left only a few over ----------------------------------
// Near convergence a failing certificate is almost always a near tie (the second nearest feature a few micrometres farther than the
// nearest: measured, scripts/cert_stats2.py) of an ordinary query whose ball touches ~10 rows and ~20 points.  Staging a whole
// chunk's box for one or two such queries costs ~25 us of single-warp latency, and so does a whole warp per query one after the
// other.  Here a group of EIGHT lanes takes one query: the run bounds of up to 32 rows in one round trip (four per lane), a prefix
// sum over the group, the points in a second round trip, group minimum — the flattened ball scan of lrr_dev.cuh at a quarter of a
// warp, four queries per warp at a time, every query of the list in (almost) one round.  Balls of more than LRR_G8_ROWS rows or
// LRR_G8_PTS points (far outliers) are queued for the one-warp-per-query pass.  Exact, with a new certificate.
__device__ __forceinline__ void lrr_match_sparse_body(const DirArgs& a, int flags, const double* s_T, TileWarp& W, double (&acc)[4]) {
  const int nfq = *reinterpret_cast<volatile int*>(a.fq_count);
  if (nfq <= 0) return;
  const double* A = s_T; const double* tt = s_T + 9; const double* cc = s_T + 12;
  const bool rigid = (flags & LRR_F_RIGID) != 0;
  const GridDev& g = a.to.grid;
  const int lane = threadIdx.x & 31, gl = lane & 7, gi = lane >> 3;
  const unsigned int gmask = 0xffu << (8 * gi);
  int* rowlo = W.rowlo + LRR_G8_ROWS * gi; int* rowbase = W.rowbase + LRR_G8_ROWS * gi;
  const int ngroups = gridDim.x * LRR_TILE_WARPS * 4;
  __syncwarp();
  for (int qi = (blockIdx.x * LRR_TILE_WARPS + (threadIdx.x >> 5)) * 4 + gi; qi < nfq; qi += ngroups) {
    // the record the certificate pass left: match position, previous neighbour (slot, original index, squared distance under the
    // current transform) and the from-location, three 16-byte loads of one hop
    const int4* rp = reinterpret_cast<const int4*>(a.fq + qi);
    const int4 ri = __ldcg(rp);
    const double2 r1 = __ldcg(reinterpret_cast<const double2*>(rp + 1)), r2 = __ldcg(reinterpret_cast<const double2*>(rp + 2));
    const int k = ri.x, cand = ri.y;
    const double x = r1.x, y = r1.y, z = r2.x;
    double qx, qy, qz;
    lrr_map_exact(A, tt, cc, x, y, z, qx, qy, qz);
    double best = r2.y, bound = 0.0; int best_s = cand, best_o = ri.z;
    const bool ok = cand >= 0 && lrr_nn_ball_g8(g, qx, qy, qz, 0.125 * g.h, rowlo, rowbase, gmask, gl, best, best_s, best_o, bound);
    if (gl == 0) {
      if (!ok) {
        const int slot = atomicAdd(a.hq_count, 1);
        a.hq_k[slot] = k; a.hq_best[slot] = cand >= 0 ? r2.y : 1e300; a.hq_s[slot] = cand; a.hq_o[slot] = cand >= 0 ? ri.z : 0x7fffffff;
      } else lrr_match_finish(a, rigid, k, k, x, y, z, qx, qy, qz, best_s, acc, (a.proof && bound > 0.0) ? 2 : 0, bound);
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(LRR_TILE_THREADS, 4) k_match_tile(const __grid_constant__ EstArgs a) {
  lrr_pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_T[15];
  __shared__ double s_out[8], s_tmp[8 * LRR_TILE_WARPS];
  const int dir = blockIdx.y;
  const DirArgs& d = a.d[dir];
  unsigned int* counter = a.ticket + 4 + dir;
  lrr_tl(0, 0);
  // certificates in use and nothing left over by k_match_cert (the steady state): its sums stand, the work list is empty already
  if (lrr_dir_certified(d) && *reinterpret_cast<volatile int*>(d.todo_count) == 0 && *reinterpret_cast<volatile int*>(d.fq_count) == 0) return;
  lrr_stage_transform(d.sc, s_T);
  double acc[4] = {0, 0, 0, 0};
  if (d.chunk_tab && d.proc && d.M == d.n_in) {
    TileWarp& W = reinterpret_cast<TileWarp*>(smem_raw)[threadIdx.x >> 5];
    lrr_match_tile_body(d, a.flags, s_T, W, counter, acc);
    if (lrr_dir_certified(d)) lrr_match_sparse_body(d, a.flags, s_T, W, acc);
  } else {
    lrr_match_a_body_dynamic(d, a.flags, s_T, counter, acc);
  }
  lrr_tl(0, 1);
  const GridScope scope = lrr_scope_of(a, dir);
  if (scope.reduce<4>(acc, s_out, s_tmp)) {
    lrr_tl(0, 2);
    if (threadIdx.x == 0) {
      Scal* sc = d.sc;
      sc->norm_dirty = 1;
      if (!lrr_dir_certified(d)) sc->match_left = 1;
      if (lrr_dir_certified(d)) {   // the certificate pass left its share of the sums in pm_*
        sc->sum_loc[0] = sc->pm_sum_loc[0] + s_out[0]; sc->sum_loc[1] = sc->pm_sum_loc[1] + s_out[1]; sc->sum_loc[2] = sc->pm_sum_loc[2] + s_out[2];
        sc->sum_r2 = sc->pm_sum_r2 + s_out[3];
        *d.todo_count = 0; *d.fq_count = 0;
      } else { sc->sum_loc[0] = s_out[0]; sc->sum_loc[1] = s_out[1]; sc->sum_loc[2] = s_out[2]; sc->sum_r2 = s_out[3]; }
      *counter = 0;
    }
  }
  lrr_tl(0, 3);
}

// ---- certificate pass ------------------------------------------------------------------------------------------------
// NEIGHBOUR CERTIFICATE of a from-feature: (y0, B) = the mapped position at which its neighbour q_prev was proven the UNIQUE nearest
// to-feature, and a lower bound B on the distance from y0 to every OTHER to-feature (written by the staged search: second smallest
// staged distance / nearest box face; by the ball scans: second smallest distance inside the inflated ball / its radius).  For a
// new mapped position y and any other feature p:  |y - p| >= |y0 - p| - |y - y0| >= B - |y - y0|,  so
//        |y - q_prev| + |y - y0| < B     =>     q_prev is still the unique nearest neighbour (no search, no tie possible).
// ICP moves the mapped positions less and less, so near convergence almost every query is kept this way; the others go to the
// staged search through the work list.  The result is the exact neighbour either way (bit-exact index sets vs the oracle).
// One warp per chunk, no shared-memory staging: full occupancy, coalesced along the tile order.
//
// TWO ROLES IN ONE PASS.  At the finest level the match of iteration i+1 is the match of iteration i for every kept query, and the
// certificate test evaluates y = T(x), q, P and e^T P e of exactly that match under exactly the transform the residual pass of
// iteration i (RMS_error / weighted_error after estimate_LS: cdcl :499, rrl sym :366) evaluates them under.  So the pass that ENDS
// one_iteration(i) is this kernel with LRR_CERT_RESID: it produces the residual sums of iteration i over all matches (weights of
// iteration i) AND runs the certificate test of find_closest(i+1); one_iteration(i+1) then starts at the weights (pre-match,
// lrr_capi.cu).  A kept match is not touched at all (index, certificate and — see DirArgs::stamp_in_match — stamp stay what they
// are); the residual r2 is rewritten for every match, as the residual pass does.  Queries that are not kept stay what they were in
// iteration i until the staged search completes the match, and the sums of the kept part wait in Scal::pm_*.
#define LRR_CERT_THREADS 256
#ifndef LRR_CERT_MINBLK
#define LRR_CERT_MINBLK 3
#endif
enum { LRR_CERT_RESID = 1, LRR_CERT_RESCALE = 2 };
#define LRR_MBOX_STRIDE 512
#define LRR_MBOX_FLAG 500
// (same protocol as lrr_mailbox_post in lrr_icp.cu: slot d = mbox + 512 d doubles, sequence number at [500])
__device__ __forceinline__ void lrr_cert_mailbox_post(const DirArgs& d, int dir, double* mbox, unsigned long long seq, int nhead) {
  __syncthreads();
  const double* src = reinterpret_cast<const double*>(d.sc);
  for (int i = threadIdx.x; i < nhead; i += blockDim.x) mbox[dir * LRR_MBOX_STRIDE + i] = __ldcg(src + i);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) { *reinterpret_cast<volatile unsigned long long*>(mbox + dir * LRR_MBOX_STRIDE + LRR_MBOX_FLAG) = seq; __threadfence_system(); }
}
__global__ void __launch_bounds__(LRR_CERT_THREADS, LRR_CERT_MINBLK) k_match_cert(const __grid_constant__ EstArgs a, int mode, double* mbox, unsigned long long seq,
                                                                                  int pull_mask, int nhead) {
  lrr_pdl_enter();
  __shared__ double s_T[15];
  __shared__ double s_out[16], s_tmp[16 * (LRR_CERT_THREADS / 32)];
  const int dir = blockIdx.y;
  const DirArgs& d = a.d[dir];
  const bool resid = (mode & LRR_CERT_RESID) != 0;
  lrr_tl(5, 0);
  if (resid && (d.sc->status & (LRR_ST_SELECT_MISS | LRR_ST_NOT_STEADY))) {
    // the estimate was not committed (order-statistic window missed / speculative plan refused): nothing changes, the host reruns
    if (mbox && blockIdx.x == 0 && (pull_mask & (1 << dir))) lrr_cert_mailbox_post(d, dir, mbox, seq, nhead);
    return;
  }
  lrr_stage_transform(d.sc, s_T);
  double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [9]: queries not kept (instrumentation); kept part of the next match: sum of from-locations (3), sum r2; residual pass: r2, w r2, w, w sqrt(r2), count
  if (lrr_dir_certified(d)) {
    const double* A = s_T; const double* tt = s_T + 9; const double* cc = s_T + 12;
    const bool rigid = (a.flags & LRR_F_RIGID) != 0;
    const GridDev& g = d.to.grid;
    const double slack = 1e-9 * g.h;
    const int lane = threadIdx.x & 31;
    const int nchunks = __ldg(d.nchunks);
    const int wpb = LRR_CERT_THREADS / 32;
    const double wfac = (mode & LRR_CERT_RESCALE) ? d.sc->wfac : 1.0;
    // per query: one hop for everything at its tile-order position (neighbour, certificate, location, match position), one hop for
    // the neighbour's location and projector and the match's weight (all issued before the certificate is checked), then the stores
    const int cstep = gridDim.x * wpb;
    int ch = blockIdx.x * wpb + (threadIdx.x >> 5);
    int2 qq = ch < nchunks ? __ldg(d.chunk_tab + ch) : make_int2(0, 0);
    while (ch < nchunks) {
      const int chn = ch + cstep;
      const int2 qqn = chn < nchunks ? __ldg(d.chunk_tab + chn) : make_int2(0, 0);   // (the next chunk's bounds are fetched a chunk ahead)
      const int tq = qq.x + lane;
      const bool valid = tq < qq.y;
      bool kept = false;
      if (valid) {
        const int k = __ldg(d.proc + tq);
        const int sp = d.cprev[tq];
        const double2* pr = reinterpret_cast<const double2*>(d.proof + tq);
        const double2 p0 = __ldcg(pr), p1 = __ldcg(pr + 1);
        const double x = __ldg(d.tfx + tq), y = __ldg(d.tfy + tq), z = __ldg(d.tfz + tq);
        if (sp >= 0 && (resid || p1.y > 0.0)) {
          const double4 q = lrr_ld_pos4(g.pos4 + sp);
          double P[6]; lrr_ld_P6(d.to.rec, sp, P);
          double w = 0.0;
          if (resid) w = d.w[k];
          double yx, yy, yz;
          lrr_map_exact(A, tt, cc, x, y, z, yx, yy, yz);
          const double r2 = lrr_proj_r2(yx - q.x, yy - q.y, yz - q.z, P[0], P[1], P[2], P[3], P[4], P[5]);
          if (resid) {
            if (mode & LRR_CERT_RESCALE) { w *= wfac; d.w[k] = w; }
            d.r2[k] = r2;
            acc[4] += r2; acc[5] += w * r2; acc[6] += w; acc[7] += sqrt(r2) * w; acc[8] += 1.0;
          }
          if (p1.y > 0.0) {
            const double dq = sqrt(lrr_dist2_exact(yx, yy, yz, q.x, q.y, q.z));
            const double ex = yx - p0.x, ey = yy - p0.y, ez = yz - p1.x;
            const double dm = sqrt(ex * ex + ey * ey + ez * ez);
            // (1e-9 relative + slack covers the rounding of the square roots and sums; B itself was rounded down when written)
            if ((dq + dm) * (1.0 + 1e-9) + slack < p1.y) {
              kept = true;
              if (!resid) d.r2[k] = r2;
              if (rigid) { acc[0] += yx; acc[1] += yy; acc[2] += yz; } else { acc[0] += x; acc[1] += y; acc[2] += z; }
              acc[3] += r2;
            }
          }
          if (g_lrr_tl) {   // development counters (scripts/cert_stats.py)
            unsigned long long* c = g_lrr_tl + 3 * 2048 * 4;
            atomicAdd(c + (kept ? 7 : 18), 1ull);
            if (!kept && p1.y > 0.0) {
              const double dq0 = sqrt(lrr_dist2_exact(p0.x, p0.y, p1.x, q.x, q.y, q.z));
              const double ex = yx - p0.x, ey = yy - p0.y, ez = yz - p1.x;
              const double dm = sqrt(ex * ex + ey * ey + ez * ez);
              if (dq0 > 2.0 * g.h) atomicAdd(c + 20, 1ull);
              if (p1.y >= 0.999 * (dq0 + 0.125 * g.h)) atomicAdd(c + 21, 1ull); else atomicAdd(c + 22, 1ull);
              atomicAdd(c + 23, (unsigned long long)(dm * 1e6));
              atomicAdd(c + 24, (unsigned long long)(fmax(p1.y - dq0, 0.0) * 1e6));
            }
          }
        } else if (g_lrr_tl) atomicAdd(g_lrr_tl + 3 * 2048 * 4 + (sp < 0 ? 16 : 17), 1ull);
      }
      const unsigned int rest = __ballot_sync(0xffffffffu, valid && !kept);
      if (rest) {
        const int nrest = __popc(rest);
        if (nrest <= a.tile_sparse) {   // a few queries: each goes to the sparse list on its own (lrr_match_sparse_body)
          int base = 0;
          if (lane == 0) base = atomicAdd(d.fq_count, nrest);
          base = __shfl_sync(0xffffffffu, base, 0);
          // A query WITHOUT a previous neighbour (a from-feature the point-count ramp has just added) borrows the candidate of the nearest
          // lane that has one: the 32 queries of a chunk lie in one sub-cell of a tile, so a neighbouring query's neighbour is a few
          // millimetres away and bounds the ball scan of the eight-lane groups; without a candidate the query would go to the
          // one-warp-per-query pass and its shell search.  Any to-feature is a valid candidate: the search is exact whatever it starts from.
          const int sp_mine = valid ? d.cprev[tq] : -1;
          const unsigned int have = __ballot_sync(0xffffffffu, sp_mine >= 0);
          int donor = lane;
          if (sp_mine < 0 && have) {
            const unsigned int below = have & ((1u << lane) - 1u), above = lane < 31 ? (have & ~((2u << lane) - 1u)) : 0u;
            const int lb = below ? 31 - __clz(below) : -1, la = above ? __ffs(above) - 1 : -1;
            donor = (lb >= 0 && (la < 0 || lane - lb <= la - lane)) ? lb : la;
          }
          const int sp_cand = __shfl_sync(0xffffffffu, sp_mine, donor);
          if (valid && !kept) {
            // (the record is put together again from the caches here, on the rare path, instead of being carried in registers through the
            //  common one: that cost the kernel 76 bytes of spills and ~6 us per launch)
            FqRec r; r.k = __ldg(d.proc + tq); r.cand = sp_cand; r.cand_o = 0x7fffffff; r.tq = tq;
            r.x = __ldg(d.tfx + tq); r.y = __ldg(d.tfy + tq); r.z = __ldg(d.tfz + tq); r.best = 1e300;
            if (r.cand >= 0) {
              const double4 q = lrr_ld_pos4(g.pos4 + r.cand);
              double yx, yy, yz; lrr_map_exact(A, tt, cc, r.x, r.y, r.z, yx, yy, yz);
              r.cand_o = lrr_pos4_orig(q); r.best = lrr_dist2_exact(yx, yy, yz, q.x, q.y, q.z);
            }
            d.fq[base + __popc(rest & ((1u << lane) - 1u))] = r;
          }
        } else if (lane == 0) { const int slot = atomicAdd(d.todo_count, 1); d.todo[slot] = make_int2(ch, (int)rest); }
        if (lane == 0) acc[9] += (double)nrest;
      }
      ch = chn; qq = qqn;
    }
  }
  lrr_tl(5, 1);
  const GridScope scope = lrr_scope_of(a, dir);
  if (scope.reduce<10>(acc, s_out, s_tmp)) {
    if (threadIdx.x == 0) {
      Scal* sc = d.sc;
      sc->cert_failed = s_out[9];
      // every query kept <=> nothing on the work list (this block is the last one of the direction: all appends are visible)
      const int left = lrr_dir_certified(d) ? *reinterpret_cast<volatile int*>(d.todo_count) + *reinterpret_cast<volatile int*>(d.fq_count) + *reinterpret_cast<volatile int*>(d.hq_count) : 1;
      sc->match_left = left;
      if (left > 0) {
        sc->norm_dirty = 1;
        sc->pm_sum_loc[0] = s_out[0]; sc->pm_sum_loc[1] = s_out[1]; sc->pm_sum_loc[2] = s_out[2]; sc->pm_sum_r2 = s_out[3];
      } else {   // nothing follows for this direction: the bookkeeping of a finished find_closest (k_match_hard repeats it when it is launched)
        const double tot[4] = {s_out[0], s_out[1], s_out[2], s_out[3]};
        lrr_match_finalize(d, a.flags, tot);
      }
      if (resid) {
        sc->res_sum_r2 = s_out[4]; sc->res_sum_wr2 = s_out[5]; sc->res_sum_w = s_out[6]; sc->res_sum_wr = s_out[7]; sc->res_count = s_out[8];
      }
    }
    if (mbox && (pull_mask & (1 << dir))) lrr_cert_mailbox_post(d, dir, mbox, seq, nhead);
  }
  lrr_tl(5, 3);
}
bool lrr_match_cert_usable(const EstArgs& a) {
  for (int d = 0; d < a.ndir; ++d) if (a.d[d].proof && a.d[d].todo && a.d[d].chunk_tab && a.d[d].proc && a.d[d].M == a.d[d].n_in) return true;
  return false;
}
void lrr_k_match_cert(const EstArgs& a, cudaStream_t st, long long* launches, int mode, double* mbox, unsigned long long seq, int pull_mask, int nhead) {
  int maxM = a.d[0].M; if (a.ndir == 2 && a.d[1].M > maxM) maxM = a.d[1].M;
  int nb = (maxM / 32 + LRR_CERT_THREADS / 32 - 1) / (LRR_CERT_THREADS / 32) + 1;   // about one chunk per warp ...
  static std::atomic<int> per_sm_dev[LRR_MAX_DEVICES];   // per device
  int per_sm = per_sm_dev[lrr_current_device()].load(std::memory_order_relaxed);
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_match_cert, LRR_CERT_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    per_sm_dev[lrr_current_device()].store(per_sm, std::memory_order_relaxed);
  }
  nb = std::max(1, std::min(nb, per_sm * lrr_num_sms() / a.ndir));   // ... but every block resident at once (no second wave)
  dim3 grid(nb, a.ndir);
  lrr_launch_pdl(k_match_cert, grid, dim3(LRR_CERT_THREADS), 0, st, a, mode, mbox, seq, pull_mask, nhead); ++*launches;
}

bool lrr_match_tile_usable(const EstArgs& a) {
  for (int d = 0; d < a.ndir; ++d) if (a.d[d].chunk_tab && a.d[d].proc && a.d[d].M == a.d[d].n_in) return true;
  return false;
}
void lrr_k_match_tile(const EstArgs& a, cudaStream_t st, long long* launches) {
  // (cudaFuncSetAttribute is per DEVICE: the 53 KB of dynamic shared memory have to be allowed on every device the library runs on)
  static std::atomic<int> attr_set[LRR_MAX_DEVICES];
  const size_t smem = sizeof(TileWarp) * LRR_TILE_WARPS;
  if (!attr_set[lrr_current_device()].load(std::memory_order_relaxed)) {
    cudaFuncSetAttribute(k_match_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[lrr_current_device()].store(1, std::memory_order_relaxed);
  }
  // every block of both directions resident at once (4 blocks of 128 threads per SM): no second wave
  dim3 grid(std::max(1, (4 * lrr_num_sms()) / a.ndir), a.ndir);
  lrr_launch_pdl(k_match_tile, grid, dim3(LRR_TILE_THREADS), smem, st, a); ++*launches;
}
