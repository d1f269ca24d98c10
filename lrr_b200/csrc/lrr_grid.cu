// K2: uniform-grid build and exact Euclidean 1-NN (sm_100a).
// Replaces rsdl_kd_tree::n_nearest(n=1) (cdcl/cdcl_estimation_ICP_matching_all.txx:572, build :113-135) and the
// Voronoi-map lookup of rrl/rrl_estimation_ICP_matching_all.txx:530-543 (map built by
// rrl/examples/compute_distance_map.cxx:55-151).  The result is the EXACT nearest neighbour: squared distance
// dx*dx+dy*dy+dz*dz evaluated left to right in fp64 without FMA contraction, ties -> lowest original index —
// the same rule as the oracle, so index sets are bit-exact.
#include "lrr_internal.h"
#include "lrr_dev.cuh"
#include <math.h>
#include <stdlib.h>

// per-device facts are cached PER DEVICE (contexts on several devices may live in one process and be driven from several threads)
#include <atomic>
static std::atomic<int> g_num_sms[LRR_MAX_DEVICES];
int lrr_current_device() { int dev = 0; cudaGetDevice(&dev); return (dev < 0 || dev >= LRR_MAX_DEVICES) ? 0 : dev; }
int lrr_num_sms() {
  const int dev = lrr_current_device();
  int n = g_num_sms[dev].load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    g_num_sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---- bounding box -----------------------------------------------------------------------------
__global__ void k_bbox(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z, int n,
                       double* __restrict__ partial /* gridDim.x * 6 */) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v[3] = {x[i], y[i], z[i]};
#pragma unroll
    for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], v[d]); mx[d] = fmax(mx[d], v[d]); }
  }
  __shared__ double s[6][32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o; o >>= 1) { mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
    if (lane == 0) { s[d][wid] = mn[d]; s[3 + d][wid] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int nw = blockDim.x >> 5; double r = s[threadIdx.x][0];
    for (int k = 1; k < nw; ++k) r = (threadIdx.x < 3) ? fmin(r, s[threadIdx.x][k]) : fmax(r, s[threadIdx.x][k]);
    partial[blockIdx.x * 6 + threadIdx.x] = r;
  }
}

__device__ __forceinline__ int cell_coord(double v, double o, double inv_h, int n) {
  int c = (int)floor((v - o) * inv_h);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void k_cell_count(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z, int n,
                             GridDev g, int* __restrict__ cell_of, int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = cell_coord(x[i], g.ox, g.inv_h, g.nx), cy = cell_coord(y[i], g.oy, g.inv_h, g.ny), cz = cell_coord(z[i], g.oz, g.inv_h, g.nz);
  int c = (cz * g.ny + cy) * g.nx + cx;
  cell_of[i] = c;
  atomicAdd(&counts[c], 1);
}

// exclusive scan of counts[0..m) into out[0..m] in three launches: per-block sums of 4096 counts, scan of the block sums (one
// block), per-block scan with its base.  (A single-block scan of the ~400 k cells of a 100 k-feature grid took 0.4 ms per upload.)
#define LRR_SCAN_ITEMS 4   // counts per thread, 1024 threads per block
__global__ void __launch_bounds__(1024) k_scan_sums(const int* __restrict__ counts, int m, int* __restrict__ bsum) {
  __shared__ int s[32];
  const int base = blockIdx.x * 1024 * LRR_SCAN_ITEMS;
  int v = 0;
#pragma unroll
  for (int u = 0; u < LRR_SCAN_ITEMS; ++u) { const int i = base + u * 1024 + threadIdx.x; if (i < m) v += counts[i]; }
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = s[threadIdx.x];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = v;
  }
}
// in-place exclusive scan of the nb block sums (one block; nb <= 32768), total -> out_total
__global__ void __launch_bounds__(1024) k_scan_bsums(int* __restrict__ bsum, int nb, int* __restrict__ out_total) {
  __shared__ int s[1024];
  const int per = (nb + 1023) / 1024, lo = min(threadIdx.x * per, nb), hi = min(lo + per, nb);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += bsum[i];
  s[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += v;
    __syncthreads();
  }
  int base = s[threadIdx.x] - sum;
  for (int i = lo; i < hi; ++i) { const int c = bsum[i]; bsum[i] = base; base += c; }
  if (threadIdx.x == 1023) *out_total = s[1023];
}
__global__ void __launch_bounds__(1024) k_scan_apply(const int* __restrict__ counts, int m, const int* __restrict__ bsum, int* __restrict__ out) {
  __shared__ int s[32];
  const int base = blockIdx.x * 1024 * LRR_SCAN_ITEMS + threadIdx.x * LRR_SCAN_ITEMS;   // consecutive counts per thread
  int c[LRR_SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int u = 0; u < LRR_SCAN_ITEMS; ++u) { c[u] = (base + u < m) ? counts[base + u] : 0; sum += c[u]; }
  int incl = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = s[lane], wi = w;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
    s[lane] = wi - w;
  }
  __syncthreads();
  int run = bsum[blockIdx.x] + s[wid] + incl - sum;
#pragma unroll
  for (int u = 0; u < LRR_SCAN_ITEMS; ++u) { if (base + u < m) out[base + u] = run; run += c[u]; }
}
static void lrr_exclusive_scan(const int* counts, int m, int* out /* m + 1 */, int* bsum /* blocks */, cudaStream_t st, long long* launches) {
  const int nb = (m + 1024 * LRR_SCAN_ITEMS - 1) / (1024 * LRR_SCAN_ITEMS);
  k_scan_sums<<<nb, 1024, 0, st>>>(counts, m, bsum); ++*launches;
  k_scan_bsums<<<1, 1024, 0, st>>>(bsum, nb, out + m); ++*launches;
  k_scan_apply<<<nb, 1024, 0, st>>>(counts, m, bsum, out); ++*launches;
}

__global__ void k_scatter(const int* __restrict__ cell_of, int n, const int* __restrict__ cell_start, int* __restrict__ fill,
                          int* __restrict__ gorig) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_of[i];
  int pos = cell_start[c] + atomicAdd(&fill[c], 1);
  gorig[pos] = i;
}

// sort each cell's original indices ascending: deterministic layout regardless of atomic order
__global__ void k_cell_sort(const int* __restrict__ cell_start, int ncell, int* __restrict__ gorig) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int lo = cell_start[c], hi = cell_start[c + 1];
  for (int i = lo + 1; i < hi; ++i) {
    int v = gorig[i], j = i - 1;
    while (j >= lo && gorig[j] > v) { gorig[j + 1] = gorig[j]; --j; }
    gorig[j + 1] = v;
  }
}

__global__ void k_fill_pos(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z, int n,
                           const int* __restrict__ gorig, double4* __restrict__ pos4, int* __restrict__ sorted_of_orig) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int o = gorig[j];
  pos4[j] = make_double4(x[o], y[o], z[o], __longlong_as_double((long long)o));
  sorted_of_orig[o] = j;
}

cudaError_t lrr_grid_build(FsetBuild* fb, cudaStream_t st, long long* launches) {
  const int n = (int)fb->n;
  GridDev& g = fb->grid;
  g.n = n;
  double* d_part = nullptr; double h_part[64 * 6];
  cudaError_t e;
  if (n == 0) {
    g.ox = g.oy = g.oz = 0; g.h = 1; g.inv_h = 1; g.nx = g.ny = g.nz = 1;
    if ((e = cudaMallocAsync((void**)&fb->cell_start, 2 * sizeof(int), st)) != cudaSuccess) return e;
    cudaMemsetAsync(fb->cell_start, 0, 2 * sizeof(int), st);
    g.cell_start = fb->cell_start; g.pos4 = nullptr;
    return cudaSuccess;
  }
  if ((e = cudaMallocAsync((void**)&d_part, 64 * 6 * sizeof(double), st)) != cudaSuccess) return e;
  k_bbox<<<64, 256, 0, st>>>(fb->x, fb->y, fb->z, n, d_part); ++*launches;
  if ((e = cudaMemcpyAsync(h_part, d_part, sizeof(h_part), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  cudaFreeAsync(d_part, st);
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int b = 0; b < 64; ++b) for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], h_part[b * 6 + d]); mx[d] = fmax(mx[d], h_part[b * 6 + 3 + d]); }
  double ext[3]; double vol = 1.0; double maxext = 0;
  for (int d = 0; d < 3; ++d) { ext[d] = mx[d] - mn[d]; if (ext[d] > maxext) maxext = ext[d]; }
  if (!(maxext > 0)) maxext = 1.0;
  for (int d = 0; d < 3; ++d) { if (ext[d] < 1e-6 * maxext) ext[d] = 1e-6 * maxext; vol *= ext[d]; }
  // Features lie on surfaces and curves, so most cells stay empty and the occupied ones hold far more than
  // the average: aim at ~4 cells per feature (LRR_GRID_DENSITY overrides, for tuning runs)
  double density = 4.0;
  if (const char* env = getenv("LRR_GRID_DENSITY")) { double v = atof(env); if (v > 0.01 && v <= 16.0) density = v; }
  double target_cells = fmax(1.0, (double)n * density);
  double h = cbrt(vol / target_cells);
  long long nx, ny, nz;
  for (;;) {
    nx = (long long)floor(ext[0] / h) + 1; ny = (long long)floor(ext[1] / h) + 1; nz = (long long)floor(ext[2] / h) + 1;
    if (nx * ny * nz <= 20LL * n + 64 && nx * ny * nz < (1LL << 27)) break;
    h *= 1.25;
  }
  g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.h = h; g.inv_h = 1.0 / h; g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz;
  const int ncell = (int)(nx * ny * nz);
  int *cell_of = nullptr, *counts = nullptr, *gorig = nullptr, *bsum = nullptr;
  if ((e = cudaMallocAsync((void**)&fb->cell_start, (size_t)(ncell + 1) * sizeof(int), st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&cell_of, (size_t)n * sizeof(int), st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&counts, (size_t)ncell * sizeof(int), st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&gorig, (size_t)n * sizeof(int), st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&bsum, (size_t)(ncell / (1024 * LRR_SCAN_ITEMS) + 2) * sizeof(int), st)) != cudaSuccess) return e;
  cudaMemsetAsync(counts, 0, (size_t)ncell * sizeof(int), st);
  const int T = 256, B = (n + T - 1) / T;
  k_cell_count<<<B, T, 0, st>>>(fb->x, fb->y, fb->z, n, g, cell_of, counts); ++*launches;
  lrr_exclusive_scan(counts, ncell, fb->cell_start, bsum, st, launches);
  cudaMemsetAsync(counts, 0, (size_t)ncell * sizeof(int), st);
  k_scatter<<<B, T, 0, st>>>(cell_of, n, fb->cell_start, counts, gorig); ++*launches;
  k_cell_sort<<<(ncell + T - 1) / T, T, 0, st>>>(fb->cell_start, ncell, gorig); ++*launches;
  k_fill_pos<<<B, T, 0, st>>>(fb->x, fb->y, fb->z, n, gorig, fb->pos4, fb->sorted_of_orig); ++*launches;
  e = cudaStreamSynchronize(st);
  cudaFreeAsync(cell_of, st); cudaFreeAsync(counts, st); cudaFreeAsync(gorig, st); cudaFreeAsync(bsum, st);
  g.cell_start = fb->cell_start; g.pos4 = fb->pos4;
  return e != cudaSuccess ? e : cudaGetLastError();
}

// ---- stand-alone query kernel (lrr_features_nearest; CompareMapTree analogue) ------------------------
__global__ void __launch_bounds__(256) k_nn_query(GridDev g, int nq, const double* __restrict__ qx, const double* __restrict__ qy,
                                                  const double* __restrict__ qz, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int s = lrr_nn_search(g, qx[i], qy[i], qz[i]);
  out[i] = (s >= 0) ? lrr_pos4_orig(lrr_ld_pos4(g.pos4 + s)) : -1;
}

cudaError_t lrr_nn_query(const GridDev& g, int64_t nq, const double* qx, const double* qy, const double* qz, int* out,
                         cudaStream_t st, long long* launches) {
  if (nq == 0) return cudaSuccess;
  k_nn_query<<<(int)((nq + 255) / 256), 256, 0, st>>>(g, (int)nq, qx, qy, qz, out); ++*launches;
  return cudaGetLastError();
}
