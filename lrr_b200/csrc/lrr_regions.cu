// FeaturesInRegions (rrl/itkLocationRegistration.cxx:1363-1522) for one volume, on the device (sm_100a): from the features inside
// the ROI cube, sorted by their distance from the query point (:1297-1311,1460-1464), the over-segmentation regions are taken in the
// order in which their first (nearest) feature appears (:1478-1483) and ALL features of a region are added, region after region,
// until at least `min_count` features (1500, :1487) are in the set.  The region label of a feature is the caller's watershed label
// image looked up at its location (:1407) — an input here, uploaded per feature set.
// A stream compaction over the resident cell-ordered SoA (the cube = a few runs of the uniform grid) followed by block-level sorts
// and scans in shared memory; one thread block per call (a few thousand candidates, once per candidate registration).
//   1  cube -> (squared distance, original index) of every feature inside, compacted
//   2  bitonic sort by (distance, index)                                  [std::sort by distance; ties: lowest index]
//   3  (label, sorted position) sorted -> runs of equal labels; a run's head holds the label's first appearance
//   4  rank k of a label = number of labels that appear earlier; prefix sums of the run lengths in rank order -> the number of labels
//      the growth loop takes; output position of a feature = (features of earlier labels) + (its position inside its run)
#include "lrr_dev.cuh"
#include "lrr_host.h"

#define REG_CAP 8192          // max features inside the cube
#define REG_T 1024

struct RegShared {
  double d2[REG_CAP];                 // 64 KB
  int idx[REG_CAP];                   // 32 KB : original index, distance order
  unsigned long long key[REG_CAP];    // 64 KB : (label, position) pairs / scratch
  int aux[REG_CAP];                   // 32 KB
  int cnt, nlabels, ntake, total;
};

// inclusive scan of n ints in place (n <= REG_CAP), REG_T threads, each owning a contiguous chunk
__device__ void reg_scan_inclusive(int* a, int n, int* s_part /* REG_T */) {
  const int t = threadIdx.x, per = (n + REG_T - 1) / REG_T, lo = min(t * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) { sum += a[i]; a[i] = sum; }
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < REG_T; o <<= 1) { const int v = t >= o ? s_part[t - o] : 0; __syncthreads(); s_part[t] += v; __syncthreads(); }
  const int base = s_part[t] - sum;
  for (int i = lo; i < hi; ++i) a[i] += base;
  __syncthreads();
}

extern __shared__ __align__(16) unsigned char reg_smem[];
__global__ void __launch_bounds__(REG_T) k_features_in_regions(FsetDev f, const int* __restrict__ labels, double x0, double y0, double z0, double x1, double y1, double z1,
                                                               double rx, double ry, double rz, int min_count, int cap, int* __restrict__ out, int* __restrict__ n_out) {
  RegShared& S = *reinterpret_cast<RegShared*>(reg_smem);
  __shared__ int s_part[REG_T];
  const GridDev& g = f.grid;
  const int t = threadIdx.x;
  if (t == 0) { S.cnt = 0; S.nlabels = 0; S.ntake = 0; S.total = 0; }
  __syncthreads();
  // ---- 1: the cube as runs of the grid ----
  {
    const int cx0 = lrr_cell_clamp(x0, g.ox, g.inv_h, g.nx), cx1 = lrr_cell_clamp(x1, g.ox, g.inv_h, g.nx);
    const int cy0 = lrr_cell_clamp(y0, g.oy, g.inv_h, g.ny), cy1 = lrr_cell_clamp(y1, g.oy, g.inv_h, g.ny);
    const int cz0 = lrr_cell_clamp(z0, g.oz, g.inv_h, g.nz), cz1 = lrr_cell_clamp(z1, g.oz, g.inv_h, g.nz);
    const int ny = cy1 - cy0 + 1, nrows = ny * (cz1 - cz0 + 1);
    const int lane = t & 31, wid = t >> 5, nw = REG_T >> 5;
    for (int r = wid; r < nrows; r += nw) {
      const int z = cz0 + r / ny, y = cy0 + r % ny, row = (z * g.ny + y) * g.nx;
      const int lo = __ldg(g.cell_start + row + cx0), hi = __ldg(g.cell_start + row + cx1 + 1);
      for (int j = lo + lane; j < hi; j += 32) {
        const double4 p = lrr_ld_pos4(g.pos4 + j);
        if (!(x0 > p.x || p.x > x1 || y0 > p.y || p.y > y1 || z0 > p.z || p.z > z1)) {   // the reference's test (:1401)
          const int slot = atomicAdd(&S.cnt, 1);
          if (slot < REG_CAP) { S.d2[slot] = lrr_dist2_exact(p.x, p.y, p.z, rx, ry, rz); S.idx[slot] = lrr_pos4_orig(p); }
        }
      }
    }
  }
  __syncthreads();
  const int n = S.cnt;
  if (n > REG_CAP) { if (t == 0) *n_out = -1; return; }   // more features in the cube than one block takes: reported to the host
  if (n == 0) { if (t == 0) *n_out = 0; return; }
  // ---- 2: sort by (distance, index) ----
  unsigned int np = 1; while ((int)np < n) np <<= 1;
  for (unsigned int i = n + t; i < np; i += REG_T) { S.d2[i] = 1e308; S.idx[i] = 0x7fffffff; }
  __syncthreads();
  for (unsigned int k = 2; k <= np; k <<= 1)
    for (unsigned int jj = k >> 1; jj > 0; jj >>= 1) {
      for (unsigned int i = t; i < np; i += REG_T) {
        const unsigned int ixj = i ^ jj;
        if (ixj > i) {
          const double a = S.d2[i], b = S.d2[ixj]; const int ia = S.idx[i], ib = S.idx[ixj];
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == ((i & k) == 0)) { S.d2[i] = b; S.d2[ixj] = a; S.idx[i] = ib; S.idx[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  // ---- 3: (label, position) pairs sorted: runs of equal labels, heads = first appearances ----
  for (unsigned int i = t; i < np; i += REG_T) S.key[i] = (int)i < n ? (((unsigned long long)(unsigned int)labels[S.idx[i]]) << 32) | i : ~0ull;
  __syncthreads();
  for (unsigned int k = 2; k <= np; k <<= 1)
    for (unsigned int jj = k >> 1; jj > 0; jj >>= 1) {
      for (unsigned int i = t; i < np; i += REG_T) {
        const unsigned int ixj = i ^ jj;
        if (ixj > i) { const unsigned long long a = S.key[i], b = S.key[ixj]; if ((a > b) == ((i & k) == 0)) { S.key[i] = b; S.key[ixj] = a; } }
      }
      __syncthreads();
    }
  // first[pos] = 1 where a label appears for the first time in distance order (the head of its run holds its smallest position)
  for (int i = t; i < n; i += REG_T) S.aux[i] = 0;
  __syncthreads();
  for (int i = t; i < n; i += REG_T) {
    const bool head = i == 0 || (S.key[i] >> 32) != (S.key[i - 1] >> 32);
    if (head) S.aux[(int)(S.key[i] & 0xffffffffull)] = 1;
  }
  __syncthreads();
  reg_scan_inclusive(S.aux, n, s_part);        // aux[pos] = number of labels that first appear at a position <= pos
  // rank of every run (label) = aux[head position] - 1; run lengths by rank -> d2 reused as an int array of counts
  int* count_by_rank = reinterpret_cast<int*>(S.d2);          // n ints
  int* head_of = reinterpret_cast<int*>(S.d2) + REG_CAP;      // n ints: index (in key order) of the head of the run an element belongs to
  for (int i = t; i < n; i += REG_T) { count_by_rank[i] = 0; const bool head = i == 0 || (S.key[i] >> 32) != (S.key[i - 1] >> 32); head_of[i] = head ? i : 0; }
  __syncthreads();
  {   // head_of: inclusive max-scan (heads hold their own index, the others 0)
    const int per = (n + REG_T - 1) / REG_T, lo = min(t * per, n), hi = min(lo + per, n);
    int m = 0;
    for (int i = lo; i < hi; ++i) { m = max(m, head_of[i]); head_of[i] = m; }
    s_part[t] = m;
    __syncthreads();
    for (int o = 1; o < REG_T; o <<= 1) { const int v = t >= o ? s_part[t - o] : 0; __syncthreads(); s_part[t] = max(s_part[t], v); __syncthreads(); }
    const int before = t > 0 ? s_part[t - 1] : 0;
    for (int i = lo; i < hi; ++i) head_of[i] = max(head_of[i], before);
    __syncthreads();
  }
  for (int i = t; i < n; i += REG_T) {
    const int h = head_of[i];
    const int rank = S.aux[(int)(S.key[h] & 0xffffffffull)] - 1;
    atomicAdd(&count_by_rank[rank], 1);
  }
  if (t == 0) S.nlabels = S.aux[n - 1];
  __syncthreads();
  const int nl = S.nlabels;
  reg_scan_inclusive(count_by_rank, nl, s_part);   // features of the labels of rank <= k
  // ---- 4: the growth loop (:1487-1495): labels are taken while the set holds fewer than min_count features ----
  for (int k = t; k < nl; k += REG_T) {
    const int before = k > 0 ? count_by_rank[k - 1] : 0;
    if (before < min_count && (count_by_rank[k] >= min_count || k == nl - 1)) { S.ntake = k + 1; S.total = count_by_rank[k]; }
  }
  __syncthreads();
  const int ntake = S.ntake, total = S.total;
  for (int i = t; i < n; i += REG_T) {
    const int h = head_of[i];
    const int rank = S.aux[(int)(S.key[h] & 0xffffffffull)] - 1;
    if (rank < ntake) {
      const int pos = (rank > 0 ? count_by_rank[rank - 1] : 0) + (i - h);   // inside a run the pairs are in distance order
      if (pos < cap) out[pos] = S.idx[(int)(S.key[i] & 0xffffffffull)];
    }
  }
  if (t == 0) *n_out = total;
}

extern "C" int lrr_features_set_labels(lrr_fset* fs, const int32_t* labels) {
  if (!fs || (fs->n > 0 && !labels)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = fs->ctx;
  cudaSetDevice(ctx->device);
  if (!fs->labels) { if (cudaMallocAsync((void**)&fs->labels, sizeof(int) * (size_t)(fs->n ? fs->n : 1), ctx->stream) != cudaSuccess) { ctx->err = "labels: allocation"; return LRR_ERR_CUDA; } }
  if (cudaMemcpyAsync(fs->labels, labels, sizeof(int) * (size_t)fs->n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "labels: copy"; return LRR_ERR_CUDA;
  }
  return LRR_OK;
}

extern "C" int lrr_features_in_regions(lrr_fset* fs, const double* x0, const double* x1, const double* reference_point, int64_t min_count, int64_t cap,
                                       int32_t* idx_out, int64_t* n_out) {
  if (!fs || !x0 || !x1 || !reference_point || !n_out || cap < 0 || (cap > 0 && !idx_out)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = fs->ctx;
  if (!fs->labels) { ctx->err = "lrr_features_in_regions: no region labels (lrr_features_set_labels)"; return LRR_ERR_STATE; }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  *n_out = 0;
  if (fs->n == 0) return LRR_OK;
  int* dout = nullptr; int* dn = nullptr;
  const size_t ocap = (size_t)std::min<int64_t>(cap, REG_CAP);
  cudaError_t e = cudaMallocAsync((void**)&dout, sizeof(int) * (ocap ? ocap : 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&dn, sizeof(int), st);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_features_in_regions, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RegShared));
  if (e == cudaSuccess) {
    k_features_in_regions<<<1, REG_T, sizeof(RegShared), st>>>(fs->dev, fs->labels, x0[0], x0[1], x0[2], x1[0], x1[1], x1[2], reference_point[0], reference_point[1],
                                                                reference_point[2], (int)std::min<int64_t>(min_count, 0x7fffffff), (int)ocap, dout, dn);
    ++ctx->launches;
    e = cudaGetLastError();
  }
  int n = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n, dn, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && n > 0) { e = cudaMemcpyAsync(idx_out, dout, sizeof(int) * std::min<size_t>((size_t)n, ocap), cudaMemcpyDeviceToHost, st); if (e == cudaSuccess) e = cudaStreamSynchronize(st); }
  if (dout) cudaFreeAsync(dout, st); if (dn) cudaFreeAsync(dn, st);
  if (e != cudaSuccess) { ctx->err = std::string("lrr_features_in_regions: ") + cudaGetErrorString(e); return LRR_ERR_CUDA; }
  if (n < 0) { ctx->err = "lrr_features_in_regions: more than 8192 features inside the cube"; return LRR_ERR_UNSUPPORTED; }
  *n_out = n;
  return LRR_OK;
}
