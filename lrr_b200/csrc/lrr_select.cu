// Sort-based pieces of K3 that run once per registration, not per iteration (sm_100a):
//   * ascending sort of the absolute residuals for the MUSE scale (rrel_muset_obj::scale; call sites
//     cdcl/cdcl_estimation_ICP_matching_all.txx:721-725, rrl/rrl_estimation_ICP_matching_all.txx:630-634);
//   * the MUSE objective over the sorted residuals;
//   * sort-based fallback of the MAD order statistic (cdcl :520-536) when the histogram selector overflows.
// The sort is a single-block stable LSD radix sort on the fp64 bit patterns (non-negative keys order like
// their bit patterns): 8 passes of 8 bits, per-thread contiguous chunks so the scatter is stable.
#include "lrr_internal.h"
#include <math.h>

#define SORT_T 1024

cudaError_t lrr_sort_alloc(SortScratch* s, size_t cap) {
  s->cap = cap;
  cudaError_t e;
  if ((e = cudaMalloc((void**)&s->keys2, (cap ? cap : 1) * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc((void**)&s->hist, 256 * SORT_T * sizeof(unsigned int))) != cudaSuccess) return e;
  return cudaSuccess;
}
void lrr_sort_free(SortScratch* s) { cudaFree(s->keys2); cudaFree(s->hist); s->keys2 = nullptr; s->hist = nullptr; }

__global__ void __launch_bounds__(SORT_T) k_radix_sort(double* __restrict__ keys, double* __restrict__ keys2, unsigned int* __restrict__ hist, int n) {
  __shared__ unsigned int s_part[SORT_T];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (n + T - 1) / T;
  const int lo = min(t * per, n), hi = min(lo + per, n);
  unsigned long long* src = reinterpret_cast<unsigned long long*>(keys);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(keys2);
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 8 * pass;
    for (int i = t; i < 256 * T; i += T) hist[i] = 0;
    __syncthreads();
    for (int i = lo; i < hi; ++i) hist[((src[i] >> shift) & 255u) * T + t] += 1;
    __syncthreads();
    // exclusive scan over hist in (digit major, thread minor) order: thread t owns entries [256 t, 256 t + 256)
    unsigned int sum = 0;
    for (int i = 0; i < 256; ++i) sum += hist[256 * t + i];
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) {
      unsigned int v = (t >= o) ? s_part[t - o] : 0;
      __syncthreads();
      s_part[t] += v;
      __syncthreads();
    }
    unsigned int base = s_part[t] - sum;
    for (int i = 0; i < 256; ++i) { unsigned int c = hist[256 * t + i]; hist[256 * t + i] = base; base += c; }
    __syncthreads();
    for (int i = lo; i < hi; ++i) {
      const unsigned long long k = src[i];
      const unsigned int pos = hist[((k >> shift) & 255u) * T + t]++;
      dst[pos] = k;
    }
    __syncthreads();
    unsigned long long* tmp = src; src = dst; dst = tmp;
  }
  // 8 passes: the result is back in `keys`
}

// n <= 2048: bitonic sort in shared memory (the MUSE call at the start of every registration: 1000 residuals)
__global__ void __launch_bounds__(1024) k_bitonic_small(double* __restrict__ keys, int n) {
  __shared__ double sh[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = (i < n) ? keys[i] : 1e308;
  __syncthreads();
  for (int k = 2; k <= 2048; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const double a = sh[i], b = sh[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { sh[i] = b; sh[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = sh[i];
}

void lrr_sort_doubles(SortScratch* s, double* keys, int n, cudaStream_t st, long long* launches) {
  if (n <= 1) return;
  if (n <= 2048) { k_bitonic_small<<<1, 1024, 0, st>>>(keys, n); ++*launches; return; }
  k_radix_sort<<<1, SORT_T, 0, st>>>(keys, s->keys2, s->hist, n); ++*launches;
}

// table: nfrac rows of {k, expected_kth, stddev_kth, divisor} with k ascending (rows failing the
// expected/stddev >= 3 test are left out by the host).  Restated from Miller & Stewart, CVPR'96 (trimmed MUSE).
__global__ void __launch_bounds__(256) k_muse(const double* __restrict__ sorted, int n, const double* __restrict__ table, int nfrac,
                                              double* __restrict__ out_scale) {
  __shared__ double s_red[256];
  __shared__ double s_seg[32];
  int prev = 0;
  for (int f = 0; f < nfrac && f < 32; ++f) {
    const int k = (int)table[4 * f];
    double a = 0;
    for (int i = prev + threadIdx.x; i < k; i += blockDim.x) a += sorted[i];
    s_red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) s_seg[f] = s_red[0];
    __syncthreads();
    prev = k;
  }
  if (threadIdx.x == 0) {
    double sum = 0, best_sk = 0, best_obj = 0; bool at_start = true;
    for (int f = 0; f < nfrac && f < 32; ++f) {
      sum += s_seg[f];
      const double sk = sum / table[4 * f + 3];
      const double obj = sk * table[4 * f + 2] / table[4 * f + 1];
      if (at_start || obj < best_obj) { at_start = false; best_sk = sk; best_obj = obj; }
    }
    *out_scale = (n == 0) ? 0.0 : best_sk;
  }
}
void lrr_muse_from_sorted(const double* sorted, int n, const double* table, int nfrac, double* out_scale, cudaStream_t st,
                          long long* launches) {
  k_muse<<<1, 256, 0, st>>>(sorted, n, table, nfrac, out_scale); ++*launches;
}

__global__ void k_mad_sorted(const double* __restrict__ sorted, int n, double* __restrict__ out) {
  if (n <= 1) { *out = 0.0; return; }
  const int loc = (n - 1) / 2 + 1;
  *out = 1.4826 * (1.0 + 5.0 / (double)(n - 1)) * sorted[loc];
}
void lrr_mad_from_sorted(const double* sorted, int n, double* out, cudaStream_t st, long long* launches) {
  k_mad_sorted<<<1, 1, 0, st>>>(sorted, n, out); ++*launches;
}
