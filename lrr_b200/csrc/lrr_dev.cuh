// Device-side helpers shared by the kernels (sm_100a).
#pragma once
#include "lrr_internal.h"

// y = A (x - c) + t with the reference's evaluation order (vnl row dot product left to right, then + t;
// cdcl/cdcl_trans_affine.txx:112-115) and NO FMA contraction, so that the mapped point — and therefore
// the neighbour index — is bit-identical to the CPU oracle for the same transform bits.
__device__ __forceinline__ void lrr_map_exact(const double* __restrict__ A, const double* __restrict__ t, const double* __restrict__ c,
                                              double x, double y, double z, double& ox, double& oy, double& oz) {
  double v0 = __dsub_rn(x, c[0]), v1 = __dsub_rn(y, c[1]), v2 = __dsub_rn(z, c[2]);
  ox = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(A[0], v0), __dmul_rn(A[1], v1)), __dmul_rn(A[2], v2)), t[0]);
  oy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(A[3], v0), __dmul_rn(A[4], v1)), __dmul_rn(A[5], v2)), t[1]);
  oz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(A[6], v0), __dmul_rn(A[7], v1)), __dmul_rn(A[8], v2)), t[2]);
}

__device__ __forceinline__ double lrr_dist2_exact(double ax, double ay, double az, double bx, double by, double bz) {
  double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// e^T P e with P symmetric given as (xx,xy,xz,yy,yz,zz), clamped at 0 (cdcl :219-220)
__device__ __forceinline__ double lrr_proj_r2(double ex, double ey, double ez, double pxx, double pxy, double pxz, double pyy,
                                              double pyz, double pzz) {
  double c0 = ex * pxx + ey * pxy + ez * pxz;
  double c1 = ex * pxy + ey * pyy + ez * pyz;
  double c2 = ex * pxz + ey * pyz + ez * pzz;
  double r = c0 * ex + c1 * ey + c2 * ez;
  return r < 0.0 ? 0.0 : r;
}

// Beaton-Tukey weight on a squared, scale-normalised residual (cdcl/cdcl_utils.h:142-152)
__device__ __forceinline__ double lrr_weight_BT(double u) {
  if (u <= 16.0) { double t = 1 - u * 0.0625; return t * t; }
  return 0.0;
}
// cdcl/cdcl_utils.h:158-178 and :184-204
__device__ __forceinline__ double lrr_rho(double r) {
  const double b_2 = 16.0, f = 2.9872, c = b_2 / 6.0;
  if (r > b_2) return f * c;
  double t = 1 - r / b_2;
  return f * c * (1 - t * t * t);
}
__device__ __forceinline__ double lrr_rho_p(double r) {
  const double b_2 = 16.0, f = 2.9872;
  if (r > b_2) return 0.0;
  double t = 1 - r / b_2;
  return f * 0.5 * t * t;
}

__device__ __forceinline__ double4 lrr_ld_pos4(const double4* __restrict__ p) {
  // two 16-byte read-only loads of one 32-byte sector
  const double2* q = reinterpret_cast<const double2*>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ int lrr_pos4_orig(const double4& p) { return (int)__double_as_longlong(p.w); }

// scan points [lo,hi) of the cell-sorted position array
__device__ __forceinline__ void lrr_nn_scan(const double4* __restrict__ pos4, int lo, int hi, double qx, double qy, double qz,
                                            double& best, int& best_s, int& best_o) {
  for (int j = lo; j < hi; ++j) {
    double4 p = lrr_ld_pos4(pos4 + j);
    double d = lrr_dist2_exact(qx, qy, qz, p.x, p.y, p.z);
    if (d < best) { best = d; best_s = j; best_o = lrr_pos4_orig(p); }
    else if (d == best) { int o = lrr_pos4_orig(p); if (o < best_o) { best_s = j; best_o = o; } }
  }
}

// Exact 1-NN over the uniform grid: Chebyshev shells around the query's (clamped) cell.  After shell r
// every unvisited point lies outside the (2r+1)^3 block, i.e. at least `bound` away (distance from the
// query to the nearest block face that still has cells behind it).  Termination requires
// best < (bound - slack)^2 strictly, slack >> the rounding of the cell assignment, so a skipped point can
// never tie or beat the result: the answer is exact, ties resolved to the lowest original index.
// Returns the SORTED slot of the neighbour (-1 for an empty set).
__device__ __forceinline__ int lrr_nn_search(const GridDev& g, double qx, double qy, double qz) {
  if (g.n <= 0) return -1;
  int cx = (int)floor((qx - g.ox) * g.inv_h), cy = (int)floor((qy - g.oy) * g.inv_h), cz = (int)floor((qz - g.oz) * g.inv_h);
  cx = cx < 0 ? 0 : (cx >= g.nx ? g.nx - 1 : cx);
  cy = cy < 0 ? 0 : (cy >= g.ny ? g.ny - 1 : cy);
  cz = cz < 0 ? 0 : (cz >= g.nz ? g.nz - 1 : cz);
  double best = 1e300; int best_s = -1, best_o = 0x7fffffff;
  const double slack = 1e-9 * g.h;
  const int rmax = max(max(g.nx, g.ny), g.nz);
  for (int r = 0; r <= rmax; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
    for (int z = z0; z <= z1; ++z) {
      const bool zface = (z == cz - r) || (z == cz + r);
      for (int y = y0; y <= y1; ++y) {
        const bool face = zface || (y == cy - r) || (y == cy + r);
        const int row = (z * g.ny + y) * g.nx;
        if (face || r == 0) {
          // whole x-run [x0,x1] is new: x-adjacent cells are contiguous in the sorted order
          lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + x0), __ldg(g.cell_start + row + x1 + 1), qx, qy, qz, best, best_s, best_o);
        } else {
          if (cx - r >= 0) lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + cx - r), __ldg(g.cell_start + row + cx - r + 1), qx, qy, qz, best, best_s, best_o);
          if (cx + r < g.nx) lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + cx + r), __ldg(g.cell_start + row + cx + r + 1), qx, qy, qz, best, best_s, best_o);
        }
      }
    }
    // distance to the nearest face of the visited block that still has unvisited cells behind it
    double bound = 1e300; bool any = false;
    if (cx - r > 0) { bound = fmin(bound, qx - (g.ox + (cx - r) * g.h)); any = true; }
    if (cx + r + 1 < g.nx) { bound = fmin(bound, (g.ox + (cx + r + 1) * g.h) - qx); any = true; }
    if (cy - r > 0) { bound = fmin(bound, qy - (g.oy + (cy - r) * g.h)); any = true; }
    if (cy + r + 1 < g.ny) { bound = fmin(bound, (g.oy + (cy + r + 1) * g.h) - qy); any = true; }
    if (cz - r > 0) { bound = fmin(bound, qz - (g.oz + (cz - r) * g.h)); any = true; }
    if (cz + r + 1 < g.nz) { bound = fmin(bound, (g.oz + (cz + r + 1) * g.h) - qz); any = true; }
    if (!any) break;                      // whole grid visited
    bound -= slack;
    if (best_s >= 0 && bound > 0.0 && best < bound * bound) break;
  }
  return best_s;
}

// ---- reduction scopes ------------------------------------------------------------------------------
// A phase body is written once against a Scope: `first()/stride()` drive the element loop, `reduce<K>`
// sums K doubles per thread over the scope.  It returns true in every thread of the FINALIZER block with
// the totals in shared memory; the finalizer then writes the derived scalars to Scal.
//   GridScope : one kernel over the whole GPU.  Warp shuffles, shared memory across warps, one row of
//               `partials` per block; the last block to finish (atomic ticket) adds the rows in fixed
//               order, so results are reproducible run to run.
//   BlockScope: one thread block owns the whole registration (batched mode): block reduction only.
struct GridScope {
  double* partials; unsigned int* ticket;
  __device__ __forceinline__ int first() const { return blockIdx.x * blockDim.x + threadIdx.x; }
  __device__ __forceinline__ int stride() const { return gridDim.x * blockDim.x; }
  // G = lanes per item: lanes with equal (lane % G) are summed; outputs are K*G values, index g*K + k.
  template <int K, int G = 1>
  __device__ __forceinline__ bool reduce(double (&v)[K], double* sh_out /* >= K*G */, double* sh_tmp /* >= K*G*nwarps */) const {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int KG = K * G;
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
      partials[(size_t)blockIdx.x * LRR_NRED + k] = s;
    }
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = atomicAdd(ticket, 1u);
      is_last = (t == gridDim.x - 1);
      if (is_last) *ticket = 0;   // re-arm for the next launch
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    for (int k = threadIdx.x; k < KG; k += blockDim.x) {
      double s = 0;
      for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * LRR_NRED + k]);
      sh_out[k] = s;
    }
    __syncthreads();
    return true;
  }
  // ticket only (no sums): true in every thread of the last block
  __device__ __forceinline__ bool last_block() const {
    __shared__ bool is_last2;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = atomicAdd(ticket, 1u);
      is_last2 = (t == gridDim.x - 1);
      if (is_last2) *ticket = 0;
    }
    __syncthreads();
    if (is_last2) __threadfence();
    return is_last2;
  }
  __device__ __forceinline__ void phase_end() const {}
};

struct BlockScope {
  __device__ __forceinline__ int first() const { return threadIdx.x; }
  __device__ __forceinline__ int stride() const { return blockDim.x; }
  template <int K, int G = 1>
  __device__ __forceinline__ bool reduce(double (&v)[K], double* sh_out, double* sh_tmp) const {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int KG = K * G;
    __syncthreads();   // sh_tmp / sh_out may still be read by the previous phase
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
      sh_out[k] = s;
    }
    __syncthreads();
    return true;
  }
  __device__ __forceinline__ bool last_block() const { __syncthreads(); return true; }
  // make the finalizer's global writes visible to the whole block before the next phase
  __device__ __forceinline__ void phase_end() const { __threadfence_block(); __syncthreads(); }
};
