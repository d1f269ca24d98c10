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

// scan points [lo,hi) of the cell-sorted position array
__device__ __forceinline__ void lrr_nn_scan(const GridDev& g, int lo, int hi, double qx, double qy, double qz, double& best,
                                            int& best_s, int& best_o) {
  for (int j = lo; j < hi; ++j) {
    const double* p = g.gpos + 3 * (size_t)j;
    double d = lrr_dist2_exact(qx, qy, qz, p[0], p[1], p[2]);
    if (d < best) { best = d; best_s = j; best_o = g.gorig[j]; }
    else if (d == best) { int o = g.gorig[j]; if (o < best_o) { best_s = j; best_o = o; } }
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
          lrr_nn_scan(g, g.cell_start[row + x0], g.cell_start[row + x1 + 1], qx, qy, qz, best, best_s, best_o);
        } else {
          if (cx - r >= 0) lrr_nn_scan(g, g.cell_start[row + cx - r], g.cell_start[row + cx - r + 1], qx, qy, qz, best, best_s, best_o);
          if (cx + r < g.nx) lrr_nn_scan(g, g.cell_start[row + cx + r], g.cell_start[row + cx + r + 1], qx, qy, qz, best, best_s, best_o);
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

// ---- deterministic two-level reduction -------------------------------------------------------------
// Each thread contributes K doubles; warp shuffles, then shared memory across warps, then one row of
// `partials` per block.  The last block to finish (atomic ticket) adds the rows in fixed order, so the
// result is reproducible run to run.  Returns true in every thread of the last block; out[K] is then
// valid in shared memory `sh_out` (K <= LRR_NRED).
template <int K>
__device__ __forceinline__ bool lrr_block_reduce(double (&v)[K], double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                 double* sh_out /* >= K */, double* sh_tmp /* >= K*nwarps */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh_tmp[wid * K + k] = x;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += sh_tmp[w * K + k];
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
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double s = 0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * LRR_NRED + k]);
    sh_out[k] = s;
  }
  __syncthreads();
  return true;
}
