// Device-side helpers shared by the kernels (sm_100a).
#pragma once
#include "lrr_internal.h"

// development aid (scripts/timeline.py): per-block %globaltimer marks of the phase kernels of this translation unit.
// Off (null pointer, one uniform load per mark) unless lrr_debug_timeline() armed it.
// Programmatic dependent launch (sm_90+): the kernels of one iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (lrr_launch_pdl, lrr_internal.h).  Each of them first waits for the complete
// predecessor grid (griddepcontrol.wait: all its memory operations are visible afterwards); once a block is through its element loop
// it lets the successor be scheduled (lrr_pdl_trigger, at the start of the final reduction), so the successor's blocks become
// resident while the last block is still adding the partial rows / solving, and the launch latency between two dependent kernels
// overlaps that serial tail.  (Triggering at kernel entry measured SLOWER: the waiting successor blocks take SM slots from the
// later waves of the running kernel.)
__device__ __forceinline__ void lrr_pdl_enter() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void lrr_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static __device__ unsigned long long* g_lrr_tl = nullptr;
__device__ __forceinline__ void lrr_tl(int kern, int slot) {
  unsigned long long* p = g_lrr_tl;
  if (p && threadIdx.x == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p[((size_t)kern * 2048 + blockIdx.y * gridDim.x + blockIdx.x) * 4 + slot] = t;
  }
}

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

// scan points [lo,hi) of the cell-sorted position array; four loads in flight per step
__device__ __forceinline__ void lrr_nn_upd(const double4& p, int j, double qx, double qy, double qz, double& best, int& best_s, int& best_o) {
  const double d = lrr_dist2_exact(qx, qy, qz, p.x, p.y, p.z);
  if (d < best) { best = d; best_s = j; best_o = lrr_pos4_orig(p); }
  else if (d == best) { const int o = lrr_pos4_orig(p); if (o < best_o) { best_s = j; best_o = o; } }
}
__device__ __forceinline__ void lrr_nn_scan(const double4* __restrict__ pos4, int lo, int hi, double qx, double qy, double qz,
                                            double& best, int& best_s, int& best_o) {
  int j = lo;
  for (; j + 4 <= hi; j += 4) {
    const double4 p0 = lrr_ld_pos4(pos4 + j), p1 = lrr_ld_pos4(pos4 + j + 1), p2 = lrr_ld_pos4(pos4 + j + 2), p3 = lrr_ld_pos4(pos4 + j + 3);
    lrr_nn_upd(p0, j, qx, qy, qz, best, best_s, best_o); lrr_nn_upd(p1, j + 1, qx, qy, qz, best, best_s, best_o);
    lrr_nn_upd(p2, j + 2, qx, qy, qz, best, best_s, best_o); lrr_nn_upd(p3, j + 3, qx, qy, qz, best, best_s, best_o);
  }
  if (j + 2 <= hi) {
    const double4 p0 = lrr_ld_pos4(pos4 + j), p1 = lrr_ld_pos4(pos4 + j + 1);
    lrr_nn_upd(p0, j, qx, qy, qz, best, best_s, best_o); lrr_nn_upd(p1, j + 1, qx, qy, qz, best, best_s, best_o);
    j += 2;
  }
  if (j < hi) { const double4 p0 = lrr_ld_pos4(pos4 + j); lrr_nn_upd(p0, j, qx, qy, qz, best, best_s, best_o); }
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

// ---- search used by the MATCH phase: bounded ball scan ------------------------------------------------------
// ICP moves the transform a little per iteration, so the neighbour found for a moving feature in the previous
// iteration is an excellent CANDIDATE for the next one.  Any candidate gives an upper bound on the squared
// nearest-neighbour distance; the exact neighbour is then the minimum over the cells that intersect the ball
// of that radius around the query — usually one to four x-runs of the cell-sorted array instead of 27 cells.
// Exactness: a fixed point p with |p-q|^2 <= best lies in a cell whose (y,z) row intersects the ball and whose
// x index is inside [floor((qx-rx-ox)/h), floor((qx+rx-ox)/h)] for rx^2 = best - (row distance)^2; all radii
// are inflated by a slack far above the rounding of the cell assignment, rows are pruned only when their
// slab distance is STRICTLY above the current best, and ties go to the lowest original index.
//   pass A (one thread per query): queries whose ball touches at most LRR_BALL_ROWS rows;
//   pass B (one warp per query): everything else — queries without a candidate first expand Chebyshev
//   shells until a point is found, then the lanes split the rows of the ball.
#ifndef LRR_BALL_ROWS
#define LRR_BALL_ROWS 9
#endif

struct NNState { double best; int best_s, best_o; };

__device__ __forceinline__ int lrr_cell_clamp(double v, double o, double inv_h, int n) {
  int c = (int)floor((v - o) * inv_h);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
// distance from coordinate v to the slab [o + c h, o + (c+1) h], shrunk by the slack
__device__ __forceinline__ double lrr_slab_dist(double v, double o, double h, int c, double slack) {
  const double lo = o + c * h, hi = lo + h;
  double d = v < lo ? lo - v : (v > hi ? v - hi : 0.0);
  d -= slack;
  return d > 0.0 ? d : 0.0;
}
// scan one (y,z) row of the ball: the x-run within sqrt(best - dyz2) of the query
__device__ __forceinline__ void lrr_ball_row(const GridDev& g, int y, int z, double dyz2, double qx, double qy, double qz, double slack,
                                             double& best, int& best_s, int& best_o) {
  const double rx = sqrt(best - dyz2) * (1.0 + 1e-9) + slack;
  const int xa = lrr_cell_clamp(qx - rx, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + rx, g.ox, g.inv_h, g.nx);
  const int row = (z * g.ny + y) * g.nx;
  lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + xa), __ldg(g.cell_start + row + xb + 1), qx, qy, qz, best, best_s, best_o);
}

// pass A.  st holds a candidate (best_s >= 0).  Returns false when the ball touches too many rows.
// The x-range of the ball's bounding box is used for every row (no square root / floor per row: for the one to four
// rows of a typical ball the extra cells cost less than the per-row arithmetic did).
__device__ __forceinline__ bool lrr_nn_ball_thread(const GridDev& g, double qx, double qy, double qz, NNState& st, int max_rows = LRR_BALL_ROWS) {
  const double slack = 1e-9 * g.h;
  const double R = sqrt(st.best) * (1.0 + 1e-9) + slack;
  const int z0 = lrr_cell_clamp(qz - R, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + R, g.oz, g.inv_h, g.nz);
  const int y0 = lrr_cell_clamp(qy - R, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + R, g.oy, g.inv_h, g.ny);
  if ((z1 - z0 + 1) * (y1 - y0 + 1) > max_rows) return false;
  const int xa = lrr_cell_clamp(qx - R, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + R, g.ox, g.inv_h, g.nx);
  for (int z = z0; z <= z1; ++z) {
    const double dz = lrr_slab_dist(qz, g.oz, g.h, z, slack), dz2 = dz * dz;
    if (dz2 > st.best) continue;
    for (int y = y0; y <= y1; ++y) {
      const double dy = lrr_slab_dist(qy, g.oy, g.h, y, slack);
      if (dz2 + dy * dy > st.best) continue;
      const int row = (z * g.ny + y) * g.nx;
      lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + xa), __ldg(g.cell_start + row + xb + 1), qx, qy, qz, st.best, st.best_s, st.best_o);
    }
  }
  return true;
}

// pass A for a query WITHOUT a candidate (first use of a from-feature): the 27 cells around the query, one thread.
// Returns 1 when that already proves the neighbour (best closer than the block boundary), 0 when a candidate was
// found but cells beyond the block may hold a closer point, -1 when the block is empty.
__device__ __forceinline__ int lrr_nn_block27_thread(const GridDev& g, double qx, double qy, double qz, NNState& st) {
  st.best = 1e300; st.best_s = -1; st.best_o = 0x7fffffff;
  const int cx = lrr_cell_clamp(qx, g.ox, g.inv_h, g.nx), cy = lrr_cell_clamp(qy, g.oy, g.inv_h, g.ny), cz = lrr_cell_clamp(qz, g.oz, g.inv_h, g.nz);
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
  for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); ++z)
    for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); ++y) {
      const int row = (z * g.ny + y) * g.nx;
      lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + x0), __ldg(g.cell_start + row + x1 + 1), qx, qy, qz, st.best, st.best_s, st.best_o);
    }
  if (st.best_s < 0) return -1;
  // distance to the nearest face of the 3x3x3 block that still has cells behind it
  double bound = 1e300; bool any = false;
  if (cx - 1 > 0) { bound = fmin(bound, qx - (g.ox + (cx - 1) * g.h)); any = true; }
  if (cx + 2 < g.nx) { bound = fmin(bound, (g.ox + (cx + 2) * g.h) - qx); any = true; }
  if (cy - 1 > 0) { bound = fmin(bound, qy - (g.oy + (cy - 1) * g.h)); any = true; }
  if (cy + 2 < g.ny) { bound = fmin(bound, (g.oy + (cy + 2) * g.h) - qy); any = true; }
  if (cz - 1 > 0) { bound = fmin(bound, qz - (g.oz + (cz - 1) * g.h)); any = true; }
  if (cz + 2 < g.nz) { bound = fmin(bound, (g.oz + (cz + 2) * g.h) - qz); any = true; }
  if (!any) return 1;
  bound -= 1e-9 * g.h;
  return (bound > 0.0 && st.best < bound * bound) ? 1 : 0;
}

__device__ __forceinline__ void lrr_warp_min(double& best, int& best_s, int& best_o) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int os = __shfl_xor_sync(0xffffffffu, best_s, o), oo = __shfl_xor_sync(0xffffffffu, best_o, o);
    if (ob < best || (ob == best && oo < best_o)) { best = ob; best_s = os; best_o = oo; }
  }
}

// Whole warp, one query without a candidate: Chebyshev shells around the query's cell until one holds a point (any point will do as
// the candidate that bounds the ball scan).  Returns false for an empty set.
__device__ __forceinline__ bool lrr_nn_shell_first(const GridDev& g, double qx, double qy, double qz, double& best, int& best_s, int& best_o) {
  const int lane = threadIdx.x & 31;
  best = 1e300; best_s = -1; best_o = 0x7fffffff;
  if (g.n <= 0) return false;
  const int cx = lrr_cell_clamp(qx, g.ox, g.inv_h, g.nx), cy = lrr_cell_clamp(qy, g.oy, g.inv_h, g.ny), cz = lrr_cell_clamp(qz, g.oz, g.inv_h, g.nz);
  const int rmax = max(max(g.nx, g.ny), g.nz);
  for (int r = 0; r <= rmax; ++r) {
    const int side = 2 * r + 1, nface = 2 * side, per_plane = 4 * r, nruns = nface + (side - 2) * per_plane;
    for (int t = lane; t < nruns; t += 32) {
      int z, y, xa, xb;
      if (t < nface) { z = (t < side) ? cz - r : cz + r; y = cy - r + (t % side); xa = cx - r; xb = cx + r; }
      else {
        const int u = t - nface, zi = u / per_plane, v = u % per_plane;
        z = cz - r + 1 + zi;
        if (v < 2) { y = (v == 0) ? cy - r : cy + r; xa = cx - r; xb = cx + r; }
        else { const int m = v - 2; y = cy - r + 1 + (m >> 1); xa = xb = (m & 1) ? cx + r : cx - r; }
      }
      if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) continue;
      xa = max(xa, 0); xb = min(xb, g.nx - 1);
      if (xa > xb) continue;
      const int row = (z * g.ny + y) * g.nx;
      lrr_nn_scan(g.pos4, __ldg(g.cell_start + row + xa), __ldg(g.cell_start + row + xb + 1), qx, qy, qz, best, best_s, best_o);
    }
    lrr_warp_min(best, best_s, best_o);
    if (best_s >= 0) break;
  }
  return best_s >= 0;
}

// pass B: whole warp, one query; st identical in every lane (best_s < 0: no candidate yet).
// Returns the sorted slot of the exact nearest neighbour (-1 for an empty set).
__device__ __forceinline__ int lrr_nn_pass_b(const GridDev& g, double qx, double qy, double qz, NNState st) {
  if (g.n <= 0) return -1;
  const int lane = threadIdx.x & 31;
  const double slack = 1e-9 * g.h;
  double best = st.best; int best_s = st.best_s, best_o = st.best_o;
  if (best_s < 0 && !lrr_nn_shell_first(g, qx, qy, qz, best, best_s, best_o)) return -1;
  // ball scan: the lanes split the rows
  const double R = sqrt(best) * (1.0 + 1e-9) + slack;
  const int z0 = lrr_cell_clamp(qz - R, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + R, g.oz, g.inv_h, g.nz);
  const int y0 = lrr_cell_clamp(qy - R, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + R, g.oy, g.inv_h, g.ny);
  const int ny = y1 - y0 + 1, nrows = (z1 - z0 + 1) * ny;
  for (int t = lane; t < nrows; t += 32) {
    const int z = z0 + t / ny, y = y0 + t % ny;
    const double dz = lrr_slab_dist(qz, g.oz, g.h, z, slack), dy = lrr_slab_dist(qy, g.oy, g.h, y, slack);
    const double dyz2 = dz * dz + dy * dy;
    if (dyz2 > best) continue;
    lrr_ball_row(g, y, z, dyz2, qx, qy, qz, slack, best, best_s, best_o);
  }
  lrr_warp_min(best, best_s, best_o);
  return best_s;
}

// pass B with a NEIGHBOUR CERTIFICATE (DirArgs::proof): the same exact search over the ball inflated by `margin`, which also
// yields the smallest distance of any OTHER point inside it.  Every point within (final distance + margin) of the query is
// examined (each lane prunes with its own current best >= the final one), so afterwards
//     bound = min( second smallest distance, final distance + margin )
// is a lower bound on the distance from the query to every feature but the winner.  `*bound_out` <= 0: no certificate (tie).
__device__ __forceinline__ void lrr_nn_upd2(const double4& p, int j, double qx, double qy, double qz, double& best, int& best_s, int& best_o, double& second) {
  if (j == best_s) return;   // the incoming candidate is scanned again by the lane that owns its row
  const double d = lrr_dist2_exact(qx, qy, qz, p.x, p.y, p.z);
  const int o = lrr_pos4_orig(p);
  if (d < best || (d == best && o < best_o)) { second = best; best = d; best_s = j; best_o = o; }
  else second = fmin(second, d);
}
// one thread, candidate known (st.best_s >= 0): the inflated ball as in lrr_nn_pass_b_cert.  Returns false when the ball touches too
// many rows for one thread (the caller queues the query for pass B).
__device__ __forceinline__ bool lrr_nn_ball_thread_cert(const GridDev& g, double qx, double qy, double qz, NNState& st, double margin, double* bound_out, int max_rows = LRR_BALL_ROWS) {
  *bound_out = 0.0;
  const double slack = 1e-9 * g.h;
  const double R = sqrt(st.best) * (1.0 + 1e-9) + slack + margin;
  const int z0 = lrr_cell_clamp(qz - R, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + R, g.oz, g.inv_h, g.nz);
  const int y0 = lrr_cell_clamp(qy - R, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + R, g.oy, g.inv_h, g.ny);
  if ((z1 - z0 + 1) * (y1 - y0 + 1) > max_rows) return false;
  double second = 1e300;
  for (int z = z0; z <= z1; ++z) {
    const double dz = lrr_slab_dist(qz, g.oz, g.h, z, slack), dz2 = dz * dz;
    for (int y = y0; y <= y1; ++y) {
      const double dy = lrr_slab_dist(qy, g.oy, g.h, y, slack);
      const double dyz2 = dz2 + dy * dy;
      const double Rm = sqrt(st.best) * (1.0 + 1e-9) + slack + margin;
      if (dyz2 > Rm * Rm) continue;
      const double rx = sqrt(Rm * Rm - dyz2) * (1.0 + 1e-9) + slack;
      const int xa = lrr_cell_clamp(qx - rx, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + rx, g.ox, g.inv_h, g.nx);
      const int row = (z * g.ny + y) * g.nx;
      const int lo = __ldg(g.cell_start + row + xa), hi = __ldg(g.cell_start + row + xb + 1);
      int j = lo;
      for (; j + 2 <= hi; j += 2) {
        const double4 p0 = lrr_ld_pos4(g.pos4 + j), p1 = lrr_ld_pos4(g.pos4 + j + 1);
        lrr_nn_upd2(p0, j, qx, qy, qz, st.best, st.best_s, st.best_o, second); lrr_nn_upd2(p1, j + 1, qx, qy, qz, st.best, st.best_s, st.best_o, second);
      }
      if (j < hi) { const double4 p0 = lrr_ld_pos4(g.pos4 + j); lrr_nn_upd2(p0, j, qx, qy, qz, st.best, st.best_s, st.best_o, second); }
    }
  }
  const double bound = fmin(sqrt(second), sqrt(st.best) + margin) * (1.0 - 1e-9) - slack;
  *bound_out = (second > st.best && bound > 0.0) ? bound : 0.0;
  return true;
}
__device__ __forceinline__ int lrr_nn_pass_b_cert(const GridDev& g, double qx, double qy, double qz, NNState st, double margin, double* bound_out) {
  *bound_out = 0.0;
  if (g.n <= 0) return -1;
  if (st.best_s < 0) return lrr_nn_pass_b(g, qx, qy, qz, st);   // no candidate yet: plain search, certificate next time
  const int lane = threadIdx.x & 31;
  const double slack = 1e-9 * g.h;
  double best = st.best; int best_s = st.best_s, best_o = st.best_o; double second = 1e300;
  const double R = sqrt(best) * (1.0 + 1e-9) + slack + margin;
  const int z0 = lrr_cell_clamp(qz - R, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + R, g.oz, g.inv_h, g.nz);
  const int y0 = lrr_cell_clamp(qy - R, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + R, g.oy, g.inv_h, g.ny);
  const int ny = y1 - y0 + 1, nrows = (z1 - z0 + 1) * ny;
  for (int t = lane; t < nrows; t += 32) {
    const int z = z0 + t / ny, y = y0 + t % ny;
    const double dz = lrr_slab_dist(qz, g.oz, g.h, z, slack), dy = lrr_slab_dist(qy, g.oy, g.h, y, slack);
    const double dyz2 = dz * dz + dy * dy;
    const double Rm = sqrt(best) * (1.0 + 1e-9) + slack + margin;   // inflated radius at this lane's current best
    if (dyz2 > Rm * Rm) continue;
    const double rx = sqrt(Rm * Rm - dyz2) * (1.0 + 1e-9) + slack;
    const int xa = lrr_cell_clamp(qx - rx, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + rx, g.ox, g.inv_h, g.nx);
    const int row = (z * g.ny + y) * g.nx;
    const int lo = __ldg(g.cell_start + row + xa), hi = __ldg(g.cell_start + row + xb + 1);
    int j = lo;
    for (; j + 4 <= hi; j += 4) {
      const double4 p0 = lrr_ld_pos4(g.pos4 + j), p1 = lrr_ld_pos4(g.pos4 + j + 1), p2 = lrr_ld_pos4(g.pos4 + j + 2), p3 = lrr_ld_pos4(g.pos4 + j + 3);
      lrr_nn_upd2(p0, j, qx, qy, qz, best, best_s, best_o, second); lrr_nn_upd2(p1, j + 1, qx, qy, qz, best, best_s, best_o, second);
      lrr_nn_upd2(p2, j + 2, qx, qy, qz, best, best_s, best_o, second); lrr_nn_upd2(p3, j + 3, qx, qy, qz, best, best_s, best_o, second);
    }
    for (; j < hi; ++j) { const double4 p0 = lrr_ld_pos4(g.pos4 + j); lrr_nn_upd2(p0, j, qx, qy, qz, best, best_s, best_o, second); }
  }
  double gb = best; int gs = best_s, go = best_o;
  lrr_warp_min(gb, gs, go);
  // lanes whose winner is the global one contribute their second; the others' best is a different point (or the shared incoming
  // candidate, which then lost and is a different point as well)
  double sec = (best_s == gs) ? second : best;
#pragma unroll
  for (int o = 16; o; o >>= 1) sec = fmin(sec, __shfl_xor_sync(0xffffffffu, sec, o));
  const double bound = fmin(sqrt(sec), sqrt(gb) + margin) * (1.0 - 1e-9) - slack;
  *bound_out = (sec > gb && bound > 0.0) ? bound : 0.0;
  return gs;
}

// FLATTENED warp-cooperative ball scan (whole warp, one query, candidate known: st.best_s >= 0; st identical in every lane).
// lrr_nn_pass_b_cert gives every lane its own rows: a far outlier's ball touches a few hundred rows of three or four points each and
// every lane walks a chain of dependent loads (run bounds -> points -> next row), 15-25 us per query.  Here the run bounds of up to
// LRR_FLAT_ROWS rows are fetched together (four independent loads per lane), a warp prefix sum lays the runs end to end in the
// warp's row table (shared memory: `rowlo`, `rowbase`, LRR_FLAT_ROWS + 1 entries each) and the lanes stride over the flattened
// points with four independent loads in flight: two round trips per LRR_FLAT_ROWS rows, whatever the shape of the ball.
// The radius is fixed by the candidate (the neighbour of the previous iteration): every point within sqrt(st.best) + margin of the
// query lies in a scanned run — the exactness argument of lrr_nn_pass_b_cert with no pruning at all.  Returns the sorted slot of
// the exact neighbour (ties: lowest original index) and the certificate bound (0: none).
#ifndef LRR_FLAT_ROWS
#define LRR_FLAT_ROWS 128
#endif
__device__ __forceinline__ int lrr_nn_ball_flat(const GridDev& g, double qx, double qy, double qz, NNState st, double margin, double* bound_out,
                                                int* rowlo, int* rowbase) {
  *bound_out = 0.0;
  const int lane = threadIdx.x & 31;
  const unsigned int full = 0xffffffffu;
  const double slack = 1e-9 * g.h;
  double best = st.best; int best_s = st.best_s, best_o = st.best_o; double second = 1e300;
  const double Rm = sqrt(st.best) * (1.0 + 1e-9) + slack + margin, Rm2 = Rm * Rm;
  const int z0 = lrr_cell_clamp(qz - Rm, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + Rm, g.oz, g.inv_h, g.nz);
  const int y0 = lrr_cell_clamp(qy - Rm, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + Rm, g.oy, g.inv_h, g.ny);
  const int ny = y1 - y0 + 1, nrows = (z1 - z0 + 1) * ny;
  for (int r0 = 0; r0 < nrows; r0 += LRR_FLAT_ROWS) {
    const int nb = min(LRR_FLAT_ROWS, nrows - r0);
    int lo4[LRR_FLAT_ROWS / 32], sz4[LRR_FLAT_ROWS / 32];
#pragma unroll
    for (int u = 0; u < LRR_FLAT_ROWS / 32; ++u) {
      const int rl = 32 * u + lane;
      lo4[u] = 0; sz4[u] = 0;
      if (rl < nb) {
        const int r = r0 + rl, z = z0 + r / ny, y = y0 + r % ny;
        const double dz = lrr_slab_dist(qz, g.oz, g.h, z, slack), dy = lrr_slab_dist(qy, g.oy, g.h, y, slack);
        const double dyz2 = dz * dz + dy * dy;
        if (dyz2 <= Rm2) {
          const double rx = sqrt(Rm2 - dyz2) * (1.0 + 1e-9) + slack;
          const int xa = lrr_cell_clamp(qx - rx, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + rx, g.ox, g.inv_h, g.nx);
          const int row = (z * g.ny + y) * g.nx;
          lo4[u] = __ldg(g.cell_start + row + xa); sz4[u] = __ldg(g.cell_start + row + xb + 1);
        }
      }
    }
    int total = 0;
    __syncwarp();   // the previous batch's points are done with the table
#pragma unroll
    for (int u = 0; u < LRR_FLAT_ROWS / 32; ++u) {
      if (32 * u < nb) {   // warp-uniform
        const int rl = 32 * u + lane, sz = (rl < nb) ? sz4[u] - lo4[u] : 0;
        int inc = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(full, inc, o); if (lane >= o) inc += v; }
        if (rl < nb) { rowlo[rl] = lo4[u]; rowbase[rl] = total + inc - sz; }
        total += __shfl_sync(full, inc, 31);
      }
    }
    if (lane == 0) rowbase[nb] = total;
    if (g_lrr_tl && lane == 0) { unsigned long long* c = g_lrr_tl + 3 * 2048 * 4; if (r0 == 0) atomicAdd(c + 25, 1ull); atomicAdd(c + 26, (unsigned long long)nb); atomicAdd(c + 27, (unsigned long long)total); }
    __syncwarp();
    for (int p0 = 0; p0 < total; p0 += 128) {
      double4 P[4]; int J[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + 32 * u + lane;
        J[u] = -1;
        if (p < total) {
          int ra = 0, rb = nb;   // last row with rowbase <= p
          while (rb - ra > 1) { const int mid = (ra + rb) >> 1; if (rowbase[mid] <= p) ra = mid; else rb = mid; }
          J[u] = rowlo[ra] + (p - rowbase[ra]);
          P[u] = lrr_ld_pos4(g.pos4 + J[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) if (J[u] >= 0) lrr_nn_upd2(P[u], J[u], qx, qy, qz, best, best_s, best_o, second);
    }
  }
  __syncwarp();
  double gb = best; int gs = best_s, go = best_o;
  lrr_warp_min(gb, gs, go);
  double sec = (best_s == gs) ? second : best;
#pragma unroll
  for (int o = 16; o; o >>= 1) sec = fmin(sec, __shfl_xor_sync(full, sec, o));
  const double bound = fmin(sqrt(sec), sqrt(gb) + margin) * (1.0 - 1e-9) - slack;
  *bound_out = (sec > gb && bound > 0.0) ? bound : 0.0;
  return gs;
}

// The flattened ball scan at a quarter of a warp: a group of EIGHT lanes (gmask = its lanes, gl = lane in the group) takes one query
// with a candidate (best, best_s, best_o on entry, identical in the group): run bounds of up to LRR_G8_ROWS rows in one round trip
// (four per lane), a prefix sum over the group into the group's row table (`rowlo`, `rowbase`: LRR_G8_ROWS ints each, shared
// memory), the points in a second round trip, group minimum.  Four queries per warp at a time.  Returns false — nothing changed —
// when the ball touches more rows or points than that (far outliers: the caller gives the query a whole warp); true with the
// exact neighbour (ties: lowest original index) and the certificate bound (0: none) otherwise.
#define LRR_G8_ROWS 32
#define LRR_G8_PTS 256
__device__ __forceinline__ bool lrr_nn_ball_g8(const GridDev& g, double qx, double qy, double qz, double margin, int* rowlo, int* rowbase, unsigned int gmask,
                                               int gl, double& best, int& best_s, int& best_o, double& bound) {
  bound = 0.0;
  const double slack = 1e-9 * g.h;
  const double Rm = sqrt(best) * (1.0 + 1e-9) + slack + margin, Rm2 = Rm * Rm;
  const int z0 = lrr_cell_clamp(qz - Rm, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + Rm, g.oz, g.inv_h, g.nz);
  const int y0 = lrr_cell_clamp(qy - Rm, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + Rm, g.oy, g.inv_h, g.ny);
  const int ny = y1 - y0 + 1, nrows = (z1 - z0 + 1) * ny;
  if (nrows > LRR_G8_ROWS) return false;
  int lo4[4], sz4[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int rl = 8 * u + gl;
    lo4[u] = 0; sz4[u] = 0;
    if (rl < nrows) {
      const int zz = z0 + rl / ny, yy = y0 + rl % ny;
      const double dz = lrr_slab_dist(qz, g.oz, g.h, zz, slack), dy = lrr_slab_dist(qy, g.oy, g.h, yy, slack);
      const double dyz2 = dz * dz + dy * dy;
      if (dyz2 <= Rm2) {
        const double rx = sqrt(Rm2 - dyz2) * (1.0 + 1e-9) + slack;
        const int xa = lrr_cell_clamp(qx - rx, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + rx, g.ox, g.inv_h, g.nx);
        const int row = (zz * g.ny + yy) * g.nx;
        lo4[u] = __ldg(g.cell_start + row + xa); sz4[u] = __ldg(g.cell_start + row + xb + 1);
      }
    }
  }
  int total = 0;
  __syncwarp(gmask);   // the group's previous query is done with the table
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (8 * u < nrows) {   // uniform in the group
      const int rl = 8 * u + gl, sz = (rl < nrows) ? sz4[u] - lo4[u] : 0;
      int inc = sz;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { const int v = __shfl_up_sync(gmask, inc, o, 8); if (gl >= o) inc += v; }
      if (rl < nrows) { rowlo[rl] = lo4[u]; rowbase[rl] = total + inc - sz; }
      total += __shfl_sync(gmask, inc, 7, 8);
    }
  }
  __syncwarp(gmask);
  if (total > LRR_G8_PTS) return false;
  double second = 1e300;
  for (int p0 = 0; p0 < total; p0 += 32) {
    double4 P[4]; int J[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + 8 * u + gl;
      J[u] = -1;
      if (p < total) {
        int ra = 0, rb = nrows;   // last row with rowbase <= p
        while (rb - ra > 1) { const int mid = (ra + rb) >> 1; if (rowbase[mid] <= p) ra = mid; else rb = mid; }
        J[u] = rowlo[ra] + (p - rowbase[ra]);
        P[u] = lrr_ld_pos4(g.pos4 + J[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) if (J[u] >= 0) lrr_nn_upd2(P[u], J[u], qx, qy, qz, best, best_s, best_o, second);
  }
  double gb = best; int gs = best_s, go = best_o;
#pragma unroll
  for (int o = 4; o; o >>= 1) {
    const double ob = __shfl_xor_sync(gmask, gb, o, 8);
    const int os = __shfl_xor_sync(gmask, gs, o, 8), oo = __shfl_xor_sync(gmask, go, o, 8);
    if (ob < gb || (ob == gb && oo < go)) { gb = ob; gs = os; go = oo; }
  }
  double sec = (best_s == gs) ? second : best;
#pragma unroll
  for (int o = 4; o; o >>= 1) sec = fmin(sec, __shfl_xor_sync(gmask, sec, o, 8));
  const double bd = fmin(sqrt(sec), sqrt(gb) + margin) * (1.0 - 1e-9) - slack;
  bound = (sec > gb && bd > 0.0) ? bd : 0.0;
  best = gb; best_s = gs; best_o = go;
  return true;
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
  __device__ __forceinline__ int block_first() const { return blockIdx.x * blockDim.x; }
  // G = lanes per item: lanes with equal (lane % G) are summed; outputs are K*G values, index g*K + k.
  template <int K, int G = 1>
  __device__ __forceinline__ bool reduce(double (&v)[K], double* sh_out /* >= K*G */, double* sh_tmp /* >= K*G*nwarps */) const {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int KG = K * G;
    lrr_pdl_trigger();
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
      __threadfence();   // only the writers of the partial row fence (a gpu-scope fence per thread is not free)
    }
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = atomicAdd(ticket, 1u);
      is_last = (t == gridDim.x - 1);
      if (is_last) *ticket = 0;   // re-arm for the next launch
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    if (KG <= 32) {
      // one warp per value: lanes stride over the block rows (independent loads in flight), fixed shuffle tree
      // (all loads of a lane are issued before the first add: one L2 round trip instead of gridDim.x/32 dependent ones;
      //  the adds keep the ascending-row order, so the sum is bitwise what the plain loop gives)
      for (int k = wid; k < KG; k += nw) {
        double s = 0;
        for (unsigned int b0 = lane; b0 < gridDim.x; b0 += 32 * 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const unsigned int b = b0 + 32 * u; v[u] = b < gridDim.x ? __ldcg(&partials[(size_t)b * LRR_NRED + k]) : 0.0; }
#pragma unroll
          for (int u = 0; u < 8; ++u) s += v[u];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh_out[k] = s;
      }
    } else {
      // many values: lanes own values (coalesced rows), warps own interleaved rows, fixed-order combine
      constexpr int KL = (KG + 31) / 32;
      double s[KL];
#pragma unroll
      for (int j = 0; j < KL; ++j) s[j] = 0;
      // (the loads of RB rows are issued before the first add — one L2 round trip per RB rows instead of one per row, which was most
      //  of a finalizer's time at two blocks per SM; the adds keep the ascending-row order, so the sums are bitwise the plain loop's)
      constexpr int RB = KL >= 3 ? 4 : 8;
      for (unsigned int b0 = wid; b0 < gridDim.x; b0 += RB * nw) {
        double v[RB][KL];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const unsigned int b = b0 + u * nw;
#pragma unroll
          for (int j = 0; j < KL; ++j) { const int k = lane + 32 * j; v[u][j] = (b < gridDim.x && k < KG) ? __ldcg(&partials[(size_t)b * LRR_NRED + k]) : 0.0; }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          if (b0 + u * nw < gridDim.x) {
#pragma unroll
            for (int j = 0; j < KL; ++j) s[j] += v[u][j];
          }
        }
      }
      __syncthreads();   // sh_tmp is free again
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
    return true;
  }
  // ticket only (no sums): true in every thread of the last block
  __device__ __forceinline__ bool last_block() const {
    __shared__ bool is_last2;
    lrr_pdl_trigger();
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
  __device__ __forceinline__ int block_first() const { return 0; }
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
