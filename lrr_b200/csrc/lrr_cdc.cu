// E-step of the full CDC estimator: cdcl_estimation<3,12>::compute_weights (cdcl/cdcl_estimation.txx:342-545) on sm_100a.
//   for every from-feature p:  y = T(p),  Cij_p = Jth Sigma Jth^T + Jp Cp Jp^T  (:366-378),  radius = 5 sqrt(lambda_max(Cij_p)) (:382),
//   every to-feature q with |y - q|^2 <= radius^2 (rsdl_kd_tree::points_in_radius, :388) — at most `max_matches` of them, the
//   reference takes the first 100 of a random shuffle (:396-398; Q4, injectable: here the lowest original indices) —
//   w0 = weight(e^T (Cij_p + Cq)^-1 e) (:411-418, cdcl_utils.h:209-224), kept when > 0 (:420), row sum and per-to-feature sum (:450-455);
//   then w = w0^2 / (row sum * to-feature sum) (:523-536).
// One WARP per from-feature: the transfer-error covariance comes from the ten symmetric blocks of Sigma (the Kronecker form of
// lrr_objcov.cu: 60 multiply-adds), the ball is scanned with the flattened run table of lrr_nn_ball_flat (run bounds of 128 grid rows
// in one round trip, the points in a second), candidates are compacted by ballot into the warp's scratch list, the cap is applied by a
// bisection on the original index, the surviving entries are evaluated one per lane and appended to the flat match set through one
// warp-aggregated reservation.  The flat set (from index, neighbour slot, weight per entry) is exactly what k_objcov / k_objpar stream.
#include "lrr_phases.cuh"

#define LRR_CDC_THREADS 128
#define LRR_CDC_WARPS (LRR_CDC_THREADS / 32)
#define LRR_CDC_CAND 4096          // fixed features inside one search ball (scratch list per warp)

struct CdcArgs {
  double A[9], t[3], c[3];
  const double* Sigma;             // device, 12 x 12 row-major
  double eval_mult; int max_matches;
  int M;
  const double *fx, *fy, *fz; const int* forig;
  FsetDev to, from_all;
  // flat match set
  int* efrom; int* eslot; double* ew; long long ecap; unsigned long long* cursor;
  int* rowcnt; int* rowstart; double* roww; double* tow;
  double* dsum;                    // [0] sum over all from-features of sqrt(lambda_max(Jth Sigma Jth^T))
  int* istat;                      // [0] rows, [1] min entries of a row, [2] max, [3] a ball held more than LRR_CDC_CAND features
  int2* cand; double* candw;       // per-warp scratch, LRR_CDC_CAND entries each
};

// largest eigenvalue of a symmetric 3x3 (xx,xy,xz,yy,yz,zz): trigonometric solution of the characteristic cubic
// (cdcl_largest_evalue, cdcl/cdcl_utils.h:284-292, calls vnl_symmetric_eigensystem_compute_eigenvals)
__device__ __forceinline__ double lrr_sym3_lmax(const double (&C)[6]) {
  const double p1 = C[1] * C[1] + C[2] * C[2] + C[4] * C[4];
  const double q = (C[0] + C[3] + C[5]) / 3.0;
  const double a = C[0] - q, b = C[3] - q, c = C[5] - q;
  const double p2 = a * a + b * b + c * c + 2.0 * p1;
  if (!(p2 > 0.0)) return q;
  const double p = sqrt(p2 / 6.0), ip = 1.0 / p;
  const double b00 = a * ip, b11 = b * ip, b22 = c * ip, b01 = C[1] * ip, b02 = C[2] * ip, b12 = C[4] * ip;
  double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
  r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  return q + 2.0 * p * cos(acos(r) / 3.0);
}

// smallest v with #{candidates of index <= v} >= want  (the list holds more than `want` candidates)
__device__ __forceinline__ int cdc_index_threshold(const int2* cand, int cnt, int want, int nto, int lane) {
  int lo = 0, hi = nto - 1;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    int c = 0;
    for (int k = lane; k < cnt; k += 32) c += cand[k].x <= mid ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (c >= want) hi = mid; else lo = mid + 1;
  }
  return lo;
}
// The warp's candidate list is full (a very large search ball): only the max_matches lowest original indices of the WHOLE ball can
// survive the cap, so everything above the max_matches-th smallest index seen so far is dropped (in place, order kept) and the scan
// goes on.  Exact whatever the size of the ball; returns the new length (<= max_matches).
__device__ __forceinline__ int cdc_prune(int2* cand, int cnt, int max_matches, int nto, int lane) {
  __syncwarp();
  if (cnt <= max_matches) return cnt;
  const int tau = cdc_index_threshold(cand, cnt, max_matches, nto, lane);
  int nk = 0;
  for (int k0 = 0; k0 < cnt; k0 += 32) {
    const int k = k0 + lane;
    int2 cd = make_int2(0, 0); bool keep = false;
    if (k < cnt) { cd = cand[k]; keep = cd.x <= tau; }
    const unsigned int m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) cand[nk + __popc(m & ((1u << lane) - 1u))] = cd;
    nk += __popc(m);
  }
  __syncwarp();
  return nk;
}

__global__ void __launch_bounds__(LRR_CDC_THREADS, 6) k_cdc_match(const __grid_constant__ CdcArgs a) {
  __shared__ double Bs[10][6];
  __shared__ int rowlo_s[LRR_CDC_WARPS][LRR_FLAT_ROWS + 1], rowbase_s[LRR_CDC_WARPS][LRR_FLAT_ROWS + 1];
  if (threadIdx.x < 60) {
    const int pair = threadIdx.x / 6, p = threadIdx.x % 6;
    int aa = 0, bb = 0, cnt = 0;
    for (int i = 0; i < 4; ++i) for (int j2 = i; j2 < 4; ++j2) { if (cnt == pair) { aa = i; bb = j2; } ++cnt; }
    const int d = p < 3 ? 0 : (p < 5 ? 1 : 2), e = p < 3 ? p : (p < 5 ? p - 2 : 2);
    auto idx = [](int dd, int ax) { return ax < 3 ? 3 * dd + ax : 9 + dd; };
    const double s1 = a.Sigma[idx(d, aa) * 12 + idx(e, bb)], s2 = a.Sigma[idx(e, aa) * 12 + idx(d, bb)];
    Bs[pair][p] = aa == bb ? s1 : s1 + s2;
  }
  __syncthreads();
  const unsigned int full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * LRR_CDC_WARPS + wib, nw = gridDim.x * LRR_CDC_WARPS;
  int* rowlo = rowlo_s[wib]; int* rowbase = rowbase_s[wib];
  int2* cand = a.cand + (size_t)gw * LRR_CDC_CAND; double* candw = a.candw + (size_t)gw * LRR_CDC_CAND;
  const GridDev& g = a.to.grid;
  const double slack = 1e-9 * g.h;
  double lev_sum = 0.0; int rows = 0, rmin = 0x7fffffff, rmax = 0;
  for (int i = gw; i < a.M; i += nw) {
    const int po = a.forig ? __ldg(a.forig + i) : i;
    const double x = __ldg(a.fx + i), y = __ldg(a.fy + i), z = __ldg(a.fz + i);
    double qx, qy, qz;
    lrr_map_exact(a.A, a.t, a.c, x, y, z, qx, qy, qz);
    // transfer-error covariance Jth Sigma Jth^T and Cij_p
    double Cp6[6];
    {
      const double xt[4] = {x - a.c[0], y - a.c[1], z - a.c[2], 1.0};
      double T6[6] = {0, 0, 0, 0, 0, 0};
      int pair = 0;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = aa; bb < 4; ++bb) {
          const double xx = xt[aa] * xt[bb];
#pragma unroll
          for (int p = 0; p < 6; ++p) T6[p] += xx * Bs[pair][p];
          ++pair;
        }
      if (lane == 0) lev_sum += sqrt(lrr_sym3_lmax(T6));
      double C[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) C[p] = __ldg(a.from_all.C6 + (size_t)p * a.from_all.n + po);
      double AC[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double a0 = a.A[3 * r], a1 = a.A[3 * r + 1], a2 = a.A[3 * r + 2];
        AC[3 * r] = a0 * C[0] + a1 * C[1] + a2 * C[2];
        AC[3 * r + 1] = a0 * C[1] + a1 * C[3] + a2 * C[4];
        AC[3 * r + 2] = a0 * C[2] + a1 * C[4] + a2 * C[5];
      }
      Cp6[0] = T6[0] + AC[0] * a.A[0] + AC[1] * a.A[1] + AC[2] * a.A[2];
      Cp6[1] = T6[1] + AC[0] * a.A[3] + AC[1] * a.A[4] + AC[2] * a.A[5];
      Cp6[2] = T6[2] + AC[0] * a.A[6] + AC[1] * a.A[7] + AC[2] * a.A[8];
      Cp6[3] = T6[3] + AC[3] * a.A[3] + AC[4] * a.A[4] + AC[5] * a.A[5];
      Cp6[4] = T6[4] + AC[3] * a.A[6] + AC[4] * a.A[7] + AC[5] * a.A[8];
      Cp6[5] = T6[5] + AC[6] * a.A[6] + AC[7] * a.A[7] + AC[8] * a.A[8];
    }
    const double radius = a.eval_mult * sqrt(lrr_sym3_lmax(Cp6)), r2 = radius * radius;
    // ---- flattened ball scan: every to-feature within the radius goes on the warp's list --------------------------------------
    int cnt = 0;
    if (g.n > 0 && radius == radius) {
      const double Rm = radius * (1.0 + 1e-9) + slack, Rm2 = Rm * Rm;
      const int z0 = lrr_cell_clamp(qz - Rm, g.oz, g.inv_h, g.nz), z1 = lrr_cell_clamp(qz + Rm, g.oz, g.inv_h, g.nz);
      const int y0 = lrr_cell_clamp(qy - Rm, g.oy, g.inv_h, g.ny), y1 = lrr_cell_clamp(qy + Rm, g.oy, g.inv_h, g.ny);
      const int ny = y1 - y0 + 1, nrows = (z1 - z0 + 1) * ny;
      for (int r0 = 0; r0 < nrows; r0 += LRR_FLAT_ROWS) {
        const int nb = min(LRR_FLAT_ROWS, nrows - r0);
        int lo4[LRR_FLAT_ROWS / 32], hi4[LRR_FLAT_ROWS / 32];
#pragma unroll
        for (int u = 0; u < LRR_FLAT_ROWS / 32; ++u) {
          const int rl = 32 * u + lane;
          lo4[u] = 0; hi4[u] = 0;
          if (rl < nb) {
            const int r = r0 + rl, zc = z0 + r / ny, yc = y0 + r % ny;
            const double dz = lrr_slab_dist(qz, g.oz, g.h, zc, slack), dy = lrr_slab_dist(qy, g.oy, g.h, yc, slack);
            const double dyz2 = dz * dz + dy * dy;
            if (dyz2 <= Rm2) {
              const double rx = sqrt(Rm2 - dyz2) * (1.0 + 1e-9) + slack;
              const int xa = lrr_cell_clamp(qx - rx, g.ox, g.inv_h, g.nx), xb = lrr_cell_clamp(qx + rx, g.ox, g.inv_h, g.nx);
              const int row = (zc * g.ny + yc) * g.nx;
              lo4[u] = __ldg(g.cell_start + row + xa); hi4[u] = __ldg(g.cell_start + row + xb + 1);
            }
          }
        }
        int total = 0;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < LRR_FLAT_ROWS / 32; ++u) {
          if (32 * u < nb) {
            const int rl = 32 * u + lane, sz = (rl < nb) ? hi4[u] - lo4[u] : 0;
            int inc = sz;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(full, inc, o); if (lane >= o) inc += v; }
            if (rl < nb) { rowlo[rl] = lo4[u]; rowbase[rl] = total + inc - sz; }
            total += __shfl_sync(full, inc, 31);
          }
        }
        if (lane == 0) rowbase[nb] = total;
        __syncwarp();
        for (int p0 = 0; p0 < total; p0 += 32) {
          const int p = p0 + lane;
          bool in = false; int J = -1, o = 0;
          if (p < total) {
            int ra = 0, rb = nb;
            while (rb - ra > 1) { const int mid = (ra + rb) >> 1; if (rowbase[mid] <= p) ra = mid; else rb = mid; }
            J = rowlo[ra] + (p - rowbase[ra]);
            const double4 P = lrr_ld_pos4(g.pos4 + J);
            in = lrr_dist2_exact(qx, qy, qz, P.x, P.y, P.z) <= r2;
            o = lrr_pos4_orig(P);
          }
          const unsigned int m = __ballot_sync(full, in);
          if (m && cnt + 32 > LRR_CDC_CAND) cnt = cdc_prune(cand, cnt, a.max_matches, a.to.n, lane);   // (warp-uniform)
          if (in) { const int pos = cnt + __popc(m & ((1u << lane) - 1u)); if (pos < LRR_CDC_CAND) cand[pos] = make_int2(o, J); }
          cnt += __popc(m);
        }
      }
    }
    __syncwarp();
    int nk = 0; double sum_w = 0.0;
    if (cnt > LRR_CDC_CAND) { if (lane == 0) atomicExch(a.istat + 3, 1); cnt = 0; }   // (only when max_matches itself does not fit the list)
    if (cnt > 0) {
      // cap: the max_matches lowest original indices (bisection on the index: smallest v with #{o <= v} >= max_matches)
      int tau = 0x7fffffff;
      if (cnt > a.max_matches) tau = cdc_index_threshold(cand, cnt, a.max_matches, a.to.n, lane);
      // evaluate the candidates, one per lane; kept entries are compacted in place (the write position never passes the chunk read)
      for (int k0 = 0; k0 < cnt; k0 += 32) {
        const int k = k0 + lane;
        double w0 = 0.0; int slot = -1;
        if (k < cnt) {
          const int2 cd = cand[k];
          if (cd.x <= tau) {
            slot = cd.y;
            const double4 P = lrr_ld_pos4(g.pos4 + slot);
            const double2* cr = reinterpret_cast<const double2*>(a.to.crec + 8 * (size_t)slot);
            const double2 c0 = __ldg(cr), c1 = __ldg(cr + 1), c2 = __ldg(cr + 2);
            const double Cij[6] = {Cp6[0] + c0.x, Cp6[1] + c0.y, Cp6[2] + c1.x, Cp6[3] + c1.y, Cp6[4] + c2.x, Cp6[5] + c2.y};
            const double c00 = Cij[3] * Cij[5] - Cij[4] * Cij[4], c01 = Cij[2] * Cij[4] - Cij[1] * Cij[5], c02 = Cij[1] * Cij[4] - Cij[2] * Cij[3];
            const double det = Cij[0] * c00 + Cij[1] * c01 + Cij[2] * c02, id = 1.0 / det;
            const double Ci[6] = {c00 * id, c01 * id, c02 * id, (Cij[0] * Cij[5] - Cij[2] * Cij[2]) * id, (Cij[1] * Cij[2] - Cij[0] * Cij[4]) * id,
                                  (Cij[0] * Cij[3] - Cij[1] * Cij[1]) * id};
            const double e0 = qx - P.x, e1 = qy - P.y, e2 = qz - P.z;
            const double v0 = Ci[0] * e0 + Ci[1] * e1 + Ci[2] * e2, v1 = Ci[1] * e0 + Ci[3] * e1 + Ci[4] * e2, v2 = Ci[2] * e0 + Ci[4] * e1 + Ci[5] * e2;
            const double res = e0 * v0 + e1 * v1 + e2 * v2;
            // weight(): Beaton-Tukey, b = 4, times f = 2.9872 (cdcl/cdcl_utils.h:209-224)
            if (res <= 16.0) { const double tt = 1.0 - res / 16.0; w0 = 2.9872 * tt * tt; }
          }
        }
        const bool keep = w0 > 0.0;
        const unsigned int m = __ballot_sync(full, keep);
        __syncwarp();
        if (keep) { const int pos = nk + __popc(m & ((1u << lane) - 1u)); cand[pos].y = slot; candw[pos] = w0; }
        nk += __popc(m);
        sum_w += w0;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum_w += __shfl_xor_sync(full, sum_w, o);
      __syncwarp();
    }
    int base = 0;
    if (nk > 0) {
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(a.cursor, (unsigned long long)nk);
      b = __shfl_sync(full, b, 0);
      if (b + (unsigned long long)nk <= (unsigned long long)a.ecap) {
        base = (int)b;
        for (int k = lane; k < nk; k += 32) {
          const int slot = cand[k].y; const double w0 = candw[k];
          a.efrom[base + k] = i; a.eslot[base + k] = slot; a.ew[base + k] = w0;
          atomicAdd(a.tow + slot, w0);
        }
      } else base = -1;   // the set did not fit: the host reruns with the capacity the cursor asks for
      ++rows; rmin = min(rmin, nk); rmax = max(rmax, nk);
    }
    if (lane == 0) { a.rowcnt[i] = nk; a.rowstart[i] = base; a.roww[i] = sum_w; }
    __syncwarp();
  }
  if (lane == 0) {
    atomicAdd(a.dsum, lev_sum);
    if (rows) { atomicAdd(a.istat, rows); atomicMin(a.istat + 1, rmin); atomicMax(a.istat + 2, rmax); }
  }
}

// w = w0^2 / (row sum * to-feature sum)   (cdcl/cdcl_estimation.txx:523-536)
__global__ void k_cdc_normalise(int n, const int* __restrict__ efrom, const int* __restrict__ eslot, double* __restrict__ ew,
                                const double* __restrict__ roww, const double* __restrict__ tow) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double w0 = ew[k];
    ew[k] = w0 * w0 / (roww[efrom[k]] * tow[eslot[k]]);
  }
}

// cdcl_normalize_matches over the many-to-many set (cdcl/cdcl_estimation_ICP.txx:751-832): centre and mean radius of the from-features
// that own a row and of the to-features that appear in any row (each once).  pass 0: sums of locations and counts; pass 1: radii.
// out[0..2] sum from, [3] rows, [4..6] sum to, [7] unique to  |  [8] sum |from - cm|, [9] sum |to - cf|
__global__ void k_cdc_norm(int pass, int M, int nto, const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
                           const int* __restrict__ rowcnt, const double* __restrict__ tow, const double4* __restrict__ pos4, double* out) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double cm[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
  if (pass == 1) {
    for (int d = 0; d < 3; ++d) { cm[d] = out[d] / out[3]; cf[d] = out[4 + d] / out[7]; }
  }
  const int n = M > nto ? M : nto;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (k < M && rowcnt[k] > 0) {
      const double x = fx[k], y = fy[k], z = fz[k];
      if (pass == 0) { acc[0] += x; acc[1] += y; acc[2] += z; acc[3] += 1.0; }
      else { const double dx = x - cm[0], dy = y - cm[1], dz = z - cm[2]; acc[0] += sqrt(dx * dx + dy * dy + dz * dz); }
    }
    if (k < nto && tow[k] > 0.0) {
      const double4 P = lrr_ld_pos4(pos4 + k);
      if (pass == 0) { acc[4] += P.x; acc[5] += P.y; acc[6] += P.z; acc[7] += 1.0; }
      else { const double dx = P.x - cf[0], dy = P.y - cf[1], dz = P.z - cf[2]; acc[1] += sqrt(dx * dx + dy * dy + dz * dz); }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if ((threadIdx.x & 31) == 0) {
    if (pass == 0) { for (int i = 0; i < 8; ++i) if (acc[i] != 0.0) atomicAdd(out + i, acc[i]); }
    else { atomicAdd(out + 8, acc[0]); atomicAdd(out + 9, acc[1]); }
  }
}

size_t lrr_cdc_scratch_entries(int blocks) { return (size_t)blocks * LRR_CDC_WARPS * LRR_CDC_CAND; }
int lrr_cdc_blocks(int M) {
  // 80 registers per thread: six blocks of four warps per SM — the kernel is bound by the latency of its dependent phases (run bounds ->
  // points -> covariance records), not by bandwidth: the fixed set is L2-resident (0.96 -> see profiles/r2_summary.md)
  // (the grid is what is resident at once: a seventh block per SM would run as a second wave behind the other six — measured 466 -> 695 us
  //  when the kernel had grown to 96 registers and only five blocks fitted)
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cdc_match, LRR_CDC_THREADS, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
  if (per_sm > 6) per_sm = 6;
  long long b = ((long long)M + LRR_CDC_WARPS - 1) / LRR_CDC_WARPS, cap = (long long)lrr_num_sms() * per_sm;
  if (b > cap) b = cap; if (b < 1) b = 1;
  return (int)b;
}
void lrr_k_cdc_match(const CdcLaunch& L, cudaStream_t st, long long* launches) {
  CdcArgs a;
  for (int i = 0; i < 9; ++i) a.A[i] = L.A[i];
  for (int i = 0; i < 3; ++i) { a.t[i] = L.t[i]; a.c[i] = L.c[i]; }
  a.Sigma = L.Sigma_dev; a.eval_mult = L.eval_mult; a.max_matches = L.max_matches; a.M = L.M;
  a.fx = L.fx; a.fy = L.fy; a.fz = L.fz; a.forig = L.forig; a.to = L.to; a.from_all = L.from_all;
  a.efrom = L.efrom; a.eslot = L.eslot; a.ew = L.ew; a.ecap = L.ecap; a.cursor = L.cursor;
  a.rowcnt = L.rowcnt; a.rowstart = L.rowstart; a.roww = L.roww; a.tow = L.tow; a.dsum = L.dsum; a.istat = L.istat;
  a.cand = L.cand; a.candw = L.candw;
  k_cdc_match<<<L.blocks, LRR_CDC_THREADS, 0, st>>>(a); ++*launches;
}
void lrr_k_cdc_normalise(int n, const int* efrom, const int* eslot, double* ew, const double* roww, const double* tow, cudaStream_t st, long long* launches) {
  if (n <= 0) return;
  int b = (n + 255) / 256, cap = lrr_num_sms() * 8; if (b > cap) b = cap;
  k_cdc_normalise<<<b, 256, 0, st>>>(n, efrom, eslot, ew, roww, tow); ++*launches;
}
void lrr_k_cdc_norm(int M, int nto, const double* fx, const double* fy, const double* fz, const int* rowcnt, const double* tow, const double4* pos4,
                    double* out10, cudaStream_t st, long long* launches) {
  const int n = M > nto ? M : nto;
  int b = (n + 255) / 256, cap = lrr_num_sms() * 4; if (b > cap) b = cap; if (b < 1) b = 1;
  cudaMemsetAsync(out10, 0, 10 * sizeof(double), st);
  k_cdc_norm<<<b, 256, 0, st>>>(0, M, nto, fx, fy, fz, rowcnt, tow, pos4, out10);
  k_cdc_norm<<<b, 256, 0, st>>>(1, M, nto, fx, fy, fz, rowcnt, tow, pos4, out10);
  *launches += 2;
}
