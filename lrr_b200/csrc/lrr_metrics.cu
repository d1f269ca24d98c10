// K5: decision / quality reductions over the current matches (sm_100a):
//   sheet_angles / tube_angles (cdcl/cdcl_estimation_ICP_matching_all.txx:273-382 via covariance eigenvectors,
//   rrl/rrl_estimation_ICP_matching_all.txx:281-375 via mapped directions_), the second moments behind
//   compute_transfer_error_covariance (cdcl :388-404, rrl :381-397), residual keys for median_error
//   (cdcl :159-190) and the MUSE sort, and the match export of the display callback.
#include "lrr_phases.cuh"

static inline int nblk_for(long long items) {
  long long b = (items + LRR_THREADS - 1) / LRR_THREADS;
  long long cap = (long long)lrr_num_sms() * 4;
  if (cap > LRR_MAX_BLOCKS) cap = LRR_MAX_BLOCKS;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void k_resid_keys(DirArgs a, double* __restrict__ keys) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.M; k += gridDim.x * blockDim.x) keys[k] = sqrt(a.r2[k]);
}
void lrr_k_resid_keys(const DirArgs& a, double* keys, cudaStream_t st, long long* launches) {
  if (a.M == 0) return;
  k_resid_keys<<<nblk_for(a.M), LRR_THREADS, 0, st>>>(a, keys); ++*launches;
}

__global__ void k_export_matches(DirArgs a, int* __restrict__ from_idx, int* __restrict__ to_idx) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.M; k += gridDim.x * blockDim.x) {
    const int i = lrr_from_index(a, k);
    from_idx[k] = a.forig ? a.forig[i] : i;
    const int s = a.sidx[k];
    to_idx[k] = s >= 0 ? lrr_pos4_orig(lrr_ld_pos4(a.to.grid.pos4 + s)) : -1;
  }
}
void lrr_k_export_matches(const DirArgs& a, int* from_idx, int* to_idx, cudaStream_t st, long long* launches) {
  if (a.M == 0) return;
  k_export_matches<<<nblk_for(a.M), LRR_THREADS, 0, st>>>(a, from_idx, to_idx); ++*launches;
}

__device__ __forceinline__ double lrr_fold_angle(double dot) {
  double ang = acos(dot);
  if (ang > 3.14159265358979323846 / 2) ang = 3.14159265358979323846 - ang;
  return ang;
}
__device__ __forceinline__ void lrr_ld_sym(const double* __restrict__ C6, int n, int o, double (&C)[9]) {
  const double xx = C6[o], xy = C6[(size_t)n + o], xz = C6[(size_t)2 * n + o], yy = C6[(size_t)3 * n + o], yz = C6[(size_t)4 * n + o],
               zz = C6[(size_t)5 * n + o];
  C[0] = xx; C[1] = xy; C[2] = xz; C[3] = xy; C[4] = yy; C[5] = yz; C[6] = xz; C[7] = yz; C[8] = zz;
}

// which: 0 = sheet (shape SHEET=2 on both sides), 1 = tube (TUBE=1)
__global__ void __launch_bounds__(LRR_THREADS) k_angles(DirArgs a, int flags, int which, double* partials, unsigned int* ticket) {
  __shared__ PhaseShared sh;
  GridScope scope; scope.partials = partials; scope.ticket = ticket;
  Scal* sc = a.sc;
  const int want = which == 0 ? 2 : 1;
  const int mapped = sc->from_mapped;
  double acc[2] = {0, 0};
  const FsetDev& F = a.from_all; const FsetDev& T = a.to;
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int s = a.sidx[k];
    if (s < 0) continue;
    const int i = lrr_from_index(a, k);
    const int po = a.forig ? a.forig[i] : i;
    const int qo = lrr_pos4_orig(lrr_ld_pos4(T.grid.pos4 + s));
    if (F.shape[po] != want || T.shape[qo] != want) continue;
    double ang;
    if (!(flags & LRR_F_RRL)) {
      if (!F.C6 || !T.C6) continue;
      double Cp[9], Cq[9], Vp[9], Vq[9], D[3];
      lrr_ld_sym(F.C6, F.n, po, Cp); lrr_ld_sym(T.C6, T.n, qo, Cq);
      if (mapped) {   // the clone's covariance is Jp C Jp^T  (cdcl :465-480)
        double Tm[9], Cn[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int m = 0; m < 3; ++m) v += sc->Ag[r * 3 + m] * Cp[m * 3 + c]; Tm[r * 3 + c] = v; }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int m = 0; m < 3; ++m) v += Tm[r * 3 + m] * sc->Ag[c * 3 + m]; Cn[r * 3 + c] = v; }
        for (int m = 0; m < 9; ++m) Cp[m] = Cn[m];
        if (mapped == 2) {   // ... and once more for a clone of a clone
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int m = 0; m < 3; ++m) v += sc->Ag2[r * 3 + m] * Cp[m * 3 + c]; Tm[r * 3 + c] = v; }
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int m = 0; m < 3; ++m) v += Tm[r * 3 + m] * sc->Ag2[c * 3 + m]; Cn[r * 3 + c] = v; }
          for (int m = 0; m < 9; ++m) Cp[m] = Cn[m];
        }
      }
      lrr_sym_eig3(Cp, Vp, D); lrr_sym_eig3(Cq, Vq, D);
      const int col = which == 0 ? 0 : 2;   // smallest / largest eigenvalue
      ang = lrr_fold_angle(Vp[col] * Vq[col] + Vp[3 + col] * Vq[3 + col] + Vp[6 + col] * Vq[6 + col]);
    } else {
      if (!F.dirs || !T.dirs) continue;
      double pd[3], qd[3];
      if (which == 0) {
        for (int d = 0; d < 3; ++d) { pd[d] = F.dirs[(size_t)d * F.n + po]; qd[d] = T.dirs[(size_t)d * T.n + qo]; }
      } else {   // tangent = d0 x d1
        double p0[3], p1[3], q0[3], q1[3];
        for (int d = 0; d < 3; ++d) {
          p0[d] = F.dirs[(size_t)d * F.n + po]; p1[d] = F.dirs[(size_t)(3 + d) * F.n + po];
          q0[d] = T.dirs[(size_t)d * T.n + qo]; q1[d] = T.dirs[(size_t)(3 + d) * T.n + qo];
        }
        pd[0] = p0[1] * p1[2] - p0[2] * p1[1]; pd[1] = p0[2] * p1[0] - p0[0] * p1[2]; pd[2] = p0[0] * p1[1] - p0[1] * p1[0];
        qd[0] = q0[1] * q1[2] - q0[2] * q1[1]; qd[1] = q0[2] * q1[0] - q0[0] * q1[2]; qd[2] = q0[0] * q1[1] - q0[1] * q1[0];
      }
      // map_loc(p + dir) - map_loc(p), normalised  (rrl :296-306)
      double x, y, z; lrr_from_loc(a, sc, mapped, i, x, y, z);
      double y0, y1, y2, n0, n1, n2;
      lrr_map_exact(sc->A, sc->t, sc->c, x, y, z, y0, y1, y2);
      lrr_map_exact(sc->A, sc->t, sc->c, x + pd[0], y + pd[1], z + pd[2], n0, n1, n2);
      double d0 = n0 - y0, d1 = n1 - y1, d2 = n2 - y2;
      const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      d0 /= nrm; d1 /= nrm; d2 /= nrm;
      ang = lrr_fold_angle(d0 * qd[0] + d1 * qd[1] + d2 * qd[2]);
    }
    const double w = a.w[k];
    acc[0] += w * ang; acc[1] += w;
  }
  if (scope.reduce<2>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x == 0) { sc->ang_num = sh.out[0]; sc->ang_den = sh.out[1]; }
  }
}
void lrr_k_angles(const DirArgs& a, int flags, int which, double* partials, unsigned int* ticket, cudaStream_t st, long long* launches) {
  k_angles<<<nblk_for(a.M), LRR_THREADS, 0, st>>>(a, flags, which, partials, ticket); ++*launches;
}

// second moments sum x~ x~^T of the Jacobian locations (x - centre, 1): 10 values -> sc->covJ[0..9) scratch
// (the host assembles mean_i J_i Sigma J_i^T from them: J is affine in x~)
__global__ void __launch_bounds__(LRR_THREADS) k_from_moments(DirArgs a, double* partials, unsigned int* ticket, double* out10) {
  __shared__ PhaseShared sh;
  GridScope scope; scope.partials = partials; scope.ticket = ticket;
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  const double c0 = sc->c[0], c1 = sc->c[1], c2 = sc->c[2];
  double acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0;
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int i = lrr_from_index(a, k);
    double x, y, z; lrr_from_loc(a, sc, mapped, i, x, y, z);
    x -= c0; y -= c1; z -= c2;
    acc[0] += x * x; acc[1] += x * y; acc[2] += x * z; acc[3] += x; acc[4] += y * y; acc[5] += y * z; acc[6] += y; acc[7] += z * z;
    acc[8] += z; acc[9] += 1.0;
  }
  if (scope.reduce<10>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x < 10) out10[threadIdx.x] = sh.out[threadIdx.x];
  }
}
void lrr_k_transfer_cov(const DirArgs& a, int dof, double* partials, unsigned int* ticket, cudaStream_t st, long long* launches) {
  (void)dof;
  k_from_moments<<<nblk_for(a.M), LRR_THREADS, 0, st>>>(a, partials, ticket, a.sc->covJ); ++*launches;
}
