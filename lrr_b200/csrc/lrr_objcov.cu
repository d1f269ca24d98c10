// K6: cdcl_obj_fun_cov<3,12>::compute (cdcl/cdcl_obj_fun_cov.txx:44-134) over the current matches (sm_100a).
//   Cij = Jp Cp Jp^T + Jt Sigma Jt^T + Cq ;  r = e^T Cij^-1 e ;  f += w (rho(r) + ln det Cij)
//   kl += w ( Jt^T Cij^-1 Jt - rho'(r) t t^T ),  t = Jt^T Cij^-1 e
// The affine Jacobian is J = x~^T (x) I3 (x~ = (x - c, 1)), so with v = Cij^-1 e and the symmetric 3x3
// S = Cij^-1 - rho'(r) v v^T the gradient kernel is the Kronecker moment  kl[(d,a),(e,b)] = sum w x~_a x~_b S_de
// — the same 4-lanes-per-match accumulation as the normal equations (24 + 1 accumulators per lane) instead of
// the reference's dense 12x3 . 3x3 . 3x12 products per match; J Sigma J^T is 144 MACs split over the 4 lanes.
// This keeps the evaluation HBM-bound: ~156 B and ~0.6 kflop per match.
#include "lrr_phases.cuh"

struct ObjArgs {
  double A[9], t[3], c[3];          // transform in the frame the objective is evaluated in
  double cm[3], rm, cf[3], rf;      // data normalisation (identity: 0,1,0,1)
  const double* Sigma;              // device, 12 x 12 row-major = U^T U
};

__device__ __forceinline__ int obj_idx(int d, int a) { return a < 3 ? 3 * d + a : 9 + d; }

__global__ void __launch_bounds__(LRR_THREADS) k_objcov(DirArgs a, ObjArgs o, double* partials, unsigned int* ticket) {
  __shared__ PhaseShared sh;
  __shared__ double Sg[144];
  GridScope scope; scope.partials = partials; scope.ticket = ticket;
  for (int i = threadIdx.x; i < 144; i += blockDim.x) Sg[i] = o.Sigma[i];
  __syncthreads();
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  const int j = threadIdx.x & 3;
  const FsetDev& F = a.from_all; const FsetDev& T = a.to;
  const double irm2 = 1.0 / (o.rm * o.rm), irf2 = 1.0 / (o.rf * o.rf);
  double acc[25];
#pragma unroll
  for (int i = 0; i < 25; ++i) acc[i] = 0.0;
  for (int k = scope.first() >> 2; k < a.M; k += scope.stride() >> 2) {
    const int s = a.sidx[k];
    if (s < 0) continue;
    const int i = lrr_from_index(a, k);
    const int po = a.forig ? a.forig[i] : i;
    double px, py, pz; lrr_from_loc(a, sc, mapped, i, px, py, pz);
    px = (px - o.cm[0]) / o.rm; py = (py - o.cm[1]) / o.rm; pz = (pz - o.cm[2]) / o.rm;
    const double4 q4 = lrr_ld_pos4(T.grid.pos4 + s);
    const int qo = lrr_pos4_orig(q4);
    const double qx = (q4.x - o.cf[0]) / o.rf, qy = (q4.y - o.cf[1]) / o.rf, qz = (q4.z - o.cf[2]) / o.rf;
    // covariances (symmetric 6: xx,xy,xz,yy,yz,zz)
    double Cp[6], Cq[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) { Cp[p] = F.C6[(size_t)p * F.n + po] * irm2; Cq[p] = T.C6[(size_t)p * T.n + qo] * irf2; }
    const double xt[4] = {px - o.c[0], py - o.c[1], pz - o.c[2], 1.0};
    // JCpJ = A Cp A^T
    double M1[9];
    {
      const double C[9] = {Cp[0], Cp[1], Cp[2], Cp[1], Cp[3], Cp[4], Cp[2], Cp[4], Cp[5]};
      double AC[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) AC[r * 3 + cc] = o.A[r * 3] * C[cc] + o.A[r * 3 + 1] * C[3 + cc] + o.A[r * 3 + 2] * C[6 + cc];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) M1[r * 3 + cc] = AC[r * 3] * o.A[cc * 3] + AC[r * 3 + 1] * o.A[cc * 3 + 1] + AC[r * 3 + 2] * o.A[cc * 3 + 2];
    }
    // J Sigma J^T: lane j sums the terms with a = j, then the 4 lanes add up
    double JS[9];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        double v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v += xt[b] * Sg[obj_idx(d, j) * 12 + obj_idx(e, b)];
        JS[d * 3 + e] = xt[j] * v;
      }
#pragma unroll
    for (int m = 0; m < 9; ++m) { JS[m] += __shfl_xor_sync(0xffffffffu, JS[m], 1); JS[m] += __shfl_xor_sync(0xffffffffu, JS[m], 2); }
    double Cij[9];
    {
      const double Q[9] = {Cq[0], Cq[1], Cq[2], Cq[1], Cq[3], Cq[4], Cq[2], Cq[4], Cq[5]};
#pragma unroll
      for (int m = 0; m < 9; ++m) Cij[m] = M1[m] + JS[m] + Q[m];
    }
    double Ci[9]; lrr_inv3(Cij, Ci);
    const double det = lrr_det3(Cij);
    // err = q - T(p)   (cdcl_obj_fun_cov.txx:84-86)
    const double yx = o.A[0] * xt[0] + o.A[1] * xt[1] + o.A[2] * xt[2] + o.t[0];
    const double yy = o.A[3] * xt[0] + o.A[4] * xt[1] + o.A[5] * xt[2] + o.t[1];
    const double yz = o.A[6] * xt[0] + o.A[7] * xt[1] + o.A[8] * xt[2] + o.t[2];
    const double e0 = qx - yx, e1 = qy - yy, e2 = qz - yz;
    const double v0 = Ci[0] * e0 + Ci[1] * e1 + Ci[2] * e2, v1 = Ci[3] * e0 + Ci[4] * e1 + Ci[5] * e2, v2 = Ci[6] * e0 + Ci[7] * e1 + Ci[8] * e2;
    const double r = e0 * v0 + e1 * v1 + e2 * v2;
    const double w = a.w[k];
    if (j == 0) acc[24] += w * (lrr_rho(r) + log(det));
    const double rp = lrr_rho_p(r);
    // S = Ci - rho' v v^T, symmetric part taken from the upper triangle (Ci is symmetric up to rounding)
    const double S[6] = {Ci[0] - rp * v0 * v0, Ci[1] - rp * v0 * v1, Ci[2] - rp * v0 * v2, Ci[4] - rp * v1 * v1, Ci[5] - rp * v1 * v2, Ci[8] - rp * v2 * v2};
    const double wxa = w * xt[j];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const double tt = wxa * xt[b];
#pragma unroll
      for (int p = 0; p < 6; ++p) acc[b * 6 + p] += tt * S[p];
    }
  }
  if (scope.reduce<25, 4>(acc, sh.out, sh.tmp)) {
    for (int i = threadIdx.x; i < 100; i += blockDim.x) sc->mom[i] = sh.out[i];
  }
}

void lrr_k_objcov_launch(const DirArgs& a, const double* A, const double* t, const double* c, const double* norm8, const double* Sigma_dev,
                         double* partials, unsigned int* ticket, cudaStream_t st, long long* launches) {
  ObjArgs o;
  for (int i = 0; i < 9; ++i) o.A[i] = A[i];
  for (int i = 0; i < 3; ++i) { o.t[i] = t[i]; o.c[i] = c[i]; o.cm[i] = norm8[i]; o.cf[i] = norm8[4 + i]; }
  o.rm = norm8[3]; o.rf = norm8[7]; o.Sigma = Sigma_dev;
  long long b = (4LL * a.M + LRR_THREADS - 1) / LRR_THREADS, cap = (long long)lrr_num_sms() * 4;
  if (cap > LRR_MAX_BLOCKS) cap = LRR_MAX_BLOCKS;
  if (b > cap) b = cap; if (b < 1) b = 1;
  k_objcov<<<(int)b, LRR_THREADS, 0, st>>>(a, o, partials, ticket); ++*launches;
}
