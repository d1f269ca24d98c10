// K6: cdcl_obj_fun_cov<3,12>::compute (cdcl/cdcl_obj_fun_cov.txx:44-134) over the current matches (sm_100a).
//   Cij = Jp Cp Jp^T + Jt Sigma Jt^T + Cq ;  r = e^T Cij^-1 e ;  f += w (rho(r) + ln det Cij)
//   kl += w ( Jt^T Cij^-1 Jt - rho'(r) t t^T ),  t = Jt^T Cij^-1 e
// The affine Jacobian is J = x~^T (x) I3 (x~ = (x - c, 1)), so with v = Cij^-1 e and the symmetric 3x3
// S = Cij^-1 - rho'(r) v v^T the gradient kernel is the Kronecker moment  kl[(d,a),(e,b)] = sum w x~_a x~_b S_de,
// and J Sigma J^T = sum_{a<=b} x~_a x~_b B_ab with the ten symmetric 3x3 blocks B_aa = Sigma_aa, B_ab = Sigma_ab + Sigma_ab^T
// (60 multiply-adds per match from 60 broadcast shared-memory values) instead of the reference's dense 3x12 . 12x12 . 12x3.
// ~0.3 kflop and 156 B per match: HBM-bound — provided the loads are: every LANE takes one match per trip
// (coalesced streams for the from-side: location, covariance, weight, neighbour slot; two gathers of one sector pair each for the
// neighbour: pos4 and the cell-ordered covariance record FsetDev::crec), the four lanes of a quad then pass (w, x~, S) round by
// shuffle and each adds row x~_j of all four matches: 24 + 1 accumulators per lane.
// Loop shape (LRR_OBJ_PIPE = 1, measured on B200 at 1 M matches: 77 -> 66 us, scripts/ab_objcov.sh): a trip issues its loads, then runs the
// quad accumulation of the PREVIOUS trip behind them, then its own per-match chain with a branch-free reciprocal and logarithm (1.0 / det and
// log() each end in a slow-path branch); one 256-thread block per SM at ~200 registers beats two at 128 (fewer rows for the finalizer, no
// register-starved schedule); keeping the NEXT trip's loads in flight as well (LRR_OBJ_PIPE = 2) was measured slower (73 us).
#include "lrr_phases.cuh"
#include <math_constants.h>

struct ObjArgs {
  double A[9], t[3], c[3];          // transform in the frame the objective is evaluated in
  double cm[3], rm, cf[3], rf;      // data normalisation (identity: 0,1,0,1)
  const double* Sigma;              // device, 12 x 12 row-major = U^T U
};

__device__ __forceinline__ int obj_idx(int d, int a) { return a < 3 ? 3 * d + a : 9 + d; }

#ifndef LRR_OBJ_THREADS
#define LRR_OBJ_THREADS 256
#endif
#ifndef LRR_OBJ_PIPE
#define LRR_OBJ_PIPE 1   // 0: plain loop, 1: accumulation of the previous trip behind this trip's loads, 2: + next trip's loads in flight (measured slower)
#endif
#ifndef LRR_OBJ_MINB
#define LRR_OBJ_MINB 1   // resident blocks per SM the register budget is set for
#endif
typedef PhaseSharedT<(LRR_OBJ_THREADS / 32 > LRR_NWARPS ? LRR_OBJ_THREADS / 32 : LRR_NWARPS)> ObjShared;   // a row of sums per warp
#if LRR_OBJ_PIPE
// Branch-free reciprocal and logarithm for the pipelined loop (positive normal arguments: determinants of positive definite 3 x 3 sums).
// 1.0 / d and log() compile to a fast path plus a slow-path branch each, which ends the basic block the scheduler can fill.
__device__ __forceinline__ double obj_rcp(double d) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0); r = fma(r, e, r);
  e = fma(-d, r, 1.0); r = fma(r, e, r);
  e = fma(-d, r, 1.0); r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double obj_log(double x) {
  // argument reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)); log m = f - (f^2/2 - s (f^2/2 + R(s^2))), f = m - 1, s = f / (2 + f)
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01, L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
               L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01, L7 = 1.479819860511658591e-01;
  int hx = __double2hiint(x); const int lx = __double2loint(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;
  hx |= (i ^ 0x3ff00000); k += (i >> 20);
  const double f = __hiloint2double(hx, lx) - 1.0;
  const double s = f * obj_rcp(2.0 + f);
  const double dk = (double)k, z = s * s, w = z * z;
  const double t1 = w * (L2 + w * (L4 + w * L6));
  const double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
  const double R = t2 + t1, hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}
__device__ __forceinline__ double obj_sel4(int k0, int k1, int k2, double a, double b, double c) {
  // a, b, c or 1.0 by all-ones masks (none set: 1.0) — logic operations, no branch and nothing on the FP64 pipe
  const int k3 = ~(k0 | k1 | k2);
  const int hi = (__double2hiint(a) & k0) | (__double2hiint(b) & k1) | (__double2hiint(c) & k2) | (0x3ff00000 & k3);
  const int lo = (__double2loint(a) & k0) | (__double2loint(b) & k1) | (__double2loint(c) & k2);
  return __hiloint2double(hi, lo);
}
#endif
__global__ void __launch_bounds__(LRR_OBJ_THREADS, LRR_OBJ_MINB) k_objcov(const __grid_constant__ DirArgs a, const __grid_constant__ ObjArgs o, double* partials, unsigned int* ticket) {
  __shared__ ObjShared sh;
  __shared__ double Bs[10][6];     // blocks B_ab (a <= b), symmetric 6: xx,xy,xz,yy,yz,zz
  GridScope scope; scope.partials = partials; scope.ticket = ticket;
  if (threadIdx.x < 60) {
    const int pair = threadIdx.x / 6, p = threadIdx.x % 6;
    // pair -> (a, b), a <= b, in the order (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3)
    int aa = 0, bb = 0, cnt = 0;
    for (int i = 0; i < 4; ++i) for (int j2 = i; j2 < 4; ++j2) { if (cnt == pair) { aa = i; bb = j2; } ++cnt; }
    const int d = p < 3 ? 0 : (p < 5 ? 1 : 2), e = p < 3 ? p : (p < 5 ? p - 2 : 2);
    const double s1 = o.Sigma[obj_idx(d, aa) * 12 + obj_idx(e, bb)];
    const double s2 = o.Sigma[obj_idx(e, aa) * 12 + obj_idx(d, bb)];   // (Sigma_ab^T)_de = (Sigma_ab)_ed
    Bs[pair][p] = aa == bb ? s1 : s1 + s2;
  }
  __syncthreads();
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  const int lane = threadIdx.x & 31, j = lane & 3, qbase = lane & ~3;
  const FsetDev& F = a.from_all; const FsetDev& T = a.to;
  const double irm = 1.0 / o.rm, irf = 1.0 / o.rf, irm2 = irm * irm, irf2 = irf * irf;
  double acc[25];
#pragma unroll
  for (int i = 0; i < 25; ++i) acc[i] = 0.0;
  const int wfirst = scope.first() - lane, wstride = scope.stride();
  // (the neighbour slot is fetched a trip ahead: the two gathers behind it start with the trip instead of a round trip into it)
#if LRR_OBJ_PIPE
  int s_next = (wfirst + lane) < a.M ? a.sidx[wfirst + lane] : -1;
  // (software pipeline: the quad accumulation of trip t-1 has no dependence on the per-match chain of trip t — Cij, inverse, log, rho —
  //  and sits in the same basic block, so the scheduler fills the chain's FP64 latency with it; lanes without a match run the chain on
  //  a clamped index with weight 0 and + I on their Cij, so that it stays finite)
  double pw = 0.0, p0 = 0.0, p1 = 0.0, p2 = 0.0, pS[6] = {0, 0, 0, 0, 0, 0};
  bool det_bad = false;
  const int k0 = -(j == 0), k1 = -(j == 1), k2 = -(j == 2);
  auto accumulate = [&]() {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
    const int src = qbase + m;
    const double wm = __shfl_sync(0xffffffffu, pw, src);
    const double xm[4] = {__shfl_sync(0xffffffffu, p0, src), __shfl_sync(0xffffffffu, p1, src), __shfl_sync(0xffffffffu, p2, src), 1.0};
    double Sm[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) Sm[p] = __shfl_sync(0xffffffffu, pS[p], src);
    const double wxa = wm * obj_sel4(k0, k1, k2, xm[0], xm[1], xm[2]);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const double tt = wxa * xm[b];
#pragma unroll
      for (int p = 0; p < 6; ++p) acc[b * 6 + p] += tt * Sm[p];
    }
  }
  };
  // raw loads of one trip (the only part with branches: index policy, optional original-index table)
  struct ObjRaw { double f[3], q[3], cq[6], cp[6], w; int valid; };
  auto fetch = [&](int k, int s_raw, ObjRaw& R) {
    R.valid = s_raw >= 0;
    const int s = R.valid ? s_raw : 0;
    const int i = lrr_from_index(a, R.valid ? k : 0);
    const int po = a.forig ? a.forig[i] : i;
    R.f[0] = __ldg(a.fx + i); R.f[1] = __ldg(a.fy + i); R.f[2] = __ldg(a.fz + i);
    const double4 q4 = lrr_ld_pos4(T.grid.pos4 + s);
    R.q[0] = q4.x; R.q[1] = q4.y; R.q[2] = q4.z;
    const double2* cr = reinterpret_cast<const double2*>(T.crec + 8 * (size_t)s);
    const double2 c0 = __ldg(cr), c1 = __ldg(cr + 1), c2 = __ldg(cr + 2);
    R.cq[0] = c0.x; R.cq[1] = c0.y; R.cq[2] = c1.x; R.cq[3] = c1.y; R.cq[4] = c2.x; R.cq[5] = c2.y;
#pragma unroll
    for (int p = 0; p < 6; ++p) R.cp[p] = __ldg(F.C6 + (size_t)p * F.n + po);
    R.w = R.valid ? a.w[k] : 0.0;
  };
#if LRR_OBJ_PIPE == 2
  // (the loads of trip t+1 are in flight during the chain of trip t: the slot index runs two trips ahead)
  int s_next2 = (wfirst + lane + wstride) < a.M ? a.sidx[wfirst + lane + wstride] : -1;
  ObjRaw R;
  if (wfirst < a.M) fetch(wfirst + lane, s_next, R);
  s_next = s_next2;
#endif
  for (int base = wfirst; base < a.M; base += wstride) {
    const int k = base + lane;
    double w, x0, x1, x2, S[6], fterm;
#if LRR_OBJ_PIPE == 2
    const ObjRaw C = R;
    s_next2 = (k + 2 * wstride) < a.M ? a.sidx[k + 2 * wstride] : -1;
    if (base + wstride < a.M) fetch(k + wstride, s_next, R);
    s_next = s_next2;
#else
    ObjRaw C;
    { const int s_raw = s_next; s_next = (k + wstride) < a.M ? a.sidx[k + wstride] : -1; fetch(k, s_raw, C); }
#endif
    double px = C.f[0], py = C.f[1], pz = C.f[2];   // lrr_from_loc on the values already loaded
    if (mapped) { double ox, oy, oz; lrr_map_exact(sc->Ag, sc->tg, sc->cg, px, py, pz, ox, oy, oz); px = ox; py = oy; pz = oz; }
    if (mapped == 2) { double ox, oy, oz; lrr_map_exact(sc->Ag2, sc->tg2, sc->cg2, px, py, pz, ox, oy, oz); px = ox; py = oy; pz = oz; }
    accumulate();   // trip t-1, here: after the loads are issued, in one basic block with the chain below
    {
      const bool valid = C.valid != 0;
      w = C.w;
      double Cq[6], Cp[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) { Cq[p] = C.cq[p] * irf2; Cp[p] = C.cp[p] * irm2; }
      const double4 q4 = make_double4(C.q[0], C.q[1], C.q[2], 0.0);
      px = (px - o.cm[0]) * irm; py = (py - o.cm[1]) * irm; pz = (pz - o.cm[2]) * irm;
      const double qx = (q4.x - o.cf[0]) * irf, qy = (q4.y - o.cf[1]) * irf, qz = (q4.z - o.cf[2]) * irf;
      x0 = px - o.c[0]; x1 = py - o.c[1]; x2 = pz - o.c[2];
      // Cij (symmetric 6) = A Cp A^T + J Sigma J^T + Cq
      double Cij[6];
      {
        // rows of A Cp
        double AC[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double a0 = o.A[3 * r], a1 = o.A[3 * r + 1], a2 = o.A[3 * r + 2];
          AC[3 * r] = a0 * Cp[0] + a1 * Cp[1] + a2 * Cp[2];
          AC[3 * r + 1] = a0 * Cp[1] + a1 * Cp[3] + a2 * Cp[4];
          AC[3 * r + 2] = a0 * Cp[2] + a1 * Cp[4] + a2 * Cp[5];
        }
        const double pad = __hiloint2double(valid ? 0 : 0x3ff00000, 0);   // + I for a lane without a match: its (weight 0) chain stays finite
        Cij[0] = AC[0] * o.A[0] + AC[1] * o.A[1] + AC[2] * o.A[2] + Cq[0] + pad;
        Cij[1] = AC[0] * o.A[3] + AC[1] * o.A[4] + AC[2] * o.A[5] + Cq[1];
        Cij[2] = AC[0] * o.A[6] + AC[1] * o.A[7] + AC[2] * o.A[8] + Cq[2];
        Cij[3] = AC[3] * o.A[3] + AC[4] * o.A[4] + AC[5] * o.A[5] + Cq[3] + pad;
        Cij[4] = AC[3] * o.A[6] + AC[4] * o.A[7] + AC[5] * o.A[8] + Cq[4];
        Cij[5] = AC[6] * o.A[6] + AC[7] * o.A[7] + AC[8] * o.A[8] + Cq[5] + pad;
        const double xt[4] = {x0, x1, x2, 1.0};
        int pair = 0;
#pragma unroll
        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
          for (int bb = aa; bb < 4; ++bb) {
            const double xx = xt[aa] * xt[bb];
#pragma unroll
            for (int p = 0; p < 6; ++p) Cij[p] += xx * Bs[pair][p];
            ++pair;
          }
      }
      // inverse and determinant of the symmetric 3x3
      const double c00 = Cij[3] * Cij[5] - Cij[4] * Cij[4], c01 = Cij[2] * Cij[4] - Cij[1] * Cij[5], c02 = Cij[1] * Cij[4] - Cij[2] * Cij[3];
      const double det = Cij[0] * c00 + Cij[1] * c01 + Cij[2] * c02;
      det_bad |= !(det > 0.0 && det < CUDART_INF);   // indefinite / singular Cij: f is poisoned after the loop, as vcl_log would (no select in the chain)
      const double id = obj_rcp(det);
      const double Ci[6] = {c00 * id, c01 * id, c02 * id, (Cij[0] * Cij[5] - Cij[2] * Cij[2]) * id, (Cij[1] * Cij[2] - Cij[0] * Cij[4]) * id,
                            (Cij[0] * Cij[3] - Cij[1] * Cij[1]) * id};
      // err = q - T(p)   (cdcl_obj_fun_cov.txx:84-86)
      const double yx = o.A[0] * x0 + o.A[1] * x1 + o.A[2] * x2 + o.t[0];
      const double yy = o.A[3] * x0 + o.A[4] * x1 + o.A[5] * x2 + o.t[1];
      const double yz = o.A[6] * x0 + o.A[7] * x1 + o.A[8] * x2 + o.t[2];
      const double e0 = qx - yx, e1 = qy - yy, e2 = qz - yz;
      const double v0 = Ci[0] * e0 + Ci[1] * e1 + Ci[2] * e2, v1 = Ci[1] * e0 + Ci[3] * e1 + Ci[4] * e2, v2 = Ci[2] * e0 + Ci[4] * e1 + Ci[5] * e2;
      const double r = e0 * v0 + e1 * v1 + e2 * v2;
      fterm = w * (lrr_rho(r) + obj_log(det));
      const double rp = lrr_rho_p(r);
      S[0] = Ci[0] - rp * v0 * v0; S[1] = Ci[1] - rp * v0 * v1; S[2] = Ci[2] - rp * v0 * v2;
      S[3] = Ci[3] - rp * v1 * v1; S[4] = Ci[4] - rp * v1 * v2; S[5] = Ci[5] - rp * v2 * v2;
    }
    acc[24] += fterm;
    pw = w; p0 = x0; p1 = x1; p2 = x2;
#pragma unroll
    for (int p = 0; p < 6; ++p) pS[p] = S[p];
  }
  accumulate();
  if (det_bad) acc[24] = CUDART_NAN;
#else
  int s_next = (wfirst + lane) < a.M ? a.sidx[wfirst + lane] : -1;
  for (int base = wfirst; base < a.M; base += wstride) {
    const int k = base + lane;
    double w = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0, S[6] = {0, 0, 0, 0, 0, 0}, fterm = 0.0;
    const int s = s_next;
    s_next = (k + wstride) < a.M ? a.sidx[k + wstride] : -1;
    if (s >= 0) {
      const int i = lrr_from_index(a, k);
      const int po = a.forig ? a.forig[i] : i;
      double px, py, pz; lrr_from_loc(a, sc, mapped, i, px, py, pz);
      const double4 q4 = lrr_ld_pos4(T.grid.pos4 + s);
      double Cq[6];
      {
        const double2* cr = reinterpret_cast<const double2*>(T.crec + 8 * (size_t)s);
        const double2 c0 = __ldg(cr), c1 = __ldg(cr + 1), c2 = __ldg(cr + 2);
        Cq[0] = c0.x * irf2; Cq[1] = c0.y * irf2; Cq[2] = c1.x * irf2; Cq[3] = c1.y * irf2; Cq[4] = c2.x * irf2; Cq[5] = c2.y * irf2;
      }
      double Cp[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) Cp[p] = __ldg(F.C6 + (size_t)p * F.n + po) * irm2;
      w = a.w[k];
      px = (px - o.cm[0]) * irm; py = (py - o.cm[1]) * irm; pz = (pz - o.cm[2]) * irm;
      const double qx = (q4.x - o.cf[0]) * irf, qy = (q4.y - o.cf[1]) * irf, qz = (q4.z - o.cf[2]) * irf;
      x0 = px - o.c[0]; x1 = py - o.c[1]; x2 = pz - o.c[2];
      // Cij (symmetric 6) = A Cp A^T + J Sigma J^T + Cq
      double Cij[6];
      {
        // rows of A Cp
        double AC[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double a0 = o.A[3 * r], a1 = o.A[3 * r + 1], a2 = o.A[3 * r + 2];
          AC[3 * r] = a0 * Cp[0] + a1 * Cp[1] + a2 * Cp[2];
          AC[3 * r + 1] = a0 * Cp[1] + a1 * Cp[3] + a2 * Cp[4];
          AC[3 * r + 2] = a0 * Cp[2] + a1 * Cp[4] + a2 * Cp[5];
        }
        Cij[0] = AC[0] * o.A[0] + AC[1] * o.A[1] + AC[2] * o.A[2] + Cq[0];
        Cij[1] = AC[0] * o.A[3] + AC[1] * o.A[4] + AC[2] * o.A[5] + Cq[1];
        Cij[2] = AC[0] * o.A[6] + AC[1] * o.A[7] + AC[2] * o.A[8] + Cq[2];
        Cij[3] = AC[3] * o.A[3] + AC[4] * o.A[4] + AC[5] * o.A[5] + Cq[3];
        Cij[4] = AC[3] * o.A[6] + AC[4] * o.A[7] + AC[5] * o.A[8] + Cq[4];
        Cij[5] = AC[6] * o.A[6] + AC[7] * o.A[7] + AC[8] * o.A[8] + Cq[5];
        const double xt[4] = {x0, x1, x2, 1.0};
        int pair = 0;
#pragma unroll
        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
          for (int bb = aa; bb < 4; ++bb) {
            const double xx = xt[aa] * xt[bb];
#pragma unroll
            for (int p = 0; p < 6; ++p) Cij[p] += xx * Bs[pair][p];
            ++pair;
          }
      }
      // inverse and determinant of the symmetric 3x3
      const double c00 = Cij[3] * Cij[5] - Cij[4] * Cij[4], c01 = Cij[2] * Cij[4] - Cij[1] * Cij[5], c02 = Cij[1] * Cij[4] - Cij[2] * Cij[3];
      const double det = Cij[0] * c00 + Cij[1] * c01 + Cij[2] * c02;
      const double id = 1.0 / det;
      const double Ci[6] = {c00 * id, c01 * id, c02 * id, (Cij[0] * Cij[5] - Cij[2] * Cij[2]) * id, (Cij[1] * Cij[2] - Cij[0] * Cij[4]) * id,
                            (Cij[0] * Cij[3] - Cij[1] * Cij[1]) * id};
      // err = q - T(p)   (cdcl_obj_fun_cov.txx:84-86)
      const double yx = o.A[0] * x0 + o.A[1] * x1 + o.A[2] * x2 + o.t[0];
      const double yy = o.A[3] * x0 + o.A[4] * x1 + o.A[5] * x2 + o.t[1];
      const double yz = o.A[6] * x0 + o.A[7] * x1 + o.A[8] * x2 + o.t[2];
      const double e0 = qx - yx, e1 = qy - yy, e2 = qz - yz;
      const double v0 = Ci[0] * e0 + Ci[1] * e1 + Ci[2] * e2, v1 = Ci[1] * e0 + Ci[3] * e1 + Ci[4] * e2, v2 = Ci[2] * e0 + Ci[4] * e1 + Ci[5] * e2;
      const double r = e0 * v0 + e1 * v1 + e2 * v2;
      fterm = w * (lrr_rho(r) + log(det));
      const double rp = lrr_rho_p(r);
      S[0] = Ci[0] - rp * v0 * v0; S[1] = Ci[1] - rp * v0 * v1; S[2] = Ci[2] - rp * v0 * v2;
      S[3] = Ci[3] - rp * v1 * v1; S[4] = Ci[4] - rp * v1 * v2; S[5] = Ci[5] - rp * v2 * v2;
    }
    acc[24] += fterm;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int src = qbase + m;
      const double wm = __shfl_sync(0xffffffffu, w, src);
      const double xm[4] = {__shfl_sync(0xffffffffu, x0, src), __shfl_sync(0xffffffffu, x1, src), __shfl_sync(0xffffffffu, x2, src), 1.0};
      double Sm[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) Sm[p] = __shfl_sync(0xffffffffu, S[p], src);
      const double wxa = wm * (j == 0 ? xm[0] : (j == 1 ? xm[1] : (j == 2 ? xm[2] : 1.0)));
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double tt = wxa * xm[b];
#pragma unroll
        for (int p = 0; p < 6; ++p) acc[b * 6 + p] += tt * Sm[p];
      }
    }
  }
#endif
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
  // one match per lane and trip; LRR_OBJ_MINB resident blocks per SM (LRR_MAX_BLOCKS rows of partial sums)
  // (the pipelined loop lets a lane without a match load entry 0 of both sets: an empty set means no match, so no trip at all)
  DirArgs aa = a;
  if (a.to.n <= 0 || a.from_all.n <= 0) aa.M = 0;
  long long b = ((long long)aa.M + LRR_OBJ_THREADS - 1) / LRR_OBJ_THREADS, cap = (long long)lrr_num_sms() * LRR_OBJ_MINB;
  if (cap > LRR_MAX_BLOCKS) cap = LRR_MAX_BLOCKS;
  if (b > cap) b = cap; if (b < 1) b = 1;
  k_objcov<<<(int)b, LRR_OBJ_THREADS, 0, st>>>(aa, o, partials, ticket); ++*launches;
}

// ---- cdcl_obj_fun_par<3,12>::compute (cdcl/cdcl_obj_fun_par.txx:46-144): the parameter objective of the CDC M-step -------------------
// Same streams, same Cij as above except that the transfer-error covariance J Sigma J^T is frozen at the transform the objective was
// built with (o.Sigma, :27-38) while A, t are the parameters being evaluated.  With v = Cij^-1 e and the unit-element matrices
// Mkl = dJp/dth_(3k+l) (cdcl_trans_affine.txx:199-214) the correction terms collapse to rank-one forms:
//   dCij/dth_(3k+l) = (A Cp)[:,l] e_k^T + e_k (A Cp)[:,l]^T
//   correct_(3k+l) = -v^T dCij v = -2 v_k (Cp A^T v)_l          correctS_(3k+l) = tr(Cij^-1 dCij) = 2 (Cij^-1 A Cp)_(kl)
//   g_(3k+l) += w ( rho' (2 x~_l v_k - 2 v_k (Cp A^T v)_l) + 2 (Cij^-1 A Cp)_(kl) ),   g_(9+k) += w rho' 2 v_k      (inliers only, :97)
// 12 + 1 sums per lane, no exchange between lanes.
__global__ void __launch_bounds__(LRR_OBJ_THREADS, 2) k_objpar(const __grid_constant__ DirArgs a, const __grid_constant__ ObjArgs o, double* partials, unsigned int* ticket) {
  __shared__ ObjShared sh;
  __shared__ double Bs[10][6];
  GridScope scope; scope.partials = partials; scope.ticket = ticket;
  if (threadIdx.x < 60) {
    const int pair = threadIdx.x / 6, p = threadIdx.x % 6;
    int aa = 0, bb = 0, cnt = 0;
    for (int i = 0; i < 4; ++i) for (int j2 = i; j2 < 4; ++j2) { if (cnt == pair) { aa = i; bb = j2; } ++cnt; }
    const int d = p < 3 ? 0 : (p < 5 ? 1 : 2), e = p < 3 ? p : (p < 5 ? p - 2 : 2);
    const double s1 = o.Sigma[obj_idx(d, aa) * 12 + obj_idx(e, bb)], s2 = o.Sigma[obj_idx(e, aa) * 12 + obj_idx(d, bb)];
    Bs[pair][p] = aa == bb ? s1 : s1 + s2;
  }
  __syncthreads();
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  const FsetDev& F = a.from_all; const FsetDev& T = a.to;
  const double irm = 1.0 / o.rm, irf = 1.0 / o.rf, irm2 = irm * irm, irf2 = irf * irf;
  double acc[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) acc[i] = 0.0;
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int s = a.sidx[k];
    if (s < 0) continue;
    const int i = lrr_from_index(a, k);
    const int po = a.forig ? a.forig[i] : i;
    double px, py, pz; lrr_from_loc(a, sc, mapped, i, px, py, pz);
    const double4 q4 = lrr_ld_pos4(T.grid.pos4 + s);
    double Cq[6];
    {
      const double2* cr = reinterpret_cast<const double2*>(T.crec + 8 * (size_t)s);
      const double2 c0 = __ldg(cr), c1 = __ldg(cr + 1), c2 = __ldg(cr + 2);
      Cq[0] = c0.x * irf2; Cq[1] = c0.y * irf2; Cq[2] = c1.x * irf2; Cq[3] = c1.y * irf2; Cq[4] = c2.x * irf2; Cq[5] = c2.y * irf2;
    }
    double Cp[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) Cp[p] = __ldg(F.C6 + (size_t)p * F.n + po) * irm2;
    const double w = a.w[k];
    px = (px - o.cm[0]) * irm; py = (py - o.cm[1]) * irm; pz = (pz - o.cm[2]) * irm;
    const double qx = (q4.x - o.cf[0]) * irf, qy = (q4.y - o.cf[1]) * irf, qz = (q4.z - o.cf[2]) * irf;
    const double xt[4] = {px - o.c[0], py - o.c[1], pz - o.c[2], 1.0};
    double AC[9], Cij[6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double a0 = o.A[3 * r], a1 = o.A[3 * r + 1], a2 = o.A[3 * r + 2];
      AC[3 * r] = a0 * Cp[0] + a1 * Cp[1] + a2 * Cp[2];
      AC[3 * r + 1] = a0 * Cp[1] + a1 * Cp[3] + a2 * Cp[4];
      AC[3 * r + 2] = a0 * Cp[2] + a1 * Cp[4] + a2 * Cp[5];
    }
    Cij[0] = AC[0] * o.A[0] + AC[1] * o.A[1] + AC[2] * o.A[2] + Cq[0];
    Cij[1] = AC[0] * o.A[3] + AC[1] * o.A[4] + AC[2] * o.A[5] + Cq[1];
    Cij[2] = AC[0] * o.A[6] + AC[1] * o.A[7] + AC[2] * o.A[8] + Cq[2];
    Cij[3] = AC[3] * o.A[3] + AC[4] * o.A[4] + AC[5] * o.A[5] + Cq[3];
    Cij[4] = AC[3] * o.A[6] + AC[4] * o.A[7] + AC[5] * o.A[8] + Cq[4];
    Cij[5] = AC[6] * o.A[6] + AC[7] * o.A[7] + AC[8] * o.A[8] + Cq[5];
    {
      int pair = 0;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = aa; bb < 4; ++bb) {
          const double xx = xt[aa] * xt[bb];
#pragma unroll
          for (int p = 0; p < 6; ++p) Cij[p] += xx * Bs[pair][p];
          ++pair;
        }
    }
    const double c00 = Cij[3] * Cij[5] - Cij[4] * Cij[4], c01 = Cij[2] * Cij[4] - Cij[1] * Cij[5], c02 = Cij[1] * Cij[4] - Cij[2] * Cij[3];
    const double det = Cij[0] * c00 + Cij[1] * c01 + Cij[2] * c02;
    const double id = 1.0 / det;
    const double Ci[6] = {c00 * id, c01 * id, c02 * id, (Cij[0] * Cij[5] - Cij[2] * Cij[2]) * id, (Cij[1] * Cij[2] - Cij[0] * Cij[4]) * id,
                          (Cij[0] * Cij[3] - Cij[1] * Cij[1]) * id};
    // err = T(p) - q   (cdcl_obj_fun_par.txx:84)
    const double e0 = o.A[0] * xt[0] + o.A[1] * xt[1] + o.A[2] * xt[2] + o.t[0] - qx;
    const double e1 = o.A[3] * xt[0] + o.A[4] * xt[1] + o.A[5] * xt[2] + o.t[1] - qy;
    const double e2 = o.A[6] * xt[0] + o.A[7] * xt[1] + o.A[8] * xt[2] + o.t[2] - qz;
    const double v[3] = {Ci[0] * e0 + Ci[1] * e1 + Ci[2] * e2, Ci[1] * e0 + Ci[3] * e1 + Ci[4] * e2, Ci[2] * e0 + Ci[4] * e1 + Ci[5] * e2};
    const double r = e0 * v[0] + e1 * v[1] + e2 * v[2];
    const double rp = lrr_rho_p(r);
    if (!(rp > 0)) continue;   // inliers only: a point may become an outlier during the parameter search (:94-97)
    acc[12] += w * (lrr_rho(r) + log(det));
    // u = Cp A^T v = (A Cp)^T v
    const double u[3] = {AC[0] * v[0] + AC[3] * v[1] + AC[6] * v[2], AC[1] * v[0] + AC[4] * v[1] + AC[7] * v[2], AC[2] * v[0] + AC[5] * v[1] + AC[8] * v[2]};
    const double Cif[9] = {Ci[0], Ci[1], Ci[2], Ci[1], Ci[3], Ci[4], Ci[2], Ci[4], Ci[5]};
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double Mkl = Cif[3 * kk] * AC[l] + Cif[3 * kk + 1] * AC[3 + l] + Cif[3 * kk + 2] * AC[6 + l];   // (Cij^-1 A Cp)_(kl)
        acc[3 * kk + l] += w * (rp * (2.0 * xt[l] * v[kk] - 2.0 * v[kk] * u[l]) + 2.0 * Mkl);
      }
      acc[9 + kk] += w * rp * 2.0 * v[kk];
    }
  }
  if (scope.reduce<13>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x < 13) sc->mom[threadIdx.x] = sh.out[threadIdx.x];
  }
}

void lrr_k_objpar_launch(const DirArgs& a, const double* A, const double* t, const double* c, const double* norm8, const double* Sigma_dev,
                         double* partials, unsigned int* ticket, cudaStream_t st, long long* launches) {
  ObjArgs o;
  for (int i = 0; i < 9; ++i) o.A[i] = A[i];
  for (int i = 0; i < 3; ++i) { o.t[i] = t[i]; o.c[i] = c[i]; o.cm[i] = norm8[i]; o.cf[i] = norm8[4 + i]; }
  o.rm = norm8[3]; o.rf = norm8[7]; o.Sigma = Sigma_dev;
  long long b = ((long long)a.M + LRR_OBJ_THREADS - 1) / LRR_OBJ_THREADS, cap = (long long)lrr_num_sms() * 2;
  if (cap > LRR_MAX_BLOCKS) cap = LRR_MAX_BLOCKS;
  if (b > cap) b = cap; if (b < 1) b = 1;
  k_objpar<<<(int)b, LRR_OBJ_THREADS, 0, st>>>(a, o, partials, ticket); ++*launches;
}
