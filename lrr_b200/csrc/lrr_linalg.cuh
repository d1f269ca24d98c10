// Small dense fp64 linear algebra, usable from host and device (sm_100a).  Replaces the VXL calls on the
// path: vnl_svd<double>::inverse() on the dof x dof normal matrix (cdcl/cdcl_estimation_ICP_matching_all.txx:1073-1074,
// cdcl sym :643-644, rrl :839-840, rrl sym :880-881), vnl_svd of 3x3 in cdcl_trans_*::inverse()
// (cdcl/cdcl_trans_affine.txx:56-58) and in recompose_increment (cdcl/cdcl_trans_rigid3d.cxx:318-320).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LRR_HD __host__ __device__ __forceinline__
#else
#define LRR_HD inline
#endif

// One-sided Jacobi SVD, n <= 12, row-major.  M = U diag(w) V^T.
template <int NMAX>
LRR_HD void lrr_jacobi_svd(int n, const double* M, double* U, double* w, double* V) {
  double G[NMAX * NMAX];
  for (int i = 0; i < n * n; ++i) G[i] = M[i];
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
      double alpha = 0, beta = 0, gamma = 0;
      for (int i = 0; i < n; ++i) {
        double gp = G[i * n + p], gq = G[i * n + q];
        alpha += gp * gp; beta += gq * gq; gamma += gp * gq;
      }
      if (gamma == 0.0) continue;
      double denom = sqrt(alpha * beta);
      if (denom > 0) { double r = fabs(gamma) / denom; if (r > off) off = r; }
      double zeta = (beta - alpha) / (2.0 * gamma);
      double t = ((zeta >= 0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
      for (int i = 0; i < n; ++i) {
        double gp = G[i * n + p], gq = G[i * n + q];
        G[i * n + p] = c * gp - s * gq; G[i * n + q] = s * gp + c * gq;
        double vp = V[i * n + p], vq = V[i * n + q];
        V[i * n + p] = c * vp - s * vq; V[i * n + q] = s * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
  for (int j = 0; j < n; ++j) {
    double nrm = 0;
    for (int i = 0; i < n; ++i) nrm += G[i * n + j] * G[i * n + j];
    nrm = sqrt(nrm);
    w[j] = nrm;
    for (int i = 0; i < n; ++i) U[i * n + j] = (nrm > 0) ? G[i * n + j] / nrm : 0.0;
  }
}

// Pseudo-inverse as vnl_svd::inverse(): V diag(1/w) U^T with exactly-zero singular values dropped.
template <int NMAX>
LRR_HD void lrr_svd_inverse(int n, const double* M, double* inv) {
  double U[NMAX * NMAX], V[NMAX * NMAX], w[NMAX];
  lrr_jacobi_svd<NMAX>(n, M, U, w, V);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int k = 0; k < n; ++k) if (w[k] > 0.0) s += V[i * n + k] * (1.0 / w[k]) * U[j * n + k];
    inv[i * n + j] = s;
  }
}

// SPD inverse through Cholesky (A = L L^T).  Returns false when A is not safely positive definite, in
// which case the caller falls back to lrr_svd_inverse (the reference's pseudo-inverse semantics).
template <int NMAX>
LRR_HD bool lrr_chol_inverse(int n, const double* A, double* inv) {
  double L[NMAX * NMAX];
  double maxd = 0;
  for (int i = 0; i < n; ++i) if (A[i * n + i] > maxd) maxd = A[i * n + i];
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 1e-13 * maxd)) return false;
    double d = sqrt(s);
    L[j * n + j] = d;
    double id = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t * id;
    }
  }
  // inv = L^-T L^-1, column by column
  for (int c = 0; c < n; ++c) {
    double y[NMAX];
    for (int i = 0; i < n; ++i) {
      double t = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) t -= L[i * n + k] * y[k];
      y[i] = t / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double t = y[i];
      for (int k = i + 1; k < n; ++k) t -= L[k * n + i] * y[k];
      y[i] = t / L[i * n + i];
    }
    for (int i = 0; i < n; ++i) inv[i * n + c] = y[i];
  }
  return true;
}

LRR_HD double lrr_det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
LRR_HD void lrr_inv3(const double* m, double* o) {
  double id = 1.0 / lrr_det3(m);
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// U V^T (orthogonal polar factor) of a 3x3
LRR_HD void lrr_orthogonalize3(const double* M, double* R) {
  double U[9], V[9], w[3];
  lrr_jacobi_svd<3>(3, M, U, w, V);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += U[i * 3 + k] * V[j * 3 + k];
    R[i * 3 + j] = s;
  }
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi), eigenvalues ascending, eigenvectors in columns of V
// (vnl_symmetric_eigensystem_compute as used by cdcl sheet_angles/tube_angles, cdcl :300-306,:351-362)
LRR_HD void lrr_sym_eig3(const double* C, double* V, double* D) {
  double a[9];
  for (int i = 0; i < 9; ++i) { a[i] = C[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 50; ++sweep) {
    double off = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      double apq = a[p * 3 + q];
      if (apq == 0.0) continue;
      double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
      double t = ((theta >= 0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) { double akp = a[k * 3 + p], akq = a[k * 3 + q]; a[k * 3 + p] = c * akp - s * akq; a[k * 3 + q] = s * akp + c * akq; }
      for (int k = 0; k < 3; ++k) { double apk = a[p * 3 + k], aqk = a[q * 3 + k]; a[p * 3 + k] = c * apk - s * aqk; a[q * 3 + k] = s * apk + c * aqk; }
      for (int k = 0; k < 3; ++k) { double vkp = V[k * 3 + p], vkq = V[k * 3 + q]; V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq; }
    }
  }
  double d[3] = {a[0], a[4], a[8]};
  int o0 = 0, o1 = 1, o2 = 2, tmp;
  if (d[o0] > d[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
  if (d[o1] > d[o2]) { tmp = o1; o1 = o2; o2 = tmp; }
  if (d[o0] > d[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
  int ord[3] = {o0, o1, o2};
  double Vs[9];
  for (int j = 0; j < 3; ++j) { D[j] = d[ord[j]]; for (int i = 0; i < 3; ++i) Vs[i * 3 + j] = V[i * 3 + ord[j]]; }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}
