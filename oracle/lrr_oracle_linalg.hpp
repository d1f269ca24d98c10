// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Small dense linear algebra used by the CPU restatement of the LRR hot path.
//
// The reference calls into VXL (vnl_svd / vnl_inverse / vnl_determinant /
// vnl_symmetric_eigensystem), which is NOT vendored under /root/reference and is
// unpinned (CMakeLists.txt:12 FIND_PACKAGE(VXL), no version).  Mathematically these
// are the (pseudo-)inverse, determinant and symmetric eigen-decomposition, so any
// accurate fp64 implementation pins them to tolerance.
// Here: one-sided Jacobi SVD, cyclic Jacobi symmetric eigensolver, cofactor 3x3.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

namespace lrro {

// One-sided (Hestenes) Jacobi SVD of an n x n row-major matrix: M = U diag(w) V^T.
// Replaces vnl_svd<double> (call sites cdcl/cdcl_estimation_ICP_matching_all.txx:1073,
// cdcl/cdcl_trans_affine.txx:56, cdcl/cdcl_trans_rigid3d.cxx:115,318).
inline void jacobi_svd(int n, const double* M, double* U, double* w, double* V) {
  std::vector<double> G(M, M + n * n);  // columns get orthogonalised
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
      double alpha = 0, beta = 0, gamma = 0;
      for (int i = 0; i < n; ++i) {
        alpha += G[i * n + p] * G[i * n + p];
        beta  += G[i * n + q] * G[i * n + q];
        gamma += G[i * n + p] * G[i * n + q];
      }
      if (gamma == 0.0) continue;
      double denom = std::sqrt(alpha * beta);
      if (denom > 0) off = std::max(off, std::fabs(gamma) / denom);
      double zeta = (beta - alpha) / (2.0 * gamma);
      double t = ((zeta >= 0) ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
      double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
      for (int i = 0; i < n; ++i) {
        double gp = G[i * n + p], gq = G[i * n + q];
        G[i * n + p] = c * gp - s * gq;
        G[i * n + q] = s * gp + c * gq;
        double vp = V[i * n + p], vq = V[i * n + q];
        V[i * n + p] = c * vp - s * vq;
        V[i * n + q] = s * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
  for (int j = 0; j < n; ++j) {
    double nrm = 0;
    for (int i = 0; i < n; ++i) nrm += G[i * n + j] * G[i * n + j];
    nrm = std::sqrt(nrm);
    w[j] = nrm;
    for (int i = 0; i < n; ++i) U[i * n + j] = (nrm > 0) ? G[i * n + j] / nrm : 0.0;
  }
}

// vnl_svd<double>::inverse(): V diag(1/w) U^T, singular values that are exactly zero dropped
// (vnl_svd's default zero_out_absolute(0.0)).
inline void svd_inverse(int n, const double* M, double* inv) {
  std::vector<double> U(n * n), V(n * n), w(n);
  jacobi_svd(n, M, U.data(), w.data(), V.data());
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int k = 0; k < n; ++k) if (w[k] > 0.0) s += V[i * n + k] * (1.0 / w[k]) * U[j * n + k];
    inv[i * n + j] = s;
  }
}

inline void svd_singular_values(int n, const double* M, double* w) {
  std::vector<double> U(n * n), V(n * n);
  jacobi_svd(n, M, U.data(), w, V.data());
}

// U V^T of a 3x3 (orthogonalisation in cdcl_trans_rigid3d::recompose_increment, rigid3d.cxx:318-320).
inline void svd_orthogonalize3(const double* M, double* R) {
  double U[9], V[9], w[3];
  jacobi_svd(3, M, U, w, V);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += U[i * 3 + k] * V[j * 3 + k];
    R[i * 3 + j] = s;
  }
}

inline double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// vnl_inverse of a 3x3 (cofactor formula, as vnl_inverse does for fixed 3x3).
inline void inv3(const double* m, double* o) {
  double d = det3(m);
  double id = 1.0 / d;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3; eigenvalues ascending, eigenvectors in
// the COLUMNS of V (vnl_symmetric_eigensystem_compute convention used at
// cdcl/cdcl_estimation_ICP_matching_all.txx:300-306: column 0 = smallest eigenvalue).
// Error projector of a plain 2-D cdcl_feature: outer product of the eigenvector of the SMALLEST eigenvalue of its covariance
// (the point normal; cdcl/cdcl_estimation_ICP.txx:470-477 via vnl_symmetric_eigensystem_compute).  Closed form for the symmetric
// 2x2 case; P3 = (xx, xy, yy).  Isotropic covariances (no defined normal) give the x axis, like an identity eigenvector matrix.
inline void projector2_from_cov(double a, double b, double c, double* P3) {
  const double half = 0.5 * (a - c), rad = std::sqrt(half * half + b * b);
  const double lam = 0.5 * (a + c) - rad;
  double v0, v1;
  if (a >= c) { v0 = b; v1 = lam - a; } else { v0 = lam - c; v1 = b; }
  const double n2 = v0 * v0 + v1 * v1;
  if (!(n2 > 0.0)) { P3[0] = 1.0; P3[1] = 0.0; P3[2] = 0.0; return; }
  const double inv = 1.0 / std::sqrt(n2);
  v0 *= inv; v1 *= inv;
  P3[0] = v0 * v0; P3[1] = v0 * v1; P3[2] = v1 * v1;
}

inline void sym_eig3(const double* C, double* V, double* D) {
  double a[9];
  std::memcpy(a, C, sizeof(a));
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 50; ++sweep) {
    double off = std::fabs(a[1]) + std::fabs(a[2]) + std::fabs(a[5]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      double apq = a[p * 3 + q];
      if (apq == 0.0) continue;
      double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
      double t = ((theta >= 0) ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) {  // A <- A J
        double akp = a[k * 3 + p], akq = a[k * 3 + q];
        a[k * 3 + p] = c * akp - s * akq;
        a[k * 3 + q] = s * akp + c * akq;
      }
      for (int k = 0; k < 3; ++k) {  // A <- J^T A
        double apk = a[p * 3 + k], aqk = a[q * 3 + k];
        a[p * 3 + k] = c * apk - s * aqk;
        a[q * 3 + k] = s * apk + c * aqk;
      }
      for (int k = 0; k < 3; ++k) {
        double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {a[0], a[4], a[8]};
  std::sort(order, order + 3, [&](int x, int y) { return d[x] < d[y]; });
  double Vs[9];
  for (int j = 0; j < 3; ++j) {
    D[j] = d[order[j]];
    for (int i = 0; i < 3; ++i) Vs[i * 3 + j] = V[i * 3 + order[j]];
  }
  std::memcpy(V, Vs, sizeof(Vs));
}

// Upper Cholesky factor U (C = U^T U) of an SPD n x n (vnl_cholesky::upper_triangle,
// cdcl/cdcl_estimation.txx:284-285).  Returns false if not positive definite.
inline bool cholesky_upper(int n, const double* C, double* U) {
  std::fill(U, U + n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = i; j < n; ++j) {
      double s = C[i * n + j];
      for (int k = 0; k < i; ++k) s -= U[k * n + i] * U[k * n + j];
      if (i == j) {
        if (s <= 0) return false;
        U[i * n + i] = std::sqrt(s);
      } else {
        U[i * n + j] = s / U[i * n + i];
      }
    }
  }
  return true;
}

}  // namespace lrro
