// vnl_shim_algo.h — stand-in for the vnl/algo routines the reference's hot path calls: vnl_svd, vnl_symmetric_eigensystem,
// vnl_inverse, vnl_determinant, vnl_cholesky (+ interface-only vnl_cost_function / vnl_lbfgs).  TEST INFRASTRUCTURE ONLY.
// Independent implementations of textbook algorithms (upstream VXL wraps netlib dsvdc / rs / lbfgs, none of which is
// available here — SURVEY.md §8(c) items (3),(4)):
//   SVD                    : one-sided Jacobi (Hestenes) on the columns, singular values sorted descending
//   symmetric eigensystem  : cyclic two-sided Jacobi, eigenvalues ascending (vnl convention), V columns = eigenvectors
//   inverse / determinant  : Gaussian elimination with partial pivoting
#ifndef vnl_shim_algo_h_
#define vnl_shim_algo_h_

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>
#include "../vnl_shim_core.h"

template <class T>
class vnl_svd {
 public:
  // zero_out_tol >= 0: singular values <= tol are treated as zero (absolute); < 0: relative to the largest.
  vnl_svd(const vnl_matrix<T>& M, double zero_out_tol = 0.0) : m_(M.rows()), n_(M.cols()), U_(M), V_(M.cols(), M.cols()), w_(M.cols()) {
    V_.set_identity();
    const unsigned m = m_, n = n_;
    for (int sweep = 0; sweep < 100; ++sweep) {
      double off = 0.0;
      for (unsigned p = 0; p + 1 < n; ++p)
        for (unsigned q = p + 1; q < n; ++q) {
          double a = 0, b = 0, c = 0;
          for (unsigned i = 0; i < m; ++i) { a += U_(i, p) * U_(i, p); b += U_(i, q) * U_(i, q); c += U_(i, p) * U_(i, q); }
          if (c == 0.0) continue;
          off = std::max(off, std::fabs(c) / std::sqrt(a * b));
          if (std::fabs(c) <= 1e-17 * std::sqrt(a * b)) continue;
          double zeta = (b - a) / (2.0 * c);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
          for (unsigned i = 0; i < m; ++i) { double up = U_(i, p), uq = U_(i, q); U_(i, p) = cs * up - sn * uq; U_(i, q) = sn * up + cs * uq; }
          for (unsigned i = 0; i < n; ++i) { double vp = V_(i, p), vq = V_(i, q); V_(i, p) = cs * vp - sn * vq; V_(i, q) = sn * vp + cs * vq; }
        }
      if (off < 1e-15) break;
    }
    for (unsigned j = 0; j < n; ++j) {
      double s = 0; for (unsigned i = 0; i < m; ++i) s += U_(i, j) * U_(i, j);
      w_[j] = std::sqrt(s);
      if (w_[j] > 0) for (unsigned i = 0; i < m; ++i) U_(i, j) /= w_[j];
    }
    // sort descending
    std::vector<unsigned> idx(n); std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return w_[a] > w_[b]; });
    vnl_matrix<T> U2(m, n), V2(n, n); std::vector<T> w2(n);
    for (unsigned j = 0; j < n; ++j) { w2[j] = w_[idx[j]]; for (unsigned i = 0; i < m; ++i) U2(i, j) = U_(i, idx[j]); for (unsigned i = 0; i < n; ++i) V2(i, j) = V_(i, idx[j]); }
    U_ = U2; V_ = V2; w_ = w2;
    double tol = zero_out_tol >= 0 ? zero_out_tol : -zero_out_tol * (n ? std::fabs(w_[0]) : 0.0);
    winv_.assign(n, 0.0); rank_ = 0;
    for (unsigned j = 0; j < n; ++j) if (std::fabs(w_[j]) > tol) { winv_[j] = 1.0 / w_[j]; ++rank_; } else w_[j] = 0;
    // a zero column of U (exact zero singular value) must still be orthonormal for U()*V^T users; not needed by pinverse
  }
  const vnl_matrix<T>& U() const { return U_; }
  const vnl_matrix<T>& V() const { return V_; }
  T U(unsigned i, unsigned j) const { return U_(i, j); }
  T V(unsigned i, unsigned j) const { return V_(i, j); }
  vnl_matrix<T> W() const { vnl_matrix<T> W(n_, n_, 0.0); for (unsigned i = 0; i < n_; ++i) W(i, i) = w_[i]; return W; }
  T W(unsigned i) const { return w_[i]; }
  T W(unsigned i, unsigned j) const { return i == j ? w_[i] : T(0); }
  unsigned rank() const { return rank_; }
  T sigma_max() const { return w_[0]; }
  T sigma_min() const { return w_[n_ - 1]; }
  T well_condition() const { return w_[n_ - 1] / w_[0]; }
  T determinant_magnitude() const { T d = 1; for (unsigned i = 0; i < n_; ++i) d *= w_[i]; return d; }
  vnl_matrix<T> pinverse() const {
    vnl_matrix<T> r(n_, m_, 0.0);
    for (unsigned i = 0; i < n_; ++i) for (unsigned j = 0; j < m_; ++j) { T s = 0; for (unsigned k = 0; k < n_; ++k) s += V_(i, k) * winv_[k] * U_(j, k); r(i, j) = s; }
    return r;
  }
  vnl_matrix<T> inverse() const { return pinverse(); }
  vnl_vector<T> solve(const vnl_vector<T>& y) const { return pinverse() * y; }
  vnl_vector<T> nullvector() const { return V_.get_column(n_ - 1); }
 private:
  unsigned m_, n_, rank_;
  vnl_matrix<T> U_, V_;
  std::vector<T> w_, winv_;
};

// eigenvalues ascending in D, eigenvectors in the columns of V
inline bool vnl_symmetric_eigensystem_compute(const vnl_matrix<double>& A, vnl_matrix<double>& V, vnl_vector<double>& D) {
  const unsigned n = A.rows();
  vnl_matrix<double> a(A);
  V.set_size(n, n); V.set_identity(); D.set_size(n);
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0; for (unsigned p = 0; p < n; ++p) for (unsigned q = p + 1; q < n; ++q) off += a(p, q) * a(p, q);
    double diag = 0; for (unsigned p = 0; p < n; ++p) diag += a(p, p) * a(p, p);
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (unsigned p = 0; p + 1 < n; ++p)
      for (unsigned q = p + 1; q < n; ++q) {
        if (a(p, q) == 0.0) continue;
        double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (unsigned k = 0; k < n; ++k) { double akp = a(k, p), akq = a(k, q); a(k, p) = c * akp - s * akq; a(k, q) = s * akp + c * akq; }
        for (unsigned k = 0; k < n; ++k) { double apk = a(p, k), aqk = a(q, k); a(p, k) = c * apk - s * aqk; a(q, k) = s * apk + c * aqk; }
        for (unsigned k = 0; k < n; ++k) { double vkp = V(k, p), vkq = V(k, q); V(k, p) = c * vkp - s * vkq; V(k, q) = s * vkp + c * vkq; }
      }
  }
  std::vector<unsigned> idx(n); std::iota(idx.begin(), idx.end(), 0u);
  std::stable_sort(idx.begin(), idx.end(), [&](unsigned x, unsigned y) { return a(x, x) < a(y, y); });
  vnl_matrix<double> V2(n, n);
  for (unsigned j = 0; j < n; ++j) { D[j] = a(idx[j], idx[j]); for (unsigned i = 0; i < n; ++i) V2(i, j) = V(i, idx[j]); }
  V = V2;
  return true;
}
inline void vnl_symmetric_eigensystem_compute_eigenvals(double M11, double M12, double M13, double M22, double M23, double M33,
                                                        double& l1, double& l2, double& l3) {
  vnl_matrix<double> A(3, 3), V; vnl_vector<double> D;
  A(0, 0) = M11; A(0, 1) = A(1, 0) = M12; A(0, 2) = A(2, 0) = M13; A(1, 1) = M22; A(1, 2) = A(2, 1) = M23; A(2, 2) = M33;
  vnl_symmetric_eigensystem_compute(A, V, D);
  l1 = D[0]; l2 = D[1]; l3 = D[2];
}
template <class T>
class vnl_symmetric_eigensystem {
 public:
  vnl_symmetric_eigensystem(const vnl_matrix<T>& M) : n_(M.rows()) { vnl_vector<T> d; vnl_symmetric_eigensystem_compute(M, V, d); D = d; }
  vnl_matrix<T> V;
  vnl_vector<T> D;   // upstream: vnl_diag_matrix; D(i) / D[i] / get_eigenvalue(i) are what call sites use
  T get_eigenvalue(unsigned i) const { return D[i]; }
  vnl_vector<T> get_eigenvector(unsigned i) const { return V.get_column(i); }
  T determinant() const { T d = 1; for (unsigned i = 0; i < n_; ++i) d *= D[i]; return d; }
 private:
  unsigned n_;
};
// interface only (appears in commented-out / disabled reference code paths)
class vnl_generalized_eigensystem {
 public:
  vnl_generalized_eigensystem(const vnl_matrix<double>&, const vnl_matrix<double>&) { std::abort(); }
  vnl_matrix<double> V, D;
};

namespace vnl_shim_detail {
// Gauss-Jordan with partial pivoting on an n x n row-major copy; returns determinant, writes the inverse if inv != 0
inline double gauss(unsigned n, std::vector<double> a, double* inv) {
  std::vector<double> b((size_t)n * n, 0.0); for (unsigned i = 0; i < n; ++i) b[(size_t)i * n + i] = 1.0;
  double det = 1.0;
  for (unsigned c = 0; c < n; ++c) {
    unsigned piv = c; for (unsigned r = c + 1; r < n; ++r) if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
    if (a[(size_t)piv * n + c] == 0.0) { if (inv) for (size_t i = 0; i < (size_t)n * n; ++i) inv[i] = 0.0; return 0.0; }
    if (piv != c) { for (unsigned k = 0; k < n; ++k) { std::swap(a[(size_t)piv * n + k], a[(size_t)c * n + k]); std::swap(b[(size_t)piv * n + k], b[(size_t)c * n + k]); } det = -det; }
    double p = a[(size_t)c * n + c]; det *= p;
    for (unsigned k = 0; k < n; ++k) { a[(size_t)c * n + k] /= p; b[(size_t)c * n + k] /= p; }
    for (unsigned r = 0; r < n; ++r) if (r != c) {
      double f = a[(size_t)r * n + c]; if (f == 0.0) continue;
      for (unsigned k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; b[(size_t)r * n + k] -= f * b[(size_t)c * n + k]; }
    }
  }
  if (inv) for (size_t i = 0; i < (size_t)n * n; ++i) inv[i] = b[i];
  return det;
}
}  // namespace vnl_shim_detail

template <class T, unsigned n> inline vnl_matrix_fixed<T, n, n> vnl_inverse(const vnl_matrix_fixed<T, n, n>& m) {
  vnl_matrix_fixed<T, n, n> r;
  if (n == 1) { r(0, 0) = T(1) / m(0, 0); return r; }
  if (n == 2) {
    T det = m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
    r(0, 0) = m(1, 1) / det; r(0, 1) = -m(0, 1) / det; r(1, 0) = -m(1, 0) / det; r(1, 1) = m(0, 0) / det; return r;
  }
  if (n == 3) {   // adjugate / determinant (the closed form vnl_inverse uses for 3x3)
    T det = m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) + m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
    r(0, 0) = (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) / det; r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) / det; r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) / det;
    r(1, 0) = (m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2)) / det; r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) / det; r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) / det;
    r(2, 0) = (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0)) / det; r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) / det; r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) / det;
    return r;
  }
  std::vector<double> a(m.data_block(), m.data_block() + n * n);
  vnl_shim_detail::gauss(n, a, r.data_block());
  return r;
}
template <class T> inline vnl_matrix<T> vnl_inverse(const vnl_matrix<T>& m) {
  vnl_matrix<T> r(m.rows(), m.cols());
  std::vector<double> a(m.data_block(), m.data_block() + m.size());
  vnl_shim_detail::gauss(m.rows(), a, r.data_block());
  return r;
}
template <class T, unsigned n> inline T vnl_determinant(const vnl_matrix_fixed<T, n, n>& m) {
  if (n == 1) return m(0, 0);
  if (n == 2) return m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
  if (n == 3) return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) + m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  std::vector<double> a(m.data_block(), m.data_block() + n * n);
  return vnl_shim_detail::gauss(n, a, 0);
}
template <class T> inline T vnl_determinant(const vnl_matrix<T>& m) {
  std::vector<double> a(m.data_block(), m.data_block() + m.size());
  return vnl_shim_detail::gauss(m.rows(), a, 0);
}

class vnl_cholesky {
 public:
  vnl_cholesky(const vnl_matrix<double>& M) : n_(M.rows()), L_(M.rows(), M.rows(), 0.0), ok_(true) {
    for (unsigned j = 0; j < n_; ++j) {
      double s = M(j, j); for (unsigned k = 0; k < j; ++k) s -= L_(j, k) * L_(j, k);
      if (s <= 0) { ok_ = false; s = 0; }
      L_(j, j) = std::sqrt(s);
      for (unsigned i = j + 1; i < n_; ++i) { double t = M(i, j); for (unsigned k = 0; k < j; ++k) t -= L_(i, k) * L_(j, k); L_(i, j) = L_(j, j) != 0 ? t / L_(j, j) : 0.0; }
    }
  }
  vnl_matrix<double> lower_triangle() const { return L_; }
  vnl_matrix<double> upper_triangle() const { return L_.transpose(); }
  bool ok() const { return ok_; }
 private:
  unsigned n_; vnl_matrix<double> L_; bool ok_;
};

class vnl_cost_function {
 public:
  vnl_cost_function() : dim_(0) {}
  explicit vnl_cost_function(int n) : dim_(n) {}
  virtual ~vnl_cost_function() {}
  virtual void compute(const vnl_vector<double>& x, double* f, vnl_vector<double>* g) { if (f) *f = this->f(x); if (g) gradf(x, *g); }
  virtual double f(const vnl_vector<double>& x) { double v; compute(x, &v, 0); return v; }
  virtual void gradf(const vnl_vector<double>& x, vnl_vector<double>& g) { compute(x, 0, &g); }
  int get_number_of_unknowns() const { return dim_; }
  void set_number_of_unknowns(int n) { dim_ = n; }
 protected:
  int dim_;
};

// Interface only: the L-BFGS iterate trajectory of upstream (netlib lbfgs) is unpinned;
// f and g of the objective are checked pointwise instead.  minimize() is therefore not provided.
class vnl_lbfgs {
 public:
  vnl_lbfgs() : f_(0) { init(); }
  vnl_lbfgs(vnl_cost_function& f) : f_(&f) { init(); }
  virtual ~vnl_lbfgs() {}
  bool minimize(vnl_vector<double>&) { std::abort(); return false; }
  void set_max_function_evals(int n) { maxfev_ = n; }
  void set_verbose(bool) {}
  void set_check_derivatives(bool) {}
  void set_trace(bool) {}
  void set_f_tolerance(double) {}
  void set_x_tolerance(double) {}
  void set_g_tolerance(double) {}
  double get_start_error() const { return start_error_; }
  double get_end_error() const { return end_error_; }
  int get_failure_code() const { return failure_code_; }
  int get_num_evaluations() const { return num_evaluations_; }
  int memory; double line_search_accuracy; double default_step_length;
 protected:
  virtual bool report_iter() { return false; }
  void init() { memory = 5; line_search_accuracy = 0.9; default_step_length = 1.0; maxfev_ = 2000; start_error_ = end_error_ = 0; failure_code_ = 0; num_evaluations_ = 0; }
  vnl_cost_function* f_;
  int maxfev_;
  double start_error_, end_error_;
  int failure_code_, num_evaluations_;
};

#endif
