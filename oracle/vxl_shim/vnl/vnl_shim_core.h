// vnl_shim_core.h — stand-in for the part of VXL's vnl (vnl_vector, vnl_matrix, vnl_vector_fixed, vnl_matrix_fixed
// and small free functions) that the reference's hot-path sources use.  TEST INFRASTRUCTURE ONLY, written from
// scratch against the call sites in /root/reference (see ../vcl_shim_all.h); evaluation order of products is the
// plain textbook triple loop in IEEE fp64 (no FMA contraction: the recipe compiles with -ffp-contract=off).
#ifndef vnl_shim_core_h_
#define vnl_shim_core_h_

#include <cassert>
#include <cmath>
#include <iostream>
#include <vector>

template <class T> class vnl_matrix;
template <class T, unsigned n> class vnl_vector_fixed;
template <class T, unsigned R, unsigned C> class vnl_matrix_fixed;

// ------------------------------------------------------------------------------------------ vnl_vector
template <class T>
class vnl_vector {
 public:
  typedef T* iterator;
  typedef const T* const_iterator;
  vnl_vector() {}
  explicit vnl_vector(unsigned n) : d_(n) {}
  vnl_vector(unsigned n, const T& v) : d_(n, v) {}
  vnl_vector(const T* p, unsigned n) : d_(p, p + n) {}
  vnl_vector(unsigned n, unsigned m, const T* p) : d_(n) { for (unsigned i = 0; i < m && i < n; ++i) d_[i] = p[i]; }
  unsigned size() const { return (unsigned)d_.size(); }
  bool set_size(unsigned n) { d_.assign(n, T()); return true; }
  T& operator[](unsigned i) { return d_[i]; }
  const T& operator[](unsigned i) const { return d_[i]; }
  T& operator()(unsigned i) { return d_[i]; }
  const T& operator()(unsigned i) const { return d_[i]; }
  T get(unsigned i) const { return d_[i]; }
  void put(unsigned i, const T& v) { d_[i] = v; }
  T* data_block() { return d_.data(); }
  const T* data_block() const { return d_.data(); }
  iterator begin() { return d_.data(); }
  iterator end() { return d_.data() + d_.size(); }
  const_iterator begin() const { return d_.data(); }
  const_iterator end() const { return d_.data() + d_.size(); }
  vnl_vector& fill(const T& v) { for (auto& x : d_) x = v; return *this; }
  vnl_vector& copy_in(const T* p) { for (unsigned i = 0; i < size(); ++i) d_[i] = p[i]; return *this; }
  void copy_out(T* p) const { for (unsigned i = 0; i < size(); ++i) p[i] = d_[i]; }
  vnl_vector operator+(const vnl_vector& o) const { vnl_vector r(size()); for (unsigned i = 0; i < size(); ++i) r[i] = d_[i] + o[i]; return r; }
  vnl_vector operator-(const vnl_vector& o) const { vnl_vector r(size()); for (unsigned i = 0; i < size(); ++i) r[i] = d_[i] - o[i]; return r; }
  vnl_vector operator-() const { vnl_vector r(size()); for (unsigned i = 0; i < size(); ++i) r[i] = -d_[i]; return r; }
  vnl_vector operator*(const T& s) const { vnl_vector r(size()); for (unsigned i = 0; i < size(); ++i) r[i] = d_[i] * s; return r; }
  vnl_vector operator/(const T& s) const { vnl_vector r(size()); for (unsigned i = 0; i < size(); ++i) r[i] = d_[i] / s; return r; }
  vnl_vector& operator+=(const vnl_vector& o) { for (unsigned i = 0; i < size(); ++i) d_[i] += o[i]; return *this; }
  vnl_vector& operator-=(const vnl_vector& o) { for (unsigned i = 0; i < size(); ++i) d_[i] -= o[i]; return *this; }
  vnl_vector& operator*=(const T& s) { for (auto& x : d_) x *= s; return *this; }
  vnl_vector& operator/=(const T& s) { for (auto& x : d_) x /= s; return *this; }
  bool operator==(const vnl_vector& o) const { return d_ == o.d_; }
  bool operator!=(const vnl_vector& o) const { return !(d_ == o.d_); }
  T squared_magnitude() const { T s = 0; for (auto& x : d_) s += x * x; return s; }
  T magnitude() const { return std::sqrt(squared_magnitude()); }
  T two_norm() const { return magnitude(); }
  T sum() const { T s = 0; for (auto& x : d_) s += x; return s; }
  vnl_vector& normalize() { T m = magnitude(); if (m != 0) for (auto& x : d_) x /= m; return *this; }
  vnl_vector extract(unsigned len, unsigned start = 0) const { vnl_vector r(len); for (unsigned i = 0; i < len; ++i) r[i] = d_[start + i]; return r; }
  vnl_vector& update(const vnl_vector& v, unsigned start = 0) { for (unsigned i = 0; i < v.size(); ++i) d_[start + i] = v[i]; return *this; }
  // v <- v^T M  (row vector times matrix), in place
  vnl_vector& post_multiply(const vnl_matrix<T>& M);
  // v <- M v, in place
  vnl_vector& pre_multiply(const vnl_matrix<T>& M);
  const vnl_vector& as_ref() const { return *this; }
  vnl_vector& as_ref() { return *this; }
 private:
  std::vector<T> d_;
};
template <class T> inline vnl_vector<T> operator*(const T& s, const vnl_vector<T>& v) { return v * s; }
template <class T> inline T dot_product(const vnl_vector<T>& a, const vnl_vector<T>& b) { T s = 0; for (unsigned i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }
template <class T> inline std::ostream& operator<<(std::ostream& os, const vnl_vector<T>& v) { for (unsigned i = 0; i < v.size(); ++i) os << (i ? " " : "") << v[i]; return os; }
template <class T> using vnl_vector_ref = vnl_vector<T>;

// ------------------------------------------------------------------------------------------ vnl_matrix
template <class T>
class vnl_matrix {
 public:
  vnl_matrix() : r_(0), c_(0) {}
  vnl_matrix(unsigned r, unsigned c) : r_(r), c_(c), d_((size_t)r * c) {}
  vnl_matrix(unsigned r, unsigned c, const T& v) : r_(r), c_(c), d_((size_t)r * c, v) {}
  vnl_matrix(const T* p, unsigned r, unsigned c) : r_(r), c_(c), d_(p, p + (size_t)r * c) {}
  unsigned rows() const { return r_; }
  unsigned cols() const { return c_; }
  unsigned columns() const { return c_; }
  unsigned size() const { return r_ * c_; }
  bool set_size(unsigned r, unsigned c) { r_ = r; c_ = c; d_.assign((size_t)r * c, T()); return true; }
  T& operator()(unsigned i, unsigned j) { return d_[(size_t)i * c_ + j]; }
  const T& operator()(unsigned i, unsigned j) const { return d_[(size_t)i * c_ + j]; }
  T* operator[](unsigned i) { return d_.data() + (size_t)i * c_; }
  const T* operator[](unsigned i) const { return d_.data() + (size_t)i * c_; }
  T get(unsigned i, unsigned j) const { return (*this)(i, j); }
  void put(unsigned i, unsigned j, const T& v) { (*this)(i, j) = v; }
  T* data_block() { return d_.data(); }
  const T* data_block() const { return d_.data(); }
  vnl_matrix& fill(const T& v) { for (auto& x : d_) x = v; return *this; }
  vnl_matrix& fill_diagonal(const T& v) { for (unsigned i = 0; i < r_ && i < c_; ++i) (*this)(i, i) = v; return *this; }
  vnl_matrix& set_identity() { fill(T(0)); return fill_diagonal(T(1)); }
  vnl_matrix transpose() const { vnl_matrix t(c_, r_); for (unsigned i = 0; i < r_; ++i) for (unsigned j = 0; j < c_; ++j) t(j, i) = (*this)(i, j); return t; }
  vnl_matrix operator+(const vnl_matrix& o) const { vnl_matrix r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] += o.d_[i]; return r; }
  vnl_matrix operator-(const vnl_matrix& o) const { vnl_matrix r(*this); for (size_t i = 0; i < d_.size(); ++i) r.d_[i] -= o.d_[i]; return r; }
  vnl_matrix operator-() const { vnl_matrix r(*this); for (auto& x : r.d_) x = -x; return r; }
  vnl_matrix operator*(const T& s) const { vnl_matrix r(*this); for (auto& x : r.d_) x *= s; return r; }
  vnl_matrix operator/(const T& s) const { vnl_matrix r(*this); for (auto& x : r.d_) x /= s; return r; }
  vnl_matrix& operator+=(const vnl_matrix& o) { for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i]; return *this; }
  vnl_matrix& operator-=(const vnl_matrix& o) { for (size_t i = 0; i < d_.size(); ++i) d_[i] -= o.d_[i]; return *this; }
  vnl_matrix& operator*=(const T& s) { for (auto& x : d_) x *= s; return *this; }
  vnl_matrix& operator/=(const T& s) { for (auto& x : d_) x /= s; return *this; }
  vnl_matrix operator*(const vnl_matrix& o) const {
    assert(c_ == o.r_);
    vnl_matrix r(r_, o.c_, T(0));
    for (unsigned i = 0; i < r_; ++i) for (unsigned j = 0; j < o.c_; ++j) { T s = 0; for (unsigned k = 0; k < c_; ++k) s += (*this)(i, k) * o(k, j); r(i, j) = s; }
    return r;
  }
  vnl_vector<T> operator*(const vnl_vector<T>& v) const {
    vnl_vector<T> r(r_);
    for (unsigned i = 0; i < r_; ++i) { T s = 0; for (unsigned k = 0; k < c_; ++k) s += (*this)(i, k) * v[k]; r[i] = s; }
    return r;
  }
  bool operator==(const vnl_matrix& o) const { return r_ == o.r_ && c_ == o.c_ && d_ == o.d_; }
  vnl_vector<T> get_column(unsigned j) const { vnl_vector<T> v(r_); for (unsigned i = 0; i < r_; ++i) v[i] = (*this)(i, j); return v; }
  vnl_vector<T> get_row(unsigned i) const { vnl_vector<T> v(c_); for (unsigned j = 0; j < c_; ++j) v[j] = (*this)(i, j); return v; }
  vnl_matrix& set_column(unsigned j, const vnl_vector<T>& v) { for (unsigned i = 0; i < r_; ++i) (*this)(i, j) = v[i]; return *this; }
  vnl_matrix& set_column(unsigned j, const T* v) { for (unsigned i = 0; i < r_; ++i) (*this)(i, j) = v[i]; return *this; }
  vnl_matrix& set_row(unsigned i, const vnl_vector<T>& v) { for (unsigned j = 0; j < c_; ++j) (*this)(i, j) = v[j]; return *this; }
  vnl_matrix& set_row(unsigned i, const T* v) { for (unsigned j = 0; j < c_; ++j) (*this)(i, j) = v[j]; return *this; }
  vnl_matrix extract(unsigned r, unsigned c, unsigned top = 0, unsigned left = 0) const {
    vnl_matrix m(r, c); for (unsigned i = 0; i < r; ++i) for (unsigned j = 0; j < c; ++j) m(i, j) = (*this)(top + i, left + j); return m;
  }
  vnl_matrix& update(const vnl_matrix& m, unsigned top = 0, unsigned left = 0) {
    for (unsigned i = 0; i < m.rows(); ++i) for (unsigned j = 0; j < m.cols(); ++j) (*this)(top + i, left + j) = m(i, j); return *this;
  }
  T frobenius_norm() const { T s = 0; for (auto& x : d_) s += x * x; return std::sqrt(s); }
  T fro_norm() const { return frobenius_norm(); }
  const vnl_matrix& as_ref() const { return *this; }
  vnl_matrix& as_ref() { return *this; }
 private:
  unsigned r_, c_;
  std::vector<T> d_;
};
template <class T> inline vnl_matrix<T> operator*(const T& s, const vnl_matrix<T>& m) { return m * s; }
template <class T> inline std::ostream& operator<<(std::ostream& os, const vnl_matrix<T>& m) {
  for (unsigned i = 0; i < m.rows(); ++i) { for (unsigned j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j); os << '\n'; }
  return os;
}
template <class T> using vnl_matrix_ref = vnl_matrix<T>;
// <A,B> = sum_ij A_ij B_ij
template <class T> inline T dot_product(const vnl_matrix<T>& a, const vnl_matrix<T>& b) {
  T s = 0; for (unsigned i = 0; i < a.rows(); ++i) for (unsigned j = 0; j < a.cols(); ++j) s += a(i, j) * b(i, j); return s;
}
template <class T> inline vnl_vector<T>& vnl_vector<T>::post_multiply(const vnl_matrix<T>& M) {
  vnl_vector<T> r(M.cols());
  for (unsigned j = 0; j < M.cols(); ++j) { T s = 0; for (unsigned k = 0; k < M.rows(); ++k) s += d_[k] * M(k, j); r[j] = s; }
  d_.assign(r.begin(), r.end());
  return *this;
}
template <class T> inline vnl_vector<T>& vnl_vector<T>::pre_multiply(const vnl_matrix<T>& M) {
  vnl_vector<T> r = M * (*this);
  d_.assign(r.begin(), r.end());
  return *this;
}
template <class T> inline vnl_matrix<T> outer_product(const vnl_vector<T>& a, const vnl_vector<T>& b) {
  vnl_matrix<T> m(a.size(), b.size()); for (unsigned i = 0; i < a.size(); ++i) for (unsigned j = 0; j < b.size(); ++j) m(i, j) = a[i] * b[j]; return m;
}

// ------------------------------------------------------------------------------------------ vnl_vector_fixed
template <class T, unsigned n>
class vnl_vector_fixed {
 public:
  typedef T* iterator;
  typedef const T* const_iterator;
  enum { SIZE = n };
  vnl_vector_fixed() { for (unsigned i = 0; i < n; ++i) d_[i] = T(); }
  explicit vnl_vector_fixed(const T& v) { for (unsigned i = 0; i < n; ++i) d_[i] = v; }
  explicit vnl_vector_fixed(const T* p) { for (unsigned i = 0; i < n; ++i) d_[i] = p[i]; }
  vnl_vector_fixed(const T& a, const T& b) { static_assert(n == 2, "size"); d_[0] = a; d_[1] = b; }
  vnl_vector_fixed(const T& a, const T& b, const T& c) { static_assert(n == 3, "size"); d_[0] = a; d_[1] = b; d_[2] = c; }
  vnl_vector_fixed(const T& a, const T& b, const T& c, const T& e) { static_assert(n == 4, "size"); d_[0] = a; d_[1] = b; d_[2] = c; d_[3] = e; }
  vnl_vector_fixed(const vnl_vector<T>& v) { assert(v.size() == n); for (unsigned i = 0; i < n; ++i) d_[i] = v[i]; }
  vnl_vector_fixed& operator=(const vnl_vector<T>& v) { assert(v.size() == n); for (unsigned i = 0; i < n; ++i) d_[i] = v[i]; return *this; }
  unsigned size() const { return n; }
  T& operator[](unsigned i) { return d_[i]; }
  const T& operator[](unsigned i) const { return d_[i]; }
  T& operator()(unsigned i) { return d_[i]; }
  const T& operator()(unsigned i) const { return d_[i]; }
  T get(unsigned i) const { return d_[i]; }
  void put(unsigned i, const T& v) { d_[i] = v; }
  T* data_block() { return d_; }
  const T* data_block() const { return d_; }
  iterator begin() { return d_; }
  iterator end() { return d_ + n; }
  const_iterator begin() const { return d_; }
  const_iterator end() const { return d_ + n; }
  vnl_vector_fixed& fill(const T& v) { for (unsigned i = 0; i < n; ++i) d_[i] = v; return *this; }
  vnl_vector<T> as_vector() const { return vnl_vector<T>(d_, n); }
  vnl_vector<T> as_ref() const { return vnl_vector<T>(d_, n); }
  operator vnl_vector<T>() const { return vnl_vector<T>(d_, n); }
  vnl_vector_fixed operator-() const { vnl_vector_fixed r; for (unsigned i = 0; i < n; ++i) r[i] = -d_[i]; return r; }
  vnl_vector_fixed& operator+=(const vnl_vector_fixed& o) { for (unsigned i = 0; i < n; ++i) d_[i] += o[i]; return *this; }
  vnl_vector_fixed& operator-=(const vnl_vector_fixed& o) { for (unsigned i = 0; i < n; ++i) d_[i] -= o[i]; return *this; }
  vnl_vector_fixed& operator+=(const T& s) { for (unsigned i = 0; i < n; ++i) d_[i] += s; return *this; }
  vnl_vector_fixed& operator-=(const T& s) { for (unsigned i = 0; i < n; ++i) d_[i] -= s; return *this; }
  vnl_vector_fixed& operator*=(const T& s) { for (unsigned i = 0; i < n; ++i) d_[i] *= s; return *this; }
  vnl_vector_fixed& operator/=(const T& s) { for (unsigned i = 0; i < n; ++i) d_[i] /= s; return *this; }
  bool operator==(const vnl_vector_fixed& o) const { for (unsigned i = 0; i < n; ++i) if (d_[i] != o[i]) return false; return true; }
  bool operator!=(const vnl_vector_fixed& o) const { return !(*this == o); }
  T squared_magnitude() const { T s = 0; for (unsigned i = 0; i < n; ++i) s += d_[i] * d_[i]; return s; }
  T magnitude() const { return std::sqrt(squared_magnitude()); }
  T two_norm() const { return magnitude(); }
  T sum() const { T s = 0; for (unsigned i = 0; i < n; ++i) s += d_[i]; return s; }
  vnl_vector_fixed& normalize() { T m = magnitude(); if (m != 0) for (unsigned i = 0; i < n; ++i) d_[i] /= m; return *this; }
  vnl_vector<T> extract(unsigned len, unsigned start = 0) const { vnl_vector<T> r(len); for (unsigned i = 0; i < len; ++i) r[i] = d_[start + i]; return r; }
  vnl_vector_fixed& update(const vnl_vector<T>& v, unsigned start = 0) { for (unsigned i = 0; i < v.size(); ++i) d_[start + i] = v[i]; return *this; }
 private:
  T d_[n];
};
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator+(const vnl_vector_fixed<T, n>& a, const vnl_vector_fixed<T, n>& b) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] + b[i]; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator-(const vnl_vector_fixed<T, n>& a, const vnl_vector_fixed<T, n>& b) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] - b[i]; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator*(const vnl_vector_fixed<T, n>& a, const T& s) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] * s; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator*(const T& s, const vnl_vector_fixed<T, n>& a) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] * s; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator/(const vnl_vector_fixed<T, n>& a, const T& s) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] / s; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator+(const vnl_vector_fixed<T, n>& a, const vnl_vector<T>& b) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] + b[i]; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator-(const vnl_vector_fixed<T, n>& a, const vnl_vector<T>& b) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] - b[i]; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator+(const vnl_vector<T>& a, const vnl_vector_fixed<T, n>& b) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] + b[i]; return r; }
template <class T, unsigned n> inline vnl_vector_fixed<T, n> operator-(const vnl_vector<T>& a, const vnl_vector_fixed<T, n>& b) { vnl_vector_fixed<T, n> r; for (unsigned i = 0; i < n; ++i) r[i] = a[i] - b[i]; return r; }
template <class T, unsigned n> inline T dot_product(const vnl_vector_fixed<T, n>& a, const vnl_vector_fixed<T, n>& b) { T s = 0; for (unsigned i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
template <class T, unsigned n> inline T dot_product(const vnl_vector_fixed<T, n>& a, const vnl_vector<T>& b) { T s = 0; for (unsigned i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
template <class T, unsigned n> inline T dot_product(const vnl_vector<T>& a, const vnl_vector_fixed<T, n>& b) { T s = 0; for (unsigned i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
template <class T, unsigned n> inline std::ostream& operator<<(std::ostream& os, const vnl_vector_fixed<T, n>& v) { for (unsigned i = 0; i < n; ++i) os << (i ? " " : "") << v[i]; return os; }
template <class T, unsigned n> inline std::istream& operator>>(std::istream& is, vnl_vector_fixed<T, n>& v) { for (unsigned i = 0; i < n; ++i) is >> v[i]; return is; }
template <class T> inline vnl_vector_fixed<T, 3> vnl_cross_3d(const vnl_vector_fixed<T, 3>& a, const vnl_vector_fixed<T, 3>& b) {
  return vnl_vector_fixed<T, 3>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <class T> inline T vnl_cross_2d(const vnl_vector_fixed<T, 2>& a, const vnl_vector_fixed<T, 2>& b) { return a[0] * b[1] - a[1] * b[0]; }

// ------------------------------------------------------------------------------------------ vnl_matrix_fixed
template <class T, unsigned R, unsigned C>
class vnl_matrix_fixed {
 public:
  vnl_matrix_fixed() { for (unsigned i = 0; i < R * C; ++i) d_[i] = T(); }
  explicit vnl_matrix_fixed(const T& v) { for (unsigned i = 0; i < R * C; ++i) d_[i] = v; }
  explicit vnl_matrix_fixed(const T* p) { for (unsigned i = 0; i < R * C; ++i) d_[i] = p[i]; }
  vnl_matrix_fixed(const vnl_matrix<T>& m) { assert(m.rows() == R && m.cols() == C); for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) (*this)(i, j) = m(i, j); }
  vnl_matrix_fixed& operator=(const vnl_matrix<T>& m) { assert(m.rows() == R && m.cols() == C); for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) (*this)(i, j) = m(i, j); return *this; }
  vnl_matrix_fixed& operator=(const T& v) { return fill(v); }
  unsigned rows() const { return R; }
  unsigned cols() const { return C; }
  unsigned columns() const { return C; }
  unsigned size() const { return R * C; }
  T& operator()(unsigned i, unsigned j) { return d_[i * C + j]; }
  const T& operator()(unsigned i, unsigned j) const { return d_[i * C + j]; }
  T* operator[](unsigned i) { return d_ + i * C; }
  const T* operator[](unsigned i) const { return d_ + i * C; }
  T get(unsigned i, unsigned j) const { return (*this)(i, j); }
  void put(unsigned i, unsigned j, const T& v) { (*this)(i, j) = v; }
  T* data_block() { return d_; }
  const T* data_block() const { return d_; }
  vnl_matrix_fixed& fill(const T& v) { for (unsigned i = 0; i < R * C; ++i) d_[i] = v; return *this; }
  vnl_matrix_fixed& fill_diagonal(const T& v) { for (unsigned i = 0; i < R && i < C; ++i) (*this)(i, i) = v; return *this; }
  vnl_matrix_fixed& set_identity() { fill(T(0)); return fill_diagonal(T(1)); }
  vnl_matrix_fixed<T, C, R> transpose() const { vnl_matrix_fixed<T, C, R> t; for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) t(j, i) = (*this)(i, j); return t; }
  vnl_matrix<T> as_matrix() const { return vnl_matrix<T>(d_, R, C); }
  vnl_matrix<T> as_ref() const { return vnl_matrix<T>(d_, R, C); }
  operator vnl_matrix<T>() const { return vnl_matrix<T>(d_, R, C); }
  vnl_matrix_fixed operator-() const { vnl_matrix_fixed r; for (unsigned i = 0; i < R * C; ++i) r.d_[i] = -d_[i]; return r; }
  vnl_matrix_fixed& operator+=(const vnl_matrix_fixed& o) { for (unsigned i = 0; i < R * C; ++i) d_[i] += o.d_[i]; return *this; }
  vnl_matrix_fixed& operator-=(const vnl_matrix_fixed& o) { for (unsigned i = 0; i < R * C; ++i) d_[i] -= o.d_[i]; return *this; }
  vnl_matrix_fixed& operator+=(const vnl_matrix<T>& o) { for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) (*this)(i, j) += o(i, j); return *this; }
  vnl_matrix_fixed& operator-=(const vnl_matrix<T>& o) { for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) (*this)(i, j) -= o(i, j); return *this; }
  vnl_matrix_fixed& operator*=(const T& s) { for (unsigned i = 0; i < R * C; ++i) d_[i] *= s; return *this; }
  vnl_matrix_fixed& operator/=(const T& s) { for (unsigned i = 0; i < R * C; ++i) d_[i] /= s; return *this; }
  bool operator==(const vnl_matrix_fixed& o) const { for (unsigned i = 0; i < R * C; ++i) if (d_[i] != o.d_[i]) return false; return true; }
  bool operator!=(const vnl_matrix_fixed& o) const { return !(*this == o); }
  vnl_vector_fixed<T, R> get_column(unsigned j) const { vnl_vector_fixed<T, R> v; for (unsigned i = 0; i < R; ++i) v[i] = (*this)(i, j); return v; }
  vnl_vector_fixed<T, C> get_row(unsigned i) const { vnl_vector_fixed<T, C> v; for (unsigned j = 0; j < C; ++j) v[j] = (*this)(i, j); return v; }
  vnl_matrix_fixed& set_column(unsigned j, const vnl_vector_fixed<T, R>& v) { for (unsigned i = 0; i < R; ++i) (*this)(i, j) = v[i]; return *this; }
  vnl_matrix_fixed& set_column(unsigned j, const vnl_vector<T>& v) { for (unsigned i = 0; i < R; ++i) (*this)(i, j) = v[i]; return *this; }
  vnl_matrix_fixed& set_column(unsigned j, const T* v) { for (unsigned i = 0; i < R; ++i) (*this)(i, j) = v[i]; return *this; }
  vnl_matrix_fixed& set_row(unsigned i, const vnl_vector_fixed<T, C>& v) { for (unsigned j = 0; j < C; ++j) (*this)(i, j) = v[j]; return *this; }
  vnl_matrix_fixed& set_row(unsigned i, const vnl_vector<T>& v) { for (unsigned j = 0; j < C; ++j) (*this)(i, j) = v[j]; return *this; }
  vnl_matrix_fixed& set_row(unsigned i, const T* v) { for (unsigned j = 0; j < C; ++j) (*this)(i, j) = v[j]; return *this; }
  vnl_matrix<T> extract(unsigned r, unsigned c, unsigned top = 0, unsigned left = 0) const {
    vnl_matrix<T> m(r, c); for (unsigned i = 0; i < r; ++i) for (unsigned j = 0; j < c; ++j) m(i, j) = (*this)(top + i, left + j); return m;
  }
  vnl_matrix_fixed& update(const vnl_matrix<T>& m, unsigned top = 0, unsigned left = 0) {
    for (unsigned i = 0; i < m.rows(); ++i) for (unsigned j = 0; j < m.cols(); ++j) (*this)(top + i, left + j) = m(i, j); return *this;
  }
  template <unsigned r2, unsigned c2> vnl_matrix_fixed& update(const vnl_matrix_fixed<T, r2, c2>& m, unsigned top = 0, unsigned left = 0) {
    for (unsigned i = 0; i < r2; ++i) for (unsigned j = 0; j < c2; ++j) (*this)(top + i, left + j) = m(i, j); return *this;
  }
  T frobenius_norm() const { T s = 0; for (unsigned i = 0; i < R * C; ++i) s += d_[i] * d_[i]; return std::sqrt(s); }
  T fro_norm() const { return frobenius_norm(); }
 private:
  T d_[R * C];
};
template <class T, unsigned R, unsigned C> inline vnl_matrix_fixed<T, R, C> operator+(const vnl_matrix_fixed<T, R, C>& a, const vnl_matrix_fixed<T, R, C>& b) { vnl_matrix_fixed<T, R, C> r(a); r += b; return r; }
template <class T, unsigned R, unsigned C> inline vnl_matrix_fixed<T, R, C> operator-(const vnl_matrix_fixed<T, R, C>& a, const vnl_matrix_fixed<T, R, C>& b) { vnl_matrix_fixed<T, R, C> r(a); r -= b; return r; }
template <class T, unsigned R, unsigned C> inline vnl_matrix_fixed<T, R, C> operator*(const vnl_matrix_fixed<T, R, C>& a, const T& s) { vnl_matrix_fixed<T, R, C> r(a); r *= s; return r; }
template <class T, unsigned R, unsigned C> inline vnl_matrix_fixed<T, R, C> operator*(const T& s, const vnl_matrix_fixed<T, R, C>& a) { vnl_matrix_fixed<T, R, C> r(a); r *= s; return r; }
template <class T, unsigned R, unsigned C> inline vnl_matrix_fixed<T, R, C> operator/(const vnl_matrix_fixed<T, R, C>& a, const T& s) { vnl_matrix_fixed<T, R, C> r(a); r /= s; return r; }
template <class T, unsigned R, unsigned K, unsigned C>
inline vnl_matrix_fixed<T, R, C> operator*(const vnl_matrix_fixed<T, R, K>& a, const vnl_matrix_fixed<T, K, C>& b) {
  vnl_matrix_fixed<T, R, C> r;
  for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) { T s = 0; for (unsigned k = 0; k < K; ++k) s += a(i, k) * b(k, j); r(i, j) = s; }
  return r;
}
template <class T, unsigned R, unsigned C>
inline vnl_vector_fixed<T, R> operator*(const vnl_matrix_fixed<T, R, C>& a, const vnl_vector_fixed<T, C>& v) {
  vnl_vector_fixed<T, R> r;
  for (unsigned i = 0; i < R; ++i) { T s = 0; for (unsigned k = 0; k < C; ++k) s += a(i, k) * v[k]; r[i] = s; }
  return r;
}
template <class T, unsigned R, unsigned C>
inline vnl_vector<T> operator*(const vnl_matrix_fixed<T, R, C>& a, const vnl_vector<T>& v) {
  vnl_vector<T> r(R);
  for (unsigned i = 0; i < R; ++i) { T s = 0; for (unsigned k = 0; k < C; ++k) s += a(i, k) * v[k]; r[i] = s; }
  return r;
}
template <class T, unsigned R, unsigned C> inline vnl_matrix<T> operator*(const vnl_matrix_fixed<T, R, C>& a, const vnl_matrix<T>& b) { return a.as_matrix() * b; }
template <class T, unsigned R, unsigned C> inline vnl_matrix<T> operator*(const vnl_matrix<T>& a, const vnl_matrix_fixed<T, R, C>& b) { return a * b.as_matrix(); }
template <class T, unsigned n> inline vnl_vector<T> operator*(const vnl_matrix<T>& a, const vnl_vector_fixed<T, n>& v) { return a * v.as_vector(); }
template <class T, unsigned R, unsigned C> inline std::ostream& operator<<(std::ostream& os, const vnl_matrix_fixed<T, R, C>& m) {
  for (unsigned i = 0; i < R; ++i) { for (unsigned j = 0; j < C; ++j) os << (j ? " " : "") << m(i, j); os << '\n'; }
  return os;
}
template <class T, unsigned R, unsigned C> inline std::istream& operator>>(std::istream& is, vnl_matrix_fixed<T, R, C>& m) {
  for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) is >> m(i, j); return is;
}
template <class T, unsigned n, unsigned m> inline vnl_matrix_fixed<T, n, m> outer_product(const vnl_vector_fixed<T, n>& a, const vnl_vector_fixed<T, m>& b) {
  vnl_matrix_fixed<T, n, m> r; for (unsigned i = 0; i < n; ++i) for (unsigned j = 0; j < m; ++j) r(i, j) = a[i] * b[j]; return r;
}

template <class T, unsigned R, unsigned C> inline T dot_product(const vnl_matrix_fixed<T, R, C>& a, const vnl_matrix_fixed<T, R, C>& b) {
  T s = 0; for (unsigned i = 0; i < R; ++i) for (unsigned j = 0; j < C; ++j) s += a(i, j) * b(i, j); return s;
}
// vnl_transpose(M) * X : the wrapper of VXL evaluates M^T X; here simply the transposed copy.
template <class T> inline vnl_matrix<T> vnl_transpose(const vnl_matrix<T>& m) { return m.transpose(); }
template <class T, unsigned R, unsigned C> inline vnl_matrix_fixed<T, C, R> vnl_transpose(const vnl_matrix_fixed<T, R, C>& m) { return m.transpose(); }

template <class T> inline T vnl_trace(const vnl_matrix<T>& m) { T s = 0; for (unsigned i = 0; i < m.rows(); ++i) s += m(i, i); return s; }
template <class T, unsigned n> inline T vnl_trace(const vnl_matrix_fixed<T, n, n>& m) { T s = 0; for (unsigned i = 0; i < n; ++i) s += m(i, i); return s; }

typedef vnl_vector_fixed<double, 2> vnl_double_2;
typedef vnl_vector_fixed<double, 3> vnl_double_3;
typedef vnl_matrix_fixed<double, 2, 2> vnl_double_2x2;
typedef vnl_matrix_fixed<double, 3, 3> vnl_double_3x3;

namespace vnl_math {
static const double pi = 3.14159265358979323846;
static const double pi_over_2 = 1.57079632679489661923;
static const double pi_over_4 = 0.78539816339744830962;
static const double sqrt2 = 1.41421356237309504880;
static const double e = 2.7182818284590452354;
inline bool isnan(double x) { return std::isnan(x); }
inline bool isinf(double x) { return std::isinf(x); }
inline bool isfinite(double x) { return std::isfinite(x); }
inline double sqr(double x) { return x * x; }
inline double abs(double x) { return std::fabs(x); }
inline int rnd(double x) { return (int)std::lround(x); }
inline double sgn(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }
}  // namespace vnl_math
inline double vnl_math_sqr(double x) { return x * x; }
inline double vnl_math_abs(double x) { return std::fabs(x); }
inline bool vnl_math_isnan(double x) { return std::isnan(x); }
inline int vnl_math_rnd(double x) { return (int)std::lround(x); }

#endif
