// lrr_svm.cpp — the verification step of the driver (rrl/itkLocationRegistration.cxx:2087-2109, 2508-2528): a trained LIBSVM model
// decides from the eight decision measurements of a candidate whether the registration is accepted.  Host only (liblrr_io.so).
// The reference links the libsvm 2.8x sources it carries (libsvm/svm.cpp) through libsvm/svm_interface.cpp:18-102 (load_model,
// predict); its model file `training_data_all.model` is not in the repository, so the caller supplies one.  Written from the
// published formats and methods: the LIBSVM model text format (header keywords, then one line per support vector
// "coef ... index:value ..."), the one-against-one decision function of C-SVC / nu-SVC (sum of coef K(x, sv) - rho per class
// pair, majority vote), Platt's sigmoid on the decision value and the pairwise coupling of Wu, Lin and Weng (JMLR 2004, second
// method) for class probabilities.  Checked against the libsvm sources of the reference tree (tests/test_svm.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/lrr_io.h"

namespace {
enum { SVM_C_SVC = 0, SVM_NU_SVC, SVM_ONE_CLASS, SVM_EPSILON_SVR, SVM_NU_SVR };
enum { K_LINEAR = 0, K_POLY, K_RBF, K_SIGMOID, K_PRECOMPUTED };
struct Node { int index; double value; };
void seterr(char* err, int errlen, const std::string& m) { if (err && errlen > 0) { std::snprintf(err, (size_t)errlen, "%s", m.c_str()); } }
}  // namespace

struct lrr_svm_model {
  int svm_type = SVM_C_SVC, kernel_type = K_RBF, degree = 3, nr_class = 0, total_sv = 0;
  double gamma = 0, coef0 = 0;
  std::vector<double> rho, probA, probB;
  std::vector<int> label, nr_sv;
  std::vector<std::vector<double> > sv_coef;   // (nr_class - 1) x total_sv
  std::vector<std::vector<Node> > sv;          // sparse vectors, ascending index
};

namespace {
// sparse inner product / squared distance over index-sorted vectors
double sparse_dot(const std::vector<Node>& a, const std::vector<Node>& b) {
  double s = 0; size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    if (a[i].index == b[j].index) { s += a[i].value * b[j].value; ++i; ++j; }
    else if (a[i].index > b[j].index) ++j; else ++i;
  }
  return s;
}
double sparse_dist2(const std::vector<Node>& a, const std::vector<Node>& b) {
  double s = 0; size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    if (a[i].index == b[j].index) { const double d = a[i].value - b[j].value; s += d * d; ++i; ++j; }
    else if (a[i].index > b[j].index) { s += b[j].value * b[j].value; ++j; }
    else { s += a[i].value * a[i].value; ++i; }
  }
  for (; i < a.size(); ++i) s += a[i].value * a[i].value;
  for (; j < b.size(); ++j) s += b[j].value * b[j].value;
  return s;
}
double int_power(double base, int times) {   // square-and-multiply, the library's rounding order
  double t = base, r = 1.0;
  for (int k = times; k > 0; k /= 2) { if (k % 2 == 1) r *= t; t = t * t; }
  return r;
}
double kernel(const lrr_svm_model& m, const std::vector<Node>& x, const std::vector<Node>& y) {
  switch (m.kernel_type) {
    case K_LINEAR: return sparse_dot(x, y);
    case K_POLY: return int_power(m.gamma * sparse_dot(x, y) + m.coef0, m.degree);
    case K_RBF: return std::exp(-m.gamma * sparse_dist2(x, y));
    case K_SIGMOID: return std::tanh(m.gamma * sparse_dot(x, y) + m.coef0);
    default: return 0.0;
  }
}
// decision values of all class pairs (i < j), pair-major
void decision_values(const lrr_svm_model& m, const std::vector<Node>& x, std::vector<double>& dec) {
  const int k = m.nr_class, l = m.total_sv;
  std::vector<double> kv((size_t)l);
  for (int i = 0; i < l; ++i) kv[(size_t)i] = kernel(m, x, m.sv[(size_t)i]);
  if (m.svm_type == SVM_ONE_CLASS || m.svm_type == SVM_EPSILON_SVR || m.svm_type == SVM_NU_SVR) {
    double s = 0; for (int i = 0; i < l; ++i) s += m.sv_coef[0][(size_t)i] * kv[(size_t)i];
    dec.assign(1, s - m.rho[0]); return;
  }
  std::vector<int> start((size_t)k, 0);
  for (int i = 1; i < k; ++i) start[(size_t)i] = start[(size_t)i - 1] + m.nr_sv[(size_t)i - 1];
  dec.clear();
  int p = 0;
  for (int i = 0; i < k; ++i)
    for (int j = i + 1; j < k; ++j) {
      double s = 0;
      const int si = start[(size_t)i], sj = start[(size_t)j], ci = m.nr_sv[(size_t)i], cj = m.nr_sv[(size_t)j];
      const std::vector<double>& c1 = m.sv_coef[(size_t)j - 1]; const std::vector<double>& c2 = m.sv_coef[(size_t)i];
      for (int q = 0; q < ci; ++q) s += c1[(size_t)(si + q)] * kv[(size_t)(si + q)];
      for (int q = 0; q < cj; ++q) s += c2[(size_t)(sj + q)] * kv[(size_t)(sj + q)];
      dec.push_back(s - m.rho[(size_t)p]); ++p;
    }
}
double platt(double dec, double A, double B) {   // numerically safe form of 1 / (1 + exp(A f + B))
  const double f = dec * A + B;
  return f >= 0 ? std::exp(-f) / (1.0 + std::exp(-f)) : 1.0 / (1.0 + std::exp(f));
}
// Wu, Lin, Weng: minimise sum_{i != j} (r_ji p_i - r_ij p_j)^2 over the simplex by the fixed-point sweep of their Algorithm 2;
// stops when max_t |(Qp)_t - p^T Q p| < 0.005 / k or after max(100, k) sweeps
void couple(int k, const std::vector<std::vector<double> >& r, double* p) {
  std::vector<std::vector<double> > Q((size_t)k, std::vector<double>((size_t)k, 0.0));
  std::vector<double> Qp((size_t)k);
  const double eps = 0.005 / k;
  for (int t = 0; t < k; ++t) {
    p[t] = 1.0 / k;
    for (int j = 0; j < t; ++j) { Q[t][t] += r[j][t] * r[j][t]; Q[t][j] = Q[j][t]; }
    for (int j = t + 1; j < k; ++j) { Q[t][t] += r[j][t] * r[j][t]; Q[t][j] = -r[j][t] * r[t][j]; }
  }
  const int max_iter = std::max(100, k);
  for (int iter = 0; iter < max_iter; ++iter) {
    double pQp = 0;
    for (int t = 0; t < k; ++t) { Qp[t] = 0; for (int j = 0; j < k; ++j) Qp[t] += Q[t][j] * p[j]; pQp += p[t] * Qp[t]; }
    double worst = 0;
    for (int t = 0; t < k; ++t) worst = std::max(worst, std::fabs(Qp[t] - pQp));
    if (worst < eps) break;
    for (int t = 0; t < k; ++t) {
      const double diff = (-Qp[t] + pQp) / Q[t][t];
      p[t] += diff;
      pQp = (pQp + diff * (diff * Q[t][t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
      for (int j = 0; j < k; ++j) { Qp[j] = (Qp[j] + diff * Q[t][j]) / (1 + diff); p[j] /= (1 + diff); }
    }
  }
}
}  // namespace

extern "C" int lrr_svm_load(const char* path, lrr_svm_model** out, char* err, int errlen) {
  if (!path || !out) return 1;
  std::ifstream in(path);
  if (!in) { seterr(err, errlen, std::string("cannot open model file ") + path); return 1; }
  lrr_svm_model* m = new lrr_svm_model();
  static const char* types[] = {"c_svc", "nu_svc", "one_class", "epsilon_svr", "nu_svr"};
  static const char* kernels[] = {"linear", "polynomial", "rbf", "sigmoid", "precomputed"};
  std::string key;
  bool have_sv = false;
  auto fail = [&](const std::string& w) { seterr(err, errlen, w); delete m; return 1; };
  while (in >> key) {
    if (key == "svm_type") { std::string v; in >> v; int t = -1; for (int i = 0; i < 5; ++i) if (v == types[i]) t = i; if (t < 0) return fail("unknown svm type " + v); m->svm_type = t; }
    else if (key == "kernel_type") { std::string v; in >> v; int t = -1; for (int i = 0; i < 5; ++i) if (v == kernels[i]) t = i; if (t < 0) return fail("unknown kernel function " + v); m->kernel_type = t; }
    else if (key == "degree") in >> m->degree;
    else if (key == "gamma") in >> m->gamma;
    else if (key == "coef0") in >> m->coef0;
    else if (key == "nr_class") in >> m->nr_class;
    else if (key == "total_sv") in >> m->total_sv;
    else if (key == "rho" || key == "probA" || key == "probB") {
      const int n = m->nr_class * (m->nr_class - 1) / 2; std::vector<double> v((size_t)std::max(n, 0));
      for (double& x : v) in >> x;
      (key == "rho" ? m->rho : (key == "probA" ? m->probA : m->probB)) = v;
    }
    else if (key == "label" || key == "nr_sv") { std::vector<int> v((size_t)std::max(m->nr_class, 0)); for (int& x : v) in >> x; (key == "label" ? m->label : m->nr_sv) = v; }
    else if (key == "SV") { std::string rest; std::getline(in, rest); have_sv = true; break; }
    else return fail("unknown text in model file: [" + key + "]");
    if (!in) return fail("truncated model header at " + key);
  }
  if (!have_sv || m->nr_class < 1 || m->total_sv < 0) return fail("model file without an SV section");
  if (m->kernel_type == K_PRECOMPUTED) return fail("precomputed kernels are not supported");
  const int ncoef = std::max(m->nr_class - 1, 1);
  if (m->rho.empty()) return fail("model file without rho");
  const bool classifier = m->svm_type == SVM_C_SVC || m->svm_type == SVM_NU_SVC;
  if (classifier && ((int)m->label.size() != m->nr_class || (int)m->nr_sv.size() != m->nr_class)) return fail("classification model without label / nr_sv");
  m->sv_coef.assign((size_t)ncoef, std::vector<double>((size_t)m->total_sv, 0.0));
  m->sv.resize((size_t)m->total_sv);
  std::string line;
  for (int i = 0; i < m->total_sv; ++i) {
    if (!std::getline(in, line)) return fail("fewer support vectors than total_sv");
    std::istringstream ls(line);
    for (int k = 0; k < ncoef; ++k) if (!(ls >> m->sv_coef[(size_t)k][(size_t)i])) return fail("bad coefficient in an SV line");
    std::string tok;
    while (ls >> tok) {
      const size_t c = tok.find(':');
      if (c == std::string::npos) return fail("bad index:value pair in an SV line");
      Node n; n.index = std::atoi(tok.substr(0, c).c_str()); n.value = std::strtod(tok.c_str() + c + 1, nullptr);
      m->sv[(size_t)i].push_back(n);
    }
  }
  if (classifier) { int s = 0; for (int v : m->nr_sv) s += v; if (s != m->total_sv) return fail("nr_sv does not add up to total_sv"); }
  *out = m;
  return 0;
}
extern "C" void lrr_svm_free(lrr_svm_model* m) { delete m; }
extern "C" int lrr_svm_nr_class(const lrr_svm_model* m) { return m ? m->nr_class : 0; }
extern "C" int lrr_svm_has_probability(const lrr_svm_model* m) {   // svm_check_probability_model
  if (!m) return 0;
  const bool classifier = m->svm_type == SVM_C_SVC || m->svm_type == SVM_NU_SVC;
  return (classifier && !m->probA.empty() && !m->probB.empty()) ? 1 : 0;
}
extern "C" int lrr_svm_labels(const lrr_svm_model* m, int32_t* labels) { if (!m || !labels) return 1; for (size_t i = 0; i < m->label.size(); ++i) labels[i] = m->label[i]; return 0; }

// predict() of libsvm/svm_interface.cpp:18-76: every component scaled to [-1, 1] with the given ranges (:47-52), then
// svm_predict_probability (classification model with probability information, when asked) or svm_predict.
extern "C" int lrr_svm_predict(const lrr_svm_model* m, const double* x, int size, const double* feature_min, const double* feature_max, int index_base,
                               int predict_probability, double* label, double* prob) {
  if (!m || !x || size < 0 || !label) return 1;
  std::vector<Node> v((size_t)size);
  for (int i = 0; i < size; ++i) {
    v[(size_t)i].index = index_base + i;
    v[(size_t)i].value = (feature_min && feature_max) ? -1.0 + 2.0 * (x[i] - feature_min[i]) / (feature_max[i] - feature_min[i]) : x[i];
  }
  std::vector<double> dec; decision_values(*m, v, dec);
  const int k = m->nr_class;
  if (m->svm_type == SVM_ONE_CLASS) { *label = dec[0] > 0 ? 1 : -1; return 0; }
  if (m->svm_type == SVM_EPSILON_SVR || m->svm_type == SVM_NU_SVR) { *label = dec[0]; return 0; }
  if (predict_probability && lrr_svm_has_probability(m) && prob) {
    const double min_prob = 1e-7;
    std::vector<std::vector<double> > r((size_t)k, std::vector<double>((size_t)k, 0.0));
    int p = 0;
    for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j) {
      r[i][j] = std::min(std::max(platt(dec[(size_t)p], m->probA[(size_t)p], m->probB[(size_t)p]), min_prob), 1 - min_prob);
      r[j][i] = 1 - r[i][j]; ++p;
    }
    couple(k, r, prob);
    int best = 0; for (int i = 1; i < k; ++i) if (prob[i] > prob[best]) best = i;
    *label = m->label[(size_t)best];
    return 0;
  }
  std::vector<int> vote((size_t)k, 0);
  int p = 0;
  for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j) { if (dec[(size_t)p++] > 0) ++vote[(size_t)i]; else ++vote[(size_t)j]; }
  int best = 0; for (int i = 1; i < k; ++i) if (vote[(size_t)i] > vote[(size_t)best]) best = i;
  *label = m->label[(size_t)best];
  return 0;
}

// The driver's acceptance test of one candidate (rrl/itkLocationRegistration.cxx:2097-2109, 2512-2528): the eight decision
// measurements, node indices 0..7 as the driver sets them (:2099; LIBSVM data files count from 1 — kept as the driver has it), the
// scaling ranges copied from its svm-scale output (:2108-2109), verified = probability_estimates[0] > 0.5.
extern "C" int lrr_svm_verify(const lrr_svm_model* m, const double* measurements8, int* verified, double* probability_good, double* label) {
  static const double fmin[8] = {0.823246, 0.17355, 13.5656, 14.8583, 0.776814, 0.124, 10.3695, 13.049};
  static const double fmax[8] = {16.1025, 139.276, 62.7776, 63.5239, 1.64664, 60.897, 59.8739, 60.8141};
  if (!m || !measurements8 || !verified) return 1;
  if (m->nr_class != 2 || !lrr_svm_has_probability(m)) return 2;   // load_model( ..., predict_probability = 1 ) exits on such a model (:82-86)
  double prob[2] = {0, 0}, lab = 0;
  const int rc = lrr_svm_predict(m, measurements8, 8, fmin, fmax, 0, 1, &lab, prob);
  if (rc) return rc;
  *verified = prob[0] > 0.5 ? 1 : 0;
  if (probability_good) *probability_good = prob[0];
  if (label) *label = lab;
  return 0;
}
