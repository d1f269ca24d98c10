// lrr_ref_impl.hpp — generic adapter from RefEst (lrr_ref_iface.hpp) onto the reference's estimator class templates.
// TEST INFRASTRUCTURE ONLY.  Included by lrr_ref_cdcl.cpp / lrr_ref_rrl.cpp AFTER the reference headers (which are
// included with `protected`/`private` opened so that matches_, forward_/backward_ and the ramp state can be read —
// the reference sources themselves are not modified).
#ifndef LRR_REF_IMPL_HPP_
#define LRR_REF_IMPL_HPP_
#include <cstring>
#include <unordered_map>
#include "lrr_ref_iface.hpp"

typedef vnl_vector_fixed<double, 3> ref_vec3;
typedef vnl_matrix_fixed<double, 3, 3> ref_mat3;

template <unsigned dof> struct RefTransOps;
template <> struct RefTransOps<12> {
  typedef cdcl_trans<3, 12> base;
  static base::sptr make(const RefTrans& t) {
    base::sptr p = new cdcl_trans_affine<3, 12>(ref_mat3(t.A), ref_vec3(t.t), ref_vec3(t.c));
    if (!t.cov.empty()) p->set_covariance(vnl_matrix_fixed<double, 12, 12>(t.cov.data()));
    return p;
  }
  static int kind() { return LRR_TRANS_AFFINE; }
};
template <> struct RefTransOps<6> {
  typedef cdcl_trans<3, 6> base;
  static base::sptr make(const RefTrans& t) {
    base::sptr p = new cdcl_trans_rigid3d(ref_mat3(t.A), ref_vec3(t.t), ref_vec3(t.c));
    if (!t.cov.empty()) p->set_covariance(vnl_matrix_fixed<double, 6, 6>(t.cov.data()));
    return p;
  }
  static int kind() { return LRR_TRANS_RIGID3D; }
};

// Policy P supplies: feature_type, fwd_type, sym_type, make_feature(const RefFeat&), after_construct(...) (Voronoi maps), objcov.
template <class P, unsigned dof>
struct RefEstImpl : RefEst {
  typedef typename P::feature_type feature_type;
  typedef typename feature_type::sptr feature_sptr;
  typedef typename P::template fwd<dof>::type fwd_type;
  typedef typename P::template sym<dof>::type sym_type;
  typedef typename cdcl_trans<3, dof>::sptr trans_sptr;

  bool is_sym;
  std::vector<RefFeat> moving_base, fixed_base;
  std::vector<feature_sptr> moving_all, fixed_all;
  std::vector<int32_t> mi_idx, fi_idx;
  std::unordered_map<const void*, int> id_;   // feature object -> index into its *_all set
  vbl_smart_ptr<fwd_type> fwd;
  vbl_smart_ptr<sym_type> sym;
  typename P::extra extra;

  std::vector<feature_sptr> make_set(const std::vector<RefFeat>& base, const std::vector<int32_t>* idx) {
    std::vector<feature_sptr> v;
    size_t n = idx ? idx->size() : base.size();
    v.reserve(n);
    for (size_t i = 0; i < n; ++i) {
      int k = idx ? (*idx)[i] : (int)i;
      feature_sptr f = P::make_feature(base[k]);
      id_[(const void*)f.as_pointer()] = k;
      v.push_back(f);
    }
    return v;
  }
  RefTrans identity() const { RefTrans t; t.kind = RefTransOps<dof>::kind(); std::memset(t.A, 0, sizeof t.A); t.A[0] = t.A[4] = t.A[8] = 1; std::memset(t.t, 0, sizeof t.t); std::memset(t.c, 0, sizeof t.c); return t; }

  RefEstImpl(bool symmetric, const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed)
      : is_sym(symmetric), moving_base(moving), fixed_base(fixed) {
    moving_all = make_set(moving_base, 0); fixed_all = make_set(fixed_base, 0);
    trans_sptr t = RefTransOps<dof>::make(identity());
    std::vector<feature_sptr> mi = make_set(moving_base, 0);
    if (!is_sym) {
      fwd = new fwd_type(mi, fixed_all, t);
      P::after_construct_fwd(extra, *fwd, fixed_all);
    } else {
      std::vector<feature_sptr> fi = make_set(fixed_base, 0);
      sym = new sym_type(moving_all, fixed_all, t, t->inverse(), mi, fi);   // ctor #3
      P::after_construct_sym(extra, *sym, moving_all, fixed_all);
    }
  }
  fwd_type* dir(int d) { if (is_sym) return d == 0 ? sym->forward_.as_pointer() : sym->backward_.as_pointer(); return d == 0 ? fwd.as_pointer() : 0; }

  void set_subsets(const int32_t* mi, int64_t nm, const int32_t* fi, int64_t nf) override {
    mi_idx.assign(mi, mi + nm); if (fi) fi_idx.assign(fi, fi + nf); else fi_idx.clear();
    std::vector<feature_sptr> mv = make_set(moving_base, &mi_idx);
    if (!is_sym) { std::vector<std::vector<feature_sptr> > multi(1, mv); fwd->set_moving_data(multi); }
    else { std::vector<feature_sptr> fv = make_set(fixed_base, fi_idx.empty() ? 0 : &fi_idx); sym->set_features(mv, fv); }
  }
  void set_transform(const RefTrans& t) override {
    trans_sptr p = RefTransOps<dof>::make(t);
    if (is_sym) sym->set_transform(p); else fwd->set_transform(p);
  }
  void set_flags(int normalize, int wbs) override {
    for (int d = 0; d < 2; ++d) if (fwd_type* f = dir(d)) { f->normalize_matches_ = normalize != 0; f->weight_by_strength(wbs != 0); }
    if (is_sym) sym->normalize_matches_ = normalize != 0;
  }
  bool set_number_points(int d, int64_t n, int* at_max) override {
    fwd_type* f = dir(d); if (!f) return false;
    bool r = f->set_number_points((unsigned)n); if (at_max) *at_max = r; return true;
  }
  void find_closest() override { if (is_sym) sym->find_closest_euclidean(); else fwd->find_closest_euclidean(); }
  bool one_iteration(unsigned it) override { return is_sym ? sym->one_iteration(0, 0, it) : fwd->one_iteration(0, 0, it); }
  bool get_transform(int d, RefTrans& t) override {
    fwd_type* f = dir(d); if (!f) return false;
    trans_sptr p = f->get_transform();
    t.kind = RefTransOps<dof>::kind();
    ref_mat3 A = p->get_A(); ref_vec3 tr = p->get_translation();
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) t.A[3 * i + j] = A(i, j); t.t[i] = tr[i]; t.c[i] = p->center_moving_[i]; }
    const vnl_matrix_fixed<double, dof, dof>& c = p->get_covariance();
    t.cov.assign(c.data_block(), c.data_block() + dof * dof);
    return true;
  }
  int lookup(const void* p) const { auto it = id_.find(p); return it == id_.end() ? -1 : it->second; }
  int64_t get_matches(int d, int64_t cap, int32_t* from, int32_t* to, double* w) override {
    fwd_type* f = dir(d); if (!f) return -1;
    int64_t m = (int64_t)f->matches_.size();
    for (int64_t i = 0; i < m && i < cap; ++i) {
      // match i was formed from moving_[res_][i] (the rigid path replaces from_ by a mapped copy, cdcl :455-484)
      if (from) from[i] = lookup((const void*)f->moving_[f->res_][i].as_pointer());
      if (to) to[i] = lookup((const void*)f->matches_[i]->to_[0].as_pointer());
      if (w) w[i] = f->matches_[i]->w_[0];
    }
    return m;
  }
  bool metric(int which, int d, double* out) override {
    fwd_type* f = dir(d); if (!f) return false;
    switch (which) {
      case LRR_METRIC_RMS: *out = f->RMS_error(); return true;
      case LRR_METRIC_WEIGHTED: *out = f->weighted_error(); return true;
      case LRR_METRIC_MEDIAN: *out = f->median_error(); return true;
      case LRR_METRIC_SHEET_ANGLES: *out = f->sheet_angles(); return true;
      case LRR_METRIC_TUBE_ANGLES: *out = f->tube_angles(); return true;
    }
    return false;
  }
  bool estimate_ls_backward(int ep) override { if (!is_sym) return false; sym->estimate_LS_backward(ep != 0); return true; }
  bool transfer_cov(int d, double* S9) override {
    fwd_type* f = dir(d); if (!f) return false;
    f->compute_transfer_error_covariance();
    const ref_mat3& S = f->get_transform()->get_covarianceJ();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S9[3 * i + j] = S(i, j);
    return true;
  }
  void get_state(lrr_est_state* s) override {
    std::memset(s, 0, sizeof *s);
    for (int d = 0; d < 2; ++d) if (fwd_type* f = dir(d)) { s->number_points[d] = f->number_points_; s->scale_estimated[d] = f->scale_estimated_; }
    if (is_sym) {
      s->number_matches = sym->number_matches_; s->finest_level = sym->finest_level_; s->oscillation_count = (int)sym->oscillation_count_;
      s->error_difference = sym->error_difference_; s->weighted_error = sym->weighted_error_;
    }
  }
  int oscillated() override { return is_sym ? (int)sym->oscillated() : 0; }
  bool objcov_eval(int d, const double* x, double* f, double* g) override {
    fwd_type* fe = dir(d); if (!fe) return false;
    return P::template objcov<dof>(*fe, x, f, g);
  }
  bool objpar_eval(int d, const double* x, double* f, double* g) override {
    fwd_type* fe = dir(d); if (!fe) return false;
    return P::template objpar<dof>(*fe, x, f, g);
  }
};
#endif
