// lrr_ref_icp2d.cpp — compiles the reference's 2-D ICP estimator (config 1 of BASELINE.json: cdcl/examples, method 4)
// UNMODIFIED from /root/reference.  TEST INFRASTRUCTURE ONLY.
// Reference sources pulled in by #include (never copied): cdcl/cdcl_trans.txx, cdcl_trans_similarity2d.cxx, cdcl_utils.txx,
// cdcl_estimation_ICP.txx (+ their headers).  The estimator is driven the way cdcl/examples/estimate.cxx:99-101,113-132 does:
// ctor #1 (one resolution for dim == 2), set_transform, one_iteration until |dRMS| < 1e-4 or 30 iterations.
#include "vxl_shim/vcl_shim_all.h"
#include "vxl_shim/vnl/algo/vnl_shim_algo.h"
#include "vxl_shim/rsdl/rsdl_kd_tree.h"
#include "vxl_shim/rrel/rrel_muset_obj.h"
#include <cstring>
#include <unordered_map>
#define protected public
#define private public
#include <cdcl/cdcl_trans.txx>
#include <cdcl/cdcl_trans_similarity2d.cxx>
#include <cdcl/cdcl_utils.txx>
#include <cdcl/cdcl_estimation_ICP.txx>
#undef protected
#undef private
#include "lrr_ref_iface.hpp"

template class cdcl_trans<2, 4>;
template class cdcl_estimation_ICP<2, 4>;

namespace {
typedef cdcl_feature<2> feat_t;
typedef cdcl_estimation_ICP<2, 4> est_t;
typedef vnl_vector_fixed<double, 2> vec2;
typedef vnl_matrix_fixed<double, 2, 2> mat2;

struct RefIcp2d : RefEst {
  std::vector<RefFeat> moving_base, fixed_base;
  std::vector<feat_t::sptr> moving, fixed;
  std::unordered_map<const void*, int> id_;
  vbl_smart_ptr<est_t> est;

  static feat_t::sptr make_feature(const RefFeat& f) {
    vec2 loc; loc[0] = f.loc[0]; loc[1] = f.loc[1];
    mat2 C; C(0, 0) = f.cov[0]; C(0, 1) = f.cov[1]; C(1, 0) = f.cov[3]; C(1, 1) = f.cov[4];
    return new feat_t(loc, f.strength, C);
  }
  std::vector<feat_t::sptr> make_set(const std::vector<RefFeat>& base, const int32_t* idx, int64_t n) {
    std::vector<feat_t::sptr> v;
    const size_t m = idx ? (size_t)n : base.size();
    for (size_t i = 0; i < m; ++i) {
      const int k = idx ? idx[i] : (int)i;
      feat_t::sptr f = make_feature(base[k]);
      id_[(const void*)f.as_pointer()] = k;
      v.push_back(f);
    }
    return v;
  }
  static cdcl_trans<2, 4>::sptr make_trans(const RefTrans& t) {
    // (A, t) -> (s, ang, t): A = s R  (cdcl_trans_similarity2d.cxx:33-47)
    const double s = std::sqrt(t.A[0] * t.A[0] + t.A[3] * t.A[3]);
    const double ang = std::atan2(t.A[3], t.A[0]);
    vec2 tr; tr[0] = t.t[0]; tr[1] = t.t[1];
    cdcl_trans<2, 4>::sptr p = new cdcl_trans_similarity2d(s, ang, tr);
    if (!t.cov.empty()) p->set_covariance(vnl_matrix_fixed<double, 4, 4>(t.cov.data()));
    return p;
  }
  RefIcp2d(const std::vector<RefFeat>& m, const std::vector<RefFeat>& f) : moving_base(m), fixed_base(f) {
    moving = make_set(moving_base, 0, 0); fixed = make_set(fixed_base, 0, 0);
    RefTrans id; id.kind = LRR_TRANS_SIMILARITY2D; std::memset(id.A, 0, sizeof id.A); id.A[0] = id.A[4] = id.A[8] = 1;
    std::memset(id.t, 0, sizeof id.t); std::memset(id.c, 0, sizeof id.c);
    est = new est_t(moving, fixed, make_trans(id));
  }
  void set_subsets(const int32_t* mi, int64_t nm, const int32_t*, int64_t) override {
    moving = make_set(moving_base, mi, nm);
    std::vector<std::vector<feat_t::sptr> > multi(1, moving);
    est->set_moving_data(multi);
  }
  void set_transform(const RefTrans& t) override { est->set_transform(make_trans(t)); }
  void set_flags(int normalize, int wbs) override { est->normalize_matches_ = normalize != 0; est->weight_by_strength(wbs != 0); }
  bool set_number_points(int, int64_t, int*) override { return false; }   // cdcl_estimation_ICP has no point ramp
  void find_closest() override { est->find_closest_euclidean(); }
  bool one_iteration(unsigned it) override { return est->one_iteration(0, 0, it); }
  bool get_transform(int d, RefTrans& t) override {
    if (d != 0) return false;
    cdcl_trans_similarity2d* p = dynamic_cast<cdcl_trans_similarity2d*>(est->get_transform().as_pointer());
    if (!p) return false;
    t.kind = LRR_TRANS_SIMILARITY2D;
    std::memset(t.A, 0, sizeof t.A); std::memset(t.t, 0, sizeof t.t); std::memset(t.c, 0, sizeof t.c);
    const mat2 sR = p->jacobian_wrt_loc();   // s_ * R_
    t.A[0] = sR(0, 0); t.A[1] = sR(0, 1); t.A[3] = sR(1, 0); t.A[4] = sR(1, 1); t.A[8] = 1.0;
    t.t[0] = p->t_[0]; t.t[1] = p->t_[1];
    const vnl_matrix_fixed<double, 4, 4>& c = p->get_covariance();
    t.cov.assign(c.data_block(), c.data_block() + 16);
    return true;
  }
  int lookup(const void* p) const { auto it = id_.find(p); return it == id_.end() ? -1 : it->second; }
  int64_t get_matches(int d, int64_t cap, int32_t* from, int32_t* to, double* w) override {
    if (d != 0) return -1;
    const int64_t m = (int64_t)est->matches_.size();
    for (int64_t i = 0; i < m && i < cap; ++i) {
      if (from) from[i] = lookup((const void*)est->matches_[i]->from_.as_pointer());
      if (to) to[i] = lookup((const void*)est->matches_[i]->to_[0].as_pointer());
      if (w) w[i] = est->matches_[i]->w_[0];
    }
    return m;
  }
  bool metric(int which, int d, double* out) override {
    if (d != 0) return false;
    switch (which) {
      case LRR_METRIC_RMS: *out = est->RMS_error(); return true;
      case LRR_METRIC_WEIGHTED: *out = est->weighted_error(); return true;
      case LRR_METRIC_MEDIAN: *out = est->median_error(); return true;
    }
    return false;
  }
  bool estimate_ls_backward(int) override { return false; }
  bool transfer_cov(int d, double* S9) override {
    if (d != 0) return false;
    est->compute_transfer_error_covariance();
    const mat2& S = est->get_transform()->get_covarianceJ();
    std::memset(S9, 0, 72);
    S9[0] = S(0, 0); S9[1] = S(0, 1); S9[3] = S(1, 0); S9[4] = S(1, 1);
    return true;
  }
  void get_state(lrr_est_state* s) override {
    std::memset(s, 0, sizeof *s);
    s->number_points[0] = (int64_t)est->moving_[est->res_].size(); s->scale_estimated[0] = est->scale_estimated_;
  }
  int oscillated() override { return 0; }
  bool objcov_eval(int, const double*, double*, double*) override { return false; }
  bool objpar_eval(int, const double*, double*, double*) override { return false; }
};
}  // namespace

RefEst* lrr_ref_make_icp2d(const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed) { return new RefIcp2d(moving, fixed); }

// error projector the reference derives per use from the fixed feature's covariance (cdcl_estimation_ICP.txx:470-477)
void lrr_ref_projector2d(const double* C3 /* xx, xy, yy */, double* P3 /* xx, xy, yy */) {
  vnl_matrix<double> A(2, 2), V(2, 2); vnl_vector<double> D(2);
  A(0, 0) = C3[0]; A(0, 1) = C3[1]; A(1, 0) = C3[1]; A(1, 1) = C3[2];
  vnl_symmetric_eigensystem_compute(A, V, D);
  vnl_vector<double> e = V.get_column(0);
  P3[0] = e[0] * e[0]; P3[1] = e[0] * e[1]; P3[2] = e[1] * e[1];
}

// normalize (+ unnormalize) of a similarity transform and its covariance (cdcl_trans_similarity2d.cxx:227-305)
void lrr_ref_sim2d_roundtrip(double* s, double* ang, double* t2, double* cov16, double rm, double rf, const double* cm2, const double* cf2, int do_unnormalize) {
  vec2 tr; tr[0] = t2[0]; tr[1] = t2[1];
  cdcl_trans_similarity2d T(*s, *ang, tr);
  if (cov16) T.set_covariance(vnl_matrix_fixed<double, 4, 4>(cov16));
  vec2 cm, cf; cm[0] = cm2[0]; cm[1] = cm2[1]; cf[0] = cf2[0]; cf[1] = cf2[1];
  T.normalize(rm, rf, cm, cf);
  if (do_unnormalize) T.unnormalize(rm, rf, cm, cf);
  *s = T.s_; *ang = T.ang_; t2[0] = T.t_[0]; t2[1] = T.t_[1];
  if (cov16) std::memcpy(cov16, T.get_covariance().data_block(), 128);
}
