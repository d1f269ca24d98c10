// lrr_ref_cdc.cpp — compiles the reference's full CDC estimator (cdcl/cdcl_estimation.txx: compute_weights, the E-step with
// covariance-driven radius matching) UNMODIFIED from /root/reference.  TEST INFRASTRUCTURE ONLY.
// Reference sources pulled in by #include (never copied): cdcl/cdcl_trans.txx, cdcl_trans_affine.txx, cdcl_utils.txx,
// cdcl_estimation_ICP.txx, cdcl_estimation.txx, cdcl_obj_fun_cov.txx, cdcl_obj_fun_par.txx (+ their headers).
// The estimator is built with its constructor #2 (one resolution level, no subsampling: cdcl_estimation.txx:55-61,
// cdcl_estimation_ICP.txx:93-108) on plain cdcl_feature<3> sets; compute_weights( false ) is the function under test.
// vcl_random_shuffle goes through the shim's hook (identity order by default: Q4), rsdl_kd_tree::points_in_radius of the shim
// returns ascending indices, so the cap of 100 keeps the lowest indices.
#include "vxl_shim/vcl_shim_all.h"
#include "vxl_shim/vnl/algo/vnl_shim_algo.h"
#include "vxl_shim/rsdl/rsdl_kd_tree.h"
#include "vxl_shim/rrel/rrel_muset_obj.h"
#include <cstring>
#include <unordered_map>
#define protected public
#define private public
#include <cdcl/cdcl_trans.txx>
#include <cdcl/cdcl_trans_affine.txx>
#include <cdcl/cdcl_utils.txx>
#include <cdcl/cdcl_estimation_ICP.txx>
#include <cdcl/cdcl_estimation.txx>
#include <cdcl/cdcl_obj_fun_cov.txx>
#include <cdcl/cdcl_obj_fun_par.txx>
#undef protected
#undef private
#include "lrr_ref_iface.hpp"

namespace {
typedef cdcl_feature<3> feat_t;
typedef cdcl_match<3> match_t;
typedef cdcl_estimation<3, 12> est_t;
typedef vnl_vector_fixed<double, 3> vec3;
typedef vnl_matrix_fixed<double, 3, 3> mat3;

struct RefCdcImpl : RefCdc {
  std::vector<feat_t::sptr> moving, fixed;
  std::unordered_map<const void*, int> id_;
  vbl_smart_ptr<est_t> est;
  static feat_t::sptr mk(const RefFeat& f) { return new feat_t(vec3(f.loc), f.strength, mat3(f.cov)); }
  RefCdcImpl(const std::vector<RefFeat>& m, const int32_t* mi, int64_t nm, const std::vector<RefFeat>& f) {
    const size_t n = mi ? (size_t)nm : m.size();
    for (size_t i = 0; i < n; ++i) { const int k = mi ? mi[i] : (int)i; feat_t::sptr p = mk(m[k]); id_[(const void*)p.as_pointer()] = k; moving.push_back(p); }
    for (size_t i = 0; i < f.size(); ++i) { feat_t::sptr p = mk(f[i]); id_[(const void*)p.as_pointer()] = (int)i; fixed.push_back(p); }
  }
  void set_transform(const RefTrans& t) override {
    vnl_matrix_fixed<double, 3, 3> A(t.A); vec3 tr(t.t), c(t.c);
    cdcl_trans<3, 12>::sptr tp = new cdcl_trans_affine<3, 12>(A, tr, c);
    if (!t.cov.empty()) tp->set_covariance(vnl_matrix_fixed<double, 12, 12>(t.cov.data()));
    std::vector<std::vector<feat_t::sptr> > mm(1, moving), ff(1, fixed);
    std::vector<double> spacing(1, 1.0);
    est = new est_t(mm, ff, tp, spacing);
  }
  bool compute_weights(lrr_cdc_summary* out) override {
    if (!est) return false;
    est->compute_weights(false);
    if (out) {
      out->rows = (int64_t)est->matches_.size(); out->entries = 0; out->min_matches = 0; out->max_matches = 0; out->avg_largest_eval = 0;
      int mn = 0x7fffffff, mx = 0;
      for (size_t i = 0; i < est->matches_.size(); ++i) { const int k = (int)est->matches_[i]->to_.size(); out->entries += k; mn = std::min(mn, k); mx = std::max(mx, k); }
      if (out->rows) { out->min_matches = mn; out->max_matches = mx; }
      // (avg_largest_eval is a local of compute_weights: recomputed here the way :375 and :480 do)
      double s = 0;
      for (size_t i = 0; i < est->moving_[est->res_].size(); ++i) {
        vnl_matrix_fixed<double, 3, 12> const& Jth = est->trans_->jacobian_wrt_par(est->moving_[est->res_][i]->location_);
        s += vcl_sqrt(cdcl_largest_evalue(est->transfer_covar(Jth)));
      }
      out->avg_largest_eval = s / (double)est->matches_.size();
    }
    return true;
  }
  int64_t rows() const override { return est ? (int64_t)est->matches_.size() : 0; }
  int64_t entries() const override { int64_t e = 0; if (est) for (size_t i = 0; i < est->matches_.size(); ++i) e += (int64_t)est->matches_[i]->to_.size(); return e; }
  void get_matches(int32_t* row_from, int64_t* row_start, int32_t* to_idx, double* w) override {
    int64_t k = 0;
    for (size_t i = 0; i < est->matches_.size(); ++i) {
      if (row_from) row_from[i] = id_[(const void*)est->matches_[i]->from_.as_pointer()];
      if (row_start) row_start[i] = k;
      for (size_t j = 0; j < est->matches_[i]->to_.size(); ++j) {
        if (to_idx) to_idx[k] = id_[(const void*)est->matches_[i]->to_[j].as_pointer()];
        if (w) w[k] = est->matches_[i]->w_[j];
        ++k;
      }
    }
    if (row_start) row_start[est->matches_.size()] = k;
  }
  bool objcov(const double* x, double* f, double* g) override {
    if (!est || est->matches_.empty()) return false;
    cdcl_trans<3, 12>::sptr t = est->trans_->clone();
    cdcl_obj_fun_cov<3, 12> obj(est->matches_, t);
    vnl_vector<double> xv(x, 78), gv(78, 0.0); double fv = 0;
    obj.compute(xv, &fv, &gv);
    if (f) *f = fv; if (g) for (unsigned k = 0; k < 78; ++k) g[k] = gv[k];
    return true;
  }
  bool objpar(const double* x, double* f, double* g) override {
    if (!est || est->matches_.empty()) return false;
    cdcl_trans<3, 12>::sptr t = est->trans_->clone();
    cdcl_obj_fun_par<3, 12> obj(est->matches_, t);
    vnl_vector<double> xv(x, 12), gv(12, 0.0); double fv = 0;
    obj.compute(xv, &fv, &gv);
    if (f) *f = fv; if (g) for (unsigned k = 0; k < 12; ++k) g[k] = gv[k];
    return true;
  }
  // cdcl_normalize_matches over the many-to-many set (cdcl_estimation_ICP.txx:751-832), undone again at once
  void normalisation(double* norm8) override {
    vec3 cm, cf; double rm = 0, rf = 0;
    est->cdcl_normalize_matches(cm, rm, cf, rf);
    est->cdcl_unnormalize_matches(cm, rm, cf, rf);
    for (int d = 0; d < 3; ++d) { norm8[d] = cm[d]; norm8[4 + d] = cf[d]; }
    norm8[3] = rm; norm8[7] = rf;
  }
};
}  // namespace

RefCdc* lrr_ref_make_cdc(const std::vector<RefFeat>& moving, const int32_t* mi, int64_t nm, const std::vector<RefFeat>& fixed) { return new RefCdcImpl(moving, mi, nm, fixed); }
