// lrr_ref_cdcl.cpp — compiles the reference's cdcl estimators UNMODIFIED from /root/reference.  TEST INFRASTRUCTURE ONLY.
// Reference sources pulled in by #include (never copied): cdcl/cdcl_trans.txx, cdcl_trans_affine.txx, cdcl_trans_rigid3d.cxx,
// cdcl_estimation_ICP_matching_all.txx, cdcl_estimation_symmetric_ICP_matching_all.txx, cdcl_obj_fun_cov.txx, cdcl_obj_fun_par.txx (+ their headers).
#include "vxl_shim/vcl_shim_all.h"
#include "vxl_shim/vnl/algo/vnl_shim_algo.h"
#include "vxl_shim/rsdl/rsdl_kd_tree.h"
#include "vxl_shim/rrel/rrel_muset_obj.h"
#include <unordered_map>
#define protected public
#define private public
#include <cdcl/cdcl_trans.txx>
#include <cdcl/cdcl_trans_affine.txx>
#include <cdcl/cdcl_trans_rigid3d.cxx>
#include <cdcl/cdcl_estimation_ICP_matching_all.txx>
#include <cdcl/cdcl_estimation_symmetric_ICP_matching_all.txx>
#include <cdcl/cdcl_obj_fun_cov.txx>
#include <cdcl/cdcl_obj_fun_par.txx>
#undef protected
#undef private
#include "lrr_ref_impl.hpp"

template class cdcl_trans<3, 12>;
template class cdcl_trans_affine<3, 12>;

struct CdclPolicy {
  typedef cdcl_feature_with_shape<3> feature_type;
  template <unsigned dof> struct fwd { typedef cdcl_estimation_ICP_matching_all<3, dof> type; };
  template <unsigned dof> struct sym { typedef cdcl_estimation_symmetric_ICP_matching_all<3, dof> type; };
  struct extra {};
  static feature_type::sptr make_feature(const RefFeat& f) {
    std::vector<ref_vec3> dirs;
    int nd = f.shape == 2 ? 1 : (f.shape == 1 ? 2 : 3);   // sheet: normal; tube: binormal, normal; corner: three (cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:191-224)
    for (int a = 0; a < nd; ++a) dirs.push_back(ref_vec3(f.dirs[a]));
    return new feature_type(ref_vec3(f.loc), f.strength, ref_mat3(f.cov), ref_mat3(f.P), (feature_type::shape_type)f.shape, dirs);
  }
  template <class F, class V> static void after_construct_fwd(extra&, F&, const V&) {}
  template <class S, class V> static void after_construct_sym(extra&, S&, const V&, const V&) {}
  // cdcl_obj_fun_cov<3,dof>::compute over the estimator's current 1-NN matches and weights (composition of SURVEY §3.4 / A13:
  // the objective takes cdcl_match<dim> over plain cdcl_feature<dim>, so the matches are re-expressed in that type)
  template <unsigned dof, class F> static bool objcov(F& fe, const double* x, double* f, double* g) {
    typedef cdcl_feature<3> bf; typedef cdcl_match<3> bm;
    std::vector<bm::sptr> ms;
    for (size_t i = 0; i < fe.matches_.size(); ++i) {
      bm::sptr m = new bm;
      m->from_ = new bf(fe.matches_[i]->from_->location_, fe.matches_[i]->from_->strength_, fe.matches_[i]->from_->covariance_);
      for (size_t j = 0; j < fe.matches_[i]->to_.size(); ++j) {
        m->to_.push_back(new bf(fe.matches_[i]->to_[j]->location_, fe.matches_[i]->to_[j]->strength_, fe.matches_[i]->to_[j]->covariance_));
        m->w_.push_back(fe.matches_[i]->w_[j]);
      }
      ms.push_back(m);
    }
    typename cdcl_trans<3, dof>::sptr t = fe.get_transform()->clone();
    cdcl_obj_fun_cov<3, dof> obj(ms, t);
    const unsigned n = dof * (dof + 1) / 2;
    vnl_vector<double> xv(x, n), gv(n, 0.0);
    double fv = 0;
    obj.compute(xv, &fv, &gv);
    if (f) *f = fv;
    if (g) for (unsigned k = 0; k < n; ++k) g[k] = gv[k];
    return true;
  }
  // cdcl_obj_fun_par<3,dof>::compute over the same matches (the objective of the CDC M-step, cdcl/cdcl_obj_fun_par.txx:46-144)
  template <unsigned dof, class F> static bool objpar(F& fe, const double* x, double* f, double* g) {
    typedef cdcl_feature<3> bf; typedef cdcl_match<3> bm;
    std::vector<bm::sptr> ms;
    for (size_t i = 0; i < fe.matches_.size(); ++i) {
      bm::sptr m = new bm;
      m->from_ = new bf(fe.matches_[i]->from_->location_, fe.matches_[i]->from_->strength_, fe.matches_[i]->from_->covariance_);
      for (size_t j = 0; j < fe.matches_[i]->to_.size(); ++j) {
        m->to_.push_back(new bf(fe.matches_[i]->to_[j]->location_, fe.matches_[i]->to_[j]->strength_, fe.matches_[i]->to_[j]->covariance_));
        m->w_.push_back(fe.matches_[i]->w_[j]);
      }
      ms.push_back(m);
    }
    typename cdcl_trans<3, dof>::sptr t = fe.get_transform()->clone();
    cdcl_obj_fun_par<3, dof> obj(ms, t);
    vnl_vector<double> xv(x, dof), gv(dof, 0.0);
    double fv = 0;
    obj.compute(xv, &fv, &gv);
    if (f) *f = fv;
    if (g) for (unsigned k = 0; k < dof; ++k) g[k] = gv[k];
    return true;
  }
};

RefEst* lrr_ref_make_cdcl(int kind, int dof, const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed) {
  bool s = kind == LRR_CDCL_SYM;
  if (dof == 12) return new RefEstImpl<CdclPolicy, 12>(s, moving, fixed);
  if (dof == 6) return new RefEstImpl<CdclPolicy, 6>(s, moving, fixed);
  return 0;
}
void lrr_ref_set_shuffle(int mode, unsigned long long seed) { vxl_shim::shuffle_hook().mode = mode; vxl_shim::shuffle_hook().s = seed; }
double lrr_ref_weight_BT(double u) { return weight_BT(u); }
double lrr_ref_rho(double u) { return rho(u); }
double lrr_ref_rho_p(double u) { return rho_p(u); }
double lrr_ref_mad_scale(const double* v, int64_t n) { std::vector<double> a(v, v + n); return rrel_util_median_abs_dev_scale(a.begin(), a.end()); }

static cdcl_trans<3, 12>::sptr mk12(const double* A9, const double* t3, const double* c3, const double* cov) {
  RefTrans t; t.kind = 0; std::memcpy(t.A, A9, 72); std::memcpy(t.t, t3, 24); std::memcpy(t.c, c3, 24); if (cov) t.cov.assign(cov, cov + 144);
  return RefTransOps<12>::make(t);
}
static cdcl_trans<3, 6>::sptr mk6(const double* A9, const double* t3, const double* c3, const double* cov) {
  RefTrans t; t.kind = 1; std::memcpy(t.A, A9, 72); std::memcpy(t.t, t3, 24); std::memcpy(t.c, c3, 24); if (cov) t.cov.assign(cov, cov + 36);
  return RefTransOps<6>::make(t);
}
template <class S> static void rd(const S& p, double* A9, double* t3, double* c3) {
  ref_mat3 A = p->get_A(); ref_vec3 tr = p->get_translation();
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) A9[3 * i + j] = A(i, j); t3[i] = tr[i]; c3[i] = p->center_moving_[i]; }
}
void lrr_ref_trans_roundtrip(int kind, double* A9, double* t3, double* c3, double* cov, double rm, double rf, const double* cm, const double* cf, int do_un) {
  if (kind == LRR_TRANS_AFFINE) {
    cdcl_trans<3, 12>::sptr p = mk12(A9, t3, c3, cov);
    p->normalize(rm, rf, ref_vec3(cm), ref_vec3(cf)); if (do_un) p->unnormalize(rm, rf, ref_vec3(cm), ref_vec3(cf));
    rd(p, A9, t3, c3); if (cov) std::memcpy(cov, p->get_covariance().data_block(), 144 * 8);
  } else {
    cdcl_trans<3, 6>::sptr p = mk6(A9, t3, c3, cov);
    p->normalize(rm, rf, ref_vec3(cm), ref_vec3(cf)); if (do_un) p->unnormalize(rm, rf, ref_vec3(cm), ref_vec3(cf));
    rd(p, A9, t3, c3); if (cov) std::memcpy(cov, p->get_covariance().data_block(), 36 * 8);
  }
}
void lrr_ref_trans_inverse(int kind, double* A9, double* t3, double* c3) {
  if (kind == LRR_TRANS_AFFINE) { cdcl_trans<3, 12>::sptr p = mk12(A9, t3, c3, 0)->inverse(); rd(p, A9, t3, c3); }
  else { cdcl_trans<3, 6>::sptr p = mk6(A9, t3, c3, 0)->inverse(); rd(p, A9, t3, c3); }
}
void lrr_ref_recompose(double* R9, double* t3, const double* dR9, const double* dt3) {
  double z[3] = {0, 0, 0};
  cdcl_trans<3, 6>::sptr g = mk6(R9, t3, z, 0), inc = mk6(dR9, dt3, z, 0);
  g->recompose_increment(inc);
  double c[3]; rd(g, R9, t3, c);
}
