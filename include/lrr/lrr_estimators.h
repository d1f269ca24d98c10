// Header-only C++ host side above the C ABI (include/lrr_gpu.h): the reference's template estimator API
// (same class names, template parameters <dim, dof>, method names, argument meaning and bool/void returns) so that
//     typedef rrl_estimation_symmetric_ICP_matching_all<dim,dof> estimation_symmetric_type;
// in rrl/itkLocationRegistration.h:336-337 can be pointed at lrr_b200::rrl_estimation_symmetric_ICP_matching_all.
// Everything lives in namespace lrr_b200 to coexist with the reference's own classes in one translation unit.
// Mirrors: cdcl/cdcl_feature_with_shape.h:23-89, cdcl/cdcl_feature_ICP.h:23-96, cdcl/cdcl_match.h:18-44,
// cdcl/cdcl_trans.h:21-148, cdcl/cdcl_trans_affine.h, cdcl/cdcl_trans_rigid3d.h, cdcl/cdcl_estimation_abs.h:20-65,
// cdcl/cdcl_estimation_ICP_matching_all.h, cdcl/cdcl_estimation_symmetric_ICP_matching_all.h,
// rrl/rrl_estimation_ICP_matching_all.h, rrl/rrl_estimation_symmetric_ICP_matching_all.h.
// No VXL: vnl fixed vectors/matrices become std::array, vbl_smart_ptr becomes std::shared_ptr.
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <vector>
#include "../lrr_gpu.h"

namespace lrr_b200 {

typedef std::array<double, 3> vec3;
typedef std::array<double, 9> mat3;   // row-major

// cdcl_feature_with_shape<3> / cdcl_feature_ICP<3>: public data members, as in the reference
template <unsigned int dim>
struct cdcl_feature_with_shape {
  static_assert(dim == 3, "the backend is built for 3-D features");
  enum shape_type { CORNER = 0, TUBE = 1, SHEET = 2 };   // cdcl_feature_with_shape.h:32
  typedef std::shared_ptr<cdcl_feature_with_shape> sptr;
  vec3 location_{};
  double strength_ = 0.0;
  mat3 covariance_{};
  mat3 error_projector_{};
  shape_type shape_ = CORNER;
  std::vector<vec3> directions_;   // sheet: normal; tube: binormal, normal; corner: tangent, binormal, normal
};
template <unsigned int dim> using cdcl_feature_ICP = cdcl_feature_with_shape<dim>;

// cdcl_match: one-to-many in the reference, exactly one `to` on this path (cdcl :551)
template <unsigned int dim, class feature_typeT = cdcl_feature_with_shape<dim> >
struct cdcl_match {
  typedef std::shared_ptr<cdcl_match> sptr;
  typename feature_typeT::sptr from_;
  std::vector<typename feature_typeT::sptr> to_;
  std::vector<double> w_;
};

// cdcl_trans<dim,dof> with the two concrete kinds of the path (affine <3,12>, rigid3d <3,6>)
template <unsigned int dim, unsigned int dof>
class cdcl_trans {
 public:
  static_assert(dim == 3 && (dof == 12 || dof == 6), "instantiated for <3,12> and <3,6> (cdcl/Templates)");
  typedef std::shared_ptr<cdcl_trans> sptr;
  cdcl_trans() { A_ = {1, 0, 0, 0, 1, 0, 0, 0, 1}; covar_.assign(dof * dof, 0.0); }
  cdcl_trans(mat3 const& A, vec3 const& t, vec3 const& center_moving) : cdcl_trans() { A_ = A; t_ = t; center_moving_ = center_moving; }
  mat3 const& get_A() const { return A_; }
  vec3 const& get_translation() const { return t_; }
  std::vector<double> const& get_covariance() const { return covar_; }   // dof x dof row-major
  void set_covariance(std::vector<double> const& c) { covar_ = c; }
  mat3 const& get_covarianceJ() const { return covarJ_; }
  void set_covarianceJ(mat3 const& c) { covarJ_ = c; }
  // map_loc: A (x - center_moving_) + t   (cdcl_trans_affine.txx:112-115)
  vec3 map_loc(vec3 const& x) const {
    vec3 y{};
    for (int r = 0; r < 3; ++r) y[r] = A_[3 * r] * (x[0] - center_moving_[0]) + A_[3 * r + 1] * (x[1] - center_moving_[1]) + A_[3 * r + 2] * (x[2] - center_moving_[2]) + t_[r];
    return y;
  }
  mat3 jacobian_wrt_loc() const { return A_; }
  sptr clone() const { return sptr(new cdcl_trans(*this)); }
  void print(FILE* f = stdout) const {
    std::fprintf(f, "A: %g %g %g / %g %g %g / %g %g %g  t: %g %g %g  center: %g %g %g\n", A_[0], A_[1], A_[2], A_[3], A_[4], A_[5], A_[6], A_[7], A_[8],
                 t_[0], t_[1], t_[2], center_moving_[0], center_moving_[1], center_moving_[2]);
  }
  vec3 center_moving_{};   // public in the reference (cdcl_trans.h:147)
  mat3 A_{}; vec3 t_{}; std::vector<double> covar_; mat3 covarJ_{};
};
typedef cdcl_trans<3, 12> cdcl_trans_affine3d;
typedef cdcl_trans<3, 6> cdcl_trans_rigid3d;

// One CUDA device + the uploaded feature sets are shared by the estimators of a process
class device_context {
 public:
  explicit device_context(int device = 0) { check(lrr_ctx_create(device, &ctx_), "lrr_ctx_create"); }
  ~device_context() { if (ctx_) lrr_ctx_destroy(ctx_); }
  device_context(device_context const&) = delete;
  lrr_ctx* get() const { return ctx_; }
  void check(int rc, const char* what) const {
    if (rc != LRR_OK) throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(rc) + "): " + (ctx_ ? lrr_ctx_last_error(ctx_) : "no context; is a CUDA device present? (no CPU fallback)"));
  }
 private:
  lrr_ctx* ctx_ = nullptr;
};

// vector<feature_sptr> -> structure-of-arrays upload (the AoS heap objects of the reference, cdcl_feature.h:17-64)
template <unsigned int dim>
class feature_set {
 public:
  typedef typename cdcl_feature_with_shape<dim>::sptr feature_sptr_type;
  feature_set(device_context& ctx, std::vector<feature_sptr_type> const& f, bool with_covariance) : ctx_(ctx) {
    const size_t n = f.size();
    std::vector<double> xyz(3 * n), P6(6 * n), C6(6 * n), s(n), dirs(9 * n, 0.0);
    std::vector<uint8_t> shape(n);
    static const int iu[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    for (size_t i = 0; i < n; ++i) {
      for (int d = 0; d < 3; ++d) xyz[d * n + i] = f[i]->location_[d];
      for (int k = 0; k < 6; ++k) { P6[k * n + i] = f[i]->error_projector_[3 * iu[k][0] + iu[k][1]]; C6[k * n + i] = f[i]->covariance_[3 * iu[k][0] + iu[k][1]]; }
      s[i] = f[i]->strength_; shape[i] = (uint8_t)f[i]->shape_;
      for (size_t a = 0; a < f[i]->directions_.size() && a < 3; ++a) for (int d = 0; d < 3; ++d) dirs[(3 * a + d) * n + i] = f[i]->directions_[a][d];
    }
    ctx_.check(lrr_features_upload(ctx_.get(), (int64_t)n, xyz.data(), P6.data(), with_covariance ? C6.data() : nullptr, s.data(), shape.data(), dirs.data(), &fs_), "lrr_features_upload");
  }
  ~feature_set() { if (fs_) lrr_features_destroy(fs_); }
  feature_set(feature_set const&) = delete;
  lrr_fset* get() const { return fs_; }
 private:
  device_context& ctx_; lrr_fset* fs_ = nullptr;
};

// cdcl_estimation_abs<dim,dof,feature> (cdcl/cdcl_estimation_abs.h:20-65)
template <unsigned int dim, unsigned int dof, class feature_typeT = cdcl_feature_with_shape<dim> >
class cdcl_estimation_abs {
 public:
  typedef feature_typeT feature_type;
  typedef typename feature_type::sptr feature_sptr_type;
  typedef cdcl_match<dim, feature_type> match_type;
  typedef typename match_type::sptr match_sptr_type;
  typedef cdcl_trans<dim, dof> trans_type;
  typedef typename trans_type::sptr trans_sptr_type;
  typedef void (*display_callback_type)(void* caller, std::vector<feature_sptr_type> const& moving, std::vector<feature_sptr_type> const& fixed,
                                        std::vector<match_sptr_type> const& matches, trans_sptr_type const& trans, unsigned int iteration);
  virtual ~cdcl_estimation_abs() {}
  virtual void initialize() = 0;
  virtual bool run(void* caller = 0, display_callback_type display_points = 0) = 0;
};

// Shared implementation of the four estimators of the path; KIND selects the reference class.
template <unsigned int dim, unsigned int dof, int KIND>
class estimation_impl : public cdcl_estimation_abs<dim, dof> {
 public:
  typedef cdcl_estimation_abs<dim, dof> base;
  typedef typename base::feature_sptr_type feature_sptr_type;
  typedef typename base::match_type match_type;
  typedef typename base::match_sptr_type match_sptr_type;
  typedef typename base::trans_type trans_type;
  typedef typename base::trans_sptr_type trans_sptr_type;
  typedef typename base::display_callback_type display_callback_type;
  typedef std::shared_ptr<estimation_impl> sptr;
  static const bool symmetric = (KIND == LRR_CDCL_SYM || KIND == LRR_RRL_SYM);

  // forward ctor: (moving, fixed, trans)                      cdcl :73-91, rrl :89-113
  estimation_impl(device_context& ctx, std::vector<feature_sptr_type> const& moving, std::vector<feature_sptr_type> const& fixed, trans_sptr_type const& trans)
      : ctx_(ctx), moving_all_(moving), fixed_all_(fixed), moving_(ctx, moving, KIND == LRR_CDCL_FWD || KIND == LRR_CDCL_SYM), fixed_(ctx, fixed, KIND == LRR_CDCL_FWD || KIND == LRR_CDCL_SYM) {
    ctx_.check(lrr_est_create(ctx_.get(), KIND, (int)dim, (int)dof, moving_.get(), fixed_.get(), &est_), "lrr_est_create");
    set_transform(trans);
  }
  // symmetric ctor #3: (moving_all, fixed_all, trans, trans_inverse, moving_inside, fixed_inside)   cdcl sym :144-164, rrl sym .h:78-83
  estimation_impl(device_context& ctx, std::vector<feature_sptr_type> const& moving_all, std::vector<feature_sptr_type> const& fixed_all, trans_sptr_type const& trans,
                  trans_sptr_type const& /*trans_inverse: recomputed as trans->inverse(), as set_transform does*/, std::vector<int32_t> const& moving_inside,
                  std::vector<int32_t> const& fixed_inside)
      : estimation_impl(ctx, moving_all, fixed_all, trans) { set_features(moving_inside, fixed_inside); set_transform(trans); }
  ~estimation_impl() { if (est_) lrr_est_destroy(est_); }

  void initialize() override {}
  // while( !converged && iteration < max ) one_iteration   (cdcl :413-428; driver loop rrl/itkLocationRegistration.cxx:2418)
  bool run(void* caller = 0, display_callback_type display_points = 0) override {
    bool converged = false; unsigned int iteration = 0;
    while (!converged && iteration < max_iterations_) { converged = one_iteration(caller, display_points, iteration); ++iteration; }
    return converged;
  }
  bool one_iteration(void* caller, display_callback_type display_points, unsigned int iteration) {
    int conv = 0;
    ctx_.check(lrr_est_one_iteration(est_, iteration, &conv), "lrr_est_one_iteration");
    if (caller && display_points) {   // the callback wants host cdcl_match objects: download indices and weights
      std::vector<match_sptr_type> m = matches(LRR_DIR_FORWARD);
      display_points(caller, moving_all_, fixed_all_, m, get_transform(), iteration);
    }
    return conv != 0;
  }
  void find_closest_euclidean() { ctx_.check(lrr_est_find_closest(est_), "lrr_est_find_closest"); }
  void set_transform(trans_sptr_type const& trans) {
    ctx_.check(lrr_est_set_transform(est_, dof == 12 ? LRR_TRANS_AFFINE : LRR_TRANS_RIGID3D, trans->get_A().data(), trans->get_translation().data(), trans->center_moving_.data(),
                                     trans->get_covariance().data()), "lrr_est_set_transform");
  }
  // set_features(moving_inside, fixed_inside): index lists into the *_all sets (the driver's deep clones, :1346,1357)
  void set_features(std::vector<int32_t> const& moving_inside, std::vector<int32_t> const& fixed_inside) {
    ctx_.check(lrr_est_set_subsets(est_, moving_inside.data(), (int64_t)moving_inside.size(), fixed_inside.empty() ? nullptr : fixed_inside.data(), (int64_t)fixed_inside.size()), "lrr_est_set_subsets");
  }
  void set_moving_data(std::vector<int32_t> const& moving_inside) { set_features(moving_inside, std::vector<int32_t>()); }
  void weight_by_strength(bool w) { weight_by_strength_ = w; ctx_.check(lrr_est_set_flags(est_, normalize_matches_, weight_by_strength_), "lrr_est_set_flags"); }
  void normalize_matches() { normalize_matches_ = true; ctx_.check(lrr_est_set_flags(est_, normalize_matches_, weight_by_strength_), "lrr_est_set_flags"); }
  void set_max_iterations(unsigned int n) { max_iterations_ = n; }
  bool set_number_points(unsigned int n, int dir = LRR_DIR_FORWARD) { int at = 0; ctx_.check(lrr_est_set_number_points(est_, dir, n, &at), "lrr_est_set_number_points"); return at != 0; }
  void set_shuffle(bool reshuffle, uint64_t seed) { ctx_.check(lrr_est_set_shuffle(est_, reshuffle ? LRR_ORDER_RESHUFFLE : LRR_ORDER_FIXED, seed), "lrr_est_set_shuffle"); }

  trans_sptr_type get_transform() { return transform(LRR_DIR_FORWARD); }
  trans_sptr_type get_transform_backward() { return transform(LRR_DIR_BACKWARD); }
  double RMS_error() { return metric(LRR_METRIC_RMS, LRR_DIR_FORWARD); }
  double weighted_error() { return metric(LRR_METRIC_WEIGHTED, LRR_DIR_FORWARD); }
  double median_error() { return metric(LRR_METRIC_MEDIAN, LRR_DIR_FORWARD); }
  double sheet_angles() { return metric(LRR_METRIC_SHEET_ANGLES, LRR_DIR_FORWARD); }
  double tube_angles() { return metric(LRR_METRIC_TUBE_ANGLES, LRR_DIR_FORWARD); }
  double RMS_error_backward() { return metric(LRR_METRIC_RMS, LRR_DIR_BACKWARD); }
  double weighted_error_backward() { return metric(LRR_METRIC_WEIGHTED, LRR_DIR_BACKWARD); }
  double sheet_angles_backward() { return metric(LRR_METRIC_SHEET_ANGLES, LRR_DIR_BACKWARD); }
  double tube_angles_backward() { return metric(LRR_METRIC_TUBE_ANGLES, LRR_DIR_BACKWARD); }
  bool oscillated() { int o = 0; ctx_.check(lrr_est_oscillated(est_, &o), "lrr_est_oscillated"); return o != 0; }
  void estimate_LS_backward(bool estimate_parameters) { ctx_.check(lrr_est_estimate_ls_backward(est_, estimate_parameters), "lrr_est_estimate_ls_backward"); }
  // compute_transfer_error_covariance(): stores covarianceJ on the returned transform (cdcl :388-404)
  mat3 compute_transfer_error_covariance(int dir = LRR_DIR_FORWARD) { mat3 S{}; ctx_.check(lrr_est_transfer_cov(est_, dir, S.data()), "lrr_est_transfer_cov"); return S; }
  std::vector<match_sptr_type> matches(int dir) {
    int64_t n = 0;
    ctx_.check(lrr_est_get_matches(est_, dir, 0, &n, nullptr, nullptr, nullptr), "lrr_est_get_matches");
    std::vector<int32_t> f((size_t)n), t((size_t)n); std::vector<double> w((size_t)n);
    ctx_.check(lrr_est_get_matches(est_, dir, n, &n, f.data(), t.data(), w.data()), "lrr_est_get_matches");
    std::vector<feature_sptr_type> const& from = dir == LRR_DIR_FORWARD ? moving_all_ : fixed_all_;
    std::vector<feature_sptr_type> const& to = dir == LRR_DIR_FORWARD ? fixed_all_ : moving_all_;
    std::vector<match_sptr_type> out((size_t)n);
    for (int64_t i = 0; i < n; ++i) { out[i].reset(new match_type); out[i]->from_ = from[f[i]]; out[i]->to_.push_back(to[t[i]]); out[i]->w_.push_back(w[i]); }
    return out;
  }
  lrr_est* handle() const { return est_; }

 private:
  trans_sptr_type transform(int dir) {
    trans_sptr_type t(new trans_type); int kind = 0;
    t->covar_.assign(dof * dof, 0.0);
    ctx_.check(lrr_est_get_transform(est_, dir, &kind, t->A_.data(), t->t_.data(), t->center_moving_.data(), t->covar_.data()), "lrr_est_get_transform");
    return t;
  }
  double metric(int which, int dir) { double v = 0; ctx_.check(lrr_est_metric(est_, which, dir, &v), "lrr_est_metric"); return v; }
  device_context& ctx_;
  std::vector<feature_sptr_type> moving_all_, fixed_all_;
  feature_set<dim> moving_, fixed_;
  lrr_est* est_ = nullptr;
  bool normalize_matches_ = symmetric, weight_by_strength_ = symmetric;
  unsigned int max_iterations_ = symmetric ? 70 : 30;   // driver :2418 / cdcl :419
};

template <unsigned int dim, unsigned int dof> using cdcl_estimation_ICP_matching_all = estimation_impl<dim, dof, LRR_CDCL_FWD>;
template <unsigned int dim, unsigned int dof> using cdcl_estimation_symmetric_ICP_matching_all = estimation_impl<dim, dof, LRR_CDCL_SYM>;
template <unsigned int dim, unsigned int dof> using rrl_estimation_ICP_matching_all = estimation_impl<dim, dof, LRR_RRL_FWD>;
template <unsigned int dim, unsigned int dof> using rrl_estimation_symmetric_ICP_matching_all = estimation_impl<dim, dof, LRR_RRL_SYM>;

}  // namespace lrr_b200
