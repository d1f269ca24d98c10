// Header-only C++ host side above the C ABI (include/lrr_gpu.h): the reference's template estimator API
// (same class names, template parameters <dim, dof>, method names, argument meaning and bool/void returns) so that
//     typedef rrl_estimation_symmetric_ICP_matching_all<dim,dof> estimation_symmetric_type;
// in rrl/itkLocationRegistration.h:336-337 can be pointed at lrr_b200::rrl_estimation_symmetric_ICP_matching_all.
// Everything lives in namespace lrr_b200 to coexist with the reference's own classes in one translation unit.
// Mirrors: cdcl/cdcl_feature_with_shape.h:23-89, cdcl/cdcl_feature_ICP.h:23-96, cdcl/cdcl_match.h:18-44,
// cdcl/cdcl_trans.h:21-148, cdcl/cdcl_trans_affine.h, cdcl/cdcl_trans_rigid3d.h, cdcl/cdcl_estimation_abs.h:20-65,
// cdcl/cdcl_estimation_ICP_matching_all.h, cdcl/cdcl_estimation_symmetric_ICP_matching_all.h,
// rrl/rrl_estimation_ICP_matching_all.h, rrl/rrl_estimation_symmetric_ICP_matching_all.h.
// Config 1 (cdcl/examples/estimate.cxx -method 4): cdcl/cdcl_feature.h:17-64, cdcl/cdcl_trans_similarity2d.h,
// cdcl/cdcl_estimation_ICP.h — the planar section at the end of this file.
// No VXL: vnl fixed vectors/matrices become std::array, vbl_smart_ptr becomes std::shared_ptr.
#pragma once
#include <array>
#include <cmath>
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

// ---- the full CDC estimator (cdcl/cdcl_estimation.h:20-100, .txx:66-545): EM over many-to-many matches ---------------------------------
// cdcl_estimation<dim,dof> derives from the ICP estimator, which supplies its initialisation; here it sits on the matching-all forward
// estimator of the path.  compute_weights = the E-step on the device (lrr_cdc_compute_weights), estimate_parameters / estimate_covariance
// = L-BFGS over cdcl_obj_fun_par / cdcl_obj_fun_cov evaluated on the device over the many-to-many set.  One resolution level per
// object: compute_weights reports avg_largest_eval, switch_resolution() gives the reference's decision (:494-508) and the caller
// re-invokes set_moving_data with that level's features, as the driver of a multi-resolution run does.  <3,12> only.
template <unsigned int dim, unsigned int dof>
class cdcl_estimation : public estimation_impl<dim, dof, LRR_CDCL_FWD> {
 public:
  typedef estimation_impl<dim, dof, LRR_CDCL_FWD> icp;
  typedef typename icp::feature_sptr_type feature_sptr_type;
  typedef typename icp::trans_sptr_type trans_sptr_type;
  typedef typename icp::display_callback_type display_callback_type;
  cdcl_estimation(device_context& ctx, std::vector<feature_sptr_type> const& moving, std::vector<feature_sptr_type> const& fixed, trans_sptr_type const& trans)
      : icp(ctx, moving, fixed, trans), cctx_(ctx) {}
  // initialize(): initial weights by find_closest_euclidean, initial covariance by estimate_LS  (:69-84)
  void initialize() override { icp::one_iteration(0, 0, 0); }
  // compute_weights( resolution_switch_allowed )  (:342-545); returns false when the covariance is too small for any match — the
  // reference then falls back to find_closest_euclidean (:544), and so does this
  bool compute_weights(bool /*resolution_switch_allowed*/ = false) {
    cctx_.check(lrr_cdc_compute_weights(this->handle(), LRR_DIR_FORWARD, 100, 5.0, &last_), "lrr_cdc_compute_weights");
    if (last_.entries == 0) { this->find_closest_euclidean(); return false; }
    return true;
  }
  lrr_cdc_summary const& last_weights() const { return last_; }
  int switch_resolution(std::vector<double> const& fixed_spacing, int res, bool allowed) const {
    int r = res;
    lrr_cdc_resolution_switch(last_.avg_largest_eval, fixed_spacing.data(), res, (int)fixed_spacing.size() - 1, allowed, &r);
    return r;
  }
  bool estimate_parameters() {   // (:207-262)
    double fs = 0, fe = 0; int ev = 0, conv = 0;
    cctx_.check(lrr_objpar_refine(this->handle(), LRR_DIR_FORWARD, 5, &fs, &fe, &ev, &conv), "lrr_objpar_refine");
    return conv != 0;
  }
  bool estimate_covariance() {   // (:264-337)
    double fs = 0, fe = 0; int ev = 0;
    cctx_.check(lrr_objcov_refine(this->handle(), LRR_DIR_FORWARD, 5, &fs, &fe, &ev), "lrr_objcov_refine");
    return (fs - fe) < 1e-5;
  }
  // one EM iteration: weights; past startEM = 5 parameters and weights again; covariance  (:88-186)
  bool one_iteration(void* /*caller*/, display_callback_type /*display_points*/, unsigned int iteration) {
    int conv = 0;
    cctx_.check(lrr_cdc_one_iteration(this->handle(), LRR_DIR_FORWARD, iteration, 100, 5.0, &conv, &last_), "lrr_cdc_one_iteration");
    return conv != 0;
  }
  bool run(void* caller = 0, display_callback_type display_points = 0) override {   // (:190-204): initialize, then at most 85 iterations
    initialize();
    bool converged = false; unsigned int iteration = 0;
    while (!converged && iteration < 85) { converged = one_iteration(caller, display_points, iteration); ++iteration; }
    return converged;
  }
 private:
  device_context& cctx_;
  lrr_cdc_summary last_ = lrr_cdc_summary();
};

// ---- planar demo estimator (BASELINE.json configs[0]) -----------------------------------------------------------
typedef std::array<double, 2> vec2;
typedef std::array<double, 4> mat2;   // row-major

// cdcl_feature<2>: location, strength, covariance (cdcl/cdcl_feature.h:17-64); the normal is the eigenvector of the smallest
// eigenvalue of covariance_ (cdcl_estimation_ICP.txx:470-477) and is derived on the device at upload
struct cdcl_feature2 {
  typedef std::shared_ptr<cdcl_feature2> sptr;
  cdcl_feature2() {}
  cdcl_feature2(vec2 const& location, double strength, mat2 const& covariance) : location_(location), strength_(strength), covariance_(covariance) {}
  vec2 location_{}; double strength_ = 0.0; mat2 covariance_{};
};

// cdcl_trans_similarity2d (cdcl/cdcl_trans_similarity2d.{h,cxx}): y = s R(ang) x + t, parameters (s cos, s sin, tx, ty)
class cdcl_trans_similarity2d {
 public:
  static const unsigned int dim = 2, dof = 4;
  typedef std::shared_ptr<cdcl_trans_similarity2d> sptr;
  cdcl_trans_similarity2d() { covar_.assign(16, 0.0); }
  cdcl_trans_similarity2d(double s, double ang, vec2 const& t) : s_(s), ang_(ang), t_(t) { covar_.assign(16, 0.0); }
  vec2 map_loc(vec2 const& x) const { const mat2 A = jacobian_wrt_loc(); return vec2{A[0] * x[0] + A[1] * x[1] + t_[0], A[2] * x[0] + A[3] * x[1] + t_[1]}; }
  mat2 jacobian_wrt_loc() const { const double c = std::cos(ang_), s = std::sin(ang_); return mat2{s_ * c, -s_ * s, s_ * s, s_ * c}; }
  std::vector<double> get_parameterization() const { return {s_ * std::cos(ang_), s_ * std::sin(ang_), t_[0], t_[1]}; }
  void set_parameterization(std::vector<double> const& p) { s_ = std::sqrt(p[0] * p[0] + p[1] * p[1]); ang_ = std::atan2(p[1] / s_, p[0] / s_); t_ = {p[2], p[3]}; }
  std::vector<double> const& get_covariance() const { return covar_; }   // 4 x 4 row-major
  void set_covariance(std::vector<double> const& c) { covar_ = c; }
  mat2 const& get_covarianceJ() const { return covarJ_; }
  // normalize / unnormalize of the parameters (similarity2d.cxx:227-305); the covariance follows B C B^T with the same B
  void normalize(double rm, double rf, vec2 const& cm, vec2 const& cf) {
    s_ = rm / rf * s_;
    const mat2 sR = jacobian_wrt_loc();
    for (int r = 0; r < 2; ++r) t_[r] = t_[r] / rf + (sR[2 * r] * cm[0] + sR[2 * r + 1] * cm[1]) / rm - cf[r] / rf;
    double B[16] = {0}; B[0] = B[5] = rm / rf; B[8] = cm[0] / rf; B[9] = -cm[1] / rf; B[12] = cm[1] / rf; B[13] = cm[0] / rf; B[10] = B[15] = 1.0 / rf;
    congruence(B);
  }
  void unnormalize(double rm, double rf, vec2 const& cm, vec2 const& cf) {
    s_ = rf / rm * s_;
    const mat2 sR = jacobian_wrt_loc();
    for (int r = 0; r < 2; ++r) t_[r] = -(sR[2 * r] * cm[0] + sR[2 * r + 1] * cm[1]) + rf * t_[r] + cf[r];
    double B[16] = {0}; B[0] = B[5] = rf / rm; B[10] = B[15] = rf;
    congruence(B);
  }
  void print(FILE* f = stdout) const { std::fprintf(f, "s: %.12g ang: %.12g t: %.12g, %.12g\n", s_, ang_, t_[0], t_[1]); }
  double s_ = 1.0, ang_ = 0.0; vec2 t_{}; std::vector<double> covar_; mat2 covarJ_{};
 private:
  void congruence(const double* B) {
    double T[16], R[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double v = 0; for (int k = 0; k < 4; ++k) v += B[4 * i + k] * covar_[4 * k + j]; T[4 * i + j] = v; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double v = 0; for (int k = 0; k < 4; ++k) v += T[4 * i + k] * B[4 * j + k]; R[4 * i + j] = v; }
    covar_.assign(R, R + 16);
  }
};

// cdcl_normalize_data (cdcl/cdcl_utils.txx): centre and mean radius of a feature vector; locations and covariances rescaled in place
inline void cdcl_normalize_data(std::vector<cdcl_feature2::sptr>& f, vec2& center, double& avg_rad) {
  center = {0, 0}; avg_rad = 0;
  for (auto& p : f) { center[0] += p->location_[0]; center[1] += p->location_[1]; }
  center[0] /= (double)f.size(); center[1] /= (double)f.size();
  for (auto& p : f) { p->location_[0] -= center[0]; p->location_[1] -= center[1]; avg_rad += std::sqrt(p->location_[0] * p->location_[0] + p->location_[1] * p->location_[1]); }
  avg_rad /= (double)f.size();
  for (auto& p : f) { for (int d = 0; d < 2; ++d) p->location_[d] /= avg_rad; for (int k = 0; k < 4; ++k) p->covariance_[k] /= avg_rad * avg_rad; }
}

// cdcl_estimation_ICP<2,4> (cdcl/cdcl_estimation_ICP.{h,txx}): ctor #1, run / one_iteration, the setters of the header
template <unsigned int dim, unsigned int dof>
class cdcl_estimation_ICP {
 public:
  static_assert(dim == 2 && dof == 4, "the device path of this estimator is built for <2,4> (cdcl/Templates/cdcl_estimation_ICP+2.4-.cxx)");
  typedef cdcl_feature2 feature_type;
  typedef feature_type::sptr feature_sptr_type;
  typedef cdcl_trans_similarity2d::sptr trans_sptr_type;
  typedef std::shared_ptr<cdcl_estimation_ICP> sptr;
  cdcl_estimation_ICP(device_context& ctx, std::vector<feature_sptr_type> const& moving, std::vector<feature_sptr_type> const& fixed, trans_sptr_type const& trans) : ctx_(ctx) {
    moving_ = upload(moving); fixed_ = upload(fixed);
    ctx_.check(lrr_est_create(ctx_.get(), LRR_CDCL_ICP, 2, 4, moving_, fixed_, &est_), "lrr_est_create");
    set_transform(trans);
  }
  ~cdcl_estimation_ICP() { if (est_) lrr_est_destroy(est_); if (moving_) lrr_features_destroy(moving_); if (fixed_) lrr_features_destroy(fixed_); }
  cdcl_estimation_ICP(cdcl_estimation_ICP const&) = delete;
  void initialize() {}
  bool run(void* = 0, void* = 0) {   // cdcl_estimation_ICP.txx:330-346
    bool converged = false; unsigned int iteration = 0;
    while (!converged && iteration < 30) { converged = one_iteration(0, 0, iteration); ++iteration; }
    return converged;
  }
  bool one_iteration(void*, void*, unsigned int iteration) { int c = 0; ctx_.check(lrr_est_one_iteration(est_, iteration, &c), "lrr_est_one_iteration"); return c != 0; }
  void set_transform(trans_sptr_type const& trans) {
    const mat2 A = trans->jacobian_wrt_loc();
    const double A9[9] = {A[0], A[1], 0, A[2], A[3], 0, 0, 0, 1}, t3[3] = {trans->t_[0], trans->t_[1], 0};
    ctx_.check(lrr_est_set_transform(est_, LRR_TRANS_SIMILARITY2D, A9, t3, nullptr, trans->get_covariance().data()), "lrr_est_set_transform");
  }
  trans_sptr_type get_transform() {
    double A9[9], t3[3], c3[3]; int kind = 0; std::vector<double> cov(16);
    ctx_.check(lrr_est_get_transform(est_, 0, &kind, A9, t3, c3, cov.data()), "lrr_est_get_transform");
    trans_sptr_type t(new cdcl_trans_similarity2d(std::sqrt(A9[0] * A9[0] + A9[3] * A9[3]), std::atan2(A9[3], A9[0]), vec2{t3[0], t3[1]}));
    t->set_covariance(cov);
    return t;
  }
  void normalize_matches() { normalize_ = true; ctx_.check(lrr_est_set_flags(est_, normalize_, strength_), "lrr_est_set_flags"); }
  void weight_by_strength(bool w) { strength_ = w; ctx_.check(lrr_est_set_flags(est_, normalize_, strength_), "lrr_est_set_flags"); }
  double RMS_error() { return metric(LRR_METRIC_RMS); }
  double weighted_error() { return metric(LRR_METRIC_WEIGHTED); }
  double median_error() { return metric(LRR_METRIC_MEDIAN); }
  mat2 compute_transfer_error_covariance() { double S[9]; ctx_.check(lrr_est_transfer_cov(est_, 0, S), "lrr_est_transfer_cov"); return mat2{S[0], S[1], S[3], S[4]}; }
 private:
  lrr_fset* upload(std::vector<feature_sptr_type> const& f) {
    const size_t n = f.size();
    std::vector<double> loc(2 * n), cov(3 * n), s(n);
    for (size_t i = 0; i < n; ++i) {
      loc[i] = f[i]->location_[0]; loc[n + i] = f[i]->location_[1];
      cov[i] = f[i]->covariance_[0]; cov[n + i] = f[i]->covariance_[1]; cov[2 * n + i] = f[i]->covariance_[3];
      s[i] = f[i]->strength_;
    }
    lrr_fset* fs = nullptr;
    ctx_.check(lrr_features_upload_cov(ctx_.get(), 2, (int64_t)n, loc.data(), cov.data(), s.data(), &fs), "lrr_features_upload_cov");
    return fs;
  }
  double metric(int which) { double v = 0; ctx_.check(lrr_est_metric(est_, which, 0, &v), "lrr_est_metric"); return v; }
  device_context& ctx_;
  lrr_fset *moving_ = nullptr, *fixed_ = nullptr; lrr_est* est_ = nullptr;
  bool normalize_ = false, strength_ = false;
};

}  // namespace lrr_b200
