// lrr_dropin.h — the DROP-IN estimators: the reference's own class names, template parameters, nested typedefs, constructor and
// method signatures, over the reference's OWN feature / match / transform types, backed by the CUDA library behind
// include/lrr_gpu.h.  This is the header a maintainer of msofka/LRR includes: it needs the reference's cdcl
// headers and VXL on the include path (both are there in an LRR build tree); the product library itself needs neither.
//
//   rrl/itkLocationRegistration.h:336-337
//     -  typedef rrl_estimation_symmetric_ICP_matching_all< dim, dof >  estimation_symmetric_type;
//     +  typedef lrr_b200::rrl_estimation_symmetric_ICP_matching_all< dim, dof >  estimation_symmetric_type;
//
// and every call site of the driver compiles unchanged (replayed verbatim by oracle/dropin_replay.cpp, which also runs the
// reference's own class on the same objects and compares):
//   ctor (rrl sym .h:78-83) + weight_by_strength / normalize_matches / set_voronoi_maps    rrl/itkLocationRegistration.cxx:1640-1643
//   set_features / set_transform                                                           :1636-1637
//   find_closest_euclidean + sheet / tube angles (forward, backward)                        :2303-2313
//   one_iteration( caller, display, iteration )                                            :2446
//   sheet_angles / tube_angles / weighted_error (+ _backward), estimate_LS_backward        :1828-1845
//   get_transform / get_transform_backward, oscillated                                     :1833,1840,2504
//   compute_transfer_error_covariance + get_covarianceJ                                    :1952
// Feature vectors are vectors of the reference's ref-counted smart pointers; the *_inside sets are the driver's deep CLONES of
// elements of the *_all sets (:1346,1357) and are mapped back to indices by their locations at set_features.
// The Voronoi maps are accepted and ignored: the backend searches the exact nearest neighbour (north_star item 2, Q5).
#pragma once
#include <vcl_vector.h>
#include <vbl/vbl_ref_count.h>
#include <vbl/vbl_smart_ptr.h>
#include <vnl/vnl_vector.h>
#include <vnl/vnl_vector_fixed.h>
#include <vnl/vnl_matrix_fixed.h>
#include <cdcl/cdcl_feature.h>
#include <cdcl/cdcl_feature_with_shape.h>
#include <cdcl/cdcl_feature_ICP.h>
#include <cdcl/cdcl_match.h>
#include <cdcl/cdcl_trans.h>
#include <cdcl/cdcl_trans_affine.h>
#include <cdcl/cdcl_trans_rigid3d.h>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../lrr_gpu.h"

namespace lrr_b200 {

// One context per device for the whole process, created on first use (the reference has no notion of a device: the estimators
// are constructed without one).
inline lrr_ctx* default_context(int device = 0) {
  static lrr_ctx* ctx[64] = {0};
  if (device < 0 || device >= 64) throw std::runtime_error("lrr_b200: bad device index");
  if (!ctx[device]) {
    const int rc = lrr_ctx_create(device, &ctx[device]);
    if (rc != LRR_OK) throw std::runtime_error("lrr_b200: lrr_ctx_create failed (status " + std::to_string(rc) + "); a CUDA device is required, there is no CPU fallback");
  }
  return ctx[device];
}
inline void dropin_check(lrr_ctx* ctx, int rc, const char* what) {
  if (rc != LRR_OK) throw std::runtime_error(std::string("lrr_b200: ") + what + " failed (status " + std::to_string(rc) + "): " + lrr_ctx_last_error(ctx));
}

// covariance_ exists in cdcl_feature / cdcl_feature_with_shape, not in cdcl_feature_ICP (cdcl/cdcl_feature_ICP.h:23-96)
template <class F> struct dropin_has_covariance { static const bool value = true; };
template <unsigned int dim> struct dropin_has_covariance< cdcl_feature_ICP<dim> > { static const bool value = false; };
template <bool has> struct dropin_cov { template <class P> static double get(P const& f, int r, int c) { return f->covariance_(r, c); } };
template <> struct dropin_cov<false> { template <class P> static double get(P const&, int, int) { return 0.0; } };

// the two transform models of the path, constructed from (A, t, centre) (cdcl_trans_affine.h:38, cdcl_trans_rigid3d.h:42)
template <unsigned int dim, unsigned int dof> struct dropin_trans;
template <unsigned int dim> struct dropin_trans<dim, 12> {
  static int kind() { return LRR_TRANS_AFFINE; }
  static typename cdcl_trans<dim, 12>::sptr make(vnl_matrix_fixed<double, dim, dim> const& A, vnl_vector_fixed<double, dim> const& t, vnl_vector_fixed<double, dim> const& c) {
    return new cdcl_trans_affine<dim, 12>(A, t, c);
  }
};
template <unsigned int dim> struct dropin_trans<dim, 6> {
  static int kind() { return LRR_TRANS_RIGID3D; }
  static typename cdcl_trans<dim, 6>::sptr make(vnl_matrix_fixed<double, dim, dim> const& A, vnl_vector_fixed<double, dim> const& t, vnl_vector_fixed<double, dim> const& c) {
    return new cdcl_trans_rigid3d(A, t, c);
  }
};

template <unsigned int dim, unsigned int dof, class feature_typeT, int KIND>
class estimation_dropin : public vbl_ref_count {
 public:
  typedef vbl_smart_ptr<estimation_dropin> sptr;
  typedef feature_typeT feature_type;
  typedef typename feature_type::sptr feature_sptr_type;
  typedef cdcl_match<dim, feature_type> match_type;
  typedef typename match_type::sptr match_sptr_type;
  typedef cdcl_trans<dim, dof> trans_type;
  typedef typename trans_type::sptr trans_sptr_type;
  typedef vnl_vector_fixed<double, dim> location_type;
  typedef void (*display_callback_type)(void* caller, vcl_vector<feature_sptr_type> const& moving, vcl_vector<feature_sptr_type> const& fixed,
                                        vcl_vector<match_sptr_type> const& matches, trans_sptr_type const& trans, unsigned int iteration);
  static const bool symmetric = (KIND == LRR_CDCL_SYM || KIND == LRR_RRL_SYM);

  // forward estimators: (moving, fixed, trans)   cdcl/cdcl_estimation_ICP_matching_all.h, rrl/rrl_estimation_ICP_matching_all.h
  estimation_dropin(vcl_vector<feature_sptr_type> const& moving, vcl_vector<feature_sptr_type> const& fixed, trans_sptr_type const& trans)
      : ctx_(default_context()), moving_all_(moving), fixed_all_(fixed) { construct(trans); }
  // symmetric estimators, ctor #3 (rrl sym .h:78-83, cdcl sym :144-164)
  estimation_dropin(vcl_vector<feature_sptr_type> const& moving, vcl_vector<feature_sptr_type> const& fixed, trans_sptr_type const& trans,
                    trans_sptr_type const& /* trans_inverse: set_transform recomputes trans->inverse(), as the reference does (cdcl sym :231-242) */,
                    vcl_vector<feature_sptr_type> const& moving_inside, vcl_vector<feature_sptr_type> const& fixed_inside)
      : ctx_(default_context()), moving_all_(moving), fixed_all_(fixed) {
    construct(trans);
    set_features(moving_inside, fixed_inside);
    set_transform(trans);
  }
  ~estimation_dropin() {
    if (est_) lrr_est_destroy(est_);
    if (fs_moving_) lrr_features_destroy(fs_moving_);
    if (fs_fixed_) lrr_features_destroy(fs_fixed_);
  }

  void initialize() {}
  bool run(void* caller = 0, display_callback_type display_points = 0) {
    bool converged = false; unsigned int iteration = 0;
    while (!converged && iteration < (symmetric ? 70u : 30u)) { converged = one_iteration(caller, display_points, iteration); ++iteration; }
    return converged;
  }
  bool one_iteration(void* caller, display_callback_type display_points, unsigned int iteration) {
    int conv = 0;
    dropin_check(ctx_, lrr_est_one_iteration(est_, iteration, &conv), "one_iteration");
    if (display_points) {   // the callback wants host cdcl_match objects (cdcl/cdcl_estimation_abs.h:50-55): indices and weights come back
      vcl_vector<match_sptr_type> m = matches(LRR_DIR_FORWARD);
      display_points(caller, moving_all_, fixed_all_, m, get_transform(), iteration);
    }
    return conv != 0;
  }
  void find_closest_euclidean() { dropin_check(ctx_, lrr_est_find_closest(est_), "find_closest_euclidean"); }
  void set_transform(trans_sptr_type const& trans) {
    double A[9], t[3], c[3], cov[dof * dof];
    for (unsigned r = 0; r < 3; ++r) { for (unsigned k = 0; k < 3; ++k) A[3 * r + k] = trans->get_A()(r, k); t[r] = trans->get_translation()[r]; c[r] = trans->center_moving_[r]; }
    for (unsigned i = 0; i < dof; ++i) for (unsigned j = 0; j < dof; ++j) cov[i * dof + j] = trans->get_covariance()(i, j);
    dropin_check(ctx_, lrr_est_set_transform(est_, dropin_trans<dim, dof>::kind(), A, t, c, cov), "set_transform");
  }
  // set_features( moving_inside, fixed_inside )   (rrl sym :269-289, cdcl sym :206-226)
  void set_features(vcl_vector<feature_sptr_type> const& moving_inside, vcl_vector<feature_sptr_type> const& fixed_inside) {
    std::vector<int32_t> mi = indices_of(moving_inside, index_moving_, moving_all_), fi = indices_of(fixed_inside, index_fixed_, fixed_all_);
    dropin_check(ctx_, lrr_est_set_subsets(est_, mi.data(), (int64_t)mi.size(), fi.empty() ? 0 : fi.data(), (int64_t)fi.size()), "set_features");
  }
  // forward estimators: set_moving_data( multires vector with one level )   (cdcl :.h, rrl .h)
  void set_moving_data(vcl_vector< vcl_vector<feature_sptr_type> > const& moving) {
    std::vector<int32_t> mi = indices_of(moving[0], index_moving_, moving_all_);
    dropin_check(ctx_, lrr_est_set_subsets(est_, mi.data(), (int64_t)mi.size(), 0, 0), "set_moving_data");
  }
  void weight_by_strength(bool w) { weight_by_strength_ = w; dropin_check(ctx_, lrr_est_set_flags(est_, normalize_matches_, weight_by_strength_), "weight_by_strength"); }
  void normalize_matches() { normalize_matches_ = true; dropin_check(ctx_, lrr_est_set_flags(est_, normalize_matches_, weight_by_strength_), "normalize_matches"); }
  template <class MapPointer> void set_voronoi_maps(MapPointer const&, MapPointer const&) {}   // exact neighbours instead (Q5)
  template <class MapPointer> void set_fixed_voronoi_map(MapPointer const&) {}

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
  bool oscillated() { int o = 0; dropin_check(ctx_, lrr_est_oscillated(est_, &o), "oscillated"); return o != 0; }
  void estimate_LS_backward(bool estimate_parameters) { dropin_check(ctx_, lrr_est_estimate_ls_backward(est_, estimate_parameters), "estimate_LS_backward"); }
  // compute_transfer_error_covariance(): the reference stores the result on the transform (set_covarianceJ, cdcl :388-404); here it
  // is stored on the transforms get_transform() returns from now on
  void compute_transfer_error_covariance() {
    double S[9];
    dropin_check(ctx_, lrr_est_transfer_cov(est_, LRR_DIR_FORWARD, S), "compute_transfer_error_covariance");
    for (unsigned r = 0; r < 3; ++r) for (unsigned c = 0; c < 3; ++c) covarJ_(r, c) = S[3 * r + c];
    have_covarJ_ = true;
  }
  // the current matches as host objects (from_ / to_ are the caller's own feature objects of the *_all sets)
  vcl_vector<match_sptr_type> matches(int dir) {
    int64_t n = 0;
    dropin_check(ctx_, lrr_est_get_matches(est_, dir, 0, &n, 0, 0, 0), "get_matches");
    std::vector<int32_t> f((size_t)n), t((size_t)n); std::vector<double> w((size_t)n);
    dropin_check(ctx_, lrr_est_get_matches(est_, dir, n, &n, f.data(), t.data(), w.data()), "get_matches");
    vcl_vector<feature_sptr_type> const& from = dir == LRR_DIR_FORWARD ? moving_all_ : fixed_all_;
    vcl_vector<feature_sptr_type> const& to = dir == LRR_DIR_FORWARD ? fixed_all_ : moving_all_;
    vcl_vector<match_sptr_type> out;
    out.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      vcl_vector<feature_sptr_type> tv(1, to[t[i]]); vcl_vector<double> wv(1, w[i]);
      out.push_back(new match_type(from[f[i]], tv, wv));
    }
    return out;
  }
  lrr_est* handle() const { return est_; }

 private:
  estimation_dropin(estimation_dropin const&);
  struct key3 { uint64_t b[3]; bool operator==(key3 const& o) const { return b[0] == o.b[0] && b[1] == o.b[1] && b[2] == o.b[2]; } };
  struct key3_hash { size_t operator()(key3 const& k) const { return (size_t)(k.b[0] * 0x9E3779B97F4A7C15ULL ^ (k.b[1] + 0xC2B2AE3D27D4EB4FULL) * 0xD6E8FEB86659FD93ULL ^ k.b[2] * 0xFF51AFD7ED558CCDULL); } };
  typedef std::unordered_map<key3, int32_t, key3_hash> index_type;
  static key3 key_of(feature_sptr_type const& f) { key3 k; for (int d = 0; d < 3; ++d) { double v = f->location_[d]; std::memcpy(&k.b[d], &v, 8); } return k; }

  lrr_fset* upload(vcl_vector<feature_sptr_type> const& f, index_type& index) {
    const size_t n = f.size();
    std::vector<double> xyz(3 * n), P6(6 * n), C6(6 * n), s(n), dirs(9 * n, 0.0);
    std::vector<uint8_t> shape(n);
    static const int iu[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    const bool with_cov = dropin_has_covariance<feature_type>::value && (KIND == LRR_CDCL_FWD || KIND == LRR_CDCL_SYM);
    for (size_t i = 0; i < n; ++i) {
      for (int d = 0; d < 3; ++d) xyz[d * n + i] = f[i]->location_[d];
      for (int k = 0; k < 6; ++k) {
        P6[k * n + i] = f[i]->error_projector_(iu[k][0], iu[k][1]);
        C6[k * n + i] = dropin_cov<dropin_has_covariance<feature_type>::value>::get(f[i], iu[k][0], iu[k][1]);
      }
      s[i] = f[i]->strength_; shape[i] = (uint8_t)f[i]->shape_;
      for (size_t a = 0; a < f[i]->directions_.size() && a < 3; ++a) for (int d = 0; d < 3; ++d) dirs[(3 * a + d) * n + i] = f[i]->directions_[a][d];
      index.insert(std::make_pair(key_of(f[i]), (int32_t)i));   // (the first feature at a location wins)
    }
    lrr_fset* fs = 0;
    dropin_check(ctx_, lrr_features_upload(ctx_, (int64_t)n, xyz.data(), P6.data(), with_cov ? C6.data() : 0, s.data(), shape.data(), dirs.data(), &fs), "lrr_features_upload");
    return fs;
  }
  void construct(trans_sptr_type const& trans) {
    proto_ = trans;
    fs_moving_ = upload(moving_all_, index_moving_); fs_fixed_ = upload(fixed_all_, index_fixed_);
    dropin_check(ctx_, lrr_est_create(ctx_, KIND, (int)dim, (int)dof, fs_moving_, fs_fixed_, &est_), "lrr_est_create");
    normalize_matches_ = symmetric; weight_by_strength_ = symmetric;   // the symmetric ctor switches both on (cdcl sym :151-156)
    set_transform(trans);
  }
  // the driver's *_inside features are deep clones of elements of the *_all sets: found again by their location bits
  std::vector<int32_t> indices_of(vcl_vector<feature_sptr_type> const& inside, index_type const& index, vcl_vector<feature_sptr_type> const& all) {
    std::vector<int32_t> out; out.reserve(inside.size());
    for (size_t i = 0; i < inside.size(); ++i) {
      typename index_type::const_iterator it = index.find(key_of(inside[i]));
      if (it == index.end()) throw std::runtime_error("lrr_b200: set_features: a feature of the inside set is not a (clone of a) feature of the set given at construction");
      out.push_back(it->second);
    }
    (void)all;
    return out;
  }
  trans_sptr_type transform(int dir) {
    int kind = 0; double A[9], t[3], c[3], cov[dof * dof];
    dropin_check(ctx_, lrr_est_get_transform(est_, dir, &kind, A, t, c, cov), "get_transform");
    vnl_matrix_fixed<double, dim, dim> Am; vnl_vector_fixed<double, dim> tv, cv;
    for (unsigned r = 0; r < 3; ++r) { for (unsigned k = 0; k < 3; ++k) Am(r, k) = A[3 * r + k]; tv[r] = t[r]; cv[r] = c[r]; }
    trans_sptr_type out = dropin_trans<dim, dof>::make(Am, tv, cv);
    vnl_matrix_fixed<double, dof, dof> C;
    for (unsigned i = 0; i < dof; ++i) for (unsigned j = 0; j < dof; ++j) C(i, j) = cov[i * dof + j];
    out->set_covariance(C);
    if (have_covarJ_ && dir == LRR_DIR_FORWARD) out->set_covarianceJ(covarJ_);
    return out;
  }
  double metric(int which, int dir) { double v = 0; dropin_check(ctx_, lrr_est_metric(est_, which, dir, &v), "metric"); return v; }

  lrr_ctx* ctx_;
  vcl_vector<feature_sptr_type> moving_all_, fixed_all_;
  index_type index_moving_, index_fixed_;
  lrr_fset *fs_moving_ = 0, *fs_fixed_ = 0;
  lrr_est* est_ = 0;
  trans_sptr_type proto_;
  bool normalize_matches_ = false, weight_by_strength_ = false, have_covarJ_ = false;
  vnl_matrix_fixed<double, dim, dim> covarJ_;
};

// the four classes of the path under the reference's names
template <unsigned int dim, unsigned int dof> class cdcl_estimation_ICP_matching_all
    : public estimation_dropin<dim, dof, cdcl_feature_with_shape<dim>, LRR_CDCL_FWD> {
 public:
  typedef estimation_dropin<dim, dof, cdcl_feature_with_shape<dim>, LRR_CDCL_FWD> impl;
  typedef vbl_smart_ptr<cdcl_estimation_ICP_matching_all> sptr;
  cdcl_estimation_ICP_matching_all(vcl_vector<typename impl::feature_sptr_type> const& moving, vcl_vector<typename impl::feature_sptr_type> const& fixed,
                                   typename impl::trans_sptr_type const& trans) : impl(moving, fixed, trans) {}
};
template <unsigned int dim, unsigned int dof> class cdcl_estimation_symmetric_ICP_matching_all
    : public estimation_dropin<dim, dof, cdcl_feature_with_shape<dim>, LRR_CDCL_SYM> {
 public:
  typedef estimation_dropin<dim, dof, cdcl_feature_with_shape<dim>, LRR_CDCL_SYM> impl;
  typedef vbl_smart_ptr<cdcl_estimation_symmetric_ICP_matching_all> sptr;
  cdcl_estimation_symmetric_ICP_matching_all(vcl_vector<typename impl::feature_sptr_type> const& moving, vcl_vector<typename impl::feature_sptr_type> const& fixed,
                                             typename impl::trans_sptr_type const& trans, typename impl::trans_sptr_type const& trans_inverse,
                                             vcl_vector<typename impl::feature_sptr_type> const& moving_inside, vcl_vector<typename impl::feature_sptr_type> const& fixed_inside)
      : impl(moving, fixed, trans, trans_inverse, moving_inside, fixed_inside) {}
};
template <unsigned int dim, unsigned int dof> class rrl_estimation_ICP_matching_all
    : public estimation_dropin<dim, dof, cdcl_feature_ICP<dim>, LRR_RRL_FWD> {
 public:
  typedef estimation_dropin<dim, dof, cdcl_feature_ICP<dim>, LRR_RRL_FWD> impl;
  typedef vbl_smart_ptr<rrl_estimation_ICP_matching_all> sptr;
  rrl_estimation_ICP_matching_all(vcl_vector<typename impl::feature_sptr_type> const& moving, vcl_vector<typename impl::feature_sptr_type> const& fixed,
                                  typename impl::trans_sptr_type const& trans) : impl(moving, fixed, trans) {}
};
template <unsigned int dim, unsigned int dof> class rrl_estimation_symmetric_ICP_matching_all
    : public estimation_dropin<dim, dof, cdcl_feature_ICP<dim>, LRR_RRL_SYM> {
 public:
  typedef estimation_dropin<dim, dof, cdcl_feature_ICP<dim>, LRR_RRL_SYM> impl;
  typedef vbl_smart_ptr<rrl_estimation_symmetric_ICP_matching_all> sptr;
  rrl_estimation_symmetric_ICP_matching_all(vcl_vector<typename impl::feature_sptr_type> const& moving, vcl_vector<typename impl::feature_sptr_type> const& fixed,
                                            typename impl::trans_sptr_type const& trans, typename impl::trans_sptr_type const& trans_inverse,
                                            vcl_vector<typename impl::feature_sptr_type> const& moving_inside, vcl_vector<typename impl::feature_sptr_type> const& fixed_inside)
      : impl(moving, fixed, trans, trans_inverse, moving_inside, fixed_inside) {}
};

}  // namespace lrr_b200
