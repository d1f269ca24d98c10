// ORACLE — TEST INFRASTRUCTURE ONLY (see lrr_oracle_core.hpp header).  C API of the CPU restatement.
// Exports the same functions as include/lrr_gpu.h with the prefix `lrro_`, so one parity script drives
// both libraries.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
#include <cstring>
#include <string>
#include <thread>
#include <atomic>
#include <set>
#include "../include/lrr_gpu.h"
#include "lrr_oracle_core.hpp"

using namespace lrro;

struct lrro_ctx { std::string err; };
struct lrro_fset {
  lrro_ctx* ctx;
  std::vector<Feature> base;   // pristine copies
  std::shared_ptr<KdTree> kd;
  bool has_cov = false, has_dirs = false;
  int dim = 3;
  std::vector<int32_t> labels;   // region label per feature (FeaturesInRegions)
};
struct lrro_est {
  lrro_ctx* ctx; int kind, dof; lrro_fset *moving_all, *fixed_all;
  std::vector<FeatureP> moving_all_f, fixed_all_f;       // the estimator's own *_all objects
  std::vector<int32_t> moving_inside_idx, fixed_inside_idx;
  std::shared_ptr<FwdEst> fwd; std::shared_ptr<SymEst> sym;
  int shuffle_mode = 0; uint64_t seed = 0;
  int flavor() const { return kind == LRR_CDCL_ICP ? FLAVOR_ICP : ((kind == LRR_RRL_FWD || kind == LRR_RRL_SYM) ? FLAVOR_RRL : FLAVOR_CDCL); }
  int trans_kind() const { return dof == 12 ? TRANS_AFFINE : (dof == 6 ? TRANS_RIGID3D : TRANS_SIM2D); }
  bool is_sym() const { return kind == LRR_CDCL_SYM || kind == LRR_RRL_SYM; }
  FwdEst* dir(int d) { if (is_sym()) return d == 0 ? sym->forward_.get() : sym->backward_.get(); return d == 0 ? fwd.get() : nullptr; }
};

static std::vector<FeatureP> clone_set(const std::vector<Feature>& base) {
  std::vector<FeatureP> v; v.reserve(base.size());
  for (auto& f : base) v.push_back(std::make_shared<Feature>(f));
  return v;
}
static std::vector<FeatureP> clone_subset(const std::vector<Feature>& base, const std::vector<int32_t>& idx) {
  std::vector<FeatureP> v; v.reserve(idx.size());
  for (int32_t i : idx) v.push_back(std::make_shared<Feature>(base[i]));
  return v;
}
static void apply_policy(lrro_est* e) {
  for (int d = 0; d < 2; ++d) if (FwdEst* f = e->dir(d)) {
    f->shuffle_mode = e->shuffle_mode; f->shuffle_seed = e->seed; f->shuffle_dir = (uint64_t)d; f->shuffle_calls = 0;
  }
}

extern "C" {

int lrro_ctx_create(int, lrro_ctx** out) { *out = new lrro_ctx(); return LRR_OK; }
void lrro_ctx_destroy(lrro_ctx* c) { delete c; }
int lrro_ctx_set_stream(lrro_ctx*, void*) { return LRR_OK; }
int lrro_ctx_set_async_upload(lrro_ctx*, int) { return LRR_OK; }
int lrro_ctx_set_option(lrro_ctx*, const char*, int) { return LRR_OK; }   // (switches of the CUDA path: nothing to switch here)   // (host memory: nothing to overlap)
int lrro_ctx_synchronize(lrro_ctx*) { return LRR_OK; }
const char* lrro_ctx_last_error(lrro_ctx* c) { return c->err.c_str(); }
int64_t lrro_ctx_launch_count(lrro_ctx*) { return 0; }

int lrro_features_upload(lrro_ctx* ctx, int64_t n, const double* xyz, const double* P6, const double* C6,
                         const double* strength, const uint8_t* shape, const double* dirs, lrro_fset** out) {
  if (!ctx || n < 0 || !xyz || !P6 || !out) return LRR_ERR_INVALID;
  auto* fs = new lrro_fset(); fs->ctx = ctx; fs->base.resize(n);
  fs->has_cov = C6 != nullptr; fs->has_dirs = dirs != nullptr;
  static const int sym6[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
  std::vector<double> pts(3 * n);
  for (int64_t i = 0; i < n; ++i) {
    Feature& f = fs->base[i];
    for (int d = 0; d < 3; ++d) { f.loc[d] = xyz[d * n + i]; pts[3 * i + d] = f.loc[d]; }
    f.strength = strength ? strength[i] : 0.0; f.strength_f = (float)f.strength;
    for (int k = 0; k < 9; ++k) { f.P[k] = P6[sym6[k] * n + i]; f.cov[k] = C6 ? C6[sym6[k] * n + i] : 0.0; }
    f.shape = shape ? shape[i] : CORNER;
    f.ndirs = (f.shape == SHEET) ? 1 : (f.shape == TUBE ? 2 : 3);
    for (int a = 0; a < 3; ++a) for (int d = 0; d < 3; ++d) f.dirs[a][d] = dirs ? dirs[(3 * a + d) * n + i] : 0.0;
    f.id = (int)i;
  }
  fs->kd = std::make_shared<KdTree>(); fs->kd->build(pts);
  *out = fs; return LRR_OK;
}
// plain cdcl_feature<2> sets for cdcl_estimation_ICP<2,4>: the projector is the normal's outer product (cdcl_estimation_ICP.txx:470-477)
int lrro_features_upload_cov(lrro_ctx* ctx, int dim, int64_t n, const double* loc, const double* cov, const double* strength, lrro_fset** out) {
  if (!ctx || n < 0 || !loc || !cov || !out) return LRR_ERR_INVALID;
  if (dim != 2) return LRR_ERR_UNSUPPORTED;
  auto* fs = new lrro_fset(); fs->ctx = ctx; fs->base.resize(n); fs->has_cov = true; fs->dim = 2;
  std::vector<double> pts(3 * n);
  for (int64_t i = 0; i < n; ++i) {
    Feature& f = fs->base[i];
    f.loc[0] = loc[i]; f.loc[1] = loc[n + i]; f.loc[2] = 0.0;
    for (int d = 0; d < 3; ++d) pts[3 * i + d] = f.loc[d];
    f.strength = strength ? strength[i] : 0.0; f.strength_f = (float)f.strength;
    for (int k = 0; k < 9; ++k) { f.P[k] = 0.0; f.cov[k] = 0.0; }
    f.cov[0] = cov[i]; f.cov[1] = f.cov[3] = cov[n + i]; f.cov[4] = cov[2 * n + i];
    double P3[3]; projector2_from_cov(cov[i], cov[n + i], cov[2 * n + i], P3);
    f.P[0] = P3[0]; f.P[1] = f.P[3] = P3[1]; f.P[4] = P3[2];
    f.shape = CORNER; f.ndirs = 0;
    for (int a = 0; a < 3; ++a) for (int d = 0; d < 3; ++d) f.dirs[a][d] = 0.0;
    f.id = (int)i;
  }
  fs->kd = std::make_shared<KdTree>(); fs->kd->build(pts);
  *out = fs; return LRR_OK;
}
void lrro_features_destroy(lrro_fset* fs) { delete fs; }
int64_t lrro_features_size(const lrro_fset* fs) { return (int64_t)fs->base.size(); }
int lrro_features_nearest(lrro_fset* fs, int64_t nq, const double* q, int32_t* idx) {
  for (int64_t i = 0; i < nq; ++i) { double p[3] = {q[i], q[nq + i], q[2 * nq + i]}; idx[i] = fs->kd->nearest(p); }
  return LRR_OK;
}
// FeaturesInRegions for one volume (rrl/itkLocationRegistration.cxx:1363-1522), restated literally: features inside the cube with their
// region labels (:1398-1412), std::sort by distance from the reference point (:1297-1311,1460-1464; ties: lowest index — the
// reference leaves them to std::sort), labels in the order of their first appearance (:1478-1483), all features of a label added,
// label after label, while the set holds fewer than min_count (:1487-1495).
int lrro_features_set_labels(lrro_fset* fs, const int32_t* labels) { fs->labels.assign(labels, labels + fs->base.size()); return LRR_OK; }
int lrro_features_in_regions(lrro_fset* fs, const double* x0, const double* x1, const double* ref, int64_t min_count, int64_t cap, int32_t* idx_out, int64_t* n_out) {
  if (fs->labels.size() != fs->base.size()) return LRR_ERR_STATE;
  struct Item { double d2; int32_t idx; int32_t label; };
  std::vector<Item> in_roi;
  for (size_t i = 0; i < fs->base.size(); ++i) {
    const double* c = fs->base[i].loc;
    bool inside = true;
    for (int d = 0; d < 3; ++d) if (x0[d] > c[d] || c[d] > x1[d]) inside = false;
    if (!inside) continue;
    const double dx = c[0] - ref[0], dy = c[1] - ref[1], dz = c[2] - ref[2];
    in_roi.push_back(Item{dx * dx + dy * dy + dz * dz, (int32_t)i, fs->labels[i]});
  }
  std::sort(in_roi.begin(), in_roi.end(), [](const Item& a, const Item& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); });
  std::set<int32_t> labels_set; std::vector<int32_t> labels;
  for (auto& it : in_roi) if (labels_set.insert(it.label).second) labels.push_back(it.label);
  std::vector<int32_t> inside;
  auto label_it = labels.begin();
  while ((int64_t)inside.size() < min_count && label_it != labels.end()) {
    for (auto& it : in_roi) if (it.label == *label_it) inside.push_back(it.idx);
    ++label_it;
  }
  *n_out = (int64_t)inside.size();
  for (int64_t i = 0; i < (int64_t)inside.size() && i < cap; ++i) idx_out[i] = inside[(size_t)i];
  return LRR_OK;
}
// test hook: brute-force NN (cross-check of the kd-tree, K-f)
int lrro_features_nearest_brute(lrro_fset* fs, int64_t nq, const double* q, int32_t* idx) {
  for (int64_t i = 0; i < nq; ++i) { double p[3] = {q[i], q[nq + i], q[2 * nq + i]}; idx[i] = fs->kd->nearest_brute(p); }
  return LRR_OK;
}

static void rebuild(lrro_est* e, TransP trans) {
  auto mi = e->moving_inside_idx.empty() ? clone_set(e->moving_all->base) : clone_subset(e->moving_all->base, e->moving_inside_idx);
  if (!e->is_sym()) {
    e->fwd = std::make_shared<FwdEst>(e->flavor(), mi, e->fixed_all_f, trans, e->fixed_all->kd);
  } else {
    auto fi = e->fixed_inside_idx.empty() ? clone_set(e->fixed_all->base) : clone_subset(e->fixed_all->base, e->fixed_inside_idx);
    e->sym = std::make_shared<SymEst>(e->flavor(), e->moving_all_f, e->fixed_all_f, trans, trans->inverse(), mi, fi, e->fixed_all->kd, e->moving_all->kd);
  }
  apply_policy(e);
}

int lrro_est_create(lrro_ctx* ctx, int kind, int dim, int dof, lrro_fset* moving_all, lrro_fset* fixed_all, lrro_est** out) {
  if (!ctx || !moving_all || !fixed_all || !out) return LRR_ERR_INVALID;
  if (kind < 0 || kind > 4) return LRR_ERR_INVALID;
  if (kind == LRR_CDCL_ICP) { if (dim != 2 || dof != 4 || moving_all->dim != 2 || fixed_all->dim != 2) return LRR_ERR_UNSUPPORTED; }
  else {
    if (dim != 3 || (dof != 12 && dof != 6) || moving_all->dim != 3 || fixed_all->dim != 3) return LRR_ERR_UNSUPPORTED;
    if (dof == 6 && (kind == LRR_RRL_FWD || kind == LRR_RRL_SYM)) return LRR_ERR_UNSUPPORTED;  // Q6
  }
  auto* e = new lrro_est(); e->ctx = ctx; e->kind = kind; e->dof = dof; e->moving_all = moving_all; e->fixed_all = fixed_all;
  e->moving_all_f = clone_set(moving_all->base); e->fixed_all_f = clone_set(fixed_all->base);
  rebuild(e, std::make_shared<Trans>(e->trans_kind()));
  *out = e; return LRR_OK;
}
void lrro_est_destroy(lrro_est* e) { delete e; }

int lrro_est_set_subsets(lrro_est* e, const int32_t* mi, int64_t nm, const int32_t* fi, int64_t nf) {
  e->moving_inside_idx.assign(mi, mi + nm);
  if (fi) e->fixed_inside_idx.assign(fi, fi + nf); else e->fixed_inside_idx.clear();
  auto mv = clone_subset(e->moving_all->base, e->moving_inside_idx);
  if (!e->is_sym()) { e->fwd->set_moving_data(mv); }
  else {
    auto fv = e->fixed_inside_idx.empty() ? clone_set(e->fixed_all->base) : clone_subset(e->fixed_all->base, e->fixed_inside_idx);
    e->sym->set_features(mv, fv);
  }
  return LRR_OK;
}
int lrro_est_set_transform(lrro_est* e, int tk, const double* A9, const double* t3, const double* c3, const double* cov) {
  if ((tk == LRR_TRANS_AFFINE) != (e->dof == 12) || (tk == LRR_TRANS_SIMILARITY2D) != (e->dof == 4)) return LRR_ERR_INVALID;
  auto t = std::make_shared<Trans>(e->trans_kind());
  if (tk == LRR_TRANS_SIMILARITY2D) {   // (A = s R, t) -> (s, ang, t): ctor #2 of cdcl_trans_similarity2d
    t->sim2d_set(std::sqrt(A9[0] * A9[0] + A9[3] * A9[3]), std::atan2(A9[3], A9[0]), t3[0], t3[1]);
  } else {
    std::memcpy(t->A, A9, sizeof(double) * 9); std::memcpy(t->t, t3, sizeof(double) * 3);
    if (c3) std::memcpy(t->c, c3, sizeof(double) * 3);
  }
  if (cov) t->covar.assign(cov, cov + e->dof * e->dof);
  if (e->is_sym()) e->sym->set_transform(t); else e->fwd->set_transform(t);
  return LRR_OK;
}
int lrro_est_set_flags(lrro_est* e, int normalize, int wbs) {
  for (int d = 0; d < 2; ++d) if (FwdEst* f = e->dir(d)) { f->normalize_matches_ = normalize != 0; f->weight_by_strength_ = wbs != 0; }
  if (e->is_sym()) e->sym->normalize_matches_ = normalize != 0;
  return LRR_OK;
}
int lrro_est_set_order(lrro_est* e, int d, const int32_t* perm, int64_t n) {
  FwdEst* f = e->dir(d); if (!f || (size_t)n != f->moving_base_.size()) return LRR_ERR_INVALID;
  f->set_order(perm, (size_t)n); return LRR_OK;
}
int lrro_est_set_shuffle(lrro_est* e, int mode, uint64_t seed) { e->shuffle_mode = mode; e->seed = seed; apply_policy(e); return LRR_OK; }
int lrro_est_set_initial_scale(lrro_est* e, int d, double s) { FwdEst* f = e->dir(d); if (!f) return LRR_ERR_INVALID; f->initial_scale = s; return LRR_OK; }
int lrro_est_set_number_points(lrro_est* e, int d, int64_t m, int* at_max) {
  FwdEst* f = e->dir(d); if (!f || e->kind == LRR_CDCL_ICP) return LRR_ERR_INVALID;   // (cdcl_estimation_ICP has no point ramp)
  bool r = f->set_number_points((size_t)m); if (at_max) *at_max = r; return LRR_OK;
}
int lrro_est_find_closest(lrro_est* e) { if (e->is_sym()) e->sym->find_closest_euclidean(); else e->fwd->find_closest_euclidean(); return LRR_OK; }
int lrro_est_one_iteration(lrro_est* e, unsigned it, int* conv) {
  bool c = e->is_sym() ? e->sym->one_iteration(it) : e->fwd->one_iteration();
  if (conv) *conv = c; return LRR_OK;
}
int lrro_est_run(lrro_est* e, int max_iter, int* conv, int* iters) {
  bool c = e->is_sym() ? e->sym->run(max_iter, iters) : e->fwd->run(max_iter, iters);
  if (conv) *conv = c; return LRR_OK;
}
int lrro_est_get_transform(lrro_est* e, int d, int* tk, double* A9, double* t3, double* c3, double* cov) {
  FwdEst* f = e->dir(d); if (!f) return LRR_ERR_INVALID;
  const Trans& t = *f->trans_;
  if (tk) *tk = t.kind == TRANS_AFFINE ? LRR_TRANS_AFFINE : (t.kind == TRANS_RIGID3D ? LRR_TRANS_RIGID3D : LRR_TRANS_SIMILARITY2D);
  if (A9) std::memcpy(A9, t.A, 72); if (t3) std::memcpy(t3, t.t, 24); if (c3) std::memcpy(c3, t.c, 24);
  if (cov) std::memcpy(cov, t.covar.data(), sizeof(double) * t.covar.size());
  return LRR_OK;
}
int lrro_est_get_matches(lrro_est* e, int d, int64_t cap, int64_t* n, int32_t* from_idx, int32_t* to_idx, double* w) {
  FwdEst* f = e->dir(d); if (!f) return LRR_ERR_INVALID;
  int64_t m = (int64_t)f->matches_.size(); if (n) *n = m;
  for (int64_t i = 0; i < m && i < cap; ++i) {
    if (from_idx) from_idx[i] = f->matches_[i].from->id;
    if (to_idx) to_idx[i] = f->matches_[i].to->id;
    if (w) w[i] = f->matches_[i].w;
  }
  return LRR_OK;
}
int lrro_est_metric(lrro_est* e, int which, int d, double* out) {
  FwdEst* f = e->dir(d); if (!f) return LRR_ERR_INVALID;
  switch (which) {
    case LRR_METRIC_RMS: *out = f->RMS_error(); break;
    case LRR_METRIC_WEIGHTED: *out = f->weighted_error(); break;
    case LRR_METRIC_MEDIAN: *out = f->median_error(); break;
    case LRR_METRIC_SHEET_ANGLES: *out = f->sheet_angles(); break;
    case LRR_METRIC_TUBE_ANGLES: *out = f->tube_angles(); break;
    default: return LRR_ERR_INVALID;
  }
  return LRR_OK;
}
int lrro_est_estimate_ls_backward(lrro_est* e, int ep) { if (!e->is_sym()) return LRR_ERR_INVALID; e->sym->estimate_LS_backward(ep != 0); return LRR_OK; }
int lrro_est_transfer_cov(lrro_est* e, int d, double* S9) {
  FwdEst* f = e->dir(d); if (!f) return LRR_ERR_INVALID;
  f->compute_transfer_error_covariance(); std::memcpy(S9, f->trans_->covarJ, 72); return LRR_OK;
}
// (instrumentation of the GPU neighbour search: nothing to report on the CPU)
int lrro_est_get_stats(lrro_est*, lrr_est_stats* s) { std::memset(s, 0, sizeof(*s)); return LRR_OK; }
int lrro_est_get_state(lrro_est* e, lrr_est_state* s) {
  std::memset(s, 0, sizeof(*s));
  for (int d = 0; d < 2; ++d) if (FwdEst* f = e->dir(d)) {
    s->number_points[d] = (int64_t)f->number_points_; s->scale[d] = f->last_scale; s->scale_estimated[d] = f->scale_estimated_;
  }
  FwdEst* f0 = e->dir(0);
  for (int k = 0; k < 3; ++k) { s->norm_center_moving[k] = f0->last_norm[k]; s->norm_center_fixed[k] = f0->last_norm[4 + k]; }
  s->norm_radius_moving = f0->last_norm[3]; s->norm_radius_fixed = f0->last_norm[7];
  if (e->is_sym()) {
    s->number_matches = e->sym->number_matches_; s->finest_level = e->sym->finest_level_;
    s->oscillation_count = (int)e->sym->oscillation_count_; s->error_difference = e->sym->error_difference_;
    s->weighted_error = e->sym->weighted_error_; s->iterations_done = e->sym->iterations_done;
  }
  return LRR_OK;
}
int lrro_est_oscillated(lrro_est* e, int* out) { *out = e->is_sym() ? e->sym->oscillated() : 0; return LRR_OK; }

int lrro_objcov_eval(lrro_est* e, int d, const double* x, double* f, double* g) {
  FwdEst* fe = e->dir(d); if (!fe) return LRR_ERR_INVALID;
  std::vector<double> saved = fe->trans_->covar;
  ObjFunCov obj(fe->matches_, *fe->trans_);
  obj.compute(x, f, g);
  fe->trans_->covar = saved;
  return LRR_OK;
}
// estimate_covariance (cdcl/cdcl_estimation.txx:264-337) on the ICP matches of direction d.
int lrro_objcov_refine(lrro_est* e, int d, int max_evals, double* f_start, double* f_end, int* n_evals) {
  FwdEst* fe = e->dir(d); if (!fe) return LRR_ERR_INVALID;
  Trans& tr = *fe->trans_; const int dof = tr.dof;
  double cm[3] = {0, 0, 0}, cf[3] = {0, 0, 0}, rm = 0, rf = 0;
  if (fe->normalize_matches_) { fe->normalize_matches_fn(cm, rm, cf, rf); tr.normalize(rm, rf, cm, cf); }
  ObjFunCov obj(fe->matches_, tr);
  std::vector<double> U((size_t)dof * dof);
  if (!cholesky_upper(dof, tr.covar.data(), U.data())) {
    if (fe->normalize_matches_) { tr.unnormalize(rm, rf, cm, cf); fe->unnormalize_matches_fn(cm, rm, cf, rf); }
    return LRR_ERR_STATE;
  }
  std::vector<double> x(dof * (dof + 1) / 2); int ind = 0;
  for (int d1 = 0; d1 < dof; ++d1) for (int d2 = d1; d2 < dof; ++d2) x[ind++] = U[d1 * dof + d2];
  double fs = 0, fe_ = 0;
  int ev = lbfgs_minimize([&](const double* xx, double* f, double* g) { obj.compute(xx, f, g); }, x, max_evals, &fs, &fe_);
  std::vector<double> up((size_t)dof * dof, 0.0); ind = 0;
  for (int d1 = 0; d1 < dof; ++d1) for (int d2 = d1; d2 < dof; ++d2) up[d1 * dof + d2] = x[ind++];
  for (int i = 0; i < dof; ++i) for (int j = 0; j < dof; ++j) { double s = 0; for (int k = 0; k < dof; ++k) s += up[k * dof + i] * up[k * dof + j]; tr.covar[i * dof + j] = s; }
  if (fe->normalize_matches_) { tr.unnormalize(rm, rf, cm, cf); fe->unnormalize_matches_fn(cm, rm, cf, rf); }
  if (f_start) *f_start = fs; if (f_end) *f_end = fe_; if (n_evals) *n_evals = ev;
  return LRR_OK;
}

// cdcl_obj_fun_par<3,12>::compute over the current matches / weights of direction d, at the parameter vector theta (row-major A, then
// t; cdcl_trans_affine.txx:65-106), with the transfer-error covariances frozen at the direction's current transform.  The transform
// itself is left as it was.
int lrro_objpar_eval(lrro_est* e, int d, const double* theta, double* f, double* g) {
  FwdEst* fe = e->dir(d); if (!fe || fe->trans_->dof != 12) return LRR_ERR_UNSUPPORTED;
  Trans saved = *fe->trans_;
  ObjFunPar obj(fe->matches_, *fe->trans_); obj.n_objects = (double)fe->cdc_rows;
  obj.compute(theta, f, g);
  *fe->trans_ = saved;
  return LRR_OK;
}
// estimate_parameters (cdcl/cdcl_estimation.txx:207-262) on the ICP matches of direction d: normalise, L-BFGS over theta (<= max_evals
// evaluations, 5 % early stop of cdcl_lbfgs), set_parameterization, unnormalise.  converged = (f_start - f_end) < 1e-4 (:239).
int lrro_objpar_refine(lrro_est* e, int d, int max_evals, double* f_start, double* f_end, int* n_evals, int* converged) {
  FwdEst* fe = e->dir(d); if (!fe || fe->trans_->dof != 12) return LRR_ERR_UNSUPPORTED;
  Trans& tr = *fe->trans_;
  double cm[3] = {0, 0, 0}, cf[3] = {0, 0, 0}, rm = 0, rf = 0;
  if (fe->normalize_matches_) { fe->normalize_matches_fn(cm, rm, cf, rf); tr.normalize(rm, rf, cm, cf); }
  ObjFunPar obj(fe->matches_, tr); obj.n_objects = (double)fe->cdc_rows;
  std::vector<double> x(12);
  for (int i = 0; i < 9; ++i) x[i] = tr.A[i];
  for (int i = 0; i < 3; ++i) x[9 + i] = tr.t[i];
  double fs = 0, fe_ = 0;
  int ev = lbfgs_minimize([&](const double* xx, double* f, double* g) { obj.compute(xx, f, g); }, x, max_evals, &fs, &fe_);
  tr.set_parameterization(x.data());
  if (fe->normalize_matches_) { tr.unnormalize(rm, rf, cm, cf); fe->unnormalize_matches_fn(cm, rm, cf, rf); }
  if (f_start) *f_start = fs; if (f_end) *f_end = fe_; if (n_evals) *n_evals = ev; if (converged) *converged = (fs - fe_) < 1e-4;
  return LRR_OK;
}

// ---- CDC E-step (cdcl_estimation<3,12>::compute_weights, cdcl/cdcl_estimation.txx:342-545) ----------------------------------------------
int lrro_cdc_compute_weights(lrro_est* e, int d, int max_matches, double eval_multiplier, lrr_cdc_summary* out) {
  FwdEst* fe = e->dir(d); if (!fe || fe->trans_->dof != 12) return LRR_ERR_UNSUPPORTED;
  if (max_matches <= 0 || !(eval_multiplier > 0)) return LRR_ERR_INVALID;
  FwdEst::CdcSummary s = fe->cdc_compute_weights((unsigned)max_matches, eval_multiplier);
  if (out) { out->rows = (int64_t)s.rows; out->entries = (int64_t)s.entries; out->min_matches = (int32_t)s.min_matches; out->max_matches = (int32_t)s.max_matches; out->avg_largest_eval = s.avg_largest_eval; }
  return LRR_OK;
}
int lrro_cdc_get_matches(lrro_est* e, int d, int64_t cap_rows, int64_t cap_entries, int64_t* n_rows, int64_t* n_entries, int32_t* row_from,
                         int64_t* row_start, int32_t* to_idx, double* w) {
  FwdEst* fe = e->dir(d); if (!fe) return LRR_ERR_INVALID;
  if (fe->cdc_row_of.empty()) return LRR_ERR_STATE;
  if (n_rows) *n_rows = (int64_t)fe->cdc_rows; if (n_entries) *n_entries = (int64_t)fe->matches_.size();
  if (!row_from && !row_start && !to_idx && !w) return LRR_OK;
  if (cap_rows < (int64_t)fe->cdc_rows || cap_entries < (int64_t)fe->matches_.size()) return LRR_ERR_INVALID;
  for (size_t k = 0; k < fe->matches_.size(); ++k) {
    if (fe->first_of_row(k)) { const int r = fe->cdc_row_of[k]; if (row_from) row_from[r] = fe->matches_[k].from->id; if (row_start) row_start[r] = (int64_t)k; }
    if (to_idx) to_idx[k] = fe->matches_[k].to->id;
    if (w) w[k] = fe->matches_[k].w;
  }
  if (row_start) row_start[fe->cdc_rows] = (int64_t)fe->matches_.size();
  return LRR_OK;
}
int lrro_cdc_resolution_switch(double avg_largest_eval, const double* fixed_spacing, int res, int max_res, int allowed, int* new_res) {
  if (!fixed_spacing || !new_res || res < 0 || res > max_res) return LRR_ERR_INVALID;
  int r = res; bool not_restarted = true;
  if (avg_largest_eval < 2.0 * fixed_spacing[r] && allowed && not_restarted) { if (r > 0) { --r; not_restarted = false; } }
  if (r + 1 <= max_res && avg_largest_eval > 2.0 * fixed_spacing[r + 1] && allowed && not_restarted) { ++r; }
  *new_res = r; return LRR_OK;
}
int lrro_cdc_one_iteration(lrro_est* e, int d, unsigned iteration, int max_matches, double eval_multiplier, int* converged, lrr_cdc_summary* out) {
  FwdEst* fe = e->dir(d); if (!fe || fe->trans_->dof != 12) return LRR_ERR_UNSUPPORTED;
  lrr_cdc_summary s; int rc;
  if ((rc = lrro_cdc_compute_weights(e, d, max_matches, eval_multiplier, &s)) != LRR_OK) return rc;
  if (s.entries == 0) return LRR_ERR_STATE;
  int par_conv = 0; double fs = 0, fe_ = 0; int ev = 0;
  if (iteration > 5) {
    if ((rc = lrro_objpar_refine(e, d, 5, &fs, &fe_, &ev, &par_conv)) != LRR_OK) return rc;
    if ((rc = lrro_cdc_compute_weights(e, d, max_matches, eval_multiplier, &s)) != LRR_OK) return rc;
    if (s.entries == 0) return LRR_ERR_STATE;
  }
  if ((rc = lrro_objcov_refine(e, d, 5, &fs, &fe_, &ev)) != LRR_OK) return rc;
  const int cov_conv = (fs - fe_) < 1e-5;
  double fro = 0; for (double v : fe->trans_->covar) fro += v * v;
  if (converged) *converged = par_conv && cov_conv && std::sqrt(fro) < 1e-3;
  if (out) *out = s;
  return LRR_OK;
}

// validTransform (rrl/itkLocationRegistration.cxx:2914-2926)
static bool valid_transform(const Trans& t) {
  double w[3]; svd_singular_values(3, t.A, w);
  for (int d = 0; d < 3; ++d) if (w[d] > 10.0 || w[d] < 0.1) return false;
  return true;
}

static void run_one_job(const lrro_fset* moving_all, const lrro_fset* fixed_all, const std::vector<FeatureP>& mall,
                        const std::vector<FeatureP>& fall, std::shared_ptr<const KdTree> kd_fixed, std::shared_ptr<const KdTree> kd_moving,
                        const lrr_job& job, int max_iter, uint64_t seed, lrr_result& r) {
  std::memset(&r, 0, sizeof(r));
  auto in_cube = [&](const Feature& f, const double* c) {
    for (int d = 0; d < 3; ++d) if (c[d] - job.roi_half_size > f.loc[d] || f.loc[d] > c[d] + job.roi_half_size) return false;
    return true;
  };
  std::vector<FeatureP> mi, fi;   // FeaturesInROIs (rrl/itkLocationRegistration.cxx:1324-1360)
  for (auto& f : fixed_all->base) if (in_cube(f, job.roi_center_fixed)) fi.push_back(std::make_shared<Feature>(f));
  for (auto& f : moving_all->base) if (in_cube(f, job.roi_center_moving)) mi.push_back(std::make_shared<Feature>(f));
  r.n_moving_inside = (int)mi.size(); r.n_fixed_inside = (int)fi.size();
  if (mi.size() < 1000 || fi.size() < 1000) { r.status = 1; return; }   // :1632
  auto t = std::make_shared<Trans>(TRANS_AFFINE);
  std::memcpy(t->A, job.A, 72); std::memcpy(t->t, job.t, 24); std::memcpy(t->c, job.c, 24);
  // prebuilt kd-trees are shared: the driver keeps one estimator alive across locations (:1635-1638)
  SymEst est(FLAVOR_RRL, mall, fall, t, t->inverse(), mi, fi, kd_fixed, kd_moving);
  est.forward_->shuffle_mode = est.backward_->shuffle_mode = 1;
  est.forward_->shuffle_seed = est.backward_->shuffle_seed = seed; est.forward_->shuffle_dir = 0; est.backward_->shuffle_dir = 1;
  est.set_transform(t);
  bool converged = false; int it = 0;
  while (!converged && it < max_iter) {                               // :2418
    converged = est.one_iteration((unsigned)it); ++it;
    if (!valid_transform(*est.forward_->trans_)) { r.status = 2; break; }   // :2482
  }
  const Trans& T = *est.forward_->trans_;
  r.converged = converged; r.oscillated = est.oscillated(); r.iterations = it;
  std::memcpy(r.A, T.A, 72); std::memcpy(r.t, T.t, 24); std::memcpy(r.c, T.c, 24);
  std::memcpy(r.cov, T.covar.data(), sizeof(double) * 144);
  r.rms = est.forward_->RMS_error(); r.weighted_error = est.forward_->weighted_error();
}

// ---- the driver's candidate loop (rrl/itkLocationRegistration.cxx:2230-2332, 1816-1871) -----------------------------------------
// ComputeTransferErrorCovariance (:2793-2858): max over a 5 mm grid of the ROI of trace( J(x) Sigma J(x)^T ).  The ROI region of the
// driver is an image region; here start = centre - half, size = 2 half, end = start + size - spacing.
static double transfer_max_trace(const Trans& tr, const double* centre, double half, double spacing) {
  const int dof = tr.dof;
  const double interval = 5.0;
  double start[3], end[3];
  for (int d = 0; d < 3; ++d) { start[d] = centre[d] - half; end[d] = start[d] + 2.0 * half - spacing; }
  double maxTrace = 0.0;
  std::vector<double> J(3 * dof);
  for (double k = start[2]; k <= end[2]; k += interval)
    for (double j = start[1]; j <= end[1]; j += interval)
      for (double i = start[0]; i <= end[0]; i += interval) {
        const double p[3] = {i, j, k};
        tr.jacobian_wrt_par(p, J.data());
        double trc = 0.0;
        for (int d = 0; d < 3; ++d) {
          double s = 0.0;
          for (int a = 0; a < dof; ++a) { double v = 0.0; for (int b = 0; b < dof; ++b) v += tr.covar[a * dof + b] * J[d * dof + b]; s += J[d * dof + a] * v; }
          trc += s;
        }
        if (trc > maxTrace) maxTrace = trc;
      }
  return maxTrace;
}
static const uint64_t kRankStream = 1ULL << 40;   // the ranking find_closest draws its permutation from its own stream (Q4)

// stage 0: ranking angles only; stage 1: the registration (+ decision measurements)
static void run_candidate(const lrro_fset* moving_all, const lrro_fset* fixed_all, const std::vector<FeatureP>& mall, const std::vector<FeatureP>& fall,
                          const lrr_job& job, int max_iter, uint64_t seed, double spacing, int stage, lrr_result& r, lrr_decision& dec) {
  auto in_cube = [&](const Feature& f, const double* c) {
    for (int d = 0; d < 3; ++d) if (c[d] - job.roi_half_size > f.loc[d] || f.loc[d] > c[d] + job.roi_half_size) return false;
    return true;
  };
  std::vector<FeatureP> mi, fi;
  for (auto& f : fixed_all->base) if (in_cube(f, job.roi_center_fixed)) fi.push_back(std::make_shared<Feature>(f));
  for (auto& f : moving_all->base) if (in_cube(f, job.roi_center_moving)) mi.push_back(std::make_shared<Feature>(f));
  if (stage == 0) { std::memset(&r, 0, sizeof(r)); std::memset(&dec, 0, sizeof(dec)); dec.rank = -1; }
  r.n_moving_inside = (int)mi.size(); r.n_fixed_inside = (int)fi.size();
  if (mi.size() < 1000 || fi.size() < 1000) { r.status = 1; return; }
  auto t = std::make_shared<Trans>(TRANS_AFFINE);
  std::memcpy(t->A, job.A, 72); std::memcpy(t->t, job.t, 24); std::memcpy(t->c, job.c, 24);
  SymEst est(FLAVOR_RRL, mall, fall, t, t->inverse(), mi, fi, fixed_all->kd, moving_all->kd);
  est.forward_->shuffle_mode = est.backward_->shuffle_mode = 1;
  est.forward_->shuffle_seed = est.backward_->shuffle_seed = seed; est.forward_->shuffle_dir = 0; est.backward_->shuffle_dir = 1;
  est.set_transform(t);
  const double deg = 180.0 / 3.14159265358979323846;
  if (stage == 0) {   // :2303-2313
    est.forward_->shuffle_calls = est.backward_->shuffle_calls = kRankStream;
    est.find_closest_euclidean();
    dec.rank_angles[0] = est.forward_->sheet_angles() * deg; dec.rank_angles[1] = est.forward_->tube_angles() * deg;
    dec.rank_angles[2] = est.backward_->sheet_angles() * deg; dec.rank_angles[3] = est.backward_->tube_angles() * deg;
    dec.rank_angle_average = (dec.rank_angles[0] + dec.rank_angles[1] + dec.rank_angles[2] + dec.rank_angles[3]) / 4;
    return;
  }
  bool converged = false; int it = 0;
  while (!converged && it < max_iter) {
    converged = est.one_iteration((unsigned)it); ++it;
    if (!valid_transform(*est.forward_->trans_)) { r.status = 2; break; }
  }
  const Trans& T = *est.forward_->trans_;
  r.converged = converged; r.oscillated = est.oscillated(); r.iterations = it;
  std::memcpy(r.A, T.A, 72); std::memcpy(r.t, T.t, 24); std::memcpy(r.c, T.c, 24);
  std::memcpy(r.cov, T.covar.data(), sizeof(double) * 144);
  r.rms = est.forward_->RMS_error(); r.weighted_error = est.forward_->weighted_error();
  dec.ran = 1; dec.verified = (r.status == 0 && r.converged && !r.oscillated) ? 1 : 0;
  if (r.status == 0 && converged && !est.oscillated()) {   // :2495-2507 -> ComputeDecisionMeasurements :1828-1860
    dec.sheet_angles[0] = est.forward_->sheet_angles() * deg; dec.tube_angles[0] = est.forward_->tube_angles() * deg;
    dec.weighted_error[0] = est.forward_->weighted_error();
    dec.max_trace[0] = transfer_max_trace(*est.forward_->trans_, job.roi_center_moving, job.roi_half_size, spacing);
    est.estimate_LS_backward(false);
    dec.max_trace[1] = transfer_max_trace(*est.backward_->trans_, job.roi_center_fixed, job.roi_half_size, spacing);
    dec.sheet_angles[1] = est.backward_->sheet_angles() * deg; dec.tube_angles[1] = est.backward_->tube_angles() * deg;
    dec.weighted_error[1] = est.backward_->weighted_error();
    const double* q[4] = {dec.weighted_error, dec.max_trace, dec.sheet_angles, dec.tube_angles};
    for (int k = 0; k < 4; ++k) { dec.measurements[k] = std::max(q[k][0], q[k][1]); dec.measurements[4 + k] = std::min(q[k][0], q[k][1]); }
    dec.measured = 1;
  }
}
int lrro_batch_run_locations(lrro_ctx*, lrro_fset* moving_all, lrro_fset* fixed_all, const lrr_job* jobs, int64_t n_locations, int32_t n_candidates,
                             int max_iter, uint64_t shuffle_seed, int until_verified, double voxel_spacing, lrr_result* out, lrr_decision* dec) {
  auto mall = clone_set(moving_all->base); auto fall = clone_set(fixed_all->base);
  for (int64_t l = 0; l < n_locations; ++l) {
    std::vector<std::pair<double, int> > cand;
    for (int c = 0; c < n_candidates; ++c) {
      const int64_t j = l * n_candidates + c;
      run_candidate(moving_all, fixed_all, mall, fall, jobs[j], max_iter, shuffle_seed + (uint64_t)j, voxel_spacing, 0, out[j], dec[j]);
      if (out[j].status == 0) cand.push_back(std::make_pair(dec[j].rank_angle_average, c));
    }
    std::stable_sort(cand.begin(), cand.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; });   // :2326 (ties: candidate order)
    bool verified = false;
    for (size_t k = 0; k < cand.size(); ++k) {
      const int64_t j = l * n_candidates + cand[k].second;
      dec[j].rank = (int)k;
      if (until_verified && verified) continue;   // :2332
      run_candidate(moving_all, fixed_all, mall, fall, jobs[j], max_iter, shuffle_seed + (uint64_t)j, voxel_spacing, 1, out[j], dec[j]);
      verified = verified || dec[j].verified;
    }
  }
  return LRR_OK;
}

// Batched location registration on the CPU.  nthreads <= 0 -> 1 (the reference's execution model).
int lrro_batch_run_mt(lrro_ctx*, lrro_fset* moving_all, lrro_fset* fixed_all, const lrr_job* jobs, int64_t njobs,
                      int max_iter, uint64_t shuffle_seed, lrr_result* out, int nthreads) {
  if (nthreads <= 0) nthreads = 1;
  std::atomic<int64_t> next{0};
  auto worker = [&]() {
    // per-thread *_all clones: features are mutated in place during normalisation (Q10)
    auto mall = clone_set(moving_all->base); auto fall = clone_set(fixed_all->base);
    for (;;) {
      int64_t j = next.fetch_add(1); if (j >= njobs) break;
      run_one_job(moving_all, fixed_all, mall, fall, fixed_all->kd, moving_all->kd, jobs[j], max_iter, shuffle_seed + (uint64_t)j, out[j]);
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  return LRR_OK;
}
int lrro_batch_run(lrro_ctx* c, lrro_fset* m, lrro_fset* f, const lrr_job* jobs, int64_t njobs, int max_iter, uint64_t seed, lrr_result* out) {
  return lrro_batch_run_mt(c, m, f, jobs, njobs, max_iter, seed, out, 1);
}
// shard of a larger list: job i is seeded with seed + its global index
int lrro_batch_run_ids(lrro_ctx*, lrro_fset* moving_all, lrro_fset* fixed_all, const lrr_job* jobs, const int64_t* ids, int64_t njobs, int max_iter,
                       uint64_t seed, lrr_result* out) {
  auto mall = clone_set(moving_all->base); auto fall = clone_set(fixed_all->base);
  for (int64_t j = 0; j < njobs; ++j)
    run_one_job(moving_all, fixed_all, mall, fall, fixed_all->kd, moving_all->kd, jobs[j], max_iter, seed + (uint64_t)ids[j], out[j]);
  return LRR_OK;
}
// expected cost of a job = the features inside its two ROI cubes (FeaturesInROIs, rrl/itkLocationRegistration.cxx:1324-1360), 0 when
// a cube fails the 1000-feature gate (:1632).  (The CUDA backend counts by whole grid cells: a scheduling heuristic, not a result.)
int lrro_batch_expected_cost(lrro_ctx*, lrro_fset* moving_all, lrro_fset* fixed_all, const lrr_job* jobs, int64_t njobs, int32_t* cost) {
  auto inside = [](const std::vector<Feature>& v, const double* c, double h) {
    int n = 0;
    for (const Feature& f : v) if (!(c[0] - h > f.loc[0] || f.loc[0] > c[0] + h || c[1] - h > f.loc[1] || f.loc[1] > c[1] + h || c[2] - h > f.loc[2] || f.loc[2] > c[2] + h)) ++n;
    return n;
  };
  for (int64_t j = 0; j < njobs; ++j) {
    const int cm = inside(moving_all->base, jobs[j].roi_center_moving, jobs[j].roi_half_size), cf = inside(fixed_all->base, jobs[j].roi_center_fixed, jobs[j].roi_half_size);
    cost[j] = (cm < 1000 || cf < 1000) ? 0 : cm + cf;
  }
  return LRR_OK;
}
// "devices" of the CPU checker = host threads over the first replica
int lrro_batch_run_multi(lrro_ctx** c, int ndev, lrro_fset** m, lrro_fset** f, const lrr_job* jobs, int64_t njobs, int max_iter, uint64_t seed, lrr_result* out) {
  return lrro_batch_run_mt(c[0], m[0], f[0], jobs, njobs, max_iter, seed, out, ndev);
}

// ---- test hooks -------------------------------------------------------------------------------
double lrro_test_muse_scale(const double* r, int64_t n) { return muse_scale(std::vector<double>(r, r + n)); }
double lrro_test_mad_scale(const double* r, int64_t n) { std::vector<double> v(r, r + n); return median_abs_dev_scale(v); }
double lrro_test_weight_BT(double u) { return weight_BT(u); }
double lrro_test_rho(double u) { return rho(u); }
double lrro_test_rho_p(double u) { return rho_p(u); }
double lrro_test_weight(double u) { return weight_rho(u); }
void lrro_test_svd_inverse(int n, const double* M, double* inv) { svd_inverse(n, M, inv); }
double lrro_test_gauss_inv_cdf(double p) { return gauss_inv_cdf(p); }
// normalize followed by unnormalize of (A,t,c,cov) — K-d
void lrro_test_trans_roundtrip(int kind, double* A9, double* t3, double* c3, double* cov, double rm, double rf,
                               const double* cm, const double* cf, int do_unnormalize) {
  Trans t(kind == LRR_TRANS_AFFINE ? TRANS_AFFINE : TRANS_RIGID3D);
  std::memcpy(t.A, A9, 72); std::memcpy(t.t, t3, 24); std::memcpy(t.c, c3, 24);
  if (cov) t.covar.assign(cov, cov + t.dof * t.dof);
  t.normalize(rm, rf, cm, cf);
  if (do_unnormalize) t.unnormalize(rm, rf, cm, cf);
  std::memcpy(A9, t.A, 72); std::memcpy(t3, t.t, 24); std::memcpy(c3, t.c, 24);
  if (cov) std::memcpy(cov, t.covar.data(), sizeof(double) * t.dof * t.dof);
}
void lrro_test_trans_inverse(int kind, double* A9, double* t3, double* c3) {
  Trans t(kind == LRR_TRANS_AFFINE ? TRANS_AFFINE : TRANS_RIGID3D);
  std::memcpy(t.A, A9, 72); std::memcpy(t.t, t3, 24); std::memcpy(t.c, c3, 24);
  auto r = t.inverse();
  std::memcpy(A9, r->A, 72); std::memcpy(t3, r->t, 24); std::memcpy(c3, r->c, 24);
}
void lrro_test_projector2d(const double* C3, double* P3) { projector2_from_cov(C3[0], C3[1], C3[2], P3); }
void lrro_test_sim2d_roundtrip(double* s, double* ang, double* t2, double* cov16, double rm, double rf, const double* cm2, const double* cf2, int do_un) {
  Trans t(TRANS_SIM2D); t.sim2d_set(*s, *ang, t2[0], t2[1]);
  if (cov16) t.covar.assign(cov16, cov16 + 16);
  const double cm[3] = {cm2[0], cm2[1], 0}, cf[3] = {cf2[0], cf2[1], 0};
  t.normalize(rm, rf, cm, cf);
  if (do_un) t.unnormalize(rm, rf, cm, cf);
  *s = t.s_; *ang = t.ang_; t2[0] = t.t[0]; t2[1] = t.t[1];
  if (cov16) std::memcpy(cov16, t.covar.data(), 128);
}
void lrro_test_recompose(double* R9, double* t3, const double* dR9, const double* dt3) {
  Trans g(TRANS_RIGID3D), inc(TRANS_RIGID3D);
  std::memcpy(g.A, R9, 72); std::memcpy(g.t, t3, 24); std::memcpy(inc.A, dR9, 72); std::memcpy(inc.t, dt3, 24);
  g.recompose_increment(inc);
  std::memcpy(R9, g.A, 72); std::memcpy(t3, g.t, 24);
}
// dense AtA/Atb over the current forward matches of a forward estimator, no weighting changes (K-e)
int lrro_test_dense_normal_equations(lrro_est* e, int d, double* AtA, double* Atb) {
  FwdEst* f = e->dir(d); if (!f) return LRR_ERR_INVALID;
  const int dof = f->trans_->dof;
  std::vector<double> A((size_t)dof * dof, 0.0), b(dof, 0.0);
  for (auto& m : f->matches_) FwdEst::accumulate_dense(*f->trans_, m.from->loc, m.to->P, m.to->loc, m.w, A, b);
  std::memcpy(AtA, A.data(), sizeof(double) * dof * dof); std::memcpy(Atb, b.data(), sizeof(double) * dof);
  return LRR_OK;
}
int lrro_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
extern "C" {
int lrro_ctx_set_profiling(lrro_ctx*, int) { return LRR_OK; }
int lrro_ctx_phase_times(lrro_ctx*, int cap, double* ms, int64_t* count) { for (int i = 0; i < cap; ++i) { ms[i] = 0; count[i] = 0; } return LRR_OK; }
}
