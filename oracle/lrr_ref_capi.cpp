// lrr_ref_capi.cpp — C ABI `lrrr_*` over the compiled REFERENCE (same entry points as include/lrr_gpu.h where the reference
// has a counterpart).  TEST INFRASTRUCTURE ONLY: loaded by tests/ and tests/golden/make_ref_golden.py, never by the product.
#include <cmath>
#include <cstring>
#include <string>
#include "lrr_ref_iface.hpp"

struct lrrr_ctx { std::string err; };
struct lrrr_fset { lrrr_ctx* ctx; std::vector<RefFeat> base; int dim = 3; };
struct lrrr_est {
  lrrr_ctx* ctx; int kind, dof; RefEst* impl;
  // for the CDC E-step (lrrr_cdc_*): the parent sets, the moving subset and the many-to-many estimator built on demand
  lrrr_fset *moving_all = nullptr, *fixed_all = nullptr; std::vector<int32_t> msub; bool has_sub = false; RefCdc* cdc = nullptr;
  void drop_cdc() { delete cdc; cdc = nullptr; }
};

extern "C" {
int lrrr_ctx_create(int, lrrr_ctx** out) { *out = new lrrr_ctx(); return LRR_OK; }
void lrrr_ctx_destroy(lrrr_ctx* c) { delete c; }
int lrrr_ctx_set_stream(lrrr_ctx*, void*) { return LRR_OK; }
int lrrr_ctx_synchronize(lrrr_ctx*) { return LRR_OK; }
const char* lrrr_ctx_last_error(lrrr_ctx* c) { return c->err.c_str(); }
int64_t lrrr_ctx_launch_count(lrrr_ctx*) { return 0; }
int lrrr_ctx_set_profiling(lrrr_ctx*, int) { return LRR_OK; }
int lrrr_ctx_phase_times(lrrr_ctx*, int cap, double* ms, int64_t* count) { for (int i = 0; i < cap; ++i) { ms[i] = 0; count[i] = 0; } return LRR_OK; }

int lrrr_features_upload(lrrr_ctx* ctx, int64_t n, const double* xyz, const double* P6, const double* C6, const double* strength,
                         const uint8_t* shape, const double* dirs, lrrr_fset** out) {
  if (!ctx || n < 0 || !xyz || !P6 || !out) return LRR_ERR_INVALID;
  static const int sym6[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
  lrrr_fset* fs = new lrrr_fset(); fs->ctx = ctx; fs->base.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    RefFeat& f = fs->base[i];
    for (int d = 0; d < 3; ++d) f.loc[d] = xyz[d * n + i];
    f.strength = strength ? strength[i] : 0.0;
    for (int k = 0; k < 9; ++k) { f.P[k] = P6[sym6[k] * n + i]; f.cov[k] = C6 ? C6[sym6[k] * n + i] : 0.0; }
    f.shape = shape ? shape[i] : 0;
    for (int a = 0; a < 3; ++a) for (int d = 0; d < 3; ++d) f.dirs[a][d] = dirs ? dirs[(3 * a + d) * n + i] : 0.0;
  }
  *out = fs; return LRR_OK;
}
// plain cdcl_feature<2> sets (location, strength, 2x2 covariance) for cdcl_estimation_ICP<2,4>
int lrrr_features_upload_cov(lrrr_ctx* ctx, int dim, int64_t n, const double* loc, const double* cov, const double* strength, lrrr_fset** out) {
  if (!ctx || n < 0 || !loc || !cov || !out) return LRR_ERR_INVALID;
  if (dim != 2) return LRR_ERR_UNSUPPORTED;
  lrrr_fset* fs = new lrrr_fset(); fs->ctx = ctx; fs->dim = 2; fs->base.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    RefFeat& f = fs->base[i];
    std::memset(&f, 0, sizeof f);
    f.loc[0] = loc[i]; f.loc[1] = loc[n + i];
    f.cov[0] = cov[i]; f.cov[1] = f.cov[3] = cov[n + i]; f.cov[4] = cov[2 * n + i];
    f.strength = strength ? strength[i] : 0.0;
  }
  *out = fs; return LRR_OK;
}
void lrrr_features_destroy(lrrr_fset* fs) { delete fs; }
int64_t lrrr_features_size(const lrrr_fset* fs) { return (int64_t)fs->base.size(); }

int lrrr_est_create(lrrr_ctx* ctx, int kind, int dim, int dof, lrrr_fset* moving_all, lrrr_fset* fixed_all, lrrr_est** out) {
  if (!ctx || !moving_all || !fixed_all || !out) return LRR_ERR_INVALID;
  if (kind == LRR_CDCL_ICP) {
    if (dim != 2 || dof != 4 || moving_all->dim != 2 || fixed_all->dim != 2) return LRR_ERR_UNSUPPORTED;
    lrrr_est* e = new lrrr_est(); e->ctx = ctx; e->kind = kind; e->dof = dof; e->impl = lrr_ref_make_icp2d(moving_all->base, fixed_all->base);
    *out = e; return LRR_OK;
  }
  if (dim != 3 || moving_all->dim != 3 || fixed_all->dim != 3) return LRR_ERR_UNSUPPORTED;
  RefEst* impl = (kind == LRR_CDCL_FWD || kind == LRR_CDCL_SYM) ? lrr_ref_make_cdcl(kind, dof, moving_all->base, fixed_all->base)
                                                                 : lrr_ref_make_rrl(kind, dof, moving_all->base, fixed_all->base);
  if (!impl) return LRR_ERR_UNSUPPORTED;
  lrrr_est* e = new lrrr_est(); e->ctx = ctx; e->kind = kind; e->dof = dof; e->impl = impl; e->moving_all = moving_all; e->fixed_all = fixed_all;
  *out = e; return LRR_OK;
}
void lrrr_est_destroy(lrrr_est* e) { if (e) { delete e->cdc; delete e->impl; delete e; } }
int lrrr_est_set_subsets(lrrr_est* e, const int32_t* mi, int64_t nm, const int32_t* fi, int64_t nf) {
  e->impl->set_subsets(mi, nm, fi, nf); e->drop_cdc();
  e->has_sub = mi != nullptr; e->msub.assign(mi ? mi : nullptr, mi ? mi + nm : nullptr);
  return LRR_OK;
}
int lrrr_est_set_transform(lrrr_est* e, int tk, const double* A9, const double* t3, const double* c3, const double* cov) {
  if ((tk == LRR_TRANS_AFFINE) != (e->dof == 12) || (tk == LRR_TRANS_SIMILARITY2D) != (e->dof == 4)) return LRR_ERR_INVALID;
  RefTrans t; t.kind = tk; std::memcpy(t.A, A9, 72); std::memcpy(t.t, t3, 24);
  if (c3) std::memcpy(t.c, c3, 24); else std::memset(t.c, 0, 24);
  if (cov) t.cov.assign(cov, cov + e->dof * e->dof);
  e->impl->set_transform(t); e->drop_cdc(); return LRR_OK;
}
int lrrr_est_set_flags(lrrr_est* e, int normalize, int wbs) { e->impl->set_flags(normalize, wbs); return LRR_OK; }
// the reference's order policy is vcl_random_shuffle (Q4); the stand-in routes it through a hook: mode 0 = keep the given order
int lrrr_est_set_shuffle(lrrr_est*, int mode, uint64_t seed) { lrr_ref_set_shuffle(mode, seed); return LRR_OK; }
int lrrr_est_set_number_points(lrrr_est* e, int d, int64_t m, int* at_max) { return e->impl->set_number_points(d, m, at_max) ? LRR_OK : LRR_ERR_INVALID; }
int lrrr_est_find_closest(lrrr_est* e) { e->drop_cdc(); e->impl->find_closest(); return LRR_OK; }
int lrrr_est_one_iteration(lrrr_est* e, unsigned it, int* conv) { e->drop_cdc(); bool c = e->impl->one_iteration(it); if (conv) *conv = c; return LRR_OK; }
int lrrr_est_run(lrrr_est* e, int max_iter, int* conv, int* iters) {
  bool c = false; int it = 0;
  while (!c && it < max_iter) { c = e->impl->one_iteration((unsigned)it); ++it; }   // cdcl :413-428 / driver loop rrl/itkLocationRegistration.cxx:2418
  if (conv) *conv = c; if (iters) *iters = it; return LRR_OK;
}
int lrrr_est_get_transform(lrrr_est* e, int d, int* tk, double* A9, double* t3, double* c3, double* cov) {
  RefTrans t; if (!e->impl->get_transform(d, t)) return LRR_ERR_INVALID;
  if (tk) *tk = t.kind;
  if (A9) std::memcpy(A9, t.A, 72); if (t3) std::memcpy(t3, t.t, 24); if (c3) std::memcpy(c3, t.c, 24);
  if (cov) std::memcpy(cov, t.cov.data(), sizeof(double) * t.cov.size());
  return LRR_OK;
}
int lrrr_est_get_matches(lrrr_est* e, int d, int64_t cap, int64_t* n, int32_t* from_idx, int32_t* to_idx, double* w) {
  int64_t m = e->impl->get_matches(d, cap, from_idx, to_idx, w); if (m < 0) return LRR_ERR_INVALID;
  if (n) *n = m; return LRR_OK;
}
int lrrr_est_metric(lrrr_est* e, int which, int d, double* out) { return e->impl->metric(which, d, out) ? LRR_OK : LRR_ERR_INVALID; }
int lrrr_est_estimate_ls_backward(lrrr_est* e, int ep) { return e->impl->estimate_ls_backward(ep) ? LRR_OK : LRR_ERR_INVALID; }
int lrrr_est_transfer_cov(lrrr_est* e, int d, double* S9) { return e->impl->transfer_cov(d, S9) ? LRR_OK : LRR_ERR_INVALID; }
int lrrr_est_get_state(lrrr_est* e, lrr_est_state* s) { e->impl->get_state(s); return LRR_OK; }
int lrrr_est_oscillated(lrrr_est* e, int* out) { *out = e->impl->oscillated(); return LRR_OK; }
int lrrr_objcov_eval(lrrr_est* e, int d, const double* x, double* f, double* g) {
  if (e->cdc && d == 0) return e->cdc->objcov(x, f, g) ? LRR_OK : LRR_ERR_STATE;
  return e->impl->objcov_eval(d, x, f, g) ? LRR_OK : LRR_ERR_UNSUPPORTED;
}
int lrrr_objpar_eval(lrrr_est* e, int d, const double* x, double* f, double* g) {
  if (e->cdc && d == 0) return e->cdc->objpar(x, f, g) ? LRR_OK : LRR_ERR_STATE;
  return e->impl->objpar_eval(d, x, f, g) ? LRR_OK : LRR_ERR_UNSUPPORTED;
}
// cdcl_estimation<3,12>::compute_weights of the REFERENCE (lrr_ref_cdc.cpp) on the estimator's moving subset / fixed set under its
// current transform and covariance.  The cap of 100 and the multiplier 5.0 are constants of the reference's source (:356-359).
int lrrr_cdc_compute_weights(lrrr_est* e, int d, int max_matches, double eval_multiplier, lrr_cdc_summary* out) {
  if (d != 0 || e->dof != 12 || e->kind == LRR_CDCL_ICP) return LRR_ERR_UNSUPPORTED;
  if (max_matches != 100 || eval_multiplier != 5.0) return LRR_ERR_UNSUPPORTED;
  RefTrans t; if (!e->impl->get_transform(0, t)) return LRR_ERR_INVALID;
  e->drop_cdc();
  e->cdc = lrr_ref_make_cdc(e->moving_all->base, e->has_sub ? e->msub.data() : nullptr, (int64_t)e->msub.size(), e->fixed_all->base);
  e->cdc->set_transform(t);
  return e->cdc->compute_weights(out) ? LRR_OK : LRR_ERR_STATE;
}
int lrrr_cdc_get_matches(lrrr_est* e, int d, int64_t cap_rows, int64_t cap_entries, int64_t* n_rows, int64_t* n_entries, int32_t* row_from,
                         int64_t* row_start, int32_t* to_idx, double* w) {
  if (d != 0 || !e->cdc) return LRR_ERR_STATE;
  if (n_rows) *n_rows = e->cdc->rows(); if (n_entries) *n_entries = e->cdc->entries();
  if (!row_from && !row_start && !to_idx && !w) return LRR_OK;
  if (cap_rows < e->cdc->rows() || cap_entries < e->cdc->entries()) return LRR_ERR_INVALID;
  e->cdc->get_matches(row_from, row_start, to_idx, w);
  return LRR_OK;
}
// cdcl_normalize_matches of the reference over its many-to-many set: (cm, rm, cf, rf)
int lrrr_test_cdc_normalisation(lrrr_est* e, double* norm8) { if (!e->cdc) return LRR_ERR_STATE; e->cdc->normalisation(norm8); return LRR_OK; }

double lrrr_test_weight_BT(double u) { return lrr_ref_weight_BT(u); }
double lrrr_test_rho(double u) { return lrr_ref_rho(u); }
double lrrr_test_rho_p(double u) { return lrr_ref_rho_p(u); }
double lrrr_test_mad_scale(const double* r, int64_t n) { return lrr_ref_mad_scale(r, n); }
void lrrr_test_trans_roundtrip(int kind, double* A9, double* t3, double* c3, double* cov, double rm, double rf, const double* cm, const double* cf, int do_un) {
  lrr_ref_trans_roundtrip(kind, A9, t3, c3, cov, rm, rf, cm, cf, do_un);
}
void lrrr_test_trans_inverse(int kind, double* A9, double* t3, double* c3) { lrr_ref_trans_inverse(kind, A9, t3, c3); }
void lrrr_test_projector2d(const double* C3, double* P3) { lrr_ref_projector2d(C3, P3); }
void lrrr_test_sim2d_roundtrip(double* s, double* ang, double* t2, double* cov16, double rm, double rf, const double* cm2, const double* cf2, int do_un) {
  lrr_ref_sim2d_roundtrip(s, ang, t2, cov16, rm, rf, cm2, cf2, do_un);
}
void lrrr_test_recompose(double* R9, double* t3, const double* dR9, const double* dt3) { lrr_ref_recompose(R9, t3, dR9, dt3); }
}
