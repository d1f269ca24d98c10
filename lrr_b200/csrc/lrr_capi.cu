// C ABI of liblrr_b200.so (include/lrr_gpu.h): contexts, feature sets, estimators and the host-side
// iteration control (ramp / oscillation state machine of the symmetric estimators), driving the sm_100a
// kernels of lrr_grid.cu, lrr_icp.cu, lrr_select.cu, lrr_metrics.cu and lrr_batch.cu.
// There is NO CPU fallback: every entry point that computes needs a CUDA device.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/lrr_gpu.h"
#include "lrr_internal.h"
#include "lrr_host.h"

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return LRR_ERR_CUDA; } \
  } while (0)

// Stream-ordered, pooled device memory: estimators and feature sets come and go per registration, and plain
// cudaFree costs milliseconds (implicit device synchronisation + unmapping).
static thread_local cudaStream_t g_alloc_stream = nullptr;
static void enter(lrr_ctx* ctx) { cudaSetDevice(ctx->device); g_alloc_stream = ctx->stream; }
template <class T> static cudaError_t dmalloc(T** p, size_t n) { return cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), g_alloc_stream); }
static void dfree(void* p) { if (p) cudaFreeAsync(p, g_alloc_stream); }

// ---- instrumentation ------------------------------------------------------------------------------------
struct PhaseTimer {
  lrr_ctx* ctx; int phase; cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(lrr_ctx* c, int ph) : ctx(c), phase(ph) {
    if (!ctx->profiling) return;
    cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream);
  }
  ~PhaseTimer() { if (!a) return; cudaEventRecord(b, ctx->stream); ctx->spans.push_back({phase, a, b}); }
};
#define PHASE(ph) PhaseTimer phase_timer_##ph(ctx, ph)
static void drain_spans(lrr_ctx* ctx) {
  for (auto& s : ctx->spans) {
    cudaEventSynchronize(s.b);
    float ms = 0; cudaEventElapsedTime(&ms, s.a, s.b);
    ctx->phase_ms[s.phase] += ms; ctx->phase_count[s.phase] += 1;
    cudaEventDestroy(s.a); cudaEventDestroy(s.b);
  }
  ctx->spans.clear();
}
extern "C" int lrr_ctx_set_profiling(lrr_ctx* ctx, int enable) {
  if (!ctx) return LRR_ERR_INVALID;
  enter(ctx);
  drain_spans(ctx);
  ctx->profiling = enable != 0;
  if (enable == 1) for (int i = 0; i < LRR_NPHASE; ++i) { ctx->phase_ms[i] = 0; ctx->phase_count[i] = 0; }   // 1: reset and start; 2: resume; 0: pause (counters kept)
  return LRR_OK;
}
extern "C" int lrr_ctx_phase_times(lrr_ctx* ctx, int cap, double* ms, int64_t* count) {
  if (!ctx || !ms || !count) return LRR_ERR_INVALID;
  enter(ctx);
  drain_spans(ctx);
  for (int i = 0; i < cap && i < LRR_NPHASE; ++i) { ms[i] = ctx->phase_ms[i]; count[i] = ctx->phase_count[i]; }
  return LRR_OK;
}

// ---- context -------------------------------------------------------------------------------------------
extern "C" int lrr_ctx_create(int device, lrr_ctx** out) {
  if (!out) return LRR_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return LRR_ERR_NO_DEVICE; }
  lrr_ctx* ctx = new lrr_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return LRR_ERR_CUDA; }
  bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ctx->stream = ctx->own_stream;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming) == cudaSuccess;
  {
    cudaMemPool_t pool; unsigned long long thr = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    g_alloc_stream = ctx->own_stream;
  }
  ok = ok && dmalloc(&ctx->partials, (size_t)2 * LRR_MAX_BLOCKS * LRR_NRED) == cudaSuccess;
  ok = ok && dmalloc(&ctx->ticket, 8) == cudaSuccess;
  ok = ok && cudaMemsetAsync(ctx->ticket, 0, 8 * sizeof(unsigned int), ctx->own_stream) == cudaSuccess;
  // four selectors (G and S for two groups)
  const size_t sel_uints = (size_t)LRR_SEL_FINE + LRR_SEL_COARSE;
  ok = ok && dmalloc(&ctx->sel_hist, 4 * sel_uints) == cudaSuccess;
  ok = ok && cudaMemsetAsync(ctx->sel_hist, 0, 4 * sel_uints * sizeof(unsigned int), ctx->own_stream) == cudaSuccess;
  ok = ok && dmalloc(&ctx->sel_cand, (size_t)4 * LRR_SEL_CAP) == cudaSuccess;
  ok = ok && dmalloc(&ctx->sel_misc, (size_t)4 * 16) == cudaSuccess;   // 16 x 8 bytes per selector
  ok = ok && cudaMemsetAsync(ctx->sel_misc, 0, 4 * 16 * sizeof(unsigned long long), ctx->own_stream) == cudaSuccess;
  ok = ok && dmalloc(&ctx->muse_table, 4 * 32) == cudaSuccess;
  ok = ok && dmalloc(&ctx->scratch_d, 1024) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&ctx->pinned, 2 * sizeof(Scal) + 4096) == cudaSuccess;
  if (ok && getenv("LRR_NO_MAILBOX") == nullptr) {   // result mailbox: mapped pinned memory the resid kernel writes directly (optional: plain copies otherwise)
    if (cudaHostAlloc((void**)&ctx->mailbox, 2 * 512 * sizeof(double), cudaHostAllocMapped) == cudaSuccess) {
      memset(ctx->mailbox, 0, 2 * 512 * sizeof(double));
      if (cudaHostGetDevicePointer((void**)&ctx->mailbox_dev, ctx->mailbox, 0) != cudaSuccess) { cudaFreeHost(ctx->mailbox); ctx->mailbox = nullptr; ctx->mailbox_dev = nullptr; }
    } else { ctx->mailbox = nullptr; cudaGetLastError(); }
  }
  ctx->sort.keys2 = nullptr; ctx->sort.hist = nullptr; ctx->sort.cap = 0; ctx->sort_keys = nullptr;
  {   // initial values of the behaviour switches (A/B measurements): environment, then lrr_ctx_set_option
    lrr_ctx::Options& o = ctx->opt;
    if (getenv("LRR_NO_PROOF")) o.certificates = 0;
    if (getenv("LRR_NO_TILE")) o.tile = 0;
    if (getenv("LRR_NO_SPEC")) o.spec = 0; else if (getenv("LRR_SPEC_FORCE")) o.spec = 2;
    if (getenv("LRR_NO_PIPE")) o.pipe = 0;
    if (getenv("LRR_NO_NORM_SKIP")) o.norm_skip = 0;
    if (getenv("LRR_NO_FLAT")) o.flat = 0;
    if (const char* ts = getenv("LRR_TILE_SPARSE")) o.tile_sparse = atoi(ts);
    if (getenv("LRR_NO_FIRST_SPARSE")) o.first_sparse = 0;
    if (getenv("LRR_NO_BATCH_SORT")) o.batch_sort = 0;
    if (const char* q = getenv("LRR_BATCH_QUANTUM")) o.batch_quantum = atoi(q);
  }
  ok = ok && cudaStreamSynchronize(ctx->own_stream) == cudaSuccess;   // (allocations and clears are ordered on the context's own stream)
  if (!ok) { cudaGetLastError(); lrr_ctx_destroy(ctx); return LRR_ERR_CUDA; }
  for (int i = 0; i < 4; ++i) {
    SelDev& s = ctx->sel[i];
    s.fine = ctx->sel_hist + (size_t)i * sel_uints; s.coarse = s.fine + LRR_SEL_FINE;
    s.cand = ctx->sel_cand + (size_t)i * LRR_SEL_CAP;
    unsigned long long* m = ctx->sel_misc + (size_t)i * 16;
    s.cand_minmax = m; s.cand_count = (unsigned int*)(m + 2); s.bin = (int*)(m + 3); s.rank = (unsigned int*)(m + 4);
    s.count = (double*)(m + 5); s.result = (double*)(m + 6); s.bracket = (double*)(m + 8);
    s.inv_width = (i < 2) ? (double)LRR_SEL_FINE / LRR_STRENGTH_CAP : (double)LRR_SEL_FINE;   // G keys in [0,9e4], S keys in (0,1]
  }
  *out = ctx;
  return LRR_OK;
}
extern "C" void lrr_ctx_destroy(lrr_ctx* ctx) {
  if (!ctx) return;
  enter(ctx);
  dfree(ctx->partials); dfree(ctx->ticket); dfree(ctx->sel_hist); dfree(ctx->sel_cand); dfree(ctx->sel_misc);
  dfree(ctx->muse_table); dfree(ctx->scratch_d); dfree(ctx->sort_keys);
  for (auto& w : ctx->batch_ws) dfree(w.p);
  ctx->batch_ws.clear();
  if (ctx->sort.keys2) lrr_sort_free(&ctx->sort);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
  if (ctx->copy_ev) cudaEventDestroy(ctx->copy_ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->dirs_stream) cudaStreamDestroy(ctx->dirs_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}
extern "C" int lrr_ctx_set_stream(lrr_ctx* ctx, void* s) { if (!ctx) return LRR_ERR_INVALID; ctx->stream = s ? (cudaStream_t)s : ctx->own_stream; return LRR_OK; }
extern "C" int lrr_ctx_synchronize(lrr_ctx* ctx) {
  if (!ctx) return LRR_ERR_INVALID;
  enter(ctx); CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->dirs_stream) CK(cudaStreamSynchronize(ctx->dirs_stream));   // asynchronous uploads: the callers' `dirs` buffers are free again
  return LRR_OK;
}
extern "C" const char* lrr_ctx_last_error(lrr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t lrr_ctx_launch_count(lrr_ctx* ctx) { return ctx ? (int64_t)ctx->launches : 0; }

int lrr_ctx_need_sort(lrr_ctx* ctx, size_t n) {
  if (ctx->sort.cap >= n && ctx->sort_keys) return LRR_OK;
  if (ctx->sort.keys2) lrr_sort_free(&ctx->sort);
  dfree(ctx->sort_keys); ctx->sort_keys = nullptr;
  CK(lrr_sort_alloc(&ctx->sort, n));
  CK(dmalloc(&ctx->sort_keys, n));
  return LRR_OK;
}

// ---- feature sets ----------------------------------------------------------------------------------------
// cov2 != null: planar set, P6 / C6 are derived on the device from the 2x2 covariances (lrr_features_upload_cov)
static int features_upload_impl(lrr_ctx* ctx, int64_t n, const double* xyz, const double* P6, const double* C6, const double* strength,
                                const uint8_t* shape, const double* dirs, const double* cov2, lrr_fset** out) {
  enter(ctx);
  lrr_fset* fs = new lrr_fset(); fs->ctx = ctx; fs->n = n; fs->dim = cov2 ? 2 : 3;
  std::vector<double> zeros; std::vector<uint8_t> zshape;
  if (!strength) { zeros.assign((size_t)n, 0.0); strength = zeros.data(); }
  if (!shape) { zshape.assign((size_t)n, 0); shape = zshape.data(); }
  cudaStream_t st = ctx->stream;
#define FCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cudaStreamSynchronize(ctx->copy_stream); lrr_features_destroy(fs); return LRR_ERR_CUDA; } } while (0)
  // the grid is built from the locations alone: they go first on the work stream, everything else (5x the bytes) is copied on
  // ctx->copy_stream while the grid kernels run, and joins before the cell-ordered records are filled
  cudaStream_t cs = ctx->copy_stream;
  FCK(dmalloc(&fs->xyz, (size_t)3 * n)); FCK(dmalloc(&fs->s, (size_t)n)); FCK(dmalloc(&fs->P6, (size_t)6 * n));
  FCK(dmalloc(&fs->shape, (size_t)n));
  FCK(cudaMemcpyAsync(fs->xyz, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  double* cov2_d = nullptr;
  if (cov2) {
    FCK(dmalloc(&fs->C6, (size_t)6 * n)); FCK(dmalloc(&cov2_d, (size_t)3 * n));
    FCK(cudaMemcpyAsync(cov2_d, cov2, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    lrr_k_projector2_from_cov(cov2_d, (int)n, fs->P6, fs->C6, st, &ctx->launches);
  } else FCK(cudaMemcpyAsync(fs->P6, P6, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, cs));
  FCK(cudaMemcpyAsync(fs->s, strength, sizeof(double) * n, cudaMemcpyHostToDevice, cs));
  FCK(cudaMemcpyAsync(fs->shape, shape, (size_t)n, cudaMemcpyHostToDevice, cs));
  if (C6 && !cov2) { FCK(dmalloc(&fs->C6, (size_t)6 * n)); FCK(cudaMemcpyAsync(fs->C6, C6, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, cs)); }
  FCK(cudaEventRecord(ctx->copy_ev, cs));
  // the directions (47 % of a cdcl_feature_ICP set's bytes) are read by the angle metrics only: last, on their own stream, and in
  // asynchronous mode not waited for at all (lrr_ctx_set_async_upload; fset_wait_dirs)
  if (dirs) {
    FCK(dmalloc(&fs->dirs, (size_t)9 * n));
    if (ctx->async_upload && !ctx->dirs_stream) FCK(cudaStreamCreateWithFlags(&ctx->dirs_stream, cudaStreamNonBlocking));
    cudaStream_t ds = ctx->async_upload ? ctx->dirs_stream : cs;
    if (ctx->async_upload) FCK(cudaStreamWaitEvent(ds, ctx->copy_ev, 0));   // behind this set's own projectors, not in front of them
    FCK(cudaMemcpyAsync(fs->dirs, dirs, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, ds));
    if (ctx->async_upload) {
      FCK(cudaEventCreateWithFlags(&fs->dirs_ev, cudaEventDisableTiming)); FCK(cudaEventRecord(fs->dirs_ev, ds)); fs->dirs_pending = true;
    }
  }
  FCK(dmalloc(&fs->pos4, (size_t)n)); FCK(dmalloc(&fs->rec, (size_t)8 * n)); FCK(dmalloc(&fs->sorted_of_orig, (size_t)n));
  FsetBuild fb; memset(&fb, 0, sizeof(fb));
  fb.n = n; fb.x = fs->xyz; fb.y = fs->xyz + n; fb.z = fs->xyz + 2 * n; fb.pos4 = fs->pos4; fb.sorted_of_orig = fs->sorted_of_orig;
  FCK(lrr_grid_build(&fb, st, &ctx->launches));
  fs->cell_start = fb.cell_start;
  FsetDev& d = fs->dev;
  d.n = (int)n; d.x = fs->xyz; d.y = fs->xyz + n; d.z = fs->xyz + 2 * n; d.s = fs->s; d.P6 = fs->P6; d.C6 = fs->C6; d.dirs = fs->dirs;
  if (fs->C6) FCK(dmalloc(&fs->crec, (size_t)8 * n));
  d.shape = fs->shape; d.grid = fb.grid; d.rec = fs->rec; d.crec = fs->crec; d.sorted_of_orig = fs->sorted_of_orig;
  FCK(cudaStreamWaitEvent(st, ctx->copy_ev, 0));
  lrr_k_fill_records(d, fs->rec, fs->crec, st, &ctx->launches);
  FCK(cudaStreamSynchronize(st));
  if (!fs->dirs_pending) FCK(cudaStreamSynchronize(cs));   // synchronous contract: every host buffer has been read when the call returns
  FCK(cudaGetLastError());
  dfree(cov2_d);
#undef FCK
  *out = fs;
  return LRR_OK;
}
extern "C" int lrr_features_upload(lrr_ctx* ctx, int64_t n, const double* xyz, const double* P6, const double* C6, const double* strength,
                                   const uint8_t* shape, const double* dirs, lrr_fset** out) {
  if (!ctx || n < 0 || n > 0x7fffff00LL || !out || (n > 0 && (!xyz || !P6))) return LRR_ERR_INVALID;
  return features_upload_impl(ctx, n, xyz, P6, C6, strength, shape, dirs, nullptr, out);
}
// plain cdcl_feature<dim> sets (cdcl/cdcl_feature.h:17-64) for LRR_CDCL_ICP: planar locations stored with z = 0, so the uniform
// grid degenerates to one layer of cells and every kernel of the 3-D path runs unchanged
extern "C" int lrr_features_upload_cov(lrr_ctx* ctx, int dim, int64_t n, const double* loc, const double* cov, const double* strength, lrr_fset** out) {
  if (!ctx || n < 0 || n > 0x7fffff00LL || !out || (n > 0 && (!loc || !cov))) return LRR_ERR_INVALID;
  if (dim != 2) return LRR_ERR_UNSUPPORTED;
  std::vector<double> xyz((size_t)3 * n, 0.0);
  if (n) memcpy(xyz.data(), loc, sizeof(double) * 2 * n);
  static const double dummy = 0.0;
  return features_upload_impl(ctx, n, xyz.data(), nullptr, nullptr, strength, nullptr, nullptr, n ? cov : &dummy, out);
}
// consumers of the direction arrays: wait (on the device, stream-ordered) for an asynchronous upload still in flight
static void fset_wait_dirs(lrr_fset* fs, cudaStream_t st) { if (fs && fs->dirs_pending) cudaStreamWaitEvent(st, fs->dirs_ev, 0); }
extern "C" int lrr_ctx_set_option(lrr_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return LRR_ERR_INVALID;
  lrr_ctx::Options& o = ctx->opt;
  const std::string k(key);
  if (k == "certificates") o.certificates = value != 0;
  else if (k == "tile") o.tile = value != 0;
  else if (k == "speculative") { if (value < 0 || value > 2) return LRR_ERR_INVALID; o.spec = value; }
  else if (k == "pipeline") o.pipe = value != 0;
  else if (k == "norm_skip") o.norm_skip = value != 0;
  else if (k == "flat_scan") o.flat = value != 0;
  else if (k == "tile_sparse") { if (value < -1 || value > 32) return LRR_ERR_INVALID; o.tile_sparse = value; }
  else if (k == "first_sparse") o.first_sparse = value != 0;
  else if (k == "batch_sort") o.batch_sort = value != 0;
  else if (k == "batch_quantum") { if (value < 0 || value > 250) return LRR_ERR_INVALID; o.batch_quantum = value; }
  else if (k == "batch_slots") { if (value < 1 || value > 65536) return LRR_ERR_INVALID; o.batch_slots = value; }
  else { ctx->err = "lrr_ctx_set_option: unknown key " + k; return LRR_ERR_INVALID; }
  return LRR_OK;
}
extern "C" int lrr_ctx_set_async_upload(lrr_ctx* ctx, int on) { if (!ctx) return LRR_ERR_INVALID; ctx->async_upload = on != 0; return LRR_OK; }
extern "C" void lrr_features_destroy(lrr_fset* fs) {
  if (!fs) return;
  enter(fs->ctx);
  if (fs->dirs_ev) { cudaEventSynchronize(fs->dirs_ev); cudaEventDestroy(fs->dirs_ev); }
  dfree(fs->xyz); dfree(fs->s); dfree(fs->P6); dfree(fs->C6); dfree(fs->dirs); dfree(fs->shape);
  dfree(fs->labels); dfree(fs->pos4); dfree(fs->rec); dfree(fs->crec); dfree(fs->sorted_of_orig); dfree(fs->cell_start);
  delete fs;
}
extern "C" int64_t lrr_features_size(const lrr_fset* fs) { return fs ? fs->n : 0; }
extern "C" int lrr_features_nearest(lrr_fset* fs, int64_t nq, const double* q, int32_t* idx) {
  if (!fs || nq < 0 || (nq > 0 && (!q || !idx))) return LRR_ERR_INVALID;
  lrr_ctx* ctx = fs->ctx;
  enter(ctx);
  if (nq == 0) return LRR_OK;
  double* dq = nullptr; int* di = nullptr;
  CK(dmalloc(&dq, (size_t)3 * nq)); CK(dmalloc(&di, (size_t)nq));
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(dq, q, sizeof(double) * 3 * nq, cudaMemcpyHostToDevice, st));
  CK(lrr_nn_query(fs->dev.grid, nq, dq, dq + nq, dq + 2 * nq, di, st, &ctx->launches));
  CK(cudaMemcpyAsync(idx, di, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  dfree(dq); dfree(di);
  return LRR_OK;
}

// ---- estimator helpers -----------------------------------------------------------------------------------
static void identity_scal(Scal& s, int dof) {
  memset(&s, 0, sizeof(s));
  s.A[0] = s.A[4] = s.A[8] = 1.0; s.Ag[0] = s.Ag[4] = s.Ag[8] = 1.0;
  s.rfrom = s.rto = 1.0; s.G = s.S = 1.0;
  (void)dof;
}

static int dir_alloc(lrr_ctx* ctx, lrr_dir& d, lrr_fset* from_all, lrr_fset* to_all) {
  d.from_all = from_all; d.to_all = to_all;
  const size_t cap = (size_t)from_all->n, nto = (size_t)to_all->n;
  d.cap = cap;
  CK(dmalloc(&d.fx, 4 * cap)); d.fy = d.fx + cap; d.fz = d.fx + 2 * cap; d.fs = d.fx + 3 * cap;
  CK(dmalloc(&d.fP6, 6 * cap)); CK(dmalloc(&d.forig, cap)); CK(dmalloc(&d.order, cap));
  CK(dmalloc(&d.sidx, cap)); CK(dmalloc(&d.w, 4 * cap)); d.r2 = d.w + cap; d.gkey = d.w + 2 * cap; d.skey = d.w + 3 * cap;
  CK(dmalloc(&d.stamp, nto)); CK(cudaMemsetAsync(d.stamp, 0xFF, nto * sizeof(unsigned long long), ctx->stream));
  CK(dmalloc(&d.sc, 1));
  CK(dmalloc(&d.hq_k, 3 * cap + 1)); d.hq_s = d.hq_k + cap; d.hq_o = d.hq_k + 2 * cap; d.hq_count = d.hq_k + 3 * cap;
  CK(dmalloc(&d.prev_s, cap)); CK(dmalloc(&d.proc, cap)); CK(dmalloc(&d.proc_mark, 2 * cap));
  CK(dmalloc(&d.hq_best, cap)); CK(cudaMemsetAsync(d.hq_count, 0, sizeof(int), ctx->stream));
  CK(dmalloc(&d.proof, cap ? cap : 1)); CK(dmalloc(&d.pos_of, cap ? cap : 1)); CK(dmalloc(&d.cprev, cap ? cap : 1)); CK(dmalloc(&d.tf, 3 * (cap ? cap : 1)));
  {  // tiles of the from-features: cubes of LRR_TILE_T cells of the to-grid over the bounding box of the parent set
    const GridDev& gf = from_all->dev.grid; const GridDev& gt = to_all->dev.grid;
    double edge = LRR_TILE_T * gt.h;
    if (!(edge > 0)) edge = 1.0;
    TileGeom& tg = d.tgeom;
    for (;;) {
      tg.ntx = std::max(1, (int)std::ceil(gf.nx * gf.h / edge)); tg.nty = std::max(1, (int)std::ceil(gf.ny * gf.h / edge)); tg.ntz = std::max(1, (int)std::ceil(gf.nz * gf.h / edge));
      if ((double)tg.ntx * tg.nty * tg.ntz <= 4194304.0) break;
      edge *= 2.0;
    }
    tg.ox = gf.ox; tg.oy = gf.oy; tg.oz = gf.oz; tg.inv_edge = 1.0 / edge;
    d.ntiles = tg.ntx * tg.nty * tg.ntz;
    CK(dmalloc(&d.tile_cnt, (size_t)d.ntiles)); CK(dmalloc(&d.tile_start, (size_t)d.ntiles + 1)); CK(dmalloc(&d.chunk_start, (size_t)d.ntiles + 1));
    CK(dmalloc(&d.tile_of, cap ? cap : 1));
    CK(dmalloc(&d.chunk_tab, cap / 32 + (size_t)d.ntiles + 1));   // every tile adds at most one partly filled chunk
    CK(dmalloc(&d.todo, cap / 32 + (size_t)d.ntiles + 1)); CK(dmalloc(&d.todo_count, 2)); CK(cudaMemsetAsync(d.todo_count, 0, 2 * sizeof(int), ctx->stream)); CK(dmalloc(&d.fq, cap ? cap : 1));
  }
  d.epoch = 0; d.has_order = false; d.matches = 0; d.weights_fresh = false;
  d.initial_scale = std::numeric_limits<double>::quiet_NaN();
  return LRR_OK;
}
static void dir_free(lrr_dir& d) {
  dfree(d.fx); dfree(d.fP6); dfree(d.forig); dfree(d.order); dfree(d.sidx); dfree(d.w); dfree(d.stamp); dfree(d.sc); dfree(d.hq_k); dfree(d.hq_best); dfree(d.prev_s); dfree(d.proc); dfree(d.proc_mark); dfree(d.proof); dfree(d.pos_of); dfree(d.cprev); dfree(d.tf);
  dfree(d.tile_cnt); dfree(d.tile_start); dfree(d.chunk_start); dfree(d.tile_of); dfree(d.chunk_tab); dfree(d.todo); dfree(d.todo_count); dfree(d.fq);
  dfree(d.cdc_efrom); dfree(d.cdc_eslot); dfree(d.cdc_ew); dfree(d.cdc_rowcnt); dfree(d.cdc_roww); dfree(d.cdc_tow); dfree(d.cdc_cand); dfree(d.cdc_candw); dfree(d.cdc_cursor);
}
static int dir_set_subset(lrr_ctx* ctx, lrr_dir& d, const int32_t* idx, int64_t n) {
  if (n < 0 || (size_t)n > d.cap) return LRR_ERR_INVALID;
  cudaStream_t st = ctx->stream;
  if (idx) {
    for (int64_t i = 0; i < n; ++i) if (idx[i] < 0 || idx[i] >= d.from_all->n) return LRR_ERR_INVALID;
    CK(cudaMemcpyAsync(d.forig, idx, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // idx is caller memory
    lrr_k_gather_from(d.from_all->dev, d.forig, (int)n, d.fx, d.fy, d.fz, d.fs, d.fP6, st, &ctx->launches);
    d.subset = true;
  } else {
    n = d.from_all->n;
    lrr_k_gather_from(d.from_all->dev, nullptr, (int)n, d.fx, d.fy, d.fz, d.fs, d.fP6, st, &ctx->launches);
    d.subset = false;
  }
  d.n_in = n; d.has_order = false; d.matches = 0; d.weights_fresh = false; d.cdc_active = false;
  CK(cudaMemsetAsync(d.prev_s, 0xFF, sizeof(int) * (size_t)(n ? n : 1), st));   // no search candidates yet
  CK(cudaMemsetAsync(d.proof, 0, sizeof(double4) * (size_t)(n ? n : 1), st));   // ... and no neighbour certificates
  CK(cudaMemsetAsync(d.cprev, 0xFF, sizeof(int) * (size_t)(n ? n : 1), st));
  // processing order of the match phase = tile order (spatially coherent; the finest level stages each tile's cells in shared memory)
  lrr_k_build_tiles(d.fx, d.fy, d.fz, (int)n, d.tgeom, d.tile_cnt, d.tile_start, d.chunk_start, d.tile_of, d.proc_mark, d.chunk_tab, d.proc, st, &ctx->launches);
  lrr_k_tile_mirror(d.proc, d.fx, d.fy, d.fz, (int)n, d.pos_of, d.tf, st, &ctx->launches);
  return LRR_OK;
}

static DirArgs dir_args(const lrr_dir& d) {
  DirArgs a; memset(&a, 0, sizeof(a));
  const size_t n = (size_t)d.n_in;
  a.fx = d.fx; a.fy = d.fy; a.fz = d.fz; a.fs = d.fs;
  a.fP6 = d.fP6; a.forig = d.subset ? d.forig : nullptr; a.n_in = (int)d.n_in;
  a.order = d.has_order ? d.order : nullptr;
  for (int r = 0; r < 4; ++r) a.perm_key[r] = d.perm_key[r];
  a.perm_half = d.has_order ? 0 : d.perm_half;
  a.M = (int)std::min<int64_t>(d.number_points, d.n_in);
  // few queries (the ramp): the thread-level ball scan of pass A is the critical path of the launch (up to nine rows one after the
  // other) while pass B has a warp to spare for every queued query; measured 62 -> 45 us per find_closest at 1000 points
  a.ball_rows = a.M <= 16384 ? 2 : 0;
  a.to = d.to_all->dev; a.from_all = d.from_all->dev;
  a.sidx = d.sidx; a.w = d.w; a.r2 = d.r2; a.gkey = d.gkey; a.skey = d.skey;
  a.stamp = d.stamp; a.tag = (~(unsigned long long)d.epoch) << 32;
  const lrr_ctx::Options& opt = d.from_all->ctx->opt;
  const bool no_proof = !opt.certificates;
  a.proof = no_proof ? nullptr : d.proof; a.todo = no_proof ? nullptr : d.todo; a.todo_count = d.todo_count; a.fq = d.fq; a.fq_count = d.todo_count + 1;
  if (a.proof) { a.pos_of = d.pos_of; a.cprev = d.cprev; a.tfx = d.tf; a.tfy = d.tf + n; a.tfz = d.tf + 2 * n; }
  a.prev_s = d.prev_s; a.proc = (d.has_order || d.perm_half) ? nullptr : d.proc; a.chunk_tab = (a.proc && d.n_in > 0 && opt.tile) ? d.chunk_tab : nullptr; a.nchunks = d.chunk_start; a.hq_k = d.hq_k; a.hq_best = d.hq_best; a.hq_s = d.hq_s; a.hq_o = d.hq_o; a.hq_count = d.hq_count;
  a.sc = d.sc;
  // certified finest level (lrr_dir_certified, lrr_tile.cu): kept matches are not touched, the stamps come from k_restamp
  a.stamp_in_match = !(a.proof && a.todo && a.chunk_tab && a.proc && a.M == a.n_in);
  return a;
}

static EstArgs est_args(lrr_est* e, int src = 0) {
  EstArgs a; memset(&a, 0, sizeof(a));
  lrr_ctx* ctx = e->ctx;
  a.ndir = e->sym() ? 2 : 1;
  for (int d = 0; d < a.ndir; ++d) a.d[d] = dir_args(e->d[d]);
  int f = 0;
  if (e->rrl()) f |= LRR_F_RRL;
  if (e->dof == 6) f |= LRR_F_RIGID;
  if (e->normalize) f |= LRR_F_NORMALIZE;
  if (e->sym()) f |= LRR_F_SYM;
  if (e->strength) f |= LRR_F_STRENGTH;
  if (e->kind == LRR_CDCL_SYM) f |= LRR_F_MERGED;
  if (e->icp()) f |= LRR_F_ICP;
  if (e->dof == 4) f |= LRR_F_SIM2D;
  if (!ctx->opt.flat) f |= LRR_F_NOFLAT;
  a.flags = f; a.dof = e->dof; a.src = src; a.estimate_parameters = 1;
  // chunks with at most tile_sparse failed certificates put their queries on the per-query sparse list, the others are staged whole.
  // While few certificates fail (sparse_hint: under a quarter of the queries in the last certificate pass — inside a registration
  // 2-6 %, near ties) EVERY failed query goes to the list: a single staged chunk costs ~25 us of one warp's latency and set the
  // duration of k_match_tile (43-61 -> 32-37 us per in-registration iteration; config 5 after convergence 309 -> 128 us)
  { const int ts = ctx->opt.tile_sparse;
    // (without the hint — a level's first find_closest — whole chunks are staged: with candidates for every query the list would win
    //  up to ~300k features per side, 160 -> 140 us at 100k, but in a fresh estimator the 36 % of the queries that have no previous
    //  neighbour yet make it lose, 0.36 -> 0.54 ms: profiles/r2_summary.md)
    // (round 2, later: queries without a previous neighbour borrow the candidate of a neighbouring lane of their chunk, k_match_cert;
    //  with that the list is tried for a level's first find_closest again up to 300k features per side: opt.first_sparse)
    int64_t total_in = 0; for (int d = 0; d < a.ndir; ++d) total_in += e->d[d].n_in;
    const bool first_list = ctx->opt.first_sparse && total_in <= 300000 * (int64_t)a.ndir;
    a.tile_sparse = (f & LRR_F_NOFLAT) ? 0 : (ts >= 0 ? ts : ((e->sparse_hint || first_list) ? 32 : 8)); }
  a.selG[0] = ctx->sel[0]; a.selG[1] = ctx->sel[1]; a.selS[0] = ctx->sel[2]; a.selS[1] = ctx->sel[3];
  a.partials = ctx->partials; a.ticket = ctx->ticket;
  return a;
}

// upload / download the head of a direction's scalars (everything before the moments)
static const size_t kScalHead = offsetof(Scal, mom);
static int push_scal(lrr_est* e, int d) {
  lrr_ctx* ctx = e->ctx;
  CK(cudaMemcpyAsync(e->d[d].sc, &e->h[d], kScalHead, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LRR_OK;
}
static int pull_scal(lrr_est* e, int mask) {
  lrr_ctx* ctx = e->ctx;
  char* pin = (char*)ctx->pinned;
  for (int d = 0; d < 2; ++d) if (mask & (1 << d))
    CK(cudaMemcpyAsync(pin + d * sizeof(Scal), e->d[d].sc, kScalHead, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  for (int d = 0; d < 2; ++d) if (mask & (1 << d)) memcpy(&e->h[d], pin + d * sizeof(Scal), kScalHead);
  return LRR_OK;
}
static int dmask(const lrr_est* e) { return e->sym() ? 3 : 1; }

// Read the scalars the resid kernel posted into the mailbox (sequence number seq): poll, no copy, no stream synchronisation.
// Falls back to the plain copies when the stream has drained without the number showing up (cannot happen) or reports an error.
static int pull_scal_mailbox(lrr_est* e, int mask, unsigned long long seq) {
  lrr_ctx* ctx = e->ctx;
  volatile unsigned long long* flag[2] = {reinterpret_cast<volatile unsigned long long*>(ctx->mailbox + 500),
                                          reinterpret_cast<volatile unsigned long long*>(ctx->mailbox + 512 + 500)};
  auto ready = [&]() { for (int d = 0; d < 2; ++d) if ((mask & (1 << d)) && *flag[d] != seq) return false; return true; };
  for (unsigned long long spins = 1; !ready(); ++spins) {
    if ((spins & 0xfff) == 0) {
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q == cudaSuccess) { if (ready()) break; return pull_scal(e, mask); }
      if (q != cudaErrorNotReady) { ctx->err = cudaGetErrorString(q); return LRR_ERR_CUDA; }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  for (int d = 0; d < 2; ++d) if (mask & (1 << d)) memcpy(&e->h[d], ctx->mailbox + 512 * d, kScalHead);
  return LRR_OK;
}

static void inverse_trans(const Scal& f, Scal& b) {   // cdcl_trans_affine::inverse / rigid3d::inverse
  double invA[9]; lrr_inv3(f.A, invA);
  for (int i = 0; i < 9; ++i) b.A[i] = invA[i];
  for (int i = 0; i < 3; ++i) {
    double acc = (-invA[3 * i]) * f.t[0]; acc += (-invA[3 * i + 1]) * f.t[1]; acc += (-invA[3 * i + 2]) * f.t[2];
    b.t[i] = acc + f.c[i]; b.c[i] = 0.0;
  }
  memset(b.cov, 0, sizeof(b.cov));
}

static void reset_control(lrr_est* e) { e->oscillation_count = 0; e->error_difference = 0.0; e->number_matches = 1000; e->finest_level = 0; }

extern "C" int lrr_est_create(lrr_ctx* ctx, int kind, int dim, int dof, lrr_fset* moving_all, lrr_fset* fixed_all, lrr_est** out) {
  if (!ctx || !moving_all || !fixed_all || !out) return LRR_ERR_INVALID;
  if (moving_all->ctx != ctx || fixed_all->ctx != ctx) return LRR_ERR_INVALID;
  if (kind < 0 || kind > 4) return LRR_ERR_INVALID;
  if (kind == LRR_CDCL_ICP) {   // cdcl_estimation_ICP<2,4> over planar cdcl_feature<2> sets (cdcl/Templates/cdcl_estimation_ICP+2.4-.cxx)
    if (dim != 2 || dof != 4 || moving_all->dim != 2 || fixed_all->dim != 2) return LRR_ERR_UNSUPPORTED;
  } else {
    if (dim != 3 || (dof != 12 && dof != 6) || moving_all->dim != 3 || fixed_all->dim != 3) return LRR_ERR_UNSUPPORTED;
    if (dof == 6 && (kind == LRR_RRL_FWD || kind == LRR_RRL_SYM)) return LRR_ERR_UNSUPPORTED;   // Q6
  }
  enter(ctx);
  lrr_est* e = new lrr_est();
  e->ctx = ctx; e->kind = kind; e->dof = dof; e->dim = dim; e->moving_all = moving_all; e->fixed_all = fixed_all;
  e->normalize = e->sym(); e->strength = e->sym();   // the symmetric ctor switches both on (cdcl sym :151-156)
  reset_control(e);
  int rc = dir_alloc(ctx, e->d[0], moving_all, fixed_all);
  if (rc == LRR_OK && e->sym()) rc = dir_alloc(ctx, e->d[1], fixed_all, moving_all);
  if (rc == LRR_OK) rc = dir_set_subset(ctx, e->d[0], nullptr, 0);
  if (rc == LRR_OK && e->sym()) rc = dir_set_subset(ctx, e->d[1], nullptr, 0);
  if (rc != LRR_OK) { lrr_est_destroy(e); return rc; }
  for (int d = 0; d < 2; ++d) {
    identity_scal(e->h[d], dof);
    e->d[d].number_points = e->icp() ? e->d[d].n_in : std::min<int64_t>(1000, e->d[d].n_in);   // (cdcl_estimation_ICP: no point ramp)
    e->d[d].scale_estimated = false;
  }
  for (int d = 0; d < (e->sym() ? 2 : 1); ++d) {
    cudaMemsetAsync(e->d[d].sc, 0, sizeof(Scal), ctx->stream);
    if ((rc = push_scal(e, d)) != LRR_OK) { lrr_est_destroy(e); return rc; }
  }
  *out = e;
  return LRR_OK;
}
extern "C" void lrr_est_destroy(lrr_est* e) {
  if (!e) return;
  enter(e->ctx);
  cudaStreamSynchronize(e->ctx->stream);
  dir_free(e->d[0]); if (e->sym()) dir_free(e->d[1]);
  delete e;
}

// Called by every entry point that changes what the next find_closest / estimate depends on: the device-side normalisation is no
// longer known to be this estimator's (norm_token), and a pre-match (the certificate pass of the next find_closest, run as the trailing
// pass of the last iteration) is dropped together with the work list it may have left.
static void touch(lrr_est* e) {
  if (!e) return;
  e->norm_token = false; e->steady_hint = false; e->sparse_hint = false;
  if (e->pre_matched) {
    enter(e->ctx);
    for (int d = 0; d < (e->sym() ? 2 : 1); ++d) cudaMemsetAsync(e->d[d].todo_count, 0, 2 * sizeof(int), e->ctx->stream);
    e->pre_matched = false;
  }
}
// the certificates assume that match position k holds from-feature k: void them whenever the order policy changes
static void void_certificates(lrr_est* e) {
  enter(e->ctx);
  for (int d = 0; d < (e->sym() ? 2 : 1); ++d) {
    lrr_dir& D = e->d[d];
    const size_t n = (size_t)(D.n_in ? D.n_in : 1);
    cudaMemsetAsync(D.proof, 0, sizeof(double4) * n, e->ctx->stream);
    cudaMemsetAsync(D.cprev, 0xFF, sizeof(int) * n, e->ctx->stream);
    cudaMemsetAsync(D.prev_s, 0xFF, sizeof(int) * n, e->ctx->stream);
  }
}

extern "C" int lrr_est_set_subsets(lrr_est* e, const int32_t* mi, int64_t nm, const int32_t* fi, int64_t nf) {
  touch(e);
  if (!e || !mi || nm < 0) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  int rc = dir_set_subset(ctx, e->d[0], mi, nm);
  if (rc != LRR_OK) return rc;
  if (e->sym()) {
    rc = (fi && nf > 0) ? dir_set_subset(ctx, e->d[1], fi, nf) : dir_set_subset(ctx, e->d[1], nullptr, 0);
    if (rc != LRR_OK) return rc;
    reset_control(e);   // set_features (cdcl sym :206-226)
  }
  if (e->icp()) e->d[0].number_points = e->d[0].n_in;   // set_moving_data (cdcl_estimation_ICP.h:80): every point, every iteration
  e->werr_cache_valid = false;
  return LRR_OK;
}

static void set_identity_norm(Scal& s) {
  for (int d = 0; d < 3; ++d) { s.cfrom[d] = 0; s.cto[d] = 0; }
  s.rfrom = 1; s.rto = 1;
}

extern "C" int lrr_est_set_transform(lrr_est* e, int tk, const double* A9, const double* t3, const double* c3, const double* cov) {
  touch(e);
  if (!e || !A9 || !t3) return LRR_ERR_INVALID;
  if ((tk == LRR_TRANS_AFFINE) != (e->dof == 12) || (tk == LRR_TRANS_SIMILARITY2D) != (e->dof == 4)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  int rc = pull_scal(e, dmask(e)); if (rc != LRR_OK) return rc;
  Scal& f = e->h[0];
  memcpy(f.A, A9, 72); memcpy(f.t, t3, 24);
  if (c3) memcpy(f.c, c3, 24); else memset(f.c, 0, 24);
  if (tk == LRR_TRANS_SIMILARITY2D) {
    // (A = s R, t) -> cdcl_trans_similarity2d(s, ang, t) -> s * R(ang)  (ctor #2, similarity2d.cxx:33-47): the planar block is
    // re-formed from scale and angle, everything outside it is fixed (z stays 0), the centre is always 0
    const double s_ = std::sqrt(A9[0] * A9[0] + A9[3] * A9[3]), ang = std::atan2(A9[3], A9[0]);
    const double sa = std::sin(ang), ca = std::cos(ang);
    memset(f.A, 0, 72); f.A[0] = s_ * ca; f.A[1] = s_ * -sa; f.A[3] = s_ * sa; f.A[4] = s_ * ca; f.A[8] = 1.0;
    f.t[2] = 0.0; memset(f.c, 0, 24);
  }
  memset(f.cov, 0, sizeof(f.cov));
  if (cov) for (int i = 0; i < e->dof * e->dof; ++i) f.cov[i] = cov[i];
  f.from_mapped = 0; f.status = 0; f.scale_estimated = 0;
  if (!e->normalize) set_identity_norm(f);
  for (int d = 0; d < 2; ++d) {   // set_transform (cdcl :141-153): scale_estimated_ = false, number_points_ = 1000
    e->d[d].scale_estimated = false;
    e->d[d].number_points = e->icp() ? e->d[d].n_in : std::min<int64_t>(1000, e->d[d].n_in);   // (cdcl_estimation_ICP.h:70 only resets scale_estimated_)
  }
  if ((rc = push_scal(e, 0)) != LRR_OK) return rc;
  if (e->sym()) {
    Scal& b = e->h[1];
    inverse_trans(f, b); b.from_mapped = 0; b.status = 0; b.scale_estimated = 0;
    if (!e->normalize) set_identity_norm(b);
    if ((rc = push_scal(e, 1)) != LRR_OK) return rc;
    reset_control(e);
  }
  e->iterations_done = 0;
  e->werr_cache_valid = false;
  return LRR_OK;
}

extern "C" int lrr_est_set_flags(lrr_est* e, int normalize, int wbs) {
  touch(e);
  if (!e) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  e->normalize = normalize != 0; e->strength = wbs != 0;
  if (!e->normalize) {
    int rc = pull_scal(e, dmask(e)); if (rc != LRR_OK) return rc;
    for (int d = 0; d < (e->sym() ? 2 : 1); ++d) { set_identity_norm(e->h[d]); if ((rc = push_scal(e, d)) != LRR_OK) return rc; }
  }
  return LRR_OK;
}

extern "C" int lrr_est_set_order(lrr_est* e, int dir, const int32_t* perm, int64_t n) {
  touch(e);
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  void_certificates(e);
  if (!perm) { d.has_order = false; return LRR_OK; }
  if (n != d.n_in) return LRR_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i) if (perm[i] < 0 || perm[i] >= n) return LRR_ERR_INVALID;
  CK(cudaMemcpyAsync(d.order, perm, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  d.has_order = true; d.perm_half = 0;
  return LRR_OK;
}
extern "C" int lrr_est_set_shuffle(lrr_est* e, int mode, uint64_t seed) {
  touch(e);
  if (!e || (mode != LRR_ORDER_FIXED && mode != LRR_ORDER_RESHUFFLE)) return LRR_ERR_INVALID;
  if (mode != e->shuffle_mode) void_certificates(e);
  e->shuffle_mode = mode; e->seed = seed;
  for (int d = 0; d < 2; ++d) { e->d[d].shuffle_calls = 0; e->d[d].perm_half = 0; }
  return LRR_OK;
}
extern "C" int lrr_est_set_initial_scale(lrr_est* e, int dir, double s) {
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  e->d[dir].initial_scale = s; return LRR_OK;
}
extern "C" int lrr_est_set_number_points(lrr_est* e, int dir, int64_t m, int* at_max) {
  touch(e);
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0) || m < 0) return LRR_ERR_INVALID;
  if (e->icp()) return LRR_ERR_INVALID;   // cdcl_estimation_ICP has no set_number_points (cdcl_estimation_ICP.h)
  lrr_dir& d = e->d[dir];
  bool r;
  if (m >= d.n_in) { d.number_points = d.n_in; r = true; } else { d.number_points = m; r = false; }   // cdcl :95-108
  if (at_max) *at_max = r;
  return LRR_OK;
}

// ---- MUSE table (rrel_muset_obj(0,false); Miller & Stewart, CVPR'96; trimmed statistics of the half-normal) ----
static double gauss_inv_cdf(double p) {   // Acklam's rational approximation + one Halley step
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
  double x;
  if (p < 0.02425) { double q = std::sqrt(-2 * std::log(p)); x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
  else if (p <= 1 - 0.02425) { double q = p - 0.5, r = q * q; x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1); }
  else { double q = std::sqrt(-2 * std::log(1 - p)); x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
  const double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
  const double u = e * std::sqrt(2 * M_PI) * std::exp(x * x / 2);
  return x - u / (1 + x * u / 2);
}
static int build_muse_table(unsigned n, double* table /* 4 x 32 */) {
  int rows = 0;
  for (double frac = 0.25; frac <= 0.95 + 0.00001; frac += 0.05) {
    long kk = std::lround(frac * n);
    unsigned k = (kk <= 0) ? 1u : (unsigned)kk; if (k > n) k = n;
    const double ek = gauss_inv_cdf(0.5 * (1.0 + double(k) / double(n + 1)));
    const double pk = double(k) / double(n + 1), qk = 1.0 - pk;
    const double Qk = ek, pQk = std::sqrt(2 / M_PI) * std::exp(-0.5 * Qk * Qk);
    const double Q1 = 1.0 / pQk, Q2 = Qk / (pQk * pQk), Q3 = (1.0 + 2.0 * Qk * Qk) / (pQk * pQk * pQk);
    const double vrk = (pk * qk / double(n + 2)) * Q1 * Q1 +
                       (pk * qk / (double(n + 2) * double(n + 2))) * (2.0 * (qk - pk) * Q1 * Q2 + pk * qk * (Q1 * Q3 + 0.5 * Q2 * Q2));
    const double sd = std::sqrt(vrk);
    if (ek / sd < 3.0) continue;
    const double mean = gauss_inv_cdf(0.5 * (1.0 + double(k) / double(n)));
    const double div = double(n) * std::sqrt(2 / M_PI) * (1.0 - std::exp(-0.5 * mean * mean));
    table[4 * rows] = (double)k; table[4 * rows + 1] = ek; table[4 * rows + 2] = sd; table[4 * rows + 3] = div;
    if (++rows == 32) break;
  }
  return rows;
}

// ---- Feistel cycle-walking permutation: the RESHUFFLE policy (stand-in for vcl_random_shuffle, Q4) --------
static inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static int apply_shuffle(lrr_est* e, int dir) {
  lrr_dir& d = e->d[dir];
  if (e->shuffle_mode != LRR_ORDER_RESHUFFLE) { d.perm_half = 0; return LRR_OK; }
  uint64_t st = e->seed + (uint64_t)dir * 0x632BE59BD9B4E019ULL + d.shuffle_calls * 0xD1342543DE82EF95ULL;
  d.shuffle_calls += 1;
  for (int r = 0; r < 4; ++r) d.perm_key[r] = (unsigned int)(splitmix64(st) >> 32);
  int bits = 2; while ((1ull << bits) < (uint64_t)d.n_in) ++bits;
  if (bits & 1) ++bits;
  d.perm_half = bits / 2;
  d.has_order = false;
  return LRR_OK;
}

// a fresh stamp epoch for direction d (the tag of the "unique matched to-feature" stamps, DirArgs::tag)
static int bump_epoch(lrr_est* e, int d) {
  lrr_ctx* ctx = e->ctx;
  e->d[d].epoch += 1;
  if (e->d[d].epoch == 0xFFFFFFFFu) {   // stamp epochs exhausted: start over
    CK(cudaMemsetAsync(e->d[d].stamp, 0xFF, (size_t)e->d[d].to_all->n * sizeof(unsigned long long), ctx->stream));
    e->d[d].epoch = 1;
  }
  return LRR_OK;
}
// find_closest_euclidean of every direction (cdcl :541-654, sym wrappers): host part (permutation keys, stamp epochs)
static int prepare_find_closest(lrr_est* e) {
  lrr_ctx* ctx = e->ctx;
  const int nd = e->sym() ? 2 : 1;
  for (int d = 0; d < nd; ++d) {
    if (e->d[d].to_all->n == 0 || e->d[d].n_in == 0) { ctx->err = "find_closest_euclidean: empty feature set"; return LRR_ERR_STATE; }
    int rc = apply_shuffle(e, d); if (rc != LRR_OK) return rc;
    if ((rc = bump_epoch(e, d)) != LRR_OK) return rc;
    e->d[d].cdc_active = false;   // a new 1-NN match replaces the many-to-many set of the CDC E-step
  }
  return LRR_OK;
}
static void finish_find_closest(lrr_est* e, const EstArgs& a) {
  const int nd = e->sym() ? 2 : 1;
  for (int d = 0; d < nd; ++d) { e->d[d].matches = a.d[d].M; e->d[d].weights_fresh = true; }
  e->werr_cache_valid = false;
}
static int do_find_closest(lrr_est* e, bool cert_only = false) {
  lrr_ctx* ctx = e->ctx;
  int rc = prepare_find_closest(e); if (rc != LRR_OK) return rc;
  EstArgs a = est_args(e);
  { PHASE(PH_MATCH); lrr_k_match(a, ctx->stream, &ctx->launches, cert_only); }
  finish_find_closest(e, a);
  return LRR_OK;
}

// Every direction of the next find_closest goes through the certificate pass (finest level, identity order, certificates on):
// the condition under which Scal::match_left means "queries not kept", and the structural half of the speculative plan.
static bool cert_path(const lrr_est* e) {
  const bool no_proof = !e->ctx->opt.certificates, no_tile = !e->ctx->opt.tile;
  if (no_proof || no_tile || e->icp() || e->shuffle_mode == LRR_ORDER_RESHUFFLE) return false;
  for (int d = 0; d < (e->sym() ? 2 : 1); ++d) {
    const lrr_dir& D = e->d[d];
    if (D.n_in <= 0 || D.number_points < D.n_in || D.has_order || D.perm_half) return false;
  }
  return true;
}
// SPECULATIVE PLAN.  When the last iteration kept every query of every direction (steady_hint, read back with the scalars) the next
// one is launched as certificate pass -> weight -> accumulate -> resid: the staged search, pass B and both normalisation kernels
// would find nothing to do (4 launches, ~9 us of a ~92 us iteration).  The bet is verified on the device: weight / collect /
// accumulate / resid first check that no direction left a query over and that the normalisation is clean (lrr_spec_bad) and commit
// nothing otherwise; the host then sees LRR_ST_NOT_STEADY, completes the match (lrr_k_match_rest) and runs the ordinary plan.
// LRR_SPEC_FORCE=1 bets in every iteration the certificate pass can run (tests: the fallback in every non-steady iteration),
// LRR_NO_SPEC=1 never.
static bool spec_ok(const lrr_est* e) {
  const bool off = e->ctx->opt.spec == 0, force = e->ctx->opt.spec == 2;
  if (off || !e->ctx->mailbox_dev || !e->normalize || e->dof == 6 || !e->norm_token || !cert_path(e)) return false;
  for (int d = 0; d < (e->sym() ? 2 : 1); ++d) if (!e->d[d].scale_estimated || e->d[d].matches <= 0) return false;
  return force || e->steady_hint;
}

static int muse_for_dir(lrr_est* e, const EstArgs& a, int d) {
  lrr_ctx* ctx = e->ctx;
  const int M = a.d[d].M;
  int rc = lrr_ctx_need_sort(ctx, (size_t)std::max(M, 1)); if (rc != LRR_OK) return rc;
  lrr_k_resid_keys(a.d[d], ctx->sort_keys, ctx->stream, &ctx->launches);
  lrr_sort_doubles(&ctx->sort, ctx->sort_keys, M, ctx->stream, &ctx->launches);
  double table[4 * 32];
  const int rows = build_muse_table((unsigned)M, table);
  CK(cudaMemcpyAsync(ctx->muse_table, table, sizeof(double) * 4 * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // `table` is a stack buffer
  lrr_muse_from_sorted(ctx->sort_keys, M, ctx->muse_table, rows, &e->d[d].sc->scale, ctx->stream, &ctx->launches);
  return LRR_OK;
}

// estimate_LS (cdcl :1009-1093, cdcl sym :398-701, rrl :783-859, rrl sym :699-918) on the device
static int do_estimate(lrr_est* e, int src, int estimate_parameters, bool force_hist = false, bool cdcl_backward = false, bool spec = false) {
  lrr_ctx* ctx = e->ctx;
  const int nd = e->sym() ? 2 : 1;
  EstArgs a = est_args(e, src);
  // stale weights (not straight after find_closest): residuals under the current transform first (cdcl :698-701)
  int stale = 0;
  for (int d = 0; d < nd; ++d) if (!e->d[d].weights_fresh) stale |= 1 << d;
  // The "unique matched to-feature" stamps exist already only when this estimate follows its own find_closest at a ramp level
  // (DirArgs::stamp_in_match).  Everywhere else — certified finest level, estimate_LS_backward — the normalisation lays them down
  // itself (k_restamp), in a fresh epoch.
  const bool restamp = e->normalize && !spec && (!a.d[src].stamp_in_match || stale || src != 0 || cdcl_backward);
  if (restamp) {
    int rc = bump_epoch(e, src); if (rc != LRR_OK) return rc;
    a = est_args(e, src);
  }
  a.estimate_parameters = estimate_parameters;
  if (cdcl_backward) a.flags |= LRR_F_NOSPATIAL | LRR_F_NO_OTHER_RENORM;
  for (int d = 0; d < nd; ++d) if (e->d[d].matches <= 0 || a.d[d].M <= 0) { ctx->err = "estimate_LS without matches"; return LRR_ERR_STATE; }
  // estimate_LS works on the matches that EXIST (the reference iterates over matches_): number_points_ may have moved on since the
  // last find_closest (a ramp step at the end of one_iteration, then estimate_LS_backward)
  for (int d = 0; d < nd; ++d) a.d[d].M = (int)e->d[d].matches;
  if (stale) { PHASE(PH_RESID); lrr_k_resid(a, stale, ctx->stream, &ctx->launches); a.flags |= LRR_F_USE_OLD_W; }
  if (cdcl_backward && e->dof == 6) {
    // rigid: make_incremental on matches whose from_ are mapped clones already (cdcl sym :708-731): mapped once more, through the
    // current global transform (the residuals above are those of estimate_scale_and_assign_weight on exactly these clones)
    lrr_k_map_clones_again(e->d[src].sc, ctx->stream, &ctx->launches);
    if (estimate_parameters) a.flags |= LRR_F_INV_OF_INCREMENT;
  }
  int mode[2] = {0, 0}; double sin[2] = {0, 0};
  for (int d = 0; d < nd; ++d) {
    lrr_dir& D = e->d[d];
    if (!D.scale_estimated) {
      if (std::isnan(D.initial_scale)) { PHASE(PH_MUSE); int rc = muse_for_dir(e, a, d); if (rc != LRR_OK) return rc; mode[d] = 1; }
      else { mode[d] = 2; sin[d] = D.initial_scale; }
      D.scale_estimated = true;
    }
  }
  if (e->normalize && cdcl_backward) {
    if (restamp) { PHASE(PH_RESTAMP); lrr_k_restamp(a, ctx->stream, &ctx->launches, 0); }
    PHASE(PH_NORM1); lrr_k_norm_merged(a, ctx->stream, &ctx->launches);
  }
  else if (e->normalize && !spec) {   // (speculative plan: the normalisation is clean or the kernels below refuse)
    // (rigid increments normalise mapped clones of the from-features, which move with the transform: never skipped)
    const bool no_skip = !ctx->opt.norm_skip;
    const int skip_clean = (e->norm_token && src == 0 && e->dof != 6 && !stale && !no_skip) ? 1 : 0;
    if (restamp) { PHASE(PH_RESTAMP); lrr_k_restamp(a, ctx->stream, &ctx->launches, skip_clean); }
    { PHASE(PH_NORM1); lrr_k_norm1(a, ctx->stream, &ctx->launches, skip_clean); }
    { PHASE(PH_NORM2); lrr_k_norm2(a, ctx->stream, &ctx->launches, skip_clean); }
  }
  e->norm_token = e->normalize && !cdcl_backward && src == 0 && e->dof != 6;
  // order statistics: window around the previous iteration's values when this estimator was the last user of the
  // context's selectors with the same configuration, else (first estimate, other estimator) histograms + collect pass
  // (the signature also covers the point counts: a ramp step resamples the keys, so the first estimate after it uses histograms)
  const unsigned long long sel_sig = (unsigned long long)(uintptr_t)e ^ ((unsigned long long)a.flags << 48) ^ ((unsigned long long)src << 40) ^
                                     ((unsigned long long)a.d[0].M * 0x9E3779B97F4A7C15ULL) ^ ((unsigned long long)(nd == 2 ? a.d[1].M : 0) * 0xC2B2AE3D27D4EB4FULL);
  // (small problems are latency-bound either way, and while the transform still moves fast the medians move too: after a
  //  miss the window is not tried again for a growing number of estimates)
  const int maxM = std::max(a.d[0].M, nd == 2 ? a.d[1].M : 0);
  bool window = !force_hist && !cdcl_backward && !stale && ctx->sel_owner == sel_sig && (e->strength || e->sym()) && maxM >= 16384 && !e->icp();
  if (window && e->window_holdoff > 0) { e->window_holdoff -= 1; window = false; }
  if (window) a.flags |= LRR_F_BRACKET;
  ctx->sel_owner = sel_sig;
  { PHASE(PH_WEIGHT); lrr_k_weight_ex(a, mode, sin, ctx->stream, &ctx->launches, spec ? 1 : 0); }
  if ((e->strength || (e->sym() && !cdcl_backward)) && !window && !e->icp()) { PHASE(PH_COLLECT); lrr_k_collect(a, ctx->stream, &ctx->launches, spec ? 1 : 0); }
  { PHASE(PH_ACCUM); lrr_k_accumulate(a, ctx->stream, &ctx->launches, spec ? 1 : 0); }
  for (int d = 0; d < nd; ++d) e->d[d].weights_fresh = false;
  e->werr_cache_valid = false;
  return LRR_OK;
}

// PIPELINED FINEST LEVEL.  The pass that ends an iteration (residuals of the matches under the new transform) and the certificate
// pass that starts the next find_closest read the same data and evaluate the same e^T P e under the same transform, so at the
// certified finest level of an affine estimator ONE launch does both (k_match_cert with LRR_CERT_RESID, lrr_tile.cu): the next
// one_iteration then starts at the weights, and the host already knows from this iteration's scalars whether every query was kept
// (steady state: certificate/residual pass -> weight -> accumulate+solve, three launches) or the match has to be completed first
// (staged search + pass B, then the normalisation).  LRR_NO_PIPE=1 switches it off (A/B).
static bool pipe_ok(const lrr_est* e) {
  const bool off = !e->ctx->opt.pipe;
  if (off || e->dof != 12 || !cert_path(e)) return false;
  for (int d = 0; d < (e->sym() ? 2 : 1); ++d) if (e->d[d].matches != e->d[d].n_in) return false;   // the finished iteration matched every point
  return true;
}

// estimate_LS + residuals under the new transform + scalar read-back.  A miss of the order-statistic window (rare: the
// key distribution jumped) leaves the transform untouched on the device; the estimate is then repeated with histograms.
static int estimate_and_pull(lrr_est* e, int resid_mask, bool spec = false) {
  lrr_ctx* ctx = e->ctx;
  const int nd = e->sym() ? 2 : 1, mask = dmask(e);
  bool fresh0[2], se0[2];
  for (int d = 0; d < 2; ++d) { fresh0[d] = e->d[d].weights_fresh; se0[d] = e->d[d].scale_estimated; }
  const int holdoff0 = e->window_holdoff; const unsigned long long owner0 = ctx->sel_owner;
  const bool used_cert = cert_path(e);   // this iteration's find_closest went through the certificate pass in every direction
  const bool pipe = pipe_ok(e);          // ... and the trailing pass is the certificate pass of the next one
  int rc = do_estimate(e, 0, 1, false, false, spec); if (rc != LRR_OK) return rc;
  static_assert(kScalHead % 8 == 0 && kScalHead / 8 <= 500, "Scal head does not fit a mailbox slot");
  auto resid_and_pull = [&](bool sp) -> int {
    const unsigned long long seq = ctx->mailbox_dev ? ++ctx->mailbox_seq : 0;
    if (pipe) {
      int r = prepare_find_closest(e); if (r != LRR_OK) return r;
      EstArgs a = est_args(e); PHASE(PH_RESID);
      lrr_k_match_cert(a, ctx->stream, &ctx->launches, 1 /* LRR_CERT_RESID */, ctx->mailbox_dev, seq, mask, (int)(kScalHead / 8));
    } else {
      EstArgs a = est_args(e); PHASE(PH_RESID);
      lrr_k_resid(a, resid_mask, ctx->stream, &ctx->launches, ctx->mailbox_dev, seq, mask, (int)(kScalHead / 8), sp ? 1 : 0);
    }
    return ctx->mailbox_dev ? pull_scal_mailbox(e, mask, seq) : pull_scal(e, mask);
  };
  if ((rc = resid_and_pull(spec)) != LRR_OK) return rc;
  auto restore = [&](int bit) -> int {
    for (int d = 0; d < nd; ++d) {
      e->h[d].status &= ~bit;
      int r = push_scal(e, d); if (r != LRR_OK) return r;
      e->d[d].weights_fresh = fresh0[d]; e->d[d].scale_estimated = se0[d];
    }
    return LRR_OK;
  };
  if (spec) {
    bool refused = false;
    for (int d = 0; d < nd; ++d) refused = refused || (e->h[d].status & LRR_ST_NOT_STEADY);
    if (refused) {   // the bet is lost: nothing was committed; complete the match and run the ordinary plan
      ctx->spec_misses += 1;
      e->steady_hint = false;
      if ((rc = restore(LRR_ST_NOT_STEADY)) != LRR_OK) return rc;
      e->window_holdoff = holdoff0; ctx->sel_owner = owner0;
      { EstArgs a = est_args(e); { PHASE(PH_SEARCH); lrr_k_match_search(a, ctx->stream, &ctx->launches); } { PHASE(PH_QUEUE); lrr_k_match_queue(a, ctx->stream, &ctx->launches); } }
      return estimate_and_pull(e, resid_mask, false);
    }
  }
  bool miss = false;
  for (int d = 0; d < nd; ++d) miss = miss || (e->h[d].status & LRR_ST_SELECT_MISS);
  if (miss) {
    ctx->window_misses += 1;
    e->window_backoff = std::min(2 * (e->window_backoff + 1), 16); e->window_holdoff = e->window_backoff;
    if ((rc = restore(LRR_ST_SELECT_MISS)) != LRR_OK) return rc;
    if ((rc = do_estimate(e, 0, 1, true, false, spec)) != LRR_OK) return rc;   // (a speculative plan was verified by the first attempt)
    if ((rc = resid_and_pull(spec)) != LRR_OK) return rc;
  } else if (e->window_holdoff == 0) e->window_backoff = 0;
  bool all_kept = used_cert;
  for (int d = 0; d < nd; ++d) all_kept = all_kept && e->h[d].match_left == 0;
  // pipelined: match_left is that of the NEXT find_closest (its certificate pass has just run); otherwise that of this iteration's,
  // the bet of the speculative plan
  e->steady_hint = all_kept;
  { double failed = 0, total = 0; for (int d = 0; d < nd; ++d) { failed += e->h[d].cert_failed; total += (double)e->d[d].n_in; }
    e->sparse_hint = used_cert && failed < 0.25 * total; }
  e->pre_matched = pipe;
  return LRR_OK;
}

// find_closest + estimate_LS + residuals + read-back of one iteration
static int match_estimate_pull(lrr_est* e, int resid_mask) {
  lrr_ctx* ctx = e->ctx;
  int rc;
  bool spec;
  if (e->pre_matched) {
    // the certificate pass of this find_closest ran as the trailing pass of the last iteration, and its outcome came back with that
    // iteration's scalars: every query kept (steady_hint) -> nothing to complete, and the normalisation is unchanged
    e->pre_matched = false;
    spec = spec_ok(e);
    EstArgs a = est_args(e);
    if (!spec) { { PHASE(PH_SEARCH); lrr_k_match_search(a, ctx->stream, &ctx->launches); } { PHASE(PH_QUEUE); lrr_k_match_queue(a, ctx->stream, &ctx->launches); } }
    finish_find_closest(e, a);
  } else {
    spec = spec_ok(e);
    if ((rc = do_find_closest(e, spec)) != LRR_OK) return rc;
  }
  return estimate_and_pull(e, resid_mask, spec);
}

static int check_status(lrr_est* e) {
  const int nd = e->sym() ? 2 : 1;
  for (int d = 0; d < nd; ++d) if (e->h[d].status & LRR_ST_SELECT_OVERFLOW) {
    e->ctx->err = "order-statistic selector overflow (more than LRR_SEL_CAP distinct keys in one histogram bin)";
    return LRR_ERR_STATE;
  }
  return LRR_OK;
}

static double rms_of(const lrr_est* e, const Scal& s, bool after) {   // RMS_error (cdcl :196-230; rrl :211-241, Q1)
  const double sum = after ? s.res_sum_r2 : s.used_sum_r2;   // (sum_r2 may already belong to the next match: pre-match, lrr_tile.cu)
  if (e->rrl()) return sum;
  const double cnt = s.res_count;
  return cnt != 0 ? std::sqrt(sum / cnt) : sum;
}
static double werr_of(const Scal& s) { return s.res_sum_w != 0 ? s.res_sum_wr / s.res_sum_w : s.res_sum_wr; }

void lrr_k_fill_ones(double* w, int n, cudaStream_t st, long long* launches);
extern "C" int lrr_est_find_closest(lrr_est* e) {
  if (!e) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  if (e->pre_matched) {   // the certificate pass has run already (trailing pass of the last iteration): complete the match
    e->pre_matched = false;
    EstArgs a = est_args(e);
    { PHASE(PH_SEARCH); lrr_k_match_search(a, ctx->stream, &ctx->launches); }
    { PHASE(PH_QUEUE); lrr_k_match_queue(a, ctx->stream, &ctx->launches); }
    finish_find_closest(e, a);
  } else {
    int rc = do_find_closest(e); if (rc != LRR_OK) return rc;
  }
  // find_closest resets every weight to 1 (cdcl :645); matches kept by their certificate are not rewritten, so where certificates
  // are in use the weights are reset here (inside one_iteration the weights are recomputed before anything reads them)
  EstArgs a = est_args(e);
  for (int d = 0; d < a.ndir; ++d) if (!a.d[d].stamp_in_match) lrr_k_fill_ones(a.d[d].w, a.d[d].M, ctx->stream, &ctx->launches);
  e->steady_hint = false;
  return LRR_OK;
}

static int one_iteration_fwd(lrr_est* e, int* converged) {
  int rc;
  if ((rc = match_estimate_pull(e, 1)) != LRR_OK) return rc;
  if ((rc = check_status(e)) != LRR_OK) return rc;
  const double before = rms_of(e, e->h[0], false), after = rms_of(e, e->h[0], true);
  *converged = std::fabs(before - after) < 1e-4;   // cdcl :506
  e->iterations_done += 1;
  return LRR_OK;
}

static int weighted_error_now(lrr_est* e, int dir, double* out) {
  lrr_ctx* ctx = e->ctx;
  if (e->d[dir].matches <= 0) { *out = 0.0; return LRR_OK; }
  EstArgs a = est_args(e);
  lrr_k_resid(a, 1 << dir, ctx->stream, &ctx->launches);
  int rc = pull_scal(e, 1 << dir); if (rc != LRR_OK) return rc;
  *out = werr_of(e->h[dir]);
  return LRR_OK;
}

// one_iteration of the symmetric estimators (cdcl sym :268-363, rrl sym :341-438)
static int one_iteration_sym(lrr_est* e, unsigned iteration, int* converged_out) {
  int rc;
  double before;
  if (e->werr_cache_valid) before = e->werr_cache;
  else if ((rc = weighted_error_now(e, 0, &before)) != LRR_OK) return rc;
  if ((rc = match_estimate_pull(e, 1)) != LRR_OK) return rc;
  if ((rc = check_status(e)) != LRR_OK) return rc;
  const double after = werr_of(e->h[0]);
  e->werr_cache = after; e->werr_cache_valid = true;
  const double curr = after - before;
  const double diff = (after - before) / after;
  bool converged = std::fabs(diff) < 1e-4;
  e->weighted_error = after;
  if (!e->rrl() || iteration > 0) {
    if (!e->finest_level && (e->oscillation_count > 1 || std::fabs(diff) < 0.01)) converged = true;
    else if (!converged) {
      if (after > 1.5) {
        if (curr * e->error_difference < 0.0) ++e->oscillation_count;
        else if (e->oscillation_count > 0) --e->oscillation_count;
        if (e->oscillation_count > 10) converged = true;
      } else {
        if (curr > 0.0) ++e->oscillation_count;
        if (e->oscillation_count > 3) converged = true;
      }
    }
  }
  e->error_difference = curr;
  if (converged && !e->finest_level) {
    converged = false;
    e->number_matches *= 2;
    int fmax = 0, bmax = 0;
    lrr_est_set_number_points(e, 0, e->number_matches, &fmax);
    lrr_est_set_number_points(e, 1, e->number_matches, &bmax);
    e->finest_level = fmax && bmax;
    e->oscillation_count = 0; e->error_difference = 0.0;
  }
  e->iterations_done += 1;
  *converged_out = converged && e->finest_level;
  return LRR_OK;
}

extern "C" int lrr_est_one_iteration(lrr_est* e, unsigned iteration, int* converged) {
  if (!e) return LRR_ERR_INVALID;
  enter(e->ctx);
  int c = 0;
  int rc = e->sym() ? one_iteration_sym(e, iteration, &c) : one_iteration_fwd(e, &c);
  if (converged) *converged = c;
  return rc;
}
extern "C" int lrr_est_run(lrr_est* e, int max_iter, int* converged, int* iterations) {
  if (!e) return LRR_ERR_INVALID;
  int c = 0, it = 0, rc = LRR_OK;
  while (!c && it < max_iter) { rc = lrr_est_one_iteration(e, (unsigned)it, &c); if (rc != LRR_OK) break; ++it; }
  if (converged) *converged = c;
  if (iterations) *iterations = it;
  return rc;
}

extern "C" int lrr_est_get_transform(lrr_est* e, int dir, int* tk, double* A9, double* t3, double* c3, double* cov) {
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  enter(e->ctx);
  int rc = pull_scal(e, 1 << dir); if (rc != LRR_OK) return rc;
  const Scal& s = e->h[dir];
  if (tk) *tk = e->dof == 12 ? LRR_TRANS_AFFINE : (e->dof == 6 ? LRR_TRANS_RIGID3D : LRR_TRANS_SIMILARITY2D);
  if (A9) memcpy(A9, s.A, 72); if (t3) memcpy(t3, s.t, 24); if (c3) memcpy(c3, s.c, 24);
  if (cov) memcpy(cov, s.cov, sizeof(double) * e->dof * e->dof);
  return LRR_OK;
}

extern "C" int lrr_est_get_matches(lrr_est* e, int dir, int64_t cap, int64_t* n, int32_t* from_idx, int32_t* to_idx, double* w) {
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  const int64_t m = d.matches;
  if (n) *n = m;
  const int64_t c = std::min(m, cap);
  if (c <= 0 || (!from_idx && !to_idx && !w)) return LRR_OK;
  EstArgs a = est_args(e);
  a.d[dir].M = (int)m;
  int* tmp = nullptr;
  CK(dmalloc(&tmp, (size_t)2 * m));
  lrr_k_export_matches(a.d[dir], tmp, tmp + m, ctx->stream, &ctx->launches);
  if (from_idx) CK(cudaMemcpyAsync(from_idx, tmp, sizeof(int) * c, cudaMemcpyDeviceToHost, ctx->stream));
  if (to_idx) CK(cudaMemcpyAsync(to_idx, tmp + m, sizeof(int) * c, cudaMemcpyDeviceToHost, ctx->stream));
  if (w) CK(cudaMemcpyAsync(w, d.w, sizeof(double) * c, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  dfree(tmp);
  return LRR_OK;
}

extern "C" int lrr_est_metric(lrr_est* e, int which, int dir, double* out) {
  if (!e || !out || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  EstArgs a = est_args(e);
  a.d[dir].M = (int)d.matches;   // metrics run over the existing matches
  int rc;
  switch (which) {
    case LRR_METRIC_RMS:
    case LRR_METRIC_WEIGHTED: {
      if (d.matches <= 0) { *out = 0.0; return LRR_OK; }
      lrr_k_resid(a, 1 << dir, ctx->stream, &ctx->launches);
      if ((rc = pull_scal(e, 1 << dir)) != LRR_OK) return rc;
      *out = which == LRR_METRIC_RMS ? rms_of(e, e->h[dir], true) : werr_of(e->h[dir]);
      return LRR_OK;
    }
    case LRR_METRIC_MEDIAN: {
      if (d.matches <= 0) { *out = 0.0; return LRR_OK; }
      const int M = (int)d.matches;
      lrr_k_resid(a, 1 << dir, ctx->stream, &ctx->launches);
      if ((rc = lrr_ctx_need_sort(ctx, (size_t)M)) != LRR_OK) return rc;
      lrr_k_resid_keys(a.d[dir], ctx->sort_keys, ctx->stream, &ctx->launches);
      lrr_sort_doubles(&ctx->sort, ctx->sort_keys, M, ctx->stream, &ctx->launches);
      int loc = (M - 1) / 2 + 1; if (loc >= M) loc = M - 1;   // cdcl :186-188
      CK(cudaMemcpyAsync(ctx->pinned, ctx->sort_keys + loc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      *out = *(double*)ctx->pinned;
      return LRR_OK;
    }
    case LRR_METRIC_SHEET_ANGLES:
    case LRR_METRIC_TUBE_ANGLES: {
      if (d.matches <= 0) { ctx->err = "angles without matches"; return LRR_ERR_STATE; }
      fset_wait_dirs(e->d[dir].from_all, ctx->stream); fset_wait_dirs(e->d[dir].to_all, ctx->stream);
      lrr_k_angles(a.d[dir], a.flags, which == LRR_METRIC_SHEET_ANGLES ? 0 : 1, ctx->partials, ctx->ticket, ctx->stream, &ctx->launches);
      CK(cudaMemcpyAsync(ctx->pinned, &d.sc->ang_num, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      const double* p = (const double*)ctx->pinned;
      *out = p[0] / p[1];
      return LRR_OK;
    }
    default: return LRR_ERR_INVALID;
  }
}

extern "C" int lrr_est_estimate_ls_backward(lrr_est* e, int estimate_parameters) {
  if (!e || !e->sym()) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  touch(e);   // (the transforms go through normalize / unnormalize at least: a pre-match of the next find_closest is dropped)
  if (e->kind == LRR_CDCL_SYM) {
    // cdcl variant (cdcl sym :705-863): the backward matches plus the forward matches reversed, normalised as ONE list,
    // strength re-weighting over the merged list, no spatial weighting; only backward_->trans_ is (un)normalised.
    int rc = do_estimate(e, 1, estimate_parameters, true, true); if (rc != LRR_OK) return rc;
    if ((rc = pull_scal(e, 3)) != LRR_OK) return rc;
    return check_status(e);
  }
  // rrl: swap forward_/backward_, estimate_LS, swap back (rrl sym :922-934)
  int rc = do_estimate(e, 1, estimate_parameters, true); if (rc != LRR_OK) return rc;
  if ((rc = pull_scal(e, 3)) != LRR_OK) return rc;
  return check_status(e);
}

// compute_transfer_error_covariance: covarJ = mean_i J_i Sigma J_i^T  (cdcl :388-404, rrl :381-397)
extern "C" int lrr_est_transfer_cov(lrr_est* e, int dir, double* S9) {
  if (!e || !S9 || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  if (d.matches <= 0) { ctx->err = "transfer covariance without matches"; return LRR_ERR_STATE; }
  EstArgs a = est_args(e);
  a.d[dir].M = (int)d.matches;
  lrr_k_transfer_cov(a.d[dir], e->dof, ctx->partials, ctx->ticket, ctx->stream, &ctx->launches);
  int rc = pull_scal(e, 1 << dir); if (rc != LRR_OK) return rc;
  double m10[10];
  CK(cudaMemcpyAsync(ctx->pinned, d.sc->covJ, 10 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(m10, ctx->pinned, sizeof(m10));
  // M[a][b] = sum x~_a x~_b
  double Mm[4][4];
  Mm[0][0] = m10[0]; Mm[0][1] = Mm[1][0] = m10[1]; Mm[0][2] = Mm[2][0] = m10[2]; Mm[0][3] = Mm[3][0] = m10[3];
  Mm[1][1] = m10[4]; Mm[1][2] = Mm[2][1] = m10[5]; Mm[1][3] = Mm[3][1] = m10[6]; Mm[2][2] = m10[7]; Mm[2][3] = Mm[3][2] = m10[8]; Mm[3][3] = m10[9];
  const int dof = e->dof; const double* Sg = e->h[dir].cov;
  // J[dd][i] = sum_a x~_a E[a][dd][i]
  auto E = [&](int a_, int dd, int i) -> double {
    if (dof == 12) { const int di = i < 9 ? i / 3 : i - 9, ai = i < 9 ? i % 3 : 3; return (di == dd && ai == a_) ? 1.0 : 0.0; }
    if (dof == 4) {   // similarity2d: J = [x -y 1 0; y x 0 1]
      if (dd > 1) return 0.0;
      if (i == 0) return a_ == dd ? 1.0 : 0.0;
      if (i == 1) return (dd == 0 && a_ == 1) ? -1.0 : ((dd == 1 && a_ == 0) ? 1.0 : 0.0);
      return (a_ == 3 && dd == i - 2) ? 1.0 : 0.0;
    }
    if (i >= 3) return (a_ == 3 && dd == i - 3) ? 1.0 : 0.0;
    if (a_ == 3 || dd == i || a_ != 3 - dd - i) return 0.0;
    return ((dd == 2 && i == 0) || (dd == 0 && i == 1) || (dd == 1 && i == 2)) ? 1.0 : -1.0;
  };
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    double s = 0;
    for (int a_ = 0; a_ < 4; ++a_) for (int i = 0; i < dof; ++i) { const double ea = E(a_, r, i); if (ea == 0) continue;
      for (int b_ = 0; b_ < 4; ++b_) for (int j = 0; j < dof; ++j) { const double eb = E(b_, c, j); if (eb == 0) continue; s += Mm[a_][b_] * ea * eb * Sg[i * dof + j]; } }
    S9[r * 3 + c] = s * (1.0 / (double)d.matches);
  }
  return LRR_OK;
}

extern "C" int lrr_est_get_state(lrr_est* e, lrr_est_state* s) {
  if (!e || !s) return LRR_ERR_INVALID;
  enter(e->ctx);
  int rc = pull_scal(e, dmask(e)); if (rc != LRR_OK) return rc;
  memset(s, 0, sizeof(*s));
  const int nd = e->sym() ? 2 : 1;
  for (int d = 0; d < nd; ++d) {
    s->number_points[d] = std::min<int64_t>(e->d[d].number_points, e->d[d].n_in);
    s->scale[d] = e->h[d].scale; s->scale_estimated[d] = e->d[d].scale_estimated;
  }
  for (int k = 0; k < 3; ++k) { s->norm_center_moving[k] = e->h[0].cfrom[k]; s->norm_center_fixed[k] = e->h[0].cto[k]; }
  s->norm_radius_moving = e->h[0].rfrom; s->norm_radius_fixed = e->h[0].rto;
  if (e->sym()) {
    s->number_matches = e->number_matches; s->finest_level = e->finest_level; s->oscillation_count = (int)e->oscillation_count;
    s->error_difference = e->error_difference; s->weighted_error = e->weighted_error;
  }
  s->iterations_done = e->iterations_done;
  return LRR_OK;
}
extern "C" int lrr_est_get_stats(lrr_est* e, lrr_est_stats* s) {
  if (!e || !s) return LRR_ERR_INVALID;
  enter(e->ctx);
  int rc = pull_scal(e, dmask(e)); if (rc != LRR_OK) return rc;
  memset(s, 0, sizeof(*s));
  for (int d = 0; d < (e->sym() ? 2 : 1); ++d) {
    s->cert_failed[d] = (int64_t)e->h[d].cert_failed; s->passb_queries[d] = (int64_t)e->h[d].passb_queries; s->work_list_entries[d] = e->h[d].match_left;
  }
  s->window_misses = e->ctx->window_misses; s->spec_misses = e->ctx->spec_misses;
  return LRR_OK;
}
extern "C" int lrr_est_oscillated(lrr_est* e, int* out) {
  if (!e || !out) return LRR_ERR_INVALID;
  *out = e->sym() ? (e->oscillation_count > 10) : 0;   // sym .h:156
  return LRR_OK;
}

// ---- covariance objective (A13) ---------------------------------------------------------------------------
void lrr_k_objcov_launch(const DirArgs& a, const double* A, const double* t, const double* c, const double* norm8, const double* Sigma_dev,
                         double* partials, unsigned int* ticket, cudaStream_t st, long long* launches);

struct HostTrans { double A[9], t[3], c[3], cov[144]; };
static void mat_BCBt12(const double* B, const double* C, double* R) {
  double T[144];
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { double s = 0; for (int k = 0; k < 12; ++k) s += B[i * 12 + k] * C[k * 12 + j]; T[i * 12 + j] = s; }
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { double s = 0; for (int k = 0; k < 12; ++k) s += T[i * 12 + k] * B[j * 12 + k]; R[i * 12 + j] = s; }
}
// cdcl_trans_affine::normalize / unnormalize (cdcl/cdcl_trans_affine.txx:220-286)
static void affine_normalize(HostTrans& T, double rm, double rf, const double* cm, const double* cf) {
  const double s = rm / rf;
  for (int i = 0; i < 9; ++i) T.A[i] = s * T.A[i];
  const double dc[3] = {cm[0] - T.c[0], cm[1] - T.c[1], cm[2] - T.c[2]};
  for (int r = 0; r < 3; ++r) {
    double acc = T.A[3 * r] * dc[0]; acc += T.A[3 * r + 1] * dc[1]; acc += T.A[3 * r + 2] * dc[2];
    T.t[r] = T.t[r] / rf + acc / rm - cf[r] / rf;
  }
  double B[144] = {0};
  for (int d = 0; d < 9; ++d) B[d * 12 + d] = s;
  for (int d1 = 0; d1 < 3; ++d1) for (int d2 = 0; d2 < 3; ++d2) B[(9 + d1) * 12 + d1 * 3 + d2] = dc[d2] / rf;
  for (int d = 9; d < 12; ++d) B[d * 12 + d] = 1.0 / rf;
  double R[144]; mat_BCBt12(B, T.cov, R); memcpy(T.cov, R, sizeof(R));
  T.c[0] = T.c[1] = T.c[2] = 0.0;
}
static void affine_unnormalize(HostTrans& T, double rm, double rf, const double* cm, const double* cf) {
  const double s = rf / rm;
  for (int i = 0; i < 9; ++i) T.A[i] = s * T.A[i];
  for (int r = 0; r < 3; ++r) T.t[r] = rf * T.t[r] + cf[r];
  double B[144] = {0};
  for (int d = 0; d < 9; ++d) B[d * 12 + d] = rf / rm;
  for (int d = 9; d < 12; ++d) B[d * 12 + d] = rf;
  double R[144]; mat_BCBt12(B, T.cov, R); memcpy(T.cov, R, sizeof(R));
  for (int d = 0; d < 3; ++d) T.c[d] = cm[d];
}
static bool cholesky_upper12(const double* C, double* U) {   // C = U^T U  (vnl_cholesky::upper_triangle)
  for (int i = 0; i < 144; ++i) U[i] = 0.0;
  for (int i = 0; i < 12; ++i) for (int j = i; j < 12; ++j) {
    double s = C[i * 12 + j];
    for (int k = 0; k < i; ++k) s -= U[k * 12 + i] * U[k * 12 + j];
    if (i == j) { if (s <= 0) return false; U[i * 12 + i] = std::sqrt(s); } else U[i * 12 + j] = s / U[i * 12 + i];
  }
  return true;
}

// the flat many-to-many set of the CDC E-step as the match arrays the objective kernels stream (position -> from index, slot, weight)
static void cdc_dir_args(const lrr_dir& d, DirArgs& a) {
  a.order = d.cdc_efrom; a.perm_half = 0; a.sidx = d.cdc_eslot; a.w = d.cdc_ew; a.M = (int)d.cdc_entries;
}
// cdcl_normalize_matches over the many-to-many set (cdcl/cdcl_estimation_ICP.txx:751-832) -> norm8 = (cm, rm, cf, rf)
static int cdc_normalisation(lrr_est* e, int dir, double* norm8) {
  lrr_ctx* ctx = e->ctx; lrr_dir& d = e->d[dir];
  double* out = ctx->scratch_d + 160;
  lrr_k_cdc_norm((int)d.n_in, (int)d.to_all->n, d.fx, d.fy, d.fz, d.cdc_rowcnt, d.cdc_tow, d.to_all->dev.grid.pos4, out, ctx->stream, &ctx->launches);
  double* pin = (double*)ctx->pinned + 512;
  CK(cudaMemcpyAsync(pin, out, 10 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 3; ++k) { norm8[k] = pin[k] / pin[3]; norm8[4 + k] = pin[4 + k] / pin[7]; }
  norm8[3] = pin[8] / pin[3]; norm8[7] = pin[9] / pin[7];
  return LRR_OK;
}

// one evaluation of f and g at x (length 78) for direction `dir`, transform / normalisation given by the caller
static int objcov_device_eval(lrr_est* e, int dir, const HostTrans& T, const double* norm8, const double* x, double* f, double* g) {
  lrr_ctx* ctx = e->ctx;
  lrr_dir& d = e->d[dir];
  double U[144] = {0}, Sg[144];
  int ind = 0;
  for (int d1 = 0; d1 < 12; ++d1) for (int d2 = d1; d2 < 12; ++d2) U[d1 * 12 + d2] = x[ind++];
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { double s = 0; for (int k = 0; k < 12; ++k) s += U[k * 12 + i] * U[k * 12 + j]; Sg[i * 12 + j] = s; }
  double* pin = (double*)ctx->pinned;
  memcpy(pin, Sg, sizeof(Sg));
  CK(cudaMemcpyAsync(ctx->scratch_d, pin, sizeof(Sg), cudaMemcpyHostToDevice, ctx->stream));
  EstArgs a = est_args(e);
  a.d[dir].M = (int)d.matches;
  if (d.cdc_active) cdc_dir_args(d, a.d[dir]);
  { PHASE(PH_OBJCOV); lrr_k_objcov_launch(a.d[dir], T.A, T.t, T.c, norm8, ctx->scratch_d, ctx->partials, ctx->ticket, ctx->stream, &ctx->launches); }
  CK(cudaMemcpyAsync(pin + 256, d.sc->mom, 100 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  const double* mom = pin + 256;
  const double n = d.cdc_active ? (double)d.cdc_entries : (double)d.matches;   // num_of_all_matches (cdcl_obj_fun_cov.txx:76,124)
  static const int sym6[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  if (f) { double fv = 0; for (int j = 0; j < 4; ++j) fv += mom[j * 25 + 24]; *f = fv / n; }
  if (g) {
    double kl[144];
    for (int i = 0; i < 12; ++i) for (int jx = 0; jx < 12; ++jx) {
      const int di = i < 9 ? i / 3 : i - 9, ai = i < 9 ? i % 3 : 3, dj = jx < 9 ? jx / 3 : jx - 9, aj = jx < 9 ? jx % 3 : 3;
      kl[i * 12 + jx] = mom[ai * 25 + aj * 6 + sym6[di][dj]];
    }
    ind = 0;   // g[ind] = < kl , U^T E + E^T U >   (cdcl_obj_fun_cov.txx:106-120)
    for (int d1 = 0; d1 < 12; ++d1) for (int d2 = d1; d2 < 12; ++d2) {
      double s = 0;
      for (int i = 0; i < 12; ++i) s += kl[i * 12 + d2] * U[d1 * 12 + i];
      for (int jx = 0; jx < 12; ++jx) s += kl[d2 * 12 + jx] * U[d1 * 12 + jx];
      g[ind++] = s * (1.0 / n);
    }
  }
  return LRR_OK;
}

static int objcov_check(lrr_est* e, int dir) {
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx;
  if (e->dof != 12) { ctx->err = "cdcl_obj_fun_cov: only the affine <3,12> instantiation is built"; return LRR_ERR_UNSUPPORTED; }
  if (e->d[dir].matches <= 0 && !e->d[dir].cdc_active) { ctx->err = "objcov without matches"; return LRR_ERR_STATE; }
  if (!e->d[dir].from_all->C6 || !e->d[dir].to_all->C6) { ctx->err = "objcov needs feature covariances (C6)"; return LRR_ERR_STATE; }
  return LRR_OK;
}

extern "C" int lrr_objcov_eval(lrr_est* e, int dir, const double* x, double* f, double* g) {
  touch(e);
  int rc = objcov_check(e, dir); if (rc != LRR_OK) return rc;
  if (!x) return LRR_ERR_INVALID;
  enter(e->ctx);
  if ((rc = pull_scal(e, 1 << dir)) != LRR_OK) return rc;
  HostTrans T; memcpy(T.A, e->h[dir].A, 72); memcpy(T.t, e->h[dir].t, 24); memcpy(T.c, e->h[dir].c, 24);
  const double norm8[8] = {0, 0, 0, 1, 0, 0, 0, 1};
  return objcov_device_eval(e, dir, T, norm8, x, f, g);
}

// L-BFGS (stands in for vnl_lbfgs: m = 5, backtracking Armijo line search) with cdcl_lbfgs::report_iter's early stop:
// finish as soon as |f_start - f| > 5 % |f_start| (cdcl/cdcl_lbfgs.h:30-36), at most max_evals evaluations
// (cdcl/cdcl_estimation.txx:299).  The iterate trajectory of the netlib routine is unpinned upstream.
template <class FG>
static int lbfgs_minimize(FG&& fg, std::vector<double>& x, int max_evals, double* f_start, double* f_end, int* evals_out) {
  const int n = (int)x.size(), mem = 5;
  std::vector<double> g(n), gn(n), xn(n), d(n);
  std::vector<std::vector<double>> S, Y; std::vector<double> RHO;
  double f = 0; int evals = 0, rc;
  if ((rc = fg(x.data(), &f, g.data())) != LRR_OK) return rc;
  ++evals; *f_start = f; *f_end = f;
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
  int iter = 0;
  while (evals < max_evals) {
    d = g;
    std::vector<double> al(S.size());
    for (int i = (int)S.size() - 1; i >= 0; --i) { al[i] = RHO[i] * dot(S[i], d); for (int k = 0; k < n; ++k) d[k] -= al[i] * Y[i][k]; }
    if (!S.empty()) { double gam = dot(S.back(), Y.back()) / dot(Y.back(), Y.back()); for (int k = 0; k < n; ++k) d[k] *= gam; }
    for (size_t i = 0; i < S.size(); ++i) { double be = RHO[i] * dot(Y[i], d); for (int k = 0; k < n; ++k) d[k] += S[i][k] * (al[i] - be); }
    for (int k = 0; k < n; ++k) d[k] = -d[k];
    const double gd = dot(g, d);
    if (!(gd < 0)) break;
    const double gnorm = std::sqrt(dot(g, g));
    double step = (iter == 0) ? 1.0 / (gnorm > 0 ? gnorm : 1.0) : 1.0;
    bool accepted = false; double fn = f;
    while (evals < max_evals) {
      for (int k = 0; k < n; ++k) xn[k] = x[k] + step * d[k];
      if ((rc = fg(xn.data(), &fn, gn.data())) != LRR_OK) return rc;
      ++evals;
      if (std::isfinite(fn) && fn <= f + 1e-4 * step * gd) { accepted = true; break; }
      step *= 0.5;
    }
    if (!accepted) break;
    std::vector<double> s(n), yv(n);
    for (int k = 0; k < n; ++k) { s[k] = xn[k] - x[k]; yv[k] = gn[k] - g[k]; }
    const double sy = dot(s, yv);
    if (sy > 1e-300) { S.push_back(s); Y.push_back(yv); RHO.push_back(1.0 / sy); if ((int)S.size() > mem) { S.erase(S.begin()); Y.erase(Y.begin()); RHO.erase(RHO.begin()); } }
    x = xn; g = gn; f = fn; *f_end = f; ++iter;
    if (std::fabs(*f_start - *f_end) > 0.05 * std::fabs(*f_start)) break;   // report_iter()
    if (std::sqrt(dot(g, g)) < 1e-12) break;
  }
  *evals_out = evals;
  return LRR_OK;
}

// estimate_covariance (cdcl/cdcl_estimation.txx:264-337) on the ICP matches of direction `dir`
extern "C" int lrr_objcov_refine(lrr_est* e, int dir, int max_evals, double* f_start, double* f_end, int* n_evals) {
  touch(e);
  int rc = objcov_check(e, dir); if (rc != LRR_OK) return rc;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  double norm8[8] = {0, 0, 0, 1, 0, 0, 0, 1};
  bool cdc_norm = false;
  if (e->normalize && d.cdc_active) { if ((rc = cdc_normalisation(e, dir, norm8)) != LRR_OK) return rc; cdc_norm = true; }
  else if (e->normalize) {   // cdcl_normalize_matches over the current matches (:271-277)
    if ((rc = bump_epoch(e, dir)) != LRR_OK) return rc;
    EstArgs a = est_args(e, dir);
    a.d[dir].M = (int)d.matches;
    a.ndir = 1; if (dir == 1) { a.d[0] = a.d[1]; a.src = 0; }
    lrr_k_restamp(a, ctx->stream, &ctx->launches);
    lrr_k_norm1(a, ctx->stream, &ctx->launches); lrr_k_norm2(a, ctx->stream, &ctx->launches);
  }
  if ((rc = pull_scal(e, 1 << dir)) != LRR_OK) return rc;
  Scal& h = e->h[dir];
  HostTrans T; memcpy(T.A, h.A, 72); memcpy(T.t, h.t, 24); memcpy(T.c, h.c, 24); memcpy(T.cov, h.cov, sizeof(T.cov));
  if (e->normalize) {
    if (!cdc_norm) {
      for (int k = 0; k < 3; ++k) { norm8[k] = h.cfrom[k]; norm8[4 + k] = h.cto[k]; }
      norm8[3] = h.rfrom; norm8[7] = h.rto;
    } else {   // (reported by lrr_est_get_state like the normalisation of any other estimate)
      for (int k = 0; k < 3; ++k) { h.cfrom[k] = norm8[k]; h.cto[k] = norm8[4 + k]; }
      h.rfrom = norm8[3]; h.rto = norm8[7];
    }
    affine_normalize(T, norm8[3], norm8[7], norm8, norm8 + 4);
  }
  double U[144];
  if (!cholesky_upper12(T.cov, U)) { ctx->err = "estimate_covariance: covariance is not positive definite"; return LRR_ERR_STATE; }
  std::vector<double> x(78); int ind = 0;
  for (int d1 = 0; d1 < 12; ++d1) for (int d2 = d1; d2 < 12; ++d2) x[ind++] = U[d1 * 12 + d2];
  double fs = 0, fe = 0; int ev = 0;
  rc = lbfgs_minimize([&](const double* xx, double* f, double* g) { return objcov_device_eval(e, dir, T, norm8, xx, f, g); }, x, max_evals, &fs, &fe, &ev);
  if (rc != LRR_OK) return rc;
  double Un[144] = {0}; ind = 0;
  for (int d1 = 0; d1 < 12; ++d1) for (int d2 = d1; d2 < 12; ++d2) Un[d1 * 12 + d2] = x[ind++];
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) { double s = 0; for (int k = 0; k < 12; ++k) s += Un[k * 12 + i] * Un[k * 12 + j]; T.cov[i * 12 + j] = s; }
  if (e->normalize) affine_unnormalize(T, norm8[3], norm8[7], norm8, norm8 + 4);
  memcpy(h.A, T.A, 72); memcpy(h.t, T.t, 24); memcpy(h.c, T.c, 24); memcpy(h.cov, T.cov, sizeof(T.cov));
  if ((rc = push_scal(e, dir)) != LRR_OK) return rc;
  e->werr_cache_valid = false;
  if (f_start) *f_start = fs; if (f_end) *f_end = fe; if (n_evals) *n_evals = ev;
  return LRR_OK;
}

// ---- parameter objective of the CDC M-step (cdcl_obj_fun_par<3,12>, cdcl/cdcl_obj_fun_par.txx:46-144) --------------------------------
void lrr_k_objpar_launch(const DirArgs& a, const double* A, const double* t, const double* c, const double* norm8, const double* Sigma_dev,
                         double* partials, unsigned int* ticket, cudaStream_t st, long long* launches);
// one evaluation at theta = (A row-major, t) in the frame of T (its centre, its covariance = the frozen Sigma_theta) / norm8
static int objpar_device_eval(lrr_est* e, int dir, const HostTrans& T, const double* norm8, const double* theta, double* f, double* g) {
  lrr_ctx* ctx = e->ctx;
  lrr_dir& d = e->d[dir];
  double* pin = (double*)ctx->pinned;
  memcpy(pin, T.cov, sizeof(T.cov));
  CK(cudaMemcpyAsync(ctx->scratch_d, pin, sizeof(T.cov), cudaMemcpyHostToDevice, ctx->stream));
  EstArgs a = est_args(e);
  a.d[dir].M = (int)d.matches;
  if (d.cdc_active) cdc_dir_args(d, a.d[dir]);
  { PHASE(PH_OBJCOV); lrr_k_objpar_launch(a.d[dir], theta, theta + 9, T.c, norm8, ctx->scratch_d, ctx->partials, ctx->ticket, ctx->stream, &ctx->launches); }
  CK(cudaMemcpyAsync(pin + 256, d.sc->mom, 13 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  const double* mom = pin + 256;
  const double n = d.cdc_active ? (double)d.cdc_rows : (double)d.matches;   // matches_.size() (cdcl_obj_fun_par.txx:131-137)
  if (f) *f = mom[12] / n;
  if (g) for (int k = 0; k < 12; ++k) g[k] = mom[k] * (1.0 / n);
  return LRR_OK;
}
extern "C" int lrr_objpar_eval(lrr_est* e, int dir, const double* theta, double* f, double* g) {
  touch(e);
  int rc = objcov_check(e, dir); if (rc != LRR_OK) return rc;
  if (!theta) return LRR_ERR_INVALID;
  enter(e->ctx);
  if ((rc = pull_scal(e, 1 << dir)) != LRR_OK) return rc;
  HostTrans T; memcpy(T.A, e->h[dir].A, 72); memcpy(T.t, e->h[dir].t, 24); memcpy(T.c, e->h[dir].c, 24); memcpy(T.cov, e->h[dir].cov, sizeof(T.cov));
  const double norm8[8] = {0, 0, 0, 1, 0, 0, 0, 1};
  return objpar_device_eval(e, dir, T, norm8, theta, f, g);
}
// estimate_parameters (cdcl/cdcl_estimation.txx:207-262) on the ICP matches of direction `dir`: normalise, L-BFGS over theta
// (<= max_evals evaluations, the 5 % early stop of cdcl_lbfgs), set_parameterization, unnormalise
extern "C" int lrr_objpar_refine(lrr_est* e, int dir, int max_evals, double* f_start, double* f_end, int* n_evals, int* converged) {
  touch(e);
  int rc = objcov_check(e, dir); if (rc != LRR_OK) return rc;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  double norm8[8] = {0, 0, 0, 1, 0, 0, 0, 1};
  bool cdc_norm = false;
  if (e->normalize && d.cdc_active) { if ((rc = cdc_normalisation(e, dir, norm8)) != LRR_OK) return rc; cdc_norm = true; }
  else if (e->normalize) {   // cdcl_normalize_matches over the current matches (:214-217)
    if ((rc = bump_epoch(e, dir)) != LRR_OK) return rc;
    EstArgs a = est_args(e, dir);
    a.d[dir].M = (int)d.matches;
    a.ndir = 1; if (dir == 1) { a.d[0] = a.d[1]; a.src = 0; }
    lrr_k_restamp(a, ctx->stream, &ctx->launches);
    lrr_k_norm1(a, ctx->stream, &ctx->launches); lrr_k_norm2(a, ctx->stream, &ctx->launches);
  }
  if ((rc = pull_scal(e, 1 << dir)) != LRR_OK) return rc;
  Scal& h = e->h[dir];
  HostTrans T; memcpy(T.A, h.A, 72); memcpy(T.t, h.t, 24); memcpy(T.c, h.c, 24); memcpy(T.cov, h.cov, sizeof(T.cov));
  if (e->normalize) {
    if (!cdc_norm) {
      for (int k = 0; k < 3; ++k) { norm8[k] = h.cfrom[k]; norm8[4 + k] = h.cto[k]; }
      norm8[3] = h.rfrom; norm8[7] = h.rto;
    } else {   // (reported by lrr_est_get_state like the normalisation of any other estimate)
      for (int k = 0; k < 3; ++k) { h.cfrom[k] = norm8[k]; h.cto[k] = norm8[4 + k]; }
      h.rfrom = norm8[3]; h.rto = norm8[7];
    }
    affine_normalize(T, norm8[3], norm8[7], norm8, norm8 + 4);
  }
  std::vector<double> x(12);
  for (int i = 0; i < 9; ++i) x[i] = T.A[i];
  for (int i = 0; i < 3; ++i) x[9 + i] = T.t[i];
  double fs = 0, fe = 0; int ev = 0;
  rc = lbfgs_minimize([&](const double* xx, double* f, double* g) { return objpar_device_eval(e, dir, T, norm8, xx, f, g); }, x, max_evals, &fs, &fe, &ev);
  if (rc != LRR_OK) return rc;
  for (int i = 0; i < 9; ++i) T.A[i] = x[i];
  for (int i = 0; i < 3; ++i) T.t[i] = x[9 + i];
  if (e->normalize) affine_unnormalize(T, norm8[3], norm8[7], norm8, norm8 + 4);
  memcpy(h.A, T.A, 72); memcpy(h.t, T.t, 24); memcpy(h.c, T.c, 24); memcpy(h.cov, T.cov, sizeof(T.cov));
  if ((rc = push_scal(e, dir)) != LRR_OK) return rc;
  e->werr_cache_valid = false;
  if (f_start) *f_start = fs; if (f_end) *f_end = fe; if (n_evals) *n_evals = ev; if (converged) *converged = (fs - fe) < 1e-4;
  return LRR_OK;
}

// ---- CDC E-step: covariance-driven radius matching (cdcl_estimation<3,12>::compute_weights, cdcl/cdcl_estimation.txx:342-545) ---------
static int cdc_check(lrr_est* e, int dir) {
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx;
  if (e->dof != 12 || e->dim != 3) { ctx->err = "cdcl_estimation::compute_weights: only the affine <3,12> instantiation is built"; return LRR_ERR_UNSUPPORTED; }
  if (!e->d[dir].from_all->C6 || !e->d[dir].to_all->C6) { ctx->err = "compute_weights needs feature covariances (C6)"; return LRR_ERR_STATE; }
  if (e->d[dir].n_in <= 0 || e->d[dir].to_all->n <= 0) { ctx->err = "compute_weights: empty feature set"; return LRR_ERR_STATE; }
  return LRR_OK;
}
extern "C" int lrr_cdc_compute_weights(lrr_est* e, int dir, int max_matches, double eval_multiplier, lrr_cdc_summary* out) {
  touch(e);
  int rc = cdc_check(e, dir); if (rc != LRR_OK) return rc;
  if (max_matches <= 0 || !(eval_multiplier > 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  cudaStream_t st = ctx->stream;
  if ((rc = pull_scal(e, 1 << dir)) != LRR_OK) return rc;
  const Scal& h = e->h[dir];
  double* pin = (double*)ctx->pinned;
  memcpy(pin, h.cov, 144 * sizeof(double));
  CK(cudaMemcpyAsync(ctx->scratch_d, pin, 144 * sizeof(double), cudaMemcpyHostToDevice, st));
  const int M = (int)d.n_in; const size_t nto = (size_t)d.to_all->n;
  if (!d.cdc_rowcnt) {
    CK(dmalloc(&d.cdc_rowcnt, 2 * d.cap + 4)); d.cdc_rowstart = d.cdc_rowcnt + d.cap; d.cdc_istat = d.cdc_rowcnt + 2 * d.cap;
    CK(dmalloc(&d.cdc_roww, d.cap)); CK(dmalloc(&d.cdc_tow, nto + 1)); d.cdc_dsum = d.cdc_tow + nto;
    CK(dmalloc(&d.cdc_cursor, 1));
  }
  const int blocks = lrr_cdc_blocks(M);
  const size_t scratch = lrr_cdc_scratch_entries(blocks);
  if (scratch > d.cdc_scratch) {
    dfree(d.cdc_cand); dfree(d.cdc_candw); d.cdc_cand = nullptr; d.cdc_candw = nullptr; d.cdc_scratch = 0;
    CK(dmalloc(&d.cdc_cand, scratch)); CK(dmalloc(&d.cdc_candw, scratch)); d.cdc_scratch = scratch;
  }
  int64_t want = (int64_t)M * std::min(max_matches, 48);
  d.cdc_active = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (want > d.cdc_ecap) {
      dfree(d.cdc_efrom); dfree(d.cdc_eslot); dfree(d.cdc_ew); d.cdc_efrom = d.cdc_eslot = nullptr; d.cdc_ew = nullptr; d.cdc_ecap = 0;
      if (want > 0x7fffffffLL) { ctx->err = "compute_weights: more than 2^31 match entries"; return LRR_ERR_UNSUPPORTED; }
      CK(dmalloc(&d.cdc_efrom, (size_t)want)); CK(dmalloc(&d.cdc_eslot, (size_t)want)); CK(dmalloc(&d.cdc_ew, (size_t)want)); d.cdc_ecap = want;
    }
    CK(cudaMemsetAsync(d.cdc_cursor, 0, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(d.cdc_tow, 0, (nto + 1) * sizeof(double), st));
    { int* ip = (int*)(pin + 256); ip[0] = 0; ip[1] = 0x7fffffff; ip[2] = 0; ip[3] = 0;
      CK(cudaMemcpyAsync(d.cdc_istat, ip, 4 * sizeof(int), cudaMemcpyHostToDevice, st)); }
    CdcLaunch L;
    L.A = h.A; L.t = h.t; L.c = h.c; L.Sigma_dev = ctx->scratch_d; L.eval_mult = eval_multiplier; L.max_matches = max_matches; L.M = M;
    L.fx = d.fx; L.fy = d.fy; L.fz = d.fz; L.forig = d.subset ? d.forig : nullptr; L.to = d.to_all->dev; L.from_all = d.from_all->dev;
    L.efrom = d.cdc_efrom; L.eslot = d.cdc_eslot; L.ew = d.cdc_ew; L.ecap = d.cdc_ecap; L.cursor = d.cdc_cursor;
    L.rowcnt = d.cdc_rowcnt; L.rowstart = d.cdc_rowstart; L.roww = d.cdc_roww; L.tow = d.cdc_tow; L.dsum = d.cdc_dsum; L.istat = d.cdc_istat;
    L.cand = d.cdc_cand; L.candw = d.cdc_candw; L.blocks = blocks;
    { PHASE(PH_MATCH); lrr_k_cdc_match(L, st, &ctx->launches); }
    CK(cudaMemcpyAsync(pin + 260, d.cdc_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pin + 261, d.cdc_dsum, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pin + 262, d.cdc_istat, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const int64_t entries = (int64_t) * (unsigned long long*)(pin + 260);
    const int* is = (const int*)(pin + 262);
    if (is[3]) { ctx->err = "compute_weights: max_matches does not fit the per-warp candidate list (4064)"; return LRR_ERR_UNSUPPORTED; }
    if (entries > d.cdc_ecap) { want = entries; continue; }   // the flat set did not fit: once more with the size the cursor reports
    d.cdc_entries = entries; d.cdc_rows = is[0];
    lrr_k_cdc_normalise((int)entries, d.cdc_efrom, d.cdc_eslot, d.cdc_ew, d.cdc_roww, d.cdc_tow, st, &ctx->launches);
    d.cdc_active = entries > 0;   // no match at all: the caller falls back to find_closest_euclidean (:544)
    if (out) {
      out->rows = d.cdc_rows; out->entries = entries; out->min_matches = is[0] ? is[1] : 0; out->max_matches = is[2];
      out->avg_largest_eval = pin[261] / (double)d.cdc_rows;   // divided by matches_.size(), summed over every from-feature (:375, :480)
    }
    return LRR_OK;
  }
  ctx->err = "compute_weights: the match set grew between two passes"; return LRR_ERR_STATE;
}
extern "C" int lrr_cdc_get_matches(lrr_est* e, int dir, int64_t cap_rows, int64_t cap_entries, int64_t* n_rows, int64_t* n_entries, int32_t* row_from,
                                   int64_t* row_start, int32_t* to_idx, double* w) {
  if (!e || dir < 0 || dir > (e->sym() ? 1 : 0)) return LRR_ERR_INVALID;
  lrr_ctx* ctx = e->ctx; enter(ctx);
  lrr_dir& d = e->d[dir];
  if (!d.cdc_active) { ctx->err = "lrr_cdc_get_matches: no CDC match set"; return LRR_ERR_STATE; }
  if (n_rows) *n_rows = d.cdc_rows; if (n_entries) *n_entries = d.cdc_entries;
  if (!row_from && !row_start && !to_idx && !w) return LRR_OK;
  if (cap_rows < d.cdc_rows || cap_entries < d.cdc_entries) return LRR_ERR_INVALID;
  const size_t M = (size_t)d.n_in, E = (size_t)d.cdc_entries;
  std::vector<int> rowcnt(M), rowstart(M), eslot(E), orig(d.to_all->n), forig;
  std::vector<double> ew(E); std::vector<double4> pos(d.to_all->n);
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(rowcnt.data(), d.cdc_rowcnt, M * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(rowstart.data(), d.cdc_rowstart, M * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(eslot.data(), d.cdc_eslot, E * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ew.data(), d.cdc_ew, E * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pos.data(), d.to_all->pos4, pos.size() * sizeof(double4), cudaMemcpyDeviceToHost, st));
  if (d.subset) { forig.resize(M); CK(cudaMemcpyAsync(forig.data(), d.forig, M * sizeof(int), cudaMemcpyDeviceToHost, st)); }
  CK(cudaStreamSynchronize(st));
  int64_t r = 0, pos_out = 0;
  std::vector<std::pair<int, double>> row;
  for (size_t i = 0; i < M; ++i) {
    if (rowcnt[i] <= 0) continue;
    row.clear();
    for (int k = 0; k < rowcnt[i]; ++k) { const int sl = eslot[rowstart[i] + k]; long long ob; memcpy(&ob, &pos[sl].w, 8); row.push_back({(int)ob, ew[rowstart[i] + k]}); }
    std::sort(row.begin(), row.end());
    if (row_from) row_from[r] = d.subset ? forig[i] : (int32_t)i;
    if (row_start) row_start[r] = pos_out;
    for (auto& pr : row) { if (to_idx) to_idx[pos_out] = pr.first; if (w) w[pos_out] = pr.second; ++pos_out; }
    ++r;
  }
  if (row_start) row_start[r] = pos_out;
  return LRR_OK;
}
// the resolution switch at the end of the matching pass (cdcl/cdcl_estimation.txx:494-508); bitwise `&` of the reference kept as logic
extern "C" int lrr_cdc_resolution_switch(double avg_largest_eval, const double* fixed_spacing, int res, int max_res, int allowed, int* new_res) {
  if (!fixed_spacing || !new_res || res < 0 || res > max_res) return LRR_ERR_INVALID;
  int r = res; bool not_restarted = true;
  if (avg_largest_eval < 2.0 * fixed_spacing[r] && allowed && not_restarted) { if (r > 0) { --r; not_restarted = false; } }
  if (r + 1 <= max_res && avg_largest_eval > 2.0 * fixed_spacing[r + 1] && allowed && not_restarted) { ++r; }
  *new_res = r;
  return LRR_OK;
}
// one EM iteration of cdcl_estimation<3,12>::one_iteration (cdcl/cdcl_estimation.txx:88-186) at a single resolution: weights; past
// iteration startEM = 5 parameters and weights again; covariance.  converged = parameters && covariance && |Sigma|_F < 1e-3 (:185).
extern "C" int lrr_cdc_one_iteration(lrr_est* e, int dir, unsigned iteration, int max_matches, double eval_multiplier, int* converged, lrr_cdc_summary* out) {
  int rc = cdc_check(e, dir); if (rc != LRR_OK) return rc;
  const unsigned startEM = 5;
  lrr_cdc_summary s; memset(&s, 0, sizeof(s));
  if ((rc = lrr_cdc_compute_weights(e, dir, max_matches, eval_multiplier, &s)) != LRR_OK) return rc;
  if (s.entries == 0) { e->ctx->err = "cdc one_iteration: the covariance is too small for any match (fall back to find_closest_euclidean, :544)"; return LRR_ERR_STATE; }
  int par_conv = 0, cov_conv = 0;
  double fs = 0, fe = 0; int ev = 0;
  if (iteration > startEM) {
    if ((rc = lrr_objpar_refine(e, dir, 5, &fs, &fe, &ev, &par_conv)) != LRR_OK) return rc;
    if ((rc = lrr_cdc_compute_weights(e, dir, max_matches, eval_multiplier, &s)) != LRR_OK) return rc;
    if (s.entries == 0) { e->ctx->err = "cdc one_iteration: no match after the parameter step"; return LRR_ERR_STATE; }
  }
  if ((rc = lrr_objcov_refine(e, dir, 5, &fs, &fe, &ev)) != LRR_OK) return rc;
  cov_conv = (fs - fe) < 1e-5;   // (:300)
  double fro = 0; for (int i = 0; i < 144; ++i) fro += e->h[dir].cov[i] * e->h[dir].cov[i];
  if (converged) *converged = par_conv && cov_conv && std::sqrt(fro) < 1e-3;
  if (out) *out = s;
  return LRR_OK;
}

// validTransform (rrl/itkLocationRegistration.cxx:2914-2926): singular values of A inside [0.1, 10]
static bool valid_transform(const double* A) {
  double U[9], V[9], w[3];
  lrr_jacobi_svd<3>(3, A, U, w, V);
  for (int d = 0; d < 3; ++d) if (w[d] > 10.0 || w[d] < 0.1) return false;
  return true;
}
// FeaturesInROIs (rrl/itkLocationRegistration.cxx:1324-1360): indices of the features inside the cube, in order
static void roi_indices(lrr_fset* fs, const double* c, double half, std::vector<int32_t>& out) {
  out.clear();
  if (fs->host_xyz.empty() && fs->n > 0) {   // first host-side ROI selection on this set: fetch the locations once
    fs->host_xyz.resize((size_t)3 * fs->n);
    cudaMemcpyAsync(fs->host_xyz.data(), fs->xyz, sizeof(double) * 3 * fs->n, cudaMemcpyDeviceToHost, fs->ctx->stream);
    cudaStreamSynchronize(fs->ctx->stream);
  }
  const int64_t n = fs->n; const double* x = fs->host_xyz.data(); const double* y = x + n; const double* z = y + n;
  const double x0 = c[0] - half, x1 = c[0] + half, y0 = c[1] - half, y1 = c[1] + half, z0 = c[2] - half, z1 = c[2] + half;
  for (int64_t i = 0; i < n; ++i)
    if (!(x0 > x[i] || x[i] > x1 || y0 > y[i] || y[i] > y1 || z0 > z[i] || z[i] > z1)) out.push_back((int32_t)i);
}

// one job through the per-iteration kernels of the single-registration path (estimator reused, like the driver :1635-1638)
static int batch_job_host(lrr_est* est, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job& job, int64_t j, int max_iter, uint64_t shuffle_seed,
                          lrr_result& r, std::vector<int32_t>& mi, std::vector<int32_t>& fi) {
  int rc;
  memset(&r, 0, sizeof(r));
  roi_indices(fixed_all, job.roi_center_fixed, job.roi_half_size, fi);
  roi_indices(moving_all, job.roi_center_moving, job.roi_half_size, mi);
  r.n_moving_inside = (int32_t)mi.size(); r.n_fixed_inside = (int32_t)fi.size();
  if (mi.size() < 1000 || fi.size() < 1000) { r.status = 1; return LRR_OK; }   // :1632
  if ((rc = lrr_est_set_subsets(est, mi.data(), (int64_t)mi.size(), fi.data(), (int64_t)fi.size())) != LRR_OK) return rc;
  if ((rc = lrr_est_set_shuffle(est, LRR_ORDER_RESHUFFLE, shuffle_seed + (uint64_t)j)) != LRR_OK) return rc;
  if ((rc = lrr_est_set_transform(est, LRR_TRANS_AFFINE, job.A, job.t, job.c, nullptr)) != LRR_OK) return rc;
  int converged = 0, it = 0;
  while (!converged && it < max_iter) {                                   // :2418
    if ((rc = lrr_est_one_iteration(est, (unsigned)it, &converged)) != LRR_OK) return rc;
    ++it;
    if (!valid_transform(est->h[0].A)) { r.status = 2; break; }           // :2482
  }
  const Scal& T = est->h[0];
  r.converged = converged; r.oscillated = est->oscillation_count > 10; r.iterations = it;
  memcpy(r.A, T.A, 72); memcpy(r.t, T.t, 24); memcpy(r.c, T.c, 24); memcpy(r.cov, T.cov, sizeof(double) * 144);
  r.rms = rms_of(est, T, true); r.weighted_error = werr_of(T);
  return LRR_OK;
}

cudaError_t lrr_batch_launch(const FsetDev& moving_all, const FsetDev& fixed_all, const lrr_job* jobs_dev, lrr_result* out_dev, long long njobs,
                             const long long* job_ids_dev, int max_iter, unsigned long long shuffle_seed, const double* muse_table, int muse_rows, int nblocks, void* workspace,
                             size_t workspace_bytes, cudaStream_t st, long long* launches, int nslots = 0, int quantum = 0);
size_t lrr_batch_workspace_bytes(int nblocks, int n_moving, int n_fixed);
int lrr_batch_blocks_per_sm();

// One launch of the block-per-registration kernel over a run of jobs, with its own stream, workspace and staging buffers: the
// single-device entry points use one slot on the context's stream, lrr_batch_run_multi keeps two in flight per device.
struct BatchSlot {
  cudaStream_t st = nullptr; bool own_stream = false;
  void* work = nullptr; lrr_job* djobs = nullptr; lrr_result* dout = nullptr; long long* dids = nullptr;
  int64_t a = 0, n = 0; bool busy = false;
  lrr_ctx* ctx = nullptr; size_t work_bytes = 0;   // the workspace goes back to the context's cache
};
// workspace cache of the context (released workspaces are idle: slot_collect has synchronised their stream)
static cudaError_t ws_acquire(lrr_ctx* ctx, size_t bytes, cudaStream_t st, void** out, size_t* got) {
  int best = -1;
  for (size_t i = 0; i < ctx->batch_ws.size(); ++i) if (ctx->batch_ws[i].bytes >= bytes && (best < 0 || ctx->batch_ws[i].bytes < ctx->batch_ws[(size_t)best].bytes)) best = (int)i;
  if (best >= 0) { *out = ctx->batch_ws[(size_t)best].p; *got = ctx->batch_ws[(size_t)best].bytes; ctx->batch_ws.erase(ctx->batch_ws.begin() + best); return cudaSuccess; }
  *got = bytes;
  return cudaMallocAsync(out, bytes, st);
}
static void ws_release(lrr_ctx* ctx, void* p, size_t bytes, cudaStream_t st) {
  if (!p) return;
  if (ctx->batch_ws.size() < 2) { ctx->batch_ws.push_back({p, bytes}); return; }
  size_t small = 0; for (size_t i = 1; i < ctx->batch_ws.size(); ++i) if (ctx->batch_ws[i].bytes < ctx->batch_ws[small].bytes) small = i;
  if (ctx->batch_ws[small].bytes < bytes) { cudaFreeAsync(ctx->batch_ws[small].p, st); ctx->batch_ws[small] = {p, bytes}; }
  else cudaFreeAsync(p, st);
}
static void slot_release(BatchSlot& s) {
  if (s.work) { if (s.ctx) ws_release(s.ctx, s.work, s.work_bytes, s.st); else cudaFreeAsync(s.work, s.st); } if (s.djobs) cudaFreeAsync(s.djobs, s.st); if (s.dout) cudaFreeAsync(s.dout, s.st); if (s.dids) cudaFreeAsync(s.dids, s.st);
  s.work = nullptr; s.djobs = nullptr; s.dout = nullptr; s.dids = nullptr;
}
static int slot_launch_impl(lrr_ctx* ctx, BatchSlot& s, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t a, int64_t n,
                            const int64_t* ids, int64_t id_base, int max_iter, uint64_t shuffle_seed, lrr_result* out);
static int slot_launch(lrr_ctx* ctx, BatchSlot& s, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t a, int64_t n,
                       const int64_t* ids, int64_t id_base, int max_iter, uint64_t shuffle_seed, lrr_result* out) {
  const int rc = slot_launch_impl(ctx, s, moving_all, fixed_all, jobs, a, n, ids, id_base, max_iter, shuffle_seed, out);
  if (rc != LRR_OK) { cudaStreamSynchronize(s.st); slot_release(s); s.busy = false; }   // every early return of the body ends here: nothing leaks
  return rc;
}
static int slot_launch_impl(lrr_ctx* ctx, BatchSlot& s, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t a, int64_t n,
                            const int64_t* ids, int64_t id_base, int max_iter, uint64_t shuffle_seed, lrr_result* out) {
  const int nblocks = (int)std::min<int64_t>(n, (int64_t)lrr_num_sms() * lrr_batch_blocks_per_sm());
  // time slicing (lrr_batch.cu): more workspace slots than blocks, a job is parked after `batch_quantum` iterations and the block takes
  // the next unit — every job is started early, so the jobs that run into the iteration cap no longer start late and end the launch
  // alone.  Slots: up to 1024 (one per job in flight) within an eighth of the free device memory.
  int nslots = nblocks;
  const int quantum = ctx->opt.batch_quantum;
  if (quantum > 0 && n > nblocks) {
    const size_t per_slot = lrr_batch_workspace_bytes(2, (int)moving_all->n, (int)fixed_all->n) - lrr_batch_workspace_bytes(1, (int)moving_all->n, (int)fixed_all->n);
    // (the budget is fixed per context, from the device's total memory: cudaMemGetInfo per launch was measured at 0.1 ... 50 ms)
    if (!ctx->batch_budget) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
      ctx->batch_budget = std::max<size_t>(std::min<size_t>(total_b / 16, (size_t)16 << 30), 1);
    }
    const size_t budget = ctx->batch_budget;
    const int64_t fit = per_slot ? (int64_t)(budget / per_slot) : 0;
    nslots = (int)std::max<int64_t>(nblocks, std::min<int64_t>(std::min<int64_t>(n, ctx->opt.batch_slots), fit));
  }
  const size_t wbytes = lrr_batch_workspace_bytes(nslots, (int)moving_all->n, (int)fixed_all->n);
  cudaStream_t st = s.st;
  s.ctx = ctx;
  const bool dbg_time = getenv("LRR_BATCH_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tl0 = now();
  CK(ws_acquire(ctx, wbytes, st, &s.work, &s.work_bytes));
  const double tl1 = now();
  CK(cudaMallocAsync((void**)&s.djobs, sizeof(lrr_job) * n, st)); CK(cudaMallocAsync((void**)&s.dout, sizeof(lrr_result) * n, st));
  CK(cudaMallocAsync((void**)&s.dids, sizeof(long long) * n, st));
  CK(cudaMemcpyAsync(s.djobs, jobs + a, sizeof(lrr_job) * n, cudaMemcpyHostToDevice, st));
  std::vector<long long> hid((size_t)n);
  for (int64_t i = 0; i < n; ++i) hid[(size_t)i] = ids ? (long long)ids[a + i] : (long long)(id_base + a + i);
  CK(cudaMemcpyAsync(s.dids, hid.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));   // `hid` is a local buffer (and pageable copies are staged anyway)
  double table[4 * 32] = {0};
  const int rows = build_muse_table(1000u, table);
  fset_wait_dirs(moving_all, st); fset_wait_dirs(fixed_all, st);
  if (dbg_time) { cudaStreamSynchronize(st); fprintf(stderr, "[batch host] workspace %.2f ms, job copies %.2f ms\n", tl1 - tl0, now() - tl1); }
  CK(lrr_batch_launch(moving_all->dev, fixed_all->dev, s.djobs, s.dout, n, s.dids, max_iter, shuffle_seed, table, rows, nblocks, s.work, s.work_bytes, st, &ctx->launches, nslots, quantum));
  CK(cudaMemcpyAsync(out + a, s.dout, sizeof(lrr_result) * n, cudaMemcpyDeviceToHost, st));
  s.a = a; s.n = n; s.busy = true;
  return LRR_OK;
}
static int slot_collect(lrr_ctx* ctx, BatchSlot& s) {
  if (!s.busy) return LRR_OK;
  s.busy = false;
  cudaError_t e1 = cudaStreamSynchronize(s.st), e2 = cudaGetLastError();
  slot_release(s);
  CK(e1); CK(e2);
  return LRR_OK;
}

// jobs the block kernel handed back (ROI above its capacity, selector overflow; status < 0) and the LRR_BATCH_MODE=host debugging
// mode go through one reused estimator object and the per-iteration kernels
static int batch_host_jobs(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, const std::vector<int64_t>& todo, const int64_t* ids,
                           int64_t id_base, int max_iter, uint64_t shuffle_seed, lrr_result* out) {
  if (todo.empty()) return LRR_OK;
  lrr_est* est = nullptr;
  int rc = lrr_est_create(ctx, LRR_RRL_SYM, 3, 12, moving_all, fixed_all, &est);
  if (rc != LRR_OK) return rc;
  std::vector<int32_t> mi, fi;
  for (int64_t j : todo) if ((rc = batch_job_host(est, moving_all, fixed_all, jobs[j], ids ? ids[j] : id_base + j, max_iter, shuffle_seed, out[j], mi, fi)) != LRR_OK) break;
  lrr_est_destroy(est);
  return rc;
}

void lrr_batch_roi_cost(const FsetDev& moving_all, const FsetDev& fixed_all, const lrr_job* jobs_dev, long long njobs, int* cost_dev, cudaStream_t st, long long* launches);
// Expected cost of every job (features covered by its two ROI cubes; 0 when an ROI cannot pass the 1000-feature gate): what the
// batched entry points order their queue by, and what a host that shards the list over processes can balance the shares with.
extern "C" int lrr_batch_expected_cost(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t njobs, int32_t* cost) {
  if (!ctx || !moving_all || !fixed_all || njobs < 0 || (njobs > 0 && (!jobs || !cost))) return LRR_ERR_INVALID;
  if (moving_all->ctx != ctx || fixed_all->ctx != ctx) return LRR_ERR_INVALID;
  if (njobs == 0) return LRR_OK;
  enter(ctx);
  cudaStream_t st = ctx->stream;
  lrr_job* dj = nullptr; int* dc = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&dj, sizeof(lrr_job) * njobs, st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&dc, sizeof(int) * njobs, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dj, jobs, sizeof(lrr_job) * njobs, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) { lrr_batch_roi_cost(moving_all->dev, fixed_all->dev, dj, njobs, dc, st, &ctx->launches); e = cudaMemcpyAsync(cost, dc, sizeof(int) * njobs, cudaMemcpyDeviceToHost, st); }
  const cudaError_t e2 = cudaStreamSynchronize(st);
  if (dj) cudaFreeAsync(dj, st); if (dc) cudaFreeAsync(dc, st);
  CK(e); CK(e2); CK(cudaGetLastError());
  return LRR_OK;
}
// queue order of a batched run: longest expected first (stable, so equal costs keep the list order)
static int batch_queue_order(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t njobs, std::vector<int64_t>& perm) {
  perm.resize((size_t)njobs);
  for (int64_t j = 0; j < njobs; ++j) perm[(size_t)j] = j;
  if (!ctx->opt.batch_sort || njobs < 2) return LRR_OK;
  std::vector<int32_t> cost((size_t)njobs);
  int rc = lrr_batch_expected_cost(ctx, moving_all, fixed_all, jobs, njobs, cost.data()); if (rc != LRR_OK) return rc;
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
  return LRR_OK;
}

// Batched location registration: the outer loops of LocationRegistration::Run (rrl/itkLocationRegistration.cxx:2138,
// 2230,2332) over one resident volume pair.  Device path: one persistent thread block per registration (lrr_batch.cu).
static int batch_run_impl(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t njobs, const int64_t* ids, int max_iter,
                          uint64_t shuffle_seed, lrr_result* out) {
  if (!ctx || !moving_all || !fixed_all || njobs < 0 || (njobs > 0 && (!jobs || !out))) return LRR_ERR_INVALID;
  if (moving_all->ctx != ctx || fixed_all->ctx != ctx) return LRR_ERR_INVALID;
  enter(ctx);
  if (njobs == 0) return LRR_OK;
  const char* mode = getenv("LRR_BATCH_MODE");
  const bool host_only = (mode && std::string(mode) == "host") || max_iter > 250;
  std::vector<int64_t> todo;
  if (!host_only) {
    // the queue in the order longest-expected-first: permuted copies of the jobs and of their ids (the id seeds a job's shuffle and
    // must stay the job's own), results scattered back to the list order afterwards
    std::vector<int64_t> perm;
    const bool dbg_time = getenv("LRR_BATCH_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double th0 = now();
    int rc = batch_queue_order(ctx, moving_all, fixed_all, jobs, njobs, perm); if (rc != LRR_OK) return rc;
    const double th1 = now();
    std::vector<lrr_job> pj((size_t)njobs); std::vector<int64_t> pid((size_t)njobs); std::vector<lrr_result> pout((size_t)njobs);
    for (int64_t i = 0; i < njobs; ++i) { const int64_t j = perm[(size_t)i]; pj[(size_t)i] = jobs[j]; pid[(size_t)i] = ids ? ids[j] : j; }
    BatchSlot s; s.st = ctx->stream;
    { PHASE(PH_BATCH); rc = slot_launch(ctx, s, moving_all, fixed_all, pj.data(), 0, njobs, pid.data(), 0, max_iter, shuffle_seed, pout.data()); }
    const double th2 = now();
    const int rc2 = slot_collect(ctx, s);
    if (dbg_time) fprintf(stderr, "[batch host] order %.2f ms, launch (workspace, copies, clears, kernel enqueue) %.2f ms, collect %.2f ms\n", th1 - th0, th2 - th1, now() - th2);
    if (rc != LRR_OK) return rc;
    if (rc2 != LRR_OK) return rc2;
    for (int64_t i = 0; i < njobs; ++i) out[perm[(size_t)i]] = pout[(size_t)i];
    for (int64_t j = 0; j < njobs; ++j) if (out[j].status < 0) todo.push_back(j);
  } else {
    for (int64_t j = 0; j < njobs; ++j) todo.push_back(j);
  }
  return batch_host_jobs(ctx, moving_all, fixed_all, jobs, todo, ids, 0, max_iter, shuffle_seed, out);
}
extern "C" int lrr_batch_run(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t njobs, int max_iter,
                             uint64_t shuffle_seed, lrr_result* out) {
  return batch_run_impl(ctx, moving_all, fixed_all, jobs, njobs, nullptr, max_iter, shuffle_seed, out);
}
extern "C" int lrr_batch_run_ids(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, const int64_t* job_ids, int64_t njobs,
                                 int max_iter, uint64_t shuffle_seed, lrr_result* out) {
  if (njobs > 0 && !job_ids) return LRR_ERR_INVALID;
  return batch_run_impl(ctx, moving_all, fixed_all, jobs, njobs, job_ids, max_iter, shuffle_seed, out);
}

// The whole job list over ndev devices of THIS process (the reference's outer loops are one sequential program: a host that links
// the backend has no process per GPU).  One host thread per device, one launch per device over its share of the list (dealt from the
// jobs sorted by expected cost); the kernel schedules inside a launch (job cursor, time slices).  Results are written in place; job j is
// seeded with shuffle_seed + j whichever device runs it, so the result does not depend on the number of devices.
#include <thread>
extern "C" int lrr_batch_run_multi(lrr_ctx** ctxs, int ndev, lrr_fset** moving_all, lrr_fset** fixed_all, const lrr_job* jobs, int64_t njobs,
                                   int max_iter, uint64_t shuffle_seed, lrr_result* out) {
  if (!ctxs || ndev <= 0 || !moving_all || !fixed_all || njobs < 0 || (njobs > 0 && (!jobs || !out))) return LRR_ERR_INVALID;
  for (int d = 0; d < ndev; ++d) {
    if (!ctxs[d] || !moving_all[d] || !fixed_all[d] || moving_all[d]->ctx != ctxs[d] || fixed_all[d]->ctx != ctxs[d]) return LRR_ERR_INVALID;
    if (moving_all[d]->n != moving_all[0]->n || fixed_all[d]->n != fixed_all[0]->n) return LRR_ERR_INVALID;   // replicas of one pair
  }
  if (njobs == 0) return LRR_OK;
  if (max_iter > 250 || ndev == 1) return batch_run_impl(ctxs[0], moving_all[0], fixed_all[0], jobs, njobs, nullptr, max_iter, shuffle_seed, out);
  std::vector<int64_t> perm;
  { enter(ctxs[0]); const int rc = batch_queue_order(ctxs[0], moving_all[0], fixed_all[0], jobs, njobs, perm); if (rc != LRR_OK) return rc; }
  // One launch per device over a share dealt from the list sorted by expected cost (device d takes positions d, d + ndev, ... : every
  // device gets the same mix of large and small ROIs); inside a launch the time-sliced kernel does the dynamic scheduling.  (A first
  // version drew runs of a few hundred jobs from a shared cursor, two launches in flight per device: every launch ends in a sparse tail
  // of long jobs, and sixteen of them cost more than they balanced — 0.64 of 2 x the one-device rate on two B200, 0.45 with time slicing.)
  std::vector<std::vector<lrr_job> > pj((size_t)ndev); std::vector<std::vector<int64_t> > pid((size_t)ndev); std::vector<std::vector<lrr_result> > pout((size_t)ndev);
  for (int64_t i = 0; i < njobs; ++i) { const size_t d = (size_t)(i % ndev); pj[d].push_back(jobs[perm[(size_t)i]]); pid[d].push_back(perm[(size_t)i]); }
  for (int d = 0; d < ndev; ++d) pout[(size_t)d].resize(pj[(size_t)d].size());
  std::vector<int> rcs((size_t)ndev, LRR_OK);
  auto worker = [&](int d) {
    lrr_ctx* ctx = ctxs[d];
    enter(ctx);
    const int64_t n = (int64_t)pj[(size_t)d].size();
    if (n == 0) { rcs[(size_t)d] = LRR_OK; return; }
    BatchSlot slot;
    int rc = LRR_OK;
    if (cudaStreamCreateWithFlags(&slot.st, cudaStreamNonBlocking) != cudaSuccess) { rcs[(size_t)d] = LRR_ERR_CUDA; ctx->err = "cudaStreamCreate"; return; }
    slot.own_stream = true;
    cudaStreamSynchronize(ctx->stream);   // the uploads of the replicas are complete
    const cudaStream_t keep = g_alloc_stream;
    rc = slot_launch(ctx, slot, moving_all[d], fixed_all[d], pj[(size_t)d].data(), 0, n, pid[(size_t)d].data(), 0, max_iter, shuffle_seed, pout[(size_t)d].data());
    g_alloc_stream = keep;
    { const int r2 = slot_collect(ctx, slot); if (rc == LRR_OK) rc = r2; }
    cudaStreamDestroy(slot.st);
    rcs[(size_t)d] = rc;
  };
  std::vector<std::thread> th;
  for (int d = 1; d < ndev; ++d) th.emplace_back(worker, d);
  worker(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < ndev; ++d) if (rcs[(size_t)d] != LRR_OK) return rcs[(size_t)d];
  for (int d = 0; d < ndev; ++d) for (size_t i = 0; i < pid[(size_t)d].size(); ++i) out[pid[(size_t)d][i]] = pout[(size_t)d][i];
  std::vector<int64_t> todo;
  for (int64_t j = 0; j < njobs; ++j) if (out[j].status < 0) todo.push_back(j);
  enter(ctxs[0]);
  return batch_host_jobs(ctxs[0], moving_all[0], fixed_all[0], jobs, todo, nullptr, 0, max_iter, shuffle_seed, out);
}

// The driver's candidate loop on the device (include/lrr_gpu.h): one launch ranks every candidate (find_closest_euclidean + four
// angle measures), the host orders each location's candidates by the mean angle (std::sort of rrl/itkLocationRegistration.cxx:2326;
// ties keep the candidate order), a second launch runs one LOCATION per thread block, best angle first, with the decision
// measurements of every converged, non-oscillating registration.
extern "C" int lrr_batch_run_locations(lrr_ctx* ctx, lrr_fset* moving_all, lrr_fset* fixed_all, const lrr_job* jobs, int64_t n_locations, int32_t n_candidates,
                                       int max_iter, uint64_t shuffle_seed, int until_verified, double voxel_spacing, lrr_result* out, lrr_decision* decisions) {
  if (!ctx || !moving_all || !fixed_all || n_locations < 0 || n_candidates <= 0 || max_iter > 250) return LRR_ERR_INVALID;
  if (moving_all->ctx != ctx || fixed_all->ctx != ctx) return LRR_ERR_INVALID;
  const int64_t njobs = n_locations * (int64_t)n_candidates;
  if (njobs > 0 && (!jobs || !out || !decisions)) return LRR_ERR_INVALID;
  if (njobs == 0) return LRR_OK;
  if (!(voxel_spacing > 0.0)) voxel_spacing = 1.0;
  enter(ctx);
  cudaStream_t st = ctx->stream;
  const int per_sm = lrr_batch_blocks_per_sm();
  const int nb_rank = (int)std::min<int64_t>(njobs, (int64_t)lrr_num_sms() * per_sm), nb_run = (int)std::min<int64_t>(n_locations, (int64_t)lrr_num_sms() * per_sm);
  const size_t wbytes = lrr_batch_workspace_bytes(std::max(nb_rank, nb_run), (int)moving_all->n, (int)fixed_all->n);
  void* work = nullptr; lrr_job* djobs = nullptr; lrr_result* dout = nullptr; lrr_decision* ddec = nullptr; int* dorder = nullptr;
  int rc = LRR_OK;
  auto fail = [&](int code, const char* what) { if (rc == LRR_OK) { rc = code; if (what) ctx->err = what; } };
  auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && rc == LRR_OK) { rc = LRR_ERR_CUDA; ctx->err = std::string(what) + ": " + cudaGetErrorString(e); } };
  cu(cudaMallocAsync(&work, wbytes, st), "workspace");
  cu(cudaMallocAsync((void**)&djobs, sizeof(lrr_job) * njobs, st), "jobs"); cu(cudaMallocAsync((void**)&dout, sizeof(lrr_result) * njobs, st), "results");
  cu(cudaMallocAsync((void**)&ddec, sizeof(lrr_decision) * njobs, st), "decisions"); cu(cudaMallocAsync((void**)&dorder, sizeof(int) * njobs, st), "order");
  double table[4 * 32] = {0};
  const int rows = build_muse_table(1000u, table);
  if (rc == LRR_OK) {
    cu(cudaMemcpyAsync(djobs, jobs, sizeof(lrr_job) * njobs, cudaMemcpyHostToDevice, st), "jobs H2D");
    // ---- stage 1: ranking ----
    lrr_batch_loc_args = BatchLocArgs{1, n_candidates, 0, nullptr, ddec, voxel_spacing};
    fset_wait_dirs(moving_all, st); fset_wait_dirs(fixed_all, st);
    { PHASE(PH_BATCH); cu(lrr_batch_launch(moving_all->dev, fixed_all->dev, djobs, dout, njobs, nullptr, max_iter, shuffle_seed, table, rows, nb_rank, work, wbytes, st, &ctx->launches), "ranking launch"); }
    lrr_batch_loc_args.mode = 0;
    cu(cudaMemcpyAsync(out, dout, sizeof(lrr_result) * njobs, cudaMemcpyDeviceToHost, st), "results D2H");
    cu(cudaMemcpyAsync(decisions, ddec, sizeof(lrr_decision) * njobs, cudaMemcpyDeviceToHost, st), "decisions D2H");
    cu(cudaStreamSynchronize(st), "ranking");
  }
  std::vector<int> order((size_t)njobs, -1);
  if (rc == LRR_OK) {
    for (int64_t l = 0; l < n_locations && rc == LRR_OK; ++l) {
      std::vector<std::pair<double, int> > cand;
      for (int c = 0; c < n_candidates; ++c) {
        const int64_t j = l * n_candidates + c;
        if (out[j].status < 0) fail(LRR_ERR_UNSUPPORTED, "lrr_batch_run_locations: an ROI holds more features than the block kernel takes (16384)");
        if (out[j].status == 0) cand.push_back(std::make_pair(decisions[j].rank_angle_average, c));
      }
      std::stable_sort(cand.begin(), cand.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; });
      for (size_t k = 0; k < cand.size(); ++k) { order[(size_t)(l * n_candidates) + k] = cand[k].second; decisions[l * n_candidates + cand[k].second].rank = (int32_t)k; }
    }
  }
  if (rc == LRR_OK) {
    // ---- stage 2: one location per block, best angle first ----
    cu(cudaMemcpyAsync(dorder, order.data(), sizeof(int) * njobs, cudaMemcpyHostToDevice, st), "order H2D");
    cu(cudaMemcpyAsync(ddec, decisions, sizeof(lrr_decision) * njobs, cudaMemcpyHostToDevice, st), "decisions H2D");
    // until verified: a unit of work is a LOCATION (its candidates one after the other, best angle first, stop at the first verified one).
    // All candidates: nothing orders the registrations of a location, so every candidate is its own unit — 8192 units on 148 blocks instead
    // of 256 (1.7 waves of 32-registration units: 2.32 s for the C4 list against 1.43 s for the plain list)
    const bool per_job = !until_verified;
    lrr_batch_loc_args = BatchLocArgs{2, per_job ? 1 : n_candidates, until_verified ? 1 : 0, per_job ? nullptr : dorder, ddec, voxel_spacing};
    { PHASE(PH_BATCH); cu(lrr_batch_launch(moving_all->dev, fixed_all->dev, djobs, dout, njobs, nullptr, max_iter, shuffle_seed, table, rows, per_job ? nb_rank : nb_run, work, wbytes, st, &ctx->launches), "registration launch"); }
    lrr_batch_loc_args.mode = 0;
    cu(cudaMemcpyAsync(out, dout, sizeof(lrr_result) * njobs, cudaMemcpyDeviceToHost, st), "results D2H");
    cu(cudaMemcpyAsync(decisions, ddec, sizeof(lrr_decision) * njobs, cudaMemcpyDeviceToHost, st), "decisions D2H");
    cu(cudaStreamSynchronize(st), "registrations");
    cu(cudaGetLastError(), "registrations");
    for (int64_t j = 0; j < njobs && rc == LRR_OK; ++j) if (out[j].status < 0) fail(LRR_ERR_STATE, "lrr_batch_run_locations: order-statistic selector overflow in the block kernel");
  }
  lrr_batch_loc_args.mode = 0;
  cudaStreamSynchronize(st);
  if (work) cudaFreeAsync(work, st); if (djobs) cudaFreeAsync(djobs, st); if (dout) cudaFreeAsync(dout, st); if (ddec) cudaFreeAsync(ddec, st); if (dorder) cudaFreeAsync(dorder, st);
  return rc;
}
