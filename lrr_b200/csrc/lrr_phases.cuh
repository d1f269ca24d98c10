// Phase bodies of one robust-ICP iteration, written once against a reduction Scope (lrr_dev.cuh) so that
// the same code runs (a) as one kernel per phase over the whole GPU for a single large registration and
// (b) inside one thread block per registration in batched mode.
//
// One iteration = MATCH -> NORM1 -> NORM2 -> WEIGHT -> COLLECT -> ACCUMULATE(+SOLVE) -> RESID; each phase
// ends in a global reduction whose finalizer writes the derived scalars into the direction's Scal.
// Reference maths: SURVEY.md §3.6; citations at each phase.
#pragma once
#include "lrr_dev.cuh"
#include "lrr_linalg.cuh"

#define LRR_THREADS 256
#define LRR_SPATIAL(flags) (((flags) & LRR_F_SYM) && !((flags) & LRR_F_NOSPATIAL))   // weight_spatially is part of this estimate
#define LRR_NWARPS (LRR_THREADS / 32)

template <int NW>
struct PhaseSharedT {
  double out[LRR_NRED];
  double tmp[LRR_NRED * NW];
};
typedef PhaseSharedT<LRR_NWARPS> PhaseShared;

// position k of the (possibly permuted) moving vector -> compact index.  RESHUFFLE = keyed 4-round Feistel
// permutation with cycle walking, evaluated on the fly (the stand-in for vcl_random_shuffle, Q4).
__device__ __forceinline__ int lrr_from_index(const DirArgs& a, int k) {
  if (a.order) return __ldg(a.order + k);
  if (a.perm_half == 0) return k;
  const unsigned int hmask = (1u << a.perm_half) - 1u, n = (unsigned int)a.n_in;
  unsigned int x = (unsigned int)k;
  do {
    unsigned int L = x >> a.perm_half, R = x & hmask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      unsigned int F = R * 0x9E3779B1u + a.perm_key[r];
      F ^= F >> 15; F *= 0x85EBCA6Bu; F ^= F >> 13;
      F &= hmask;
      const unsigned int nl = R; R = L ^ F; L = nl;
    }
    x = (L << a.perm_half) | R;
  } while (x >= n);
  return (int)x;
}

// location of the match's from_ feature: the original, or — while a rigid increment is being estimated —
// the clone mapped through the previous global transform (cdcl :455-484)
__device__ __forceinline__ void lrr_from_loc(const DirArgs& a, const Scal* __restrict__ sc, int mapped, int i, double& x, double& y,
                                             double& z) {
  x = __ldg(a.fx + i); y = __ldg(a.fy + i); z = __ldg(a.fz + i);
  if (mapped) { double ox, oy, oz; lrr_map_exact(sc->Ag, sc->tg, sc->cg, x, y, z, ox, oy, oz); x = ox; y = oy; z = oz; }
  if (mapped == 2) { double ox, oy, oz; lrr_map_exact(sc->Ag2, sc->tg2, sc->cg2, x, y, z, ox, oy, oz); x = ox; y = oy; z = oz; }
}

__device__ __forceinline__ void lrr_ld_P6(const double* __restrict__ rec, int s, double (&P)[6]) {
  const double2* q = reinterpret_cast<const double2*>(rec + 8 * (size_t)s);
  double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  P[0] = a.x; P[1] = a.y; P[2] = b.x; P[3] = b.y; P[4] = c.x; P[5] = c.y;
}

// ---- MATCH: find_closest_euclidean (cdcl :541-654, rrl :494-571) fused with the residuals of
// RMS_error (cdcl :196-230) and the first normalisation sum (cdcl :1159-1190) ---------------------------
// residual, stamp and sums of one finished match
// pmode: what happens to the neighbour certificate of from-feature i (DirArgs::proof): 0 = none comes with this match (any old one
// is void), 1 = the match was accepted BY the certificate (it stays as it is), 2 = new certificate (bound) at this mapped position
__device__ __forceinline__ void lrr_match_finish(const DirArgs& a, bool rigid, int k, int i, double x, double y, double z, double yx,
                                                 double yy, double yz, int s, double (&acc)[4], int pmode = 0, double bound = 0.0) {
  a.sidx[k] = s;
  a.prev_s[i] = s;   // candidate of this from-feature for the next find_closest
  if (a.proof && pmode != 1) {   // (certificates live at the feature's tile-order position, with a mirror of prev_s: DirArgs::pos_of)
    const int tq = __ldg(a.pos_of + i);
    if (pmode == 2) { double2* pr = reinterpret_cast<double2*>(a.proof + tq); pr[0] = make_double2(yx, yy); pr[1] = make_double2(yz, bound); }
    else reinterpret_cast<double*>(a.proof + tq)[3] = 0.0;
    a.cprev[tq] = s;
  }
  double r2 = 0.0;
  if (s >= 0) {
    const double4 q = lrr_ld_pos4(a.to.grid.pos4 + s);
    double P[6]; lrr_ld_P6(a.to.rec, s, P);
    r2 = lrr_proj_r2(yx - q.x, yy - q.y, yz - q.z, P[0], P[1], P[2], P[3], P[4], P[5]);
    if (a.stamp_in_match) atomicMin(a.stamp + s, a.tag | (unsigned long long)(unsigned int)k);
  }
  a.r2[k] = r2;
  a.w[k] = 1.0;   // cdcl :645
  if (rigid) { acc[0] += yx; acc[1] += yy; acc[2] += yz; } else { acc[0] += x; acc[1] += y; acc[2] += z; }
  acc[3] += r2;
}

// stage the (uniform) transform of a direction in shared memory: broadcast loads instead of 30 registers per thread
__device__ __forceinline__ void lrr_stage_transform(const Scal* sc, double* s_T /* 15 doubles shared */) {
  __syncthreads();
  if (threadIdx.x < 15) s_T[threadIdx.x] = threadIdx.x < 9 ? sc->A[threadIdx.x] : (threadIdx.x < 12 ? sc->t[threadIdx.x - 9] : sc->c[threadIdx.x - 12]);
  __syncthreads();
}

// pass A: one thread per query, ball scan bounded by the previous neighbour of the same from-feature;
// queries without a candidate take the 27 cells around them; unfinished queries go to the direction's queue
__device__ __forceinline__ void lrr_match_a_one(const DirArgs& a, int flags, const double* s_T, const int* __restrict__ proc, int tq, double (&acc)[4]) {
  const double* A = s_T; const double* t = s_T + 9; const double* c = s_T + 12;
  const bool rigid = (flags & LRR_F_RIGID) != 0;
  const GridDev& g = a.to.grid;
  const int k = proc ? __ldg(proc + tq) : tq;
  const int i = lrr_from_index(a, k);
  const double x = __ldg(a.fx + i), y = __ldg(a.fy + i), z = __ldg(a.fz + i);
  double yx, yy, yz;
  lrr_map_exact(A, t, c, x, y, z, yx, yy, yz);
  NNState st; st.best = 1e300; st.best_s = a.prev_s[i]; st.best_o = 0x7fffffff;
  bool done = false;
  double bound = 0.0;
  const bool cert = a.proof && a.todo && a.M == a.n_in && proc;   // certificates are read at the finest level in tile order only
  if (st.best_s >= 0) {
    const double4 p = lrr_ld_pos4(g.pos4 + st.best_s);
    st.best = lrr_dist2_exact(yx, yy, yz, p.x, p.y, p.z); st.best_o = lrr_pos4_orig(p);
    const int max_rows = a.ball_rows > 0 ? a.ball_rows : LRR_BALL_ROWS;
    done = cert ? lrr_nn_ball_thread_cert(g, yx, yy, yz, st, 0.125 * g.h, &bound, max_rows) : lrr_nn_ball_thread(g, yx, yy, yz, st, max_rows);
  } else if (g.n > 0) {
    const int r27 = lrr_nn_block27_thread(g, yx, yy, yz, st);
    if (r27 == 1) done = true;
    else if (r27 == 0) done = lrr_nn_ball_thread(g, yx, yy, yz, st);   // rescans the block's rows inside the ball: still exact
  }
  if (done) lrr_match_finish(a, rigid, k, i, x, y, z, yx, yy, yz, st.best_s, acc, (bound > 0.0 && i == k) ? 2 : 0, bound);
  else {
    const int slot = atomicAdd(a.hq_count, 1);
    a.hq_k[slot] = k; a.hq_best[slot] = st.best; a.hq_s[slot] = st.best_s; a.hq_o[slot] = st.best_o;
  }
}

template <class Scope>
__device__ __forceinline__ void lrr_match_a_body(const Scope& scope, const DirArgs& a, int flags, const double* s_T, double (&acc)[4]) {
  const int* __restrict__ proc = (a.proc && a.M == a.n_in) ? a.proc : nullptr;
  for (int tq = scope.first(); tq < a.M; tq += scope.stride()) lrr_match_a_one(a, flags, s_T, proc, tq, acc);
}
// persistent-kernel variant: warps pull 32-query chunks from a global counter, so blocks whose queries fall into
// dense regions of the fixed set do not make everybody else wait at the next grid barrier
__device__ __forceinline__ void lrr_match_a_body_dynamic(const DirArgs& a, int flags, const double* s_T, unsigned int* counter, double (&acc)[4]) {
  const int* __restrict__ proc = (a.proc && a.M == a.n_in) ? a.proc : nullptr;
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned int chunk = 0;
    if (lane == 0) chunk = atomicAdd(counter, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    const long long base = (long long)chunk * 32;
    if (base >= a.M) break;
    const int tq = (int)base + lane;
    if (tq < a.M) lrr_match_a_one(a, flags, s_T, proc, tq, acc);
    __syncwarp();
  }
}

// one queued query, whole warp.  rowtab: the warp's row table in shared memory (2 x (LRR_FLAT_ROWS + 1) ints) for the flattened
// ball scan, or null (batched mode: the lane-per-row scan)
__device__ __forceinline__ void lrr_match_b_one(const DirArgs& a, int flags, const double* s_T, int qi, double (&acc)[4], int* rowtab = nullptr) {
  const double* A = s_T; const double* t = s_T + 9; const double* c = s_T + 12;
  const bool rigid = (flags & LRR_F_RIGID) != 0;
  const int lane = threadIdx.x & 31;
  const int kq = a.hq_k[qi];
  const int i = lrr_from_index(a, kq);
  const double x = __ldg(a.fx + i), y = __ldg(a.fy + i), z = __ldg(a.fz + i);
  double yx, yy, yz;
  lrr_map_exact(A, t, c, x, y, z, yx, yy, yz);
  NNState st; st.best = a.hq_best[qi]; st.best_s = a.hq_s[qi]; st.best_o = a.hq_o[qi];
  int s; double bound = 0.0;
  // (certificates only where they are read: finest level in tile order; margin = an eighth of a cell, far above the per-iteration
  //  motion near convergence and ~1.4x the ball volume at worst)
  const bool cert = a.proof && a.todo && a.M == a.n_in;
  if (rowtab) {
    bool have = st.best_s >= 0;
    if (!have) have = lrr_nn_shell_first(a.to.grid, yx, yy, yz, st.best, st.best_s, st.best_o);   // (certificate next time, as before)
    const bool had = a.hq_s[qi] >= 0;
    s = have ? lrr_nn_ball_flat(a.to.grid, yx, yy, yz, st, (cert && had) ? 0.125 * a.to.grid.h : 0.0, &bound, rowtab, rowtab + LRR_FLAT_ROWS + 1) : -1;
    if (!(cert && had)) bound = 0.0;
  }
  else if (cert) s = lrr_nn_pass_b_cert(a.to.grid, yx, yy, yz, st, 0.125 * a.to.grid.h, &bound);
  else s = lrr_nn_pass_b(a.to.grid, yx, yy, yz, st);
  if (lane == 0) lrr_match_finish(a, rigid, kq, i, x, y, z, yx, yy, yz, s, acc, (bound > 0.0 && i == kq) ? 2 : 0, bound);
}
// first_item >= 0: the warp's first query is assigned statically (its number among the nstatic warps of the launch) and only the
// following ones are pulled from the counter — a launch whose queue is shorter than the number of warps (the steady state: empty)
// does not touch the counter at all (a few thousand same-address atomics are several microseconds of an otherwise empty kernel)
__device__ __forceinline__ void lrr_match_b_body_dynamic(const DirArgs& a, int flags, const double* s_T, unsigned int* counter, double (&acc)[4],
                                                         int first_item = -1, int nstatic = 0, int* rowtab = nullptr) {
  const int lane = threadIdx.x & 31;
  const int nq = *a.hq_count;
  if (nq <= 0) return;
  unsigned int qi = (unsigned int)first_item;
  bool pull = first_item < 0;
  for (;;) {
    if (pull) {
      if (lane == 0) qi = atomicAdd(counter, 1u) + (unsigned int)nstatic;
      qi = __shfl_sync(0xffffffffu, qi, 0);
    }
    pull = true;
    if ((int)qi >= nq) break;
    lrr_match_b_one(a, flags, s_T, (int)qi, acc, rowtab);
    __syncwarp();
  }
}

// pass B: one warp per queued query
template <class Scope>
__device__ __forceinline__ void lrr_match_b_body(const Scope& scope, const DirArgs& a, int flags, const double* s_T, double (&acc)[4]) {
  const int nq = *a.hq_count;
  for (int qi = scope.first() >> 5; qi < nq; qi += scope.stride() >> 5) lrr_match_b_one(a, flags, s_T, qi, acc);
}

// bookkeeping once every match of the direction exists (sums of both passes in `tot`)
__device__ __forceinline__ void lrr_match_finalize(const DirArgs& a, int flags, const double* tot) {
  Scal* sc = a.sc;
  sc->sum_loc[0] = tot[0]; sc->sum_loc[1] = tot[1]; sc->sum_loc[2] = tot[2]; sc->sum_r2 = tot[3];
  sc->res_sum_r2 = tot[3]; sc->res_count = (double)a.M;
  sc->passb_queries = (double)*a.hq_count;
  *a.hq_count = 0;
  if (flags & LRR_F_RIGID) {   // make_incremental: the matches' from_ are now clones mapped through the current transform
    for (int i = 0; i < 9; ++i) sc->Ag[i] = sc->A[i];
    for (int i = 0; i < 3; ++i) { sc->tg[i] = sc->t[i]; sc->cg[i] = sc->c[i]; }
    sc->from_mapped = 1;
  } else sc->from_mapped = 0;
}

// `counter` (grid scope only): work counter for dynamic 32-query chunks — warps that drew cheap queries keep pulling
// instead of waiting at the block reduction for the warp that drew a dense region (23 % barrier stalls before).
template <class Scope, class SH>
__device__ void lrr_phase_match(const Scope& scope, const DirArgs& a, int flags, SH& sh, unsigned int* counter = nullptr) {
  __shared__ double s_T[15];
  lrr_stage_transform(a.sc, s_T);
  double acc[4] = {0, 0, 0, 0};
  if (counter) lrr_match_a_body_dynamic(a, flags, s_T, counter, acc);
  else lrr_match_a_body(scope, a, flags, s_T, acc);
  lrr_tl(0, 1);
  if (scope.template reduce<4>(acc, sh.out, sh.tmp)) {
    lrr_tl(0, 2);
    if (threadIdx.x == 0) {
      Scal* sc = a.sc; sc->sum_loc[0] = sh.out[0]; sc->sum_loc[1] = sh.out[1]; sc->sum_loc[2] = sh.out[2]; sc->sum_r2 = sh.out[3];
      sc->norm_dirty = 1; sc->match_left = 1;
      if (counter) *counter = 0;
    }
  }
  scope.phase_end();
}

template <class Scope, class SH>
__device__ void lrr_phase_match_hard(const Scope& scope, const DirArgs& a, int flags, SH& sh, unsigned int* counter = nullptr) {
  __shared__ double s_T[15];
  lrr_stage_transform(a.sc, s_T);
  double acc[4] = {0, 0, 0, 0};
  if (counter) lrr_match_b_body_dynamic(a, flags, s_T, counter, acc);
  else lrr_match_b_body(scope, a, flags, s_T, acc);
  lrr_tl(1, 1);
  if (scope.template reduce<4>(acc, sh.out, sh.tmp)) {
    lrr_tl(1, 2);
    if (threadIdx.x == 0) {
      const Scal* sc = a.sc;
      const double tot[4] = {sc->sum_loc[0] + sh.out[0], sc->sum_loc[1] + sh.out[1], sc->sum_loc[2] + sh.out[2], sc->sum_r2 + sh.out[3]};
      lrr_match_finalize(a, flags, tot);
      if (counter) *counter = 0;
    }
  }
  scope.phase_end();
}

// Batched mode: the queue pass A leaves (queries whose ball touches more than DirArgs::ball_rows rows) is drained by GROUPS OF
// EIGHT LANES (lrr_nn_ball_g8, lrr_dev.cuh): 64 queries of a 512-thread block in flight, two round trips each, instead of 16 warps
// walking their rows lane by lane.  What a group cannot take (no candidate yet, a ball of more than 32 rows) moves on to the second
// queue `a2` (same DirArgs with the hq_* members pointing at it) for the one-warp-per-query pass.  rows: 64 ints per group.
template <class Scope, class SH>
__device__ void lrr_phase_match_queue_g8(const Scope& scope, const DirArgs& a, const DirArgs& a2, int flags, SH& sh, int* rows) {
  __shared__ double s_T[15];
  lrr_stage_transform(a.sc, s_T);
  const double* A = s_T; const double* t = s_T + 9; const double* c = s_T + 12;
  const bool rigid = (flags & LRR_F_RIGID) != 0;
  const GridDev& g = a.to.grid;
  const int lane = threadIdx.x & 31, gl = lane & 7, gi = lane >> 3;
  const unsigned int gmask = 0xffu << (8 * gi);
  int* rowlo = rows + ((threadIdx.x >> 5) * 4 + gi) * 2 * LRR_G8_ROWS; int* rowbase = rowlo + LRR_G8_ROWS;
  const int nq = *a.hq_count;
  double acc[4] = {0, 0, 0, 0};
  for (int qi = scope.first() >> 3; qi < nq; qi += scope.stride() >> 3) {
    const int kq = a.hq_k[qi];
    const int i = lrr_from_index(a, kq);
    const double x = __ldg(a.fx + i), y = __ldg(a.fy + i), z = __ldg(a.fz + i);
    double qx, qy, qz;
    lrr_map_exact(A, t, c, x, y, z, qx, qy, qz);
    const int cand = a.hq_s[qi]; const double cbest = a.hq_best[qi]; const int co = a.hq_o[qi];
    double best = cbest, bound; int best_s = cand, best_o = co;
    const bool ok = cand >= 0 && lrr_nn_ball_g8(g, qx, qy, qz, 0.0, rowlo, rowbase, gmask, gl, best, best_s, best_o, bound);
    if (gl == 0) {
      if (ok) lrr_match_finish(a, rigid, kq, i, x, y, z, qx, qy, qz, best_s, acc, 0, 0.0);
      else { const int slot = atomicAdd(a2.hq_count, 1); a2.hq_k[slot] = kq; a2.hq_best[slot] = cbest; a2.hq_s[slot] = cand; a2.hq_o[slot] = co; }
    }
  }
  if (scope.template reduce<4>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x == 0) {
      Scal* sc = a.sc;
      sc->sum_loc[0] += sh.out[0]; sc->sum_loc[1] += sh.out[1]; sc->sum_loc[2] += sh.out[2]; sc->sum_r2 += sh.out[3];
      *a.hq_count = 0;
    }
  }
  scope.phase_end();
}

// ---- sheet_angles / tube_angles of the rrl estimators (rrl/rrl_estimation_ICP_matching_all.txx:281-375): weighted mean over the
// SHEET-SHEET (which = 0) / TUBE-TUBE (1) matches of acos |< T-mapped normal / tangent of p, normal / tangent of q >| folded to
// [0, pi/2].  Scope form for the batched mode (the single-registration path has its own kernel, lrr_metrics.cu).  Result in
// sc->ang_num / sc->ang_den.
template <class Scope, class SH>
__device__ void lrr_phase_angles_rrl(const Scope& scope, const DirArgs& a, int which, SH& sh) {
  Scal* sc = a.sc;
  const int want = which == 0 ? 2 : 1;
  const int mapped = sc->from_mapped;
  const FsetDev& F = a.from_all; const FsetDev& T = a.to;
  double acc[2] = {0, 0};
  if (F.dirs && T.dirs) for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int s = a.sidx[k];
    if (s < 0) continue;
    const int i = lrr_from_index(a, k);
    const int po = a.forig ? a.forig[i] : i;
    const int qo = lrr_pos4_orig(lrr_ld_pos4(T.grid.pos4 + s));
    if (F.shape[po] != want || T.shape[qo] != want) continue;
    double pd[3], qd[3];
    if (which == 0) {
      for (int d = 0; d < 3; ++d) { pd[d] = F.dirs[(size_t)d * F.n + po]; qd[d] = T.dirs[(size_t)d * T.n + qo]; }
    } else {   // tangent = d0 x d1
      double p0[3], p1[3], q0[3], q1[3];
      for (int d = 0; d < 3; ++d) {
        p0[d] = F.dirs[(size_t)d * F.n + po]; p1[d] = F.dirs[(size_t)(3 + d) * F.n + po];
        q0[d] = T.dirs[(size_t)d * T.n + qo]; q1[d] = T.dirs[(size_t)(3 + d) * T.n + qo];
      }
      pd[0] = p0[1] * p1[2] - p0[2] * p1[1]; pd[1] = p0[2] * p1[0] - p0[0] * p1[2]; pd[2] = p0[0] * p1[1] - p0[1] * p1[0];
      qd[0] = q0[1] * q1[2] - q0[2] * q1[1]; qd[1] = q0[2] * q1[0] - q0[0] * q1[2]; qd[2] = q0[0] * q1[1] - q0[1] * q1[0];
    }
    double x, y, z; lrr_from_loc(a, sc, mapped, i, x, y, z);
    double y0, y1, y2, n0, n1, n2;
    lrr_map_exact(sc->A, sc->t, sc->c, x, y, z, y0, y1, y2);
    lrr_map_exact(sc->A, sc->t, sc->c, x + pd[0], y + pd[1], z + pd[2], n0, n1, n2);
    double d0 = n0 - y0, d1 = n1 - y1, d2 = n2 - y2;
    const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    d0 /= nrm; d1 /= nrm; d2 /= nrm;
    double ang = acos(d0 * qd[0] + d1 * qd[1] + d2 * qd[2]);
    if (ang > 3.14159265358979323846 / 2) ang = 3.14159265358979323846 - ang;
    const double w = a.w[k];
    acc[0] += w * ang; acc[1] += w;
  }
  if (scope.template reduce<2>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x == 0) { sc->ang_num = sh.out[0]; sc->ang_den = sh.out[1]; }
  }
  scope.phase_end();
}

// ---- NORM1 / NORM2: cdcl_normalize_matches (cdcl :1159-1240, rrl :864-943).  The std::set of matched
// fixed features becomes "the match with the lowest position wins the stamp of its to-feature". ----------
template <class Scope, class SH>
__device__ void lrr_phase_norm1(const Scope& scope, const DirArgs& a, SH& sh) {
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  const double invM = 1.0;
  (void)invM;
  const double cm0 = sc->sum_loc[0] / (double)a.M, cm1 = sc->sum_loc[1] / (double)a.M, cm2 = sc->sum_loc[2] / (double)a.M;
  double acc[5] = {0, 0, 0, 0, 0};
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int i = lrr_from_index(a, k);
    double x, y, z; lrr_from_loc(a, sc, mapped, i, x, y, z);
    const double dx = x - cm0, dy = y - cm1, dz = z - cm2;
    acc[4] += sqrt(dx * dx + dy * dy + dz * dz);
    const int s = a.sidx[k];
    if (s >= 0 && a.stamp[s] == (a.tag | (unsigned long long)(unsigned int)k)) {
      const double4 q = lrr_ld_pos4(a.to.grid.pos4 + s);
      acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += 1.0;
    }
  }
  lrr_tl(6, 1);
  if (scope.template reduce<5>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x == 0) {
      sc->sum_q[0] = sh.out[0]; sc->sum_q[1] = sh.out[1]; sc->sum_q[2] = sh.out[2]; sc->n_unique = sh.out[3];
      sc->sum_rad_from = sh.out[4];
      sc->cfrom[0] = cm0; sc->cfrom[1] = cm1; sc->cfrom[2] = cm2;
      sc->rfrom = sh.out[4] / (double)a.M;
      for (int d = 0; d < 3; ++d) sc->cto[d] = sh.out[d] / sh.out[3];
    }
  }
  scope.phase_end();
}

template <class Scope, class SH>
__device__ void lrr_phase_norm2(const Scope& scope, const DirArgs& a, Scal* other /* null or the opposite direction */, SH& sh) {
  Scal* sc = a.sc;
  const double c0 = sc->cto[0], c1 = sc->cto[1], c2 = sc->cto[2];
  double acc[1] = {0};
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int s = a.sidx[k];
    if (s >= 0 && a.stamp[s] == (a.tag | (unsigned long long)(unsigned int)k)) {
      const double4 q = lrr_ld_pos4(a.to.grid.pos4 + s);
      const double dx = q.x - c0, dy = q.y - c1, dz = q.z - c2;
      acc[0] += sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
  lrr_tl(7, 1);
  if (scope.template reduce<1>(acc, sh.out, sh.tmp)) {
    if (threadIdx.x == 0) {
      sc->sum_rad_to = sh.out[0];
      sc->rto = sh.out[0] / sc->n_unique;
      sc->norm_dirty = 0;
      if (other) {   // cdcl_normalize_matches_known with the centres and radii swapped (cdcl sym :470-478)
        for (int d = 0; d < 3; ++d) { other->cfrom[d] = sc->cto[d]; other->cto[d] = sc->cfrom[d]; }
        other->rfrom = sc->rto; other->rto = sc->rfrom;
      }
    }
  }
  scope.phase_end();
}

// ---- exact order statistic: histogram side --------------------------------------------------------------
__device__ __forceinline__ int lrr_sel_bin(const SelDev& s, double key) {
  int b = (int)(key * s.inv_width);
  return b < 0 ? 0 : (b >= LRR_SEL_FINE ? LRR_SEL_FINE - 1 : b);
}
// Histogram update with warp aggregation: lanes that hit the same fine bin (the strength cap puts ~1.5 % of all
// keys into the last bin) elect a leader that adds their count once, so no address sees more than one atomic
// per warp.  Called under divergence: the mask is the set of currently active lanes.
__device__ __forceinline__ void lrr_sel_add(const SelDev& s, double key) {
  const int b = lrr_sel_bin(s, key);
  const unsigned int active = __activemask();
  const unsigned int same = __match_any_sync(active, b);
  const int lane = threadIdx.x & 31;
  if (lane == __ffs(same) - 1) atomicAdd(s.fine + b, (unsigned int)__popc(same));
  const int cb = b >> 6;
  const unsigned int samec = __match_any_sync(active, cb);
  if (lane == __ffs(samec) - 1) atomicAdd(s.coarse + cb, (unsigned int)__popc(samec));
}
// Called by ONE block (all LRR_THREADS threads) once every key of the group is in the histogram: locate
// the fine bin of the (n-1)/2+1 -th smallest key (0-based) — the element the reference's private
// rrel_util_median_abs_dev_scale picks with nth_element (cdcl :520-536, Q12).
static __device__ void lrr_sel_locate(const SelDev& s, double n_keys, double* shd /* >= 2*LRR_THREADS doubles */, bool reset_cand = true) {
  unsigned int* shu = reinterpret_cast<unsigned int*>(shd);
  __shared__ unsigned int s_target, s_cbin, s_cbase;
  const long long n = (long long)n_keys;
  if (threadIdx.x == 0) { *s.count = n_keys; if (reset_cand) { *s.cand_count = 0; s.cand_minmax[0] = ~0ull; s.cand_minmax[1] = 0ull; } }
  if (n <= 1) { if (threadIdx.x == 0) { *s.bin = -1; *s.rank = 0; *s.result = 0.0; } __syncthreads(); return; }
  const unsigned int target = (unsigned int)((n - 1) / 2 + 1);
  // coarse level: 4 bins per thread, inclusive scan over threads
  unsigned int c4[4], loc = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { c4[j] = __ldcg(s.coarse + 4 * threadIdx.x + j); loc += c4[j]; }
  shu[threadIdx.x] = loc;
  __syncthreads();
  for (int o = 1; o < LRR_THREADS; o <<= 1) {
    unsigned int v = (threadIdx.x >= o) ? shu[threadIdx.x - o] : 0;
    __syncthreads();
    shu[threadIdx.x] += v;
    __syncthreads();
  }
  const unsigned int incl = shu[threadIdx.x], excl = incl - loc;
  if (target >= excl && target < incl) {
    unsigned int base = excl; int cb = 4 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) { if (target >= base + c4[j]) { base += c4[j]; ++cb; } else break; }
    s_cbin = cb; s_cbase = base; s_target = target;
  }
  __syncthreads();
  // fine level inside the coarse bin: 64 bins, one per thread of the first two warps
  unsigned int f = 0;
  if (threadIdx.x < 64) f = __ldcg(s.fine + s_cbin * 64 + threadIdx.x);
  shu[threadIdx.x] = f;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int base = s_cbase; int fb = 0;
    for (; fb < 64; ++fb) { if (s_target >= base + shu[fb]) base += shu[fb]; else break; }
    *s.bin = s_cbin * 64 + fb;
    *s.rank = s_target - base;
  }
  __syncthreads();
}
__device__ __forceinline__ void lrr_sel_offer(const SelDev& s, int bin, double key) {
  if (lrr_sel_bin(s, key) != bin) return;
  const unsigned int idx = atomicAdd(s.cand_count, 1u);
  if (idx < LRR_SEL_CAP) s.cand[idx] = key;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(key);
  atomicMin(s.cand_minmax, bits); atomicMax(s.cand_minmax + 1, bits);
}
// Called by ONE block after every candidate is in the list: exact rank selection, then the MAD scale.
// Returns (in every thread) false on candidate overflow with distinct values (host falls back to a sort).
static __device__ bool lrr_sel_finish(const SelDev& s, double* sh_cand /* >= LRR_SEL_CAP doubles */, bool reset_cand = true) {
  __shared__ double s_val; __shared__ int s_ok;
  const double n_keys = *s.count;
  if (n_keys <= 1.0) { if (threadIdx.x == 0) *s.result = 0.0; __syncthreads(); return true; }
  const unsigned int cnt = __ldcg(s.cand_count), rank = *s.rank;
  if (threadIdx.x == 0) { s_ok = 1; s_val = 0.0; }
  __syncthreads();
  if (cnt > LRR_SEL_CAP) {
    if (threadIdx.x == 0) {
      const unsigned long long mn = __ldcg(s.cand_minmax), mx = __ldcg(s.cand_minmax + 1);
      if (mn == mx) s_val = __longlong_as_double((long long)mn); else s_ok = 0;
    }
  } else {
    for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x) sh_cand[i] = __ldcg(s.cand + i);
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const double v = sh_cand[i]; unsigned int r = 0;
      for (unsigned int j = 0; j < cnt; ++j) { const double u = sh_cand[j]; r += (u < v) || (u == v && j < i); }
      if (r == rank) s_val = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_ok) {
      *s.result = 1.4826 * (1.0 + 5.0 / (n_keys - 1.0)) * s_val;
      // search window of the next call: sized from the local key density (cnt keys in one fine bin) to hold ~1.5 % of the keys
      double want = n_keys / 512.0; want = want < 64.0 ? 64.0 : (want > 256.0 ? 256.0 : want);
      const double density = (double)(cnt > 0 ? cnt : 1) * s.inv_width;
      double d = s_val > 0.0 ? 0.5 * want / density / s_val : 0.02;
      d = d < 1e-5 ? 1e-5 : (d > 0.25 ? 0.25 : d);
      s.bracket[0] = s_val * (1.0 - d); s.bracket[1] = s_val * (1.0 + d); s.bracket[2] = d; s.bracket[3] = s_val;
    }
    if (reset_cand) *s.cand_count = 0;
  }
  const bool ok = s_ok != 0;
  __syncthreads();
  return ok;
}

// ---- exact order statistic without a histogram: window predicted from the previous iteration ---------------
// ICP changes the key distributions slowly, so the wanted element almost always lies inside a narrow window around
// its previous value.  The weight pass counts the keys below the window and appends the keys inside it (a few hundred
// atomics instead of two per key); if the wanted rank falls into the window the exact element is picked by rank
// counting, otherwise the call reports a miss and the host reruns the estimate on the histogram path.
__device__ __forceinline__ void lrr_sel_window_offer(const SelDev& s, double key, double lo, double hi, double& below) {
  if (key < lo) below += 1.0;
  else if (key <= hi) { const unsigned int slot = atomicAdd(s.cand_count, 1u); if (slot < LRR_SEL_CAP) s.cand[slot] = key; }
}
// Called by ONE block once every key was offered.  Returns (every thread) true on a hit.
// preloaded >= 0: the caller already copied that many candidates into sh_cand (k_weight loads the lists of both selectors in one
// round trip before it finishes them one after the other)
static __device__ bool lrr_sel_window_finish(const SelDev& s, double n_keys, double below, double* sh_cand, int preloaded = -1) {
  __shared__ double s_val; __shared__ int s_hit;
  const long long n = (long long)n_keys;
  const unsigned int cnt = preloaded >= 0 ? (unsigned int)preloaded : __ldcg(s.cand_count);
  if (threadIdx.x == 0) { s_hit = 0; s_val = 0.0; *s.count = n_keys; }
  __syncthreads();
  if (n <= 1) {
    if (threadIdx.x == 0) { *s.result = 0.0; *s.cand_count = 0; }
    __syncthreads();
    return true;
  }
  const long long target = (n - 1) / 2 + 1, nb = (long long)below;
  const bool hit = cnt <= LRR_SEL_CAP && nb <= target && target < nb + (long long)cnt;
  if (hit) {
    const unsigned int rank = (unsigned int)(target - nb);
    if (preloaded < 0) for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x) sh_cand[i] = __ldcg(s.cand + i);
    __syncthreads();
    if (cnt <= 1024) {
      // exact element of the wanted rank by rank counting (broadcast shared-memory reads, no barriers): the window holds
      // a few hundred keys, and 45 barrier-separated bitonic stages cost more than cnt^2 / blockDim comparisons
      for (unsigned int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const double v = sh_cand[i];
        unsigned int r = 0;
        for (unsigned int j = 0; j < cnt; ++j) { const double u = sh_cand[j]; r += (u < v) || (u == v && j < i); }
        if (r == rank) { s_val = v; s_hit = 1; }
      }
      __syncthreads();
    } else {
      // bitonic sort of the window's keys in shared memory, then the element of the wanted rank
      unsigned int np = 1; while (np < cnt) np <<= 1;
      for (unsigned int i = cnt + threadIdx.x; i < np; i += blockDim.x) sh_cand[i] = 1e308;
      __syncthreads();
      for (unsigned int k = 2; k <= np; k <<= 1)
        for (unsigned int jj = k >> 1; jj > 0; jj >>= 1) {
          for (unsigned int i = threadIdx.x; i < np; i += blockDim.x) {
            const unsigned int ixj = i ^ jj;
            if (ixj > i) {
              const double x = sh_cand[i], y = sh_cand[ixj];
              if ((x > y) == ((i & k) == 0)) { sh_cand[i] = y; sh_cand[ixj] = x; }
            }
          }
          __syncthreads();
        }
      if (threadIdx.x == 0) { s_val = sh_cand[rank]; s_hit = 1; }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    if (s_hit) {
      *s.result = 1.4826 * (1.0 + 5.0 / (n_keys - 1.0)) * s_val;
      double want = n_keys / 512.0; want = want < 64.0 ? 64.0 : (want > 256.0 ? 256.0 : want);   // ~0.2 % of the keys inside the window
      double d = s.bracket[2] * want / (double)(cnt > 0 ? cnt : 1);
      d = d < 1e-5 ? 1e-5 : (d > 0.25 ? 0.25 : d);
      s.bracket[0] = s_val * (1.0 - d); s.bracket[1] = s_val * (1.0 + d); s.bracket[2] = d; s.bracket[3] = s_val;
    }
    *s.cand_count = 0;
  }
  const bool res = s_hit != 0;
  __syncthreads();
  return res;
}

// squared scale of the direction in the units the weights are computed in
__device__ __forceinline__ double lrr_scale_of(const Scal* sc, int M, int flags, double unit, int scale_mode, double scale_in) {
  if (scale_mode == 2) return scale_in;                 // injected initial scale, estimator units
  if (scale_mode == 1) return sc->scale / unit;         // MUSE over the unnormalised residuals (written by k_muse)
  if (flags & LRR_F_USE_OLD_W) return sqrt(sc->res_sum_wr2 / sc->res_sum_w) / unit;   // cdcl :698-701,:731
  return sqrt(sc->sum_r2 / (double)M) / unit;           // all weights were reset to 1 (Q2)
}

// ---- WEIGHT: estimate_scale_and_assign_weight (cdcl :660-822, rrl :577-718), the first half of
// weight_by_strength (cdcl :826-862, rrl :722-752) and of weight_spatially (cdcl :973-1004, rrl :756-778) --
// dir = 0/1; reversed = this list joins the merged list with from/to swapped (cdcl sym add_constraints)
template <class Scope, class SH>
__device__ bool lrr_phase_weight(const Scope& scope, const DirArgs& a, int flags, int reversed, const SelDev& selG, const SelDev& selS,
                                 int scale_mode, double scale_in, SH& sh) {
  Scal* sc = a.sc;
  const bool sym = (flags & LRR_F_SYM) != 0, spatial = LRR_SPATIAL(flags);
  const double unit = sym ? 1.0 : sc->rto;   // forward estimators weight in normalised coordinates (Q3)
  // cdcl_estimation_ICP doubles whatever scale it found (`constant = 2.0`, cdcl_estimation_ICP.txx:580-581)
  const double scale = lrr_scale_of(sc, a.M, flags, unit, scale_mode, scale_in) * ((flags & LRR_F_ICP) ? 2.0 : 1.0);
  const double scale2 = scale * scale, unit2 = unit * unit;
  const double inv_unit2 = 1.0 / unit2, inv_scale2 = 1.0 / scale2;
  const int mapped = sc->from_mapped;
  // spatial weight uses the normalised location of the match's to_ feature
  double sc0, sc1, sc2, sr;
  if (!reversed) { sc0 = sc->cto[0]; sc1 = sc->cto[1]; sc2 = sc->cto[2]; sr = sc->rto; }
  else { sc0 = sc->cfrom[0]; sc1 = sc->cfrom[1]; sc2 = sc->cfrom[2]; sr = sc->rfrom; }
  const double isr = 1.0 / sr;
  const bool window = (flags & LRR_F_BRACKET) != 0;
  double gLo = 0, gHi = 0, sLo = 0, sHi = 0;
  if (window) { if (selG.bracket) { gLo = selG.bracket[0]; gHi = selG.bracket[1]; } if (selS.bracket) { sLo = selS.bracket[0]; sHi = selS.bracket[1]; } }
  double acc[3] = {0, 0, 0};   // keys of G, keys of G below its window, keys of S below its window
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int s = a.sidx[k];
    const double r2n = a.r2[k] * inv_unit2;
    double w = lrr_weight_BT(r2n * inv_scale2) * inv_scale2;
    double gk = -1.0, sk = -1.0;
    const int i = lrr_from_index(a, k);
    if ((flags & LRR_F_ICP) && (flags & LRR_F_STRENGTH)) {
      // cdcl_estimation_ICP.txx:607-611,629-637: every match is multiplied by p.strength * q.strength (no cap, zero weights
      // included) and the weights are then divided by the SUM of those products (acc[0]; written to sc->G by k_weight)
      if (s >= 0) { gk = __ldg(a.fs + i) * __ldg(a.to.rec + 8 * (size_t)s + 6); acc[0] += gk; }
    } else if ((flags & LRR_F_STRENGTH) && w != 0.0 && s >= 0) {   // Q13
      const double sq = __ldg(a.to.rec + 8 * (size_t)s + 6);
      const double sp = __ldg(a.fs + i);
      double g;
      if (flags & LRR_F_RRL) g = (double)__fmul_rn((float)sp, (float)sq);   // cdcl_feature_ICP::strength_ is float
      else g = sp * sq;
      if (g > LRR_STRENGTH_CAP) g = LRR_STRENGTH_CAP;
      gk = g;
      if (window) lrr_sel_window_offer(selG, fabs(g), gLo, gHi, acc[1]);
      else if (selG.fine) lrr_sel_add(selG, fabs(g));
      acc[0] += 1.0;
    }
    if (spatial && s >= 0) {
      double qx, qy, qz;
      if (!reversed) { const double4 q = lrr_ld_pos4(a.to.grid.pos4 + s); qx = q.x; qy = q.y; qz = q.z; }
      else lrr_from_loc(a, sc, mapped, i, qx, qy, qz);
      qx = (qx - sc0) * isr; qy = (qy - sc1) * isr; qz = (qz - sc2) * isr;
      sk = 1.0 / (1.0 + 5.0 * (qx * qx + qy * qy + qz * qz));   // reference: sqrt then square (cdcl :984-985)
      if (window) lrr_sel_window_offer(selS, sk, sLo, sHi, acc[2]);
      else if (selS.fine) lrr_sel_add(selS, sk);
    }
    a.w[k] = w; a.gkey[k] = gk; a.skey[k] = sk;
  }
  const bool fin = scope.template reduce<3>(acc, sh.out, sh.tmp);
  if (fin && threadIdx.x == 0) { sc->n_nonzero = sh.out[0]; sc->sel_below[0] = sh.out[1]; sc->sel_below[1] = sh.out[2]; sc->scale = scale; sc->scale_estimated = 1; sc->used_sum_r2 = sc->sum_r2; }
  return fin;
}

// ---- COLLECT: candidates of the two order statistics ----------------------------------------------------
template <class Scope>
__device__ void lrr_phase_collect(const Scope& scope, const DirArgs& a, int flags, const SelDev& selG, const SelDev& selS) {
  const int binG = (flags & LRR_F_STRENGTH) ? *selG.bin : -1;
  const int binS = LRR_SPATIAL(flags) ? *selS.bin : -1;
  if (binG < 0 && binS < 0) return;
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    if (binG >= 0) { const double g = a.gkey[k]; if (g >= 0.0) lrr_sel_offer(selG, binG, fabs(g)); }
    if (binS >= 0) { const double s = a.skey[k]; if (s >= 0.0) lrr_sel_offer(selS, binS, s); }
  }
}

// ---- ACCUMULATE: second half of the re-weighting + the weighted normal equations
// (cdcl :1029-1069, cdcl sym :551-628, rrl :803-835, rrl sym ThreadedGenerateDataForward/Backward :441-655).
// Four lanes share one match; lane j owns row x~_j of the Kronecker moments
//   T[j][b][p] = sum w x~_j x~_b P_p   (b = 0..3, p = xx,xy,xz,yy,yz,zz)      V[j][d] = sum w x~_j (P b^)_d
// from which J^T W J / J^T W b of ANY transform whose Jacobian is affine in x (affine, small-angle rigid) is
// assembled once at the end — 27 fp64 accumulators per lane instead of 90 per thread.
template <class Scope, class SH>
__device__ bool lrr_phase_accumulate(const Scope& scope, const DirArgs& a, int flags, int reversed, int keep_w, double G, double S,
                                     const double* jc /* centre subtracted inside the Jacobian */, SH& sh) {
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  const bool sym = (flags & LRR_F_SYM) != 0, strength = (flags & LRR_F_STRENGTH) != 0, spatial = LRR_SPATIAL(flags);
  const bool skip_small = sym && (flags & LRR_F_RRL);
  // normalised coordinates (x - c) / r as a multiplication by 1/r: one rounding apart from the reference's
  // division, far inside the 1e-5 gate, and it removes six fp64 divisions (~40 instructions each) per match.
  // The thirteen uniform scalars of the loop live in shared memory (broadcast reads): as registers they pushed the 27 accumulators
  // of every lane into local memory (ncu: 21 MB of spill traffic per launch next to 26 MB of algorithmic bytes).
  __shared__ double s_u[13];
  __syncthreads();
  if (threadIdx.x == 0) {
    s_u[0] = sc->cfrom[0]; s_u[1] = sc->cfrom[1]; s_u[2] = sc->cfrom[2]; s_u[3] = 1.0 / sc->rfrom;
    s_u[4] = sc->cto[0]; s_u[5] = sc->cto[1]; s_u[6] = sc->cto[2]; s_u[7] = 1.0 / sc->rto;
    s_u[8] = 1.0 / G; s_u[9] = 1.0 / S; s_u[10] = jc[0]; s_u[11] = jc[1]; s_u[12] = jc[2];
  }
  __syncthreads();
#define cf0 s_u[0]
#define cf1 s_u[1]
#define cf2 s_u[2]
#define irf s_u[3]
#define ct0 s_u[4]
#define ct1 s_u[5]
#define ct2 s_u[6]
#define irt s_u[7]
#define invG s_u[8]
#define invS s_u[9]
  const int lane = threadIdx.x & 31, j = lane & 3, qbase = lane & ~3;
  double acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.0;
  // Every LANE loads one match per trip (consecutive lanes = consecutive k: coalesced streams, one gather hop behind sidx, which
  // is fetched a trip ahead), finishes its weight and forms (w, x~, P, P b^); the four lanes of a quad then pass their four
  // matches round by shuffle and each adds row x~_j of all four.  Against "the four lanes of a quad load the same match":
  // four times the matches in flight per warp for the same registers, a quarter of the dependent trips.
  const int wfirst = scope.first() - lane, wstride = scope.stride();
  int s_next = (wfirst + lane) < a.M ? a.sidx[wfirst + lane] : -1;
  for (int base = wfirst; base < a.M; base += wstride) {
    const int k = base + lane;
    const int s = s_next;
    s_next = (k + wstride) < a.M ? a.sidx[k + wstride] : -1;
    double w = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0, P[6] = {0, 0, 0, 0, 0, 0}, pb0 = 0.0, pb1 = 0.0, pb2 = 0.0;
    if (k < a.M) {
      w = a.w[k];                                    // Beaton-Tukey weight / scale^2
      if (strength) {
        const double g = a.gkey[k];
        if (g >= 0.0 || (flags & LRR_F_ICP)) w *= g; // only where w != 0   (cdcl :839-848); cdcl_estimation_ICP: every match
        w *= invG;                                   // every match         (cdcl :857-860, Q13)
      }
      if (spatial) { w *= a.skey[k]; w *= invS; }    // cdcl :985-1002
      // cdcl symmetric: the re-weighted reversed constraints are clones that get erased again (Q11), the
      // backward originals keep their Beaton-Tukey weights
      if (!keep_w) a.w[k] = w;
      if (s < 0 || (skip_small && w < 1e-4)) w = 0.0;   // contributes nothing (rrl sym :463,:571)
      else {
        const int i = lrr_from_index(a, k);
        double px, py, pz; lrr_from_loc(a, sc, mapped, i, px, py, pz);
        px = (px - cf0) * irf; py = (py - cf1) * irf; pz = (pz - cf2) * irf;
        const double4 q4 = lrr_ld_pos4(a.to.grid.pos4 + s);
        const double qx = (q4.x - ct0) * irt, qy = (q4.y - ct1) * irt, qz = (q4.z - ct2) * irt;
        double b0, b1, b2;
        if (!reversed) {
          lrr_ld_P6(a.to.rec, s, P);
          x0 = px - s_u[10]; x1 = py - s_u[11]; x2 = pz - s_u[12]; b0 = qx; b1 = qy; b2 = qz;
        } else {   // J at the to_ location, projector and right-hand side of the from_ feature
#pragma unroll
          for (int p = 0; p < 6; ++p) P[p] = __ldg(a.fP6 + (size_t)p * a.n_in + i);
          x0 = qx - s_u[10]; x1 = qy - s_u[11]; x2 = qz - s_u[12]; b0 = px; b1 = py; b2 = pz;
        }
        pb0 = P[0] * b0 + P[1] * b1 + P[2] * b2;
        pb1 = P[1] * b0 + P[3] * b1 + P[4] * b2;
        pb2 = P[2] * b0 + P[4] * b1 + P[5] * b2;
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int src = qbase + m;
      const double wm = __shfl_sync(0xffffffffu, w, src);
      const double xm[4] = {__shfl_sync(0xffffffffu, x0, src), __shfl_sync(0xffffffffu, x1, src), __shfl_sync(0xffffffffu, x2, src), 1.0};
      double Pm[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) Pm[p] = __shfl_sync(0xffffffffu, P[p], src);
      const double c0 = __shfl_sync(0xffffffffu, pb0, src), c1 = __shfl_sync(0xffffffffu, pb1, src), c2 = __shfl_sync(0xffffffffu, pb2, src);
      const double wxa = wm * (j == 0 ? xm[0] : (j == 1 ? xm[1] : (j == 2 ? xm[2] : 1.0)));
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double tt = wxa * xm[b];
#pragma unroll
        for (int p = 0; p < 6; ++p) acc[b * 6 + p] += tt * Pm[p];
      }
      acc[24] += wxa * c0; acc[25] += wxa * c1; acc[26] += wxa * c2;
    }
  }
#undef cf0
#undef cf1
#undef cf2
#undef irf
#undef ct0
#undef ct1
#undef ct2
#undef irt
#undef invG
#undef invS
  lrr_tl(2, 1);
  const bool fin = scope.template reduce<27, 4>(acc, sh.out, sh.tmp);
  lrr_tl(2, 2);
  if (fin) {
    for (int i = threadIdx.x; i < 108; i += blockDim.x) sc->mom[i] = sh.out[i];
    __syncthreads();
  }
  return fin;
}

// ---- SOLVE: mirror / invert / set_parameterization / unnormalize / recompose / inverse
// (cdcl :1071-1089, cdcl sym :630-695, rrl sym :867-912; cdcl_trans_affine.txx:220-286,
//  cdcl_trans_rigid3d.cxx:140-166,260-326).  Run by ONE block; F = source direction, B = other or null. ----
__device__ __forceinline__ int lrr_sym6(int d, int e) {   // index of P_de in (xx,xy,xz,yy,yz,zz)
  const int lo = d < e ? d : e, hi = d < e ? e : d;
  return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
}

// Gauss-Jordan inverse of an SPD n x n matrix in shared memory (n <= 12); n*n threads cooperate.
// Returns false when a pivot is not safely positive (caller falls back to the Jacobi-SVD pseudo-inverse,
// the reference's vnl_svd::inverse semantics).
static __device__ bool lrr_gj_inverse(int n, double* Msh /* n*n, destroyed */, double* inv /* n*n out */) {
  __shared__ int s_ok;
  __shared__ double s_piv;
  const int tid = threadIdx.x, r = tid / n, cidx = tid % n;
  const bool act = tid < n * n;
  double maxd = 0;
  if (tid == 0) { for (int i = 0; i < n; ++i) maxd = fmax(maxd, Msh[i * n + i]); s_ok = 1; s_piv = maxd; }
  if (act) inv[tid] = (r == cidx) ? 1.0 : 0.0;
  __syncthreads();
  maxd = s_piv;
  for (int p = 0; p < n; ++p) {
    const double piv = Msh[p * n + p];
    // (every thread reads the same pivot after the barrier that ended the previous step: the decision is uniform, no flag round trip)
    if (!(piv > 1e-13 * maxd)) { s_ok = 0; break; }
    // one reciprocal per pivot instead of three divisions per thread (~40 dependent instructions each) on the critical path
    // of every iteration; a rounding apart from the quotient, far inside the 1e-5 gate
    const double ipv = 1.0 / piv;
    double f = 0, mp = 0, ip = 0;
    if (act) { f = Msh[r * n + p] * ipv; mp = Msh[p * n + cidx]; ip = inv[p * n + cidx]; }
    __syncthreads();
    if (act) {
      if (r == p) { Msh[tid] = mp * ipv; inv[tid] = ip * ipv; }
      else { Msh[tid] -= f * mp; inv[tid] -= f * ip; }
    }
    __syncthreads();
  }
  const bool ok = s_ok != 0;
  __syncthreads();
  return ok;
}

static __device__ void lrr_solve(Scal* F, Scal* B, int flags, int dof, int estimate_parameters) {
  __shared__ double T[4][4][6], V[4][3];
  __shared__ double AtA[144], Atb[12], Msh[144], inv[144], theta[12];
  const int tid = threadIdx.x;
  const bool rigid = (flags & LRR_F_RIGID) != 0, normalize = (flags & LRR_F_NORMALIZE) != 0, sym = (flags & LRR_F_SYM) != 0;
  const bool sim2d = (flags & LRR_F_SIM2D) != 0;
  for (int i = tid; i < 108; i += blockDim.x) {
    double v = __ldcg(&F->mom[i]);
    if (sym && B) v += __ldcg(&B->mom[i]);
    const int a = i / 27, r = i % 27;
    if (r < 24) T[a][r / 6][r % 6] = v; else V[a][r - 24] = v;
  }
  // The scalars the serial part below works on are staged in shared memory by 46 different threads (one L2 round trip)
  // instead of being read and written one by one by thread 0 (~40 dependent round trips on the critical path of every
  // iteration): sF/sB = A[9] t[3] c[3], sFg = Ag[9] tg[3] cg[3], sFn/sBn = cfrom[3] rfrom cto[3] rto.
  __shared__ double sF[15], sFg[15], sFn[8], sB[15], sBn[8];
  // (the global transform an increment is recomposed onto is the LAST map the clones went through)
  if (tid < 15) { sF[tid] = __ldcg(&F->A[0] + tid); sFg[tid] = F->from_mapped == 2 ? __ldcg(&F->Ag2[0] + tid) : __ldcg(&F->Ag[0] + tid); if (B) sB[tid] = __ldcg(&B->A[0] + tid); }
  else if (tid >= 32 && tid < 40) { sFn[tid - 32] = __ldcg(&F->cfrom[0] + (tid - 32)); if (B) sBn[tid - 32] = __ldcg(&B->cfrom[0] + (tid - 32)); }
  __syncthreads();
  lrr_tl(9, 1);
  if (sim2d) {
    // similarity2d: J = [x -y 1 0; y x 0 1] (cdcl_trans_similarity2d.cxx:150-164), b = q.  J[d][i] = sum_a x~_a E[a][d][i] with
    // E[0][0][0] = E[1][1][0] = 1, E[1][0][1] = -1, E[0][1][1] = 1, E[3][0][2] = E[3][1][3] = 1
    if (tid < 16 + 4) {
      auto Jc = [](int a, int d, int i) -> double {
        if (d > 1) return 0.0;
        if (i == 0) return (a == d) ? 1.0 : 0.0;
        if (i == 1) return (d == 0 && a == 1) ? -1.0 : ((d == 1 && a == 0) ? 1.0 : 0.0);
        return (a == 3 && d == i - 2) ? 1.0 : 0.0;
      };
      if (tid < 16) {
        const int i = tid / 4, jx = tid % 4;
        double s = 0;
        for (int a = 0; a < 4; ++a) for (int d = 0; d < 2; ++d) {
          const double ca = Jc(a, d, i); if (ca == 0.0) continue;
          for (int b = 0; b < 4; ++b) for (int e = 0; e < 2; ++e) {
            const double cb = Jc(b, e, jx); if (cb == 0.0) continue;
            s += ca * cb * T[a][b][lrr_sym6(d, e)];
          }
        }
        AtA[i * 4 + jx] = s;
      } else {
        const int i = tid - 16;
        double s = 0;
        for (int a = 0; a < 4; ++a) for (int d = 0; d < 2; ++d) { const double ca = Jc(a, d, i); if (ca != 0.0) s += ca * V[a][d]; }
        Atb[i] = s;
      }
    }
  } else if (!rigid) {
    // affine: parameter index of (row d, column a of x~) = a<3 ? 3d+a : 9+d   (cdcl_trans_affine.txx:65-106,141-161)
    for (int e = tid; e < 144; e += blockDim.x) {
      const int i = e / 12, jx = e % 12;
      const int di = i < 9 ? i / 3 : i - 9, ai = i < 9 ? i % 3 : 3;
      const int dj = jx < 9 ? jx / 3 : jx - 9, aj = jx < 9 ? jx % 3 : 3;
      AtA[e] = T[ai][aj][lrr_sym6(di, dj)];
    }
    if (tid < 12) { const int di = tid < 9 ? tid / 3 : tid - 9, ai = tid < 9 ? tid % 3 : 3; Atb[tid] = V[ai][di]; }
  } else {
    // small-angle rigid: J = [ -[x]x | I ]  (cdcl_trans_rigid3d.cxx:197-225), b = q - p  (cdcl :1051-1054)
    // J[d][i] = sum_a x~_a E[a][d][i]; non-zero E: rotation columns i<3 from the cross-product matrix, E[3][d][3+d] = 1
    if (tid < 36 + 6) {
      // coefficient tensor on the fly: Jc(a,d,i)
      auto Jc = [](int a, int d, int i) -> double {
        if (i >= 3) return (a == 3 && d == i - 3) ? 1.0 : 0.0;
        if (a == 3 || d == i) return 0.0;
        // column i of -[x]x: col0 = (0,-z,y), col1 = (z,0,-x), col2 = (-y,x,0)
        // entry (d,i) = eps-like sign * x_a with a = 3 - d - i
        if (a != 3 - d - i) return 0.0;
        // (d,i): (1,0)=-z (2,0)=+y (0,1)=+z (2,1)=-x (0,2)=-y (1,2)=+x
        return ((d == 2 && i == 0) || (d == 0 && i == 1) || (d == 1 && i == 2)) ? 1.0 : -1.0;
      };
      if (tid < 36) {
        const int i = tid / 6, jx = tid % 6;
        double s = 0;
        for (int a = 0; a < 4; ++a) for (int d = 0; d < 3; ++d) {
          const double ca = Jc(a, d, i); if (ca == 0.0) continue;
          for (int b = 0; b < 4; ++b) for (int e = 0; e < 3; ++e) {
            const double cb = Jc(b, e, jx); if (cb == 0.0) continue;
            s += ca * cb * T[a][b][lrr_sym6(d, e)];
          }
        }
        AtA[i * 6 + jx] = s;
      } else {
        const int i = tid - 36;
        double s = 0;
        for (int a = 0; a < 4; ++a) for (int d = 0; d < 3; ++d) {
          const double ca = Jc(a, d, i); if (ca == 0.0) continue;
          double v = V[a][d];
          for (int b = 0; b < 3; ++b) v -= T[a][b][lrr_sym6(d, b)];   // (P x^)_d
          s += ca * v;
        }
        Atb[i] = s;
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < dof * dof; e += blockDim.x) { F->AtA[e] = AtA[e]; Msh[e] = AtA[e]; }
  if (tid < dof) F->Atb[tid] = Atb[tid];
  __syncthreads();
  lrr_tl(9, 2);
  bool ok = lrr_gj_inverse(dof, Msh, inv);
  lrr_tl(9, 3);
  if (!ok) {   // singular or indefinite normal matrix: pseudo-inverse like vnl_svd::inverse (slow path, one thread)
    if (tid == 0) { lrr_svd_inverse<12>(dof, AtA, inv); F->status |= LRR_ST_SINGULAR; }
    __syncthreads();
  }
  if (tid < dof) { double s = 0; for (int k = 0; k < dof; ++k) s += inv[tid * dof + k] * Atb[k]; theta[tid] = s; }
  __syncthreads();
  const double rm = sFn[3], rfx = sFn[7];
  // covariance: set_covariance(invAA), then affine unnormalize scales it by B = diag((rf/rm) I9, rf I3)
  for (int e = tid; e < dof * dof; e += blockDim.x) {
    double v = inv[e];
    if (sim2d && normalize) {   // similarity2d unnormalize: B = diag(rf/rm, rf/rm, rf, rf)  (similarity2d.cxx:293-304)
      const int i = e / 4, jx = e % 4;
      v *= ((i < 2) ? rfx / rm : rfx) * ((jx < 2) ? rfx / rm : rfx);
    } else if (!rigid && normalize) {
      const int i = e / 12, jx = e % 12;
      v *= ((i < 9) ? rfx / rm : rfx) * ((jx < 9) ? rfx / rm : rfx);
    }
    F->cov[e] = v;
  }
  __syncthreads();
  lrr_tl(10, 0);
  if (tid == 0) {
    double A[9], t[3], c[3];
    const double cm[3] = {sFn[0], sFn[1], sFn[2]}, cf[3] = {sFn[4 + 0], sFn[4 + 1], sFn[4 + 2]};
    if (sim2d) {
      if (estimate_parameters) {
        // set_parameterization (similarity2d.cxx:93-110): s = |(p0,p1)|, R from (p0,p1)/s, ang = atan2; then unnormalize (:271-291)
        // rebuilds s R from cos / sin of that angle
        double s_ = sqrt(theta[0] * theta[0] + theta[1] * theta[1]);
        double ca = theta[0] / s_, sa = theta[1] / s_;
        t[0] = theta[2]; t[1] = theta[3]; t[2] = 0.0;
        if (normalize) {
          const double ang = atan2(sa, ca);
          s_ = rfx / rm * s_;
          const double c2 = cos(ang), s2 = sin(ang);
          const double sR[4] = {s_ * c2, -s_ * s2, s_ * s2, s_ * c2};
          for (int r = 0; r < 2; ++r) {
            double acc = (-sR[2 * r]) * cm[0]; acc += (-sR[2 * r + 1]) * cm[1];
            t[r] = acc + rfx * t[r] + cf[r];
          }
        }
        for (int i = 0; i < 9; ++i) A[i] = 0.0;
        A[0] = s_ * ca; A[1] = s_ * -sa; A[3] = s_ * sa; A[4] = s_ * ca; A[8] = 1.0;
        for (int i = 0; i < 9; ++i) sF[i] = A[i];
        for (int d = 0; d < 3; ++d) { sF[9 + d] = t[d]; sF[12 + d] = 0.0; }
      }
    } else if (!rigid) {
      if (estimate_parameters) {
        for (int i = 0; i < 9; ++i) A[i] = theta[i];
        for (int d = 0; d < 3; ++d) t[d] = theta[9 + d];
        for (int d = 0; d < 3; ++d) c[d] = normalize ? 0.0 : sF[12 + d];
      } else {
        for (int i = 0; i < 9; ++i) A[i] = sF[i];
        for (int d = 0; d < 3; ++d) { t[d] = sF[9 + d]; c[d] = sF[12 + d]; }
        if (normalize) {   // cdcl_trans_affine::normalize  (affine.txx:220-253)
          const double s = rm / rfx;
          for (int i = 0; i < 9; ++i) A[i] = s * A[i];
          const double dc[3] = {cm[0] - c[0], cm[1] - c[1], cm[2] - c[2]};
          for (int r = 0; r < 3; ++r) {
            double acc = A[3 * r] * dc[0]; acc += A[3 * r + 1] * dc[1]; acc += A[3 * r + 2] * dc[2];
            t[r] = t[r] / rfx + acc / rm - cf[r] / rfx;
          }
          c[0] = c[1] = c[2] = 0.0;
        }
      }
      if (normalize) {   // cdcl_trans_affine::unnormalize  (affine.txx:259-286)
        const double s = rfx / rm;
        for (int i = 0; i < 9; ++i) A[i] = s * A[i];
        for (int r = 0; r < 3; ++r) t[r] = rfx * t[r] + cf[r];
        for (int d = 0; d < 3; ++d) c[d] = cm[d];
      }
      for (int i = 0; i < 9; ++i) sF[i] = A[i];
      for (int d = 0; d < 3; ++d) { sF[9 + d] = t[d]; sF[12 + d] = c[d]; }
    } else if (estimate_parameters) {
      // increment from the small-angle parameterisation (rigid3d.cxx:140-166)
      const double da = theta[0], db = theta[1], dg = theta[2];
      A[0] = 1.0; A[1] = -dg; A[2] = db; A[3] = dg; A[4] = 1.0; A[5] = -da; A[6] = -db; A[7] = da; A[8] = 1.0;
      t[0] = theta[3]; t[1] = theta[4]; t[2] = theta[5];
      if (normalize) {   // cdcl_trans_rigid3d::unnormalize  (rigid3d.cxx:285-305)
        const double s = rfx / rm;
        for (int i = 0; i < 9; ++i) A[i] = s * A[i];
        for (int r = 0; r < 3; ++r) {
          double acc = (-A[3 * r]) * cm[0]; acc += (-A[3 * r + 1]) * cm[1]; acc += (-A[3 * r + 2]) * cm[2];
          t[r] = acc + rfx * t[r] + cf[r];
        }
      }
      if (B && (flags & LRR_F_INV_OF_INCREMENT)) {   // the other direction <- inverse() of the increment as it stands here (cdcl sym :849-851)
        double invA[9]; lrr_inv3(A, invA);
        for (int i = 0; i < 9; ++i) sB[i] = invA[i];
        for (int i = 0; i < 3; ++i) {
          double acc = (-invA[3 * i]) * t[0]; acc += (-invA[3 * i + 1]) * t[1]; acc += (-invA[3 * i + 2]) * t[2];
          sB[9 + i] = acc + (normalize ? cm[i] : 0.0); sB[12 + i] = 0.0;   // (the increment's centre: unnormalize left center_moving_ = the normalisation centre, rigid3d.cxx:305)
        }
      }
      // recompose_increment (rigid3d.cxx:309-326): dR = U V^T, R <- R dR, t <- dR t + dt   (Q7)
      double dR[9]; lrr_orthogonalize3(A, dR);
      double Rn[9], tn[3];
      for (int i = 0; i < 3; ++i) for (int jx = 0; jx < 3; ++jx) {
        double acc = sFg[3 * i] * dR[jx]; acc += sFg[3 * i + 1] * dR[3 + jx]; acc += sFg[3 * i + 2] * dR[6 + jx];
        Rn[3 * i + jx] = acc;
      }
      for (int i = 0; i < 3; ++i) {
        double acc = dR[3 * i] * sFg[9 + 0]; acc += dR[3 * i + 1] * sFg[9 + 1]; acc += dR[3 * i + 2] * sFg[9 + 2];
        tn[i] = acc + t[i];
      }
      for (int i = 0; i < 9; ++i) sF[i] = Rn[i];
      for (int d = 0; d < 3; ++d) { sF[9 + d] = tn[d]; sF[12 + d] = sFg[12 + d]; }
    }
    if (B && !estimate_parameters && normalize && !rigid && !(flags & LRR_F_NO_OTHER_RENORM)) {
      // the other direction's transform only goes through normalize() / unnormalize() (cdcl sym :470-478,:668-675):
      // same map, but its centre moves to the centre of its own from-features
      const double brm = sBn[3], brf = sBn[7];
      const double bcm[3] = {sBn[0], sBn[1], sBn[2]}, bcf[3] = {sBn[4 + 0], sBn[4 + 1], sBn[4 + 2]};
      double A2[9], t2[3];
      const double s1 = brm / brf;
      for (int i = 0; i < 9; ++i) A2[i] = s1 * sB[i];
      const double dc[3] = {bcm[0] - sB[12 + 0], bcm[1] - sB[12 + 1], bcm[2] - sB[12 + 2]};
      for (int r = 0; r < 3; ++r) {
        double acc = A2[3 * r] * dc[0]; acc += A2[3 * r + 1] * dc[1]; acc += A2[3 * r + 2] * dc[2];
        t2[r] = sB[9 + r] / brf + acc / brm - bcf[r] / brf;
      }
      const double s2 = brf / brm;
      for (int i = 0; i < 9; ++i) sB[i] = s2 * A2[i];
      for (int r = 0; r < 3; ++r) { sB[9 + r] = brf * t2[r] + bcf[r]; }
      // covariance: B2 (B1 Sigma B1^T) B2 with B1 = diag(s1 I9, 1/rf I3) + N/rf, B2 = diag(s2 I9, rf I3); N[9+d][3d+k] = dc[k]
      double K[144], Tm[144];
      for (int i = 0; i < 144; ++i) K[i] = 0.0;
      for (int i = 0; i < 9; ++i) K[i * 12 + i] = s2 * s1;
      for (int d = 0; d < 3; ++d) { K[(9 + d) * 12 + 9 + d] = brf * (1.0 / brf); for (int k = 0; k < 3; ++k) K[(9 + d) * 12 + 3 * d + k] = brf * (dc[k] / brf); }
      for (int i = 0; i < 12; ++i) for (int jx = 0; jx < 12; ++jx) { double v = 0; for (int k = 0; k < 12; ++k) v += K[i * 12 + k] * B->cov[k * 12 + jx]; Tm[i * 12 + jx] = v; }
      for (int i = 0; i < 12; ++i) for (int jx = 0; jx < 12; ++jx) { double v = 0; for (int k = 0; k < 12; ++k) v += Tm[i * 12 + k] * K[jx * 12 + k]; B->cov[i * 12 + jx] = v; }
      for (int d = 0; d < 3; ++d) sB[12 + d] = bcm[d];
    }
    if (B && estimate_parameters && !(rigid && (flags & LRR_F_INV_OF_INCREMENT))) {   // backward = forward->inverse()  (affine.txx:54-59, rigid3d.cxx:113-118)
      double invA[9]; lrr_inv3(sF, invA);
      for (int i = 0; i < 9; ++i) sB[i] = invA[i];
      for (int i = 0; i < 3; ++i) {
        double acc = (-invA[3 * i]) * sF[9 + 0]; acc += (-invA[3 * i + 1]) * sF[9 + 1]; acc += (-invA[3 * i + 2]) * sF[9 + 2];
        sB[9 + i] = acc + sF[12 + i]; sB[12 + i] = 0.0;
      }
    }
  }
  __syncthreads();
  lrr_tl(10, 1);
  if (tid < 15) { (&F->A[0])[tid] = sF[tid]; if (B) (&B->A[0])[tid] = sB[tid]; }
  if (B && estimate_parameters) for (int e = tid; e < dof * dof; e += blockDim.x) B->cov[e] = 0.0;   // inverse(): fresh covariance
  __syncthreads();
}

// ---- RESID: residuals of the current matches under the current transform: RMS_error (cdcl :196-230,
// rrl :211-241 Q1), weighted_error (cdcl :236-268, rrl :247-276), scale-from-weights sums (cdcl :698-701) ----
template <class Scope, class SH>
__device__ bool lrr_phase_resid(const Scope& scope, const DirArgs& a, SH& sh) {
  Scal* sc = a.sc;
  const int mapped = sc->from_mapped;
  double A[9], t[3], c[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = sc->A[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { t[i] = sc->t[i]; c[i] = sc->c[i]; }
  double acc[5] = {0, 0, 0, 0, 0};
  for (int k = scope.first(); k < a.M; k += scope.stride()) {
    const int s = a.sidx[k];
    if (s < 0) continue;
    const int i = lrr_from_index(a, k);
    double x, y, z; lrr_from_loc(a, sc, mapped, i, x, y, z);
    double yx, yy, yz; lrr_map_exact(A, t, c, x, y, z, yx, yy, yz);
    const double4 q = lrr_ld_pos4(a.to.grid.pos4 + s);
    double P[6]; lrr_ld_P6(a.to.rec, s, P);
    const double r2 = lrr_proj_r2(yx - q.x, yy - q.y, yz - q.z, P[0], P[1], P[2], P[3], P[4], P[5]);
    const double w = a.w[k];
    a.r2[k] = r2;
    acc[0] += r2; acc[1] += w * r2; acc[2] += w; acc[3] += sqrt(r2) * w; acc[4] += 1.0;
  }
  lrr_tl(8, 1);
  const bool fin = scope.template reduce<5>(acc, sh.out, sh.tmp);
  if (fin) {
    if (threadIdx.x == 0) {
      sc->res_sum_r2 = sh.out[0]; sc->res_sum_wr2 = sh.out[1]; sc->res_sum_w = sh.out[2]; sc->res_sum_wr = sh.out[3];
      sc->res_count = sh.out[4];
    }
  }
  scope.phase_end();
  return fin;
}
