"""ctypes binding of the C ABI declared in include/lrr_gpu.h.

The binding is generic over (shared library, symbol prefix): the product loads
``lrr_b200/liblrr_b200.so`` with prefix ``lrr_``; the parity tests instantiate the same class over the
CPU checker library (same ABI, its own prefix) so that one script drives both.  Nothing in this package
imports or loads that checker — only tests/, __graft_entry__.smoke() and bench.py do.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

LRR_OK = 0
KIND = {"cdcl_fwd": 0, "cdcl_sym": 1, "rrl_fwd": 2, "rrl_sym": 3, "cdcl_icp": 4}
TRANS_AFFINE, TRANS_RIGID3D, TRANS_SIMILARITY2D = 0, 1, 2
DIR_FORWARD, DIR_BACKWARD = 0, 1
ORDER_FIXED, ORDER_RESHUFFLE = 0, 1
METRIC = {"rms": 0, "weighted": 1, "median": 2, "sheet_angles": 3, "tube_angles": 4}

# every symbol include/lrr_gpu.h declares (checked by tests/test_abi.py against the header text)
ABI_SYMBOLS = [
    "ctx_create", "ctx_destroy", "ctx_set_stream", "ctx_synchronize", "ctx_last_error", "ctx_launch_count",
    "ctx_set_profiling", "ctx_phase_times", "ctx_set_async_upload", "ctx_set_option",
    "features_upload", "features_upload_cov", "features_destroy", "features_size", "features_nearest", "features_set_labels", "features_in_regions",
    "est_create", "est_destroy", "est_set_subsets", "est_set_transform", "est_set_flags", "est_set_order",
    "est_set_shuffle", "est_set_initial_scale", "est_set_number_points", "est_find_closest", "est_one_iteration",
    "est_run", "est_get_transform", "est_get_matches", "est_metric", "est_estimate_ls_backward", "est_transfer_cov",
    "est_get_state", "est_get_stats", "est_oscillated", "objcov_eval", "objcov_refine", "objpar_eval", "objpar_refine",
    "cdc_compute_weights", "cdc_get_matches", "cdc_resolution_switch", "cdc_one_iteration", "batch_run", "batch_run_ids", "batch_run_multi", "batch_run_locations", "batch_expected_cost",
]


class EstState(C.Structure):
    _fields_ = [
        ("number_points", C.c_int64 * 2), ("number_matches", C.c_int64), ("finest_level", C.c_int32),
        ("oscillation_count", C.c_int32), ("error_difference", C.c_double), ("weighted_error", C.c_double),
        ("scale", C.c_double * 2), ("scale_estimated", C.c_int32 * 2), ("iterations_done", C.c_int32),
        ("norm_center_moving", C.c_double * 3), ("norm_center_fixed", C.c_double * 3),
        ("norm_radius_moving", C.c_double), ("norm_radius_fixed", C.c_double),
    ]


class EstStats(C.Structure):
    _fields_ = [("cert_failed", C.c_int64 * 2), ("passb_queries", C.c_int64 * 2), ("work_list_entries", C.c_int64 * 2),
                ("window_misses", C.c_int64), ("spec_misses", C.c_int64)]


class CdcSummary(C.Structure):
    _fields_ = [("rows", C.c_int64), ("entries", C.c_int64), ("min_matches", C.c_int32), ("max_matches", C.c_int32),
                ("avg_largest_eval", C.c_double)]


class Job(C.Structure):
    _fields_ = [("A", C.c_double * 9), ("t", C.c_double * 3), ("c", C.c_double * 3),
                ("roi_center_moving", C.c_double * 3), ("roi_center_fixed", C.c_double * 3),
                ("roi_half_size", C.c_double)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("converged", C.c_int32), ("oscillated", C.c_int32), ("iterations", C.c_int32),
                ("n_moving_inside", C.c_int32), ("n_fixed_inside", C.c_int32),
                ("A", C.c_double * 9), ("t", C.c_double * 3), ("c", C.c_double * 3), ("cov", C.c_double * 144),
                ("rms", C.c_double), ("weighted_error", C.c_double)]


JOB_DTYPE = np.dtype([("A", "f8", 9), ("t", "f8", 3), ("c", "f8", 3), ("roi_center_moving", "f8", 3),
                      ("roi_center_fixed", "f8", 3), ("roi_half_size", "f8")])
RESULT_DTYPE = np.dtype([("status", "i4"), ("converged", "i4"), ("oscillated", "i4"), ("iterations", "i4"),
                         ("n_moving_inside", "i4"), ("n_fixed_inside", "i4"), ("A", "f8", 9), ("t", "f8", 3),
                         ("c", "f8", 3), ("cov", "f8", 144), ("rms", "f8"), ("weighted_error", "f8")])
DECISION_DTYPE = np.dtype([("rank_angles", "f8", 4), ("rank_angle_average", "f8"), ("measurements", "f8", 8), ("weighted_error", "f8", 2),
                           ("max_trace", "f8", 2), ("sheet_angles", "f8", 2), ("tube_angles", "f8", 2), ("rank", "i4"), ("ran", "i4"),
                           ("verified", "i4"), ("measured", "i4")])
assert JOB_DTYPE.itemsize == C.sizeof(Job) and RESULT_DTYPE.itemsize == C.sizeof(Result) and DECISION_DTYPE.itemsize == 8 * 21 + 16


class LrrError(RuntimeError):
    pass


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


class Lib:
    """One loaded shared library exposing the lrr_gpu.h ABI under `prefix`."""

    def __init__(self, path: str, prefix: str = "lrr_"):
        if not os.path.exists(path):
            raise LrrError(f"shared library not found: {path} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        self._declare()

    def fn(self, name):
        return getattr(self.dll, self.prefix + name)

    def has(self, name) -> bool:
        return hasattr(self.dll, self.prefix + name)

    def _declare(self):
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
        pd, pi32 = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        sig = {
            "ctx_create": (i32, [i32, C.POINTER(vp)]), "ctx_destroy": (None, [vp]),
            "ctx_set_stream": (i32, [vp, vp]), "ctx_synchronize": (i32, [vp]), "ctx_set_async_upload": (i32, [vp, i32]), "ctx_set_option": (i32, [vp, C.c_char_p, i32]),
            "ctx_last_error": (C.c_char_p, [vp]), "ctx_launch_count": (i64, [vp]),
            "ctx_set_profiling": (i32, [vp, i32]), "ctx_phase_times": (i32, [vp, i32, pd, C.POINTER(C.c_int64)]),
            "features_upload": (i32, [vp, i64, pd, pd, pd, pd, C.POINTER(C.c_uint8), pd, C.POINTER(vp)]),
            "features_upload_cov": (i32, [vp, i32, i64, pd, pd, pd, C.POINTER(vp)]),
            "features_destroy": (None, [vp]), "features_size": (i64, [vp]),
            "features_nearest": (i32, [vp, i64, pd, pi32]),
            "features_set_labels": (i32, [vp, pi32]), "features_in_regions": (i32, [vp, pd, pd, pd, i64, i64, pi32, C.POINTER(i64)]),
            "est_create": (i32, [vp, i32, i32, i32, vp, vp, C.POINTER(vp)]), "est_destroy": (None, [vp]),
            "est_set_subsets": (i32, [vp, pi32, i64, pi32, i64]),
            "est_set_transform": (i32, [vp, i32, pd, pd, pd, pd]),
            "est_set_flags": (i32, [vp, i32, i32]), "est_set_order": (i32, [vp, i32, pi32, i64]),
            "est_set_shuffle": (i32, [vp, i32, u64]), "est_set_initial_scale": (i32, [vp, i32, dbl]),
            "est_set_number_points": (i32, [vp, i32, i64, C.POINTER(i32)]),
            "est_find_closest": (i32, [vp]), "est_one_iteration": (i32, [vp, C.c_uint, C.POINTER(i32)]),
            "est_run": (i32, [vp, i32, C.POINTER(i32), C.POINTER(i32)]),
            "est_get_transform": (i32, [vp, i32, C.POINTER(i32), pd, pd, pd, pd]),
            "est_get_matches": (i32, [vp, i32, i64, C.POINTER(i64), pi32, pi32, pd]),
            "est_metric": (i32, [vp, i32, i32, pd]), "est_estimate_ls_backward": (i32, [vp, i32]),
            "est_transfer_cov": (i32, [vp, i32, pd]), "est_get_state": (i32, [vp, C.POINTER(EstState)]),
            "est_oscillated": (i32, [vp, C.POINTER(i32)]), "est_get_stats": (i32, [vp, C.POINTER(EstStats)]),
            "objcov_eval": (i32, [vp, i32, pd, pd, pd]),
            "objcov_refine": (i32, [vp, i32, i32, pd, pd, C.POINTER(i32)]),
            "objpar_eval": (i32, [vp, i32, pd, pd, pd]),
            "objpar_refine": (i32, [vp, i32, i32, pd, pd, C.POINTER(i32), C.POINTER(i32)]),
            "cdc_compute_weights": (i32, [vp, i32, i32, dbl, C.POINTER(CdcSummary)]),
            "cdc_get_matches": (i32, [vp, i32, i64, i64, C.POINTER(i64), C.POINTER(i64), pi32, C.POINTER(i64), pi32, pd]),
            "cdc_resolution_switch": (i32, [dbl, pd, i32, i32, i32, C.POINTER(i32)]),
            "cdc_one_iteration": (i32, [vp, i32, C.c_uint, i32, dbl, C.POINTER(i32), C.POINTER(CdcSummary)]),
            "test_cdc_normalisation": (i32, [vp, pd]),
            "batch_run": (i32, [vp, vp, vp, vp, i64, i32, u64, vp]),
            "batch_run_ids": (i32, [vp, vp, vp, vp, C.POINTER(C.c_int64), i64, i32, u64, vp]),
            "batch_run_locations": (i32, [vp, vp, vp, vp, i64, C.c_int32, i32, u64, i32, dbl, vp, vp]),
            "batch_run_multi": (i32, [C.POINTER(vp), i32, C.POINTER(vp), C.POINTER(vp), vp, i64, i32, u64, vp]),
            "batch_expected_cost": (i32, [vp, vp, vp, vp, i64, pi32]),
        }
        for name, (res, args) in sig.items():
            if self.has(name):
                f = self.fn(name)
                f.restype, f.argtypes = res, args

    # ---- object-style helpers ------------------------------------------------------------------
    def context(self, device: int = 0) -> "Context":
        return Context(self, device)

    def batch_run_multi(self, contexts, movings, fixeds, jobs: np.ndarray, max_iter: int = 70, seed: int = 0):
        """The whole job list over several devices of this process (contexts[d], movings[d], fixeds[d] = replicas on device d)."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        n = len(contexts)
        vp = C.c_void_p
        ca = (vp * n)(*[c.h for c in contexts]); ma = (vp * n)(*[m.h for m in movings]); fa = (vp * n)(*[f.h for f in fixeds])
        rc = self.fn("batch_run_multi")(ca, n, ma, fa, jobs.ctypes.data_as(vp), len(jobs), max_iter, seed, out.ctypes.data_as(vp))
        contexts[0].check(rc, "batch_run_multi")
        return out


class Context:
    def __init__(self, lib: Lib, device: int = 0):
        self.lib = lib
        h = C.c_void_p()
        rc = lib.fn("ctx_create")(device, C.byref(h))
        if rc != LRR_OK:
            raise LrrError(f"{lib.prefix}ctx_create(device={device}) failed with status {rc} "
                           "(4 = no CUDA device; this backend has no CPU fallback)")
        self.h = h

    def check(self, rc, what):
        if rc != LRR_OK:
            msg = self.lib.fn("ctx_last_error")(self.h)
            raise LrrError(f"{self.lib.prefix}{what} -> status {rc}: {msg.decode() if msg else ''}")

    def synchronize(self):
        self.check(self.lib.fn("ctx_synchronize")(self.h), "ctx_synchronize")

    def set_option(self, key: str, value: int):
        """behaviour switch of the context (include/lrr_gpu.h: certificates, tile, speculative, pipeline, norm_skip, flat_scan, tile_sparse)"""
        self.check(self.lib.fn("ctx_set_option")(self.h, key.encode(), int(value)), "ctx_set_option")

    def set_async_upload(self, on: bool):
        """the direction arrays of later uploads are copied while the registration already runs (host buffers must stay valid)"""
        self.check(self.lib.fn("ctx_set_async_upload")(self.h, int(on)), "ctx_set_async_upload")

    def set_stream(self, stream_ptr: int):
        self.check(self.lib.fn("ctx_set_stream")(self.h, C.c_void_p(stream_ptr)), "ctx_set_stream")

    def launch_count(self) -> int:
        return int(self.lib.fn("ctx_launch_count")(self.h))

    PHASES = ["match", "norm1", "norm2", "weight", "collect", "accumulate", "resid", "muse", "metric", "objcov", "batch", "restamp", "search", "queue"]

    def set_profiling(self, on: bool):
        self.check(self.lib.fn("ctx_set_profiling")(self.h, int(on)), "ctx_set_profiling")

    def set_profiling_keep(self, on: bool):
        """pause / resume the per-kernel timers without resetting the counters"""
        self.check(self.lib.fn("ctx_set_profiling")(self.h, 2 if on else 0), "ctx_set_profiling")

    def phase_times(self) -> dict:
        n = len(self.PHASES)
        ms = np.zeros(n); cnt = np.zeros(n, dtype=np.int64)
        self.check(self.lib.fn("ctx_phase_times")(self.h, n, _dp(ms), cnt.ctypes.data_as(C.POINTER(C.c_int64))), "ctx_phase_times")
        return {p: (float(ms[i]), int(cnt[i])) for i, p in enumerate(self.PHASES) if cnt[i]}

    def upload(self, feats: dict) -> "FeatureSet":
        return FeatureSet(self, feats)

    def estimator(self, kind: str, moving: "FeatureSet", fixed: "FeatureSet", dof: int = 12, dim: int = 3) -> "Estimator":
        return Estimator(self, kind, moving, fixed, dof, dim)

    def batch_run(self, moving: "FeatureSet", fixed: "FeatureSet", jobs: np.ndarray, max_iter: int = 70, seed: int = 0, ids=None):
        """ids: global indices of `jobs` in a larger list (a shard): job i is seeded with seed + ids[i]."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE)
        if ids is None:
            rc = self.lib.fn("batch_run")(self.h, moving.h, fixed.h, jobs.ctypes.data_as(C.c_void_p), len(jobs), max_iter,
                                          seed, out.ctypes.data_as(C.c_void_p))
        else:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            assert len(ids) == len(jobs)
            rc = self.lib.fn("batch_run_ids")(self.h, moving.h, fixed.h, jobs.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.POINTER(C.c_int64)),
                                              len(jobs), max_iter, seed, out.ctypes.data_as(C.c_void_p))
        self.check(rc, "batch_run")
        return out

    def batch_expected_cost(self, moving: "FeatureSet", fixed: "FeatureSet", jobs: np.ndarray) -> np.ndarray:
        """features covered by the two ROI cubes of every job (0: an ROI cannot pass the 1000-feature gate)"""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        cost = np.zeros(len(jobs), dtype=np.int32)
        rc = self.lib.fn("batch_expected_cost")(self.h, moving.h, fixed.h, jobs.ctypes.data_as(C.c_void_p), len(jobs),
                                                cost.ctypes.data_as(C.POINTER(C.c_int32)))
        self.check(rc, "batch_expected_cost")
        return cost

    def batch_run_locations(self, moving: "FeatureSet", fixed: "FeatureSet", jobs: np.ndarray, n_candidates: int, max_iter: int = 70, seed: int = 0,
                            until_verified: bool = False, voxel_spacing: float = 1.0):
        """The driver's candidate loop: jobs[l * n_candidates + c]; returns (results, decisions), indexed like jobs."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        assert len(jobs) % n_candidates == 0
        out = np.zeros(len(jobs), dtype=RESULT_DTYPE); dec = np.zeros(len(jobs), dtype=DECISION_DTYPE)
        rc = self.lib.fn("batch_run_locations")(self.h, moving.h, fixed.h, jobs.ctypes.data_as(C.c_void_p), len(jobs) // n_candidates, n_candidates,
                                                max_iter, seed, int(until_verified), float(voxel_spacing), out.ctypes.data_as(C.c_void_p),
                                                dec.ctypes.data_as(C.c_void_p))
        self.check(rc, "batch_run_locations")
        return out, dec

    def close(self):
        if self.h:
            self.lib.fn("ctx_destroy")(self.h)
            self.h = None


class FeatureSet:
    """SoA host arrays -> lrr_fset.  feats: dict(xyz(3,n), P6(6,n), C6(6,n)|None, strength(n), shape(n) u8, dirs(9,n)|None),
    or — plain 2-D cdcl_feature sets of the cdcl_icp kind — dict(xy(2,n), C3(3,n) = (xx,xy,yy), strength(n)|None)."""

    def __init__(self, ctx: Context, feats: dict):
        self.ctx = ctx
        if "xy" in feats:
            n = int(feats["xy"].shape[1]); self.n = n
            xy = _f64(feats["xy"], (2, n)); C3 = _f64(feats["C3"], (3, n))
            st = _f64(feats["strength"], (n,)) if feats.get("strength") is not None else None
            h = C.c_void_p()
            ctx.check(ctx.lib.fn("features_upload_cov")(ctx.h, 2, n, _dp(xy), _dp(C3), _dp(st), C.byref(h)), "features_upload_cov")
            self.h = h
            return
        n = int(feats["xyz"].shape[1])
        self.n = n
        xyz = _f64(feats["xyz"], (3, n)); P6 = _f64(feats["P6"], (6, n))
        C6 = _f64(feats.get("C6"), (6, n)) if feats.get("C6") is not None else None
        st = _f64(feats["strength"], (n,))
        shape = np.ascontiguousarray(feats["shape"], dtype=np.uint8)
        dirs = _f64(feats.get("dirs"), (9, n)) if feats.get("dirs") is not None else None
        h = C.c_void_p()
        rc = ctx.lib.fn("features_upload")(ctx.h, n, _dp(xyz), _dp(P6), _dp(C6), _dp(st),
                                           shape.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(dirs), C.byref(h))
        ctx.check(rc, "features_upload")
        self.h = h
        self._dirs_host = dirs   # asynchronous upload (Context.set_async_upload): the copy may still be reading this buffer

    def nearest(self, q_xyz: np.ndarray) -> np.ndarray:
        q = _f64(q_xyz); nq = q.shape[1]
        out = np.empty(nq, dtype=np.int32)
        self.ctx.check(self.ctx.lib.fn("features_nearest")(self.h, nq, _dp(q), out.ctypes.data_as(C.POINTER(C.c_int32))),
                       "features_nearest")
        return out

    def set_labels(self, labels):
        """region label of every feature (the watershed over-segmentation looked up at its location)"""
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        self._labels = lab
        self.ctx.check(self.ctx.lib.fn("features_set_labels")(self.h, lab.ctypes.data_as(C.POINTER(C.c_int32))), "features_set_labels")

    def in_regions(self, x0, x1, reference_point, min_count: int = 1500, cap: int = 8192) -> np.ndarray:
        """FeaturesInRegions for this volume: indices in the reference's order (region after region, nearest first)."""
        out = np.empty(cap, dtype=np.int32); n = C.c_int64(0)
        x0 = _f64(x0, (3,)); x1 = _f64(x1, (3,)); rp = _f64(reference_point, (3,))
        self.ctx.check(self.ctx.lib.fn("features_in_regions")(self.h, _dp(x0), _dp(x1), _dp(rp), min_count, cap, out.ctypes.data_as(C.POINTER(C.c_int32)),
                                                               C.byref(n)), "features_in_regions")
        return out[: min(int(n.value), cap)].copy()

    def close(self):
        if self.h:
            self.ctx.lib.fn("features_destroy")(self.h)
            self.h = None


class Estimator:
    def __init__(self, ctx: Context, kind: str, moving: FeatureSet, fixed: FeatureSet, dof: int = 12, dim: int = 3):
        self.ctx, self.lib, self.kind, self.dof = ctx, ctx.lib, kind, dof
        self.moving, self.fixed = moving, fixed
        h = C.c_void_p()
        rc = self.lib.fn("est_create")(ctx.h, KIND[kind], dim, dof, moving.h, fixed.h, C.byref(h))
        ctx.check(rc, "est_create")
        self.h = h

    def _c(self, rc, what):
        self.ctx.check(rc, what)

    def set_subsets(self, moving_inside, fixed_inside=None):
        mi = np.ascontiguousarray(moving_inside, dtype=np.int32)
        p32 = C.POINTER(C.c_int32)
        if fixed_inside is None:
            self._c(self.lib.fn("est_set_subsets")(self.h, mi.ctypes.data_as(p32), len(mi), None, 0), "est_set_subsets")
        else:
            fi = np.ascontiguousarray(fixed_inside, dtype=np.int32)
            self._c(self.lib.fn("est_set_subsets")(self.h, mi.ctypes.data_as(p32), len(mi), fi.ctypes.data_as(p32), len(fi)),
                    "est_set_subsets")

    def set_transform(self, A, t, c=None, cov=None, kind=None):
        if kind is None:
            kind = TRANS_AFFINE if self.dof == 12 else (TRANS_RIGID3D if self.dof == 6 else TRANS_SIMILARITY2D)
        A = _f64(np.asarray(A).reshape(9)); t = _f64(t, (3,))
        c = _f64(c if c is not None else np.zeros(3), (3,))
        cov = _f64(np.asarray(cov).reshape(-1)) if cov is not None else None
        self._c(self.lib.fn("est_set_transform")(self.h, kind, _dp(A), _dp(t), _dp(c), _dp(cov)), "est_set_transform")

    def set_flags(self, normalize_matches: bool, weight_by_strength: bool):
        self._c(self.lib.fn("est_set_flags")(self.h, int(normalize_matches), int(weight_by_strength)), "est_set_flags")

    def set_order(self, perm, direction=0):
        p = np.ascontiguousarray(perm, dtype=np.int32)
        self._c(self.lib.fn("est_set_order")(self.h, direction, p.ctypes.data_as(C.POINTER(C.c_int32)), len(p)), "est_set_order")

    def set_shuffle(self, mode: int, seed: int = 0):
        self._c(self.lib.fn("est_set_shuffle")(self.h, mode, seed), "est_set_shuffle")

    def set_initial_scale(self, scale: float, direction=0):
        self._c(self.lib.fn("est_set_initial_scale")(self.h, direction, float(scale)), "est_set_initial_scale")

    def set_number_points(self, m: int, direction=0) -> bool:
        at = C.c_int(0)
        self._c(self.lib.fn("est_set_number_points")(self.h, direction, int(m), C.byref(at)), "est_set_number_points")
        return bool(at.value)

    def find_closest(self):
        self._c(self.lib.fn("est_find_closest")(self.h), "est_find_closest")

    def one_iteration(self, iteration: int = 0) -> bool:
        conv = C.c_int(0)
        self._c(self.lib.fn("est_one_iteration")(self.h, iteration, C.byref(conv)), "est_one_iteration")
        return bool(conv.value)

    def run(self, max_iter: int = 30):
        conv, it = C.c_int(0), C.c_int(0)
        self._c(self.lib.fn("est_run")(self.h, max_iter, C.byref(conv), C.byref(it)), "est_run")
        return bool(conv.value), int(it.value)

    def get_transform(self, direction=0):
        A = np.zeros(9); t = np.zeros(3); c = np.zeros(3); cov = np.zeros(self.dof * self.dof); k = C.c_int(0)
        self._c(self.lib.fn("est_get_transform")(self.h, direction, C.byref(k), _dp(A), _dp(t), _dp(c), _dp(cov)), "est_get_transform")
        return {"kind": k.value, "A": A.reshape(3, 3), "t": t, "c": c, "cov": cov.reshape(self.dof, self.dof)}

    def get_matches(self, direction=0):
        n = C.c_int64(0)
        self._c(self.lib.fn("est_get_matches")(self.h, direction, 0, C.byref(n), None, None, None), "est_get_matches")
        m = int(n.value)
        fi = np.empty(m, dtype=np.int32); ti = np.empty(m, dtype=np.int32); w = np.empty(m)
        p32 = C.POINTER(C.c_int32)
        self._c(self.lib.fn("est_get_matches")(self.h, direction, m, C.byref(n), fi.ctypes.data_as(p32), ti.ctypes.data_as(p32), _dp(w)),
                "est_get_matches")
        return fi, ti, w

    def metric(self, which: str, direction=0) -> float:
        out = np.zeros(1)
        self._c(self.lib.fn("est_metric")(self.h, METRIC[which], direction, _dp(out)), "est_metric")
        return float(out[0])

    def estimate_ls_backward(self, estimate_parameters: bool = False):
        self._c(self.lib.fn("est_estimate_ls_backward")(self.h, int(estimate_parameters)), "est_estimate_ls_backward")

    def transfer_cov(self, direction=0) -> np.ndarray:
        S = np.zeros(9)
        self._c(self.lib.fn("est_transfer_cov")(self.h, direction, _dp(S)), "est_transfer_cov")
        return S.reshape(3, 3)

    def state(self) -> EstState:
        s = EstState()
        self._c(self.lib.fn("est_get_state")(self.h, C.byref(s)), "est_get_state")
        return s

    def stats(self) -> dict:
        s = EstStats()
        self._c(self.lib.fn("est_get_stats")(self.h, C.byref(s)), "est_get_stats")
        return {"cert_failed": list(s.cert_failed), "passb_queries": list(s.passb_queries), "work_list_entries": list(s.work_list_entries),
                "window_misses": int(s.window_misses), "spec_misses": int(s.spec_misses)}

    def oscillated(self) -> bool:
        o = C.c_int(0)
        self._c(self.lib.fn("est_oscillated")(self.h, C.byref(o)), "est_oscillated")
        return bool(o.value)

    def objcov_eval(self, x, direction=0, want_g=True):
        x = _f64(x); f = np.zeros(1); g = np.zeros_like(x) if want_g else None
        self._c(self.lib.fn("objcov_eval")(self.h, direction, _dp(x), _dp(f), _dp(g)), "objcov_eval")
        return float(f[0]), g

    def objcov_refine(self, max_evals=5, direction=0):
        fs, fe = np.zeros(1), np.zeros(1); ne = C.c_int(0)
        self._c(self.lib.fn("objcov_refine")(self.h, direction, max_evals, _dp(fs), _dp(fe), C.byref(ne)), "objcov_refine")
        return float(fs[0]), float(fe[0]), int(ne.value)

    def objpar_eval(self, theta, direction=0, want_g=True):
        """cdcl_obj_fun_par::compute at theta = (A row-major, t): f and the 12 gradient entries"""
        x = _f64(theta); f = np.zeros(1); g = np.zeros_like(x) if want_g else None
        self._c(self.lib.fn("objpar_eval")(self.h, direction, _dp(x), _dp(f), _dp(g)), "objpar_eval")
        return float(f[0]), g

    def objpar_refine(self, max_evals=5, direction=0):
        fs, fe = np.zeros(1), np.zeros(1); ne = C.c_int(0); cv = C.c_int(0)
        self._c(self.lib.fn("objpar_refine")(self.h, direction, max_evals, _dp(fs), _dp(fe), C.byref(ne), C.byref(cv)), "objpar_refine")
        return float(fs[0]), float(fe[0]), int(ne.value), bool(cv.value)

    # ---- E-step of the full CDC estimator (cdcl_estimation::compute_weights, cdcl/cdcl_estimation.txx:342-545) ----
    @staticmethod
    def _cdc_dict(s):
        return {"rows": int(s.rows), "entries": int(s.entries), "min_matches": int(s.min_matches), "max_matches": int(s.max_matches),
                "avg_largest_eval": float(s.avg_largest_eval)}

    def cdc_compute_weights(self, max_matches=100, eval_multiplier=5.0, direction=0):
        s = CdcSummary()
        self._c(self.lib.fn("cdc_compute_weights")(self.h, direction, max_matches, eval_multiplier, C.byref(s)), "cdc_compute_weights")
        return self._cdc_dict(s)

    def cdc_get_matches(self, direction=0):
        """(row_from[rows], row_start[rows + 1], to_idx[entries], w[entries]): the many-to-many match set in CSR form"""
        nr, ne = C.c_int64(0), C.c_int64(0)
        fn = self.lib.fn("cdc_get_matches")
        self._c(fn(self.h, direction, 0, 0, C.byref(nr), C.byref(ne), None, None, None, None), "cdc_get_matches")
        rf = np.zeros(nr.value, np.int32); rs = np.zeros(nr.value + 1, np.int64); ti = np.zeros(ne.value, np.int32); w = np.zeros(ne.value)
        self._c(fn(self.h, direction, nr.value, ne.value, None, None, rf.ctypes.data_as(C.POINTER(C.c_int32)),
                   rs.ctypes.data_as(C.POINTER(C.c_int64)), ti.ctypes.data_as(C.POINTER(C.c_int32)), _dp(w)), "cdc_get_matches")
        return rf, rs, ti, w

    def cdc_one_iteration(self, iteration, max_matches=100, eval_multiplier=5.0, direction=0):
        s = CdcSummary(); cv = C.c_int(0)
        self._c(self.lib.fn("cdc_one_iteration")(self.h, direction, iteration, max_matches, eval_multiplier, C.byref(cv), C.byref(s)), "cdc_one_iteration")
        return bool(cv.value), self._cdc_dict(s)

    def close(self):
        if self.h:
            self.lib.fn("est_destroy")(self.h)
            self.h = None
