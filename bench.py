#!/usr/bin/env python
"""bench.py — ICP iterations/s of the CDCL robust-ICP hot path on B200 (BASELINE.json metric).

A "step" is one `one_iteration` call of the product-default estimator (rrl_estimation_symmetric_ICP_matching_all<3,12>,
SURVEY.md §3.2) at the FINEST level (all points of both directions) on a synthetic lung-CT-sized pair with 100 000 shape features
per volume INSIDE a ramped registration: every registration starts from the initial transform, runs the
reference's own point-count ramp 1000 -> x2 -> all and stops at the driver's convergence test
(rrl/itkLocationRegistration.cxx:2418); the timed steps are its finest-level iterations — the transform still moves in every one of
them, the first has no neighbour certificates at all — gathered over as many registrations as it takes (K steps after W untimed
ones).  `value` = K / (sum of their CUDA-event times), L2 flushed before each.  The post-convergence fixed point that round 1
reported as the headline (iterate on until the transform stops changing: every certificate holds, three launches per
iteration) is kept as `value_fixed_point`: the driver loop never gets there.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun): the path shards only across independent registrations, so every rank runs its own replica of
the single registration on its own GPU, no data-path collective; `value` = all ranks' iterations divided by the max-over-ranks
time ("weak" scaling, replicas).  The batched workload C4 (`registrations`) is the FIXED list of 256 locations x 32
initialisations sharded block-cyclically over the ranks (strong scaling), one result gather at the end.

One JSON line on stdout (rank 0).  `e2e` = the same metric through the C ABI with HOST buffers (upload of both volumes, estimator
construction, full ramped run, result read-back inside the timed region).  `parity` compares the GPU registration with the CPU arm
run on the same pair in the same process (iteration counts, neighbour index sets, transform, covariance) and the run FAILS
(exit 3) when it is outside the north_star tolerance.  `c5` is config 5 (1 M features per side).  `--impl reference` times the
reference's own estimator code on the host (oracle/_ref: its sources compiled unmodified against oracle/vxl_shim; the oracle port
if absent).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FEATURES = 100_000
SEED = 3
WORKLOAD = "C3: rrl symmetric ICP matching-all <3,12>, 100k shape features per volume, finest-level iterations inside ramped registrations"
METRIC = "icp_iterations_per_s_100k_features"
UNIT = "iterations/s"
# compulsory bytes per forward affine iteration = 124 B per match + 24 B per fixed feature
BYTES_PER_MATCH, BYTES_PER_FIXED = 124, 24


def algorithmic_bytes_iteration(m_f, n_f, m_b, n_b):
    return BYTES_PER_MATCH * (m_f + m_b) + BYTES_PER_FIXED * (n_f + n_b)


# per-kernel algorithmic bytes, per match of one direction / per to-feature
KERNEL_BYTES = {
    # x 24 + q 24 + P 48 (reads) + idx 4 + r2 8 + w 8 (writes); + 24 B per to-feature read once by the NN search
    "match": (116, 24),
    "norm1": (24 + 4 + 8 + 24, 0),
    "norm2": (4 + 8 + 24, 0),
    "weight": (8 + 4 + 8 + 8 + 24 + 8 + 8 + 8, 0),
    "collect": (16, 0),
    "accumulate": (24 + 4 + 8 + 8 + 8 + 24 + 48 + 8, 0),
    "resid": (24 + 4 + 24 + 48 + 8 + 8, 0),
    # certificate pass = residual pass of the finished iteration + neighbour test of the next: x 24 + certificate 32 + previous neighbour 4 +
    # q 24 + P 48 + w 8 (reads) + r2 8 (write)
    "cert_resid": (24 + 32 + 4 + 24 + 48 + 8 + 8, 0),
    "restamp": (4 + 8, 0),
    # match completion: per RE-SEARCHED query (not per match): x 24 + q 24 + P 48 + idx 4 + r2 8 + certificate 32; the candidates read
    # by the search itself are not compulsory traffic
    "search": (140, 0), "queue": (140, 0),
}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML in a thread (nvidia-smi -lms
    block-buffers its pipe, so a short run only ever delivered one line); nvidia-smi one-shot calls as the fallback."""

    def __init__(self, gpu_index):
        self.rows, self.gpu, self.stop_flag, self.t, self.mode = [], gpu_index, threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.mode = pynvml, "nvml"
        except Exception:
            self.mode = "smi"
        self.t = threading.Thread(target=self._loop, daemon=True); self.t.start()

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                if self.mode == "nvml":
                    nv = self.nv
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM); mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    names = []
                    for nm, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
                        bit = getattr(nv, attr, None)
                        if bit is not None and (r & bit):
                            names.append(nm)
                    self.rows.append((float(sm), float(mx), names))
                else:
                    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    names = [nm for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:6])
                             if v.strip().lower().startswith("active")]
                    self.rows.append((float(out[0]), float(out[1]), names))
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler not started"]}
        self.stop_flag.set(); self.t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples (NVML / nvidia-smi unavailable)"]}
        reasons = sorted({n for r in self.rows for n in r[2]})
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows), "source": self.mode}


def make_workload():
    from lrr_b200 import synth
    return synth.make_pair(SEED, N_FEATURES, N_FEATURES)


def ramp_to_finest(est, max_iter=70):
    """The reference's driver loop (rrl/itkLocationRegistration.cxx:2418): one_iteration until converged."""
    it, conv = 0, False
    while not conv and it < max_iter:
        conv = est.one_iteration(it)
        it += 1
    return it, conv


def cpu_library():
    """CPU arm library: oracle/_ref/liblrr_ref.so when it exists = the REFERENCE's own estimator sources compiled unmodified
    against oracle/vxl_shim (kind "reference"; the prebuilt .so travels to the GPU box), else the oracle port (kind "port")."""
    from lrr_b200.capi import Lib
    ref = os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so")
    if os.path.exists(ref) and not os.environ.get("LRR_BENCH_CPU_PORT"):
        return Lib(ref, "lrrr_"), "reference"
    so = os.path.join(ROOT, "oracle", "liblrr_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liblrr_oracle.so"])
    return Lib(so, "lrro_"), "port"


CPU_KIND_NOTE = {
    "reference": "the reference's own rrl_estimation_symmetric_ICP_matching_all<3,12> (sources compiled unmodified, oracle/_ref; "
                 "VXL/ITK third-party calls served by oracle/vxl_shim, exact-NN k-d tree instead of the Voronoi image)",
    "port": "oracle/ CPU port of the same algorithm (oracle/_ref not built on this machine)",
}
CPU_RASTER_NOTE = ("the reference's own rrl_estimation_symmetric_ICP_matching_all<3,12> (sources compiled unmodified, oracle/_ref) on its STOCK matching path: "
                   "O(1) lookup in a rasterised 1 mm Voronoi label volume (rrl/rrl_estimation_ICP_matching_all.txx:530-543; the volumes are built "
                   "untimed, like the voronoi.mhd input files upstream).  Its match sets differ from the exact neighbour in ~10 % of the queries "
                   "(1 mm voxels, SURVEY Q5), so parity is pinned on the exact-NN variant (exact_nn_variant) and this figure is timing only")


def cpu_measure(pair, steps, warmup, keep=False, voronoi="exact"):
    """CPU arm, single thread: the reference pins its estimators to one thread
    (rrl/rrl_estimation_symmetric_ICP_matching_all.txx:760) and one registration has no other parallel axis.
    Same metric as the b200 arm: the finest-level one_iteration calls INSIDE ramped registrations (every registration starts again
    from the initial transform and stops at the driver's convergence test, rrl/itkLocationRegistration.cxx:2418), `steps` of them
    timed after `warmup` untimed ones.  keep: also return the final state of the first registration (bench parity block)."""
    lib, kind = cpu_library()
    ctx = lib.context()
    # voronoi = "raster": the stock reference path (O(1) voxel lookup), timing only; "exact": the exact-NN map parity is pinned on
    raster_s = None
    if kind == "reference" and lib.has("voronoi_raster_seconds"):
        os.environ["LRR_REF_VORONOI"] = voronoi
        raster_s = lib.fn("voronoi_raster_seconds"); raster_s.restype = ctypes.c_double; raster_s.argtypes = []
    elif voronoi != "exact":
        return None
    r0 = raster_s() if raster_s else 0.0
    t0 = time.perf_counter()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator("rrl_sym", mv, fx, 12)
    t_setup = time.perf_counter() - t0 - ((raster_s() - r0) if raster_s else 0.0)   # (the label volumes are input files upstream)
    os.environ["LRR_REF_VORONOI"] = "exact"
    A, t, c = pair["init"]
    samples, seen, regs, first = [], 0, [], None
    while len(samples) < steps:
        t0 = time.perf_counter()
        est.set_transform(A, t, c)
        it, conv = 0, False
        while not conv and it < 70:
            finest = est.state().finest_level == 1
            ta = time.perf_counter(); conv = est.one_iteration(it); dt = time.perf_counter() - ta
            it += 1
            if finest:
                seen += 1
                if seen > warmup and len(samples) < steps:
                    samples.append(dt)
        regs.append((time.perf_counter() - t0 + (t_setup if not regs else 0.0), it, bool(conv)))
        if first is None and keep:
            first = {"T": est.get_transform(0), "Tb": est.get_transform(1), "m0": est.get_matches(0), "m1": est.get_matches(1), "iters": it, "conv": bool(conv)}
        if it == 0:
            break
    dt = float(np.sum(samples))
    out = {"kind": kind, "finest_iters_per_s": len(samples) / dt, "ms_per_step": 1e3 * dt / len(samples), "samples": len(samples),
           "registration_s": regs[0][0], "registration_iters": regs[0][1], "registration_iters_per_s": regs[0][1] / regs[0][0], "converged": regs[0][2]}
    if keep:
        out["first"] = first
    return out


BATCH_SEED, BATCH_SIZE_MM, BATCH_LOCATIONS, BATCH_INITS = 4, 160.0, 256, 32


def make_batch_workload(world):
    """C4: one 100k/100k pair, the FIXED list of 256 locations x 32 initialisations (first near-correct, the
    rest perturbed up to 30 deg / 20 mm) = 8192 jobs at every N, 50 mm ROI cubes, rrl symmetric <3,12>, <= 70 iterations each."""
    from lrr_b200 import synth
    pair = synth.make_pair(BATCH_SEED, N_FEATURES, N_FEATURES, size=BATCH_SIZE_MM)
    jobs = synth.make_jobs(pair, BATCH_LOCATIONS, BATCH_INITS, seed=BATCH_SEED)
    return pair, jobs


def verified_count(res):
    """converged and not oscillated and validTransform (status 0) — rrl/itkLocationRegistration.cxx:2495-2507,2914-2926
    (the reference's SVM verification needs a model file that is not in the repository)."""
    return int(((res["status"] == 0) & (res["converged"] == 1) & (res["oscillated"] == 0)).sum())


def cpu_batch_measure(pair, jobs, n_single=4):
    from lrr_b200.capi import Lib, RESULT_DTYPE
    lib = Lib(os.path.join(ROOT, "oracle", "liblrr_oracle.so"), "lrro_")
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    t0 = time.perf_counter(); ctx.batch_run(mv, fx, jobs[:n_single], 70, BATCH_SEED); t1 = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    nall = min(len(jobs), 2 * cores)
    out = np.zeros(nall, dtype=RESULT_DTYPE)
    f = lib.dll.lrro_batch_run_mt
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
    jj = np.ascontiguousarray(jobs[:nall])
    t0 = time.perf_counter()
    f(ctx.h, mv.h, fx.h, jj.ctypes.data_as(ctypes.c_void_p), nall, 70, BATCH_SEED, out.ctypes.data_as(ctypes.c_void_p), cores)
    tall = time.perf_counter() - t0
    return {"registrations_per_s_1core": n_single / t1, "registrations_per_s_allcores": nall / tall, "cores_all": cores,
            "sample": f"{n_single} jobs on 1 thread; {nall} jobs on {cores} threads (parallel over registrations)"}


def run_reference(args, rank):
    if rank != 0:
        return
    pair = make_workload()
    steps = max(1, min(args.steps, 30))   # bounded sample: ~0.4 s of CPU per finest-level iteration, ~2.4 s per ramped registration
    warm = max(0, min(args.warmup, 5))
    rx = cpu_measure(pair, max(1, min(steps, 4)), 0)            # exact-NN variant (what parity is pinned on), short
    r = cpu_measure(pair, steps, warm, voronoi="raster") or rx     # stock path: O(1) Voronoi-volume lookup
    stock = r is not rx
    cores = 1
    line = {
        "impl": "reference", "metric": METRIC, "value": r["finest_iters_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_fixed": N_FEATURES, "n_moving": N_FEATURES, "seed": SEED, "l2": "host CPU run"},
        "cpu_baseline": {"value": r["finest_iters_per_s"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                         "sample": f"{steps} finest-level one_iteration calls inside ramped registrations ({warm} untimed first); single thread = the "
                                   "reference's execution model (its MultiThreader is pinned to 1 thread); " + (CPU_RASTER_NOTE if stock else CPU_KIND_NOTE[r["kind"]]),
                         "matching": "voronoi_raster_1mm (stock)" if stock else "exact_nn",
                         "exact_nn_variant": {"value": rx["finest_iters_per_s"], "registration_iters_per_s": rx["registration_iters_per_s"], "samples": rx["samples"]},
                         "host_cores_available": os.cpu_count()},
        # same definition as the b200 arm's e2e: iterations of one whole ramped registration / its wall time (setup included),
        # so that the driver's e2e ratio compares like with like; `value` is the finest-level figure like the b200 arm's `value`
        "e2e": {"value": r["registration_iters_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "what": "feature-set construction (k-d trees), estimator, set_transform, ramped run to convergence (<=70 one_iteration calls)"},
        "registration": {"seconds": r["registration_s"], "iterations": r["registration_iters"],
                         "iters_per_s": r["registration_iters_per_s"], "converged": r["converged"]},
        "gpu_launches": 0,
        "note": CPU_RASTER_NOTE if stock else CPU_KIND_NOTE[r["kind"]],
    }
    print(json.dumps(line), flush=True)


def feature_bytes(feats):
    return int(sum(v.nbytes for v in feats.values() if v is not None))


def pin(feats):
    import torch
    out = {}
    for k, v in feats.items():
        if v is None:
            out[k] = None
            continue
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        out[k] = t.numpy()
        out["_keep_" + k] = t
    return out


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def gpu_registration(est, pair, on_finest=None):
    """One ramped registration from the initial transform to the driver's convergence test.  on_finest(call) wraps every
    finest-level one_iteration call (timing); returns (iterations, converged, finest-level iterations)."""
    A0, t0_, c0 = pair["init"]
    est.set_transform(A0, t0_, c0)
    it, conv, nf = 0, False, 0
    while not conv and it < 70:
        finest = est.state().finest_level == 1
        if finest and on_finest is not None:
            conv = on_finest(lambda: est.one_iteration(it))
        else:
            conv = est.one_iteration(it)
        nf += int(finest)
        it += 1
    return it, conv, nf


def run_c5(ctx, stream, peak, steps):
    """Config 5: 1 M features per side, cdcl_estimation_ICP_matching_all<3,12> over every point to convergence,
    then cdcl_obj_fun_cov<3,12> + L-BFGS (<= 5 evaluations) on the final matches.  The only configuration whose working set
    (148 MB per iteration) does not fit the L2."""
    import torch
    from lrr_b200 import synth
    n = 1_000_000
    pair = synth.make_pair(5, n, n)
    with torch.cuda.stream(stream):
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        est = ctx.estimator("cdcl_fwd", mv, fx, 12)
        est.set_flags(True, True)
        times = []
        for rep in range(2):   # (the first run warms the memory pool and the order-statistic windows)
            est.set_transform(*pair["init"]); est.set_number_points(n)
            it, conv, times = 0, False, []
            while not conv and it < 30:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); conv = est.one_iteration(it); b.record(stream); b.synchronize()
                times.append(a.elapsed_time(b)); it += 1
        run_iters, run_conv = it, conv
        ctx.set_profiling(True)
        for k in range(steps):
            est.one_iteration(it); it += 1
        ph_it = ctx.phase_times()
        cov = est.get_transform(0)["cov"]
        U = np.linalg.cholesky(cov).T
        x = np.concatenate([U[i, i:] for i in range(12)])
        est.objcov_eval(x)   # untimed: the first evaluation allocates its buffers
        ctx.set_profiling(True)
        for k in range(5):
            f0, g0 = est.objcov_eval(x)
        ph_obj = ctx.phase_times(); ctx.set_profiling(False)
        t0 = time.perf_counter(); fs, fe, ev = est.objcov_refine(5); t_ref = time.perf_counter() - t0
        est.close(); fx.close(); mv.close()
    it_bytes = 124 * n + 24 * n
    in_run_ms = float(np.mean(times[1:])) if len(times) > 1 else float(times[0])
    steady_ms = float(sum(v[0] / v[1] for v in ph_it.values()))
    obj_us = 1e3 * ph_obj["objcov"][0] / ph_obj["objcov"][1]
    return {
        "workload": "C5: whole-volume pair, 1M features per side, cdcl forward affine <3,12> over every point + cdcl_obj_fun_cov L-BFGS refinement",
        "run": {"iterations": run_iters, "converged": bool(run_conv), "ms_per_iteration_in_run": in_run_ms, "ms_first_iteration": float(times[0])},
        "iteration": {"algorithmic_bytes": it_bytes, "ms_in_run": in_run_ms, "gbs_in_run": it_bytes / (in_run_ms * 1e-3) / 1e9,
                      "frac_in_run": it_bytes / (in_run_ms * 1e-3) / 1e9 / peak,
                      "ms_after_convergence_sum_of_kernels": steady_ms, "frac_after_convergence": it_bytes / (steady_ms * 1e-3) / 1e9 / peak,
                      "kernels_us": {k: 1e3 * v[0] / v[1] for k, v in ph_it.items()}},
        "objcov_eval": {"algorithmic_bytes": 156 * n, "us": obj_us, "gbs": 156 * n / (obj_us * 1e-6) / 1e9, "frac": 156 * n / (obj_us * 1e-6) / 1e9 / peak,
                        "f": float(f0)},
        "objcov_refine": {"f_start": float(fs), "f_end": float(fe), "evaluations": int(ev), "seconds": t_ref},
    }


def run_cdc(ctx, stream, peak):
    """The E-step of the full CDC estimator at C3 size: cdcl_estimation<3,12>::compute_weights over 100k / 100k features
    under the transform and covariance of three ICP iterations, then cdcl_obj_fun_cov / cdcl_obj_fun_par over the many-to-many set."""
    import torch
    from lrr_b200 import synth
    n = N_FEATURES
    pair = synth.make_pair(SEED, n, n)
    with torch.cuda.stream(stream):
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        est = ctx.estimator("cdcl_fwd", mv, fx, 12)
        est.set_flags(True, True); est.set_transform(*pair["init"])
        for it in range(3):
            est.one_iteration(it)
        s = est.cdc_compute_weights()   # untimed: allocates the flat set and the scratch lists
        ctx.set_profiling(True)
        for k in range(5):
            est.cdc_compute_weights()
        ph = ctx.phase_times()
        T = est.get_transform(0)
        U = np.linalg.cholesky(T["cov"]).T
        x = np.concatenate([U[i, i:] for i in range(12)]); th = np.concatenate([T["A"].reshape(9), T["t"]])
        est.objcov_eval(x); est.objpar_eval(th)
        ctx.set_profiling(True)
        for k in range(5):
            est.objcov_eval(x)
        ph_cov = ctx.phase_times()
        ctx.set_profiling(True)
        for k in range(5):
            est.objpar_eval(th)
        ph_par = ctx.phase_times(); ctx.set_profiling(False)
        est.close(); fx.close(); mv.close()
    us_match = 1e3 * ph["match"][0] / ph["match"][1]
    us_cov = 1e3 * ph_cov["objcov"][0] / ph_cov["objcov"][1]; us_par = 1e3 * ph_par["objcov"][0] / ph_par["objcov"][1]
    e = s["entries"]
    return {"workload": "cdcl_estimation<3,12>::compute_weights (covariance-driven radius matching, cap 100, radius 5 sqrt(lambda_max)), 100k features per side",
            "rows": s["rows"], "entries": e, "max_matches": s["max_matches"], "avg_largest_eval": s["avg_largest_eval"],
            "compute_weights": {"kernel": "k_cdc_match", "us": us_match, "entries_per_s": e / (us_match * 1e-6),
                                "note": "latency / instruction bound: the fixed set is L2-resident (profiles/r2_ncu_full_k_cdc_match.csv: 34 MB of DRAM traffic, FP64 pipe 27 %)"},
            "objcov_eval": {"us": us_cov, "algorithmic_bytes": 156 * e, "gbs": 156 * e / (us_cov * 1e-6) / 1e9, "frac": 156 * e / (us_cov * 1e-6) / 1e9 / peak},
            "objpar_eval": {"us": us_par, "algorithmic_bytes": 156 * e, "gbs": 156 * e / (us_par * 1e-6) / 1e9, "frac": 156 * e / (us_par * 1e-6) / 1e9 / peak}}


def run_b200(args, rank, world, local_rank):
    import torch
    import lrr_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    out_fd = 1
    if world > 1:
        # stdout carries exactly one JSON line: NCCL writes its version banner to fd 1, so libraries get stderr there
        sys.stdout.flush()
        out_fd = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # one process per GPU on one host: every rank spins on its result mailbox and feeds its own launch queue, so each gets its own
        # share of the host cores (a rank that lands on another rank's core was measured 20 % slower per iteration at N = 8)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if per >= 2:
                os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
        except (AttributeError, OSError):
            pass
    lib = lrr_b200.lib()
    pair = make_workload()
    stream = torch.cuda.Stream()
    ctx = lib.context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    A0, t0_, c0 = pair["init"]
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        est = ctx.estimator("rrl_sym", mv, fx, 12)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

        # ---- warm-up: W finest-level iterations = whole untimed registrations; the first one is also the parity registration ----
        seen = 0
        reg_iters, reg_conv, reg_finest = gpu_registration(est, pair)
        seen += reg_finest
        par_gpu = {"T": est.get_transform(0), "Tb": est.get_transform(1), "m0": est.get_matches(0), "m1": est.get_matches(1),
                   "iters": reg_iters, "conv": bool(reg_conv)}
        st = est.state()
        assert st.finest_level == 1 and st.number_points[0] == N_FEATURES and st.number_points[1] == N_FEATURES, "ramp did not reach the finest level"
        while seen < W:
            seen += gpu_registration(est, pair)[2]
        sampler = ClockSampler(local_rank); sampler.start()

        # ---- timed region 1 (headline): exactly K finest-level iterations inside ramped registrations, L2 flushed before each,
        #      CUDA events on the launching stream ----
        evs, warm_evs = [], []

        def timed(call, store, do_flush):
            if len(store) >= K:
                return call()
            if do_flush:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); r = call(); b.record(stream)
            store.append((a, b))
            return r

        barrier()
        launches0 = ctx.launch_count()
        n_regs = 0
        while len(evs) < K:
            gpu_registration(est, pair, lambda call: timed(call, evs, True)); n_regs += 1
        barrier()
        launches = ctx.launch_count() - launches0
        ms_iters = np.array([a.elapsed_time(b) for a, b in evs])
        total_ms = float(ms_iters.sum())
        # ---- the same without the flush (the production state: a 30 MB working set stays in the 126 MB L2) ----
        barrier()
        while len(warm_evs) < K:
            gpu_registration(est, pair, lambda call: timed(call, warm_evs, False))
        barrier()
        warm_ms_iters = np.array([a.elapsed_time(b) for a, b in warm_evs])
        warm_ms = float(warm_ms_iters.sum())
        # ---- per-kernel CUDA-event times of the finest-level iterations inside registrations (L2 flushed) ----
        def profiled(call):
            flush.fill_(1); ctx.set_profiling_keep(True); r = call(); ctx.set_profiling_keep(False); return r
        ctx.set_profiling(True); ctx.set_profiling_keep(False)
        nprof = 0
        while nprof < max(10, min(K, 20)):
            nprof += gpu_registration(est, pair, profiled)[2]
        phases = ctx.phase_times(); ctx.set_profiling(False)
        stats_last = est.stats()

        # ---- fixed point (round 1's headline): iterate past convergence until the transform has stopped changing ----
        gpu_registration(est, pair)
        it_no = 70
        settle_iters, prevT = 0, est.get_transform(0)
        while settle_iters < 24:
            est.one_iteration(it_no); it_no += 1; settle_iters += 1
            T = est.get_transform(0)
            moved = max(float(np.abs(T["A"] - prevT["A"]).max()), float(np.abs(T["t"] - prevT["t"]).max()))
            prevT = T
            if moved < 1e-12:
                break
        for _ in range(3):
            est.one_iteration(it_no); it_no += 1
        barrier()
        fp_evs = []
        for _ in range(K):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); est.one_iteration(it_no); it_no += 1; b.record(stream)
            fp_evs.append((a, b))
        barrier()
        fp_ms = float(sum(a.elapsed_time(b) for a, b in fp_evs))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(K):
            est.one_iteration(it_no); it_no += 1
        b.record(stream)
        barrier()
        fp_warm_ms = a.elapsed_time(b)
        ctx.set_profiling(True)
        for _ in range(min(K, 20)):
            flush.fill_(1); est.one_iteration(it_no); it_no += 1
        fp_phases = ctx.phase_times(); ctx.set_profiling(False)
        clocks = sampler.stop()   # sampled across both timed regions, the per-kernel passes and the fixed point

    tm = torch.tensor([total_ms, warm_ms, fp_ms, fp_warm_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms_max, warm_ms_max, fp_ms_max, fp_warm_ms_max = (float(x) for x in tm)
    value = world * K / (total_ms_max * 1e-3)
    value_warm = world * K / (warm_ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers (rank-local; max over ranks) ----
    e2e_steps = 0 if args.only_iterations else max(1, min(5, args.steps))
    # the rrl estimators work on cdcl_feature_ICP (location, float strength, shape, error projector, directions — no
    # covariance, cdcl/cdcl_feature_ICP.h:23-96), so the end-to-end call uploads exactly those arrays
    e2e_fixed = {k: v for k, v in pair["fixed"].items() if k != "C6"}
    e2e_moving = {k: v for k, v in pair["moving"].items() if k != "C6"}
    pf, pm = pin(e2e_fixed), pin(e2e_moving)
    h2d = feature_bytes(e2e_fixed) + feature_bytes(e2e_moving)
    d2h = (9 + 3 + 3 + 144) * 8
    # asynchronous upload (lrr_ctx_set_async_upload): the direction arrays — 47 % of the bytes, read by the angle metrics only — cross PCIe
    # while the registration already runs; the copy is complete before the step ends (lrr_features_destroy waits for it), so every byte of
    # h2d_bytes_per_step is moved inside the timed region
    ctx.set_async_upload(True)
    with torch.cuda.stream(stream):
        e2e_iters = 0
        t_start = time.perf_counter()
        for step in range(-1 if e2e_steps else 0, e2e_steps):      # step -1: untimed warm-up (memory pool growth, first-touch)
            if step == 0:
                barrier()
                t_start = time.perf_counter()
                e2e_iters = 0
            ta = time.perf_counter()
            fx2 = ctx.upload({k: v for k, v in pf.items() if not k.startswith("_keep_")})
            mv2 = ctx.upload({k: v for k, v in pm.items() if not k.startswith("_keep_")})
            tb = time.perf_counter()
            est2 = ctx.estimator("rrl_sym", mv2, fx2, 12)
            est2.set_transform(A0, t0_, c0)
            tc = time.perf_counter()
            n_it, _ = ramp_to_finest(est2)
            est2.get_transform(0)
            td = time.perf_counter()
            e2e_iters += n_it
            est2.close(); fx2.close(); mv2.close()
            te_ = time.perf_counter()
            print(f"[e2e] upload {tb - ta:.4f} est {tc - tb:.4f} run {td - tc:.4f} ({n_it} its) close {te_ - td:.4f}", file=sys.stderr)
        barrier()
        e2e_s = time.perf_counter() - t_start
    ctx.set_async_upload(False)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_iters / float(te[0]) if e2e_iters else None

    # ---- batched location registration: the fixed C4 list sharded block-cyclically over the ranks ----
    from lrr_b200.sharding import expected_costs, gather_results, run_batch_sharded
    bpair, bjobs = make_batch_workload(world)
    if args.only_iterations:
        bjobs = bjobs[:world]
    with torch.cuda.stream(stream):
        bfx, bmv = ctx.upload(bpair["fixed"]), ctx.upload(bpair["moving"])
        # warm-up: a rank's share of at least 1024 jobs, so that the workspace (a slot per job in flight, up to 1024: gigabytes, kept by the
        # context between calls) has its full size before the timed run
        run_batch_sharded(ctx, bmv, bfx, bjobs[: min(len(bjobs), 1024 * world)], 70, BATCH_SEED, rank, world)
        barrier()
        tb0 = time.perf_counter()
        # (N > 1: the shares are dealt over the jobs sorted by expected cost — lrr_batch_expected_cost, inside the timed region; within
        #  a rank the backend hands its jobs out longest-expected-first either way)
        bcosts = expected_costs(ctx, bmv, bfx, bjobs, world)
        local = run_batch_sharded(ctx, bmv, bfx, bjobs, 70, BATCH_SEED, rank, world, costs=bcosts)
        torch.cuda.synchronize()
        tb_local = time.perf_counter() - tb0
        full = gather_results(local, len(bjobs), rank, world, device=torch.device("cuda", local_rank) if world > 1 else None, costs=bcosts)
        barrier()
    tb = torch.tensor([tb_local], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    batch_s = float(tb[0])

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    M, N = N_FEATURES, N_FEATURES
    researched = float(sum(stats_last["cert_failed"]))   # queries of the last profiled iteration's certificate pass that had to be searched again

    def kernel_table(ph):
        tab = {}
        for name, (ms, cnt) in ph.items():
            pm_, pt_ = KERNEL_BYTES.get(name, (0, 0))
            ndirs = 1 if name in ("norm1", "norm2", "restamp") else 2
            by = ndirs * (pm_ * M + pt_ * N)
            if name in ("search", "queue"):
                by = None   # re-searched queries only: a few thousand of 200 000, not a streaming kernel (est.stats() in `neighbour_search`)
            us = 1e3 * ms / cnt
            tab[name] = {"us_per_launch": us, "launches": cnt, "algorithmic_bytes": by, "gbs": (by / (us * 1e-6) / 1e9) if (by and us > 0) else None}
        return tab
    per_phase, fp_per_phase = kernel_table(phases), kernel_table(fp_phases)
    KNAME = {"resid": "k_match_cert (residual pass of the iteration = certificate pass of the next find_closest)", "weight": "k_weight", "accumulate": "k_accumulate (+ solve)",
             "search": "k_match_tile (staged search + sparse list)", "queue": "k_match_hard", "norm1": "k_norm1", "norm2": "k_norm2", "restamp": "k_restamp", "collect": "k_collect", "match": "k_match_cert + k_match_tile + k_match_hard"}
    streaming = {k: v for k, v in per_phase.items() if v["gbs"]}
    dom = max(streaming.items(), key=lambda kv: kv[1]["us_per_launch"])[0]
    d = per_phase[dom]
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        names = {"accumulate": ["k_accumulate"], "weight": ["k_weight"], "resid": ["k_match_cert"]}.get(dom, [])
        if names and all(n in tj["dram_bytes_per_launch"] for n in names):
            traffic = sum(tj["dram_bytes_per_launch"][n] for n in names); traffic_src = tj["source"]
        first = tj.get("dram_bytes_per_launch_first_finest_iteration", {})
        if dom == "match" and all(n in first for n in ("k_match_cert", "k_match_tile", "k_match_hard")):
            # the `match` phase is timed only in the first finest-level iteration of a registration (afterwards the certificate pass is
            # the trailing pass of the previous iteration): its three launches together, from the capture of exactly that iteration
            traffic = sum(first[n] for n in ("k_match_cert", "k_match_tile", "k_match_hard")); traffic_src = tj["source"]
    except Exception:
        pass
    it_bytes = algorithmic_bytes_iteration(M, N, M, N)
    ms_step = total_ms_max / K
    fp_ms_step, fp_warm_ms_step = fp_ms_max / K, fp_warm_ms_max / K
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_fixed": N, "n_moving": M, "seed": SEED, "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2": "flushed (256 MiB write) before every timed iteration; working set ~30 MB is L2-resident otherwise (value_l2_warm)",
                   "registration_iterations": reg_iters, "finest_level_iterations_per_registration": reg_finest, "registrations_timed": n_regs,
                   "timed_steps": "the finest-level one_iteration calls of ramped registrations up to the driver's convergence test (transform still moving; the first of each registration has no certificates)"},
        "ms_per_step_median": float(np.median(ms_iters)), "ms_per_step_each": [round(float(x), 4) for x in ms_iters],
        "value_l2_warm": value_warm, "ms_per_step_l2_warm": warm_ms_max / K, "ms_per_step_l2_warm_median": float(np.median(warm_ms_iters)),
        "value_fixed_point": world * K / (fp_ms_max * 1e-3), "value_fixed_point_l2_warm": world * K / (fp_warm_ms_max * 1e-3),
        "fixed_point": {"what": "round 1's headline: the registration iterated past convergence until the transform stopped changing (every neighbour certificate holds; "
                                "certificate/residual pass -> weight -> accumulate+solve); the driver loop never reaches this state",
                        "ms_per_step": fp_ms_step, "ms_per_step_l2_warm": fp_warm_ms_step, "settle_iterations": settle_iters,
                        "roofline_iteration_frac": it_bytes / (fp_ms_step * 1e-3) / 1e9 / peak, "kernels": fp_per_phase},
        "roofline": {"bound": "hbm", "kernel": KNAME.get(dom, "k_" + dom), "achieved": d["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": (d["gbs"] / peak) if d["gbs"] else None, "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                     "us_per_launch": d["us_per_launch"], "algorithmic_bytes_per_launch": d["algorithmic_bytes"],
                     "note": "dominant streaming kernel of the in-registration finest-level iteration by CUDA-event time; per-kernel events serialise the "
                             "programmatic-dependent-launch overlap, so their sum exceeds ms_per_step"},
        "roofline_iteration": {"algorithmic_bytes": it_bytes, "achieved_gbs": it_bytes / (ms_step * 1e-3) / 1e9,
                               "frac": it_bytes / (ms_step * 1e-3) / 1e9 / peak,
                               "achieved_gbs_l2_warm": it_bytes / (warm_ms_max / K * 1e-3) / 1e9},
        "kernels": per_phase,
        "neighbour_search": {"queries_searched_again_last_profiled_iteration": researched, "of": 2 * M, "stats": stats_last},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "iterations": e2e_iters, "seconds": float(te[0]),
                "what": "per step: upload both volumes (cdcl_feature_ICP arrays: locations, projectors, strengths, shapes, directions) from pinned host memory (grid build; the directions asynchronously, lrr_ctx_set_async_upload, complete before the step ends), construct estimator, set_transform, "
                        "ramped run to convergence (<=70 one_iteration calls), read back transform+covariance"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "registrations": {"value": len(bjobs) / batch_s, "unit": "registrations/s", "jobs": int(len(bjobs)), "seconds": batch_s, "scaling": "strong (fixed list)",
                          "verified": verified_count(full), "verified_per_s": verified_count(full) / batch_s,
                          "rejected": int((full["status"] == 1).sum()), "mean_iterations": float(full["iterations"][full["status"] == 0].mean()),
                          "what": f"C4: {BATCH_LOCATIONS} locations x {BATCH_INITS} initialisations = {len(bjobs)} jobs on one resident 100k/100k pair, 50 mm ROI, "
                                  "rrl symmetric <3,12>, <= 70 iterations, block-cyclic shards over the ranks, one result gather"},
    }
    rc_exit = 0
    if world == 1 and not args.no_cpu and not args.only_iterations:
        # ---- CPU arm on the same pair in the same process: baseline AND parity of the benchmarked registration ----
        r = cpu_measure(pair, 6, 0, keep=True)
        c = r.pop("first")
        rs = cpu_measure(pair, 6, 0, voronoi="raster")   # the stock matching path (timing only); None when oracle/_ref is not built
        b = rs or r
        line["cpu_baseline"] = {"value": b["finest_iters_per_s"], "unit": UNIT, "cores": 1, "kind": b["kind"],
                                "sample": f"{b['samples']} finest-level one_iteration calls inside ramped registrations, single thread "
                                          "(the reference's execution model); " + (CPU_RASTER_NOTE if rs else CPU_KIND_NOTE[r["kind"]]),
                                "matching": "voronoi_raster_1mm (stock)" if rs else "exact_nn",
                                "registration_iters_per_s": b["registration_iters_per_s"], "registration_s": b["registration_s"],
                                "exact_nn_variant": {"value": r["finest_iters_per_s"], "registration_iters_per_s": r["registration_iters_per_s"],
                                                     "registration_s": r["registration_s"], "what": CPU_KIND_NOTE[r["kind"]]},
                                "host_cores_available": os.cpu_count()}
        g = par_gpu
        par = {"iters_equal": g["iters"] == c["iters"] and g["conv"] == c["conv"], "iterations": [g["iters"], c["iters"]],
               "idx_equal": bool(np.array_equal(g["m0"][1], c["m0"][1]) and np.array_equal(g["m1"][1], c["m1"][1]) and np.array_equal(g["m0"][0], c["m0"][0])),
               "rel_A": relerr(g["T"]["A"], c["T"]["A"]), "rel_t": relerr(g["T"]["t"], c["T"]["t"]), "rel_cov": relerr(g["T"]["cov"], c["T"]["cov"]),
               "rel_A_backward": relerr(g["Tb"]["A"], c["Tb"]["A"]), "rel_w": relerr(g["m0"][2], c["m0"][2]),
               "against": r["kind"], "tolerance": 1e-5,
               "what": "the first (untimed) GPU registration of this run against the CPU arm on the same pair: iteration count and convergence flag, neighbour index sets of both "
                       "directions at the end (np.array_equal), forward / backward transform and 12 x 12 covariance"}
        par["ok"] = bool(par["iters_equal"] and par["idx_equal"] and max(par["rel_A"], par["rel_t"], par["rel_cov"], par["rel_A_backward"]) <= 1e-5)
        line["parity"] = par
        if not par["ok"]:
            rc_exit = 3
        line["cpu_baseline"]["registrations"] = cpu_batch_measure(bpair, bjobs)
        # C4 `verified`, last clause of SURVEY.md §8(d): GPU theta / Sigma within 1e-5 of the CPU oracle — on a sample, outside the timed region
        line["registrations"]["verified_vs_oracle"] = verify_batch_sample(bpair, bjobs, full)
        if not line["registrations"]["verified_vs_oracle"]["ok"]:
            rc_exit = 3
    else:
        line["cpu_baseline"] = None
    if world == 1 and not args.only_iterations and not args.no_c5:
        try:
            line["c5"] = run_c5(ctx, stream, peak, 10)
        except Exception as ex:   # noqa: BLE001
            line["c5"] = {"error": repr(ex)}
        try:
            line["cdc_e_step"] = run_cdc(ctx, stream, peak)
        except Exception as ex:   # noqa: BLE001
            line["cdc_e_step"] = {"error": repr(ex)}
    sys.stdout.flush()
    os.write(out_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    if rc_exit:
        print("bench.py: PARITY FAILURE (see `parity` / `registrations.verified_vs_oracle` in the JSON line)", file=sys.stderr)
        sys.exit(rc_exit)


def verify_batch_sample(pair, jobs, gpu_results, n_sample=64):
    """64 jobs spread over the list, re-run by the CPU oracle (all host cores) and compared with the GPU results."""
    from lrr_b200.capi import Lib, RESULT_DTYPE
    lib = Lib(os.path.join(ROOT, "oracle", "liblrr_oracle.so"), "lrro_")
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    ids = np.unique(np.linspace(0, len(jobs) - 1, n_sample).astype(np.int64))
    ref = ctx.batch_run(mv, fx, jobs[ids], 70, BATCH_SEED, ids=ids)
    g = gpu_results[ids]
    same_int = all(np.array_equal(g[k], ref[k]) for k in ("status", "converged", "oscillated", "iterations", "n_moving_inside", "n_fixed_inside"))
    ran = ref["status"] == 0
    worst = 0.0
    for k in ("A", "t", "cov"):
        for a, b in zip(g[k][ran], ref[k][ran]):
            worst = max(worst, relerr(a, b))
    return {"jobs": int(len(ids)), "ran": int(ran.sum()), "integer_fields_equal": bool(same_int), "worst_rel_A_t_cov": worst, "tolerance": 1e-5,
            "ok": bool(same_int and worst <= 1e-5), "what": "status, converged, oscillated, iteration count and ROI sizes identical; A, t and covariance of every job that ran within 1e-5 relative"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)   # finest-level iterations inside registrations (~5 per registration)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--only-iterations", action="store_true", help="skip the e2e / batch / cpu legs (short run for ncu launch lists)")
    ap.add_argument("--no-c5", action="store_true", help="skip the config-5 leg (1 M features per side)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
