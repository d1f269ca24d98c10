#!/usr/bin/env python
"""bench.py — ICP iterations/s of the CDCL robust-ICP hot path on B200 (BASELINE.json metric).

A "step" is one `one_iteration` call of the product-default estimator
(rrl_estimation_symmetric_ICP_matching_all<3,12>) at the FINEST level (all points of both
directions) on a synthetic lung-CT-sized pair with 100 000 shape features per volume,
after the reference's own point-count ramp 1000 -> x2 -> all has been run and (untimed setup, like the ramp) the iteration has
been continued until the transform stopped changing: the metric is the steady-state iteration; `config` carries
the counts (`ramp_iterations`, `settle_iterations`).  The W warm-up steps come after that.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun): the path shards only across independent registrations, so every rank runs
its own replica of the registration on its own GPU, no data-path collective; `value` = all ranks' iterations
divided by the max-over-ranks time ("weak" scaling).

One JSON line on stdout (rank 0).  `value` = device-resident iterations/s with an L2 flush between timed
iterations; `e2e` = the same metric through the C ABI with HOST buffers (upload of both volumes, estimator
construction, full ramped run, result read-back inside the timed region).  `--impl reference` times the reference's own
estimator code on the host (oracle/_ref: its sources compiled unmodified against oracle/vxl_shim; the oracle port if absent).
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
WORKLOAD = "C3: rrl symmetric ICP matching-all <3,12>, 100k shape features per volume, finest level (all points)"
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
    # the whole iteration as one cooperative kernel: SURVEY.md §8(d) bytes per iteration (124 B per match + 24 B per to-feature)
    "fused_iteration": (BYTES_PER_MATCH, BYTES_PER_FIXED),
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


def cpu_measure(pair, steps, warmup):
    """CPU arm, single thread: the reference pins its estimators to one thread
    (rrl/rrl_estimation_symmetric_ICP_matching_all.txx:760) and one registration has no other parallel axis."""
    lib, kind = cpu_library()
    ctx = lib.context()
    t0 = time.perf_counter()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator("rrl_sym", mv, fx, 12)
    A, t, c = pair["init"]
    est.set_transform(A, t, c)
    iters, conv = ramp_to_finest(est)
    est.get_transform(0)
    t_reg = time.perf_counter() - t0
    for i in range(warmup):
        est.one_iteration(iters + i)
    t0 = time.perf_counter()
    for i in range(steps):
        est.one_iteration(iters + warmup + i)
    dt = time.perf_counter() - t0
    return {"kind": kind, "finest_iters_per_s": steps / dt, "ms_per_step": 1e3 * dt / steps, "registration_s": t_reg,
            "registration_iters": iters, "registration_iters_per_s": iters / t_reg, "converged": bool(conv)}


BATCH_SEED, BATCH_SIZE_MM, BATCH_LOCATIONS_PER_GPU, BATCH_INITS = 4, 160.0, 32, 32


def make_batch_workload(world):
    """C4, bounded: one 100k/100k pair, per GPU 32 locations x 32 initialisations (first near-correct,
    the rest perturbed up to 30 deg / 20 mm), 50 mm ROI cubes, rrl symmetric <3,12>, <= 70 iterations each."""
    from lrr_b200 import synth
    pair = synth.make_pair(BATCH_SEED, N_FEATURES, N_FEATURES, size=BATCH_SIZE_MM)
    jobs = synth.make_jobs(pair, BATCH_LOCATIONS_PER_GPU * world, BATCH_INITS, seed=BATCH_SEED)
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
    steps = max(1, min(args.steps, 30))   # bounded sample: ~0.9 s of CPU per finest-level iteration
    warm = max(1, min(args.warmup, 2))
    r = cpu_measure(pair, steps, warm)
    cores = 1
    line = {
        "impl": "reference", "metric": METRIC, "value": r["finest_iters_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_fixed": N_FEATURES, "n_moving": N_FEATURES, "seed": SEED, "l2": "host CPU run"},
        "cpu_baseline": {"value": r["finest_iters_per_s"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                         "sample": f"{steps} one_iteration calls at the finest level after the full ramp; single thread = the "
                                   "reference's execution model (its MultiThreader is pinned to 1 thread); " + CPU_KIND_NOTE[r["kind"]],
                         "host_cores_available": os.cpu_count()},
        # same definition as the b200 arm's e2e: iterations of one whole ramped registration / its wall time (setup included),
        # so that the driver's e2e ratio compares like with like; `value` is the finest-level figure like the b200 arm's `value`
        "e2e": {"value": r["registration_iters_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "what": "feature-set construction (k-d trees), estimator, set_transform, ramped run to convergence (<=70 one_iteration calls)"},
        "registration": {"seconds": r["registration_s"], "iterations": r["registration_iters"],
                         "iters_per_s": r["registration_iters_per_s"], "converged": r["converged"]},
        "gpu_launches": 0,
        "note": CPU_KIND_NOTE[r["kind"]],
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
    lib = lrr_b200.lib()
    pair = make_workload()
    stream = torch.cuda.Stream()
    ctx = lib.context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    A0, t0_, c0 = pair["init"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        est = ctx.estimator("rrl_sym", mv, fx, 12)
        est.set_transform(A0, t0_, c0)
        ramp_iters, ramp_conv = ramp_to_finest(est)
        st = est.state()
        assert st.finest_level == 1 and st.number_points[0] == N_FEATURES and st.number_points[1] == N_FEATURES, "ramp did not reach the finest level"
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

        # The metric is the STEADY-STATE iteration at the finest level.  The driver loop's convergence test fires
        # while the transform still moves by ~1e-4 mm per iteration; part of the setup (untimed, like the ramp) is therefore to
        # iterate on until the transform has stopped changing (at most 16 iterations; the count goes into `config`).
        it_no = ramp_iters
        settle_iters, prevT = 0, est.get_transform(0)
        while settle_iters < 16:
            est.one_iteration(it_no); it_no += 1; settle_iters += 1
            T = est.get_transform(0)
            moved = max(float(np.abs(T["A"] - prevT["A"]).max()), float(np.abs(T["t"] - prevT["t"]).max()))
            prevT = T
            if moved < 1e-12:
                break
        sampler = ClockSampler(local_rank); sampler.start()
        for _ in range(max(3, args.warmup)):
            est.one_iteration(it_no); it_no += 1

        # ---- timed region 1: K iterations, L2 flushed between iterations, CUDA events on the launching stream ----
        barrier()
        launches0 = ctx.launch_count()
        evs = []
        for _ in range(args.steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            est.one_iteration(it_no); it_no += 1
            b.record(stream)
            evs.append((a, b))
        barrier()
        launches = ctx.launch_count() - launches0
        ms_iters = np.array([a.elapsed_time(b) for a, b in evs])
        total_ms = float(ms_iters.sum())

        # ---- timed region 2: L2-warm back-to-back iterations (the production steady state) ----
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            est.one_iteration(it_no); it_no += 1
        b.record(stream)
        barrier()
        warm_ms = a.elapsed_time(b)

        # ---- per-kernel CUDA-event times (same conditions as region 1) ----
        ctx.set_profiling(True)
        for _ in range(args.steps):
            flush.fill_(1)
            est.one_iteration(it_no); it_no += 1
        phases = ctx.phase_times()
        ctx.set_profiling(False)
        clocks = sampler.stop()   # sampled across warm-up, both timed regions and the per-kernel pass
        final = est.get_transform(0)

    tm = torch.tensor([total_ms, warm_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms_max, warm_ms_max = float(tm[0]), float(tm[1])
    value = world * args.steps / (total_ms_max * 1e-3)
    value_warm = world * args.steps / (warm_ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers (rank-local; max over ranks) ----
    e2e_steps = 0 if args.only_iterations else max(1, min(5, args.steps))
    # the rrl estimators work on cdcl_feature_ICP (location, float strength, shape, error projector, directions — no
    # covariance, cdcl/cdcl_feature_ICP.h:23-96), so the end-to-end call uploads exactly those arrays
    e2e_fixed = {k: v for k, v in pair["fixed"].items() if k != "C6"}
    e2e_moving = {k: v for k, v in pair["moving"].items() if k != "C6"}
    pf, pm = pin(e2e_fixed), pin(e2e_moving)
    h2d = feature_bytes(e2e_fixed) + feature_bytes(e2e_moving)
    d2h = (9 + 3 + 3 + 144) * 8
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
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_iters / float(te[0]) if e2e_iters else None

    # ---- batched location registration, sharded over the ranks ----
    from lrr_b200.sharding import gather_results, run_batch_sharded
    bpair, bjobs = make_batch_workload(world)
    if args.only_iterations:
        bjobs = bjobs[:world]
    with torch.cuda.stream(stream):
        bfx, bmv = ctx.upload(bpair["fixed"]), ctx.upload(bpair["moving"])
        run_batch_sharded(ctx, bmv, bfx, bjobs[: min(len(bjobs), 512 * world)], 70, BATCH_SEED, rank, world)   # warm-up (workspace allocation)
        barrier()
        tb0 = time.perf_counter()
        local = run_batch_sharded(ctx, bmv, bfx, bjobs, 70, BATCH_SEED, rank, world)
        torch.cuda.synchronize()
        tb_local = time.perf_counter() - tb0
        full = gather_results(local, len(bjobs), rank, world, device=torch.device("cuda", local_rank) if world > 1 else None)
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
    # dominant kernel by CUDA-event time
    dom = max(phases.items(), key=lambda kv: kv[1][0])[0] if phases else "match"
    per_phase = {}
    for name, (ms, cnt) in phases.items():
        pm_, pt_ = KERNEL_BYTES.get(name, (0, 0))
        ndirs = 1 if name in ("norm1", "norm2", "resid") else 2
        by = ndirs * (pm_ * M + pt_ * N)
        us = 1e3 * ms / cnt
        per_phase[name] = {"us_per_launch": us, "launches": cnt, "algorithmic_bytes": by, "gbs": by / (us * 1e-6) / 1e9 if us > 0 else None}
        if name in ("norm1", "norm2") and os.environ.get("LRR_NO_NORM_SKIP") is None and os.environ.get("LRR_NO_PROOF") is None:
            # steady state: every query is kept by its certificate, the match set — all the normalisation depends on — is unchanged,
            # and the launch leaves at entry (Scal::norm_dirty): no bytes are moved, the time is the launch itself
            per_phase[name].update({"algorithmic_bytes": 0, "gbs": None, "note": "leaves at entry while no match changed (steady state)"})
    d = per_phase.get(dom, {"gbs": 0, "us_per_launch": 0, "algorithmic_bytes": 0})
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (bench.py cannot run under a profiler)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        names = {"match": ["k_match_cert", "k_match_tile", "k_match_hard"], "accumulate": ["k_accumulate"], "weight": ["k_weight"]}.get(dom, [])
        if names and all(n in tj["dram_bytes_per_launch"] for n in names):
            traffic = sum(tj["dram_bytes_per_launch"][n] for n in names); traffic_src = tj["source"]
    except Exception:
        pass
    it_bytes = algorithmic_bytes_iteration(M, N, M, N)
    ms_step = total_ms_max / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_fixed": N, "n_moving": M, "seed": SEED, "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2": "flushed (256 MiB write) between timed iterations; working set ~30 MB is L2-resident otherwise",
                   "ramp_iterations": ramp_iters, "ramp_converged": bool(ramp_conv), "settle_iterations": settle_iters},
        "value_l2_warm": value_warm, "ms_per_step_l2_warm": warm_ms_max / args.steps,
        "roofline": {"bound": "hbm", "kernel": "k_iteration (fused: match, normalise, weight, order statistics, accumulate, solve, residuals)" if dom == "fused_iteration" else ("k_match_cert + k_match_tile + k_match_hard (exact 1-NN: neighbour certificates, shared-memory-staged cells for the queries they do not cover, queued far queries)" if dom == "match" else "k_" + dom), "achieved": d["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": (d["gbs"] / peak) if d["gbs"] else None, "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                     "us_per_launch": d["us_per_launch"], "algorithmic_bytes_per_launch": d["algorithmic_bytes"]},
        "roofline_iteration": {"algorithmic_bytes": it_bytes, "achieved_gbs": it_bytes / (ms_step * 1e-3) / 1e9,
                               "frac": it_bytes / (ms_step * 1e-3) / 1e9 / peak,
                               "achieved_gbs_l2_warm": it_bytes / (warm_ms_max / args.steps * 1e-3) / 1e9},
        "kernels": per_phase,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "iterations": e2e_iters, "seconds": float(te[0]),
                "what": "per step: upload both volumes (cdcl_feature_ICP arrays: locations, projectors, strengths, shapes, directions) from pinned host memory (grid build), construct estimator, set_transform, "
                        "ramped run to convergence (<=70 one_iteration calls), read back transform+covariance"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "registrations": {"value": len(bjobs) / batch_s, "unit": "registrations/s", "jobs": int(len(bjobs)), "seconds": batch_s,
                          "verified": verified_count(full), "verified_per_s": verified_count(full) / batch_s,
                          "rejected": int((full["status"] == 1).sum()), "mean_iterations": float(full["iterations"][full["status"] == 0].mean()),
                          "what": f"C4 bounded: {BATCH_LOCATIONS_PER_GPU} locations x {BATCH_INITS} initialisations per GPU on one resident "
                                  "100k/100k pair, 50 mm ROI, rrl symmetric <3,12>, <= 70 iterations, jobs sharded over the ranks, one result gather"},
    }
    if world == 1 and not args.only_iterations and os.environ.get("LRR_NO_PROOF") is None:
        # the same timed loop with the neighbour certificates switched off (every finest-level query searched in every iteration):
        # what an iteration costs when nothing can be kept, e.g. the first iteration at the finest level
        try:
            env = dict(os.environ, LRR_NO_PROOF="1")
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--only-iterations", "--steps", str(args.steps), "--warmup", str(args.warmup)],
                                 env=env, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
            sub = json.loads(out[-1])
            line["certificates_off"] = {"value": sub["value"], "unit": UNIT, "ms_per_step": sub["ms_per_step"], "value_l2_warm": sub["value_l2_warm"],
                                        "match_us": sub["kernels"]["match"]["us_per_launch"],
                                        "what": "LRR_NO_PROOF=1: same loop, every query of both directions searched in every iteration (no neighbour certificates)"}
        except Exception as ex:   # noqa: BLE001
            line["certificates_off"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu and not args.only_iterations:
        r = cpu_measure(pair, 6, 1)
        line["cpu_baseline"] = {"value": r["finest_iters_per_s"], "unit": UNIT, "cores": 1, "kind": r["kind"],
                                "sample": "full ramped registration + 6 one_iteration calls at the finest level, single thread "
                                          "(the reference's execution model); " + CPU_KIND_NOTE[r["kind"]],
                                "registration_iters_per_s": r["registration_iters_per_s"], "registration_s": r["registration_s"],
                                "host_cores_available": os.cpu_count()}
        line["cpu_baseline"]["registrations"] = cpu_batch_measure(bpair, bjobs)
    else:
        line["cpu_baseline"] = None
    line["final_A00"] = float(final["A"][0, 0])
    sys.stdout.flush()
    os.write(out_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--only-iterations", action="store_true", help="skip the e2e / batch / cpu legs (short run for ncu launch lists)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
