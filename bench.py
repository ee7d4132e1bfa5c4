#!/usr/bin/env python
"""bench.py -- residual-sample CG iterations/s of the MGVI Fisher-CG path on B200 (BASELINE.json).

One "step" = one batched solve call of the hot path (sampler right-hand side + ITERS lock-step
CG iterations of  M x = J' sqrt(F) n1 + eta  over every local residual sample) on the workload
BASELINE.json's target is quoted on: the 2048 x 2048 correlated field, Normal likelihood,
32 antithetic residual samples (configs[2]), sharded by sample over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU restatement of the reference path, host cores

`value`  : sample-iterations/s with every input already resident in HBM (CUDA events on the
           library's own stream, max over ranks).
`e2e`    : the same metric through the host-buffer C ABI call a user makes
           (mgvib200_sample_residuals_cg, reference-default stopping rule, pinned host buffers,
           H2D of the draws and D2H of the residuals inside the timed region).
`roofline`: the dominant kernel of the step; algorithmic bytes per SURVEY.md section 8d.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "residual-sample CG iterations/s (2048x2048 correlated field, Normal, 32 residual samples)"
UNIT = "sample-iterations/s"


def metric_name(args, cfg=None):
    if args.workload == "c3":
        return METRIC
    dims, like, n = WORKLOADS[args.workload]
    return "residual-sample CG iterations/s (%s correlated field, %s, %d residual samples)" % (
        "x".join(str(d) for d in dims), "Normal" if like == "LIKE_NORMAL" else "Poisson", n)

WORKLOADS = {
    # name: (dims, likelihood name, n residuals)
    "c3": ((2048, 2048), "LIKE_NORMAL", 32),
    "c2": ((65536,), "LIKE_POISSON", 16),
    "c5": ((256, 256, 256), "LIKE_POISSON", 64),
    "small": ((256, 256), "LIKE_NORMAL", 8),
    "np2": ((1000, 1000), "LIKE_NORMAL", 8),      # axis lengths that are not a power of two (Bluestein passes)
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--iters-per-step", type=int, default=25)
    ap.add_argument("--field-std", type=float, default=0.5)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-maxiters", type=int, default=0, help="0 = reference default (n_theta)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-step", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-samples", type=int, default=0, help="CPU arm: residuals per step; 0 = one per host core")
    ap.add_argument("--ref-iters", type=int, default=2)
    ap.add_argument("--no-configs", action="store_true", help="skip the short C1/C2/C4/C5 measurements")
    ap.add_argument("--n-residuals", type=int, default=0,
                    help="override the workload's residual count (tuning: 4 on one GPU is the per-GPU share of C3 on 8 GPUs)")
    return ap.parse_args()


def make_config(args, syn):
    dims, like, n = WORKLOADS[args.workload]
    if getattr(args, "n_residuals", 0) > 0:
        n = args.n_residuals
    cfg = syn.field_config(dims, getattr(syn, like), n, field_std=args.field_std, draws=False)
    return cfg


def draws_for(cfg, idx):
    """Standard-normal draws of residual `idx` (global index), independent of the sharding:
    n1 (n_lambda) then eta (n_theta), as the reference consumes them."""
    rng = np.random.default_rng(42 + idx)
    return rng.standard_normal(cfg["n_lambda"]), rng.standard_normal(cfg["n_theta"])


def algorithmic_bytes(N, d):
    """SURVEY.md section 8d, per residual sample.  The step's kernels and the share of
    B_cg = 32 N (d+1) + 72 N each one carries (the three FFT kernels implement the 2d axis passes
    of the pass model; the two axis-0 passes are fused into ONE kernel here)."""
    first = 16 * N + 8 * N + 24 * N            # first axis pass + A read + fused p = r + beta p
    mid_each = 16 * N                           # every other axis pass
    fused0 = 2 * 16 * N + 8 * N                 # both axis-0 passes + w read
    last = 16 * N + 8 * N + 8 * N               # last axis pass + A read + v read (dot fused)
    update = 48 * N                             # x += a p ; r -= a Ap ; r.r
    return dict(first=first, mid_each=mid_each, fused0=fused0, last=last, update=update,
                B_mv=32 * N * (d + 1), B_cg=32 * N * (d + 1) + 72 * N)


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region: an NVML polling thread
    (every ~5 ms; nvidia-smi -lms cannot sample a 50 ms region), summarised as the recipe's
    clocks line."""

    def __init__(self, gpu_index):
        import threading
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(sm), float(pw), int(rs)))
            except Exception:           # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        sm = [x[0] for x in self.samples]
        pw = [x[1] for x in self.samples]
        med_p = float(np.median(pw))
        load = [s_ for s_, p_ in zip(sm, pw) if p_ >= med_p] or sm
        reasons = sorted(k for k, bit in names.items() if any(x[2] & bit for x in self.samples))
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/), timed on the host cores.
# The reference parallelises over residual samples (`Threads.@threads`, src/residual_samplers.jl:85-92):
# one residual per core.  Here: one PROCESS per core (immune to torchrun's OMP_NUM_THREADS=1 and to the
# GIL), each solving its own residual with single-threaded FFTs; a step is one lock-step round of all
# workers (barrier at the start of every round), timed from the first start to the last finish.
# ------------------------------------------------------------------------------------------------
def _cpu_worker(widx, workload, field_std, sample_ids, iters, rounds, barrier, queue):
    try:
        os.environ["OMP_NUM_THREADS"] = "1"
        os.environ["MKL_NUM_THREADS"] = "1"
        os.environ["OPENBLAS_NUM_THREADS"] = "1"
        sys.path.insert(0, ROOT)
        import oracle as O
        g = __import__("__graft_entry__")
        syn = g.load_package().synthetic
        dims, like, n = WORKLOADS[workload]
        cfg = syn.field_config(dims, getattr(syn, like), n, field_std=field_std, draws=False)
        O.mgvi_oracle.set_fft_workers(1)
        om = O.FieldModel(cfg["dims"], cfg["amplitude"], cfg["scale"], cfg["data"], offset=cfg["offset"],
                          link=cfg["link"], likelihood=cfg["likelihood"], sigma=cfg["sigma"], layout=cfg["layout"])
        Z1 = np.empty((cfg["n_lambda"], len(sample_ids)))
        Z2 = np.empty((cfg["n_theta"], len(sample_ids)))
        for j, i in enumerate(sample_ids):
            Z1[:, j], Z2[:, j] = draws_for(cfg, i)
        spans = []
        for _ in range(rounds):
            barrier.wait()
            t0 = time.time()
            _R, it, _ = O.sample_residuals_cg(om, cfg["center"], Z1, Z2, atol=0.0, rtol=0.0, itmax=iters)
            t1 = time.time()
            assert int(it.sum()) == len(sample_ids) * iters
            spans.append((t0, t1))
        queue.put((widx, spans, None))
    except BaseException as e:      # noqa: BLE001
        try:
            barrier.abort()
        except Exception:           # noqa: BLE001
            pass
        queue.put((widx, [], repr(e)))


def cpu_reference_rate(args, n_samples, iters, repeats=1, warmup=0):
    """sample-iterations/s of the CPU restatement of MGVI.jl's CG sampler (oracle/) on a bounded sample of
    the workload: `n_samples` residuals (default: one per host core) x `iters` CG iterations per step, one
    worker process per core.  Returns (rate, seconds per step list, workers)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_tot = WORKLOADS[args.workload][2]
    if n_samples <= 0:
        n_samples = max(1, min(cores, n_tot))
    workers = max(1, min(cores, n_samples))
    ids = [list(range(w, n_samples, workers)) for w in range(workers)]
    ctxm = mp.get_context("spawn")          # the GPU arm has CUDA initialised: never fork it
    barrier = ctxm.Barrier(workers)
    queue = ctxm.Queue()
    rounds = warmup + repeats
    procs = [ctxm.Process(target=_cpu_worker, args=(w, args.workload, args.field_std, ids[w], iters, rounds, barrier, queue))
             for w in range(workers)]
    for p_ in procs:
        p_.start()
    got = [queue.get() for _ in procs]
    for p_ in procs:
        p_.join()
    errs = [e for _, _, e in got if e]
    if errs:
        raise RuntimeError("CPU reference worker failed: " + errs[0])
    times = []
    for r in range(warmup, rounds):
        t0 = min(sp[r][0] for _, sp, _ in got)
        t1 = max(sp[r][1] for _, sp, _ in got)
        times.append(t1 - t0)
    rate = n_samples * iters * len(times) / sum(times)
    return rate, times, workers, n_samples


CPU_NOTE = ("CPU restatement of MGVI.jl v0.4.5 hot path (oracle/: NumPy + scipy.fft), one residual sample per worker "
            "process, one process per host core (the reference's Threads.@threads over samples, "
            "src/residual_samplers.jl:85-92); Julia is not installed in this image")


def run_reference(args, rank, world):
    if rank != 0:
        return
    g = __import__("__graft_entry__")
    syn = g.load_package().synthetic
    cfg = make_config(args, syn)
    K, W = args.steps, max(args.warmup, 1)
    rate, times, workers, ns = cpu_reference_rate(args, args.ref_samples, args.ref_iters, repeats=K, warmup=W)
    sample = "%d residual samples x %d CG iterations per step (%s, full field size), %d worker processes" % (
        ns, args.ref_iters, args.workload, workers)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, cfg, cfg["n_residuals"] // max(world, 1), world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(), "note": CPU_NOTE},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def bench_config(args, cfg, n_loc, world):
    """The `config` object both arms print (same keys, so the driver can compare them)."""
    N, d = cfg["n_theta"], len(cfg["dims"])
    B_cg = 32 * N * (d + 1) + 72 * N
    return {"workload": workload_name(args, cfg), "iters_per_step": args.iters_per_step, "field_std": args.field_std,
            "samples_per_gpu": n_loc, "parallelism": "sample-sharded x%d" % world,
            "l2": "inputs larger than L2 (%.1f GB streamed per step per GPU)" % (B_cg * n_loc * args.iters_per_step / 1e9),
            "timing": "CUDA events on the library stream, max over ranks"}


def workload_name(args, cfg):
    dims = "x".join(str(d) for d in cfg["dims"])
    like = "Normal(sigma=%.2g)" % cfg["sigma"] if cfg["likelihood"] == 0 else "Poisson"
    return "%s: %s correlated field, %s, exp link, %d residual samples" % (args.workload, dims, like,
                                                                         cfg["n_residuals"])


# ------------------------------------------------------------------------------------------------
# Short measurements of the other BASELINE.json configurations (the `configs` block of the line)
# ------------------------------------------------------------------------------------------------
def _prof(ctx, fn):
    lib = ctx.lib
    lib.profile_begin(ctx.handle)
    out = fn()
    tags = (C.c_int32 * 64)(); ms = (C.c_double * 64)(); cnt = (C.c_int64 * 64)()
    n = lib.profile_end(ctx.handle, 64, tags, ms, cnt)
    return out, {int(tags[i]): (float(ms[i]), int(cnt[i])) for i in range(max(n, 0))}


def quick_field(pkg, ctx, workload, n_loc, first_sample, iters, steps, warmup, hbm_peak, field_std):
    """`steps` timed solves of `iters` lock-step CG iterations over n_loc residual samples, inputs resident."""
    syn, lib = pkg.synthetic, ctx.lib
    dims, like, n_tot = WORKLOADS[workload]
    cfg = syn.field_config(dims, getattr(syn, like), n_tot, field_std=field_std, draws=False)
    N, nl, d = cfg["n_theta"], cfg["n_lambda"], len(dims)
    model = pkg.CorrelatedFieldModel.from_config(cfg, ctx)
    Z1h, Z2h = np.empty((n_loc, nl)), np.empty((n_loc, N))
    for j in range(n_loc):
        Z1h[j], Z2h[j] = draws_for(cfg, first_sample + j)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    center = np.ascontiguousarray(cfg["center"])
    bufs = [lib.dev_alloc(ctx.handle, b) for b in (8 * N, Z1h.nbytes, Z2h.nbytes, 8 * N * n_loc)]
    c_dev, z1_dev, z2_dev, r_dev = bufs
    ctx.check(lib.memcpy_h2d(ctx.handle, c_dev, vp(center), 8 * N))
    ctx.check(lib.memcpy_h2d(ctx.handle, z1_dev, vp(Z1h), Z1h.nbytes))
    ctx.check(lib.memcpy_h2d(ctx.handle, z2_dev, vp(Z2h), Z2h.nbytes))
    ctx.check(lib.linearize_dev(model.handle, c_dev))
    it = np.zeros(n_loc, dtype=np.int64)
    rn = np.zeros(n_loc)
    fixed = pkg.capi.CGOpts(0.0, 0.0, iters)

    def step():
        ctx.check(lib.sample_residuals_cg_dev(model.handle, z1_dev, z2_dev, n_loc, C.byref(fixed), r_dev,
                                              it.ctypes.data_as(pkg.capi.c_int64_p), pkg.capi.ptr(rn)))
        return ctx.last_timing()["sampler_ms"]

    for _ in range(warmup):
        step()
    ms = sum(step() for _ in range(steps))
    per_iter = ms / (steps * iters)
    B_cg = 32 * N * (d + 1) + 72 * N
    out = {"workload": workload_name(type("A", (), {"workload": workload})(), dict(cfg, n_residuals=n_tot)),
           "samples_on_this_gpu": n_loc, "iters_per_step": iters, "steps": steps,
           "value": n_loc * iters * steps / (ms * 1e-3), "unit": UNIT, "lockstep_iteration_ms": per_iter,
           "B_cg_bytes_per_sample_iteration": B_cg,
           "frac_of_hbm_peak": B_cg * n_loc / (per_iter * 1e-3) / 1e9 / hbm_peak}
    if N * 8 * 4 * n_loc < 100e6:
        out["note"] = "working set fits L2: latency-bound, the HBM fraction is not a roofline statement"
    for b in bufs:
        lib.dev_free(ctx.handle, b)
    model.close()
    return out


def quick_poly(pkg, ctx):
    """C1: the polynomial fit on the 1000-point grids, 8 residuals: whole mgvi_step through the host API."""
    cfg = pkg.synthetic.polyfit_config(8, reference_grid=False)
    m = pkg.model_from_config(cfg, ctx)
    conf = pkg.MGVIConfig()
    pkg.mgvi_step(m, cfg["data"], 8, cfg["center"], conf, ctx, draws=(cfg["Z1"], cfg["Z2"]))
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        res, _c = pkg.mgvi_step(m, cfg["data"], 8, cfg["center"], conf, ctx, draws=(cfg["Z1"], cfg["Z2"]))
    dt = (time.perf_counter() - t0) / reps
    m.close()
    return {"workload": "c1: polynomial fit, Normal, 5 parameters, 1000 data points, 8 residual samples (CG sampler + NewtonCG)",
            "mgvi_step_ms": dt * 1e3, "steps_per_s": 1.0 / dt, "mnlp": res.mnlp, "note": "latency-bound: time only"}


def quick_dense(pkg, ctx, torch):
    """C4: dense J 16384 x 4096, MatrixInversion sampler.  The FP64 tensor-pipe denominator is measured here:
    cuBLAS dgemm (torch.mm, float64) of the same full product, and cuSOLVER's Cholesky of the same matrix."""
    mm, k, n = 16384, 4096, 32
    cfg = pkg.synthetic.dense_linear_config(mm, k, n)
    model, prof_c = _prof(ctx, lambda: pkg.model_from_config(cfg, ctx))
    syrk_ms = prof_c.get(3001, (0.0, 0))[0]
    pkg.ResidualSampler(model, cfg["center"], pkg.MatrixInversion(), ctx)
    Z = np.asfortranarray(cfg["Z"])
    R = np.empty((k, n), order="F")

    def run():
        ctx.check(ctx.lib.sample_residuals_dense(model.handle, pkg.capi.ptr(Z), n, pkg.capi.ptr(R), None))
    # factor + solve on a fresh model minus solve only, wall clock of the synchronous C-ABI call (per-kernel events
    # would add their host overhead to ~200 short launches), best of 3
    t_first, t_solve = [], []
    for _ in range(3):
        m2 = pkg.model_from_config(cfg, ctx)
        pkg.ResidualSampler(m2, cfg["center"], pkg.MatrixInversion(), ctx)
        t0 = time.perf_counter()
        ctx.check(ctx.lib.sample_residuals_dense(m2.handle, pkg.capi.ptr(Z), n, pkg.capi.ptr(R), None))
        t_first.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.check(ctx.lib.sample_residuals_dense(m2.handle, pkg.capi.ptr(Z), n, pkg.capi.ptr(R), None))
        t_solve.append(time.perf_counter() - t0)
        m2.close()
    chol_ms = (min(t_first) - min(t_solve)) * 1e3
    solve_ms = min(t_solve) * 1e3
    run()
    conf = pkg.MGVIConfig(linsolver=pkg.MatrixInversion())
    pkg.mgvi_step(model, cfg["data"], n, cfg["center"], conf, ctx, draws=cfg["Z"])
    t0 = time.perf_counter()
    res, _c = pkg.mgvi_step(model, cfg["data"], n, cfg["center"], conf, ctx, draws=cfg["Z"])
    step_s = time.perf_counter() - t0
    model.close()
    # denominators (library code, measured on this box, never on the product path)
    Jt = torch.from_numpy(cfg["J"]).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        Mt = torch.mm(Jt.T, Jt)
    torch.cuda.synchronize()
    e0.record()
    reps = 3
    for _ in range(reps):
        Mt = torch.mm(Jt.T, Jt)
    e1.record(); torch.cuda.synchronize()
    dgemm_ms = e0.elapsed_time(e1) / reps
    Mt = Mt * 4.0 + torch.eye(k, dtype=torch.float64, device="cuda")
    torch.linalg.cholesky(Mt)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        torch.linalg.cholesky(Mt)
    e1.record(); torch.cuda.synchronize()
    potrf_ms = e0.elapsed_time(e1) / reps
    del Jt, Mt
    torch.cuda.empty_cache()
    full = 2.0 * mm * k * k
    T = (k + 127) // 128
    executed = 2.0 * mm * 128 * 128 * (T * (T + 1) // 2)        # upper block tiles only
    dgemm_tf = full / (dgemm_ms * 1e-3) / 1e12
    ours_tf = executed / (syrk_ms * 1e-3) / 1e12 if syrk_ms else None
    return {"workload": "c4: dense J %d x %d, Normal(0.5), %d residuals, MatrixInversion sampler" % (mm, k, n),
            "syrk_ms": syrk_ms, "syrk_tflops_executed": ours_tf,
            "syrk_full_product_equivalent_tflops": full / (syrk_ms * 1e-3) / 1e12 if syrk_ms else None,
            "cublas_dgemm_full_product_ms": dgemm_ms, "cublas_dgemm_tflops": dgemm_tf,
            "roofline": {"bound": "tensor", "kernel": "k_dmma_gemm (J'FJ, upper block tiles, mma.sync m8n8k4 f64)",
                         "achieved": ours_tf, "peak": dgemm_tf, "unit": "TFLOP/s",
                         "frac": (ours_tf / dgemm_tf) if ours_tf else None,
                         "peak_source": "cuBLAS dgemm 16384 x 4096 x 4096 (torch.mm, float64) measured in this run"},
            "cholesky_ms": chol_ms, "cusolver_potrf_ms": potrf_ms,
            "cholesky_tflops": (k ** 3 / 3.0) / (chol_ms * 1e-3) / 1e12 if chol_ms > 0 else None,
            "sample_ms_after_factor": solve_ms, "mgvi_step_s": step_s, "mnlp": res.mnlp,
            "timing": "cholesky / sample: wall clock of the synchronous C-ABI call (factor + solve minus solve only), best of 3"}


def c5_step(pkg, ctx, rank, world, per_gpu, sampler_maxiters, torch, dist, field_std):
    """BASELINE.json configs[4] at `per_gpu` residual samples per GPU (8 per GPU = 64 on 8 GPUs; weak scaling in
    N): one mgvi_step through the host-buffer C ABI -- sampler capped at `sampler_maxiters` lock-step iterations
    (a converged C5 sampler is thousands of iterations and has no communication), then the complete default
    NewtonCG phase, whose KL value / gradient / mean-weight sums cross the GPUs."""
    syn, lib = pkg.synthetic, ctx.lib
    dims, like, _n = WORKLOADS["c5"]
    n_tot = per_gpu * world
    cfg = syn.field_config(dims, getattr(syn, like), n_tot, field_std=field_std, draws=False)
    N, nl = cfg["n_theta"], cfg["n_lambda"]
    model = pkg.CorrelatedFieldModel.from_config(cfg, ctx)
    Z1h, Z2h = np.empty((per_gpu, nl)), np.empty((per_gpu, N))
    for j in range(per_gpu):
        Z1h[j], Z2h[j] = draws_for(cfg, rank * per_gpu + j)
    samples = np.empty((2 * per_gpu, N))
    c_out = np.empty(N)
    mnlp = C.c_double()
    trace = pkg.capi.NewtonCGTrace()
    iters = np.zeros(per_gpu, dtype=np.int64)
    cg = pkg.capi.CGOpts(pkg.api.SQRT_EPS, pkg.api.SQRT_EPS, int(sampler_maxiters))
    opts = pkg.capi.NewtonCGOpts(0.1, 4, 5, 50)
    center = np.ascontiguousarray(cfg["center"])

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sync()
    t0 = time.perf_counter()
    ctx.check(lib.step(model.handle, pkg.capi.ptr(center), pkg.capi.ptr(Z1h), pkg.capi.ptr(Z2h), per_gpu, 0,
                       C.byref(cg), C.byref(opts), pkg.capi.ptr(samples), pkg.capi.ptr(c_out), C.byref(mnlp),
                       iters.ctypes.data_as(pkg.capi.c_int64_p), C.byref(trace)))
    sync()
    dt = time.perf_counter() - t0
    tm = ctx.last_timing()
    red = torch.tensor([dt, tm["sampler_ms"], tm["newton_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    model.close()
    tr = trace.as_dict()
    return {"workload": "c5: 256x256x256 correlated field, Poisson, exp link, %d residual samples (%d per GPU)" % (n_tot, per_gpu),
            "scaling": "weak", "n_gpus": world, "seconds": float(red[0]), "steps_per_s": 1.0 / float(red[0]),
            "sampler_ms": float(red[1]), "sampler_lockstep_iterations_cap": int(sampler_maxiters),
            "newton_ms": float(red[2]), "newton_phases_per_s": 1e3 / float(red[2]), "mnlp": float(mnlp.value),
            "kl_evaluations": tr["f_calls"], "kl_gradients": tr["g_calls"], "newton_cg_updates": tr["cg_updates"],
            "collectives": "all-gather of canonical chunk partials per KL value / gradient / mean weight (NCCL)" if world > 1 else "none (1 GPU)",
            "h2d_bytes": int(Z1h.nbytes + Z2h.nbytes + center.nbytes), "d2h_bytes": int(samples.nbytes + c_out.nbytes)}


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line: str):
    """The driver reads ONE JSON line from stdout; libraries (NCCL prints its version banner on
    stdout) are kept off it: fd 1 is pointed at stderr for the whole run and the result line goes
    to the saved descriptor."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    g = __import__("__graft_entry__")
    pkg = g.load_package()
    syn = pkg.synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = make_config(args, syn)
    N, n_lambda, n_tot = cfg["n_theta"], cfg["n_lambda"], cfg["n_residuals"]
    d = len(cfg["dims"])
    if n_tot % world:
        raise SystemExit("residual count %d does not divide over %d ranks" % (n_tot, world))
    n_loc = n_tot // world
    mine = range(rank * n_loc, (rank + 1) * n_loc)     # contiguous column shard

    ctx = pkg.MGVIContext(device=local_rank)
    lib = ctx.lib
    if world > 1:
        box = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.init_comm(rank, world, box[0])
    model = pkg.CorrelatedFieldModel.from_config(cfg, ctx)

    # ---- host draws in pinned memory (one contiguous column per residual) ----------------------
    Z1p = torch.empty((n_loc, n_lambda), dtype=torch.float64, pin_memory=True)
    Z2p = torch.empty((n_loc, N), dtype=torch.float64, pin_memory=True)
    Rp = torch.empty((n_loc, N), dtype=torch.float64, pin_memory=True)
    Z1h, Z2h, Rh = Z1p.numpy(), Z2p.numpy(), Rp.numpy()
    for j, i in enumerate(mine):
        Z1h[j], Z2h[j] = draws_for(cfg, i)
    center = np.ascontiguousarray(cfg["center"])

    def dev(nbytes):
        p = lib.dev_alloc(ctx.handle, nbytes)
        if not p:
            raise SystemExit("device allocation failed")
        return p

    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    c_dev, z1_dev, z2_dev, r_dev = dev(8 * N), dev(Z1h.nbytes), dev(Z2h.nbytes), dev(Rh.nbytes)
    ctx.check(lib.memcpy_h2d(ctx.handle, c_dev, vp(center), 8 * N))
    ctx.check(lib.memcpy_h2d(ctx.handle, z1_dev, vp(Z1h), Z1h.nbytes))
    ctx.check(lib.memcpy_h2d(ctx.handle, z2_dev, vp(Z2h), Z2h.nbytes))
    ctx.check(lib.linearize_dev(model.handle, c_dev))

    iters = np.zeros(n_loc, dtype=np.int64)
    rn = np.zeros(n_loc)
    ITERS = args.iters_per_step
    fixed = pkg.capi.CGOpts(0.0, 0.0, ITERS)     # never converges early: exactly ITERS iterations

    def step_dev():
        ctx.check(lib.sample_residuals_cg_dev(model.handle, z1_dev, z2_dev, n_loc, C.byref(fixed), r_dev,
                                              iters.ctypes.data_as(pkg.capi.c_int64_p), pkg.capi.ptr(rn)))
        assert int(iters.sum()) == n_loc * ITERS, iters
        return ctx.last_timing()["sampler_ms"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    for _ in range(W):
        step_dev()
    barrier()
    clocks = ClockSampler(local_rank)
    ctx.launch_count(reset=True)
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(K):
        dev_ms += step_dev()
    barrier()
    wall_s = time.perf_counter() - t0
    launches = ctx.launch_count()
    clk = clocks.stop()
    # per-kernel breakdown: a separate, untimed pass with CUDA events around every launch (the timed steps
    # above replay CUDA graphs; events between graph nodes would serialise them)
    lib.profile_begin(ctx.handle)
    for _ in range(min(K, 2)):
        step_dev()
    tags = (C.c_int32 * 64)()
    tms = (C.c_double * 64)()
    tcnt = (C.c_int64 * 64)()
    ntag = lib.profile_end(ctx.handle, 64, tags, tms, tcnt)
    tt = torch.tensor([dev_ms, wall_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(tt[0]), float(tt[1])
    value = n_tot * ITERS * K / (dev_ms_max * 1e-3)

    # ---- per-kernel breakdown and roofline -----------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    ab = algorithmic_bytes(N, d)
    names = {200: ("first_pass(p=r+beta*p, A.p, H_last)", ab["first"]),
             1: ("axis0_fused(H, w., H)", ab["fused0"]),
             0: ("axis_pass(H)", ab["mid_each"]),
             10: ("last_pass(H_last, c^2 A., +p, p.Ap)", ab["last"]),
             211: ("fused_1d(p update, A., H, w., H, A., +p, p.Ap)", ab["first"] + ab["last"] - 16 * N),
             1000: ("cg_update(x+=a p, r-=a Ap, r.r)", ab["update"]),
             # deferred solution update: the update kernel streams r only (half of the update's
             # share), x += a p rides in the next iteration's axis0_fused (credited there below), one flush at the end
             1014: ("cg_update_r(r-=a Ap, r.r)", ab["update"] // 2),
             1015: ("x_flush(x+=a p, pending samples)", 24 * N),
             400: ("rhs_first(q.n1, H)", 16 * N + 16 * N), 30: ("rhs_last(H, cA., +eta, b.b)", 16 * N + 40 * N),
             430: ("rhs_1d", 16 * N + 40 * N)}
    if any(int(tags[i]) == 1014 for i in range(max(ntag, 0))):
        names[1] = ("axis0_fused(H, w., H; x+=a p of the previous update)", ab["fused0"] + ab["update"] // 2)
    kern = []
    for i in range(max(ntag, 0)):
        tag, ms, cnt = int(tags[i]), float(tms[i]), int(tcnt[i])
        nm, per_sample = names.get(tag, ("tag%d" % tag, None))
        avg_ms = ms / max(cnt, 1)
        e = {"kernel": nm, "tag": tag, "launches": cnt, "avg_ms": avg_ms, "share": None}
        if per_sample is not None:
            e["algorithmic_bytes_per_launch"] = per_sample * n_loc
            e["algorithmic_gbs"] = per_sample * n_loc / (avg_ms * 1e-3) / 1e9
        kern.append(e)
    # bytes each CG kernel really moves per sample (ncu: nothing is re-read), next to the
    # algorithmic share of SURVEY's pass model
    moved_k = {200: 32 * N, 1: 16 * N, 0: 16 * N, 10: 24 * N, 1000: 48 * N, 1014: 24 * N}
    if any(k["tag"] == 1014 for k in kern):
        moved_k[1] = 40 * N
    for k in kern:
        if k["tag"] in moved_k and d >= 2:
            k["moved_bytes_per_launch"] = moved_k[k["tag"]] * n_loc
            k["moved_gbs"] = moved_k[k["tag"]] * n_loc / (k["avg_ms"] * 1e-3) / 1e9
    tot_ms = sum(k["avg_ms"] * k["launches"] for k in kern) or 1.0
    for k in kern:
        k["share"] = k["avg_ms"] * k["launches"] / tot_ms
    kern.sort(key=lambda k: -k["share"])
    dom = kern[0] if kern else None
    roofline = None
    if dom and "algorithmic_gbs" in dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["algorithmic_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom["algorithmic_gbs"] / hbm_peak,
                    "traffic": None,
                    "traffic_note": "not measurable inside a timed run; the ncu --set full capture of this command is "
                                    "profiles/r2b_ncu_full_c3_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                    "peak_source": peak_src, "share_of_step": dom["share"],
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"]}
        if "moved_gbs" in dom:
            # the same kernel against the bytes it really moves (the algorithmic model counts the fused axis-0
            # transforms as two passes, so `frac` can exceed 1)
            roofline["moved_bytes_per_launch"] = dom["moved_bytes_per_launch"]
            roofline["frac_bytes_moved"] = dom["moved_gbs"] / hbm_peak
    # the whole Fisher-CG iteration against B_cg: what north_star's 60 % target is about
    per_iter_ms = dev_ms_max / (K * ITERS)
    cg_path = {"B_cg_bytes_per_sample_iteration": ab["B_cg"], "lockstep_iteration_ms": per_iter_ms,
               "achieved_gbs": ab["B_cg"] * n_loc / (per_iter_ms * 1e-3) / 1e9}
    cg_path["frac_of_hbm_peak"] = cg_path["achieved_gbs"] / hbm_peak
    if d >= 2:
        # bytes the kernels actually move per sample-iteration: 32N + 16N (+ 32N per middle
        # axis) + 24N + 48N -- A and w are shared by all samples and stay in L2
        moved = (120 + 32 * (d - 2)) * N
        cg_path["bytes_moved_per_sample_iteration"] = moved
        cg_path["frac_of_hbm_peak_bytes_moved"] = moved * n_loc / (per_iter_ms * 1e-3) / 1e9 / hbm_peak

    out = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, cfg, n_loc, world),
        "wall_ms_per_step": wall_ms_max / K,
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cg_path": cg_path,
        "kernels": kern,
    }

    # ---- end to end through the host-buffer C ABI --------------------------------------------
    if not args.no_e2e:
        dflt = pkg.capi.CGOpts(pkg.api.SQRT_EPS, pkg.api.SQRT_EPS, int(args.e2e_maxiters))
        it2 = np.zeros(n_loc, dtype=np.int64)

        def step_host():
            ctx.check(lib.sample_residuals_cg(model.handle, pkg.capi.ptr(Z1h), pkg.capi.ptr(Z2h), n_loc,
                                              C.byref(dflt), pkg.capi.ptr(Rh), it2.ctypes.data_as(pkg.capi.c_int64_p),
                                              pkg.capi.ptr(rn)))
            return int(it2.sum())

        step_host()
        barrier()
        t0 = time.perf_counter()
        tot_it = 0
        for _ in range(args.e2e_steps):
            tot_it += step_host()
        barrier()
        dt = time.perf_counter() - t0
        red = torch.tensor([dt, float(tot_it)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = red.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(red, op=dist.ReduceOp.SUM)
            dt, tot_it = float(mx[0]), float(red[1])
        out["e2e"] = {"value": tot_it / dt, "unit": UNIT, "h2d_bytes_per_step": int(Z1h.nbytes + Z2h.nbytes),
                      "d2h_bytes_per_step": int(Rh.nbytes), "seconds_per_step": dt / args.e2e_steps,
                      "sample_iterations_per_step": tot_it / args.e2e_steps,
                      "lockstep_iterations": int(ctx.last_timing()["lockstep_iters"]),
                      "stopping_rule": "reference default: abstol = reltol = sqrt(eps)",
                      "api": "mgvib200_sample_residuals_cg (host buffers, pinned)"}

    # ---- one full MGVI step (sampler + Newton-CG), host buffers --------------------------------
    if not args.no_full_step:
        samples = np.empty((2 * n_loc, N))
        c_out = np.empty(N)
        mnlp = C.c_double()
        trace = pkg.capi.NewtonCGTrace()
        dflt = pkg.capi.CGOpts(pkg.api.SQRT_EPS, pkg.api.SQRT_EPS, int(args.e2e_maxiters))
        opts = pkg.capi.NewtonCGOpts(0.1, 4, 5, 50)
        barrier()
        t0 = time.perf_counter()
        ctx.check(lib.step(model.handle, pkg.capi.ptr(center), pkg.capi.ptr(Z1h), pkg.capi.ptr(Z2h), n_loc, 0,
                           C.byref(dflt), C.byref(opts), pkg.capi.ptr(samples), pkg.capi.ptr(c_out), C.byref(mnlp),
                           iters.ctypes.data_as(pkg.capi.c_int64_p), C.byref(trace)))
        barrier()
        dt = time.perf_counter() - t0
        tm = ctx.last_timing()
        out["mgvi_step"] = {"seconds": dt, "steps_per_s": 1.0 / dt, "mnlp": float(mnlp.value),
                            "sampler_ms": tm["sampler_ms"], "newton_ms": tm["newton_ms"],
                            "sampler_iterations": int(iters.sum()), "trace": trace.as_dict()}

    # ---- the other BASELINE.json configurations, short (rank 0 of a single-GPU run) -------------------
    free_main = False
    if not args.no_configs and args.workload == "c3":
        for p in (c_dev, z1_dev, z2_dev, r_dev):
            lib.dev_free(ctx.handle, p)
        model.close()
        free_main = True
        cfgs = {}
        if world == 1:
            for name, fn in (("c1", lambda: quick_poly(pkg, ctx)),
                             # iteration counts of the form 1 + 8 k: the library issues lock-step iterations in chunks of
                             # 8 after the first, a count in between would time no-op launches past the cap
                             ("c2", lambda: quick_field(pkg, ctx, "c2", 16, 0, 201, 3, 2, hbm_peak, args.field_std)),
                             ("c4", lambda: quick_dense(pkg, ctx, torch)),
                             ("c5", lambda: quick_field(pkg, ctx, "c5", 8, 0, 17, 2, 1, hbm_peak, args.field_std))):
                try:
                    cfgs[name] = fn()
                except Exception as e:      # noqa: BLE001
                    cfgs[name] = {"error": repr(e)}
        # every N: one C5 mgvi_step with the Newton phase (the phase that communicates), 8 samples per GPU
        try:
            cfgs["c5_step"] = c5_step(pkg, ctx, rank, world, 8, 30, torch, dist, args.field_std)
        except Exception as e:              # noqa: BLE001
            cfgs["c5_step"] = {"error": repr(e)}
        out["configs"] = cfgs

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ----------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, times, workers, ns = cpu_reference_rate(args, args.ref_samples, args.ref_iters, repeats=2, warmup=1)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": "port",
                               "host_cores": os.cpu_count(),
                               "sample": "%d residual samples x %d CG iterations at full field size, %d worker "
                                         "processes (%.1f s per step)" % (ns, args.ref_iters, workers, float(np.mean(times))),
                               "note": CPU_NOTE}
    if rank == 0:
        emit(json.dumps(out))
    if not free_main:
        for p in (c_dev, z1_dev, z2_dev, r_dev):
            lib.dev_free(ctx.handle, p)
        model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
