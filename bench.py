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
    ap.add_argument("--ref-samples", type=int, default=4)
    ap.add_argument("--ref-iters", type=int, default=2)
    return ap.parse_args()


def make_config(args, syn):
    dims, like, n = WORKLOADS[args.workload]
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
# CPU restatement of the reference path (oracle/), timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, n_samples, iters, repeats=1, warmup=0):
    """sample-iterations/s of the CPU restatement of MGVI.jl's CG sampler (oracle/, NumPy +
    scipy.fft on all host cores) on a bounded sample of the workload: `n_samples` residuals x
    `iters` CG iterations.  Returns (rate, seconds per repeat list, cores)."""
    import oracle as O
    cores = os.cpu_count() or 1
    O.mgvi_oracle.set_fft_workers(cores)
    om = O.FieldModel(cfg["dims"], cfg["amplitude"], cfg["scale"], cfg["data"], offset=cfg["offset"],
                      link=cfg["link"], likelihood=cfg["likelihood"], sigma=cfg["sigma"], layout=cfg["layout"])
    Z1 = np.empty((cfg["n_lambda"], n_samples))
    Z2 = np.empty((cfg["n_theta"], n_samples))
    for i in range(n_samples):
        Z1[:, i], Z2[:, i] = draws_for(cfg, i)
    times = []
    for rep in range(warmup + repeats):
        t0 = time.perf_counter()
        _R, it, _ = O.sample_residuals_cg(om, cfg["center"], Z1, Z2, atol=0.0, rtol=0.0, itmax=iters)
        dt = time.perf_counter() - t0
        assert int(it.sum()) == n_samples * iters
        if rep >= warmup:
            times.append(dt)
    rate = n_samples * iters * len(times) / sum(times)
    return rate, times, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    g = __import__("__graft_entry__")
    syn = g.load_package().synthetic
    cfg = make_config(args, syn)
    K, W = args.steps, max(args.warmup, 1)
    rate, times, cores = cpu_reference_rate(cfg, args.ref_samples, args.ref_iters, repeats=K, warmup=W)
    sample = "%d residual samples x %d CG iterations per step (%s, full field size)" % (
        args.ref_samples, args.ref_iters, args.workload)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "field_std": args.field_std},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement of MGVI.jl v0.4.5 hot path (NumPy + scipy.fft, all host cores); "
                                 "Julia is not installed in this image"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def workload_name(args, cfg):
    dims = "x".join(str(d) for d in cfg["dims"])
    like = "Normal(sigma=%.2g)" % cfg["sigma"] if cfg["likelihood"] == 0 else "Poisson"
    return "%s: %s correlated field, %s, exp link, %d residual samples" % (args.workload, dims, like,
                                                                         cfg["n_residuals"])


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
             400: ("rhs_first(q.n1, H)", 16 * N + 16 * N), 30: ("rhs_last(H, cA., +eta, b.b)", 16 * N + 40 * N),
             430: ("rhs_1d", 16 * N + 40 * N)}
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
    tot_ms = sum(k["avg_ms"] * k["launches"] for k in kern) or 1.0
    for k in kern:
        k["share"] = k["avg_ms"] * k["launches"] / tot_ms
    kern.sort(key=lambda k: -k["share"])
    dom = kern[0] if kern else None
    # DRAM bytes per launch of each kernel from the committed `ncu --set full` capture of this same
    # workload (profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum, 32 samples)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload, {})
    except Exception:
        pass
    for k_ in kern:
        t_ = traffic.get(str(k_["tag"]))
        k_["dram_traffic_bytes_per_launch"] = (t_ * n_loc / 32.0) if (t_ and world >= 1) else None
    roofline = None
    if dom and "algorithmic_gbs" in dom:
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["algorithmic_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom["algorithmic_gbs"] / hbm_peak,
                    "traffic": dom.get("dram_traffic_bytes_per_launch"),
                    "peak_source": peak_src, "share_of_step": dom["share"],
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"]}
    # the whole Fisher-CG iteration against B_cg: what north_star's 60 % target is about
    per_iter_ms = dev_ms_max / (K * ITERS)
    cg_path = {"B_cg_bytes_per_sample_iteration": ab["B_cg"], "lockstep_iteration_ms": per_iter_ms,
               "achieved_gbs": ab["B_cg"] * n_loc / (per_iter_ms * 1e-3) / 1e9}
    cg_path["frac_of_hbm_peak"] = cg_path["achieved_gbs"] / hbm_peak

    out = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "iters_per_step": ITERS, "field_std": args.field_std,
                   "samples_per_gpu": n_loc, "parallelism": "sample-sharded x%d" % world,
                   "l2": "inputs larger than L2 (%.1f GB streamed per step per GPU)" % (ab["B_cg"] * n_loc * ITERS / 1e9),
                   "timing": "CUDA events on the library stream, max over ranks"},
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

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ----------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, times, cores = cpu_reference_rate(cfg, args.ref_samples, args.ref_iters, repeats=1, warmup=0)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d residual samples x %d CG iterations at full field size (%.1f s)" % (
                                   args.ref_samples, args.ref_iters, sum(times)),
                               "note": "CPU restatement of MGVI.jl v0.4.5 hot path (oracle/: NumPy + scipy.fft, "
                                       "all host cores); Julia is not installed in this image"}
    if rank == 0:
        emit(json.dumps(out))
    for p in (c_dev, z1_dev, z2_dev, r_dev):
        lib.dev_free(ctx.handle, p)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
