#!/usr/bin/env python
"""Benchmark of the jgb200 hot path (contract in the round prompt, tier section 4).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload all|nr|wls] [--scenarios S] [--draws D]

BASELINE.json's metric has two halves, both measured here on the synthetic 10k-bus meshed grid of configs[1]:

* NR  (top level of the JSON line): the N-1 contingency sweep of configs[3] — every rank takes S independent
  branch-outage scenarios, runs full Newton-Raphson power flows (mismatch!/solve! to 1e-8 from the flat start) for all
  of them in one batch, and the ranks all-gather the converged states once (jgb_allgather_states). One step = one such
  batch; metric = Newton iterations (solve! calls) per second, whole job.
* WLS (the "wls" block; top level with --workload wls): the Monte-Carlo study of configs[4] on the PMU + legacy
  measurement set of configs[2] — every rank takes D noise draws, runs Gauss-Newton WLS estimations (increment!/solve!
  to 1e-8) for all of them in one batch, same all-gather. Metric = Gauss-Newton iterations per second, whole job.

"scaling" is weak (S, D per GPU fixed as N grows); the configs as written — 10 000 outages / 1000 draws in TOTAL,
split over the N ranks — are measured in the same run and reported in the "strong_scaling" blocks. Every leg has a
device-resident figure (`value`), an end-to-end figure through the C ABI with pinned HOST buffers (`e2e`), a roofline
with SURVEY.md section 8(d)'s byte formulas, and the CPU arm on the box's host cores beside it.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1e-8
NR_METRIC, NR_UNIT, NR_MAX_ITER = "newton_iterations_per_s", "NR iterations/s", 20
WLS_METRIC, WLS_UNIT, WLS_MAX_ITER = "wls_gauss_newton_iterations_per_s", "GN iterations/s", 40


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=3)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ the WLS measurement set
def wls_monitoring(jgb200, ps, vm, va):
    """configs[2]: voltmeter at every bus, wattmeters and varmeters at every bus and both ends of every branch,
    rectangular PMUs on a seeded 10 % of the buses."""
    pw = jgb200.power(ps, vm, va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    buses = np.sort(np.random.default_rng(7).choice(ps.n, ps.n // 10, replace=False))
    jgb200.add_pmu(mon, pw, vm, va, buses=buses, polar=False)
    return mon


def wls_sigma(t):
    wd = np.ones(t.m)
    cp = t.w_colptr - 1
    for c in range(t.m):
        wd[c] = t.w_nzval[cp[c]]
    return np.sqrt(1 / wd)


def wls_draw(mean, sigma, q):
    """Draw q of the Monte-Carlo study: z = z_exact + sqrt(variance) * eps, eps ~ N(0, 1) from default_rng(1000 + q)."""
    return mean + sigma * np.random.default_rng(1000 + q).standard_normal(len(mean))


# --------------------------------------------------------------------------------------------- CPU reference arm
_CPU = {}


def _cpu_init(refactor=True):
    """Per-process setup (not timed, like the GPU arm's setup): grid, Ybus, index maps, the Newton-Raphson object. The
    factorisation object lives as long as the process: symbolic analysis once, numeric refactorisation afterwards —
    the reference's `factorization` / `factorization!` split (backend/utility.jl:470-500)."""
    import oracle
    from oracle import nr as onr
    from oracle.fast import FastNR
    s = oracle.synthetic_grid()
    base = oracle.ac_model(s)
    a = onr.newton_raphson(s, base)
    _CPU.update(s=s, base=base, f=FastNR(a, FastNR.NOPIVOT, refactor=refactor), refactor=refactor)


def _cpu_worker(args):
    """One process = one core: full NR power flows for a slice of outage scenarios with the CPU restatement."""
    ks, refactor = args
    from oracle.model import apply_outage
    if not _CPU or _CPU.get("refactor") != refactor:
        _cpu_init(refactor)
    s, base, f = _CPU["s"], _CPU["base"], _CPU["f"]
    iters = 0
    f.t_asm = f.t_fac = f.t_sol = 0.0
    t0 = time.perf_counter()
    for k in ks:
        m = apply_outage(s, base, int(k))
        f.set_y(m.nzval, m.nzval_t)
        f.reset()
        f.power_flow(NR_MAX_ITER, TOL)
        iters += f.iteration
    return iters, len(ks), time.perf_counter() - t0, f.t_asm, f.t_fac, f.t_sol


def _cpu_wls_init(refactor=True):
    import oracle
    from oracle import nr as onr, wls as owls
    from oracle.fast import FastNR, FastWLS
    import jgb200
    ps = jgb200.synthetic_grid()
    ps.model = jgb200.ac_model(ps)
    osys = oracle.synthetic_grid()
    o = onr.newton_raphson(osys)
    assert onr.power_flow(o)
    mon = wls_monitoring(jgb200, ps, o.vm, o.va)          # host-side tables only: no device is touched
    og = owls.gauss_newton(osys, mon, oracle.ac_model(osys), lu_options=FastNR.NOPIVOT)
    t = jgb200.ac_wls(ps, mon)
    _CPU.update(wls=FastWLS(og, refactor=refactor), wls_mean=t.mean.copy(), wls_sigma=wls_sigma(t), wls_refactor=refactor)


def _cpu_wls_worker(args):
    qs, refactor = args
    if "wls" not in _CPU or _CPU.get("wls_refactor") != refactor:
        _cpu_wls_init(refactor)
    fw, mean, sigma = _CPU["wls"], _CPU["wls_mean"], _CPU["wls_sigma"]
    iters = 0
    fw.t_rows = fw.t_gain = fw.t_fac = fw.t_sol = 0.0
    t0 = time.perf_counter()
    for q in qs:
        fw.mean[:] = wls_draw(mean, sigma, int(q))
        fw.reset()
        fw.state_estimation(WLS_MAX_ITER, TOL)
        iters += fw.iteration
    return iters, len(qs), time.perf_counter() - t0, fw.t_rows, fw.t_gain, fw.t_fac, fw.t_sol


class CpuArm:
    """The CPU restatement of JuliaGrid's loops on `cores` worker processes (C assembly loops; the sparse LU of the
    reference's UMFPACK / KLU restated as SuperLU for the first factorisation + a KLU-style numeric refactorisation
    afterwards). Only the solve loops are timed."""

    def __init__(self, cores, kind="nr", refactor=True):
        import multiprocessing as mp
        self.cores, self.kind, self.refactor = cores, kind, refactor
        self.worker = _cpu_worker if kind == "nr" else _cpu_wls_worker
        self.pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None
        self.run([[] for _ in range(cores)], chunked=True)        # every worker finishes its setup before the clock starts

    def run(self, items, chunked=False):
        chunks = items if chunked else [items[i::self.cores] for i in range(self.cores)]
        args = [(c, self.refactor) for c in chunks]
        t0 = time.perf_counter()
        res = self.pool.map(self.worker, args) if self.pool else [self.worker(args[0])]
        wall = time.perf_counter() - t0
        return {"iterations": sum(r[0] for r in res), "units": sum(r[1] for r in res), "wall": wall,
                "busy": sum(r[2] for r in res), "split": [sum(r[q] for r in res) for q in range(3, len(res[0]))]}

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def nr_config(S, world, total=None):
    return {"workload": "synthetic 10k-bus meshed grid (seed 20261017, BASELINE configs[1]) Newton-Raphson AC power "
                        "flow, N-1 contingency sweep (configs[3]): independent branch-outage solves to 1e-8 from the "
                        "flat start, batched per GPU",
            "buses": 10000, "branches": 12699, "dim_jacobian": 18498, "nnz_jacobian": 122308,
            "scenarios_per_gpu": S, "scenarios_total": total if total is not None else S * world, "tolerance": TOL,
            "max_iterations": NR_MAX_ITER,
            "l2_policy": "per-step working set (~4.5 MB x scenarios) is far larger than the 126 MB L2; no flush needed",
            "parallelism": f"scenario-sharded x{world}, one all-gather of converged states"}


def wls_config(D, world, m=None, total=None):
    return {"workload": "synthetic 10k-bus meshed grid, Gauss-Newton WLS state estimation on the PMU + legacy "
                        "measurement set of configs[2] (voltmeters, wattmeters, varmeters everywhere, rectangular PMUs "
                        "on 10 % of the buses), Monte-Carlo measurement-noise draws (configs[4]) batched per GPU",
            "buses": 10000, "measurement_rows": m, "draws_per_gpu": D, "draws_total": total if total is not None else D * world,
            "tolerance": TOL, "max_iterations": WLS_MAX_ITER,
            "l2_policy": "per-step working set (~25 MB x draws) is far larger than the 126 MB L2; no flush needed",
            "parallelism": f"draw-sharded x{world}, one all-gather of converged states"}


def run_reference(args):
    """`--impl reference`: the reference's CPU path (restated, see CpuArm) on all host cores, bounded samples of the same
    workloads. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import jgb200
    cores = os.cpu_count() or 1
    out = {}
    if args.workload in ("all", "nr"):
        elig = jgb200.eligible_outages(jgb200.synthetic_grid())
        per_step = 8 * cores                       # bounded sample: 8 scenarios per core per step
        arm = CpuArm(cores, "nr", refactor=True)
        tot = {"iterations": 0, "units": 0, "wall": 0.0, "split": [0.0, 0.0, 0.0]}
        for step in range(args.warmup + args.steps):
            lo = (step * per_step) % 4096
            r = arm.run(elig[lo: lo + per_step])
            if step >= args.warmup:
                tot["iterations"] += r["iterations"]; tot["units"] += r["units"]; tot["wall"] += r["wall"]
                tot["split"] = [a + b for a, b in zip(tot["split"], r["split"])]
        arm.close()
        # second number: a fresh SuperLU factorisation (ordering + symbolic + numeric) at every iteration, as in round 1
        arm2 = CpuArm(cores, "nr", refactor=False)
        r2 = arm2.run(elig[: 2 * cores])
        arm2.close()
        value = tot["iterations"] / tot["wall"]
        busy = max(1e-9, sum(tot["split"]))
        out["nr"] = {
            "impl": "reference", "metric": NR_METRIC, "value": value, "unit": NR_UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot["wall"] / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": nr_config(per_step, 1),
            "cpu_baseline": {"value": value, "unit": NR_UNIT, "cores": cores, "kind": "port",
                             "sample": f"{tot['units']} outage scenarios ({tot['iterations']} NR iterations) of the same "
                                       f"sweep on {cores} processes; CPU restatement of JuliaGrid: C assembly loops, one "
                                       "symbolic LU per process (SuperLU: MMD_AT_PLUS_A ordering, pivot order, patterns) "
                                       "and a KLU-style numeric refactorisation per iteration like lu!/klu! — Julia is "
                                       "not installed in this image",
                             "share_of_cpu_time": {"assembly": tot["split"][0] / busy, "refactor": tot["split"][1] / busy,
                                                   "solve": tot["split"][2] / busy},
                             "superlu_every_iteration": {"value": r2["iterations"] / r2["wall"], "unit": NR_UNIT,
                                                         "sample": f"{r2['units']} scenarios"}},
            "e2e": {"value": value, "unit": NR_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    if args.workload in ("all", "wls"):
        per_step = 2 * cores                       # bounded sample: 2 draws per core per step (~5 GN iterations each)
        arm = CpuArm(cores, "wls", refactor=True)
        tot = {"iterations": 0, "units": 0, "wall": 0.0, "split": [0.0] * 4}
        for step in range(args.warmup + args.steps):
            r = arm.run(np.arange(step * per_step, (step + 1) * per_step) % 1000)
            if step >= args.warmup:
                tot["iterations"] += r["iterations"]; tot["units"] += r["units"]; tot["wall"] += r["wall"]
                tot["split"] = [a + b for a, b in zip(tot["split"], r["split"])]
        arm.close()
        value = tot["iterations"] / tot["wall"]
        busy = max(1e-9, sum(tot["split"]))
        out["wls"] = {
            "impl": "reference", "metric": WLS_METRIC, "value": value, "unit": WLS_UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot["wall"] / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": wls_config(per_step, 1),
            "cpu_baseline": {"value": value, "unit": WLS_UNIT, "cores": cores, "kind": "port",
                             "sample": f"{tot['units']} Monte-Carlo draws ({tot['iterations']} GN iterations) on {cores} "
                                       "processes; CPU restatement: C normalEquation! loops, SciPy SpGEMM for H'WH (the "
                                       "reference's two SparseArrays products), symbolic LU once + numeric refactorisation",
                             "share_of_cpu_time": {"rows": tot["split"][0] / busy, "gain": tot["split"][1] / busy,
                                                   "refactor": tot["split"][2] / busy, "solve": tot["split"][3] / busy}},
            "e2e": {"value": value, "unit": WLS_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    line = out.get("wls") if args.workload == "wls" else out["nr"]
    if args.workload == "all":
        line["wls"] = out["wls"]
    emit(line)


# --------------------------------------------------------------------------------------------- our arm
class Job:
    """Process-wide plumbing: device, stream, torch.distributed (rendezvous, barrier, reductions of the timings), the
    library context and its own NCCL communicator for the all-gather of states."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import jgb200
        self.torch, self.dist, self.jgb = torch, dist, jgb200
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the JSON line only
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)     # library work and torch CUDA events share this stream
        torch.cuda.set_stream(self.stream)
        self.ctx = jgb200.Context(self.local, self.stream.cuda_stream)
        self.lib = self.ctx.lib
        if self.world > 1:
            jgb200.dist.comm_init(self.ctx)                  # the library's own communicator (jgb_comm_init)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, ms, count):
        """max over ranks of the device time, sum over ranks of the work."""
        if self.world == 1:
            return ms, count
        torch, dist = self.torch, self.dist
        t = torch.tensor([ms, float(count)], dtype=torch.float64, device=self.dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return float(tmax[0]), float(tsum[1])

    def timed(self, step_fn, finish=None, profile=False):
        """W warm-up steps, then exactly K timed steps between CUDA events on the library's stream, bracketed by a
        barrier + synchronize on both sides. step_fn(i) must not do host-side staging: inputs are prepared before."""
        torch, args = self.torch, self.args
        for w in range(args.warmup):
            step_fn(w)
        if finish:
            finish()
        if profile:
            torch.cuda.synchronize()
            self.lib.jgb_profile(self.ctx.handle, 1)     # phase timers (CUDA events, same stream): timed region only
        count = 0
        launches0 = self.ctx.stat("launches")
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            count += step_fn(args.warmup + k)
        if finish:
            finish()                                     # the last all-gather completes inside the timed region
        e1.record()
        self.barrier()
        ms, count = self.reduce(e0.elapsed_time(e1), count)
        return ms, count, self.ctx.stat("launches") - launches0

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class Gather:
    """The one collective of a sweep through the library (jgb_allgather_states, grouped NCCL all-gather on a private
    stream): it overlaps the next batch, so the outputs are double-buffered and at most one gather is in flight."""

    def __init__(self, job, rows, n):
        torch = job.torch
        self.job, self.rows, self.n = job, rows, n
        self.enabled = job.world > 1
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=job.dev)
        self.local = [(mk((rows, n), torch.float64), mk((rows, n), torch.float64), mk(rows, torch.int32), mk(rows, torch.int8))
                      for _ in range(2 if self.enabled else 1)]
        self.all = None
        if self.enabled:
            w = job.world
            self.all = (mk((w * rows, n), torch.float64), mk((w * rows, n), torch.float64), mk(w * rows, torch.int32),
                        mk(w * rows, torch.int8))
        self.turn = 0

    def buffers(self):
        return self.local[self.turn % len(self.local)]

    def submit(self):
        if self.enabled:
            vm, va, it, st = self.buffers()
            self.job.jgb.dist.allgather_states(self.job.ctx, vm, va, it, st, out=self.all)
        self.turn += 1

    def finish(self):
        if self.enabled:
            self.job.jgb.dist.comm_wait(self.job.ctx, host_blocking=False)


def nr_leg(job, S, extras=True, strong_total=None):
    """The Newton-Raphson half: device-resident, end-to-end, roofline, CPU baseline."""
    torch, jgb200, ctx, lib, args = job.torch, job.jgb, job.ctx, job.lib, job.args
    ps = jgb200.synthetic_grid()
    a = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(a) and a.method.iteration == 6, "base case must converge in 6 iterations"
    base_vm, base_va = a.voltage.magnitude.copy(), a.voltage.angle.copy()
    jgb200.set_initial_point(a)
    a._push_state()
    elig = jgb200.eligible_outages(ps)
    n = ps.n
    P = lambda t: C.c_void_p(t.data_ptr())
    tot = C.c_int64(0)

    def scenarios(step, count, rank, world):
        lo = ((step * world + rank) * count) % (len(elig) - count)
        return elig[lo: lo + count]

    def make_runner(count, nsteps, shard=None):
        """Inputs of every step staged on the device before the clock starts; outputs double-buffered for the gather."""
        ins = []
        for step in range(nsteps):
            ks = scenarios(step, count, job.rank, job.world) if shard is None else shard(step)
            of, ot, dy = jgb200.outage_arrays(ps, ks)
            ins.append((torch.from_numpy(of).to(job.dev), torch.from_numpy(ot).to(job.dev), torch.from_numpy(dy).to(job.dev)))
        g = Gather(job, count, n)

        def step_fn(i):
            of_d, ot_d, dy_d = ins[i]
            # outputs alternate between two buffer sets: the gather that read this set two steps ago has completed, because
            # the gather submitted one step ago waited for it before it started (jgb_allgather_states, stream order)
            vm_d, va_d, it_d, st_d = g.buffers()
            ctx.check(lib.jgb_nr_batch_dev(ctx.handle, count, P(of_d), P(ot_d), P(dy_d), NR_MAX_ITER, TOL, P(vm_d), P(va_d),
                                           P(it_d), P(st_d), C.byref(tot)))
            g.submit()
            return tot.value
        return step_fn, g

    nsteps = args.warmup + args.steps
    step_dev, gather = make_runner(S, nsteps)
    ms_dev, iters_dev, launches = job.timed(step_dev, finish=gather.finish, profile=True)
    t_fac, t_bs, t_asm = ctx.stat("nr.time.factor_ms"), ctx.stat("nr.time.backsolve_ms"), ctx.stat("nr.time.assemble_ms")
    n_fac, n_asm = ctx.stat("nr.time.factor_count"), ctx.stat("nr.time.assemble_count")
    lib.jgb_profile(ctx.handle, 0)
    st_last = gather.local[(gather.turn - 1) % len(gather.local)][3]
    assert bool((st_last == 0).all()), "every scenario of the sweep must converge"
    comm_calls = ctx.stat("comm.calls")

    # ---- end to end: pinned host buffers through jgb_nr_batch (H2D of the scenario list, D2H of every state)
    of_h = torch.empty(S, dtype=torch.int64).pin_memory()
    ot_h = torch.empty(S, dtype=torch.int64).pin_memory()
    dy_h = torch.empty((S, 8), dtype=torch.float64).pin_memory()
    vm_h = torch.empty((S, n), dtype=torch.float64).pin_memory()
    va_h = torch.empty((S, n), dtype=torch.float64).pin_memory()
    it_h = torch.empty(S, dtype=torch.int32).pin_memory()
    st_h = torch.empty(S, dtype=torch.int8).pin_memory()
    host_in = [jgb200.outage_arrays(ps, scenarios(step, S, job.rank, job.world)) for step in range(nsteps)]
    HP = lambda t, ct: C.cast(t.data_ptr(), C.POINTER(ct))

    def step_host(i):
        of, ot, dy = host_in[i]
        of_h.numpy()[:] = of          # the caller's arrays -> the pinned buffers handed to the C ABI (part of e2e)
        ot_h.numpy()[:] = ot
        dy_h.numpy()[:] = dy
        ctx.check(lib.jgb_nr_batch(ctx.handle, S, HP(of_h, C.c_int64), HP(ot_h, C.c_int64), HP(dy_h, C.c_double),
                                   NR_MAX_ITER, TOL, HP(vm_h, C.c_double), HP(va_h, C.c_double), HP(it_h, C.c_int32),
                                   HP(st_h, C.c_int8), C.byref(tot)))
        return tot.value
    ms_e2e, iters_e2e, _ = job.timed(step_host)
    assert bool((st_h.numpy() == 0).all())

    # ---- strong-scaling form of configs[3]: `strong_total` outages in TOTAL, split over the ranks
    strong = None
    if strong_total and job.world > 1:
        per = -(-strong_total // job.world)
        step_s, g_s = make_runner(per, nsteps)
        ms_s, it_s, _ = job.timed(step_s, finish=g_s.finish)
        strong = {"value": it_s / (ms_s * 1e-3), "unit": NR_UNIT, "ms_per_step": ms_s / args.steps,
                  "scenarios_total": per * job.world, "scenarios_per_gpu": per,
                  "note": "configs[3] as written: the whole sweep split over the ranks, one all-gather"}

    result = {"ms_dev": ms_dev, "iters_dev": iters_dev, "launches": launches, "ms_e2e": ms_e2e, "iters_e2e": iters_e2e,
              "strong": strong, "comm_calls": comm_calls}
    if job.rank != 0:
        return result, {}, (ps, base_vm, base_va), a

    # ---- roofline with SURVEY 8(d)'s formulas (per scenario-iteration, FP64 values, shared index data not counted)
    hbm, which = peaks()
    nnzj, dimj = ctx.stat("nr.nnz_j"), ctx.stat("nr.dim")
    nnz_lu, upd = ctx.stat("nr.batch.nnz_lu"), ctx.stat("nr.batch.upd_size")
    b_asm = 16.0 * n + 8.0 * dimj + 8.0 * nnzj
    b_fac = 8.0 * nnzj + 8.0 * nnz_lu
    b_sol = 8.0 * nnz_lu + 24.0 * dimj
    b_upd = 16.0 * dimj + 16.0 * n
    design = 16.0 * upd                                  # update blocks written to and read back from HBM (not in 8(d))
    phase_ms = t_fac / max(1.0, n_fac)
    achieved = b_fac * S / (phase_ms * 1e-3) / 1e9 if t_fac > 0 else 0.0
    it_per_step = iters_dev / args.steps / job.world     # per rank
    step_ms = ms_dev / args.steps
    whole = (b_asm + b_fac + b_sol + b_upd) * it_per_step / (step_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
            # ncu DRAM bytes of the factor launches of one iteration at tj["scenarios"] scenarios; the traffic is per
            # scenario (no cross-scenario reuse), so it scales linearly to this run's batch
            traffic = tj.get("mf_factor_kernel_dram_bytes_per_factor_phase")
            if traffic is not None and tj.get("scenarios"):
                traffic = traffic * (S / float(tj["scenarios"]))
    roofline = {"bound": "hbm", "kernel": "mf_factor_* (all launches of one factor phase: front assembly + partial LU + "
                                          "forward solve)",
                "achieved": achieved, "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic,
                "algorithmic_bytes_per_launch_group": b_fac * S,
                "algorithmic_bytes_formula": "SURVEY 8(d): B_fac = 8 nnzJ + 8 nnz(L+U) per scenario, x scenarios",
                "design_bytes_per_launch_group": design * S,
                "design_bytes_note": "contribution blocks written to HBM by a front and read back by its parent; not part "
                                     "of 8(d)'s formula, reported separately (achieved incl. them: "
                                     f"{(b_fac + design) * S / (phase_ms * 1e-3) / 1e9:.0f} GB/s)",
                "nnz_lu": nnz_lu, "launches_per_phase": ctx.stat("nr.batch.factor_launches"), "avg_phase_ms": phase_ms,
                "whole_step": {"achieved": whole, "frac": whole / hbm,
                               "bytes_per_scenario_iteration": {"assembly": b_asm, "factor": b_fac, "solve": b_sol,
                                                                "update": b_upd}},
                "per_phase": {"assemble": {"ms": t_asm / max(1.0, n_asm), "GBs": b_asm * S / (t_asm / max(1.0, n_asm) * 1e-3) / 1e9 if t_asm > 0 else None},
                              "backsolve": {"ms": t_bs / max(1.0, n_fac), "GBs": b_sol * S / (t_bs / max(1.0, n_fac) * 1e-3) / 1e9 if t_bs > 0 else None}},
                "share_of_step": {"factor": t_fac / ms_dev, "backsolve": t_bs / ms_dev, "assemble": t_asm / ms_dev}}

    # ---- CPU baseline: bounded sample of the same sweep on one host core (the reference is single-threaded)
    arm = CpuArm(1, "nr", refactor=True)
    r = arm.run(elig[:16] if args.headline_only else elig[:128])
    arm.close()
    busy = max(1e-9, sum(r["split"]))
    cpu = {"value": r["iterations"] / r["busy"], "unit": NR_UNIT, "cores": 1, "kind": "port",
           "sample": f"first {r['units']} outage scenarios of the sweep ({r['iterations']} NR iterations, {r['busy']:.1f} s) on "
                     "1 core; CPU restatement of JuliaGrid: C assembly loops, symbolic LU once (SuperLU, MMD_AT_PLUS_A, no "
                     "pivoting), KLU-style numeric refactorisation per iteration (lu!/klu!)",
           "share_of_cpu_time": {"assembly": r["split"][0] / busy, "refactor": r["split"][1] / busy,
                                 "solve": r["split"][2] / busy}}

    single = {}
    if extras:
        reps = 5
        jgb200.set_initial_point(a)
        a._push_state()
        jgb200.power_flow(a)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        it_sum = 0
        for _ in range(reps):
            jgb200.set_initial_point(a)
            a._push_state()
            jgb200.power_flow(a)
            it_sum += a.method.iteration
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        single["nr_single_case_iterations_per_s"] = it_sum / dt
        single["nr_single_case_iterations"] = a.method.iteration
        lpi = ctx.stat("nr.launches_per_iteration")
        single["nr_single_case"] = {"ms_per_iteration": 1e3 * dt / it_sum, "kernel_launches_per_iteration": lpi,
                                    "bytes_per_iteration_8d": 36.0 * ctx.stat("nr.nnz_y") + 53.0 * n + 8 * dimj + 8 * nnzj
                                                              + 8 * nnzj + 8 * ctx.stat("nr.nnz_lu") + 8 * ctx.stat("nr.nnz_lu") + 24 * dimj,
                                    "note": "configs[1]: one case on one GPU is bound by the dependent chain of per-level "
                                            "launches (one CTA per front), not by HBM"}
        single["nr_single_case"]["hbm_frac"] = (single["nr_single_case"]["bytes_per_iteration_8d"]
                                                / (single["nr_single_case"]["ms_per_iteration"] * 1e-3) / 1e9 / hbm)
        # where the iteration goes: CUDA-event phase times of one more (untimed) solve; the rest is the assembly / update /
        # convergence kernels, the per-iteration 16-byte read-back and launch gaps
        lib.jgb_profile(ctx.handle, 1)
        jgb200.set_initial_point(a)
        a._push_state()
        jgb200.power_flow(a)
        nfc = max(1.0, ctx.stat("nr.time.factor_count"))
        fac_us, bs_us = 1e3 * ctx.stat("nr.time.factor_ms") / nfc, 1e3 * ctx.stat("nr.time.backsolve_ms") / nfc
        lib.jgb_profile(ctx.handle, 0)
        single["nr_single_case"]["per_iteration_us"] = {
            "factor_levels": fac_us, "backsolve_levels": bs_us,
            "note": "CUDA-event spans of a separate run with the phase timers on (launches issued one by one instead of "
                    "the replayed graph, so the two spans include their launch gaps); what is left of ms_per_iteration is "
                    "assembly + update + convergence test + the 16-byte read-back",
            "launch_list": "profiles/r02_launch_summary_nr_single.txt (21 + 21 level launches, 15-45 us / 7-20 us each)"}
    return result, {"roofline": roofline, "cpu_baseline": cpu, **single}, (ps, base_vm, base_va), a


def wls_leg(job, D, truth, strong_total=None, extras=True):
    """The Gauss-Newton WLS half (configs[2] measurement set, configs[4] Monte-Carlo draws)."""
    torch, jgb200, ctx, lib, args = job.torch, job.jgb, job.ctx, job.lib, job.args
    ps, vm_true, va_true = truth
    mon = wls_monitoring(jgb200, ps, vm_true, va_true)
    se = jgb200.gauss_newton(mon, ctx)
    t = se.method.tables
    m, n = int(t.m), ps.n
    sigma = wls_sigma(t)
    jgb200.set_voltage_se(se, ps.vm, ps.va)
    se._push()
    P = lambda x: C.c_void_p(x.data_ptr())
    tot = C.c_int64(0)
    nsteps = args.warmup + args.steps
    # two distinct sets of draws alternate over the steps (pinned host copies double as the e2e inputs)
    nsets = 2

    def draws(count, which, rank, world):
        first = (which * world + rank) * count
        return np.stack([wls_draw(t.mean, sigma, (first + q) % 100000) for q in range(count)])

    def make_runner(count):
        sets = [torch.from_numpy(draws(count, w, job.rank, job.world)).pin_memory() for w in range(nsets)]
        dsets = [z.to(job.dev) for z in sets]
        g = Gather(job, count, n)
        obj = torch.empty(count, dtype=torch.float64, device=job.dev)

        def step_fn(i):
            vm_d, va_d, it_d, st_d = g.buffers()
            ctx.check(lib.jgb_wls_batch_dev(ctx.handle, count, P(dsets[i % nsets]), WLS_MAX_ITER, TOL, P(vm_d), P(va_d),
                                            P(it_d), P(st_d), P(obj), C.byref(tot)))
            g.submit()
            return tot.value
        return step_fn, g, sets

    step_dev, gather, host_sets = make_runner(D)
    ms_dev, iters_dev, launches = job.timed(step_dev, finish=gather.finish, profile=True)
    tm = {k: ctx.stat(f"wls.time.{k}_ms") for k in ("rows", "gain", "factor", "backsolve")}
    n_fac = max(1.0, ctx.stat("wls.time.factor_count"))
    lib.jgb_profile(ctx.handle, 0)
    st_last = gather.local[(gather.turn - 1) % len(gather.local)][3]
    all_ok = bool((st_last == 0).all())

    vm_h = torch.empty((D, n), dtype=torch.float64).pin_memory()
    va_h = torch.empty((D, n), dtype=torch.float64).pin_memory()
    it_h = torch.empty(D, dtype=torch.int32).pin_memory()
    st_h = torch.empty(D, dtype=torch.int8).pin_memory()
    ob_h = torch.empty(D, dtype=torch.float64).pin_memory()
    HP = lambda x, ct: C.cast(x.data_ptr(), C.POINTER(ct))

    def step_host(i):
        ctx.check(lib.jgb_wls_batch(ctx.handle, D, HP(host_sets[i % nsets], C.c_double), WLS_MAX_ITER, TOL,
                                    HP(vm_h, C.c_double), HP(va_h, C.c_double), HP(it_h, C.c_int32), HP(st_h, C.c_int8),
                                    HP(ob_h, C.c_double), C.byref(tot)))
        return tot.value
    ms_e2e, iters_e2e, _ = job.timed(step_host)

    strong = None
    if strong_total and job.world > 1:
        per = -(-strong_total // job.world)
        step_s, g_s, _ = make_runner(per)
        ms_s, it_s, _ = job.timed(step_s, finish=g_s.finish)
        strong = {"value": it_s / (ms_s * 1e-3), "unit": WLS_UNIT, "ms_per_step": ms_s / args.steps,
                  "draws_total": per * job.world, "draws_per_gpu": per,
                  "note": "configs[4] as written: the draws split over the ranks, one all-gather"}
    if job.rank != 0:
        return None

    hbm, which = peaks()
    nnzh, nnzg, nbr = ctx.stat("wls.nnz_h"), ctx.stat("wls.nnz_g"), ctx.stat("wls.nbr")
    nnz_lu = ctx.stat("wls.batch.nnz_lu")
    nnz_l = (nnz_lu + 2 * n) / 2.0                       # L of the LDL^T incl. diagonal: half of the symmetric L+U
    b_rows = 29.0 * m + 8.0 * nnzh + 16.0 * n + 56.0 * nbr
    b_gain = 12.0 * nnzh + 8.0 * m + 8.0 * nnzg + 16.0 * n
    b_solve = 8.0 * nnzg + 16.0 * nnz_l + 48.0 * n
    it_per_step = iters_dev / args.steps / job.world
    incs_per_step = it_per_step + D                      # every draw ends with one more increment! (the converged test)
    step_ms = ms_dev / args.steps
    phases = {"rows": b_rows, "gain": b_gain, "factor": b_solve * 0.5, "backsolve": b_solve * 0.5}
    dom = max(("rows", "gain", "factor", "backsolve"), key=lambda k: tm[k])
    fac_ms = (tm["factor"] + tm["backsolve"]) / n_fac
    ach_fac = b_solve * D / (fac_ms * 1e-3) / 1e9 if fac_ms > 0 else 0.0
    whole = (b_rows + b_gain + b_solve) * incs_per_step / (step_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "gain factor + solve (mf_factor_sym_* / mf_factor_bulk_* / mf_backsolve_*: LDL^T of "
                                          "H'WH on the fixed elimination tree, all launches of one increment!)",
                "achieved": ach_fac, "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": ach_fac / hbm,
                "traffic": None,
                "algorithmic_bytes_per_launch_group": b_solve * D,
                "algorithmic_bytes_formula": "SURVEY 8(d): 8 nnzG + 16 nnz(L) + 48 n per draw, x draws",
                "nnz_l": nnz_l, "avg_phase_ms": fac_ms, "dominant_phase": dom,
                "design_bytes_per_launch_group": 16.0 * ctx.stat("wls.batch.upd_size") * D,
                "whole_step": {"achieved": whole, "frac": whole / hbm,
                               "bytes_per_draw_increment": {"rows": b_rows, "gain": b_gain, "factor_solve": b_solve}},
                "per_phase": {k: {"ms": tm[k] / n_fac, "GBs": (phases[k] * D / (tm[k] / n_fac * 1e-3) / 1e9) if tm[k] > 0 else None}
                              for k in tm},
                "share_of_step": {k: tm[k] / ms_dev for k in tm}}

    arm = CpuArm(1, "wls", refactor=True)
    r = arm.run(np.arange(2 if args.headline_only else 8))
    arm.close()
    busy = max(1e-9, sum(r["split"]))
    cpu = {"value": r["iterations"] / r["busy"], "unit": WLS_UNIT, "cores": 1, "kind": "port",
           "sample": f"first {r['units']} Monte-Carlo draws ({r['iterations']} GN iterations, {r['busy']:.1f} s) on 1 core; CPU "
                     "restatement: C normalEquation! loops, SciPy SpGEMM for H'WH, symbolic LU once + numeric refactorisation",
           "share_of_cpu_time": {"rows": r["split"][0] / busy, "gain": r["split"][1] / busy, "refactor": r["split"][2] / busy,
                                 "solve": r["split"][3] / busy}}
    block = {"metric": WLS_METRIC, "value": iters_dev / (ms_dev * 1e-3), "unit": WLS_UNIT, "n_gpus": job.world,
             "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
             "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
             "config": wls_config(D, job.world, m),
             "e2e": {"value": iters_e2e / (ms_e2e * 1e-3), "unit": WLS_UNIT, "ms_per_step": ms_e2e / args.steps,
                     "h2d_bytes_per_step": int(D * m * 8) * job.world, "d2h_bytes_per_step": int(D * (16 * n + 13)) * job.world},
             "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
             "iterations_per_step": iters_dev / args.steps, "all_converged": all_ok,
             "rows": m, "nnz_h": nnzh, "nnz_g": nnzg}
    if strong:
        block["strong_scaling"] = strong

    if extras:
        try:
            # configs[2]: the single case with default_rng(1) noise, beside the CPU restatement on the same input
            z = t.mean + sigma * np.random.default_rng(1).standard_normal(m)
            jgb200.set_mean(se, z)
            jgb200.set_voltage_se(se, ps.vm, ps.va)
            jgb200.state_estimation(se)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            it_sum = 0
            for _ in range(3):
                jgb200.set_voltage_se(se, ps.vm, ps.va)
                jgb200.state_estimation(se)
                it_sum += se.method.iteration
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            block["single_case"] = {"gn_iterations_per_s": it_sum / dt, "iterations": se.method.iteration,
                                    "ms_per_iteration": 1e3 * dt / it_sum,
                                    "kernel_launches_per_iteration": ctx.stat("wls.launches_per_iteration"),
                                    "hbm_frac": (b_rows + b_gain + b_solve) / (dt / it_sum) / 1e9 / hbm}
            jgb200.residual_test(se, threshold=1e300)          # builds the lists; threshold never met: nothing removed
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rt = jgb200.residual_test(se, threshold=1e300)
            torch.cuda.synchronize()
            block["single_case"]["residual_test_ms"] = (time.perf_counter() - t0) * 1e3
            block["single_case"]["max_normalized_residual"] = rt.maxNormalizedResidual
            import oracle
            from oracle import wls as owls
            from oracle.fast import FastNR, FastWLS
            osys = oracle.synthetic_grid()
            og = owls.gauss_newton(osys, mon, oracle.ac_model(osys), lu_options=FastNR.NOPIVOT)
            og.mean[:] = z
            fw = FastWLS(og, refactor=True)
            t0 = time.perf_counter()
            fw.state_estimation()
            block["single_case"]["cpu_gn_iterations_per_s"] = fw.iteration / (time.perf_counter() - t0)
            block["single_case"]["cpu_iterations"] = fw.iteration
            block["single_case"]["cpu_vs_gpu_max_abs_voltage_difference"] = float(
                max(np.abs(fw.vm - se.voltage.magnitude).max(), np.abs(fw.va - se.voltage.angle).max()))
        except Exception as e:      # the extras must never sink the line
            block["single_case_error"] = str(e)
    return block


def other_extras(job, ps, a):
    """SURVEY 8(f) rows measured in round 1, kept as extras: PMU-only linear estimation, fast Newton-Raphson batches."""
    torch, jgb200, ctx = job.torch, job.jgb, job.ctx
    n = ps.n
    single = {}
    try:
        import scipy.sparse.linalg as spla
        pw = jgb200.power(ps, a.voltage.magnitude, a.voltage.angle)
        mon = jgb200.measurement(ps)
        jgb200.add_pmu(mon, pw, a.voltage.magnitude, a.voltage.angle, buses=range(n), branch=True, polar=False)
        keep = mon.pmu["bus"] | (mon.pmu["mag_mean"] > 0.05)
        mon.pmu = {k: v[keep] for k, v in mon.pmu.items()}
        pse = jgb200.pmu_state_estimation(mon, ctx)
        pm = pse.method
        R = 1024
        Z = pm.mean[None, :] + 1e-4 * np.random.default_rng(1).standard_normal((R, len(pm.mean)))
        dZ = torch.from_numpy(Z).cuda()
        dX = torch.empty((R, 2 * n), dtype=torch.float64, device="cuda")
        for _ in range(2):
            pm.solver.solve_dev(R, dZ.data_ptr(), dX.data_ptr(), True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            pm.solver.solve_dev(R, dZ.data_ptr(), dX.data_ptr(), True)
        torch.cuda.synchronize()
        single["pmu_se_monte_carlo_draws_per_s"] = 5 * R / (time.perf_counter() - t0)
        Zp = torch.from_numpy(Z).pin_memory()
        Xp = torch.empty((R, 2 * n), dtype=torch.float64).pin_memory()
        pm.solver.solve_projected(Zp.numpy(), out=Xp.numpy())
        t0 = time.perf_counter()
        for _ in range(3):
            X = pm.solver.solve_projected(Zp.numpy(), out=Xp.numpy())
        single["pmu_se_monte_carlo_draws_per_s_e2e"] = 3 * R / (time.perf_counter() - t0)
        single["pmu_se_rows"] = int(len(pm.mean))
        h = pm.coefficient.tocsc()
        wh = (pm.precision @ h).tocsc()
        lu = spla.splu((h.T @ wh).tocsc())
        t0 = time.perf_counter()
        xs = np.stack([lu.solve(wh.T @ Z[r]) for r in range(32)])
        single["pmu_se_cpu_baseline_draws_per_s"] = 32 / (time.perf_counter() - t0)
        single["pmu_se_cpu_vs_gpu_max_abs_difference"] = float(np.abs(xs - X[:32]).max())
    except Exception as e:
        single["pmu_se_error"] = str(e)
    try:
        fa = jgb200.fast_newton_raphson_xb(ps, ctx)
        Rf = 1024
        scale = 1.0 + 0.1 * np.random.default_rng(3).standard_normal((Rf, n))
        sp0, sq0, _ = ps.supply
        pin = sp0[None, :] - ps.pd[None, :] * scale
        qin = sq0[None, :] - ps.qd[None, :] * scale
        jgb200.fnr_batch(fa, pin[:64], qin[:64], iteration=60)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, fit, fst = jgb200.fnr_batch(fa, pin, qin, iteration=60)
        torch.cuda.synchronize()
        single["fnr_scenarios_iterations_per_s_e2e"] = float(fit.sum()) / (time.perf_counter() - t0)
        single["fnr_scenarios"] = Rf
        single["fnr_all_converged"] = bool((fst == 0).all())
        single["fnr_mean_iterations"] = float(fit.mean())
    except Exception as e:
        single["fnr_error"] = str(e)
    return single


def run_ours(args):
    job = Job(args)
    S, D = args.scenarios, args.draws
    sampler = ClockSampler(job.local)
    if job.rank == 0:
        sampler.start()
    extras = not args.headline_only
    nr_res = nr_extra = truth = a = None
    if args.workload in ("all", "nr"):
        nr_res, nr_extra, truth, a = nr_leg(job, S, extras=extras, strong_total=10000)
    wls_block = None
    if args.workload in ("all", "wls"):
        if truth is None:
            import jgb200
            ps = jgb200.synthetic_grid()
            a0 = jgb200.newton_raphson(ps, job.ctx)
            assert jgb200.power_flow(a0)
            truth = (ps, a0.voltage.magnitude.copy(), a0.voltage.angle.copy())
        try:
            wls_block = wls_leg(job, D, truth, strong_total=1000, extras=extras)
        except Exception as e:
            if args.workload == "wls":
                raise
            wls_block = {"error": str(e)}
            job.barrier()
    clocks = sampler.result() if job.rank == 0 else None
    if job.rank != 0:
        job.close()
        return
    n = 10000
    if args.workload == "wls":
        line = dict(wls_block)
        line["clocks"] = clocks
    else:
        r = nr_res
        line = {
            "metric": NR_METRIC, "value": r["iters_dev"] / (r["ms_dev"] * 1e-3), "unit": NR_UNIT, "n_gpus": job.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_dev"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": nr_config(S, job.world),
            "e2e": {"value": r["iters_e2e"] / (r["ms_e2e"] * 1e-3), "unit": NR_UNIT, "ms_per_step": r["ms_e2e"] / args.steps,
                    "h2d_bytes_per_step": int(S * (8 + 8 + 64)) * job.world,
                    "d2h_bytes_per_step": int(S * (16 * n + 5)) * job.world},
            "gpu_launches": int(r["launches"]), "clocks": clocks, "iterations_per_step": r["iters_dev"] / args.steps,
            "allgather": {"via": "jgb_allgather_states (NCCL inside libjgb200.so)", "calls": r["comm_calls"]} if job.world > 1 else None,
            **nr_extra,
        }
        if r["strong"]:
            line["strong_scaling"] = r["strong"]
        if wls_block is not None:
            line["wls"] = wls_block
        if extras and a is not None:
            try:
                line.update(other_extras(job, truth[0], a))
            except Exception as e:
                line["extras_error"] = str(e)
    emit(line)
    job.close()


_REAL_STDOUT = None


def protect_stdout():
    """stdout carries the one JSON line and nothing else: file descriptor 1 is pointed at stderr for the whole run (NCCL,
    the CUDA runtime and worker processes may print there), the line itself goes to a duplicate of the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "nr", "wls"],
                    help="all: NR line with the WLS leg in its \"wls\" block (default); nr / wls: that half only, at top level")
    ap.add_argument("--headline-only", action="store_true",
                    help="profiling runs (ncu launch lists): skip the single-case / linear extras, shorten the CPU samples")
    ap.add_argument("--scenarios", type=int, default=10000,
                    help="outage scenarios per GPU per step (default: the whole 10 000-outage sweep of configs[3])")
    ap.add_argument("--draws", type=int, default=1000,
                    help="Monte-Carlo draws per GPU per step (default: the 1000 draws of configs[4])")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
