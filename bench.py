#!/usr/bin/env python
"""Benchmark of the jgb200 hot path (contract in the round prompt, tier section ④).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--scenarios S]

Workload (config.workload): the synthetic 10k-bus meshed grid of BASELINE.json configs[1], solved
as the N-1 contingency sweep of configs[3]: every rank takes S independent branch-outage scenarios (weak scaling: S per
GPU is fixed), runs full Newton-Raphson power flows (mismatch!/solve! to 1e-8) for all of them in one batch, and the
ranks all-gather the converged states once. One step = one such batch; metric = Newton iterations (solve! calls) per
second, whole job.  Extra keys report the single-case rates of configs[1] (NR) and configs[2] (GN-WLS).
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

METRIC = "newton_iterations_per_s"
UNIT = "NR iterations/s"
TOL = 1e-8
MAX_ITER = 20


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


# --------------------------------------------------------------------------------------------- CPU reference arm
_CPU = {}


def _cpu_init():
    """Per-process setup (not timed, like the GPU arm's setup): grid, Ybus, index maps, SuperLU-backed NR object."""
    import oracle
    from oracle import nr as onr
    from oracle.fast import FastNR
    s = oracle.synthetic_grid()
    base = oracle.ac_model(s)
    a = onr.newton_raphson(s, base)
    _CPU.update(s=s, base=base, f=FastNR(a, FastNR.NOPIVOT))


def _cpu_worker(ks):
    """One process = one core: full NR power flows for a slice of outage scenarios with the CPU restatement."""
    from oracle.model import apply_outage
    if not _CPU:
        _cpu_init()
    s, base, f = _CPU["s"], _CPU["base"], _CPU["f"]
    iters = 0
    t0 = time.perf_counter()
    for k in ks:
        m = apply_outage(s, base, int(k))
        f.set_y(m.nzval, m.nzval_t)
        f.reset()
        f.power_flow(MAX_ITER, TOL)
        iters += f.iteration
    return iters, len(ks), time.perf_counter() - t0


class CpuArm:
    """CPU restatement (C assembly loops + SuperLU with the no-pivot symmetric settings = the faster of the two
    BASELINE.md §3 settings) on `cores` worker processes; only the solve loops are timed."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init) if cores > 1 else None
        if self.pool:
            self.pool.map(_cpu_worker, [[] for _ in range(cores)])     # make sure every worker finished its setup
        else:
            _cpu_init()

    def run(self, ks):
        chunks = [ks[i::self.cores] for i in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, chunks) if self.pool else [_cpu_worker(chunks[0])]
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), sum(r[1] for r in res), wall

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import jgb200
    ps = jgb200.synthetic_grid()
    elig = jgb200.eligible_outages(ps)
    cores = os.cpu_count() or 1
    per_step = 4 * cores                        # bounded sample: 4 scenarios per core per step (~1 s per step)
    arm = CpuArm(cores)
    total, iters_all, scen_all = 0.0, 0, 0
    for step in range(args.warmup + args.steps):
        lo = (step * per_step) % 4096
        iters, scen, wall = arm.run(elig[lo: lo + per_step])
        if step >= args.warmup:
            total += wall
            iters_all += iters
            scen_all += scen
    arm.close()
    value = iters_all / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(per_step, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{scen_all} outage scenarios ({iters_all} NR iterations) of the same sweep, "
                                   f"{cores} processes; CPU restatement of JuliaGrid (C loops + SuperLU no-pivot "
                                   f"instead of UMFPACK/KLU) — Julia is not installed in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(S, world):
    return {"workload": "synthetic 10k-bus meshed grid (seed 20261017, BASELINE configs[1]) Newton-Raphson AC power "
                        "flow, N-1 contingency sweep (configs[3]): independent branch-outage solves to 1e-8 from the "
                        "flat start, batched per GPU",
            "buses": 10000, "branches": 12699, "dim_jacobian": 18498, "nnz_jacobian": 122308,
            "scenarios_per_gpu": S, "scenarios_total": S * world, "tolerance": TOL, "max_iterations": MAX_ITER,
            "l2_policy": "per-step working set (~4.5 MB x scenarios) is far larger than the 126 MB L2; no flush needed",
            "parallelism": f"scenario-sharded x{world}, one all-gather of converged states"}


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import jgb200
    from jgb200._lib import ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=dev)
    S = args.scenarios
    side = torch.cuda.Stream(device=dev)          # library work and torch CUDA events share this stream
    torch.cuda.set_stream(side)
    ctx = jgb200.Context(local, side.cuda_stream)
    lib = ctx.lib

    ps = jgb200.synthetic_grid()
    a = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(a) and a.method.iteration == 6, "base case must converge in 6 iterations"
    jgb200.set_initial_point(a)
    a._push_state()
    elig = jgb200.eligible_outages(ps)
    n = ps.n

    def scenarios(step):
        lo = ((step * world + rank) * S) % (len(elig) - S)
        return elig[lo: lo + S]

    # ---- buffers: pinned host + device
    of_h = torch.empty(S, dtype=torch.int64).pin_memory()
    ot_h = torch.empty(S, dtype=torch.int64).pin_memory()
    dy_h = torch.empty((S, 8), dtype=torch.float64).pin_memory()
    vm_h = torch.empty((S, n), dtype=torch.float64).pin_memory()
    va_h = torch.empty((S, n), dtype=torch.float64).pin_memory()
    it_h = torch.empty(S, dtype=torch.int32).pin_memory()
    st_h = torch.empty(S, dtype=torch.int8).pin_memory()
    of_d, ot_d, dy_d = of_h.to(dev), ot_h.to(dev), dy_h.to(dev)
    vm_d = torch.empty((S, n), dtype=torch.float64, device=dev)
    va_d = torch.empty((S, n), dtype=torch.float64, device=dev)
    it_d = torch.empty(S, dtype=torch.int32, device=dev)
    st_d = torch.empty(S, dtype=torch.int8, device=dev)
    tot = C.c_int64(0)

    def load(step, to_device):
        of, ot, dy = jgb200.outage_arrays(ps, scenarios(step))
        of_h.numpy()[:] = of
        ot_h.numpy()[:] = ot
        dy_h.numpy()[:] = dy
        if to_device:
            of_d.copy_(of_h)
            ot_d.copy_(ot_h)
            dy_d.copy_(dy_h)

    def step_device():
        ctx.check(lib.jgb_nr_batch_dev(ctx.handle, S, C.c_void_p(of_d.data_ptr()), C.c_void_p(ot_d.data_ptr()),
                                       C.c_void_p(dy_d.data_ptr()), MAX_ITER, TOL, C.c_void_p(vm_d.data_ptr()),
                                       C.c_void_p(va_d.data_ptr()), C.c_void_p(it_d.data_ptr()),
                                       C.c_void_p(st_d.data_ptr()), C.byref(tot)))
        if world > 1:
            # the one all-gather of the sweep runs on NCCL's stream while the next batch is solved; at most one is in
            # flight, and the last one is waited for inside the timed region (finish_device)
            if pending:
                pending.pop().wait()
            pending.append(jgb200.dist.gather_batch_result_async(vm_d, va_d, it_d, st_d))
        return tot.value

    pending = []

    def finish_device():
        while pending:
            pending.pop().wait()

    def step_host():
        ctx.check(lib.jgb_nr_batch(ctx.handle, S, C.cast(of_h.data_ptr(), C.POINTER(C.c_int64)),
                                   C.cast(ot_h.data_ptr(), C.POINTER(C.c_int64)),
                                   C.cast(dy_h.data_ptr(), C.POINTER(C.c_double)), MAX_ITER, TOL,
                                   C.cast(vm_h.data_ptr(), C.POINTER(C.c_double)),
                                   C.cast(va_h.data_ptr(), C.POINTER(C.c_double)),
                                   C.cast(it_h.data_ptr(), C.POINTER(C.c_int32)),
                                   C.cast(st_h.data_ptr(), C.POINTER(C.c_int8)), C.byref(tot)))
        return tot.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, to_device, profile=False, finish=None):
        """W warm-up steps, then exactly K timed steps between CUDA events; max over ranks."""
        for w in range(args.warmup):
            load(w, to_device)
            fn()
        if finish:
            finish()
        if profile:
            torch.cuda.synchronize()
            lib.jgb_profile(ctx.handle, 1)       # phase timers (CUDA events on the same stream) cover the timed region only
        loads = []
        iters = 0
        launches0 = ctx.stat("launches")
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host = 0.0
        e0.record()
        w0 = time.perf_counter()
        for k in range(args.steps):
            th = time.perf_counter()
            load(args.warmup + k, to_device)       # device leg: staging the next batch is outside the metric...
            t_host += time.perf_counter() - th
            iters += fn()
        if finish:
            finish()                               # outstanding collectives complete inside the timed region
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        if to_device:
            ms -= 1e3 * t_host                     # ...so its host time is removed from the device-resident figure
        t = torch.tensor([ms, float(iters)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, iters = float(tmax[0]), float(tsum[1])
        return ms, iters, ctx.stat("launches") - launches0, wall

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, iters_dev, launches, _ = timed(step_device, True, profile=True, finish=finish_device)
    t_fac, t_bs, t_asm = ctx.stat("nr.time.factor_ms"), ctx.stat("nr.time.backsolve_ms"), ctx.stat("nr.time.assemble_ms")
    n_fac = ctx.stat("nr.time.factor_count")
    lib.jgb_profile(ctx.handle, 0)
    ms_e2e, iters_e2e, _, _ = timed(step_host, False)
    clocks = sampler.result() if rank == 0 else None
    assert bool((st_h.numpy() == 0).all()), "every scenario of the sweep must converge"

    # ---- single-case rates (configs[1] and [2]), rank 0 only, a few repetitions each
    single = {}
    if rank == 0 and not args.headline_only:
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
        single["nr_single_case_iterations_per_s"] = it_sum / (time.perf_counter() - t0)
        single["nr_single_case_iterations"] = a.method.iteration
        try:
            pw = jgb200.power(ps, a.voltage.magnitude, a.voltage.angle)
            mon = jgb200.measurement(ps)
            jgb200.add_voltmeter(mon, a.voltage.magnitude)
            jgb200.add_wattmeter(mon, pw)
            jgb200.add_varmeter(mon, pw)
            buses = np.sort(np.random.default_rng(7).choice(n, n // 10, replace=False))
            jgb200.add_pmu(mon, pw, a.voltage.magnitude, a.voltage.angle, buses=buses, polar=False)
            se = jgb200.gauss_newton(mon, ctx)
            t = se.method.tables
            wd = np.ones(t.m)
            cp = t.w_colptr - 1
            for c in range(t.m):
                wd[c] = t.w_nzval[cp[c]]
            z = t.mean + np.sqrt(1 / wd) * np.random.default_rng(1).standard_normal(t.m)
            jgb200.set_mean(se, z)
            jgb200.state_estimation(se)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            it_sum = 0
            for _ in range(3):
                jgb200.set_voltage_se(se, ps.vm, ps.va)
                jgb200.state_estimation(se)
                it_sum += se.method.iteration
            torch.cuda.synchronize()
            single["wls_single_case_gn_iterations_per_s"] = it_sum / (time.perf_counter() - t0)
            single["wls_single_case_iterations"] = se.method.iteration
            single["wls_rows"] = int(t.m)
            # largest normalised residual (selected inverse of the gain factor + row projection)
            jgb200.residual_test(se, threshold=1e300)          # builds the lists; threshold never met: nothing removed
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rt = jgb200.residual_test(se, threshold=1e300)
            torch.cuda.synchronize()
            single["wls_residual_test_ms"] = (time.perf_counter() - t0) * 1e3
            single["wls_max_normalized_residual"] = rt.maxNormalizedResidual
            # configs[4]: the 1000 Monte-Carlo noise draws of the same measurement set, all on this GPU
            Sm = 1000
            Z = np.stack([t.mean + np.sqrt(1 / wd) * np.random.default_rng(1000 + q).standard_normal(t.m)
                          for q in range(Sm)])
            jgb200.set_voltage_se(se, ps.vm, ps.va)
            Zpin = torch.from_numpy(Z).pin_memory().numpy()       # measurement draws in pinned host memory (662 MB)
            jgb200.wls_batch(se, Zpin)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rb = jgb200.wls_batch(se, Zpin)
            torch.cuda.synchronize()
            single["wls_monte_carlo_gn_iterations_per_s"] = rb.total_iterations / (time.perf_counter() - t0)
            single["wls_monte_carlo_draws"] = Sm
            single["wls_monte_carlo_all_converged"] = bool((rb.status == 0).all())
            # CPU restatement of the same single-case estimation (C normalEquation! loops + SciPy SpGEMM + SuperLU)
            jgb200.set_mean(se, z)
            jgb200.set_voltage_se(se, ps.vm, ps.va)
            jgb200.state_estimation(se)
            import oracle
            from oracle import wls as owls
            from oracle.fast import FastNR, FastWLS
            osys = oracle.synthetic_grid()
            og = owls.gauss_newton(osys, mon, oracle.ac_model(osys), lu_options=FastNR.NOPIVOT)
            og.mean[:] = z
            fw = FastWLS(og)
            t0 = time.perf_counter()
            fw.state_estimation()
            single["wls_cpu_baseline_gn_iterations_per_s"] = fw.iteration / (time.perf_counter() - t0)
            single["wls_cpu_vs_gpu_max_abs_voltage_difference"] = float(
                max(np.abs(fw.vm - se.voltage.magnitude).max(), np.abs(fw.va - se.voltage.angle).max()))
        except Exception as e:      # the WLS extras must never sink the headline line
            single["wls_error"] = str(e)
        try:
            # PMU-only linear state estimation, 1024 Monte-Carlo draws on one gain factorisation
            # (jgb_lin_*), beside SciPy SuperLU (one factorisation, one solve per draw) on one host core
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
            # end to end from pinned host memory, like the headline e2e leg (434 MB in, 164 MB out per call)
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
            # fast Newton-Raphson (XB) on 1024 load scenarios sharing the two device factorisations
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (mf_factor_kernel: front assembly + partial LU + forward solve)
    hbm, which = peaks()
    # algorithmic bytes of ONE factor phase for S scenarios: read J values + mismatch, write packed U rows, write and
    # read back every update block: 8 * (nnzJ + dim + u_size + 2 * upd_size) per scenario
    nnzj, dimj = ctx.stat("nr.nnz_j"), ctx.stat("nr.dim")
    bytes_fac = 8.0 * (nnzj + dimj + ctx.stat("nr.batch.u_size") + 2 * ctx.stat("nr.batch.upd_size")) * S
    fac_launches = ctx.stat("nr.batch.factor_launches")
    achieved = (bytes_fac * n_fac) / (t_fac * 1e-3) / 1e9 if t_fac > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
            # ncu DRAM bytes of the factor launches of one iteration, measured at tj["scenarios"] scenarios; the
            # traffic is per scenario (no cross-scenario reuse), so it scales linearly to this run's batch
            traffic = tj.get("mf_factor_kernel_dram_bytes_per_factor_phase")
            if traffic is not None and tj.get("scenarios"):
                traffic = traffic * (S / float(tj["scenarios"]))
    roofline = {"bound": "hbm", "kernel": "mf_factor_kernel (all launches of one factor phase)", "achieved": achieved,
                "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                "algorithmic_bytes_per_phase": bytes_fac, "launches_per_phase": fac_launches,
                "avg_phase_ms": t_fac / max(1.0, n_fac),
                "share_of_step": {"factor": t_fac / ms_dev, "backsolve": t_bs / ms_dev, "assemble": t_asm / ms_dev}}

    # ---- CPU baseline: bounded sample of the same sweep on one host core (the reference is single-threaded)
    arm = CpuArm(1)
    it_cpu, sc_cpu, busy = arm.run(elig[:8] if args.headline_only else elig[:96])
    rate = it_cpu / busy
    cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"first {sc_cpu} outage scenarios of the sweep ({it_cpu} NR iterations, {busy:.1f} s) on 1 core; "
                     "CPU restatement of JuliaGrid: C assembly loops + SciPy SuperLU (MMD_AT_PLUS_A, no pivoting) "
                     "standing in for UMFPACK/KLU"}

    value = iters_dev / (ms_dev * 1e-3)
    e2e_v = iters_e2e / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(S, world),
        "e2e": {"value": e2e_v, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(S * (8 + 8 + 64)) * world, "d2h_bytes_per_step": int(S * (16 * n + 5)) * world},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "iterations_per_step": iters_dev / args.steps, **single,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--headline-only", action="store_true",
                    help="profiling runs (ncu launch lists): skip the single-case / WLS / linear extras and shorten the "
                         "CPU baseline sample")
    ap.add_argument("--scenarios", type=int, default=10000,
                    help="outage scenarios per GPU per step (default: the whole 10 000-outage sweep of configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
