"""Scratch timing of the WLS path: single case and a Monte-Carlo batch (config 3 / 5 of BASELINE.json)."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import jgb200, torch
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ps = jgb200.synthetic_grid()
ctx = jgb200.Context(0)
a = jgb200.newton_raphson(ps, ctx); jgb200.power_flow(a)
pw = jgb200.power(ps, a.voltage.magnitude, a.voltage.angle)
mon = jgb200.measurement(ps)
jgb200.add_voltmeter(mon, a.voltage.magnitude); jgb200.add_wattmeter(mon, pw); jgb200.add_varmeter(mon, pw)
buses = np.sort(np.random.default_rng(7).choice(ps.n, ps.n // 10, replace=False))
jgb200.add_pmu(mon, pw, a.voltage.magnitude, a.voltage.angle, buses=buses, polar=False)
t0 = time.perf_counter(); se = jgb200.gauss_newton(mon, ctx); print("setup s", time.perf_counter() - t0)
for k in ("m", "nnz_h", "nnz_g", "gain_terms", "fronts", "levels", "max_front", "flops", "nnz_lu", "u_size", "upd_size"):
    print(k, ctx.stat("wls." + k), end="; ")
print()
t = se.method.tables
wd = np.array([t.w_nzval[t.w_colptr[c] - 1] for c in range(t.m)])
sig = np.sqrt(1 / wd)
z = t.mean + sig * np.random.default_rng(1).standard_normal(t.m)
jgb200.set_mean(se, z); jgb200.state_estimation(se)
ctx.lib.jgb_profile(ctx.handle, 1)
torch.cuda.synchronize(); t0 = time.perf_counter()
jgb200.set_voltage_se(se, ps.vm, ps.va); ok = jgb200.state_estimation(se)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
nf = max(1, ctx.stat("wls.time.factor_count"))
print(f"single WLS: {ok} {se.method.iteration} iters {dt*1e3:.1f} ms -> {se.method.iteration/dt:.1f} it/s | per increment: rows {ctx.stat('wls.time.rows_ms')/nf:.3f} gain {ctx.stat('wls.time.gain_ms')/nf:.3f} factor {ctx.stat('wls.time.factor_ms')/nf:.3f} backsolve {ctx.stat('wls.time.backsolve_ms')/nf:.3f} ms")
Z = np.stack([t.mean + sig * np.random.default_rng(1000 + s).standard_normal(t.m) for s in range(S)])
jgb200.set_voltage_se(se, ps.vm, ps.va)
res = jgb200.wls_batch(se, Z)
ctx.lib.jgb_profile(ctx.handle, 1)
torch.cuda.synchronize(); t0 = time.perf_counter()
res = jgb200.wls_batch(se, Z)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
nf = max(1, ctx.stat("wls.time.factor_count"))
print(f"batch WLS S={S}: {dt*1e3:.1f} ms, total iters {res.total_iterations} -> {res.total_iterations/dt:.1f} it/s | per increment: rows {ctx.stat('wls.time.rows_ms')/nf:.2f} gain {ctx.stat('wls.time.gain_ms')/nf:.2f} factor {ctx.stat('wls.time.factor_ms')/nf:.2f} backsolve {ctx.stat('wls.time.backsolve_ms')/nf:.2f} ms; converged {(res.status==0).all()}")
