"""Profiling driver for the WLS path (run under ncu): single-case stateEstimation!"""
import sys
import numpy as np
sys.path.insert(0, '.')
import jgb200
ps = jgb200.synthetic_grid()
ctx = jgb200.Context(0)
a = jgb200.newton_raphson(ps, ctx); jgb200.power_flow(a)
pw = jgb200.power(ps, a.voltage.magnitude, a.voltage.angle)
mon = jgb200.measurement(ps)
jgb200.add_voltmeter(mon, a.voltage.magnitude); jgb200.add_wattmeter(mon, pw); jgb200.add_varmeter(mon, pw)
buses = np.sort(np.random.default_rng(7).choice(ps.n, ps.n // 10, replace=False))
jgb200.add_pmu(mon, pw, a.voltage.magnitude, a.voltage.angle, buses=buses, polar=False)
se = jgb200.gauss_newton(mon, ctx)
if "batchonly" not in sys.argv:
    print("increment", jgb200.increment(se))
    print("increment", jgb200.increment(se))
if len(sys.argv) > 2 and sys.argv[1] == "batch":
    # Monte-Carlo batch of S draws, iteration cap 1 (two increments): launch list of the batch kernels
    S = int(sys.argv[2])
    t = se.method.tables
    wd = np.array([t.w_nzval[t.w_colptr[c] - 1] for c in range(t.m)])
    Z = np.stack([t.mean + np.sqrt(1 / wd) * np.random.default_rng(1000 + s).standard_normal(t.m) for s in range(S)])
    jgb200.set_voltage_se(se, ps.vm, ps.va)
    res = jgb200.wls_batch(se, Z, iteration=1)
    print("batch", S, res.total_iterations)
