"""Scratch timing of the constant-matrix solves: PMU state estimation Monte-Carlo draws and DC power-flow scenarios
on the synthetic 10k-bus grid, with the SciPy (SuperLU, one factorisation + one solve per draw) CPU time beside it."""
import sys, time
import numpy as np
import scipy.sparse.linalg as spla
sys.path.insert(0, '.')
import jgb200, torch
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ps = jgb200.synthetic_grid()
ctx = jgb200.Context(0)
a = jgb200.newton_raphson(ps, ctx); jgb200.power_flow(a)
vm, va = a.voltage.magnitude, a.voltage.angle
pw = jgb200.power(ps, vm, va)
mon = jgb200.measurement(ps)
jgb200.add_pmu(mon, pw, vm, va, buses=range(ps.n), branch=True, polar=False)
keep = mon.pmu["bus"] | (mon.pmu["mag_mean"] > 0.05)
mon.pmu = {k: v[keep] for k, v in mon.pmu.items()}
t0 = time.perf_counter(); se = jgb200.pmu_state_estimation(mon, ctx); t_setup = time.perf_counter() - t0
m = se.method
print("pmu se: rows", len(m.mean), "setup s", round(t_setup, 3), m.solver.dims())
rng = np.random.default_rng(1)
Z = m.mean[None, :] + 1e-4 * rng.standard_normal((R, len(m.mean)))
dZ = torch.from_numpy(Z).cuda(); dX = torch.empty((R, 2 * ps.n), dtype=torch.float64, device="cuda")
for _ in range(2): m.solver.solve_dev(R, dZ.data_ptr(), dX.data_ptr(), True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): m.solver.solve_dev(R, dZ.data_ptr(), dX.data_ptr(), True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
t0 = time.perf_counter(); X = m.solver.solve_projected(Z); dte = time.perf_counter() - t0
print(f"pmu se MC R={R}: resident {dt*1e3:.2f} ms -> {R/dt:.0f} draws/s; end to end {dte*1e3:.1f} ms -> {R/dte:.0f} draws/s")
h = m.coefficient.tocsc(); wh = (m.precision @ h).tocsc(); G = (h.T @ wh).tocsc()
t0 = time.perf_counter(); lu = spla.splu(G); tf = time.perf_counter() - t0
nc = min(R, 64)
t0 = time.perf_counter(); xs = np.stack([lu.solve(wh.T @ Z[r]) for r in range(nc)]); tc = (time.perf_counter() - t0) / nc
print(f"cpu: factor {tf*1e3:.1f} ms, per draw {tc*1e3:.2f} ms -> {1/tc:.0f} draws/s; max |dx| vs gpu {np.abs(xs - X[:nc]).max():.2e}")
# DC power flow scenarios
d = jgb200.dc_power_flow(ps, ctx)
dem = ps.pd[None, :] * (1 + 0.1 * rng.standard_normal((R, ps.n)))
B = d.rhs(demand=dem)
dB = torch.from_numpy(np.ascontiguousarray(B)).cuda(); dT = torch.empty((R, ps.n), dtype=torch.float64, device="cuda")
for _ in range(2): d.solver.solve_dev(R, dB.data_ptr(), dT.data_ptr(), False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.solver.solve_dev(R, dB.data_ptr(), dT.data_ptr(), False)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"dc pf R={R}: resident {dt*1e3:.2f} ms -> {R/dt:.0f} scenarios/s", d.solver.dims())
