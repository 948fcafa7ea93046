"""Single-case Newton-Raphson on the 70 000-bus fixture (tests/golden/case_ACTIVSg70k.npz): setup time and iterations/s."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jgb200, torch
from conftest import product_system
ps = product_system("case_ACTIVSg70k"); ctx = jgb200.Context(0)
t0 = time.perf_counter(); a = jgb200.newton_raphson(ps, ctx); print("setup s", round(time.perf_counter() - t0, 2))
jgb200.power_flow(a)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    jgb200.set_initial_point(a); a._push_state(); jgb200.power_flow(a)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("70k NR: iterations", a.method.iteration, "ms", round(dt * 1e3, 2), "it/s", round(a.method.iteration / dt, 1),
      {k: ctx.stat("nr." + k) for k in ("fronts", "levels", "max_front", "nnz_lu", "flops")})
