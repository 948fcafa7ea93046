"""Scratch timing: batch NR with phase timers + single-case NR/WLS."""
import sys, time, ctypes as C, os
import numpy as np
sys.path.insert(0, '.')
import jgb200
from jgb200._lib import ptr, i64
import torch

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ps = jgb200.synthetic_grid()
ctx = jgb200.Context(0)
a = jgb200.newton_raphson(ps, ctx)
lib = ctx.lib
if "single" in sys.argv:
    for rep in range(3):
        jgb200.set_initial_point(a); a._push_state()
        torch.cuda.synchronize(); t = time.perf_counter()
        ok = jgb200.power_flow(a)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    lib.jgb_profile(ctx.handle, 1)
    jgb200.set_initial_point(a); a._push_state(); jgb200.power_flow(a)
    nfc = max(1.0, ctx.stat("nr.time.factor_count"))
    print("single NR", ok, a.method.iteration, f"{dt*1e3:.2f} ms -> {a.method.iteration/dt:.0f} it/s | per iteration GPU: factor {ctx.stat('nr.time.factor_ms')/nfc*1e3:.0f} us, backsolve {ctx.stat('nr.time.backsolve_ms')/nfc*1e3:.0f} us")
    lib.jgb_profile(ctx.handle, 0)
jgb200.set_initial_point(a); a._push_state()
elig = jgb200.eligible_outages(ps)
of, ot, dy = jgb200.outage_arrays(ps, elig[:S])
dev = torch.device("cuda")
of_d, ot_d, dy_d = torch.from_numpy(of).to(dev), torch.from_numpy(ot).to(dev), torch.from_numpy(dy).to(dev)
vm_d = torch.empty((S, ps.n), dtype=torch.float64, device=dev); va_d = torch.empty_like(vm_d)
it_d = torch.empty(S, dtype=torch.int32, device=dev); st_d = torch.empty(S, dtype=torch.int8, device=dev)
tot = C.c_int64(0)
def run():
    ctx.check(lib.jgb_nr_batch_dev(ctx.handle, S, C.c_void_p(of_d.data_ptr()), C.c_void_p(ot_d.data_ptr()), C.c_void_p(dy_d.data_ptr()), 20, 1e-8,
              C.c_void_p(vm_d.data_ptr()), C.c_void_p(va_d.data_ptr()), C.c_void_p(it_d.data_ptr()), C.c_void_p(st_d.data_ptr()), C.byref(tot)))
run(); torch.cuda.synchronize()
lib.jgb_profile(ctx.handle, 1)
t = time.perf_counter(); run(); run(); torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 2
nf = ctx.stat("nr.time.factor_count")
print(f"batch S={S}: {dt*1e3:.1f} ms/step, {tot.value/dt:.0f} it/s | per iteration: factor {ctx.stat('nr.time.factor_ms')/nf:.2f} ms, backsolve {ctx.stat('nr.time.backsolve_ms')/nf:.2f} ms, assemble {ctx.stat('nr.time.assemble_ms')/ctx.stat('nr.time.assemble_count'):.2f} ms; status ok {(st_d==0).all().item()}")
# verify against single-case run for one scenario
k = int(elig[5]); jgb200.update_branch(a, k, 0); jgb200.set_initial_point(a); jgb200.power_flow(a)
print("check scenario 5: max|dVm|", float(np.abs(vm_d[5].cpu().numpy() - a.voltage.magnitude).max()), "iters", int(it_d[5]), a.method.iteration)
