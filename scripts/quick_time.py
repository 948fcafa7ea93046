"""Scratch timing of the NR path (not the bench): single-case run and a batch of outages."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import jgb200
from jgb200._lib import ptr, f64, i64, cplx
import torch

ps = jgb200.synthetic_grid()
ctx = jgb200.Context(0)
a = jgb200.newton_raphson(ps, ctx)
for k in ("nr.dim", "nr.nnz_j", "nr.nnz_lu", "nr.fronts", "nr.levels", "nr.max_front", "nr.flops", "nr.launches_per_iteration"):
    print(k, ctx.stat(k))
for rep in range(3):
    jgb200.set_initial_point(a); a._push_state()
    torch.cuda.synchronize(); t = time.perf_counter()
    ok = jgb200.power_flow(a)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("single run", ok, a.method.iteration, "iters", dt * 1e3, "ms ->", a.method.iteration / dt, "it/s")
base_vm, base_va = a.voltage.magnitude.copy(), a.voltage.angle.copy()
# batch
mdl = ps.model
lib = ctx.lib
for S in (32, 256, 1024, 2048):
    ks = np.arange(S) % ps.nbr
    of = i64(ps.frm[ks] + 1); ot = i64(ps.to[ks] + 1)
    dy = np.stack([mdl.y_ff[ks], mdl.y_ft[ks], mdl.y_tf[ks], mdl.y_tt[ks]], axis=1)
    dy = np.ascontiguousarray(dy).view(np.float64).reshape(S, 8)
    vm = np.empty((S, ps.n)); va = np.empty((S, ps.n)); it = np.empty(S, dtype=np.int32); st = np.empty(S, dtype=np.int8)
    tot = C.c_int64(0)
    jgb200.set_initial_point(a); a._push_state()
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        rc = lib.jgb_nr_batch(ctx.handle, S, ptr(of, C.c_int64), ptr(ot, C.c_int64), ptr(dy, C.c_double), 20, 1e-8,
                              ptr(vm, C.c_double), ptr(va, C.c_double), ptr(it, C.c_int32), ptr(st, C.c_int8), C.byref(tot))
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print("batch S", S, "rc", rc, "total iters", tot.value, "time", dt * 1e3, "ms ->", tot.value / dt, "it/s", "status", np.bincount(st + 3))
