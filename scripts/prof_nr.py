"""Profiling driver (run under ncu): mode single | batch [S]"""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import jgb200
from jgb200._lib import ptr, i64

mode = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ps = jgb200.synthetic_grid()
ctx = jgb200.Context(0)
a = jgb200.newton_raphson(ps, ctx)
if mode == "single":
    for rep in range(2):
        jgb200.set_initial_point(a); a._push_state()
        jgb200.power_flow(a)
    print("iters", a.method.iteration)
else:
    elig = jgb200.eligible_outages(ps)          # the sweep of the benchmark: bridges (islanding outages) are not part of it
    ks = elig[np.arange(S) % len(elig)]
    of, ot, dy = jgb200.outage_arrays(ps, ks)
    vm = np.empty((S, ps.n)); va = np.empty((S, ps.n)); it = np.empty(S, dtype=np.int32); st = np.empty(S, dtype=np.int8)
    tot = C.c_int64(0)
    rc = ctx.lib.jgb_nr_batch(ctx.handle, S, ptr(of, C.c_int64), ptr(ot, C.c_int64), ptr(dy, C.c_double), 2, 1e-8,
                              ptr(vm, C.c_double), ptr(va, C.c_double), ptr(it, C.c_int32), ptr(st, C.c_int8), C.byref(tot))
    print("rc", rc, tot.value)
