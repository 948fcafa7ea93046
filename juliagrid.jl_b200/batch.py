"""Scenario batches on one GPU: N-1 contingency sweeps (independent NR solves) and Monte-Carlo measurement draws
(independent GN-WLS solves). Scenarios share topology, index maps and the symbolic factorisation; the reference does
these as user loops of updateBranch!/powerFlow! (test/powerFlow/reusing.jl:40-84) and update*!/stateEstimation!."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import ptr, f64, i64
from .ac_power_flow import AcPowerFlow
from .ac_state_estimation import AcStateEstimation
from .cases import PowerSystem


@dataclass
class BatchResult:
    vm: np.ndarray            # S x n
    va: np.ndarray
    iterations: np.ndarray    # int32 per scenario: number of solve! calls
    status: np.ndarray        # int8: 0 converged, 1 iteration cap, -3 singular
    total_iterations: int
    objective: np.ndarray | None = None


def eligible_outages(system: PowerSystem) -> np.ndarray:
    """In-service branches whose removal keeps the grid connected (not a bridge of the simple graph, or has a
    parallel twin), in branch order. One iterative Tarjan pass on the host."""
    n = system.n
    on = np.flatnonzero(system.status == 1)
    f, t = system.frm[on], system.to[on]
    key = np.minimum(f, t) * n + np.maximum(f, t)
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    eu, ev = uniq // n, uniq % n
    adj_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(adj_ptr, eu + 1, 1)
    np.add.at(adj_ptr, ev + 1, 1)
    adj_ptr = np.cumsum(adj_ptr)
    fill = adj_ptr[:-1].copy()
    adj_v = np.empty(2 * len(uniq), dtype=np.int64)
    adj_e = np.empty(2 * len(uniq), dtype=np.int64)
    for e in range(len(uniq)):
        a, b = eu[e], ev[e]
        adj_v[fill[a]], adj_e[fill[a]] = b, e
        fill[a] += 1
        adj_v[fill[b]], adj_e[fill[b]] = a, e
        fill[b] += 1
    disc = np.full(n, -1, dtype=np.int64)
    low = np.zeros(n, dtype=np.int64)
    bridge = np.zeros(len(uniq), dtype=bool)
    timer = 0
    for root in range(n):
        if disc[root] >= 0:
            continue
        stack = [(root, -1, adj_ptr[root])]
        disc[root] = low[root] = timer
        timer += 1
        while stack:
            v, pe, it = stack[-1]
            if it < adj_ptr[v + 1]:
                stack[-1] = (v, pe, it + 1)
                w, e = adj_v[it], adj_e[it]
                if e == pe:
                    continue
                if disc[w] >= 0:
                    low[v] = min(low[v], disc[w])
                else:
                    disc[w] = low[w] = timer
                    timer += 1
                    stack.append((w, e, adj_ptr[w]))
            else:
                stack.pop()
                if stack:
                    p = stack[-1][0]
                    low[p] = min(low[p], low[v])
                    if low[v] > disc[p]:
                        bridge[pe] = True
    is_bridge = bridge[inv] & (cnt[inv] == 1)
    return on[~is_bridge]


def outage_arrays(system: PowerSystem, branches):
    """C-ABI inputs of jgb_nr_batch: 1-based end buses (0 = base case) and the four Y-parameters per scenario."""
    ks = np.asarray(branches, dtype=np.int64)
    mdl = system.model
    base = ks < 0
    kk = np.where(base, 0, ks)
    of = np.where(base, 0, system.frm[kk] + 1).astype(np.int64)
    ot = np.where(base, 0, system.to[kk] + 1).astype(np.int64)
    dy = np.stack([mdl.y_ff[kk], mdl.y_ft[kk], mdl.y_tf[kk], mdl.y_tt[kk]], axis=1)
    dy[base] = 0
    return of, ot, np.ascontiguousarray(dy).view(np.float64).reshape(len(ks), 8)


def nr_batch(a: AcPowerFlow, branches, iteration: int = 20, tolerance: float = 1e-8,
             warm_start: bool = False) -> BatchResult:
    """One powerFlow! per outage scenario (branch index, -1 = base case). Every scenario starts from the analysis' start
    point (setInitialPoint! before each powerFlow!, like fnr_batch); with warm_start=True from analysis.voltage as it
    stands — after power_flow() that is the converged base case, the reference's behaviour when the user loop does not
    call setInitialPoint!. The analysis' own voltages are left untouched either way."""
    system = a.system
    S = len(branches)
    of, ot, dy = outage_arrays(system, branches)
    if warm_start:
        if a._state_dirty:
            a._push_state()
    else:
        a.ctx.check(a.ctx.lib.jgb_nr_set_state(a.ctx.handle, ptr(f64(a._initial[0]), C.c_double),
                                               ptr(f64(a._initial[1]), C.c_double)))
        a._state_dirty = True           # the device holds the start point now, not analysis.voltage
    vm, va = np.empty((S, system.n)), np.empty((S, system.n))
    it, st = np.empty(S, dtype=np.int32), np.empty(S, dtype=np.int8)
    tot = C.c_int64(0)
    a.ctx.check(a.ctx.lib.jgb_nr_batch(a.ctx.handle, S, ptr(of, C.c_int64), ptr(ot, C.c_int64), ptr(dy, C.c_double),
                                       iteration, tolerance, ptr(vm, C.c_double), ptr(va, C.c_double),
                                       ptr(it, C.c_int32), ptr(st, C.c_int8), C.byref(tot)))
    return BatchResult(vm, va, it, st, tot.value)


def wls_batch(a: AcStateEstimation, Z, iteration: int = 40, tolerance: float = 1e-8) -> BatchResult:
    """One stateEstimation! per row of Z (S x m measurement means), all from the analysis' current start point."""
    Z = f64(Z)
    S, m = Z.shape
    if m != a.method.tables.m:
        raise ValueError("Z must be S x m")
    if a._dirty:
        a._push()
    n = a.system.n
    vm, va = np.empty((S, n)), np.empty((S, n))
    it, st, ob = np.empty(S, dtype=np.int32), np.empty(S, dtype=np.int8), np.empty(S)
    tot = C.c_int64(0)
    a.ctx.check(a.ctx.lib.jgb_wls_batch(a.ctx.handle, S, ptr(Z, C.c_double), iteration, tolerance,
                                        ptr(vm, C.c_double), ptr(va, C.c_double), ptr(it, C.c_int32),
                                        ptr(st, C.c_int8), ptr(ob, C.c_double), C.byref(tot)))
    return BatchResult(vm, va, it, st, tot.value, ob)
