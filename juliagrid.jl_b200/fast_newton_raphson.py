"""Host mirror of the reference's fast Newton-Raphson (BX / XB) operator surface, backed by libjgb200.so.

    fast_newton_raphson_bx(system) / _xb(system) <-> fastNewtonRaphsonBX / XB(system, B200)  src/powerFlow/acPowerFlow.jl:215-339
    mismatch_fnr(analysis)                       <-> mismatch!(analysis)                      :686-727
    solve_fnr(analysis)                          <-> solve!(analysis)                         :913-983
    power_flow_fnr(analysis; ...)                <-> powerFlow!(analysis; ...)                :1389-1433
    fnr_batch(analysis, supply..., demand...)    <-> the user loop updateBus!(active, reactive) + powerFlow! per injection
                                                     scenario: B' and B'' do not depend on the injections, so the two device
                                                     factorisations are shared by the whole block
The two constant Jacobians are built here like the reference builds them (fastNewtonJacobian :341-412,
fastNewtonJacobian! :414-451, jacobianCoefficient :453-480) — in the Julia drop-in they are JuliaGrid's own.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from ._lib import Context, ptr, f64, i64, i8, cplx
from .ac_power_flow import Polar, _initialize
from .cases import PowerSystem
from .model import AcModel, ac_model


class FastNewtonRaphsonMethod:
    def __init__(self):
        self.active = self.reactive = None      # B', B'' (scipy CSC on the reference's structural pattern)
        self.pq = self.pvpq = None              # 0-based, -1 when absent
        self.bx = True
        self.iteration = 0


class AcPowerFlowFast:
    def __init__(self, system: PowerSystem, ctx: Context):
        self.system, self.ctx = system, ctx
        self.voltage: Polar = None
        self.method = FastNewtonRaphsonMethod()
        self.bus_type = self.slack = None
        self._initial = None
        self._dirty = True

    def _push(self):
        self.ctx.check(self.ctx.lib.jgb_fnr_set_state(self.ctx.handle, ptr(f64(self.voltage.magnitude), C.c_double),
                                                      ptr(f64(self.voltage.angle), C.c_double)))
        self._dirty = False

    def _pull(self):
        n = self.system.n
        vm, va = np.empty(n), np.empty(n)
        self.ctx.check(self.ctx.lib.jgb_fnr_get_state(self.ctx.handle, ptr(vm, C.c_double), ptr(va, C.c_double)))
        self.voltage = Polar(vm, va)


def fast_jacobians(system: PowerSystem, mdl: AcModel, bus_type, slack, bx: bool):
    n = system.n
    pq = np.full(n, -1, dtype=np.int64)
    pvpq = np.full(n, -1, dtype=np.int64)
    is_pq, non_slack = bus_type == 1, bus_type != 3
    pq[is_pq] = np.arange(is_pq.sum())
    pvpq[non_slack] = np.arange(non_slack.sum())
    npq = int(is_pq.sum())
    # structural patterns on the Ybus pattern (explicit zeros kept): column bus non-slack x row bus non-slack / PQ x PQ
    col = np.repeat(np.arange(n), np.diff(mdl.colptr))
    row = mdl.rowval - 1
    mp = non_slack[col] & non_slack[row]
    mq = is_pq[col] & is_pq[row]
    on = system.status == 1
    i, j = system.frm[on], system.to[on]
    bsi = 0.5 * system.b[on]
    tinv = 1.0 / system.tap[on]
    s, c = np.sin(system.shift[on]), np.cos(system.shift[on])
    y = mdl.admittance[on]
    if bx:
        bmk, A, B = -1.0 / system.x[on], y.real, y.imag
    else:
        bmk, A, B = y.imag, np.zeros(len(i)), -1.0 / system.x[on]
    den = c * c + s * s
    pij, pji = (-A * s - B * c) / den, (A * s - B * c) / den
    qa, qb, qc = -bmk * tinv, (bmk + bsi) * tinv * tinv, bmk + bsi
    both = non_slack[i] & non_slack[j]
    rows = np.concatenate([pvpq[row[mp]], pvpq[i[both]], pvpq[j[both]], pvpq[i[non_slack[i]]], pvpq[j[non_slack[j]]]])
    cols = np.concatenate([pvpq[col[mp]], pvpq[j[both]], pvpq[i[both]], pvpq[i[non_slack[i]]], pvpq[j[non_slack[j]]]])
    vals = np.concatenate([np.zeros(mp.sum()), pij[both], pji[both], (B / den)[non_slack[i]], B[non_slack[j]]])
    active = sp.coo_matrix((vals, (rows, cols)), shape=(n - 1, n - 1)).tocsc()
    bq = is_pq[i] & is_pq[j]
    sh = np.flatnonzero(is_pq & (system.bs != 0))
    rows = np.concatenate([pq[row[mq]], pq[i[bq]], pq[j[bq]], pq[i[is_pq[i]]], pq[j[is_pq[j]]], pq[sh]])
    cols = np.concatenate([pq[col[mq]], pq[j[bq]], pq[i[bq]], pq[i[is_pq[i]]], pq[j[is_pq[j]]], pq[sh]])
    vals = np.concatenate([np.zeros(mq.sum()), qa[bq], qa[bq], qb[is_pq[i]], qc[is_pq[j]], system.bs[sh]])
    reactive = sp.coo_matrix((vals, (rows, cols)), shape=(npq, npq)).tocsc()
    active.sort_indices()
    reactive.sort_indices()
    return active, reactive, pq, pvpq


def _fast(system: PowerSystem, bx: bool, ctx: Context | None, device: int) -> AcPowerFlowFast:
    if system.model is None:
        system.model = ac_model(system)
    mdl: AcModel = system.model
    ctx = ctx or Context(device)
    a = AcPowerFlowFast(system, ctx)
    bus_type, slack, vm, va = _initialize(system)
    a.bus_type, a.slack = bus_type, slack
    m = a.method
    m.active, m.reactive, m.pq, m.pvpq = fast_jacobians(system, mdl, bus_type, slack, bx)
    m.bx = bx
    P = lambda v, ct: ptr(v, ct)
    bp = (i64(m.active.indptr + 1), i64(m.active.indices + 1), f64(m.active.data))
    bq = (i64(m.reactive.indptr + 1), i64(m.reactive.indices + 1), f64(m.reactive.data))
    ctx.check(ctx.lib.jgb_fnr_setup(ctx.handle, system.n, P(i64(mdl.colptr), C.c_int64), P(i64(mdl.rowval), C.c_int64),
                                    P(cplx(mdl.nzval_t), C.c_double), P(i8(bus_type), C.c_int8), slack + 1,
                                    P(bp[0], C.c_int64), P(bp[1], C.c_int64), P(bp[2], C.c_double),
                                    P(bq[0], C.c_int64), P(bq[1], C.c_int64), P(bq[2], C.c_double)))
    _set_injection(a)
    a.voltage = Polar(vm, va)
    a._initial = (vm.copy(), va.copy())
    a._push()
    return a


def _set_injection(a: AcPowerFlowFast):
    s = a.system
    sp_, sq_, _ = s.supply
    a.ctx.check(a.ctx.lib.jgb_fnr_set_injection(a.ctx.handle, ptr(f64(sp_), C.c_double), ptr(f64(sq_), C.c_double),
                                                ptr(f64(s.pd), C.c_double), ptr(f64(s.qd), C.c_double)))


def fast_newton_raphson_bx(system: PowerSystem, ctx: Context | None = None, device: int = 0) -> AcPowerFlowFast:
    return _fast(system, True, ctx, device)


def fast_newton_raphson_xb(system: PowerSystem, ctx: Context | None = None, device: int = 0) -> AcPowerFlowFast:
    return _fast(system, False, ctx, device)


def mismatch_fnr(a: AcPowerFlowFast):
    if a._dirty:
        a._push()
    sp_, sq_ = C.c_double(0), C.c_double(0)
    a.ctx.check(a.ctx.lib.jgb_fnr_mismatch(a.ctx.handle, C.byref(sp_), C.byref(sq_)))
    return sp_.value, sq_.value


def solve_fnr(a: AcPowerFlowFast):
    if a._dirty:
        a._push()
    a.ctx.check(a.ctx.lib.jgb_fnr_solve(a.ctx.handle))
    a.method.iteration += 1
    a._pull()


def power_flow_fnr(a: AcPowerFlowFast, iteration: int = 20, tolerance: float = 1e-8) -> bool:
    if a._dirty:
        a._push()
    it, sp_, sq_ = C.c_int64(0), C.c_double(0), C.c_double(0)
    rc = a.ctx.check(a.ctx.lib.jgb_fnr_run(a.ctx.handle, iteration, tolerance, C.byref(it), C.byref(sp_), C.byref(sq_)))
    a.method.iteration = it.value
    a._pull()
    return rc == 0


def fnr_batch(a: AcPowerFlowFast, p_injection, q_injection, iteration: int = 20, tolerance: float = 1e-8):
    """p_injection, q_injection [R][n]: supply - demand of every bus and scenario; every scenario starts from the
    analysis' start point. Returns (magnitude [R][n], angle [R][n], iterations [R], status [R])."""
    p, q = f64(np.atleast_2d(p_injection)), f64(np.atleast_2d(q_injection))
    R, n = p.shape
    if n != a.system.n or q.shape != p.shape:
        raise ValueError("injection blocks must be [R][n]")
    a.voltage = Polar(a._initial[0].copy(), a._initial[1].copy())
    a._push()
    vm, va = np.empty((R, n)), np.empty((R, n))
    it, st, tot = np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int8), C.c_int64(0)
    a.ctx.check(a.ctx.lib.jgb_fnr_batch(a.ctx.handle, R, ptr(p, C.c_double), ptr(q, C.c_double), iteration, tolerance,
                                        ptr(vm, C.c_double), ptr(va, C.c_double), ptr(it, C.c_int32), ptr(st, C.c_int8),
                                        C.byref(tot)))
    _set_injection(a)          # back to the system's own injections for the single-case surface
    return vm, va, it, st


fastNewtonRaphsonBX = fast_newton_raphson_bx
fastNewtonRaphsonXB = fast_newton_raphson_xb
