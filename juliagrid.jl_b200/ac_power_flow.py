"""Host mirror of the reference's Newton-Raphson operator surface, backed by libjgb200.so.

    newton_raphson(system)        <-> newtonRaphson(system, B200)      src/powerFlow/acPowerFlow.jl:39-87
    mismatch(analysis)            <-> mismatch!(analysis)              :645-685
    solve(analysis)               <-> solve!(analysis)                 :793-911
    power_flow(analysis; ...)     <-> powerFlow!(analysis; ...)        :1389-1433
    set_initial_point(analysis)   <-> setInitialPoint!(analysis)       :1226-1249
    update_branch(analysis, k, status) <-> updateBranch!(analysis; label, status)  powerSystem/branch.jl:453-475

Argument meaning and error behaviour follow the reference: non-convergence is not an error (the last iterate stays in
`analysis.voltage`), a singular Jacobian raises. All numerics run in CUDA; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Context, ptr, f64, i64, i8, cplx
from .cases import PowerSystem
from .model import AcModel, ac_model, apply_branch_status


@dataclass
class Polar:
    magnitude: np.ndarray
    angle: np.ndarray


class NewtonRaphsonMethod:
    """analysis.method — host mirrors of NewtonRaphson{T} (src/definition/analysis.jl:154-164)."""

    def __init__(self):
        self.pq = self.pvpq = self.pcount = None           # 1-based Int64 like the reference
        self.jacobian_colptr = self.jacobian_rowval = None
        self.iteration = 0
        self.dim = 0


class AcPowerFlow:
    def __init__(self, system: PowerSystem, ctx: Context):
        self.system = system
        self.ctx = ctx
        self.voltage: Polar = None
        self.method = NewtonRaphsonMethod()
        self.bus_type = None
        self.slack = None
        self._state_dirty = True
        self._initial = None

    # vectors kept on the device; fetched on demand
    def _vectors(self):
        m = self.method
        f = np.empty(m.dim)
        inc = np.empty(m.dim)
        jv = np.empty(len(m.jacobian_rowval))
        it = C.c_int64(0)
        lib = self.ctx.lib
        self.ctx.check(lib.jgb_nr_get_vectors(self.ctx.handle, ptr(f, C.c_double), ptr(inc, C.c_double),
                                              ptr(jv, C.c_double), C.byref(it)))
        return f, inc, jv, it.value

    @property
    def mismatch(self):
        return self._vectors()[0]

    @property
    def increment(self):
        return self._vectors()[1]

    @property
    def jacobian_nzval(self):
        return self._vectors()[2]

    def _push_state(self):
        lib = self.ctx.lib
        self.ctx.check(lib.jgb_nr_set_state(self.ctx.handle, ptr(f64(self.voltage.magnitude), C.c_double),
                                            ptr(f64(self.voltage.angle), C.c_double)))
        self._state_dirty = False

    def _pull_state(self):
        lib = self.ctx.lib
        vm = np.empty(self.system.n)
        va = np.empty(self.system.n)
        self.ctx.check(lib.jgb_nr_get_state(self.ctx.handle, ptr(vm, C.c_double), ptr(va, C.c_double)))
        self.voltage = Polar(vm, va)


def _initialize(system: PowerSystem):
    """initializeACPowerFlow + changeSlackBus! (acPowerFlow.jl:1312-1358), on copies."""
    _, _, first_gen = system.supply
    has_gen = first_gen >= 0
    bus_type = system.bus_type.copy()
    bus_type[(~has_gen) & (bus_type == 2)] = 1
    vm = system.vm.copy()
    sel = has_gen & (bus_type != 1)
    vm[sel] = system.gen_vm[first_gen[sel]]
    slack = system.slack
    if not has_gen[slack]:
        bus_type[slack] = 1
        cand = np.flatnonzero((bus_type == 2) & has_gen)
        if len(cand) == 0:
            raise RuntimeError("The slack bus is missing.")
        slack = int(cand[0])
        bus_type[slack] = 3
    return bus_type, slack, vm, system.va.copy()


def newton_raphson(system: PowerSystem, ctx: Context | None = None, device: int = 0) -> AcPowerFlow:
    if system.model is None:
        system.model = ac_model(system)
    mdl: AcModel = system.model
    ctx = ctx or Context(device)
    lib = ctx.lib
    a = AcPowerFlow(system, ctx)
    bus_type, slack, vm, va = _initialize(system)
    a.bus_type, a.slack = bus_type, slack
    ycp, yrv = i64(mdl.colptr), i64(mdl.rowval)
    ctx.check(lib.jgb_nr_setup(ctx.handle, system.n, ptr(ycp, C.c_int64), ptr(yrv, C.c_int64),
                               ptr(cplx(mdl.nzval), C.c_double), ptr(cplx(mdl.nzval_t), C.c_double),
                               ptr(i8(bus_type), C.c_int8), slack + 1))
    dim, nnz = C.c_int64(0), C.c_int64(0)
    ctx.check(lib.jgb_nr_dims(ctx.handle, C.byref(dim), C.byref(nnz)))
    m = a.method
    m.dim = dim.value
    m.pq = np.empty(system.n, dtype=np.int64)
    m.pvpq = np.empty(system.n, dtype=np.int64)
    m.pcount = np.empty(system.n, dtype=np.int64)
    m.jacobian_colptr = np.empty(dim.value + 1, dtype=np.int64)
    m.jacobian_rowval = np.empty(nnz.value, dtype=np.int64)
    ctx.check(lib.jgb_nr_pattern(ctx.handle, ptr(m.pq, C.c_int64), ptr(m.pvpq, C.c_int64), ptr(m.pcount, C.c_int64),
                                 ptr(m.jacobian_colptr, C.c_int64), ptr(m.jacobian_rowval, C.c_int64)))
    sp, sq, _ = system.supply
    ctx.check(lib.jgb_nr_set_injection(ctx.handle, ptr(f64(sp), C.c_double), ptr(f64(sq), C.c_double),
                                       ptr(f64(system.pd), C.c_double), ptr(f64(system.qd), C.c_double)))
    a.voltage = Polar(vm, va)
    a._initial = (vm.copy(), va.copy())
    a._push_state()
    return a


def mismatch(a: AcPowerFlow):
    """mismatch!(analysis) -> (stopP, stopQ)."""
    if a._state_dirty:
        a._push_state()
    sp, sq = C.c_double(0), C.c_double(0)
    a.ctx.check(a.ctx.lib.jgb_nr_mismatch(a.ctx.handle, C.byref(sp), C.byref(sq)))
    return sp.value, sq.value


def solve(a: AcPowerFlow):
    """solve!(analysis): Jacobian fill, refactor, solve, V/theta update, iteration += 1."""
    if a._state_dirty:
        a._push_state()
    a.ctx.check(a.ctx.lib.jgb_nr_solve(a.ctx.handle))
    a.method.iteration += 1
    a._pull_state()


def power_flow(a: AcPowerFlow, iteration: int = 20, tolerance: float = 1e-8) -> bool:
    """powerFlow!(analysis; iteration, tolerance). Returns True when converged (the reference only prints it)."""
    if a._state_dirty:
        a._push_state()
    it, sp, sq = C.c_int64(0), C.c_double(0), C.c_double(0)
    rc = a.ctx.check(a.ctx.lib.jgb_nr_run(a.ctx.handle, iteration, tolerance, C.byref(it), C.byref(sp), C.byref(sq)))
    a.method.iteration = it.value
    a.last_stop = (sp.value, sq.value)
    a._pull_state()
    return rc == 0


_POWER_KEYS = ("injection_active", "injection_reactive", "from_active", "from_reactive", "to_active", "to_reactive",
               "from_current_magnitude", "from_current_angle", "to_current_magnitude", "to_current_angle")


def power_device(a: AcPowerFlow) -> dict:
    """power!(analysis) + current!(analysis) on the device (postprocessing/acAnalysis.jl:30-79, 672-700): bus
    injections and branch from/to flows and currents at the analysis' present state."""
    sysm, mdl, lib = a.system, a.system.model, a.ctx.lib
    if a._state_dirty:
        a._push_state()
    if not getattr(a, "_branches_set", False):
        a.ctx.check(lib.jgb_nr_set_branches(a.ctx.handle, sysm.nbr, ptr(i64(sysm.frm + 1), C.c_int64),
                                            ptr(i64(sysm.to + 1), C.c_int64), ptr(cplx(mdl.y_ff), C.c_double),
                                            ptr(cplx(mdl.y_ft), C.c_double), ptr(cplx(mdl.y_tf), C.c_double),
                                            ptr(cplx(mdl.y_tt), C.c_double), ptr(i8(sysm.status), C.c_int8)))
        a._branches_set = True
    out = {k: np.empty(sysm.n if k.startswith("injection") else sysm.nbr) for k in _POWER_KEYS}
    a.ctx.check(lib.jgb_nr_power(a.ctx.handle, *[ptr(out[k], C.c_double) for k in _POWER_KEYS]))
    return out


def generator_power(a: AcPowerFlow, pw: dict | None = None):
    """generatorPower for every generator (postprocessing/acAnalysis.jl:538-629) from the device's bus injections:
    returns (active, reactive). A bus's reactive output is shared between its in-service generators in proportion
    to their capability ranges; the first generator of the slack bus takes the active balance."""
    s = a.system
    pw = pw or power_device(a)
    inj_p, inj_q = pw["injection_active"], pw["injection_reactive"]
    on = np.flatnonzero(s.gen_status == 1)
    pg, qg = np.zeros(s.ngen), np.zeros(s.ngen)
    by_bus = {}
    for g in on:
        by_bus.setdefault(int(s.gen_bus[g]), []).append(int(g))
    eps = np.finfo(float).eps
    for b, gens in by_bus.items():
        qsum = inj_q[b] + s.qd[b]
        if len(gens) == 1:
            g = gens[0]
            pg[g] = inj_p[b] + s.pd[b] if b == a.slack else s.gen_p[g]
            qg[g] = qsum
            continue
        qmin, qmax = s.gen_qmin[gens].copy(), s.gen_qmax[gens].copy()
        fin_min, fin_max = qmin[~np.isinf(qmin)].sum(), qmax[~np.isinf(qmax)].sum()
        big = abs(qsum) + abs(fin_min) + abs(fin_max)
        qmin = np.where(np.isinf(qmin), np.where(qmin > 0, big, -big), qmin)
        qmax = np.where(np.isinf(qmax), np.where(qmax < 0, -big, big), qmax)
        smin, smax = qmin.sum(), qmax.sum()
        if s.base_mva * abs(smin - smax) > 10 * eps:
            qg[gens] = qmin + ((qsum - smin) / (smax - smin)) * (qmax - qmin)
        else:
            qg[gens] = qmin + (qsum - smin) / len(gens)
        pg[gens] = s.gen_p[gens]
        if b == a.slack:
            pg[gens[0]] = inj_p[b] + s.pd[b] - s.gen_p[gens[1:]].sum()
    return pg, qg


def reactive_limit(a: AcPowerFlow, pw: dict | None = None) -> np.ndarray:
    """reactiveLimit!(analysis) (acPowerFlow.jl:1081-1156). Mutates `a.system` like the reference: violating
    generators are fixed at their limit and their bus becomes a demand bus (a converted slack hands over to the first
    generator bus); returns the flags (-1 / +1). Build a new analysis with newton_raphson(system) afterwards."""
    s = a.system
    pg, qg = generator_power(a, pw)
    s.bus_type = a.bus_type.copy()
    s.slack = a.slack
    on = s.gen_status == 1
    s.gen_p = np.where(on, pg, s.gen_p)
    sp, sq = np.zeros(s.n), np.zeros(s.n)
    np.add.at(sp, s.gen_bus[on], pg[on])
    np.add.at(sq, s.gen_bus[on], qg[on])
    violate = np.zeros(s.ngen, dtype=np.int64)
    for i in np.flatnonzero(on & (s.gen_qmin < s.gen_qmax)):
        j = int(s.gen_bus[i])
        low, high = qg[i] < s.gen_qmin[i], qg[i] > s.gen_qmax[i]
        if s.bus_type[j] != 1 and (low or high):
            violate[i] = 1 if high else -1
            new_q = s.gen_qmax[i] if high else s.gen_qmin[i]
            s.bus_type[j] = 1
            sq[j] += new_q - qg[i]
            s.gen_q[i] = new_q
            if j == s.slack:
                cand = np.flatnonzero(s.bus_type == 2)
                if len(cand):
                    s.slack = int(cand[0])
                    s.bus_type[s.slack] = 3
    if s.bus_type[s.slack] != 3:
        raise RuntimeError("The slack bus is missing.")
    s.supply_p, s.supply_q = sp, sq
    return violate


def adjust_angle(a: AcPowerFlow, slack: int):
    """adjustAngle!(analysis; slack) (acPowerFlow.jl:1186-1196); slack is a 0-based bus index."""
    a.voltage.angle = a.voltage.angle + (a.system.va[slack] - a.voltage.angle[slack])
    a._state_dirty = True


def set_initial_point(a: AcPowerFlow):
    """setInitialPoint!(analysis): back to the start point of the constructor."""
    a.voltage = Polar(a._initial[0].copy(), a._initial[1].copy())
    a._state_dirty = True


def set_voltage(a: AcPowerFlow, magnitude, angle):
    a.voltage = Polar(np.array(magnitude, dtype=float), np.array(angle, dtype=float))
    a._state_dirty = True


def update_branch(a: AcPowerFlow, k: int, status: int):
    """updateBranch!(analysis; label = k, status): in-place Ybus value update on the fixed pattern
    (updateBranchMain! branch.jl:313-431 + acNodalUpdate! model.jl:81-110). k is the 0-based branch index."""
    if status == int(a.system.status[k]):
        return
    pos, yv, ytv, _ = apply_branch_status(a.system, k, status)
    p1 = i64(pos + 1)
    a.ctx.check(a.ctx.lib.jgb_nr_update_y(a.ctx.handle, 4, ptr(p1, C.c_int64), ptr(cplx(yv), C.c_double),
                                          ptr(cplx(ytv), C.c_double)))
    a._branches_set = False       # Y-parameters / statuses changed: re-upload before the next power_device


# camelCase aliases matching the reference's exported names
newtonRaphson = newton_raphson
powerFlow = power_flow
setInitialPoint = set_initial_point
updateBranch = update_branch
reactiveLimit = reactive_limit
adjustAngle = adjust_angle
generatorPower = generator_power
reactiveLimit = reactive_limit
adjustAngle = adjust_angle
generatorPower = generator_power
