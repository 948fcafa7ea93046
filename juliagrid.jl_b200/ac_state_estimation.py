"""Host mirror of the reference's Gauss-Newton WLS operator surface, backed by libjgb200.so.

    gauss_newton(monitoring)         <-> gaussNewton(monitoring, B200)   src/stateEstimation/acStateEstimation.jl:43-75
    increment(analysis)              <-> increment!(analysis)            :878-904
    solve_se(analysis)               <-> solve!(analysis)                :1035-1047
    state_estimation(analysis; ...)  <-> stateEstimation!(analysis; ...) :1286-1329
    set_mean(analysis, z)            <-> update*!(analysis; ...) value updates (measurement/*.jl), pattern fixed
    chi_test(analysis)               <-> chiTest(analysis)               src/stateEstimation/badData.jl (chi-square part)
    residual_test(analysis)          <-> residualTest!(analysis)         src/stateEstimation/badData.jl:181-285
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from ._lib import Context, ptr, f64, i64, i8, cplx
from .ac_power_flow import Polar
from .measurement import Measurement, WlsTables, ac_wls
from .model import ac_model


class GaussNewtonMethod:
    """analysis.method — host mirrors of GaussNewton{T} (src/definition/analysis.jl:532-545)."""

    def __init__(self):
        self.tables: WlsTables = None
        self.mean = None
        self.type = None
        self.index = None
        self.range = None
        self.gain_colptr = self.gain_rowval = None
        self.objective = 0.0
        self.iteration = 0


class AcStateEstimation:
    def __init__(self, monitoring: Measurement, ctx: Context):
        self.monitoring = monitoring
        self.system = monitoring.system
        self.ctx = ctx
        self.voltage: Polar = None
        self.method = GaussNewtonMethod()
        self._dirty = True

    def _vectors(self):
        t = self.method.tables
        n = self.system.n
        res, inc = np.empty(t.m), np.empty(2 * n)
        hv = np.empty(len(t.h_rowval))
        gv = np.empty(len(self.method.gain_rowval))
        it = C.c_int64(0)
        self.ctx.check(self.ctx.lib.jgb_wls_get_vectors(self.ctx.handle, ptr(res, C.c_double), ptr(inc, C.c_double),
                                                        ptr(hv, C.c_double), ptr(gv, C.c_double), C.byref(it)))
        return res, inc, hv, gv

    residual = property(lambda self: self._vectors()[0])
    increment = property(lambda self: self._vectors()[1])
    jacobian_nzval = property(lambda self: self._vectors()[2])
    gain_nzval = property(lambda self: self._vectors()[3])

    def _push(self):
        self.ctx.check(self.ctx.lib.jgb_wls_set_state(self.ctx.handle, ptr(f64(self.voltage.magnitude), C.c_double),
                                                      ptr(f64(self.voltage.angle), C.c_double)))
        self._dirty = False

    def _pull(self):
        vm, va = np.empty(self.system.n), np.empty(self.system.n)
        self.ctx.check(self.ctx.lib.jgb_wls_get_state(self.ctx.handle, ptr(vm, C.c_double), ptr(va, C.c_double)))
        self.voltage = Polar(vm, va)


def gauss_newton(monitoring: Measurement, ctx: Context | None = None, device: int = 0) -> AcStateEstimation:
    system = monitoring.system
    if system.model is None:
        system.model = ac_model(system)
    mdl = system.model
    ctx = ctx or Context(device)
    a = AcStateEstimation(monitoring, ctx)
    t = ac_wls(system, monitoring)
    lib = ctx.lib
    P = lambda v, ct: ptr(v, ct)
    hcp, hrv, idx, rng = i64(t.h_colptr), i64(t.h_rowval), i64(t.index), i64(t.range)
    wcp, wrv, wnz = i64(t.w_colptr), i64(t.w_rowval), f64(t.w_nzval)
    ycp, yrv = i64(mdl.colptr), i64(mdl.rowval)
    frm, to = i64(system.frm + 1), i64(system.to + 1)
    ctx.check(lib.jgb_wls_setup(
        ctx.handle, system.n, t.m, system.slack + 1, P(hcp, C.c_int64), P(hrv, C.c_int64), P(i8(t.type), C.c_int8),
        P(idx, C.c_int64), P(rng, C.c_int64), P(wcp, C.c_int64), P(wrv, C.c_int64), P(wnz, C.c_double),
        P(ycp, C.c_int64), P(yrv, C.c_int64), P(cplx(mdl.nzval), C.c_double), P(cplx(mdl.nzval_t), C.c_double),
        system.nbr, P(frm, C.c_int64), P(to, C.c_int64), P(f64(system.g), C.c_double), P(f64(system.b), C.c_double),
        P(f64(system.tap), C.c_double), P(f64(system.shift), C.c_double), P(cplx(mdl.admittance), C.c_double)))
    nh, ng = C.c_int64(0), C.c_int64(0)
    ctx.check(lib.jgb_wls_dims(ctx.handle, C.byref(nh), C.byref(ng)))
    me = a.method
    me.tables, me.mean, me.type, me.index, me.range = t, t.mean.copy(), t.type, t.index, t.range
    me.gain_colptr = np.empty(2 * system.n + 1, dtype=np.int64)
    me.gain_rowval = np.empty(ng.value, dtype=np.int64)
    ctx.check(lib.jgb_wls_gain_pattern(ctx.handle, P(me.gain_colptr, C.c_int64), P(me.gain_rowval, C.c_int64)))
    ctx.check(lib.jgb_wls_set_mean(ctx.handle, P(f64(me.mean), C.c_double)))
    # GN starts from bus.voltage.{magnitude, angle} (acStateEstimation.jl:52-55), not generator set-points
    a.voltage = Polar(system.vm.copy(), system.va.copy())
    a._push()
    return a


def increment(a: AcStateEstimation) -> float:
    """increment!(analysis) -> maximum(abs, increment); sets method.objective."""
    if a._dirty:
        a._push()
    mi, ob = C.c_double(0), C.c_double(0)
    a.ctx.check(a.ctx.lib.jgb_wls_increment(a.ctx.handle, C.byref(mi), C.byref(ob)))
    a.method.objective = ob.value
    return mi.value


def solve_se(a: AcStateEstimation):
    """solve!(analysis)."""
    a.ctx.check(a.ctx.lib.jgb_wls_solve(a.ctx.handle))
    a.method.iteration += 1
    a._pull()


def state_estimation(a: AcStateEstimation, iteration: int = 40, tolerance: float = 1e-8) -> bool:
    """stateEstimation!(analysis; iteration, tolerance); True when converged."""
    if a._dirty:
        a._push()
    it, mi, ob = C.c_int64(0), C.c_double(0), C.c_double(0)
    rc = a.ctx.check(a.ctx.lib.jgb_wls_run(a.ctx.handle, iteration, tolerance, C.byref(it), C.byref(mi), C.byref(ob)))
    a.method.iteration = it.value
    a.method.objective = ob.value
    a.last_increment = mi.value
    a._pull()
    return rc == 0


def set_mean(a: AcStateEstimation, z):
    a.method.mean = np.array(z, dtype=float)
    a.ctx.check(a.ctx.lib.jgb_wls_set_mean(a.ctx.handle, ptr(f64(a.method.mean), C.c_double)))


def set_voltage_se(a: AcStateEstimation, magnitude, angle):
    a.voltage = Polar(np.array(magnitude, dtype=float), np.array(angle, dtype=float))
    a._dirty = True


class ChiTest:
    def __init__(self, detect, threshold, objective):
        self.detect, self.threshold, self.objective = detect, threshold, objective


def chi_test(a: AcStateEstimation, confidence: float = 0.95) -> ChiTest:
    """chiTest(analysis; confidence): WLS objective against the chi-square quantile, m_in_service - (2n - 1) dof."""
    import scipy.stats
    dof = int(np.count_nonzero(a.method.type)) - (2 * a.system.n - 1)
    thr = float(scipy.stats.chi2.ppf(confidence, dof))
    return ChiTest(a.method.objective > thr, thr, a.method.objective)


class ResidualTest:
    """bad = ResidualTest(detect, maxNormalizedResidual, label, index) (src/definition/analysis.jl); `label` is
    (device class, 0-based device index) here — labels are a host-side naming layer of the reference."""

    def __init__(self, detect, max_normalized_residual, label, index):
        self.detect, self.maxNormalizedResidual, self.label, self.index = detect, max_normalized_residual, label, index


def residual_test(a: AcStateEstimation, threshold: float = 3.0) -> ResidualTest:
    """residualTest!(analysis; threshold): numeric part on the device (selected inverse of the gain factor, row
    projection), status bookkeeping of badData.jl:224-282 here. index is the 0-based row, -1 when none."""
    rn, idx = C.c_double(0), C.c_int64(0)
    a.ctx.check(a.ctx.lib.jgb_wls_residual_test(a.ctx.handle, threshold, C.byref(rn), C.byref(idx), None))
    me, mon = a.method, a.monitoring
    row = idx.value - 1
    detect = rn.value > threshold
    label = None
    if row >= 0:
        r = me.range - 1                       # 0-based block starts: volt | amp | watt | var | pmu
        nv, na, nw, nq = (len(getattr(mon, d)["index"]) for d in ("volt", "amp", "watt", "var"))
        rows = [row]
        if row < r[1]:
            label = ("voltmeter", row)
            if detect:
                mon.volt["status"][row] = 0
        elif row < r[2]:
            label = ("ammeter", row - nv)
            if detect:
                mon.amp["status"][row - nv] = 0
        elif row < r[3]:
            label = ("wattmeter", row - nv - na)
            if detect:
                mon.watt["status"][row - nv - na] = 0
        elif row < r[4]:
            label = ("varmeter", row - nv - na - nw)
            if detect:
                mon.var["status"][row - nv - na - nw] = 0
        else:
            local = row - nv - na - nw - nq
            k = local // 2
            label = ("pmu", k)
            if detect:
                if mon.pmu["polar"][k]:
                    if int(me.type[row]) in (2, 3, 4, 5, 12):
                        mon.pmu["mag_status"][k] = 0
                    else:
                        mon.pmu["ang_status"][k] = 0
                else:
                    mon.pmu["mag_status"][k] = 0
                    mon.pmu["ang_status"][k] = 0
                    rows.append(row + 1 if local % 2 == 0 else row - 1)
        if detect:
            me.type = np.array(me.type, copy=True)
            for q in rows:
                a.ctx.check(a.ctx.lib.jgb_wls_remove_row(a.ctx.handle, q + 1))
                me.mean[q] = 0.0
                me.type[q] = 0
            me.iteration = 0
    return ResidualTest(detect, rn.value, label, row)


gaussNewton = gauss_newton
chiTest = chi_test
residualTest = residual_test
stateEstimation = state_estimation


# ---- update*!(analysis; ...): in-place changes of single meters, reusing the gain pattern and its factorisation --------
def _w_diag_off(t: WlsTables):
    """Diagonal of the precision matrix and W[row, row-1] per row from the CSC tables."""
    W = sp.csc_matrix((t.w_nzval, t.w_rowval - 1, t.w_colptr - 1), shape=(t.m, t.m))
    d = W.diagonal()
    off = np.zeros(t.m)
    sub = W.diagonal(-1)
    off[1:] = sub
    return d, off


def _sync_rows(a: AcStateEstimation):
    """Rebuild the acWLS tables from the (edited) monitoring and push every row whose mean / type / index / precision
    changed through jgb_wls_update_rows — what the reference's _update*!(analysis, idx) do row by row."""
    old = a.method.tables
    new = ac_wls(a.system, a.monitoring)
    if new.m != old.m or not np.array_equal(new.h_rowval, old.h_rowval) or not np.array_equal(new.h_colptr, old.h_colptr):
        raise ValueError("the update changes the measurement layout: build a new analysis (gauss_newton)")
    d0, o0 = _w_diag_off(old)
    d1, o1 = _w_diag_off(new)
    # the analysis may hold means set with set_mean (Monte-Carlo draws): only rows whose table entries changed move
    ch = (new.mean != old.mean) | (new.type != old.type) | (new.index != old.index) | (d0 != d1) | (o0 != o1)
    rows = np.flatnonzero(ch)
    if len(rows) == 0:
        return rows
    r1 = i64(rows + 1)
    a.ctx.check(a.ctx.lib.jgb_wls_update_rows(a.ctx.handle, len(rows), ptr(r1, C.c_int64),
                                              ptr(f64(new.mean[rows]), C.c_double), ptr(f64(d1[rows]), C.c_double),
                                              ptr(f64(o1[rows]), C.c_double), ptr(i8(new.type[rows]), C.c_int8),
                                              ptr(i64(new.index[rows]), C.c_int64)))
    me = a.method
    me.mean[rows] = new.mean[rows]
    me.tables, me.type, me.index = new, new.type, new.index
    return rows


def _update(a: AcStateEstimation, dev: str, k: int, **fields):
    d = getattr(a.monitoring, dev)
    if not 0 <= k < len(d["index"]):
        raise IndexError(f"{dev} {k} does not exist")
    for key, val in fields.items():
        if val is not None:
            d[key][k] = val
    return _sync_rows(a)


def update_voltmeter(a: AcStateEstimation, k: int, magnitude=None, variance=None, status=None):
    """updateVoltmeter!(analysis; label, magnitude, variance, status) (measurement/voltmeter.jl)."""
    return _update(a, "volt", k, mean=magnitude, variance=variance, status=status)


def update_ammeter(a: AcStateEstimation, k: int, magnitude=None, variance=None, status=None, square=None):
    """updateAmmeter!(analysis; ...) (measurement/ammeter.jl)."""
    return _update(a, "amp", k, mean=magnitude, variance=variance, status=status, square=square)


def update_wattmeter(a: AcStateEstimation, k: int, active=None, variance=None, status=None):
    """updateWattmeter!(analysis; label, active, variance, status) (measurement/powermeter.jl:608-677)."""
    return _update(a, "watt", k, mean=active, variance=variance, status=status)


def update_varmeter(a: AcStateEstimation, k: int, reactive=None, variance=None, status=None):
    """updateVarmeter!(analysis; ...) (measurement/powermeter.jl)."""
    return _update(a, "var", k, mean=reactive, variance=variance, status=status)


def update_pmu(a: AcStateEstimation, k: int, magnitude=None, angle=None, variance_magnitude=None, variance_angle=None,
               status_magnitude=None, status_angle=None, polar=None, correlated=None, square=None):
    """updatePmu!(analysis; ...) (measurement/pmu.jl). Switching `correlated` on for a PMU that was built without it
    adds an off-diagonal precision entry, i.e. a new gain pattern: the library answers -4 and a new analysis is needed."""
    return _update(a, "pmu", k, mag_mean=magnitude, ang_mean=angle, mag_variance=variance_magnitude,
                   ang_variance=variance_angle, mag_status=status_magnitude, ang_status=status_angle, polar=polar,
                   correlated=correlated, square=square)


def update_branch_se(a: AcStateEstimation, k: int, status: int):
    """updateBranch!(analysis; label, status) on a state-estimation analysis (powerSystem/branch.jl:453-475): the Ybus
    values the injection rows read and the parameters the flow rows of the branch read change in place; meters on a
    branch taken out of service must be switched off by the caller (update_*meter(status=0)), as in the reference."""
    from .model import apply_branch_status
    s = a.system
    pos, y, yt, adm = apply_branch_status(s, k, status)
    a.ctx.check(a.ctx.lib.jgb_wls_update_y(a.ctx.handle, len(pos), ptr(i64(pos + 1), C.c_int64), ptr(cplx(y), C.c_double),
                                           ptr(cplx(yt), C.c_double)))
    a.ctx.check(a.ctx.lib.jgb_wls_update_branch(a.ctx.handle, k + 1, float(s.g[k]), float(s.b[k]), float(s.tap[k]),
                                                float(s.shift[k]), ptr(cplx(np.array([adm])), C.c_double)))
