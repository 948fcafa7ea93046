"""update*!(analysis; ...) on a Gauss-Newton analysis (jgb_wls_update_rows / _update_y / _update_branch): after every
in-place change the updated analysis must equal an analysis built from scratch on the edited monitoring — the property
`testReusing` asserts in the reference (test/utility/utility.jl:325-360, driven by test/stateEstimation/reusing.jl) —
and the CPU oracle built from the same monitoring."""
import copy

import numpy as np
import pytest

import jgb200
from oracle import wls as owls
from conftest import oracle_system
from test_wls_gpu import _truth, _modified14, _everything

pytestmark = pytest.mark.gpu


def _check(a, mon, os_, o, ctx2):
    """testReusing: one increment from the start point gives the same type / mean / Jacobian / increment as a fresh
    analysis and as the oracle; then both converge to the same state."""
    fresh = jgb200.gauss_newton(mon, ctx2)
    jgb200.set_voltage_se(a, a.system.vm, a.system.va)
    mi_a, mi_f = jgb200.increment(a), jgb200.increment(fresh)
    assert np.array_equal(a.method.tables.type, fresh.method.tables.type)
    np.testing.assert_allclose(a.jacobian_nzval, fresh.jacobian_nzval, atol=1e-10, rtol=0)
    np.testing.assert_allclose(a.increment, fresh.increment, atol=1e-10, rtol=1e-9)
    np.testing.assert_allclose(a.residual, fresh.residual, atol=1e-12, rtol=0)
    assert mi_a == pytest.approx(mi_f, rel=1e-9, abs=1e-12)
    g = owls.gauss_newton(os_, mon, o.mdl)
    omi = owls.increment(g)
    np.testing.assert_allclose(a.increment, g.increment, rtol=1e-6, atol=1e-9 * max(1.0, omi))
    assert a.method.objective == pytest.approx(g.objective, rel=1e-9)
    jgb200.set_voltage_se(a, a.system.vm, a.system.va)
    ok_a = jgb200.state_estimation(a, iteration=100, tolerance=1e-10)
    ok_f = jgb200.state_estimation(fresh, iteration=100, tolerance=1e-10)
    assert ok_a == ok_f and a.method.iteration == fresh.method.iteration
    np.testing.assert_allclose(a.voltage.magnitude, fresh.voltage.magnitude, atol=1e-10)
    np.testing.assert_allclose(a.voltage.angle, fresh.voltage.angle, atol=1e-10)


def test_meter_updates_reuse_the_factorisation(ctx):
    ps, os_, o, pw = _truth("case14test", _modified14)
    mon = _everything(ps, o, pw)
    a = jgb200.gauss_newton(mon, ctx)
    ctx2 = jgb200.Context(0)
    nv = len(mon.volt["index"])
    # voltmeter (reusing.jl:42-61): huge variance, out of service, back with a new mean, new variance
    assert len(jgb200.update_voltmeter(a, 0, magnitude=2.6, variance=1e30)) == 1
    _check(a, mon, os_, o, ctx2)
    jgb200.update_voltmeter(a, 0, variance=1e-4, status=0)
    assert a.method.tables.type[0] == 0
    _check(a, mon, os_, o, ctx2)
    jgb200.update_voltmeter(a, 0, magnitude=float(o.vm[0]), status=1)
    assert a.method.tables.type[0] == 1
    _check(a, mon, os_, o, ctx2)
    # from-end ammeter (reusing.jl:64-96): magnitude / square toggles change the type code 2 <-> 4
    k = 0
    jgb200.update_ammeter(a, k, magnitude=3.0, variance=1e30)
    _check(a, mon, os_, o, ctx2)
    jgb200.update_ammeter(a, k, variance=1e-3, status=0)
    _check(a, mon, os_, o, ctx2)
    jgb200.update_ammeter(a, k, magnitude=float(pw["from_current_magnitude"][mon.amp["index"][k]]), status=1, square=True)
    assert a.method.tables.type[nv + k] == 4
    _check(a, mon, os_, o, ctx2)
    jgb200.update_ammeter(a, k, square=False)
    assert a.method.tables.type[nv + k] == 2
    _check(a, mon, os_, o, ctx2)
    # bus wattmeter out and in again (a row that was out of service at construction comes back: slot lists rebuilt)
    assert mon.watt["status"][5] == 0
    rows = jgb200.update_wattmeter(a, 5, status=1)
    assert len(rows) == 1 and a.method.tables.type[rows[0]] == 6
    _check(a, mon, os_, o, ctx2)
    jgb200.update_wattmeter(a, 2, active=5.3, variance=1e45)
    _check(a, mon, os_, o, ctx2)
    jgb200.update_wattmeter(a, 2, active=float(pw["injection_active"][2]), variance=1e-4)
    # branch varmeter: to-end meter off / on
    kq = int(np.flatnonzero(~mon.var["bus"] & ~mon.var["frm"])[0])
    jgb200.update_varmeter(a, kq, status=0)
    _check(a, mon, os_, o, ctx2)
    jgb200.update_varmeter(a, kq, status=1, variance=1e-2)
    _check(a, mon, os_, o, ctx2)
    # PMUs: polar bus phasor magnitude off, rectangular correlated pair with new variances (2x2 precision block)
    jgb200.update_pmu(a, 3, status_magnitude=1)
    _check(a, mon, os_, o, ctx2)
    kc = int(np.flatnonzero(mon.pmu["correlated"])[0])
    jgb200.update_pmu(a, kc, variance_magnitude=1e-6, variance_angle=1e-5)
    _check(a, mon, os_, o, ctx2)
    jgb200.update_pmu(a, kc, status_magnitude=0)
    _check(a, mon, os_, o, ctx2)
    # a new off-diagonal precision entry is a new gain pattern: refused, the analysis keeps working
    ku = int(np.flatnonzero(~mon.pmu["polar"] & ~mon.pmu["correlated"])[0])
    before = copy.deepcopy(mon.pmu)
    with pytest.raises(jgb200.JgbError) as e:
        jgb200.update_pmu(a, ku, correlated=True)
    assert e.value.rc == -4
    mon.pmu = before
    ctx2.close()


def test_update_rows_rejects_bad_arguments(ctx):
    import ctypes as C
    from jgb200._lib import ptr, i64, f64, i8
    ps, os_, o, pw = _truth("case14test", _modified14)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_wattmeter(mon, pw)
    a = jgb200.gauss_newton(mon, ctx)
    lib, h = ctx.lib, ctx.handle
    r = i64([0])
    assert lib.jgb_wls_update_rows(h, 1, ptr(r, C.c_int64), None, None, None, None, None) == -1       # row 0
    r = i64([10 ** 6])
    assert lib.jgb_wls_update_rows(h, 1, ptr(r, C.c_int64), None, None, None, None, None) == -1
    r = i64([1])
    t = i8([33])
    assert lib.jgb_wls_update_rows(h, 1, ptr(r, C.c_int64), None, None, None, ptr(t, C.c_int8), None) == -1
    t = i8([7])                                     # a voltmeter row has no branch-flow entries in the H pattern
    assert lib.jgb_wls_update_rows(h, 1, ptr(r, C.c_int64), None, None, None, ptr(t, C.c_int8), None) == -4
    assert lib.jgb_wls_update_rows(h, 0, None, None, None, None, None, None) == -1
    assert lib.jgb_wls_update_y(h, 1, None, None, None) == -1
    assert lib.jgb_wls_update_branch(h, 0, 0.0, 0.0, 1.0, 0.0, ptr(f64([0, 0]), C.c_double)) == -1
    assert jgb200.state_estimation(a)               # still usable


def test_branch_outage_on_an_estimation_reuses_the_pattern(ctx):
    """updateBranch!(analysis; status = 0) on a WLS analysis: Ybus values and the branch's parameters change in place;
    the meters on the branch are switched off like the reference does (branch.jl:384-431)."""
    ps, os_, o, pw = _truth("case14test", _modified14)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    a = jgb200.gauss_newton(mon, ctx)
    k = 4
    jgb200.update_branch_se(a, k, 0)
    for dev, upd in ((mon.watt, jgb200.update_wattmeter), (mon.var, jgb200.update_varmeter)):
        for q in np.flatnonzero(~dev["bus"] & (dev["index"] == k)):
            upd(a, int(q), status=0)
    # truth of the outaged system for the remaining meters
    from oracle import nr as onr
    os2 = os_.copy()
    os2.status[k] = 0
    o2 = onr.newton_raphson(os2)
    assert onr.power_flow(o2)
    pw2 = jgb200.power(ps, o2.vm, o2.va)
    nb = ps.n
    for dev, key, upd, arg in ((mon.watt, "active", jgb200.update_wattmeter, "active"),
                               (mon.var, "reactive", jgb200.update_varmeter, "reactive")):
        for q in range(len(dev["index"])):
            if dev["status"][q] == 0:
                continue
            idx = int(dev["index"][q])
            val = pw2["injection_" + key][idx] if dev["bus"][q] else (
                pw2["from_" + key][idx] if dev["frm"][q] else pw2["to_" + key][idx])
            upd(a, q, **{arg: float(val)})
    for q in range(nb):
        jgb200.update_voltmeter(a, q, magnitude=float(o2.vm[q]))
    assert jgb200.state_estimation(a, tolerance=1e-10)
    np.testing.assert_allclose(a.voltage.magnitude, o2.vm, atol=1e-8)
    np.testing.assert_allclose(a.voltage.angle, o2.va, atol=1e-8)
