"""Bad-data post-step on the device: selected inverse of the gain factor + row projection + largest
normalised residual, against the reference's known answers (test/stateEstimation/badData.jl) and the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import jgb200
from oracle import nr as onr, wls as owls
from conftest import oracle_system, product_system
from test_oracle_golden import _bad_data_case
from test_linear_cpu import _product_monitoring

pytestmark = pytest.mark.gpu


def _product_case(two_outliers):
    so, o, me = _bad_data_case(two_outliers)
    ps = product_system("case14test")
    ps.bus_type[0], ps.bus_type[2], ps.slack, ps.va[2] = 2, 3, 2, -0.17
    ps.model = jgb200.ac_model(ps)
    return so, o, me, _product_monitoring(ps, me)


def _projection(se):
    c = np.empty(se.method.tables.m)
    rn, idx = C.c_double(0), C.c_int64(0)
    se.ctx.check(se.ctx.lib.jgb_wls_residual_test(se.ctx.handle, 3.0, C.byref(rn), C.byref(idx),
                                                  c.ctypes.data_as(C.POINTER(C.c_double))))
    return c, rn.value, idx.value - 1


def test_one_outlier(ctx):
    """badData.jl:24-41: chi-square 3227.3 / 109.7, r_N 52.5 at Varmeter 4, recovery 1e-10 after removal."""
    so, o, me, mon = _product_case(False)
    se = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(se)
    chi = jgb200.chi_test(se)
    assert chi.detect and abs(chi.threshold - 109.7) < 0.1 and abs(chi.objective - 3227.3) < 0.1
    og = owls.gauss_newton(so, me, o.mdl)
    assert owls.state_estimation(og)
    c, rn, idx = _projection(se)
    np.testing.assert_allclose(c, owls.residual_projection(og), rtol=1e-9, atol=1e-12)
    bad = jgb200.residual_test(se, threshold=3.0)
    assert bad.detect and abs(bad.maxNormalizedResidual - 52.5) < 0.1
    assert bad.label == ("varmeter", 3) and bad.index == idx
    assert mon.var["status"][3] == 0
    assert jgb200.state_estimation(se)
    assert np.abs(se.voltage.magnitude - o.vm).max() < 1e-10 and np.abs(se.voltage.angle - o.va).max() < 1e-10
    again = jgb200.residual_test(se)
    assert not again.detect


def test_two_outliers(ctx):
    """badData.jl:58-84: PMU 10 (7713.26) first, then Varmeter 4 (78.3), then recovery."""
    so, o, me, mon = _product_case(True)
    se = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(se)
    bad = jgb200.residual_test(se)
    assert bad.detect and bad.label == ("pmu", 9) and abs(bad.maxNormalizedResidual - 7713.26) < 0.1
    assert mon.pmu["mag_status"][9] == 0 and mon.pmu["ang_status"][9] == 1
    assert jgb200.state_estimation(se)
    bad = jgb200.residual_test(se)
    assert bad.detect and bad.label == ("varmeter", 3) and abs(bad.maxNormalizedResidual - 78.3) < 0.1
    assert jgb200.state_estimation(se)
    assert np.abs(se.voltage.magnitude - o.vm).max() < 1e-10 and np.abs(se.voltage.angle - o.va).max() < 1e-10


def test_rectangular_pmu_outlier_removes_both_rows(ctx):
    """badData.jl:118-135 shape: a rectangular current phasor with a gross magnitude error takes both of its rows out."""
    so, o, me, _ = _product_case(False)
    ps = product_system("case14test")
    ps.bus_type[0], ps.bus_type[2], ps.slack, ps.va[2] = 2, 3, 2, -0.17
    ps.model = jgb200.ac_model(ps)
    pw = jgb200.power(ps, o.vm, o.va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm, variance=1e-2)
    jgb200.add_wattmeter(mon, pw, variance=1e-2)
    jgb200.add_varmeter(mon, pw, variance=1e-2)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=(), branch=True, polar=False, variance_magnitude=1e-5, variance_angle=1e-5)
    mon.pmu["mag_mean"][4] = 30.0
    se = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(se)
    bad = jgb200.residual_test(se)
    assert bad.detect and bad.label == ("pmu", 4)
    first = int(se.method.range[4] - 1) + 8
    assert se.method.type[first] == 0 and se.method.type[first + 1] == 0
    assert jgb200.state_estimation(se)
    assert np.abs(se.voltage.magnitude - o.vm).max() < 1e-9 and np.abs(se.voltage.angle - o.va).max() < 1e-9


def test_projection_10k_sample(ctx):
    """config-3 measurement set on ACTIVSg10k with one gross error: the device's c and r_N against the oracle on a
    sample of rows (one SuperLU solve per row), and the outlier is found."""
    so, ps = oracle_system("case_ACTIVSg10k"), product_system("case_ACTIVSg10k")
    ps.model = jgb200.ac_model(ps)
    o = onr.newton_raphson(so)
    assert onr.power_flow(o)
    pw = jgb200.power(ps, o.vm, o.va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    buses = np.sort(np.random.default_rng(7).choice(ps.n, ps.n // 10, replace=False))
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=buses, polar=False)
    outlier = ps.n + 1234                         # a wattmeter row
    mon.watt["mean"][1234] += 5.0
    se = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(se)
    c, rn, idx = _projection(se)
    assert idx == outlier and rn > 100
    og = owls.gauss_newton(so, mon, o.mdl)
    og.vm, og.va = se.voltage.magnitude.copy(), se.voltage.angle.copy()
    owls.normal_equation(og)
    rows = np.r_[np.arange(0, og.m, og.m // 40), outlier]
    want = owls.residual_projection_rows(og, rows)
    np.testing.assert_allclose(c[rows], want, rtol=1e-6, atol=1e-12)
    bad = jgb200.residual_test(se)
    assert bad.detect and bad.label == ("wattmeter", 1234)
    assert jgb200.state_estimation(se)
    assert np.abs(se.voltage.magnitude - o.vm).max() < 1e-8 and np.abs(se.voltage.angle - o.va).max() < 1e-8


def test_residual_test_needs_a_solved_estimation_and_valid_rows(ctx):
    fresh = jgb200.Context(0)
    rn, idx = C.c_double(0), C.c_int64(0)
    assert fresh.lib.jgb_wls_residual_test(fresh.handle, 3.0, C.byref(rn), C.byref(idx), None) == -1   # no setup
    so, o, me, mon = _product_case(False)
    se = jgb200.gauss_newton(mon, fresh)
    assert fresh.lib.jgb_wls_remove_row(fresh.handle, 0) == -1
    assert fresh.lib.jgb_wls_remove_row(fresh.handle, se.method.tables.m + 1) == -1
    assert jgb200.state_estimation(se)
    bad = jgb200.residual_test(se, threshold=1e6)          # nothing exceeds the threshold: nothing is removed
    assert not bad.detect and bad.index >= 0 and mon.var["status"][3] == 1
    assert np.count_nonzero(se.method.type) == se.method.tables.m
    fresh.close()
