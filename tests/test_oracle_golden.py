"""Pins the CPU oracle against the reference's own golden vectors and known answers (no GPU)."""
import numpy as np
import pytest
import scipy.stats

import oracle
from oracle import nr, wls, post
from conftest import golden, oracle_system

RTOL = 1.5e-8   # Julia's default `isapprox` rtol = sqrt(eps), used by test/utility/utility.jl:34-40


@pytest.mark.parametrize("case,iters", [("case14test", 7), ("case30test", 4)])
def test_newton_raphson_golden(case, iters):
    """test/powerFlow/analysis.jl:5-67 — iteration count and voltages vs results.h5 (MATPOWER)."""
    g = golden(case)["newtonRaphson"]
    s = oracle_system(case)
    a = nr.newton_raphson(s)
    assert nr.power_flow(a)
    assert a.iteration == iters == int(g["iteration"][0])
    np.testing.assert_allclose(a.vm, g["voltageMagnitude"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(a.va, g["voltageAngle"], rtol=RTOL, atol=1e-15)
    pw = post.powers(s, a.mdl, a.vm, a.va)
    for ours, theirs in (("injection_active", "injectionActive"), ("injection_reactive", "injectionReactive"),
                         ("from_active", "fromActive"), ("from_reactive", "fromReactive"),
                         ("to_active", "toActive"), ("to_reactive", "toReactive")):
        np.testing.assert_allclose(pw[ours], g[theirs], rtol=RTOL, atol=1e-12)


def test_jacobian_pattern_sizes():
    """Stored zeros are structural: out-of-service branches keep their entries (model.jl:70-71)."""
    a = nr.newton_raphson(oracle_system("case14test"))
    assert len(a.mdl.rowval) == 54 and a.dim == 22 and len(a.j_rowval) == 146
    a = nr.newton_raphson(oracle_system("case30test"))
    assert len(a.mdl.rowval) == 112 and a.dim == 53


def _test14():
    """System of test/stateEstimation/analysis.jl:7-20."""
    s = oracle_system("case14test")
    s.bus_type[0] = 2
    s.bus_type[2] = 3
    s.slack = 2
    s.va[2] = -0.25
    s.vm[0], s.vm[2], s.vm[3], s.vm[4] = 1.0, 1.2, 1.0, 1.1
    s.g[2], s.g[5] = 0.01, 0.05
    a = nr.newton_raphson(s)
    assert nr.power_flow(a)
    return s, a, post.powers(s, a.mdl, a.vm, a.va)


OFF = dict(volt=False, watt=False, var=False)
WLS_CASES = {
    "voltmeter": dict(OFF, volt=True),
    "ammeter": dict(OFF, amp=True, var_amp=1e-2),
    "ammeter_square": dict(OFF, amp=True, amp_square=True),
    "watt_bus": dict(OFF, watt=True, power_branch=False),
    "watt_branch": dict(OFF, watt=True, power_bus=False),
    "var_bus": dict(OFF, var=True, power_branch=False),
    "var_branch": dict(OFF, var=True, power_bus=False, var_power=1e-2),
    "pmu_rect_branch": dict(OFF, pmu_branch=True, pmu_polar=False, var_pmu_mag=1e-4, var_pmu_ang=1e-4),
    "pmu_rect_branch_correlated": dict(OFF, pmu_branch=True, pmu_polar=False, pmu_correlated=True),
    "all_legacy": dict(amp=True),
}


@pytest.mark.parametrize("name", sorted(WLS_CASES))
def test_wls_recovers_power_flow(name):
    """test/stateEstimation/analysis.jl:2-346 (testAcEstimation): exact measurements -> PF voltages within 1e-10."""
    s, a, pw = _test14()
    me = wls.measurements_from_solution(s, pw, a.vm, a.va, **WLS_CASES[name])
    mb = wls.measurements_from_solution(s, pw, a.vm, a.va, pmu_bus=range(s.n), pmu_polar=True, var_pmu_mag=1.0,
                                        var_pmu_ang=1.0, **OFF)
    for k in me.pmu:
        me.pmu[k] = np.concatenate([me.pmu[k], mb.pmu[k]])
    se = wls.gauss_newton(s, me, a.mdl)
    assert wls.state_estimation(se, iteration=200, tolerance=1e-12)
    np.testing.assert_allclose(se.vm, a.vm, atol=1e-10, rtol=0)
    np.testing.assert_allclose(se.va, a.va, atol=1e-10, rtol=0)


def test_wls_polar_branch_pmu_with_outages():
    """analysis.jl:110-135: from-end polar PMUs with some statuses off."""
    s, a, pw = _test14()
    me = wls.measurements_from_solution(s, pw, a.vm, a.va, pmu_branch=True, pmu_polar=True, var_pmu_mag=1e-2,
                                        var_pmu_ang=1e-2, **OFF)
    keep = me.pmu["frm"]
    for k in me.pmu:
        me.pmu[k] = me.pmu[k][keep]
    me.pmu["mag_status"][[1, 13, 17]] = 0
    me.pmu["ang_status"][[13, 17]] = 0
    mb = wls.measurements_from_solution(s, pw, a.vm, a.va, pmu_bus=range(s.n), pmu_polar=True, var_pmu_mag=1.0,
                                        var_pmu_ang=1.0, **OFF)
    for k in me.pmu:
        me.pmu[k] = np.concatenate([me.pmu[k], mb.pmu[k]])
    se = wls.gauss_newton(s, me, a.mdl)
    assert (se.type == 0).sum() == 5
    assert wls.state_estimation(se, iteration=200, tolerance=1e-12)
    np.testing.assert_allclose(se.vm, a.vm, atol=1e-10, rtol=0)
    np.testing.assert_allclose(se.va, a.va, atol=1e-10, rtol=0)


def test_bad_data_known_answer():
    """test/stateEstimation/badData.jl:5-41: objective 3227.3 +- 0.1, chi2 threshold 109.7 +- 0.1."""
    ka = golden("known_answers")["badData_one_outlier"]
    s = oracle_system("case14test")
    s.bus_type[0] = 2
    s.bus_type[2] = 3
    s.slack = 2
    s.va[2] = -0.17
    a = nr.newton_raphson(s)
    assert nr.power_flow(a)
    pw = post.powers(s, a.mdl, a.vm, a.va)
    me = wls.measurements_from_solution(s, pw, a.vm, a.va, var_volt=1e-2, var_power=1e-2)
    me.var["mean"][3] = 10.25
    se = wls.gauss_newton(s, me, a.mdl)
    assert wls.state_estimation(se)
    assert se.m == 114
    assert abs(se.objective - ka["objective"]) < ka["atol"]
    thr = scipy.stats.chi2.ppf(0.95, se.m - (2 * s.n - 1))
    assert abs(thr - ka["threshold"]) < ka["atol"]


def test_outage_reuse_matches_rebuild():
    """test/powerFlow/reusing.jl:40-84: in-place outage on the fixed pattern == freshly built model."""
    s = oracle_system("case14test")
    base = oracle.ac_model(s)
    k = 6
    mdl_inplace = oracle.model.apply_outage(s, base, k)
    s2 = s.copy()
    s2.status[k] = 0
    a1 = nr.newton_raphson(s, mdl_inplace)
    a2 = nr.newton_raphson(s2)
    assert nr.power_flow(a1) and nr.power_flow(a2)
    assert a1.iteration == a2.iteration
    np.testing.assert_allclose(a1.vm, a2.vm, atol=1e-8)
    np.testing.assert_allclose(a1.va, a2.va, atol=1e-8)
    assert np.array_equal(a1.j_rowval, a2.j_rowval)


def _bad_data_case(two_outliers):
    """Setups of test/stateEstimation/badData.jl:5-41 (one outlier) and :58-84 (two outliers)."""
    s = oracle_system("case14test")
    s.bus_type[0] = 2
    s.bus_type[2] = 3
    s.slack = 2
    s.va[2] = -0.17
    a = nr.newton_raphson(s)
    assert nr.power_flow(a)
    pw = post.powers(s, a.mdl, a.vm, a.va)
    kw = dict(pmu_bus=range(s.n), pmu_polar=True, var_pmu_mag=1e-5, var_pmu_ang=1e-5) if two_outliers else {}
    me = wls.measurements_from_solution(s, pw, a.vm, a.va, var_volt=1e-2, var_power=1e-2, **kw)
    me.var["mean"][3] = 10.25                    # "Varmeter 4"
    if two_outliers:
        me.pmu["mag_mean"][9] = 30.0             # "PMU 10"
    return s, a, me


def test_bad_data_largest_normalized_residual_known_answers():
    """residualTest!: 52.5 (Varmeter 4), then recovery; two outliers: 7713.26 (PMU 10), 78.3 (Varmeter 4), recovery
    (test/stateEstimation/badData.jl:24-41, 58-84; all atol 1e-1, recovery 1e-10)."""
    s, a, me = _bad_data_case(False)
    se = wls.gauss_newton(s, me, a.mdl)
    assert wls.state_estimation(se)
    detect, rn, idx = wls.residual_test(se)
    assert detect and abs(rn - 52.5) < 0.1 and idx == se.range[3] + 3
    assert wls.state_estimation(se)
    assert np.abs(se.vm - a.vm).max() < 1e-10 and np.abs(se.va - a.va).max() < 1e-10

    s, a, me = _bad_data_case(True)
    se = wls.gauss_newton(s, me, a.mdl)
    assert wls.state_estimation(se)
    detect, rn, idx = wls.residual_test(se)
    assert detect and abs(rn - 7713.26) < 0.1 and idx == se.range[4] + 2 * 9
    assert wls.state_estimation(se)
    detect, rn, idx = wls.residual_test(se)
    assert detect and abs(rn - 78.3) < 0.1 and idx == se.range[3] + 3
    assert wls.state_estimation(se)
    assert np.abs(se.vm - a.vm).max() < 1e-10 and np.abs(se.va - a.va).max() < 1e-10


@pytest.mark.parametrize("case,total", [("case14test", 14), ("case30test", 8)])
def test_reactive_limit_goldens(case, total):
    """test/powerFlow/limits.jl:4-43: NR, reactiveLimit!, NR again, adjustAngle! -> results.h5:/case/reactiveLimit;
    the generator outputs of the first run are the goldens' generatorActive / generatorReactive."""
    g = golden(case)
    s = oracle_system(case)
    a = nr.newton_raphson(s)
    assert nr.power_flow(a)
    pw = post.powers(s, a.mdl, a.vm, a.va)
    pg, qg = post.generator_powers(s, pw["injection_active"], pw["injection_reactive"], a.slack)
    np.testing.assert_allclose(pg, g["newtonRaphson"]["generatorActive"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(qg, g["newtonRaphson"]["generatorReactive"], rtol=0, atol=1e-10)
    first, slack0 = a.iteration, a.slack
    assert np.any(nr.reactive_limit(a) != 0)
    b = nr.newton_raphson(s)
    assert nr.power_flow(b)
    nr.adjust_angle(b, slack0)
    assert b.iteration + first == total == int(g["reactiveLimit"]["iteration"][0])
    np.testing.assert_allclose(b.vm, g["reactiveLimit"]["voltageMagnitude"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(b.va, g["reactiveLimit"]["voltageAngle"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("case,bx,iters", [("case14test", True, 23), ("case14test", False, 23),
                                           ("case30test", True, 12), ("case30test", False, 9)])
def test_fast_newton_raphson_goldens(case, bx, iters):
    """test/powerFlow/analysis.jl:70-150: fastNewtonRaphsonBX / XB iteration counts and voltages from results.h5."""
    g = golden(case)["fastNewtonRaphsonBX" if bx else "fastNewtonRaphsonXB"]
    a = nr.fast_newton_raphson(oracle_system(case), bx)
    assert nr.fnr_power_flow(a, iteration=100)
    assert a.iteration == iters == int(g["iteration"][0])
    np.testing.assert_allclose(a.vm, g["voltageMagnitude"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(a.va, g["voltageAngle"], rtol=0, atol=1e-8)
