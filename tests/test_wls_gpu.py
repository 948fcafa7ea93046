"""Parity of the CUDA Gauss-Newton WLS path (through the C ABI) against the CPU oracle and the known answers."""
import ctypes as C

import numpy as np
import pytest

import jgb200
import oracle
from oracle import nr as onr, wls as owls, post
from conftest import golden, oracle_system, product_system

pytestmark = pytest.mark.gpu
VOLT_ATOL = 1e-8


def _modified14(ps):
    """System of test/stateEstimation/analysis.jl:7-20 (works on both container kinds)."""
    ps.bus_type[0] = 2
    ps.bus_type[2] = 3
    ps.slack = 2
    ps.va[2] = -0.25
    ps.vm[0], ps.vm[2], ps.vm[3], ps.vm[4] = 1.0, 1.2, 1.0, 1.1
    ps.g[2], ps.g[5] = 0.01, 0.05
    return ps


def _truth(name, modify=None):
    ps, os_ = product_system(name), oracle_system(name)
    if modify:
        modify(ps)
        modify(os_)
    ps.model = jgb200.ac_model(ps)
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    pw = jgb200.power(ps, o.vm, o.va)
    return ps, os_, o, pw


def _bus_pmus(mon, pw, o):
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(mon.system.n), polar=True, variance_magnitude=1.0,
                   variance_angle=1.0)


def _everything(ps, o, pw):
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_ammeter(mon, pw, variance=1e-2)
    jgb200.add_ammeter(mon, pw, square=True)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(ps.n), branch=True, polar=True, variance_magnitude=1e-2,
                   variance_angle=1e-2)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=(), branch=True, polar=True, square=True, variance_magnitude=1e-2,
                   variance_angle=1e-2)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(ps.n), branch=True, polar=False, variance_magnitude=1e-4,
                   variance_angle=1e-4)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(0, ps.n, 2), branch=True, polar=False, correlated=True)
    mon.watt["status"][5] = 0
    mon.var["status"][17] = 0
    mon.pmu["mag_status"][3] = 0
    mon.pmu["ang_status"][40] = 0
    return mon


@pytest.mark.parametrize("case", ["case14test", "case30test"])
def test_all_codes_one_increment(case, ctx):
    """Every one of the 21 measurement codes (+ out-of-service rows, correlated pairs): tables bit-exact, residual,
    H values, gain values, objective and increment against the oracle for one increment! call."""
    ps, os_, o, pw = _truth(case, _modified14 if case == "case14test" else None)
    mon = _everything(ps, o, pw)
    a = jgb200.gauss_newton(mon, ctx)
    g = owls.gauss_newton(os_, mon, o.mdl)
    ex = owls.export_one_based(g)
    t = a.method.tables
    assert set(np.unique(t.type)) == set(range(22))
    assert np.array_equal(t.h_colptr, ex["h_colptr"]) and np.array_equal(t.h_rowval, ex["h_rowval"])
    assert np.array_equal(t.type, ex["type"]) and np.array_equal(t.index, ex["index"])
    assert np.array_equal(t.range, ex["range"])
    gcp, grv = owls.gain_pattern(g)
    assert np.array_equal(a.method.gain_colptr, gcp + 1) and np.array_equal(a.method.gain_rowval, grv + 1)
    mi = jgb200.increment(a)
    omi = owls.increment(g)
    np.testing.assert_allclose(a.residual, g.residual, rtol=1e-11, atol=1e-10)   # h(x) sums cancelling terms ~1e2
    hs = max(1.0, np.abs(g.h_nzval).max())
    # squared-current / current-angle derivatives cancel heavily on lightly loaded branches, which amplifies the
    # 1-ulp difference between pow(t,4) (Julia / oracle) and (t*t)*(t*t) (kernel) in the coefficients
    np.testing.assert_allclose(a.jacobian_nzval, g.h_nzval, rtol=1e-9, atol=1e-9 * hs)
    import scipy.sparse as sp
    G_ours = sp.csc_matrix((a.gain_nzval, grv, gcp), shape=(2 * ps.n, 2 * ps.n)).toarray()
    G_ref = g.gain.toarray()          # SciPy drops the explicit zeros Julia would keep; compare dense
    np.testing.assert_allclose(G_ours, G_ref, rtol=1e-10, atol=1e-12 * np.abs(G_ref).max())
    assert G_ours[ps.slack, ps.slack] == 1.0 and np.count_nonzero(G_ours[ps.slack]) == 1
    assert a.method.objective == pytest.approx(g.objective, rel=1e-10)
    np.testing.assert_allclose(a.increment, g.increment, rtol=1e-6, atol=1e-9 * max(1.0, omi))
    assert mi == pytest.approx(omi, rel=1e-6)
    assert a.increment[ps.slack] == 0.0


RECOVERY = {
    "voltmeter": lambda mon, pw, o: jgb200.add_voltmeter(mon, o.vm),
    "ammeter": lambda mon, pw, o: jgb200.add_ammeter(mon, pw, variance=1e-2),
    "ammeter_square": lambda mon, pw, o: jgb200.add_ammeter(mon, pw, square=True),
    "watt_bus": lambda mon, pw, o: jgb200.add_wattmeter(mon, pw, branch=False),
    "watt_branch": lambda mon, pw, o: jgb200.add_wattmeter(mon, pw, bus=False),
    "var_bus": lambda mon, pw, o: jgb200.add_varmeter(mon, pw, branch=False),
    "var_branch": lambda mon, pw, o: jgb200.add_varmeter(mon, pw, bus=False, variance=1e-2),
    "pmu_rect_branch": lambda mon, pw, o: jgb200.add_pmu(mon, pw, o.vm, o.va, branch=True, polar=False,
                                                         variance_magnitude=1e-4, variance_angle=1e-4),
    "pmu_rect_correlated": lambda mon, pw, o: jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(14), branch=True,
                                                             polar=False, correlated=True),
}


@pytest.mark.parametrize("name", sorted(RECOVERY))
def test_recovers_power_flow(name, ctx):
    """test/stateEstimation/analysis.jl:2-346 (testAcEstimation): exact measurements -> PF voltages within 1e-10,
    and the same number of Gauss-Newton updates as the oracle."""
    ps, os_, o, pw = _truth("case14test", _modified14)
    mon = jgb200.measurement(ps)
    RECOVERY[name](mon, pw, o)
    _bus_pmus(mon, pw, o)
    a = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(a, iteration=200, tolerance=1e-12)
    np.testing.assert_allclose(a.voltage.magnitude, o.vm, atol=1e-10, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, o.va, atol=1e-10, rtol=0)
    g = owls.gauss_newton(os_, mon, o.mdl)
    assert owls.state_estimation(g, iteration=200, tolerance=1e-12)
    # 1e-12 sits at the rounding floor of the increments, so the last update may or may not be taken there (squared
    # currents converge sub-linearly near the end: their count at 1e-12 is rounding-sensitive and not compared) ...
    if name != "ammeter_square":
        assert abs(a.method.iteration - g.iteration) <= 1
    # ... at the reference's own default tolerance the counts are equal, for every meter type
    b = jgb200.gauss_newton(mon, ctx)
    g2 = owls.gauss_newton(os_, mon, o.mdl)
    assert jgb200.state_estimation(b, iteration=200, tolerance=1e-8) and owls.state_estimation(g2, iteration=200, tolerance=1e-8)
    assert b.method.iteration == g2.iteration


def test_polar_branch_pmus_with_statuses(ctx):
    """analysis.jl:110-135: from-end polar PMUs (codes 2, 14) with some magnitude / angle statuses off."""
    ps, os_, o, pw = _truth("case14test", _modified14)
    mon = jgb200.measurement(ps)
    jgb200.add_pmu(mon, pw, o.vm, o.va, branch=True, polar=True, variance_magnitude=1e-2, variance_angle=1e-2)
    keep = mon.pmu["frm"].copy()
    for k in mon.pmu:
        mon.pmu[k] = mon.pmu[k][keep]
    mon.pmu["mag_status"][[1, 13, 17]] = 0
    mon.pmu["ang_status"][[13, 17]] = 0
    _bus_pmus(mon, pw, o)
    a = jgb200.gauss_newton(mon, ctx)
    assert (a.method.type == 0).sum() == 5
    assert jgb200.state_estimation(a, iteration=200, tolerance=1e-12)
    np.testing.assert_allclose(a.voltage.magnitude, o.vm, atol=1e-10, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, o.va, atol=1e-10, rtol=0)


def test_bad_data_known_answer(ctx):
    """test/stateEstimation/badData.jl:5-41: objective 3227.3 +- 0.1 after convergence with one gross error."""
    ka = golden("known_answers")["badData_one_outlier"]

    def mod(s):
        s.bus_type[0] = 2
        s.bus_type[2] = 3
        s.slack = 2
        s.va[2] = -0.17

    ps, os_, o, pw = _truth("case14test", mod)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm, variance=1e-2)
    jgb200.add_wattmeter(mon, pw, variance=1e-2)
    jgb200.add_varmeter(mon, pw, variance=1e-2)
    mon.var["mean"][3] = 10.25
    a = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(a)
    assert a.method.tables.m == 114
    assert abs(a.method.objective - ka["objective"]) < ka["atol"]
    g = owls.gauss_newton(os_, mon, o.mdl)
    assert owls.state_estimation(g)
    assert a.method.iteration == g.iteration == 30
    np.testing.assert_allclose(a.voltage.magnitude, g.vm, atol=VOLT_ATOL, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, g.va, atol=VOLT_ATOL, rtol=0)


def _config3(ps, o, pw, seed=1):
    """BASELINE config 3: V at every bus, P/Q injections + flows both ends, rectangular PMUs on a seeded 10 %."""
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    rng = np.random.default_rng(7)
    buses = np.sort(rng.choice(ps.n, ps.n // 10, replace=False))
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=buses, polar=False)
    return mon


@pytest.mark.parametrize("case", ["synthetic20", "synthetic10k"])
def test_config3_with_noise_matches_oracle(case, ctx):
    ps, os_, o, pw = _truth(case)
    mon = _config3(ps, o, pw)
    a = jgb200.gauss_newton(mon, ctx)
    # symmetric no-pivot SuperLU settings for the gain matrix (default pivoting explodes on G)
    g = owls.gauss_newton(os_, mon, o.mdl, lu_options=dict(permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                                                           options=dict(SymmetricMode=True)))
    t = a.method.tables
    ex = owls.export_one_based(g)
    assert np.array_equal(t.h_colptr, ex["h_colptr"]) and np.array_equal(t.h_rowval, ex["h_rowval"])
    gcp, grv = owls.gain_pattern(g)
    assert np.array_equal(a.method.gain_colptr, gcp + 1) and np.array_equal(a.method.gain_rowval, grv + 1)
    rng = np.random.default_rng(1)
    sigma = np.sqrt(1.0 / np.asarray(g.w.diagonal()))
    z = g.mean + sigma * rng.standard_normal(g.m)
    jgb200.set_mean(a, z)
    g.mean[:] = z
    ok_a = jgb200.state_estimation(a)
    ok_o = owls.state_estimation(g)
    assert ok_a and ok_o
    assert a.method.iteration == g.iteration
    assert a.method.objective == pytest.approx(g.objective, rel=1e-8)
    np.testing.assert_allclose(a.voltage.magnitude, g.vm, atol=VOLT_ATOL, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, g.va, atol=VOLT_ATOL, rtol=0)


def test_stepwise_equals_run(ctx):
    ps, os_, o, pw = _truth("case30test")
    mon = _config3(ps, o, pw)
    a = jgb200.gauss_newton(mon, ctx)
    n_it = 0
    for _ in range(41):
        if jgb200.increment(a) < 1e-8:
            break
        jgb200.solve_se(a)
        n_it += 1
    vm = a.voltage.magnitude.copy()
    b = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(b)
    assert b.method.iteration == n_it
    np.testing.assert_allclose(vm, b.voltage.magnitude, atol=1e-13)


def test_monte_carlo_batch(ctx):
    """Config 5 in small: S noise draws through jgb_wls_batch == S separate stateEstimation! runs."""
    ps, os_, o, pw = _truth("synthetic20")
    mon = _config3(ps, o, pw)
    a = jgb200.gauss_newton(mon, ctx)
    t = a.method.tables
    S = 40
    W = np.zeros(t.m)
    for c in range(t.m):
        for q in range(t.w_colptr[c] - 1, t.w_colptr[c + 1] - 1):
            if t.w_rowval[q] - 1 == c:
                W[c] = t.w_nzval[q]
    Z = np.stack([t.mean + np.sqrt(1 / W) * np.random.default_rng(1000 + s).standard_normal(t.m) for s in range(S)])
    res = jgb200.wls_batch(a, Z)
    assert (res.status == 0).all()
    for s in (0, 7, 39):
        jgb200.set_mean(a, Z[s])
        jgb200.set_voltage_se(a, ps.vm, ps.va)
        assert jgb200.state_estimation(a)
        assert a.method.iteration == res.iterations[s]
        np.testing.assert_allclose(res.vm[s], a.voltage.magnitude, atol=1e-12)
        np.testing.assert_allclose(res.va[s], a.voltage.angle, atol=1e-12)
        assert res.objective[s] == pytest.approx(a.method.objective, rel=1e-10)
    assert res.total_iterations == res.iterations.sum()
