"""Full-size (10k-bus) batch runs checked through size-independent properties (BASELINE configs 4 and 5 in small S):
the converged state of every sampled scenario satisfies the oracle's power-flow equations of the outage-modified grid,
Monte-Carlo objectives follow the chi-square expectation, and batch results equal single-case solves."""
import numpy as np
import pytest

import jgb200
import oracle
from oracle import nr as onr
from oracle.fast import FastNR
from oracle.model import apply_outage
from conftest import oracle_system, product_system

pytestmark = pytest.mark.gpu


def test_n1_sweep_10k_satisfies_oracle_equations(ctx):
    ps = product_system("synthetic10k")
    a = jgb200.newton_raphson(ps, ctx)
    elig = jgb200.eligible_outages(ps)
    assert len(elig) == 12567
    S = 512
    ks = elig[np.linspace(0, len(elig) - 1, S).astype(int)]
    res = jgb200.nr_batch(a, ks)
    assert (res.status == 0).all()
    assert res.iterations.min() >= 5 and res.iterations.max() <= 8
    assert res.total_iterations == int(res.iterations.sum())
    assert np.isfinite(res.vm).all() and res.vm.min() > 0.9 and res.vm.max() < 1.2
    # independent check: the oracle's C mismatch function on the outage-modified Ybus, evaluated at our state
    os_ = oracle_system("synthetic10k")
    base = oracle.ac_model(os_)
    o = onr.newton_raphson(os_, base)
    f = FastNR(o)
    for pos in (0, 17, 255, 511):
        m = apply_outage(os_, base, int(ks[pos]))
        f.set_y(m.nzval, m.nzval_t)
        f.vm[:] = res.vm[pos]
        f.va[:] = res.va[pos]
        dp, dq = f.mismatch()
        assert dp < 1e-8 and dq < 1e-8
    # slack and PV magnitudes are never touched by the iteration
    fixed = o.bus_type != 1
    assert np.array_equal(res.vm[:, fixed], np.broadcast_to(o.vm[fixed], (S, fixed.sum())))
    assert np.all(res.va[:, o.slack] == 0.0)
    # different outages give different states; the base-case scenario reproduces powerFlow!
    assert np.abs(res.vm[0] - res.vm[1]).max() > 1e-9
    base_res = jgb200.nr_batch(a, np.array([-1]))
    assert jgb200.power_flow(a)
    np.testing.assert_allclose(base_res.vm[0], a.voltage.magnitude, atol=1e-12)


def test_monte_carlo_wls_10k_chi_square(ctx):
    ps = product_system("synthetic10k")
    a = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(a)
    vm, va = a.voltage.magnitude, a.voltage.angle
    pw = jgb200.power(ps, vm, va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    buses = np.sort(np.random.default_rng(7).choice(ps.n, ps.n // 10, replace=False))
    jgb200.add_pmu(mon, pw, vm, va, buses=buses, polar=False)
    se = jgb200.gauss_newton(mon, ctx)
    t = se.method.tables
    assert t.m == 82796
    wd = t.w_nzval[t.w_colptr[:-1] - 1]          # diagonal precision (no correlated PMUs in this set)
    S = 48
    Z = np.stack([t.mean + np.sqrt(1 / wd) * np.random.default_rng(1000 + s).standard_normal(t.m) for s in range(S)])
    res = jgb200.wls_batch(se, Z)
    assert (res.status == 0).all()
    dof = t.m - (2 * ps.n - 1)
    # objective ~ chi2(dof): mean dof, std sqrt(2 dof); 6 sigma band per draw, 1 % band on the batch mean
    assert np.all(np.abs(res.objective - dof) < 6 * np.sqrt(2 * dof))
    assert abs(res.objective.mean() - dof) < 0.01 * dof
    # estimates stay within a few measurement sigmas of the truth, and a single-case run reproduces a batch row
    assert np.abs(res.vm - vm).max() < 5e-3 and np.abs(res.va - va).max() < 5e-3
    jgb200.set_mean(se, Z[11])
    jgb200.set_voltage_se(se, ps.vm, ps.va)
    assert jgb200.state_estimation(se)
    np.testing.assert_allclose(res.vm[11], se.voltage.magnitude, atol=1e-11)
    np.testing.assert_allclose(res.va[11], se.voltage.angle, atol=1e-11)
    assert res.iterations[11] == se.method.iteration


def test_70k_bus_case_single_and_outage_batch(ctx):
    """The largest single-interconnect case the reference ships (case_ACTIVSg70k.h5: 70 000 buses, 88 207 branches,
    dim J = 129 k): index sets bit-exact, Newton-Raphson from the stored profile equals the oracle (6 iterations,
    1e-8), and a 64-outage batch satisfies the oracle's equations of the modified grids."""
    ps, os_ = product_system("case_ACTIVSg70k"), oracle_system("case_ACTIVSg70k")
    a = jgb200.newton_raphson(ps, ctx)
    base = oracle.ac_model(os_)
    o = onr.newton_raphson(os_, base)
    m, ex = a.method, onr.export_one_based(o)
    for mine, key in ((m.pq, "pq"), (m.pvpq, "pvpq"), (m.pcount, "pcount"), (m.jacobian_colptr, "j_colptr"),
                      (m.jacobian_rowval, "j_rowval")):
        assert np.array_equal(mine, ex[key]), key
    f = FastNR(o)
    assert f.power_flow() and jgb200.power_flow(a)
    assert a.method.iteration == f.iteration == 6
    assert np.abs(a.voltage.magnitude - f.vm).max() < 1e-8 and np.abs(a.voltage.angle - f.va).max() < 1e-8
    elig = jgb200.eligible_outages(ps)
    ks = elig[np.linspace(0, len(elig) - 1, 64).astype(int)]
    res = jgb200.nr_batch(a, ks)
    ok = res.status == 0
    assert ok.sum() >= 60                      # a few heavy corridors of this case have no N-1 solution from this start
    for pos in np.flatnonzero(ok)[[0, 20, -1]]:
        mo = apply_outage(os_, base, int(ks[pos]))
        f.set_y(mo.nzval, mo.nzval_t)
        f.vm[:] = res.vm[pos]
        f.va[:] = res.va[pos]
        dp, dq = f.mismatch()
        assert dp < 1e-8 and dq < 1e-8
    # every outage the device reports as not converged also fails under the oracle (fresh partial-pivoting SuperLU per
    # iteration), and two of the converged ones take the same number of iterations there
    for pos in list(np.flatnonzero(~ok)) + list(np.flatnonzero(ok)[[5, 40]]):
        mo = apply_outage(os_, base, int(ks[pos]))
        f.set_y(mo.nzval, mo.nzval_t)
        f.reset()
        conv = f.power_flow(20, 1e-8)
        assert conv == bool(ok[pos]), f"outage {ks[pos]}: device status {res.status[pos]}, oracle converged {conv}"
        if conv:
            assert f.iteration == res.iterations[pos]
            assert np.abs(f.vm - res.vm[pos]).max() < 1e-8


def test_70k_bus_state_estimation_recovers_power_flow(ctx):
    """Config-3 measurement set on the 70 000-bus case (m = 0.57 M rows, gain fronts beyond the shared-memory LDLt
    limit): exact measurements of the power-flow state are recovered (test/stateEstimation/analysis.jl recovery
    property), and the chi-square / residual tests see no bad data."""
    ps = product_system("case_ACTIVSg70k")
    a = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(a)
    vm, va = a.voltage.magnitude, a.voltage.angle
    pw = jgb200.power(ps, vm, va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    buses = np.sort(np.random.default_rng(7).choice(ps.n, ps.n // 10, replace=False))
    jgb200.add_pmu(mon, pw, vm, va, buses=buses, polar=False)
    se = jgb200.gauss_newton(mon, ctx)
    assert jgb200.state_estimation(se)
    assert se.method.iteration <= 6
    assert np.abs(se.voltage.magnitude - vm).max() < 1e-8 and np.abs(se.voltage.angle - va).max() < 1e-8
    assert not jgb200.chi_test(se).detect
    assert not jgb200.residual_test(se).detect
