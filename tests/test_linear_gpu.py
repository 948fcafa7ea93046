"""Parity of the constant-matrix solver (`jgb_lin_*`) and the three linear analyses built on it against the CPU
oracle, the reference's dcPowerFlow goldens and its recovery tests. Tolerance: 1e-8 p.u. / rad (north star), the
small cases are asserted much tighter."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import jgb200
from oracle import linear, nr as onr, post, wls as owls
from conftest import golden, oracle_system, product_system
from test_linear_cpu import _mod14, _dc_measurements, _pmu_truth, _product_monitoring

pytestmark = pytest.mark.gpu
ATOL = 1e-8


def test_dc_power_flow_goldens():
    """test/powerFlow/analysis.jl:230-275."""
    for case in ("case14test", "case30test"):
        g = golden(case)["dcPowerFlow"]
        a = jgb200.dc_power_flow(product_system(case))
        th = jgb200.solve_dc(a)
        np.testing.assert_allclose(th, g["voltage"], rtol=0, atol=1e-12)
        pw = jgb200.power_dc(a)
        for k in ("injection", "supply", "generator", "from"):
            np.testing.assert_allclose(pw[k], g[k], rtol=0, atol=1e-11)
        a.solver.ctx.close()


@pytest.mark.parametrize("case", ["case_ACTIVSg10k", "synthetic10k"])
def test_dc_power_flow_10k_and_scenarios(case):
    so, spd = oracle_system(case), product_system(case)
    a = jgb200.dc_power_flow(spd)
    th = jgb200.solve_dc(a)
    dc = linear.dc_model(so)
    ref = linear.dc_power_flow(so, dc)
    assert np.abs(th - ref).max() < ATOL
    # 70 demand scenarios (ragged: not a multiple of the 32-wide tile), one factorisation
    rng = np.random.default_rng(5)
    dem = so.pd[None, :] * (1.0 + 0.1 * rng.standard_normal((70, so.n)))
    got = jgb200.dc_batch(a, demand=dem)
    lu = spla.splu(linear.slack_fixed(dc, so.slack))
    sup, _ = linear._supply(so)
    for r in (0, 1, 31, 32, 69):
        want = linear.add_slack_angle(so, lu.solve(sup - dem[r] - so.gs - dc.shift_power))
        assert np.abs(got[r] - want).max() < ATOL
    assert a.solver.dims()["n"] == so.n
    a.solver.ctx.close()


def test_dc_state_estimation_recovers_power_flow():
    """test/stateEstimation/analysis.jl:455-510 through the device solver."""
    so, spd = _mod14(oracle_system("case14test")), _mod14(product_system("case14test"))
    dc = linear.dc_model(so)
    th = linear.dc_power_flow(so, dc)
    pw = linear.dc_power(so, dc, th)
    for cfg in ((True, False, True), (False, True, True), (True, True, False)):
        me = _dc_measurements(so, pw, th, *cfg)
        a = jgb200.dc_state_estimation(_product_monitoring(spd, me))
        est = jgb200.solve_dc_se(a)
        np.testing.assert_allclose(est, th, rtol=0, atol=1e-10)
        np.testing.assert_allclose(est, linear.dc_state_estimation(so, me, dc), rtol=0, atol=1e-10)
        a.method.solver.ctx.close()


def test_pmu_state_estimation_recovers_power_flow():
    """test/stateEstimation/analysis.jl:350-372, :398-412."""
    so, o, pw = _pmu_truth()
    spd = _mod14(product_system("case14test"), cond=True)
    spd.model = jgb200.ac_model(spd)
    for corr in (False, True):
        me = owls.measurements_from_solution(so, pw, o.vm, o.va, volt=False, watt=False, var=False,
                                             pmu_bus=range(so.n), pmu_branch=True, pmu_polar=False, pmu_correlated=corr)
        a = jgb200.pmu_state_estimation(_product_monitoring(spd, me))
        v = jgb200.solve_pmu_se(a)
        np.testing.assert_allclose(v.magnitude, o.vm, rtol=0, atol=1e-9)
        np.testing.assert_allclose(v.angle, o.va, rtol=0, atol=1e-9)
        a.method.solver.ctx.close()


def test_pmu_state_estimation_monte_carlo_10k():
    """PMUs on every bus and branch end of ACTIVSg10k, 40 noisy draws: every draw agrees with the oracle's solve."""
    so, spd = oracle_system("case_ACTIVSg10k"), product_system("case_ACTIVSg10k")
    spd.model = jgb200.ac_model(spd)
    o = onr.newton_raphson(so)
    assert onr.power_flow(o)
    pw = jgb200.power(spd, o.vm, o.va)
    mon = jgb200.measurement(spd)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(spd.n), branch=True, polar=False)
    # current phasors of (nearly) unloaded branches have no defined rectangular variance (errorVariance in the reference)
    keep = mon.pmu["bus"] | (mon.pmu["mag_mean"] > 1e-3)
    mon.pmu = {k: v[keep] for k, v in mon.pmu.items()}
    a = jgb200.pmu_state_estimation(mon)
    v = jgb200.solve_pmu_se(a)
    assert np.abs(v.magnitude - o.vm).max() < ATOL and np.abs(v.angle - o.va).max() < ATOL
    m = a.method
    rng = np.random.default_rng(11)
    Z = m.mean[None, :] + 1e-4 * rng.standard_normal((40, len(m.mean)))
    got = jgb200.pmu_se_batch(a, Z)
    h = m.coefficient.tocsc()
    wh = (m.precision @ h).tocsc()
    lu = spla.splu((h.T @ wh).tocsc())
    for r in (0, 17, 39):
        x = lu.solve(wh.T @ Z[r])
        vv = x[:so.n] + 1j * x[so.n:]
        assert np.abs(got.magnitude[r] - np.abs(vv)).max() < ATOL
        assert np.abs(got.angle[r] - np.angle(vv)).max() < ATOL
    m.solver.ctx.close()


def test_linear_solver_refactor_and_errors():
    rng = np.random.default_rng(3)
    n = 200
    a = sp.random(n, n, density=0.02, random_state=7, format="csc")
    a = (a + a.T + sp.diags(np.full(n, 8.0))).tocsc()
    s = jgb200.LinearSolver(a)
    b = rng.standard_normal((5, n))
    np.testing.assert_allclose(s.solve(b), spla.splu(a).solve(b.T).T, rtol=0, atol=1e-12)
    a2 = a.copy()
    a2.data *= 1.5
    s.refactor(a2)
    np.testing.assert_allclose(s.solve(b[0]), spla.splu(a2).solve(b[0]), rtol=0, atol=1e-12)
    # identity row / column
    s3 = jgb200.LinearSolver(a, skip=17, ctx=s.ctx)
    x = s3.solve(b[1])
    a3 = a.tolil(); a3[17, :] = 0; a3[:, 17] = 0; a3[17, 17] = 1
    np.testing.assert_allclose(x, spla.splu(a3.tocsc()).solve(b[1]), rtol=0, atol=1e-12)
    assert x[17] == b[1][17]
    # unsymmetric values on a symmetric pattern (fast Newton-Raphson B' with phase shifters): A and A' are factored
    au = a.copy()
    au.data = au.data * (1.0 + 0.3 * rng.random(len(au.data)))
    su = jgb200.LinearSolver(au, ctx=s.ctx)
    np.testing.assert_allclose(su.solve(b), spla.splu(au).solve(b.T).T, rtol=0, atol=1e-12)
    au2 = au.copy()
    au2.data = au2.data * (1.0 + 0.1 * rng.random(len(au2.data)))
    su.refactor(au2)
    np.testing.assert_allclose(su.solve(b[2]), spla.splu(au2).solve(b[2]), rtol=0, atol=1e-12)
    # unsymmetric pattern -> bad argument; singular -> -3
    bad = sp.csc_matrix(np.array([[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [1.0, 0.0, 2.0]]))
    with pytest.raises(jgb200.JgbError) as e:
        jgb200.LinearSolver(bad, ctx=s.ctx)
    assert e.value.rc == -1
    z = sp.csc_matrix(np.array([[1.0, 1.0], [1.0, 1.0]]))
    with pytest.raises(jgb200.JgbError) as e:
        jgb200.LinearSolver(z, ctx=s.ctx)
    assert e.value.rc == -3
    s.ctx.close()


def test_lin_edge_cases_and_bad_arguments(ctx):
    """Ragged block widths (1, 33: not multiples of the 32-wide tile), calls out of order and null pointers."""
    import ctypes as C
    lib = ctx.lib
    rng = np.random.default_rng(9)
    n = 50
    a = sp.random(n, n, density=0.08, random_state=3, format="csc")
    a = (a + a.T + sp.diags(np.full(n, 6.0))).tocsc()
    fresh = jgb200.Context(0)
    # solve before setup -> bad argument / logic error, not a crash
    x = np.empty(n)
    rc = fresh.lib.jgb_lin_solve(fresh.handle, 1, x.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -1 and b"jgb_lin_setup" in fresh.lib.jgb_last_error(fresh.handle)
    s = jgb200.LinearSolver(a, ctx=fresh)
    lu = spla.splu(a)
    for R in (1, 33):
        b = rng.standard_normal((R, n))
        np.testing.assert_allclose(s.solve(b), lu.solve(b.T).T, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        s.solve(np.zeros(n + 1))
    # projected solve without a projection; null pointers; zero right-hand sides
    z = np.zeros(3)
    assert fresh.lib.jgb_lin_solve_projected(fresh.handle, 1, z.ctypes.data_as(C.POINTER(C.c_double)),
                                            x.ctypes.data_as(C.POINTER(C.c_double))) == -1
    assert fresh.lib.jgb_lin_solve(fresh.handle, 1, None, x.ctypes.data_as(C.POINTER(C.c_double))) == -1
    assert fresh.lib.jgb_lin_solve(fresh.handle, 0, x.ctypes.data_as(C.POINTER(C.c_double)),
                                  x.ctypes.data_as(C.POINTER(C.c_double))) == -1
    assert fresh.lib.jgb_lin_refactor(fresh.handle, None) == -1
    # a projection with the wrong number of columns is refused on the host side
    with pytest.raises(ValueError):
        s.set_projection(sp.identity(n + 2, format="csc"))
    s.set_projection(sp.identity(n, format="csc"))
    b = rng.standard_normal(n)
    np.testing.assert_allclose(s.solve_projected(b), lu.solve(b), rtol=0, atol=1e-12)
    fresh.close()
