"""Linear analyses, CPU side: the oracle against the reference's goldens and recovery tests, and the
product's host tables (dc_model, H / W / mean of the DC and PMU estimators) against the oracle's restatement."""
import numpy as np
import scipy.sparse as sp

import jgb200
from oracle import linear, nr as onr, post, wls as owls
from conftest import golden, oracle_system, product_system


def _mod14(s, cond=False):
    """test/stateEstimation/analysis.jl:350-357 and :455-461."""
    s.bus_type[0] = 2
    s.bus_type[2] = 3
    s.slack = 2
    s.va[2] = -0.17
    if cond:
        s.g[2], s.g[5] = 0.01, 0.05
    return s


def test_oracle_dc_power_flow_goldens():
    """test/powerFlow/analysis.jl:230-275 against results.h5:/caseNNtest/dcPowerFlow."""
    for case in ("case14test", "case30test"):
        g = golden(case)["dcPowerFlow"]
        s = oracle_system(case)
        dc = linear.dc_model(s)
        th = linear.dc_power_flow(s, dc)
        np.testing.assert_allclose(th, g["voltage"], rtol=0, atol=1e-12)
        pw = linear.dc_power(s, dc, th)
        for k in ("injection", "supply", "generator", "from"):
            np.testing.assert_allclose(pw[k], g[k], rtol=0, atol=1e-12)


def _dc_measurements(s, pw, th, bus_watt, branch_watt, pmus):
    m = owls.Measurement()
    if bus_watt:
        for i in range(s.n):
            owls._push(m.watt, index=i, bus=True, frm=False, mean=pw["injection"][i], variance=1e-4, status=1)
    if branch_watt:
        for k in range(s.nbr):
            if s.status[k] == 1:
                owls._push(m.watt, index=k, bus=False, frm=True, mean=pw["from"][k], variance=1e-4, status=1)
                owls._push(m.watt, index=k, bus=False, frm=False, mean=pw["to"][k], variance=1e-4, status=1)
    if pmus:
        for i in range(s.n):
            owls._push(m.pmu, index=i, bus=True, frm=False, polar=True, square=False, correlated=False, mag_mean=1.0,
                       mag_variance=1e-8, mag_status=1, ang_mean=th[i], ang_variance=1e-8, ang_status=1)
    return m.finalize()


def test_oracle_dc_state_estimation_recovers_power_flow():
    """test/stateEstimation/analysis.jl:455-510: estimate == DC power-flow angles for three measurement sets."""
    s = _mod14(oracle_system("case14test"))
    dc = linear.dc_model(s)
    th = linear.dc_power_flow(s, dc)
    pw = linear.dc_power(s, dc, th)
    for cfg in ((True, False, True), (False, True, True), (True, True, False)):
        est = linear.dc_state_estimation(s, _dc_measurements(s, pw, th, *cfg), dc)
        np.testing.assert_allclose(est, th, rtol=0, atol=1e-10)


def _pmu_truth():
    s = _mod14(oracle_system("case14test"), cond=True)
    a = onr.newton_raphson(s)
    assert onr.power_flow(a)
    return s, a, post.powers(s, a.mdl, a.vm, a.va)


def test_oracle_pmu_state_estimation_recovers_power_flow():
    """test/stateEstimation/analysis.jl:350-372 (+ correlated variant :398-412)."""
    s, a, pw = _pmu_truth()
    for corr in (False, True):
        me = owls.measurements_from_solution(s, pw, a.vm, a.va, volt=False, watt=False, var=False,
                                             pmu_bus=range(s.n), pmu_branch=True, pmu_polar=False, pmu_correlated=corr)
        vm, va = linear.pmu_state_estimation(s, me, a.mdl)
        np.testing.assert_allclose(vm, a.vm, rtol=0, atol=1e-10)
        np.testing.assert_allclose(va, a.va, rtol=0, atol=1e-10)


def _same_sparse(a, b, atol):
    a, b = sp.csc_matrix(a), sp.csc_matrix(b)
    assert a.shape == b.shape
    d = (a - b)
    assert (abs(d).max() if d.nnz else 0.0) <= atol


def test_product_dc_model_matches_oracle():
    for case in ("case14test", "case30test", "case_ACTIVSg10k"):
        so, spd = oracle_system(case), product_system(case)
        o, p = linear.dc_model(so), jgb200.dc_model(spd)
        assert np.array_equal(p.nodal.indptr, o.colptr) and np.array_equal(p.nodal.indices, o.rowval)
        np.testing.assert_allclose(p.nodal.data, o.nzval, rtol=1e-14, atol=1e-9 * np.abs(o.nzval).max())
        np.testing.assert_allclose(p.admittance, o.admittance, rtol=1e-15)
        np.testing.assert_allclose(p.shift_power, o.shift_power, rtol=1e-13, atol=1e-13)


def _product_monitoring(spd, me):
    mon = jgb200.measurement(spd)
    for dev in ("volt", "amp", "watt", "var", "pmu"):
        src = getattr(me, dev)
        if len(src["index"]):
            mon._append(dev, **{k: src[k] for k in src})
    return mon


def test_product_dc_wls_tables_match_oracle():
    so, spd = _mod14(oracle_system("case14test")), _mod14(product_system("case14test"))
    dc = linear.dc_model(so)
    th = linear.dc_power_flow(so, dc)
    pw = linear.dc_power(so, dc, th)
    me = _dc_measurements(so, pw, th, True, True, True)
    me.watt["status"][3] = 0
    me.pmu["ang_status"][5] = 0
    w = linear.dc_wls(so, me, dc)
    h, prec, mean = jgb200.dc_wls_tables(_product_monitoring(spd, me), jgb200.dc_model(spd))
    _same_sparse(h, w.coefficient, 1e-12)
    _same_sparse(prec, w.precision, 1e-6)
    np.testing.assert_allclose(mean, w.mean, rtol=0, atol=1e-14)


def test_product_pmu_wls_tables_match_oracle():
    so, a, pw = _pmu_truth()
    spd = _mod14(product_system("case14test"), cond=True)
    spd.model = jgb200.ac_model(spd)
    for corr in (False, True):
        me = owls.measurements_from_solution(so, pw, a.vm, a.va, volt=False, watt=False, var=False,
                                             pmu_bus=range(so.n), pmu_branch=True, pmu_polar=False, pmu_correlated=corr)
        me.pmu["mag_status"][4] = 0
        me.pmu["ang_status"][20] = 0
        w = linear.pmu_wls(so, me, a.mdl)
        h, prec, mean = jgb200.pmu_wls_tables(_product_monitoring(spd, me))
        _same_sparse(h, w.coefficient, 1e-12)
        _same_sparse(prec, w.precision, 1e-6 * abs(w.precision).max())
        np.testing.assert_allclose(mean, w.mean, rtol=0, atol=1e-14)


def test_product_fast_jacobians_match_oracle():
    """B' and B'' of fastNewtonRaphsonBX / XB: product (vectorised) vs oracle (branch loop), incl. a phase shifter."""
    from jgb200.ac_power_flow import _initialize
    for case in ("case14test", "case30test", "case_ACTIVSg10k"):
        for bx in (True, False):
            so, ps = oracle_system(case), product_system(case)
            if case == "case14test":
                so.shift[3] = ps.shift[3] = 0.1
            o = onr.fast_newton_raphson(so, bx)
            ps.model = jgb200.ac_model(ps)
            bt, sl, _, _ = _initialize(ps)
            A, R, pq, pvpq = jgb200.fast_jacobians(ps, ps.model, bt, sl, bx)
            assert np.array_equal(pq, o.pq) and np.array_equal(pvpq, o.pvpq)
            _same_sparse(A, o.active, 1e-10)
            _same_sparse(R, o.reactive, 1e-10)
