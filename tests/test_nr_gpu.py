"""Parity of the CUDA Newton-Raphson path (through the C ABI) against the CPU oracle and the golden vectors."""
import ctypes as C

import numpy as np
import pytest

import jgb200
import oracle
from oracle import nr as onr
from conftest import golden, oracle_system, product_system

pytestmark = pytest.mark.gpu

CASES = ["case14test", "case30test", "synthetic20", "synthetic10k", "case_ACTIVSg10k"]
VALUE_RTOL = 1e-11   # FP64 assembly: same formulas, CUDA vs glibc sincos differ by <= 2 ulp
VOLT_ATOL = 1e-8     # north_star: within 1e-8 p.u. / rad on voltages and angles


def _pair(name, ctx):
    ps = product_system(name)
    os_ = oracle_system(name)
    a = jgb200.newton_raphson(ps, ctx)
    o = onr.newton_raphson(os_)
    return a, o


@pytest.mark.parametrize("case", CASES)
def test_index_sets_bit_exact(case, ctx):
    a, o = _pair(case, ctx)
    ex = onr.export_one_based(o)
    m = a.method
    assert np.array_equal(m.pq, ex["pq"])
    assert np.array_equal(m.pvpq, ex["pvpq"])
    assert np.array_equal(m.pcount, ex["pcount"])
    assert np.array_equal(m.jacobian_colptr, ex["j_colptr"])
    assert np.array_equal(m.jacobian_rowval, ex["j_rowval"])
    assert m.pq.dtype == np.int64 and m.jacobian_rowval.dtype == np.int64


@pytest.mark.parametrize("case", CASES)
def test_mismatch_and_jacobian_values(case, ctx):
    a, o = _pair(case, ctx)
    sp, sq = jgb200.mismatch(a)
    op, oq = onr.mismatch(o)
    onr.fill_jacobian(o)
    assert sp == pytest.approx(op, rel=1e-12, abs=1e-14)
    assert sq == pytest.approx(oq, rel=1e-12, abs=1e-14)
    scale = max(1.0, np.abs(o.j_nzval).max())
    # the mismatch is a cancelling sum of terms as large as the Jacobian entries: absolute tolerance scales with them
    np.testing.assert_allclose(a.mismatch, o.mismatch, rtol=VALUE_RTOL, atol=1e-13 * scale)
    np.testing.assert_allclose(a.jacobian_nzval, o.j_nzval, rtol=VALUE_RTOL, atol=1e-13 * scale)


@pytest.mark.parametrize("case", CASES)
def test_single_solve_step(case, ctx):
    """One mismatch! + solve!: the increment solves J dx = f to FP64 accuracy and the state update matches."""
    a, o = _pair(case, ctx)
    jgb200.mismatch(a)
    jgb200.solve(a)
    onr.mismatch(o)
    onr.solve(o)
    inc = a.increment
    np.testing.assert_allclose(inc, o.increment, rtol=1e-7, atol=1e-9 * max(1, np.abs(o.increment).max()))
    J = onr.jacobian_csc(o)
    res = np.abs(J @ inc - o.mismatch).max()
    assert res <= 1e-10 * max(1.0, np.abs(o.mismatch).max())
    np.testing.assert_allclose(a.voltage.magnitude, o.vm, atol=VOLT_ATOL, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, o.va, atol=VOLT_ATOL, rtol=0)
    assert a.method.iteration == 1


@pytest.mark.parametrize("case,iters", [("case14test", 7), ("case30test", 4)])
def test_power_flow_golden(case, iters, ctx):
    """The reference's own known answers (test/powerFlow/analysis.jl:5-67 against results.h5)."""
    g = golden(case)["newtonRaphson"]
    a = jgb200.newton_raphson(product_system(case), ctx)
    assert jgb200.power_flow(a)
    assert a.method.iteration == iters
    np.testing.assert_allclose(a.voltage.magnitude, g["voltageMagnitude"], rtol=1.5e-8, atol=0)
    np.testing.assert_allclose(a.voltage.angle, g["voltageAngle"], rtol=1.5e-8, atol=1e-15)


@pytest.mark.parametrize("case", CASES)
def test_power_flow_matches_oracle(case, ctx):
    a, o = _pair(case, ctx)
    ok_a = jgb200.power_flow(a)
    tr = []
    ok_o = onr.power_flow(o, trace=tr)
    assert ok_a == ok_o and ok_a
    assert a.method.iteration == o.iteration
    np.testing.assert_allclose(a.voltage.magnitude, o.vm, atol=VOLT_ATOL, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, o.va, atol=VOLT_ATOL, rtol=0)
    assert a.last_stop[0] < 1e-8 and a.last_stop[1] < 1e-8


def test_stepwise_loop_equals_run(ctx):
    """User loop mismatch!/solve! (docs/src/manual/acPowerFlow.md:166-210) == powerFlow! wrapper."""
    a = jgb200.newton_raphson(product_system("case14test"), ctx)
    n_it = 0
    for _ in range(21):
        sp, sq = jgb200.mismatch(a)
        if sp < 1e-8 and sq < 1e-8:
            break
        jgb200.solve(a)
        n_it += 1
    vm, va = a.voltage.magnitude.copy(), a.voltage.angle.copy()
    b = jgb200.newton_raphson(product_system("case14test"), ctx)
    assert jgb200.power_flow(b)
    assert n_it == b.method.iteration == 7
    np.testing.assert_allclose(vm, b.voltage.magnitude, atol=1e-13)
    np.testing.assert_allclose(va, b.voltage.angle, atol=1e-13)


def test_iteration_cap_is_soft(ctx):
    a = jgb200.newton_raphson(product_system("case14test"), ctx)
    assert jgb200.power_flow(a, iteration=3) is False
    assert a.method.iteration == 3


@pytest.mark.parametrize("case,k", [("case14test", 6), ("case30test", 10), ("synthetic20", 100)])
def test_outage_reuse(case, k, ctx):
    """updateBranch!(analysis; status = 0) then re-solve == freshly built model (test/powerFlow/reusing.jl:40-84),
    and restoring the branch returns to the base solution."""
    ps = product_system(case)
    a = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(a)
    base_vm = a.voltage.magnitude.copy()
    jgb200.update_branch(a, k, 0)
    jgb200.set_initial_point(a)
    assert jgb200.power_flow(a)
    os_ = oracle_system(case)
    os_.status[k] = 0
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    assert a.method.iteration == o.iteration
    np.testing.assert_allclose(a.voltage.magnitude, o.vm, atol=VOLT_ATOL, rtol=0)
    np.testing.assert_allclose(a.voltage.angle, o.va, atol=VOLT_ATOL, rtol=0)
    jgb200.update_branch(a, k, 1)
    jgb200.set_initial_point(a)
    assert jgb200.power_flow(a)
    np.testing.assert_allclose(a.voltage.magnitude, base_vm, atol=VOLT_ATOL, rtol=0)


def test_singular_jacobian_raises(ctx):
    """An islanding outage makes J singular: the reference surfaces a SingularException; here rc = -3."""
    ps = product_system("case14test")
    a = jgb200.newton_raphson(ps, ctx)
    jgb200.update_branch(a, 13, 0)      # branch 7-15 is the only in-service link of bus 15 (label 8)
    jgb200.mismatch(a)
    with pytest.raises(jgb200.JgbError) as e:
        for _ in range(3):
            jgb200.solve(a)
    assert e.value.rc == -3


def test_bad_arguments(ctx):
    lib = ctx.lib
    assert lib.jgb_nr_setup(ctx.handle, 0, None, None, None, None, None, 1) == -1
    assert b"null" in lib.jgb_last_error(ctx.handle) or b"empty" in lib.jgb_last_error(ctx.handle)
    c2 = jgb200.Context(0)
    assert c2.lib.jgb_nr_mismatch(c2.handle, None, None) == -1     # setup not called
    c2.close()


@pytest.mark.parametrize("case", ["case14test", "case30test", "synthetic10k"])
def test_power_postprocessing_on_device(case, ctx):
    """power!/current! on the device vs the oracle and, for the IEEE cases, vs the MATPOWER goldens
    asserted by testPower (test/utility/utility.jl:43-59)."""
    from oracle import post
    a, o = _pair(case, ctx)
    assert jgb200.power_flow(a) and onr.power_flow(o)
    pw = jgb200.power_device(a)
    ref = post.powers(o.sys, o.mdl, o.vm, o.va)
    for k in ref:
        np.testing.assert_allclose(pw[k], ref[k], rtol=1e-9, atol=1e-10, err_msg=k)
    host = jgb200.power(a.system, a.voltage.magnitude, a.voltage.angle)
    for k in host:
        np.testing.assert_allclose(pw[k], host[k], rtol=1e-9, atol=1e-10, err_msg=k)
    if case != "synthetic10k":
        g = golden(case)["newtonRaphson"]
        for ours, theirs in (("injection_active", "injectionActive"), ("injection_reactive", "injectionReactive"),
                             ("from_active", "fromActive"), ("from_reactive", "fromReactive"),
                             ("to_active", "toActive"), ("to_reactive", "toReactive")):
            np.testing.assert_allclose(pw[ours], g[theirs], rtol=1.5e-8, atol=1e-10, err_msg=ours)


@pytest.mark.parametrize("case,total", [("case14test", 14), ("case30test", 8)])
def test_reactive_limit_goldens(case, total, ctx):
    """test/powerFlow/limits.jl:4-43 on the device path: generator outputs (testPower goldens), reactiveLimit!, second
    Newton-Raphson run on the changed bus types, adjustAngle! -> results.h5:/case/reactiveLimit/newtonRaphson."""
    g = golden(case)
    ps = product_system(case)
    a = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(a)
    pg, qg = jgb200.generator_power(a)
    np.testing.assert_allclose(pg, g["newtonRaphson"]["generatorActive"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(qg, g["newtonRaphson"]["generatorReactive"], rtol=0, atol=1e-9)
    first, slack0 = a.method.iteration, a.slack
    violate = jgb200.reactive_limit(a)
    os_ = oracle_system(case)
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    assert np.array_equal(violate, onr.reactive_limit(o))
    assert np.array_equal(ps.bus_type, os_.bus_type) and ps.slack == os_.slack
    b = jgb200.newton_raphson(ps, ctx)
    assert jgb200.power_flow(b)
    jgb200.adjust_angle(b, slack0)
    assert b.method.iteration + first == total
    np.testing.assert_allclose(b.voltage.magnitude, g["reactiveLimit"]["voltageMagnitude"], rtol=0, atol=VOLT_ATOL)
    np.testing.assert_allclose(b.voltage.angle, g["reactiveLimit"]["voltageAngle"], rtol=0, atol=VOLT_ATOL)
