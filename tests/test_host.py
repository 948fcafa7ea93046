"""CPU-only checks: the C-ABI library loads and exports every symbol of include/jgb200.h, host logic of the product
against the oracle, the symbolic analysis through its host replay, the C oracle against the NumPy oracle, and that
the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import jgb200
import oracle
from oracle import nr as onr, wls as owls, post
from jgb200._lib import ptr
from conftest import ROOT, oracle_system, product_system


def header_symbols():
    text = open(os.path.join(ROOT, "include", "jgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jgb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = jgb200.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/jgb200.h but not exported"
    assert sorted(jgb200.exported_symbols()) == syms, "ctypes prototypes and header disagree"
    assert lib.jgb_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device jgb_create must fail with -5 (skipped on the GPU box)."""
    lib = jgb200.load()
    rc = C.c_int32(0)
    h = lib.jgb_create(0, None, C.byref(rc))
    if h:
        lib.jgb_destroy(h)
        pytest.skip("a GPU is present")
    assert rc.value == -5
    assert b"no CPU fallback" in lib.jgb_last_error(None)
    with pytest.raises(jgb200.JgbError):
        jgb200.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "juliagrid.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


@pytest.mark.parametrize("case", ["case14test", "case30test", "synthetic20", "case_ACTIVSg10k"])
def test_product_model_matches_oracle(case):
    ps, os_ = product_system(case), oracle_system(case)
    m, om = jgb200.ac_model(ps), oracle.ac_model(os_)
    assert np.array_equal(m.colptr - 1, om.colptr) and np.array_equal(m.rowval - 1, om.rowval)
    scale = np.abs(om.nzval).max()
    np.testing.assert_allclose(m.nzval, om.nzval, rtol=0, atol=4e-16 * scale)
    np.testing.assert_allclose(m.nzval_t, om.nzval_t, rtol=0, atol=4e-16 * scale)
    from jgb200.ac_power_flow import _initialize
    bt, sl, vm, va = _initialize(ps)
    obt, osl, ovm, ova = onr.initialize(os_)
    assert np.array_equal(bt, obt) and sl == osl and np.array_equal(vm, ovm) and np.array_equal(va, ova)


def test_synthetic_grid_recipes_agree():
    a, b = jgb200.synthetic_grid(side=30), oracle.synthetic_grid(side=30)
    for k in ("pd", "qd", "frm", "to", "r", "x", "b", "tap", "bus_type", "gen_bus", "gen_p", "gen_vm"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("case", ["case14test", "case30test", "synthetic20", "synthetic10k"])
def test_symbolic_host_replay_solves_jacobian(case):
    """jgb_selfcheck_symbolic: ordering + fronts + assembly maps replayed on the host solve J x = f like SuperLU."""
    o = onr.newton_raphson(oracle_system(case))
    onr.mismatch(o)
    onr.fill_jacobian(o)
    n = o.dim
    grp = np.zeros(n, dtype=np.int64)
    for i in range(o.mdl.n):
        if o.pvpq[i] >= 0:
            grp[o.pvpq[i]] = i
        if o.pq[i] >= 0:
            grp[o.pq[i]] = i
    cp, rv = (o.j_colptr + 1).astype(np.int64), (o.j_rowval + 1).astype(np.int64)
    x, stats = np.zeros(n), np.zeros(8)
    rc = jgb200.load().jgb_selfcheck_symbolic(n, ptr(cp, C.c_int64), ptr(rv, C.c_int64), ptr(o.j_nzval, C.c_double),
                                              ptr(grp, C.c_int64), ptr(o.mismatch, C.c_double), ptr(x, C.c_double),
                                              ptr(stats, C.c_double))
    assert rc == 0
    J = onr.jacobian_csc(o)
    ref = spla.splu(J).solve(o.mismatch)
    assert np.abs(J @ x - o.mismatch).max() <= 1e-10 * max(1.0, np.abs(o.mismatch).max())
    np.testing.assert_allclose(x, ref, rtol=1e-7, atol=1e-10 * max(1.0, np.abs(ref).max()))
    assert stats[0] >= 1 and stats[1] >= 1 and stats[3] >= len(rv)


def test_symbolic_rejects_bad_input():
    lib = jgb200.load()
    assert lib.jgb_selfcheck_symbolic(0, None, None, None, None, None, None, None) == -1
    cp, rv = np.array([1, 2, 3], dtype=np.int64), np.array([1, 9], dtype=np.int64)
    assert lib.jgb_selfcheck_symbolic(2, ptr(cp, C.c_int64), ptr(rv, C.c_int64), None, None, None, None, None) == -4


@pytest.mark.parametrize("case", ["case14test", "synthetic20", "case_ACTIVSg10k"])
def test_c_oracle_matches_numpy_oracle(case):
    from oracle.fast import FastNR
    a = onr.newton_raphson(oracle_system(case))
    f = FastNR(a)
    assert f.mismatch() == onr.mismatch(a)
    f.jacobian()
    onr.fill_jacobian(a)
    assert np.array_equal(f.mism, a.mismatch) and np.array_equal(f.jnz, a.j_nzval)
    if case != "case_ACTIVSg10k":
        assert f.power_flow() and onr.power_flow(a)
        assert f.iteration == a.iteration
        np.testing.assert_allclose(f.vm, a.vm, atol=1e-12)


def test_wls_tables_match_oracle():
    """Product acWLS tables (H pattern, W, mean, type, index, range) == oracle's, every device class and flag."""
    ps, os_ = product_system("case14test"), oracle_system("case14test")
    ps.model = jgb200.ac_model(ps)
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    pw = jgb200.power(ps, o.vm, o.va)
    pw_o = post.powers(os_, o.mdl, o.vm, o.va)
    for k in pw:
        np.testing.assert_allclose(pw[k], pw_o[k], atol=1e-13)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_ammeter(mon, pw)
    jgb200.add_ammeter(mon, pw, square=True)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    for kw in (dict(polar=True), dict(polar=True, square=True), dict(polar=False), dict(polar=False, correlated=True)):
        jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(ps.n), branch=True, **kw)
    mon.watt["status"][5] = 0
    mon.pmu["mag_status"][3] = 0
    mon.pmu["ang_status"][40] = 0
    t = jgb200.ac_wls(ps, mon)
    g = owls.gauss_newton(os_, mon, o.mdl)
    ex = owls.export_one_based(g)
    assert t.m == g.m
    for k in ("h_colptr", "h_rowval", "type", "index", "range"):
        assert np.array_equal(getattr(t, k), ex[k]), k
    assert np.array_equal(t.mean, g.mean)
    assert np.array_equal(t.w_colptr - 1, g.w.indptr) and np.array_equal(t.w_rowval - 1, g.w.indices)
    np.testing.assert_allclose(t.w_nzval, g.w.data, rtol=1e-15)


def test_eligible_outages_counts():
    """12 567 non-islanding outages on the synthetic grid, 8 729 on ACTIVSg10k."""
    assert len(jgb200.eligible_outages(product_system("synthetic10k"))) == 12567
    assert len(jgb200.eligible_outages(product_system("case_ACTIVSg10k"))) == 8729
    assert 13 not in jgb200.eligible_outages(product_system("case14test"))


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 10000, 12567):
        for world in (1, 2, 3, 8):
            pieces = [jgb200.dist.shard_bounds(total, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == total
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1


def test_c_wls_oracle_matches_numpy_oracle():
    """oracle/csrc/oracle_wls.c (CPU baseline for the WLS numbers) == oracle/wls.py on the benchmark measurement set."""
    from oracle.fast import FastNR, FastWLS
    os_, ps = oracle_system("synthetic20"), product_system("synthetic20")
    ps.model = jgb200.ac_model(ps)
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    pw = jgb200.power(ps, o.vm, o.va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(0, ps.n, 10), polar=False)
    mon.watt["status"][7] = 0
    g = owls.gauss_newton(os_, mon, o.mdl, lu_options=FastNR.NOPIVOT)
    g.mean += 1e-3 * np.random.default_rng(3).standard_normal(g.m)
    fw = FastWLS(g)
    fr = FastWLS(g, refactor=True)
    assert fw.increment() == pytest.approx(owls.increment(g), rel=1e-12)
    assert np.array_equal(fw.res, g.residual) and np.array_equal(fw.hnz, g.h_nzval)
    assert fw.objective == pytest.approx(g.objective, rel=1e-13)
    fw.reset()
    assert fw.state_estimation() and owls.state_estimation(g)
    assert fw.iteration == g.iteration
    np.testing.assert_allclose(fw.vm, g.vm, atol=1e-12)
    # the refactorisation arm (symbolic once, numeric refactor afterwards, like ldlt! / lu!) gives the same run
    assert fr.state_estimation() and fr.iteration == g.iteration
    assert fr.refactor.symbolic_calls == 1 and fr.refactor.numeric_calls == fr.iteration
    np.testing.assert_allclose(fr.vm, g.vm, atol=1e-12)
    np.testing.assert_allclose(fr.va, g.va, atol=1e-12)


@pytest.mark.parametrize("case", ["case14test", "case30test", "synthetic20", "synthetic10k", "case_ACTIVSg10k"])
@pytest.mark.parametrize("opts", [None, "nopivot"])
def test_refactor_arm_matches_superlu(case, opts):
    """oracle/csrc/oracle_lu.c: the KLU-style numeric refactorisation (pattern, ordering and pivot order of the first
    factorisation reused) solves every later Newton system like a fresh SuperLU factorisation, for the default
    (COLAMD + partial pivoting) and the no-pivot settings; the power flows take the same iterations."""
    from oracle.fast import FastNR, Refactor
    lu_opts = FastNR.NOPIVOT if opts else None
    a = onr.newton_raphson(oracle_system(case))
    f = FastNR(a, lu_opts)
    r = Refactor(lu_opts)
    rng = np.random.default_rng(0)
    for it in range(3):
        f.mismatch()
        f.jacobian()
        J = sp.csc_matrix((f.jnz, f.jrowval, f.jcolptr32), shape=(len(f.mism),) * 2)
        r.factor(J)
        x = r.solve(f.mism)
        ref = spla.splu(J, **(lu_opts or {})).solve(f.mism)
        scale = max(1.0, np.abs(ref).max())
        # a frozen pivot order is less stable than fresh partial pivoting (klu_refactor has the same property): the big
        # case, perturbed far from its operating point, is held to 1e-7 relative residual, the small ones to 1e-9
        tol = 1e-7 if case == "case_ACTIVSg10k" else 1e-9
        assert np.abs(J @ x - f.mism).max() <= tol * max(1.0, np.abs(f.mism).max())
        np.testing.assert_allclose(x, ref, atol=10 * tol * scale, rtol=1e-5)
        f.vm += 1e-3 * rng.standard_normal(f.n)          # new values on the same pattern
        f.va += 1e-3 * rng.standard_normal(f.n)
    assert r.symbolic_calls == 1 and r.numeric_calls == 2
    if not case.endswith("10k"):
        g0, g1 = FastNR(a, lu_opts), FastNR(a, lu_opts, refactor=True)
        assert g0.power_flow() and g1.power_flow() and g0.iteration == g1.iteration
        np.testing.assert_allclose(g0.vm, g1.vm, atol=1e-12)
    # a singular matrix is reported, not silently factored
    Z = sp.csc_matrix((np.zeros_like(f.jnz), f.jrowval, f.jcolptr32), shape=J.shape)
    with pytest.raises(np.linalg.LinAlgError):
        r.factor(Z)


@pytest.mark.skipif(not os.path.exists("/root/reference/docs/src/examples/cases/hdf5/case_ACTIVSg10k.h5"),
                    reason="the reference's HDF5 cases are only present in the build container")
def test_product_hdf5_case_loader_matches_fixtures():
    """powerSystem("case.h5") (load.jl:141-289): the product's own reader against the committed fixtures."""
    import jgb200
    for case in ("case_ACTIVSg10k", "case_ACTIVSg70k"):
        a = jgb200.power_system(f"/root/reference/docs/src/examples/cases/hdf5/{case}.h5")
        b = product_system(case)
        assert (a.n, a.nbr, a.ngen, a.slack) == (b.n, b.nbr, b.ngen, b.slack)
        for k in ("bus_type", "pd", "qd", "gs", "bs", "vm", "va", "frm", "to", "r", "x", "g", "b", "tap", "shift", "status",
                  "gen_bus", "gen_p", "gen_q", "gen_vm", "gen_status", "gen_qmin", "gen_qmax"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_orderings_minimum_fill_beats_minimum_degree(monkeypatch):
    """Both orderings of the symbolic analysis (JGB_ORDER, JGB_FILL_K) replay to the same solution; on the 10k-bus
    Jacobian minimum local fill gives the smaller factor (the reason it is the default)."""
    o = onr.newton_raphson(oracle_system("synthetic10k"))
    onr.mismatch(o)
    onr.fill_jacobian(o)
    n = o.dim
    grp = np.zeros(n, dtype=np.int64)
    grp[o.pvpq[o.pvpq >= 0]] = np.flatnonzero(o.pvpq >= 0)
    grp[o.pq[o.pq >= 0]] = np.flatnonzero(o.pq >= 0)
    cp, rv = (o.j_colptr + 1).astype(np.int64), (o.j_rowval + 1).astype(np.int64)
    J = onr.jacobian_csc(o)
    out = {}
    for order, k in (("degree", "48"), ("fill", "48"), ("fill", "16")):
        monkeypatch.setenv("JGB_ORDER", order)
        monkeypatch.setenv("JGB_FILL_K", k)
        x, stats = np.zeros(n), np.zeros(8)
        rc = jgb200.load().jgb_selfcheck_symbolic(n, ptr(cp, C.c_int64), ptr(rv, C.c_int64), ptr(o.j_nzval, C.c_double),
                                                  ptr(grp, C.c_int64), ptr(o.mismatch, C.c_double), ptr(x, C.c_double),
                                                  ptr(stats, C.c_double))
        assert rc == 0
        assert np.abs(J @ x - o.mismatch).max() <= 1e-10 * max(1.0, np.abs(o.mismatch).max())
        out[(order, k)] = stats.copy()
    # stats: [fronts, levels, depths, nnz(L+U), flops, max front, u_size, upd_size]
    assert out[("fill", "48")][3] < 0.97 * out[("degree", "48")][3]
    assert out[("fill", "48")][4] < 0.90 * out[("degree", "48")][4]
    assert out[("fill", "16")][1] <= out[("fill", "48")][1]          # the capped rule keeps the tree shallower


@pytest.mark.parametrize("case", ["case14test", "case30test"])
def test_reactive_limit_host_logic_matches_oracle(case):
    """generator_power / reactive_limit of the host mirror (fed with the oracle's converged state instead of the device's)
    against the oracle's restatement: outputs, violation flags, bus types, slack and the rewritten bus supply."""
    from types import SimpleNamespace
    from jgb200.ac_power_flow import _initialize
    os_, ps = oracle_system(case), product_system(case)
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    ps.model = jgb200.ac_model(ps)
    bus_type, slack, _, _ = _initialize(ps)
    a = SimpleNamespace(system=ps, bus_type=bus_type, slack=slack)
    pw = jgb200.power(ps, o.vm, o.va)
    pg, qg = jgb200.generator_power(a, pw)
    pwo = post.powers(os_, o.mdl, o.vm, o.va)
    opg, oqg = post.generator_powers(os_, pwo["injection_active"], pwo["injection_reactive"], o.slack)
    np.testing.assert_allclose(pg, opg, rtol=0, atol=1e-12)
    np.testing.assert_allclose(qg, oqg, rtol=0, atol=1e-12)
    violate = jgb200.reactive_limit(a, pw)
    assert np.array_equal(violate, onr.reactive_limit(o)) and np.any(violate != 0)
    assert np.array_equal(ps.bus_type, os_.bus_type) and ps.slack == os_.slack
    sp_, sq_, _ = ps.supply
    np.testing.assert_allclose(sp_, os_.supply_p, rtol=0, atol=1e-12)
    np.testing.assert_allclose(sq_, os_.supply_q, rtol=0, atol=1e-12)
    np.testing.assert_allclose(ps.gen_q, os_.gen_q, rtol=0, atol=1e-12)


def _build_c_driver(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_c_test")
    libdir = os.path.dirname(jgb200.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c_test.c"), "-L", libdir, "-ljgb200", "-lm",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def test_plain_c_driver_links_every_header_symbol(tmp_path):
    """tests/abi_c_test.c is compiled against include/jgb200.h and linked to the .so: a symbol declared but not exported
    (or the reverse, through the table check below) fails here without ctypes in the loop."""
    import subprocess
    exe = _build_c_driver(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    first = out.stdout.splitlines()[0].split()
    assert first[0] == "symbols" and int(first[1]) == len(header_symbols())
    src = open(os.path.join(ROOT, "tests", "abi_c_test.c")).read()
    for name in header_symbols():
        assert f"SYM({name})" in src, f"{name} is missing from the C driver's table"


@pytest.mark.parametrize("case", ["case14test", "synthetic20", "synthetic10k", "case_ACTIVSg10k"])
def test_task_partition_host_replay(case, monkeypatch):
    """jgb_selfcheck_tasks: the task partition of the batch factorisation (fronts per CTA, shared-memory stack offsets,
    per-warp entry lists) replayed on the host solves J x = f like the plain multifrontal replay, for several caps."""
    o = onr.newton_raphson(oracle_system(case))
    onr.mismatch(o)
    onr.fill_jacobian(o)
    n = o.dim
    grp = np.zeros(n, dtype=np.int64)
    for i in range(o.mdl.n):
        if o.pvpq[i] >= 0:
            grp[o.pvpq[i]] = i
        if o.pq[i] >= 0:
            grp[o.pq[i]] = i
    cp, rv = (o.j_colptr + 1).astype(np.int64), (o.j_rowval + 1).astype(np.int64)
    J = onr.jacobian_csc(o)
    monkeypatch.setenv("JGB_TASKS", "1")
    for env in ({}, {"JGB_TASK_MAXNF": "12", "JGB_TASK_STACK": "96"}, {"JGB_TASK_BUNDLE": "6", "JGB_TASK_META": "1500"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        x, st = np.zeros(n), np.zeros(8)
        rc = jgb200.load().jgb_selfcheck_tasks(n, ptr(cp, C.c_int64), ptr(rv, C.c_int64), ptr(o.j_nzval, C.c_double),
                                               ptr(grp, C.c_int64), ptr(o.mismatch, C.c_double), ptr(x, C.c_double),
                                               ptr(st, C.c_double))
        assert rc == 0
        assert np.abs(J @ x - o.mismatch).max() <= 1e-10 * max(1.0, np.abs(o.mismatch).max())
        # stats: launches, tasks, fronts in tasks, update elements kept on chip, smem bytes, blob ints, fronts, upd
        assert st[0] >= 1 and st[2] >= 1 and st[4] <= 220 * 1024 and st[3] <= st[7]
        if case.endswith("10k"):
            assert st[2] > 0.8 * st[6] or env          # the default caps put most fronts into tasks
        for k in env:
            monkeypatch.delenv(k)


def test_measurement_file_layout_round_trip():
    """measurement(system, "file.h5") (src/measurement/load.jl:31-273): the datasets of saveMeasurement's layout
    (save.jl:40-118) rebuild the same Measurement, scalar datasets stand for constant vectors, the acWLS tables built
    from the reloaded set are identical, and a non-HDF5 extension is refused like load.jl:49-51."""
    ps, os_ = product_system("case14test"), oracle_system("case14test")
    ps.model = jgb200.ac_model(ps)
    o = onr.newton_raphson(os_)
    assert onr.power_flow(o)
    pw = jgb200.power(ps, o.vm, o.va)
    mon = jgb200.measurement(ps)
    jgb200.add_voltmeter(mon, o.vm)
    jgb200.add_ammeter(mon, pw, square=True)
    jgb200.add_wattmeter(mon, pw)
    jgb200.add_varmeter(mon, pw)
    jgb200.add_pmu(mon, pw, o.vm, o.va, buses=range(ps.n), branch=True, polar=False, correlated=True)
    mon.watt["status"][5] = 0
    data = jgb200.measurement_to_arrays(mon)
    assert data["voltmeter/layout/index"].min() == 1
    assert np.array_equal(data["pmu/layout/to"], ~mon.pmu["frm"] & ~mon.pmu["bus"])
    data["voltmeter/magnitude/variance"] = np.float64(mon.volt["variance"][0])      # scalar dataset = constant vector
    back = jgb200.measurement_from_arrays(ps, data)
    for dev in ("volt", "amp", "watt", "var", "pmu"):
        for k, v in getattr(mon, dev).items():
            assert np.array_equal(getattr(back, dev)[k], v), (dev, k)
    t0, t1 = jgb200.ac_wls(ps, mon), jgb200.ac_wls(ps, back)
    for k in ("h_colptr", "h_rowval", "type", "index", "range", "mean", "w_nzval"):
        assert np.array_equal(getattr(t0, k), getattr(t1, k)), k
    with pytest.raises(ValueError):
        jgb200.load_measurement(ps, "monitoring.m")
    bad = dict(data)
    bad["voltmeter/layout/index"] = data["voltmeter/layout/index"] + ps.n
    with pytest.raises(ValueError):
        jgb200.measurement_from_arrays(ps, bad)


def test_reference_measurement_file():
    """The reference's own measurement file (src/data/monitoring.h5, written by saveMeasurement for src/data/case14.h5;
    fixture tests/golden/monitoring14.json): the loader of measurement(system, "file.h5") (load.jl:31-273) rebuilds the
    five device classes with the counts of the file's attributes, scalar datasets as constant vectors and Bool bitfields as
    flags; the acWLS tables of the product and of the oracle agree on it, and the oracle's Gauss-Newton estimation of this
    real measurement set converges. Where the reference tree is present the HDF5 file itself is read and compared."""
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "monitoring14.json")
    with open(path) as fh:
        g = json.load(fh)
    ps = jgb200.power_system(path)
    ps.model = jgb200.ac_model(ps)
    data = {k: np.asarray(v) for k, v in g["datasets"].items()}
    mon = jgb200.measurement_from_arrays(ps, data)
    at = g["attrs"]
    assert len(mon.volt["index"]) == at["number of voltmeters"] == 14
    assert len(mon.amp["index"]) == at["number of ammeters"] == 2 * len(ps.status)
    assert len(mon.watt["index"]) == at["number of wattmeters"] == ps.n + 2 * len(ps.status)
    assert len(mon.var["index"]) == at["number of varmeters"] and len(mon.pmu["index"]) == at["number of pmus"]
    assert mon.volt["variance"].shape == (14,) and np.all(mon.volt["variance"] == 1e-4)      # scalar dataset
    assert mon.amp["frm"].dtype == bool and mon.amp["frm"][:4].tolist() == [True, False, True, False]
    assert mon.amp["square"].all() and not mon.pmu["polar"].any()
    assert mon.watt["bus"][:ps.n].all() and not mon.watt["bus"][ps.n:].any()
    ref_file = "/root/reference/src/data/monitoring.h5"
    if os.path.exists(ref_file):
        mon2 = jgb200.load_measurement(ps, ref_file)
        for dev in ("volt", "amp", "watt", "var", "pmu"):
            for k, v in getattr(mon, dev).items():
                assert np.array_equal(getattr(mon2, dev)[k], v), (dev, k)
    # the same tables on both sides, and a converging estimation on the CPU oracle
    os_ = oracle.system_from_arrays(g["system"])
    mdl = oracle.ac_model(os_)
    t = jgb200.ac_wls(ps, mon)
    gn = owls.gauss_newton(os_, mon, mdl)
    ex = owls.export_one_based(gn)
    assert t.m == gn.m == 14 + 40 + 54 + 54 + 2 * 54
    for k in ("h_colptr", "h_rowval", "type", "index", "range"):
        assert np.array_equal(getattr(t, k), ex[k]), k
    assert np.array_equal(t.mean, gn.mean)
    assert owls.state_estimation(gn, iteration=40, tolerance=1e-8)
    assert np.abs(gn.vm - mon.volt["mean"]).max() < 0.05          # the estimate sits on the (noisy) voltmeter readings
