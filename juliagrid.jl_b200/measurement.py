"""Host-side Measurement container, exact-measurement generators and `acWLS` table builder of the product.

In the Julia drop-in all of this stays in JuliaGrid (`src/measurement/*`, `acWLS` src/stateEstimation/
acStateEstimation.jl:77-259) and only the resulting tables cross the C ABI; the Python host mirror needs its own
copy to drive the same ABI. Vectorised NumPy; rows are produced in the reference's device order
(voltmeter | ammeter | wattmeter | varmeter | PMU with 2 rows each) with the reference's type codes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .cases import PowerSystem
from .model import AcModel, ac_model

_KEYS = {
    "volt": ("index", "mean", "variance", "status"),
    "amp": ("index", "frm", "square", "mean", "variance", "status"),
    "watt": ("index", "bus", "frm", "mean", "variance", "status"),
    "var": ("index", "bus", "frm", "mean", "variance", "status"),
    "pmu": ("index", "bus", "frm", "polar", "square", "correlated", "mag_mean", "mag_variance", "mag_status",
            "ang_mean", "ang_variance", "ang_status"),
}
_INT = {"index", "status", "mag_status", "ang_status"}
_BOOL = {"bus", "frm", "polar", "square", "correlated"}


def _dtype(k):
    return np.int64 if k in _INT else bool if k in _BOOL else np.float64


def _blank(dev):
    return {k: np.zeros(0, dtype=_dtype(k)) for k in _KEYS[dev]}


@dataclass
class Measurement:
    """monitoring: per device class arrays like the reference's Voltmeter/Ammeter/Wattmeter/Varmeter/PMU structs
    (src/definition/system.jl:274-430); `index` is a 0-based bus or branch index."""
    system: PowerSystem
    volt: dict = field(default_factory=lambda: _blank("volt"))
    amp: dict = field(default_factory=lambda: _blank("amp"))
    watt: dict = field(default_factory=lambda: _blank("watt"))
    var: dict = field(default_factory=lambda: _blank("var"))
    pmu: dict = field(default_factory=lambda: _blank("pmu"))

    def _append(self, dev: str, **cols):
        d = getattr(self, dev)
        n = len(next(iter(cols.values())))
        for k in _KEYS[dev]:
            v = np.broadcast_to(np.asarray(cols[k], dtype=_dtype(k)), (n,))
            d[k] = np.concatenate([d[k], v])


def measurement(system: PowerSystem) -> Measurement:
    return Measurement(system)


_H5_LAYOUT = {      # HDF5 group -> (device key, value group(s), layout flags), src/measurement/load.jl:190-273
    "voltmeter": ("volt", {"magnitude": ("mean", "variance", "status")}, ()),
    "ammeter": ("amp", {"magnitude": ("mean", "variance", "status")}, ("from", "square")),
    "wattmeter": ("watt", {"active": ("mean", "variance", "status")}, ("bus", "from")),
    "varmeter": ("var", {"reactive": ("mean", "variance", "status")}, ("bus", "from")),
    "pmu": ("pmu", {"magnitude": ("mag_mean", "mag_variance", "mag_status"),
                    "angle": ("ang_mean", "ang_variance", "ang_status")},
            ("bus", "from", "polar", "square", "correlated")),
}


def measurement_from_arrays(system: PowerSystem, data: dict) -> Measurement:
    """measurement(system, "file.h5") (src/measurement/load.jl:31-273) from the file's datasets keyed by their HDF5 path
    ("wattmeter/active/mean", "pmu/layout/polar", ...): positional 1-based bus / branch indices, per-unit / radian
    values, a scalar dataset stands for a constant vector. Labels are not kept (the mirror addresses meters by position)."""
    mon = Measurement(system)
    for group, (dev, values, flags) in _H5_LAYOUT.items():
        key = group + "/layout/index"
        if key not in data:
            continue
        index = np.atleast_1d(np.asarray(data[key])).astype(np.int64) - 1
        count = len(index)

        def vec(path, dtype):
            v = np.asarray(data[path])
            if v.ndim == 0 or (v.size == 1 and count != 1):
                return np.full(count, v.reshape(-1)[0], dtype=dtype)
            return v.astype(dtype)

        cols = {"index": index}
        for vgroup, (kmean, kvar, kstat) in values.items():
            cols[kmean] = vec(f"{group}/{vgroup}/mean", np.float64)
            cols[kvar] = vec(f"{group}/{vgroup}/variance", np.float64)
            cols[kstat] = vec(f"{group}/{vgroup}/status", np.int64)
        for flag in flags:
            cols["frm" if flag == "from" else flag] = vec(f"{group}/layout/{flag}", bool)
        onbus = np.ones(count, dtype=bool) if dev == "volt" else cols.get("bus", np.zeros(count, dtype=bool))
        limit = np.where(onbus, system.n, len(system.status))
        if np.any(index < 0) or np.any(index >= limit):
            raise ValueError(f"{group}: bus / branch index outside the power system")
        mon._append(dev, **cols)
    return mon


def measurement_to_arrays(mon: Measurement) -> dict:
    """The datasets `saveMeasurement` writes (src/measurement/save.jl:40-118), keyed by their HDF5 path: the inverse of
    `measurement_from_arrays` (1-based positional indices, `to` = not `from` for branch meters)."""
    out = {}
    for group, (dev, values, flags) in _H5_LAYOUT.items():
        d = getattr(mon, dev)
        if len(d["index"]) == 0:
            continue
        out[f"{group}/layout/index"] = d["index"] + 1
        for vgroup, keys in values.items():
            for name, k in zip(("mean", "variance", "status"), keys):
                out[f"{group}/{vgroup}/{name}"] = d[k].copy()
        for flag in flags:
            out[f"{group}/layout/{flag}"] = d["frm" if flag == "from" else flag].copy()
        if "from" in flags:
            bus = d["bus"] if "bus" in d else np.zeros(len(d["index"]), dtype=bool)
            out[f"{group}/layout/to"] = ~d["frm"] & ~bus
    return out


def load_measurement(system: PowerSystem, path: str) -> Measurement:
    """measurement(system, path): JuliaGrid's HDF5 measurement files (`saveMeasurement`, src/measurement/save.jl) read with
    the mirror's own pure-Python HDF5 reader; `.m` is not a measurement format (load.jl:49-51)."""
    if not path.endswith(".h5"):
        raise ValueError("The extension of the measurement file must be .h5")
    from .h5 import H5File
    f = H5File(path)
    data = {}
    for group, (_, values, flags) in _H5_LAYOUT.items():
        if group not in f.keys("/"):
            continue
        names = [f"{group}/layout/index"] + [f"{group}/layout/{fl}" for fl in flags]
        for vgroup in values:
            names += [f"{group}/{vgroup}/{k}" for k in ("mean", "variance", "status")]
        for nm in names:
            data[nm] = f["/" + nm]
    return measurement_from_arrays(system, data)


def power(system: PowerSystem, vm, va) -> dict:
    """power!/current! values the generators consume (src/postprocessing/acAnalysis.jl:30-170, 672-723)."""
    mdl: AcModel = system.model or ac_model(system)
    v = vm * np.exp(1j * va)
    n = system.n
    col = np.repeat(np.arange(n), np.diff(mdl.colptr))
    inj_i = np.zeros(n, dtype=complex)
    np.add.at(inj_i, col, mdl.nzval_t * v[mdl.rowval - 1])
    s_inj = v * np.conj(inj_i)
    vi, vj = v[system.frm], v[system.to]
    i_f = mdl.y_ff * vi + mdl.y_ft * vj
    i_t = mdl.y_tf * vi + mdl.y_tt * vj
    s_f, s_t = vi * np.conj(i_f), vj * np.conj(i_t)
    on = system.status == 1
    z = lambda a: np.where(on, a, 0.0)
    return {"injection_active": s_inj.real, "injection_reactive": s_inj.imag,
            "from_active": z(s_f.real), "from_reactive": z(s_f.imag), "to_active": z(s_t.real),
            "to_reactive": z(s_t.imag), "from_current_magnitude": z(np.abs(i_f)),
            "from_current_angle": z(np.angle(i_f)), "to_current_magnitude": z(np.abs(i_t)),
            "to_current_angle": z(np.angle(i_t))}


def add_voltmeter(mon: Measurement, vm, variance=1e-4, status=1):
    """addVoltmeter!(monitoring, analysis): one voltmeter per bus."""
    n = mon.system.n
    mon._append("volt", index=np.arange(n), mean=vm, variance=variance, status=status)


def _interleave(a, b):
    out = np.empty(2 * len(a), dtype=np.result_type(a, b))
    out[0::2], out[1::2] = a, b
    return out


def add_ammeter(mon: Measurement, pw: dict, variance=1e-4, status=1, square=False):
    """addAmmeter!(monitoring, analysis): per in-service branch a from-end then a to-end ammeter."""
    on = np.flatnonzero(mon.system.status == 1)
    k = np.repeat(on, 2)
    frm = np.tile([True, False], len(on))
    mean = _interleave(pw["from_current_magnitude"][on], pw["to_current_magnitude"][on])
    mon._append("amp", index=k, frm=frm, square=square, mean=mean, variance=variance, status=status)


def _add_power(mon, dev, inj, fr, to, variance, status, bus, branch):
    s = mon.system
    if bus:
        mon._append(dev, index=np.arange(s.n), bus=True, frm=False, mean=inj, variance=variance, status=status)
    if branch:
        on = np.flatnonzero(s.status == 1)
        mon._append(dev, index=np.repeat(on, 2), bus=False, frm=np.tile([True, False], len(on)),
                    mean=_interleave(fr[on], to[on]), variance=variance, status=status)


def add_wattmeter(mon: Measurement, pw: dict, variance=1e-4, status=1, bus=True, branch=True):
    """addWattmeter!(monitoring, analysis) (measurement/powermeter.jl:479-524): every bus, then per in-service
    branch its from-end and to-end."""
    _add_power(mon, "watt", pw["injection_active"], pw["from_active"], pw["to_active"], variance, status, bus, branch)


def add_varmeter(mon: Measurement, pw: dict, variance=1e-4, status=1, bus=True, branch=True):
    _add_power(mon, "var", pw["injection_reactive"], pw["from_reactive"], pw["to_reactive"], variance, status, bus,
               branch)


def add_pmu(mon: Measurement, pw: dict, vm, va, buses=(), branch=False, polar=True, square=False, correlated=False,
            variance_magnitude=1e-8, variance_angle=1e-8, status=1):
    """addPmu!(monitoring, analysis): bus phasors on `buses`, then (optionally) per in-service branch from / to."""
    buses = np.asarray(list(buses), dtype=np.int64)
    if len(buses):
        mon._append("pmu", index=buses, bus=True, frm=False, polar=polar, square=False, correlated=correlated,
                    mag_mean=vm[buses], mag_variance=variance_magnitude, mag_status=status, ang_mean=va[buses],
                    ang_variance=variance_angle, ang_status=status)
    if branch:
        on = np.flatnonzero(mon.system.status == 1)
        mon._append("pmu", index=np.repeat(on, 2), bus=False, frm=np.tile([True, False], len(on)), polar=polar,
                    square=square, correlated=correlated,
                    mag_mean=_interleave(pw["from_current_magnitude"][on], pw["to_current_magnitude"][on]),
                    mag_variance=variance_magnitude, mag_status=status,
                    ang_mean=_interleave(pw["from_current_angle"][on], pw["to_current_angle"][on]),
                    ang_variance=variance_angle, ang_status=status)


@dataclass
class WlsTables:
    """What acWLS returns (acStateEstimation.jl:238-258), in the reference's 1-based CSC layout."""
    m: int
    h_colptr: np.ndarray
    h_rowval: np.ndarray
    w_colptr: np.ndarray
    w_rowval: np.ndarray
    w_nzval: np.ndarray
    mean: np.ndarray
    type: np.ndarray
    index: np.ndarray       # 1-based
    range: np.ndarray       # 1-based
    correlated: bool


def ac_wls(system: PowerSystem, mon: Measurement) -> WlsTables:
    mdl: AcModel = system.model
    n = system.n
    volt, amp, watt, var, pmu = mon.volt, mon.amp, mon.watt, mon.var, mon.pmu
    nv, na, nw, nq, npmu = (len(d["index"]) for d in (volt, amp, watt, var, pmu))
    m = nv + na + nw + nq + 2 * npmu
    mean = np.zeros(m)
    typ = np.zeros(m, dtype=np.int8)
    idx = np.zeros(m, dtype=np.int64)
    prec_diag = np.zeros(m)
    rows, cols = [], []
    deg = np.diff(mdl.colptr)

    def branch_cols(r, k):
        f, t = system.frm[k], system.to[k]
        rows.append(np.repeat(r, 4))
        cols.append(np.stack([f, t, f + n, t + n], axis=1).ravel())

    def bus_cols(r, i):
        cnt = deg[i]
        rr = np.repeat(r, cnt)
        start = np.repeat(mdl.colptr[i] - 1, cnt)
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        j = mdl.rowval[start + within] - 1
        rows.append(np.repeat(rr, 2))
        cols.append(np.stack([j, j + n], axis=1).ravel())

    off = 0
    rng = np.zeros(6, dtype=np.int64)
    # voltmeters: code 1, one constant entry in the V column
    r = off + np.arange(nv)
    st = volt["status"]
    mean[r] = st * volt["mean"]
    typ[r] = st * 1
    idx[r] = volt["index"]
    prec_diag[r] = 1 / volt["variance"] if nv else 0
    rows.append(r)
    cols.append(volt["index"] + n)
    off += nv
    rng[1] = off
    # ammeters: codes 2/3 or 4/5
    r = off + np.arange(na)
    st, sq = amp["status"], amp["square"]
    mean[r] = st * np.where(sq, amp["mean"] ** 2, amp["mean"])
    prec_diag[r] = 1 / np.where(sq, 4 * amp["mean"] ** 2 * amp["variance"], amp["variance"]) if na else 0
    typ[r] = st * np.where(amp["frm"], np.where(sq, 4, 2), np.where(sq, 5, 3))
    idx[r] = amp["index"]
    branch_cols(r, amp["index"])
    off += na
    rng[2] = off
    for dev, cbus, cfrom, cto, slot in ((watt, 6, 7, 8, 3), (var, 9, 10, 11, 4)):
        cnt = len(dev["index"])
        r = off + np.arange(cnt)
        st = dev["status"]
        mean[r] = st * dev["mean"]
        prec_diag[r] = 1 / dev["variance"] if cnt else 0
        typ[r] = st * np.where(dev["bus"], cbus, np.where(dev["frm"], cfrom, cto))
        idx[r] = dev["index"]
        b = dev["bus"]
        if b.any():
            bus_cols(r[b], dev["index"][b])
        if (~b).any():
            branch_cols(r[~b], dev["index"][~b])
        off += cnt
        rng[slot] = off
    # PMUs: two rows each
    r0 = off + 2 * np.arange(npmu)
    r1 = r0 + 1
    sm, sa = pmu["mag_status"], pmu["ang_status"]
    polar, sq, bus, frm, corr = pmu["polar"], pmu["square"], pmu["bus"], pmu["frm"], pmu["correlated"]
    k = pmu["index"]
    w_extra = []
    if npmu:
        idx[r0] = k
        idx[r1] = k
        s_, c_ = np.sin(pmu["ang_mean"]), np.cos(pmu["ang_mean"])
        both = sm * sa
        mean[r0] = np.where(polar, sm * np.where(sq, pmu["mag_mean"] ** 2, pmu["mag_mean"]), both * pmu["mag_mean"] * c_)
        mean[r1] = np.where(polar, sa * pmu["ang_mean"], both * pmu["mag_mean"] * s_)
        var_re = pmu["mag_variance"] * c_ ** 2 + pmu["ang_variance"] * (pmu["mag_mean"] * s_) ** 2
        var_im = pmu["mag_variance"] * s_ ** 2 + pmu["ang_variance"] * (pmu["mag_mean"] * c_) ** 2
        prec_diag[r0] = np.where(polar, 1 / np.where(sq, 4 * pmu["mag_mean"] ** 2 * pmu["mag_variance"],
                                                     pmu["mag_variance"]), 1 / var_re)
        prec_diag[r1] = np.where(polar, 1 / pmu["ang_variance"], 1 / var_im)
        cm = (~polar) & corr
        if cm.any():      # covariancePmu + precision! (equations.jl:591-666)
            l1inv = 1 / np.sqrt(var_re[cm])
            l2 = s_[cm] * c_[cm] * (pmu["mag_variance"][cm] - pmu["ang_variance"][cm] * pmu["mag_mean"][cm] ** 2) * l1inv
            l3inv2 = 1 / (var_im[cm] - l2 ** 2)
            offd = (-l2 * l1inv) * l3inv2
            prec_diag[r0[cm]] = (l1inv - l2 * offd) * l1inv
            prec_diag[r1[cm]] = l3inv2
            w_extra = [(r0[cm], r1[cm], offd), (r1[cm], r0[cm], offd)]
        code0 = np.where(polar, np.where(bus, 12, np.where(frm, np.where(sq, 4, 2), np.where(sq, 5, 3))),
                         np.where(bus, 16, np.where(frm, 18, 19)))
        code1 = np.where(polar, np.where(bus, 13, np.where(frm, 14, 15)), np.where(bus, 17, np.where(frm, 20, 21)))
        typ[r0] = np.where(polar, sm, both) * code0
        typ[r1] = np.where(polar, sa, both) * code1
        pb = polar & bus
        rows += [r0[pb], r1[pb]]
        cols += [k[pb] + n, k[pb]]
        rb = (~polar) & bus
        for rr in (r0[rb], r1[rb]):
            rows.append(np.repeat(rr, 2))
            cols.append(np.stack([k[rb], k[rb] + n], axis=1).ravel())
        nb = ~bus
        if nb.any():
            branch_cols(r0[nb], k[nb])
            branch_cols(r1[nb], k[nb])
    off += 2 * npmu
    rng[5] = off

    rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    H = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(m, 2 * n))
    H.sort_indices()
    wr = [np.arange(m)] + [e[0] for e in w_extra]
    wc = [np.arange(m)] + [e[1] for e in w_extra]
    wv = [prec_diag] + [e[2] for e in w_extra]
    W = sp.csc_matrix((np.concatenate(wv), (np.concatenate(wr), np.concatenate(wc))), shape=(m, m))
    W.sort_indices()
    return WlsTables(m, H.indptr.astype(np.int64) + 1, H.indices.astype(np.int64) + 1,
                     W.indptr.astype(np.int64) + 1, W.indices.astype(np.int64) + 1, W.data.astype(np.float64), mean,
                     typ, idx + 1, rng + 1, bool(len(w_extra)))
