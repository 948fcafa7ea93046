"""Oracle restatement of the Gauss-Newton WLS AC state estimation (TEST INFRASTRUCTURE).

Reference lines followed (all under /root/reference/src):
  gaussNewton / acWLS                stateEstimation/acStateEstimation.jl:43-259
  oneIndices!/twoIndices!/fourIndices!/nthIndices!     :1130-1238
  normalEquation!                    stateEstimation/acStateEstimation.jl:261-583
  increment! (Normal)                stateEstimation/acStateEstimation.jl:878-904
  solve!                             stateEstimation/acStateEstimation.jl:1035-1047
  stateEstimation! loop              stateEstimation/acStateEstimation.jl:1286-1329
  scalar formulas                    backend/equations.jl:20-698
  varianceSquare / if2exp            measurement/utility.jl:115-129
  removeColumn / restoreColumn!      backend/sparse.jl:155-188
  SpGEMM + factorisation             Julia SparseArrays / SuiteSparse (third party); here SciPy / SuperLU.

`normal_equation` walks rows (each row evaluates its function once and writes all of its H entries);
the reference walks the theta-columns of H and re-evaluates every branch row twice — the values are
the same expressions on the same inputs, so results agree to the last bit except `objective`, whose
summation order differs (checked to 1e-12 relative).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import sin, cos, sqrt, atan2
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .system import System
from .model import AcModel, ac_model


# ----------------------------------------------------------------------------- measurement container
@dataclass
class Measurement:
    """Arrays mirroring the reference `Measurement` (definition/system.jl:274-430), 0-based indices.

    voltmeter: index(bus), mean, variance, status
    ammeter:   index(branch), frm(bool: from-end else to-end), square, mean, variance, status
    wattmeter / varmeter: index(bus or branch), bus(bool), frm(bool), mean, variance, status
    pmu: index, bus, frm, polar, square, correlated, mag_mean/var/status, ang_mean/var/status
    """
    volt: dict = field(default_factory=lambda: _empty(("index", "mean", "variance", "status")))
    amp: dict = field(default_factory=lambda: _empty(("index", "frm", "square", "mean", "variance", "status")))
    watt: dict = field(default_factory=lambda: _empty(("index", "bus", "frm", "mean", "variance", "status")))
    var: dict = field(default_factory=lambda: _empty(("index", "bus", "frm", "mean", "variance", "status")))
    pmu: dict = field(default_factory=lambda: _empty(
        ("index", "bus", "frm", "polar", "square", "correlated",
         "mag_mean", "mag_variance", "mag_status", "ang_mean", "ang_variance", "ang_status")))

    def finalize(self):
        ints = {"index", "status", "mag_status", "ang_status"}
        bools = {"bus", "frm", "polar", "square", "correlated"}
        for dev in (self.volt, self.amp, self.watt, self.var, self.pmu):
            for k, v in dev.items():
                if k in ints:
                    dev[k] = np.asarray(v, dtype=np.int64)
                elif k in bools:
                    dev[k] = np.asarray(v, dtype=bool)
                else:
                    dev[k] = np.asarray(v, dtype=float)
        return self


def _empty(keys):
    return {k: [] for k in keys}


def _push(dev: dict, **kw):
    for k, v in kw.items():
        dev[k].append(v)


def measurements_from_solution(sys: System, pw: dict, vm: np.ndarray, va: np.ndarray, *,
                               volt=True, watt=True, var=True, amp=False, amp_square=False,
                               pmu_bus=(), pmu_branch=False, pmu_polar=True, pmu_square=False,
                               pmu_correlated=False, power_bus=True, power_branch=True,
                               var_volt=1e-4, var_power=1e-4, var_amp=1e-4, var_pmu_mag=1e-8, var_pmu_ang=1e-8
                               ) -> Measurement:
    """Exact measurement set from a solved state, in the row order the reference's generators produce:
    `addVoltmeter!(monitoring, analysis)` (every bus), `addAmmeter!` (per in-service branch: from, to),
    `addWattmeter!/addVarmeter!` (`measurement/powermeter.jl:479-524`: every bus, then per in-service
    branch from-end then to-end), `addPmu!` (every listed bus; per in-service branch from, to).
    Default variances are the reference templates' (`definition/internal.jl:173-232`)."""
    m = Measurement()
    on = [k for k in range(sys.nbr) if sys.status[k] == 1]
    if volt:
        for i in range(sys.n):
            _push(m.volt, index=i, mean=vm[i], variance=var_volt, status=1)
    if amp:
        for k in on:
            _push(m.amp, index=k, frm=True, square=amp_square, mean=pw["from_current_magnitude"][k],
                  variance=var_amp, status=1)
            _push(m.amp, index=k, frm=False, square=amp_square, mean=pw["to_current_magnitude"][k],
                  variance=var_amp, status=1)
    for flag, dev, inj, fr, to in ((watt, m.watt, "injection_active", "from_active", "to_active"),
                                   (var, m.var, "injection_reactive", "from_reactive", "to_reactive")):
        if not flag:
            continue
        for i in range(sys.n if power_bus else 0):
            _push(dev, index=i, bus=True, frm=False, mean=pw[inj][i], variance=var_power, status=1)
        for k in (on if power_branch else ()):
            _push(dev, index=k, bus=False, frm=True, mean=pw[fr][k], variance=var_power, status=1)
            _push(dev, index=k, bus=False, frm=False, mean=pw[to][k], variance=var_power, status=1)
    for i in pmu_bus:
        _push(m.pmu, index=int(i), bus=True, frm=False, polar=pmu_polar, square=False, correlated=pmu_correlated,
              mag_mean=vm[i], mag_variance=var_pmu_mag, mag_status=1,
              ang_mean=va[i], ang_variance=var_pmu_ang, ang_status=1)
    if pmu_branch:
        for k in on:
            for frm, mg, an in ((True, "from_current_magnitude", "from_current_angle"),
                                (False, "to_current_magnitude", "to_current_angle")):
                _push(m.pmu, index=k, bus=False, frm=frm, polar=pmu_polar, square=pmu_square,
                      correlated=pmu_correlated, mag_mean=pw[mg][k], mag_variance=var_pmu_mag, mag_status=1,
                      ang_mean=pw[an][k], ang_variance=var_pmu_ang, ang_status=1)
    return m.finalize()


# ----------------------------------------------------------------------------- analysis container
@dataclass
class GaussNewton:
    sys: System
    mdl: AcModel
    slack: int
    vm: np.ndarray
    va: np.ndarray
    m: int                      # rows
    h_colptr: np.ndarray        # CSC pattern of the Jacobian H (m x 2n), 0-based, rows sorted
    h_rowval: np.ndarray
    h_nzval: np.ndarray
    w: sp.csc_matrix            # precision (m x m)
    mean: np.ndarray
    residual: np.ndarray
    type: np.ndarray            # int8 codes 0..21
    index: np.ndarray           # bus or branch index, 0-based
    range: np.ndarray           # 6 row offsets, 0-based
    increment: np.ndarray
    objective: float = 0.0
    iteration: int = 0
    correlated: bool = False
    lu_options: dict | None = None
    gain: sp.csc_matrix | None = None

    def h_position(self, row: int, col: int) -> int:
        lo, hi = self.h_colptr[col], self.h_colptr[col + 1]
        k = lo + int(np.searchsorted(self.h_rowval[lo:hi], row))
        assert k < hi and self.h_rowval[k] == row
        return int(k)

    def jacobian(self) -> sp.csc_matrix:
        return sp.csc_matrix((self.h_nzval.copy(), self.h_rowval, self.h_colptr), shape=(self.m, 2 * self.sys.n))


def _variance_square(mean, variance, square):            # measurement/utility.jl:120-126
    return 4 * mean ** 2 * variance if square else variance


def _variance_pmu(vmag, vang, mag, c, s):                # equations.jl:576-588
    var_re = vmag * c ** 2 + vang * (mag * s) ** 2
    var_im = vmag * s ** 2 + vang * (mag * c) ** 2
    return var_re, var_im


def gauss_newton(sys: System, meas: Measurement, mdl: AcModel | None = None,
                 lu_options: dict | None = None) -> GaussNewton:
    """gaussNewton(monitoring) -> acWLS (acStateEstimation.jl:43-259)."""
    if mdl is None:
        mdl = ac_model(sys)
    n = sys.n
    volt, amp, watt, var, pmu = meas.volt, meas.amp, meas.watt, meas.var, meas.pmu
    nv, na, nw, nq, npmu = (len(d["index"]) for d in (volt, amp, watt, var, pmu))
    total = nv + na + nw + nq + 2 * npmu

    rows, cols, vals = [], [], []          # H triplets (SparseModel jcb)
    prow, pcol, pval = [], [], []          # W triplets (SparseModel pcs)
    mean = np.zeros(total)
    typ = np.zeros(total, dtype=np.int8)
    idx = np.zeros(total, dtype=np.int64)
    rng = np.zeros(6, dtype=np.int64)
    state = {"row": 0}
    correlated = False

    def prec(variance):                    # precision!(pcs, variance) (equations.jl:668-677)
        r = state.get("prow", 0)
        prow.append(r); pcol.append(r); pval.append(1 / variance)
        state["prow"] = r + 1

    def one(status, col, bus, code):       # oneIndices! (:1130-1150)
        r = state["row"]
        typ[r] = status * code; idx[r] = bus
        rows.append(r); cols.append(col); vals.append(float(status))
        state["row"] = r + 1

    def two(status, bus, code):            # twoIndices! (:1152-1174)
        r = state["row"]
        typ[r] = status * code; idx[r] = bus
        rows.extend((r, r)); cols.extend((bus, bus + n)); vals.extend((0.0, 0.0))
        state["row"] = r + 1

    def four(status, location, br, code1, code2):   # fourIndices! (:1176-1210)
        r = state["row"]
        idx[r] = br
        typ[r] = status * (code1 if location else code2)
        f, t = int(sys.frm[br]), int(sys.to[br])
        rows.extend((r, r, r, r)); cols.extend((f, t, f + n, t + n)); vals.extend((0.0,) * 4)
        state["row"] = r + 1

    def nth(status, bus, code):            # nthIndices! (:1212-1238)
        r = state["row"]
        typ[r] = status * code; idx[r] = bus
        for p in range(mdl.colptr[bus], mdl.colptr[bus + 1]):
            j = int(mdl.rowval[p])
            rows.extend((r, r)); cols.extend((j, j + n)); vals.extend((0.0, 0.0))
        state["row"] = r + 1

    for i in range(nv):
        st = int(volt["status"][i]); k = int(volt["index"][i])
        mean[state["row"]] = st * volt["mean"][i]
        prec(volt["variance"][i])
        one(st, k + n, k, 1)
    rng[1] = state["row"]

    for i in range(na):
        st = int(amp["status"][i]); k = int(amp["index"][i]); sq = bool(amp["square"][i])
        mean[state["row"]] = st * (amp["mean"][i] ** (2 if sq else 1))
        prec(_variance_square(amp["mean"][i], amp["variance"][i], sq))
        if sq:
            four(st, bool(amp["frm"][i]), k, 4, 5)
        else:
            four(st, bool(amp["frm"][i]), k, 2, 3)
    rng[2] = state["row"]

    for dev, cbus, cfrom, cto, slot in ((watt, 6, 7, 8, 3), (var, 9, 10, 11, 4)):
        for i in range(len(dev["index"])):
            st = int(dev["status"][i]); k = int(dev["index"][i])
            mean[state["row"]] = st * dev["mean"][i]
            prec(dev["variance"][i])
            if dev["bus"][i]:
                nth(st, k, cbus)
            else:
                four(st, bool(dev["frm"][i]), k, cfrom, cto)
        rng[slot] = state["row"]

    for i in range(npmu):
        sm = int(pmu["mag_status"][i]); sa = int(pmu["ang_status"][i]); k = int(pmu["index"][i])
        r = state["row"]
        if pmu["polar"][i]:
            sq = bool(pmu["square"][i])
            mean[r] = sm * (pmu["mag_mean"][i] ** (2 if sq else 1))
            prec(_variance_square(pmu["mag_mean"][i], pmu["mag_variance"][i], sq))
            mean[r + 1] = sa * pmu["ang_mean"][i]
            prec(pmu["ang_variance"][i])
            if pmu["bus"][i]:
                one(sm, k + n, k, 12)
                one(sa, k, k, 13)
            else:
                if sq:
                    four(sm, bool(pmu["frm"][i]), k, 4, 5)
                else:
                    four(sm, bool(pmu["frm"][i]), k, 2, 3)
                four(sa, bool(pmu["frm"][i]), k, 14, 15)
        else:
            s, c = sin(pmu["ang_mean"][i]), cos(pmu["ang_mean"][i])
            st = sm * sa
            mean[r] = st * pmu["mag_mean"][i] * c
            mean[r + 1] = st * pmu["mag_mean"][i] * s
            var_re, var_im = _variance_pmu(pmu["mag_variance"][i], pmu["ang_variance"][i], pmu["mag_mean"][i], c, s)
            if pmu["correlated"][i]:
                correlated = True
                # covariancePmu + precision! (equations.jl:591-666)
                l1inv = 1 / sqrt(var_re)
                l2 = s * c * (pmu["mag_variance"][i] - pmu["ang_variance"][i] * pmu["mag_mean"][i] ** 2) * l1inv
                l3inv2 = 1 / (var_im - l2 ** 2)
                off = (-l2 * l1inv) * l3inv2
                pr = state.get("prow", 0)
                prow.extend((pr, pr + 1, pr, pr + 1)); pcol.extend((pr + 1, pr, pr, pr + 1))
                pval.extend((off, off, (l1inv - l2 * off) * l1inv, l3inv2))
                state["prow"] = pr + 2
            else:
                prec(var_re)
                prec(var_im)
            if pmu["bus"][i]:
                two(st, k, 16)
                two(st, k, 17)
            else:
                four(st, bool(pmu["frm"][i]), k, 18, 19)
                four(st, bool(pmu["frm"][i]), k, 20, 21)
    rng[5] = state.get("prow", 0)

    # sparse(row, col, val, m, 2n): CSC, rows sorted, explicit zeros kept
    H = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(total, 2 * n))
    H.sort_indices()
    h_colptr = H.indptr.astype(np.int64)
    h_rowval = H.indices.astype(np.int64)
    h_nzval = np.zeros(len(h_rowval))
    gn = GaussNewton(sys, mdl, sys.slack, sys.vm.copy(), sys.va.copy(), total, h_colptr, h_rowval, h_nzval,
                     None, mean, np.zeros(total), typ, idx, rng, np.zeros(2 * n), 0.0, 0, correlated, lu_options)
    for r, c, v in zip(rows, cols, vals):
        if v != 0.0:
            gn.h_nzval[gn.h_position(r, c)] = v
    W = sp.coo_matrix((pval, (prow, pcol)), shape=(total, total)).tocsc()
    W.sort_indices()
    gn.w = W
    return gn


# ----------------------------------------------------------------------------- per-row functions
def _coef(sys, mdl, k):
    y = mdl.admittance[k]
    return y.real, y.imag, 0.5 * sys.g[k], 0.5 * sys.b[k], 1 / sys.tap[k]


def normal_equation(a: GaussNewton):
    """normalEquation! (acStateEstimation.jl:261-583): residual, objective, H values."""
    sys, mdl = a.sys, a.mdl
    n = sys.n
    V, T = a.vm, a.va
    a.objective = 0.0
    wdiag = a.w.diagonal()
    H = a.h_nzval
    pos = a.h_position

    def obj(row):                                        # seobjective (equations.jl:689-698)
        a.objective += a.residual[row] ** 2 * wdiag[row]

    def obj2(row):
        a.objective += a.residual[row] ** 2 * wdiag[row] + \
            2 * a.residual[row] * a.residual[row - 1] * a.w[row, row - 1]

    for row in range(a.m):
        code = int(a.type[row])
        k = int(a.index[row])
        if code == 0:
            continue
        if code in (1, 12):
            a.residual[row] = a.mean[row] - V[k]
            obj(row)
            continue
        if code == 13:
            a.residual[row] = a.mean[row] - T[k]
            obj(row)
            continue
        if code in (6, 9):
            i = k
            cs_minus = 0.0      # sum Vj (G sin - B cos)
            cs_plus = 0.0       # sum Vj (G cos + B sin)
            Gii = Bii = 0.0
            for q in range(mdl.colptr[i], mdl.colptr[i + 1]):
                j = int(mdl.rowval[q])
                y = mdl.nzval_t[q]                       # Y[i,j]
                G, B = y.real, y.imag
                d = T[i] - T[j]
                s, c = sin(d), cos(d)
                cs_minus += V[j] * (G * s - B * c)
                cs_plus += V[j] * (G * c + B * s)
            yd = mdl.nzval[mdl.position(i, i)]
            Gii, Bii = yd.real, yd.imag
            if code == 6:
                a.residual[row] = a.mean[row] - V[i] * cs_plus
                obj(row)
                H[pos(row, i)] = V[i] * (-cs_minus) - Bii * (V[i] * V[i])        # Piθi
                H[pos(row, i + n)] = cs_plus + Gii * V[i]                    # PiVi
            else:
                a.residual[row] = a.mean[row] - V[i] * cs_minus
                obj(row)
                H[pos(row, i)] = V[i] * cs_plus - Gii * (V[i] * V[i])            # Qiθi
                H[pos(row, i + n)] = cs_minus - Bii * V[i]                   # QiVi
            for q in range(mdl.colptr[i], mdl.colptr[i + 1]):
                j = int(mdl.rowval[q])
                if j == i:
                    continue
                y = mdl.nzval[mdl.position(i, j)]        # ac.nodalMatrix[idx, col] (equations.jl:70-75)
                G, B = y.real, y.imag
                d = T[i] - T[j]
                s, c = sin(d), cos(d)
                if code == 6:
                    H[pos(row, j)] = V[i] * V[j] * (G * s - B * c)           # Piθj
                    H[pos(row, j + n)] = V[i] * (G * c + B * s)              # PiVj
                else:
                    H[pos(row, j)] = -V[i] * V[j] * (G * c + B * s)          # Qiθj
                    H[pos(row, j + n)] = V[i] * (G * s - B * c)              # QiVj
            continue
        if code in (16, 17):
            i = k
            if code == 16:
                a.residual[row] = a.mean[row] - V[i] * cos(T[i])
                obj(row)
                H[pos(row, i)] = -V[i] * sin(T[i])
                H[pos(row, i + n)] = cos(T[i])
            else:
                a.residual[row] = a.mean[row] - V[i] * sin(T[i])
                obj2(row)
                H[pos(row, i)] = V[i] * cos(T[i])
                H[pos(row, i + n)] = sin(T[i])
            continue

        # ---- branch rows
        i, j = int(sys.frm[k]), int(sys.to[k])
        g, b, gsi, bsi, tinv = _coef(sys, mdl, k)
        Vi, Vj = V[i], V[j]
        phi = sys.shift[k]
        d = T[i] - T[j] - phi                            # ViVjθijState (equations.jl:20-30)
        s, c = sin(d), cos(d)
        if code == 7:       # Pij (equations.jl:147-176)
            A, B_, C = (tinv * tinv) * (g + gsi), tinv * g, tinv * b
            h = A * (Vi * Vi) - (B_ * c + C * s) * Vi * Vj
            dti = (B_ * s - C * c) * Vi * Vj
            dvi = 2 * A * Vi - (B_ * c + C * s) * Vj
            dtj = -dti
            dvj = -(B_ * c + C * s) * Vi
        elif code == 8:     # Pji (:183-212)
            A, B_, C = g + gsi, tinv * g, tinv * b
            h = A * (Vj * Vj) - (B_ * c - C * s) * Vi * Vj
            dti = (B_ * s + C * c) * Vi * Vj
            dvi = (-B_ * c + C * s) * Vj
            dtj = -dti
            dvj = 2 * A * Vj - (B_ * c - C * s) * Vi
        elif code == 10:    # Qij (:215-244)
            A, B_, C = (tinv * tinv) * (b + bsi), tinv * g, tinv * b
            h = -A * (Vi * Vi) - (B_ * s - C * c) * Vi * Vj
            dti = -(B_ * c + C * s) * Vi * Vj
            dvi = -2 * A * Vi - (B_ * s - C * c) * Vj
            dtj = -dti
            dvj = -(B_ * s - C * c) * Vi
        elif code == 11:    # Qji (:247-276)
            A, B_, C = b + bsi, tinv * g, tinv * b
            h = -A * (Vj * Vj) + (B_ * s + C * c) * Vi * Vj
            dti = (B_ * c - C * s) * Vi * Vj
            dvi = (B_ * s + C * c) * Vj
            dtj = -dti
            dvj = -2 * A * Vj + (B_ * s + C * c) * Vi
        elif code in (2, 4, 14):   # Iij family (:279-331, 389-423)
            A = tinv ** 4 * ((g + gsi) ** 2 + (b + bsi) ** 2)
            B_ = (tinv * tinv) * (g ** 2 + b ** 2)
            C = (tinv * tinv * tinv) * (g * (g + gsi) + b * (b + bsi))
            D = (tinv * tinv * tinv) * (g * bsi - b * gsi)
            if code == 2:
                iinv = 1 / (sqrt(A * (Vi * Vi) + B_ * (Vj * Vj) - 2 * Vi * Vj * (C * c - D * s)))
                h = 1 / iinv
                dti = iinv * (C * s + D * c) * Vi * Vj
                dvi = iinv * (A * Vi - (C * c - D * s) * Vj)
                dtj = -dti
                dvj = iinv * (B_ * Vj - (C * c - D * s) * Vi)
            elif code == 4:
                h = A * (Vi * Vi) + B_ * (Vj * Vj) - 2 * Vi * Vj * (C * c - D * s)
                dti = 2 * (C * s + D * c) * Vi * Vj
                dvi = 2 * (A * Vi - (C * c - D * s) * Vj)
                dtj = -dti
                dvj = 2 * (B_ * Vj - (C * c - D * s) * Vi)
            else:
                # ψij: phasor from ψijCoefficient + ViVjθiθjState (:389-407), derivatives with Iij coefficients
                pA, pB = (tinv * tinv) * (g + gsi), (tinv * tinv) * (b + bsi)
                pC, pD = tinv * g, tinv * b
                si, ci = sin(T[i]), cos(T[i])
                sj, cj = sin(T[j] + phi), cos(T[j] + phi)
                re = (pA * ci - pB * si) * Vi - (pC * cj - pD * sj) * Vj
                im = (pA * si + pB * ci) * Vi - (pC * sj + pD * cj) * Vj
                iinv2 = 1 / (re * re + im * im)
                h = atan2(im, re)
                dti = iinv2 * (A * (Vi * Vi) - (C * c - D * s) * Vi * Vj)
                dvi = -iinv2 * (C * s + D * c) * Vj
                dtj = iinv2 * (B_ * (Vj * Vj) - (C * c - D * s) * Vi * Vj)
                dvj = iinv2 * (C * s + D * c) * Vi
        elif code in (3, 5, 15):   # Iji family (:334-386, 426-458)
            A = (tinv * tinv) * (g ** 2 + b ** 2)
            B_ = (g + gsi) ** 2 + (b + bsi) ** 2
            C = tinv * (g * (g + gsi) + b * (b + bsi))
            D = tinv * (g * bsi - gsi * b)
            if code == 3:
                iinv = 1 / sqrt(A * (Vi * Vi) + B_ * (Vj * Vj) - 2 * Vi * Vj * (C * c + D * s))
                h = 1 / iinv
                dti = iinv * (C * s - D * c) * Vi * Vj
                dvi = iinv * (A * Vi - (C * c + D * s) * Vj)
                dtj = -dti
                dvj = iinv * (B_ * Vj - (C * c + D * s) * Vi)
            elif code == 5:
                h = A * (Vi * Vi) + B_ * (Vj * Vj) - 2 * Vi * Vj * (C * c + D * s)
                dti = 2 * (C * s - D * c) * Vi * Vj
                dvi = 2 * (A * Vi - (C * c + D * s) * Vj)
                dtj = -dti
                dvj = 2 * (B_ * Vj - (C * c + D * s) * Vi)
            else:
                pA, pB = g + gsi, b + bsi                 # ψjiCoefficient (:426-436)
                pC, pD = tinv * g, tinv * b
                si, ci = sin(T[i] - phi), cos(T[i] - phi)  # VjViθjθiState (:47-60)
                sj, cj = sin(T[j]), cos(T[j])
                re = (pA * cj - pB * sj) * Vj - (pC * ci - pD * si) * Vi
                im = (pA * sj + pB * cj) * Vj - (pC * si + pD * ci) * Vi
                iinv2 = 1 / (re * re + im * im)
                h = atan2(im, re)
                dti = iinv2 * (A * (Vi * Vi) - (C * c + D * s) * Vi * Vj)
                dvi = -iinv2 * (C * s - D * c) * Vj
                dtj = iinv2 * (B_ * (Vj * Vj) - (C * c + D * s) * Vi * Vj)
                dvj = iinv2 * (C * s - D * c) * Vi
        elif code in (18, 20):     # Re/Im Iij (:466-505)
            pA, pB = (tinv * tinv) * (g + gsi), (tinv * tinv) * (b + bsi)
            pC, pD = tinv * g, tinv * b
            si, ci = sin(T[i]), cos(T[i])
            sj, cj = sin(T[j] + phi), cos(T[j] + phi)
            if code == 18:
                h = (pA * ci - pB * si) * Vi - (pC * cj - pD * sj) * Vj
                dti = -(pA * si + pB * ci) * Vi
                dvi = pA * ci - pB * si
                dtj = (pC * sj + pD * cj) * Vj
                dvj = -pC * cj + pD * sj
            else:
                h = (pA * si + pB * ci) * Vi - (pC * sj + pD * cj) * Vj
                dti = (pA * ci - pB * si) * Vi
                dvi = pA * si + pB * ci
                dtj = (-pC * cj + pD * sj) * Vj
                dvj = -pC * sj - pD * cj
        elif code in (19, 21):     # Re/Im Iji (:508-547)
            pA, pB = g + gsi, b + bsi
            pC, pD = tinv * g, tinv * b
            si, ci = sin(T[i] - phi), cos(T[i] - phi)
            sj, cj = sin(T[j]), cos(T[j])
            if code == 19:
                h = (pA * cj - pB * sj) * Vj - (pC * ci - pD * si) * Vi
                dti = (pC * si + pD * ci) * Vi
                dvi = -pC * ci + pD * si
                dtj = -(pA * sj + pB * cj) * Vj
                dvj = pA * cj - pB * sj
            else:
                h = (pA * sj + pB * cj) * Vj - (pC * si + pD * ci) * Vi
                dti = (-pC * ci + pD * si) * Vi
                dvi = -pC * si - pD * ci
                dtj = (pA * cj - pB * sj) * Vj
                dvj = pA * sj + pB * cj
        else:
            raise ValueError(f"unknown measurement code {code}")

        a.residual[row] = a.mean[row] - h
        if code in (20, 21):
            obj2(row)
        else:
            obj(row)
        H[pos(row, i)] = dti
        H[pos(row, i + n)] = dvi
        H[pos(row, j)] = dtj
        H[pos(row, j + n)] = dvj


def increment(a: GaussNewton) -> float:
    """increment!(analysis) for the normal-equation methods (acStateEstimation.jl:878-904)."""
    n = a.sys.n
    normal_equation(a)
    lo, hi = a.h_colptr[a.slack], a.h_colptr[a.slack + 1]
    saved = a.h_nzval[lo:hi].copy()                     # removeColumn (sparse.jl:155-163)
    a.h_nzval[lo:hi] = 0.0
    H = sp.csc_matrix((a.h_nzval, a.h_rowval, a.h_colptr), shape=(a.m, 2 * n))
    temp = (H.T @ a.w).tocsc()
    gain = (temp @ H).tolil()
    gain[a.slack, a.slack] = 1.0
    gain = gain.tocsc()
    a.gain = gain
    opts = a.lu_options or {}
    lu = spla.splu(gain, **opts)
    a.increment[:] = lu.solve(temp @ a.residual)
    a.increment[a.slack] = 0.0
    a.h_nzval[lo:hi] = saved                             # restoreColumn! (sparse.jl:177-188)
    return float(np.max(np.abs(a.increment)))


def solve(a: GaussNewton):
    """solve!(analysis) (acStateEstimation.jl:1035-1047)."""
    n = a.sys.n
    a.va += a.increment[:n]
    a.vm += a.increment[n:]
    a.iteration += 1


def state_estimation(a: GaussNewton, iteration: int = 40, tolerance: float = 1e-8, trace: list | None = None):
    """stateEstimation!(analysis) loop (acStateEstimation.jl:1286-1329)."""
    a.iteration = 0
    converged = False
    for _ in range(iteration + 1):
        max_inc = increment(a)
        if trace is not None:
            trace.append((max_inc, a.objective))
        if max_inc < tolerance:
            converged = True
            break
        if a.iteration == iteration:
            break
        solve(a)
    return converged


def gain_pattern(a: GaussNewton):
    """Structural pattern of G = H'WH (slack row/col kept, diagonal forced): what Julia's SpGEMM stores."""
    n = a.sys.n
    Hp = sp.csc_matrix((np.ones(len(a.h_rowval)), a.h_rowval, a.h_colptr), shape=(a.m, 2 * n))
    Wp = a.w.copy()
    Wp.data[:] = 1.0
    G = (Hp.T @ Wp @ Hp).tocsc()
    G.sort_indices()
    return G.indptr.astype(np.int64), G.indices.astype(np.int64)


def export_one_based(a: GaussNewton) -> dict:
    return {
        "h_colptr": (a.h_colptr + 1).astype(np.int64),
        "h_rowval": (a.h_rowval + 1).astype(np.int64),
        "type": a.type.astype(np.int8),
        "index": (a.index + 1).astype(np.int64),
        "range": (a.range + 1).astype(np.int64),
    }


# ----------------------------------------------------------------------------- bad data
def chi_test(a: GaussNewton, confidence: float = 0.95):
    """chiTest (stateEstimation/badData.jl:1-46 of the chi-square part): objective against the chi2 quantile with
    m_in_service - (2n - 1) degrees of freedom."""
    import scipy.stats
    inservice = int(np.count_nonzero(a.type))
    threshold = scipy.stats.chi2.ppf(confidence, inservice - (2 * a.sys.n - 1))
    return a.objective > threshold, threshold, a.objective


def residual_projection(a: GaussNewton) -> np.ndarray:
    """c[i] = h_i G^-1 h_i' with the slack column of H removed and G[slack, slack] = 1
    (residualTest!, badData.jl:198-203; badDataProjection / rowProjection, :287-362). The reference reads the entries
    of G^-1 it needs from a sparse selected inverse of the gain factor; here they come from SuperLU solves."""
    n = a.sys.n
    H = a.jacobian().tolil()
    H[:, a.slack] = 0.0
    H = H.tocsc()
    G = (H.T @ a.w @ H).tolil()
    G[a.slack, a.slack] = 1.0
    lu = spla.splu(G.tocsc())
    X = lu.solve(H.T.toarray())                 # G^-1 H'   (2n x m), dense: small / medium cases only
    return np.asarray(H.multiply(X.T).sum(axis=1)).ravel()


def residual_projection_rows(a: GaussNewton, rows) -> np.ndarray:
    """Same quantity for a sample of rows (large cases): one SuperLU solve per row."""
    H = a.jacobian().tolil()
    H[:, a.slack] = 0.0
    H = H.tocsr()
    G = (H.T @ a.w @ H).tolil()
    G[a.slack, a.slack] = 1.0
    lu = spla.splu(G.tocsc())
    out = np.zeros(len(rows))
    for q, r in enumerate(rows):
        h = H.getrow(int(r)).toarray().ravel()
        out[q] = h @ lu.solve(h)
    return out


def residual_test(a: GaussNewton, threshold: float = 3.0):
    """residualTest! for Gauss-Newton WLS (badData.jl:181-285). Returns (detect, max normalised residual, row index
    or -1); on detection the row (and its partner for a rectangular PMU) is taken out of service exactly like the
    reference: H row, mean and residual zeroed, type 0, iteration 0."""
    c = residual_projection(a)
    wdiag = a.w.diagonal()
    best, index = 0.0, -1
    for i in range(a.m):
        if a.residual[i] != 0.0:
            rn = abs(a.residual[i]) / sqrt(abs(1 / wdiag[i] - c[i]))
            if rn > best:
                best, index = rn, i
    detect = best > threshold
    if detect:
        rows = [index]
        if index >= a.range[4]:                               # PMU block: rectangular pairs go together
            code = int(a.type[index])
            if code in (16, 17, 18, 19, 20, 21):
                local = index - a.range[4]
                rows.append(index - 1 if local % 2 == 1 else index + 1)
        for r in rows:
            for col in range(2 * a.sys.n):
                lo, hi = a.h_colptr[col], a.h_colptr[col + 1]
                k = lo + int(np.searchsorted(a.h_rowval[lo:hi], r))
                if k < hi and a.h_rowval[k] == r:
                    a.h_nzval[k] = 0.0
            a.mean[r] = 0.0
            a.residual[r] = 0.0
            a.type[r] = 0
        a.iteration = 0
    return detect, best, index
