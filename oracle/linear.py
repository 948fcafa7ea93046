"""Oracle restatement of the reference's linear analyses (TEST INFRASTRUCTURE — never imported by the product).

Reference lines followed (all under /root/reference/src):
  dcModel!                        powerSystem/model.jl:161-212   (CSC builder: backend/sparse.jl:2-101)
  dcPowerFlow / solve!            powerFlow/dcPowerFlow.jl:43-134 (removeRowColumn + unit diagonal, sparse.jl:165-202;
                                  addSlackAngle!, backend/utility.jl:610-622)
  power!(::DcPowerFlow)           postprocessing/dcAnalysis.jl:27-76, 353-390
  dcStateEstimationWls / solve!   stateEstimation/dcStateEstimation.jl:41-140, 342-371 (meanPi/meanPij/meanθi,
                                  backend/equations.jl:121-123, 178-180, 461-463)
  pmuEstimationWls / solve!       stateEstimation/pmuStateEstimation.jl:66-166, 369-399
                                  (ReImIijCoefficient / ReImIjiCoefficient, backend/expressions.jl:291-302, 338-349;
                                  variancePmu / covariancePmu / precision!, backend/equations.jl:576-677)
  factorisation + solution!       Julia SparseArrays / SuiteSparse (third party); here SciPy SuperLU.
Pinned by the reference's own goldens: results.h5:/case14test|case30test/dcPowerFlow (tests/golden/*.json) and the
recovery tests of test/stateEstimation/analysis.jl:350-440, 455-560 (estimate == power-flow state).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import sin, cos, sqrt
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .system import System
from .model import ac_model
from .wls import Measurement, _variance_pmu


@dataclass
class DcModel:
    n: int
    colptr: np.ndarray       # 0-based CSC of dc.nodalMatrix, sorted rows, explicit zeros kept
    rowval: np.ndarray
    nzval: np.ndarray
    admittance: np.ndarray   # dc.admittance (0 when out of service)
    shift_power: np.ndarray  # dc.shiftPower

    def matrix(self) -> sp.csc_matrix:
        return sp.csc_matrix((self.nzval.copy(), self.rowval, self.colptr), shape=(self.n, self.n))


def dc_model(sys: System) -> DcModel:
    n, m = sys.n, sys.nbr
    adm = np.zeros(m)
    shift_power = np.zeros(n)
    for i in range(m):
        if sys.status[i] == 1:
            adm[i] = 1 / (sys.tap[i] * sys.x[i])
            shift = sys.shift[i] * adm[i]
            shift_power[sys.frm[i]] -= shift
            shift_power[sys.to[i]] += shift
    cols_rows = [[i] for i in range(n)]
    cols_vals = [[0.0] for _ in range(n)]
    for i in range(m):
        f, t = int(sys.frm[i]), int(sys.to[i])
        a = 0.0
        if sys.status[i] == 1:
            a = adm[i]
            cols_vals[f][0] += a
            cols_vals[t][0] += a
        cols_rows[t].append(f); cols_vals[t].append(-a)     # addEntry!(builder, from, to, -admittance)
        cols_rows[f].append(t); cols_vals[f].append(-a)     # addEntry!(builder, to, from, -admittance)
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval, nzval = [], []
    for c in range(n):
        rows = np.array(cols_rows[c])
        prev = None
        for k in np.argsort(rows, kind="stable"):           # canonicalize!: stable sort, duplicates summed in order
            r = int(rows[k])
            if prev is not None and r == prev:
                nzval[-1] += cols_vals[c][k]
            else:
                rowval.append(r); nzval.append(cols_vals[c][k]); prev = r
        colptr[c + 1] = len(rowval)
    return DcModel(n, colptr, np.array(rowval, dtype=np.int64), np.array(nzval), adm, shift_power)


def _supply(sys: System):
    sp_ = np.zeros(sys.n)
    gens = {}
    for g in range(sys.ngen):
        if sys.gen_status[g] == 1:
            b = int(sys.gen_bus[g])
            sp_[b] += sys.gen_p[g]
            gens.setdefault(b, []).append(g)
    return sp_, gens


def dc_rhs(sys: System, dc: DcModel, supply=None) -> np.ndarray:
    """pf.rhs (dcPowerFlow.jl:99-102)."""
    if supply is None:
        supply, _ = _supply(sys)
    return supply - sys.pd - sys.gs - dc.shift_power


def slack_fixed(dc: DcModel, slack: int) -> sp.csc_matrix:
    """removeRowColumn + nodalMatrix[slack, slack] = 1 (dcPowerFlow.jl:113-114)."""
    B = dc.matrix().tolil()
    B[slack, :] = 0.0
    B[:, slack] = 0.0
    B[slack, slack] = 1.0
    return B.tocsc()


def add_slack_angle(sys: System, angle: np.ndarray) -> np.ndarray:
    angle = angle.copy()
    angle[sys.slack] = 0.0
    if sys.va[sys.slack] != 0.0:
        angle += sys.va[sys.slack]
    return angle


def dc_power_flow(sys: System, dc: DcModel | None = None, rhs: np.ndarray | None = None) -> np.ndarray:
    dc = dc or dc_model(sys)
    b = dc_rhs(sys, dc) if rhs is None else rhs
    theta = spla.splu(slack_fixed(dc, sys.slack)).solve(b)
    return add_slack_angle(sys, theta)


def dc_injection(sys: System, dc: DcModel, angle: np.ndarray, i: int) -> float:
    p = 0.0
    for j in range(dc.colptr[i], dc.colptr[i + 1]):
        p += dc.nzval[j] * angle[dc.rowval[j]]
    return p + sys.gs[i] + dc.shift_power[i]


def dc_power(sys: System, dc: DcModel, angle: np.ndarray) -> dict:
    supply, gens = _supply(sys)
    inj = supply - sys.pd
    s = sys.slack
    inj[s] = dc_injection(sys, dc, angle, s)
    sup = supply.copy()
    sup[s] = sys.pd[s] + inj[s]
    gen = np.zeros(sys.ngen)
    for g in range(sys.ngen):
        if sys.gen_status[g] == 1:
            b = int(sys.gen_bus[g])
            if b == s and gens[b][0] == g:
                gen[g] = dc_injection(sys, dc, angle, b) + sys.pd[b]
                for o in gens[b][1:]:
                    gen[g] -= sys.gen_p[o]
            else:
                gen[g] = sys.gen_p[g]
    frm = dc.admittance * (angle[sys.frm] - angle[sys.to] - sys.shift)
    return {"injection": inj, "supply": sup, "generator": gen, "from": frm, "to": -frm}


# ----------------------------------------------------------------------------- DC state estimation
@dataclass
class LinearWls:
    coefficient: sp.csc_matrix      # H
    precision: sp.csc_matrix        # W
    mean: np.ndarray                # z
    slack: int                      # -1 when no column is removed


def dc_wls(sys: System, meas: Measurement, dc: DcModel | None = None) -> LinearWls:
    """dcStateEstimationWls: wattmeter rows first, then the angle rows of bus PMUs."""
    dc = dc or dc_model(sys)
    watt, pmu = meas.watt, meas.pmu
    nw = len(watt["index"])
    pmu_rows = [i for i in range(len(pmu["index"])) if pmu["bus"][i]]
    total = nw + len(pmu_rows)
    mean = np.zeros(total)
    prec = np.zeros(total)
    rows, cols, vals = [], [], []
    for i in range(nw):
        k = int(watt["index"][i]); st = int(watt["status"][i])
        prec[i] = 1 / watt["variance"][i]
        if watt["bus"][i]:
            mean[i] = st * (watt["mean"][i] - dc.shift_power[k] - sys.gs[k])            # meanPi
            for j in range(dc.colptr[k], dc.colptr[k + 1]):
                rows.append(i); cols.append(int(dc.rowval[j])); vals.append(st * dc.nzval[j])
        else:
            a = st * dc.admittance[k] if watt["frm"][i] else -st * dc.admittance[k]
            mean[i] = st * (watt["mean"][i] + sys.shift[k] * a)                         # meanPij
            rows += [i, i]; cols += [int(sys.frm[k]), int(sys.to[k])]; vals += [a, -a]
    for q, i in enumerate(pmu_rows):
        r = nw + q
        st = int(pmu["ang_status"][i])
        mean[r] = st * (pmu["ang_mean"][i] - sys.va[sys.slack])                         # meanθi
        prec[r] = 1 / pmu["ang_variance"][i]
        rows.append(r); cols.append(int(pmu["index"][i])); vals.append(float(st))
    H = sp.coo_matrix((vals, (rows, cols)), shape=(total, sys.n)).tocsc()
    H.sort_indices()
    return LinearWls(H, sp.diags(prec).tocsc(), mean, sys.slack)


def wls_matrices(w: LinearWls):
    """What `solve!` forms: H with the slack column removed, G = H'WH with G[slack, slack] = 1, P = W H."""
    H = w.coefficient.copy().tolil()
    if w.slack >= 0:
        H[:, w.slack] = 0.0
    H = H.tocsc()
    P = (w.precision @ H).tocsc()
    G = (H.T @ P).tolil()
    if w.slack >= 0:
        G[w.slack, w.slack] = 1.0
    return H, G.tocsc(), P


def wls_solve(w: LinearWls, mean: np.ndarray | None = None) -> np.ndarray:
    H, G, P = wls_matrices(w)
    z = w.mean if mean is None else mean
    return spla.splu(G).solve(P.T @ z)


def dc_state_estimation(sys: System, meas: Measurement, dc: DcModel | None = None) -> np.ndarray:
    w = dc_wls(sys, meas, dc)
    return add_slack_angle(sys, wls_solve(w))


# ----------------------------------------------------------------------------- PMU state estimation
def _reim_from(sys: System, adm, k):
    g, b = adm[k].real, adm[k].imag
    tinv = 1 / sys.tap[k]
    phi = sys.shift[k]
    return (tinv * tinv * (g + 0.5 * sys.g[k]), -(tinv * tinv) * (b + 0.5 * sys.b[k]),
            -tinv * (g * cos(phi) - b * sin(phi)), tinv * (b * cos(phi) + g * sin(phi)))


def _reim_to(sys: System, adm, k):
    g, b = adm[k].real, adm[k].imag
    tinv = 1 / sys.tap[k]
    phi = sys.shift[k]
    return (-tinv * (g * cos(phi) + b * sin(phi)), tinv * (b * cos(phi) - g * sin(phi)),
            g + 0.5 * sys.g[k], -b - 0.5 * sys.b[k])


def pmu_wls(sys: System, meas: Measurement, mdl=None) -> LinearWls:
    """pmuEstimationWls: two rows (Re, Im) per PMU; state = [Re V; Im V]."""
    mdl = mdl or ac_model(sys)
    pmu = meas.pmu
    n, npmu = sys.n, len(pmu["index"])
    mean = np.zeros(2 * npmu)
    rows, cols, vals = [], [], []
    prow, pcol, pval = [], [], []
    for i in range(npmu):
        k = int(pmu["index"][i])
        r = 2 * i
        s, c = sin(pmu["ang_mean"][i]), cos(pmu["ang_mean"][i])
        var_re, var_im = _variance_pmu(pmu["mag_variance"][i], pmu["ang_variance"][i], pmu["mag_mean"][i], c, s)
        if pmu["correlated"][i]:
            l1inv = 1 / sqrt(var_re)
            l2 = s * c * (pmu["mag_variance"][i] - pmu["ang_variance"][i] * pmu["mag_mean"][i] ** 2) * l1inv
            l3inv2 = 1 / (var_im - l2 ** 2)
            off = (-l2 * l1inv) * l3inv2
            prow += [r, r + 1, r, r + 1]; pcol += [r + 1, r, r, r + 1]
            pval += [off, off, (l1inv - l2 * off) * l1inv, l3inv2]
        else:
            prow += [r, r + 1]; pcol += [r, r + 1]; pval += [1 / var_re, 1 / var_im]
        on = pmu["mag_status"][i] == 1 and pmu["ang_status"][i] == 1
        if on:
            mean[r] = pmu["mag_mean"][i] * c
            mean[r + 1] = pmu["mag_mean"][i] * s
        if pmu["bus"][i]:
            v = 1.0 if on else 0.0
            rows += [r, r + 1]; cols += [k, k + n]; vals += [v, v]
        else:
            A, B, C, D = (_reim_from(sys, mdl.admittance, k) if pmu["frm"][i] else _reim_to(sys, mdl.admittance, k)) \
                if on else (0.0, 0.0, 0.0, 0.0)
            f, t = int(sys.frm[k]), int(sys.to[k])
            rows += [r, r + 1, r, r + 1, r, r + 1, r, r + 1]
            cols += [f, f + n, t, t + n, f + n, f, t + n, t]
            vals += [A, A, C, C, B, -B, D, -D]
    H = sp.coo_matrix((vals, (rows, cols)), shape=(2 * npmu, 2 * n)).tocsc()
    H.sort_indices()
    W = sp.coo_matrix((pval, (prow, pcol)), shape=(2 * npmu, 2 * npmu)).tocsc()
    return LinearWls(H, W, mean, -1)


def pmu_state_estimation(sys: System, meas: Measurement, mdl=None):
    w = pmu_wls(sys, meas, mdl)
    x = wls_solve(w)
    v = x[:sys.n] + 1j * x[sys.n:]
    return np.abs(v), np.angle(v)
