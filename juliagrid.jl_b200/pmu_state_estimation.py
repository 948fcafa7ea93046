"""Host mirror of the reference's PMU-only (linear WLS, rectangular state) state estimation, backed by libjgb200.so.

    pmu_state_estimation(monitoring)  <-> pmuStateEstimation(monitoring, B200)  src/stateEstimation/pmuStateEstimation.jl:36-166
    solve_pmu_se(analysis)            <-> solve!(analysis)                      :369-399
    pmu_se_batch(analysis, Z)         <-> `updatePmu!(...)` + `solve!` per Monte-Carlo draw on a fixed H and W
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from ._lib import Context
from .ac_power_flow import Polar
from .dc_state_estimation import LinearWls
from .measurement import Measurement
from .model import ac_model


class PmuStateEstimation:
    def __init__(self, monitoring: Measurement, method: LinearWls):
        self.monitoring, self.system, self.method = monitoring, monitoring.system, method
        self.voltage: Polar = None


def pmu_wls_tables(monitoring: Measurement):
    """pmuEstimationWls: rows 2i, 2i+1 = Re, Im of PMU i; columns = [Re V | Im V]."""
    s = monitoring.system
    mdl = s.model or ac_model(s)
    p = monitoring.pmu
    n, npmu = s.n, len(p["index"])
    k = p["index"]
    sn, cs = np.sin(p["ang_mean"]), np.cos(p["ang_mean"])
    mag, vm_, va_ = p["mag_mean"], p["mag_variance"], p["ang_variance"]
    var_re = vm_ * cs * cs + va_ * (mag * sn) * (mag * sn)            # variancePmu (equations.jl:576-588)
    var_im = vm_ * sn * sn + va_ * (mag * cs) * (mag * cs)
    on = ((p["mag_status"] == 1) & (p["ang_status"] == 1)).astype(float)
    mean = np.empty(2 * npmu)
    mean[0::2], mean[1::2] = on * mag * cs, on * mag * sn
    r = 2 * np.arange(npmu)
    corr = p["correlated"]
    l1inv = 1.0 / np.sqrt(var_re)                                     # covariancePmu / precision! (:591-666)
    l2 = sn * cs * (vm_ - va_ * mag * mag) * l1inv
    l3inv2 = 1.0 / (var_im - l2 * l2)
    off = (-l2 * l1inv) * l3inv2
    d0 = np.where(corr, (l1inv - l2 * off) * l1inv, 1.0 / var_re)
    d1 = np.where(corr, l3inv2, 1.0 / var_im)
    rc = r[corr]
    w = sp.coo_matrix((np.concatenate([d0, d1, off[corr], off[corr]]),
                       (np.concatenate([r, r + 1, rc, rc + 1]), np.concatenate([r, r + 1, rc + 1, rc]))),
                      shape=(2 * npmu, 2 * npmu)).tocsc()
    ib = np.flatnonzero(p["bus"])
    ibr = np.flatnonzero(~p["bus"])
    kb = k[ibr]
    g, b = mdl.admittance[kb].real, mdl.admittance[kb].imag
    tinv = 1.0 / s.tap[kb]
    cphi, sphi = np.cos(s.shift[kb]), np.sin(s.shift[kb])
    frm = p["frm"][ibr]
    # ReImIijCoefficient / ReImIjiCoefficient (backend/expressions.jl:291-302, 338-349)
    A = np.where(frm, tinv * tinv * (g + 0.5 * s.g[kb]), -tinv * (g * cphi + b * sphi))
    B = np.where(frm, -(tinv * tinv) * (b + 0.5 * s.b[kb]), tinv * (b * cphi - g * sphi))
    Cc = np.where(frm, -tinv * (g * cphi - b * sphi), g + 0.5 * s.g[kb])
    D = np.where(frm, tinv * (b * cphi + g * sphi), -b - 0.5 * s.b[kb])
    o = on[ibr]
    A, B, Cc, D = o * A, o * B, o * Cc, o * D
    f, t = s.frm[kb], s.to[kb]
    rr = r[ibr]
    rows = np.concatenate([r[ib], r[ib] + 1, rr, rr + 1, rr, rr + 1, rr, rr + 1, rr, rr + 1])
    cols = np.concatenate([k[ib], k[ib] + n, f, f + n, t, t + n, f + n, f, t + n, t])
    vals = np.concatenate([on[ib], on[ib], A, A, Cc, Cc, B, -B, D, -D])
    h = sp.coo_matrix((vals, (rows, cols)), shape=(2 * npmu, 2 * n)).tocsc()
    h.sort_indices()
    return h, w, mean


def pmu_state_estimation(monitoring: Measurement, ctx: Context | None = None, device: int = 0) -> PmuStateEstimation:
    h, w, z = pmu_wls_tables(monitoring)
    return PmuStateEstimation(monitoring, LinearWls(h, w, z, -1, ctx, device))


def _polar(x, n) -> Polar:
    v = x[..., :n] + 1j * x[..., n:]
    return Polar(np.abs(v), np.angle(v))


def solve_pmu_se(a: PmuStateEstimation) -> Polar:
    a.voltage = _polar(a.method.solve(), a.system.n)
    return a.voltage


def pmu_se_batch(a: PmuStateEstimation, Z) -> Polar:
    """Z [R][2*npmu]: rectangular means (Re, Im interleaved per PMU) of every draw; magnitudes / angles [R][n]."""
    return _polar(a.method.solve(np.atleast_2d(Z)), a.system.n)
