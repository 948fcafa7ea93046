"""Host mirror of the reference's DC (linear WLS) state estimation, backed by libjgb200.so.

    dc_state_estimation(monitoring)  <-> dcStateEstimation(monitoring, B200)  src/stateEstimation/dcStateEstimation.jl:41-140
    solve_dc_se(analysis)            <-> solve!(analysis)                     :342-371
    dc_se_batch(analysis, Z)         <-> the user loop `updateWattmeter!(active = ...)` + `solve!` per Monte-Carlo draw:
                                         H and W do not change, so the gain matrix is factored once
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from ._lib import Context
from .dc_power_flow import DcModel, dc_model, _add_slack_angle
from .linear_solver import LinearSolver
from .measurement import Measurement


class LinearWls:
    """analysis.method of the linear estimators: coefficient (H), precision (W), mean (z) and the device solver."""

    def __init__(self, coefficient, precision, mean, slack, ctx=None, device=0):
        self.coefficient, self.precision, self.mean, self.slack = coefficient, precision, mean, slack
        h = coefficient.tocsc(copy=True)
        if slack >= 0:                                   # removeColumn (sparse.jl:155-163)
            h.data[h.indptr[slack]:h.indptr[slack + 1]] = 0.0
        wh = (precision @ h).tocsc()
        gain = (h.T @ wh).tolil()
        if slack >= 0:
            gain[slack, slack] = 1.0                     # dcStateEstimation.jl:354
        gain = gain.tocsc()
        gain = ((gain + gain.T) * 0.5).tocsc()           # exact symmetry of the stored values
        self.solver = LinearSolver(gain, skip=slack, ctx=ctx, device=device)
        self.solver.set_projection(wh)

    def solve(self, mean=None) -> np.ndarray:
        return self.solver.solve_projected(self.mean if mean is None else mean)


class DcStateEstimation:
    def __init__(self, monitoring: Measurement, dc: DcModel, method: LinearWls):
        self.monitoring, self.system, self.dc, self.method = monitoring, monitoring.system, dc, method
        self.angle = None


def dc_wls_tables(monitoring: Measurement, dc: DcModel):
    """dcStateEstimationWls: rows = wattmeters, then the angle of every bus PMU."""
    s = monitoring.system
    watt, pmu = monitoring.watt, monitoring.pmu
    nw = len(watt["index"])
    k = watt["index"]
    st = watt["status"].astype(float)
    bus = watt["bus"]
    ib = np.flatnonzero(bus)
    nodal = dc.nodal
    cnt = np.diff(nodal.indptr)[k[ib]]
    rows_b = np.repeat(ib, cnt)
    start = nodal.indptr[k[ib]]
    pos = np.repeat(start, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
    cols_b = nodal.indices[pos]
    vals_b = np.repeat(st[ib], cnt) * nodal.data[pos]
    ibr = np.flatnonzero(~bus)
    a = np.where(watt["frm"][ibr], 1.0, -1.0) * st[ibr] * dc.admittance[k[ibr]]
    mean = np.zeros(nw)
    mean[ib] = st[ib] * (watt["mean"][ib] - dc.shift_power[k[ib]] - s.gs[k[ib]])
    mean[ibr] = st[ibr] * (watt["mean"][ibr] + s.shift[k[ibr]] * a)
    pb = np.flatnonzero(pmu["bus"])
    pst = pmu["ang_status"][pb].astype(float)
    total = nw + len(pb)
    rows = np.concatenate([rows_b, ibr, ibr, nw + np.arange(len(pb))])
    cols = np.concatenate([cols_b, s.frm[k[ibr]], s.to[k[ibr]], pmu["index"][pb]])
    vals = np.concatenate([vals_b, a, -a, pst])
    h = sp.coo_matrix((vals, (rows, cols)), shape=(total, s.n)).tocsc()
    h.sort_indices()
    prec = np.concatenate([1.0 / watt["variance"], 1.0 / pmu["ang_variance"][pb]])
    mean = np.concatenate([mean, pst * (pmu["ang_mean"][pb] - s.va[s.slack])])
    return h, sp.diags(prec).tocsc(), mean


def dc_state_estimation(monitoring: Measurement, ctx: Context | None = None, device: int = 0) -> DcStateEstimation:
    dc = dc_model(monitoring.system)
    h, w, z = dc_wls_tables(monitoring, dc)
    return DcStateEstimation(monitoring, dc, LinearWls(h, w, z, monitoring.system.slack, ctx, device))


def solve_dc_se(a: DcStateEstimation) -> np.ndarray:
    a.angle = _add_slack_angle(a.system, a.method.solve())
    return a.angle


def dc_se_batch(a: DcStateEstimation, Z) -> np.ndarray:
    """Z [R][m]: one mean vector per draw (already in the estimator's row order and offsets); angles [R][n]."""
    return _add_slack_angle(a.system, a.method.solve(np.atleast_2d(Z)))
