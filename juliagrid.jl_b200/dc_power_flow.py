"""Host mirror of the reference's DC power flow, backed by libjgb200.so.

    dc_model(system)          <-> dcModel!(system)            src/powerSystem/model.jl:161-212
    dc_power_flow(system)     <-> dcPowerFlow(system, B200)   src/powerFlow/dcPowerFlow.jl:43-70
    solve_dc(analysis)        <-> solve!(analysis)            :93-134
    power_dc(analysis)        <-> power!(analysis)            src/postprocessing/dcAnalysis.jl:27-76, 353-390
    dc_batch(analysis, P)     <-> the user loop `updateBus!(active = ...)` + `solve!` per injection scenario: the nodal
                                  matrix is constant, so it is factored once and every scenario is one right-hand side
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from ._lib import Context
from .cases import PowerSystem
from .linear_solver import LinearSolver


@dataclass
class DcModel:
    nodal: sp.csc_matrix       # dc.nodalMatrix (explicit zeros of out-of-service branches kept)
    admittance: np.ndarray     # dc.admittance
    shift_power: np.ndarray    # dc.shiftPower


def dc_model(system: PowerSystem) -> DcModel:
    n, m = system.n, system.nbr
    on = system.status == 1
    adm = np.zeros(m)
    adm[on] = 1.0 / (system.tap[on] * system.x[on])
    shift = system.shift * adm
    shift_power = np.zeros(n)
    np.subtract.at(shift_power, system.frm, shift)
    np.add.at(shift_power, system.to, shift)
    rows = np.concatenate([np.arange(n), system.frm, system.to, system.frm, system.to])
    cols = np.concatenate([np.arange(n), system.frm, system.to, system.to, system.frm])
    vals = np.concatenate([np.zeros(n), adm, adm, -adm, -adm])
    nodal = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()    # duplicates summed, zeros kept
    nodal.sort_indices()
    return DcModel(nodal, adm, shift_power)


class DcPowerFlow:
    def __init__(self, system: PowerSystem, dc: DcModel, solver: LinearSolver):
        self.system, self.dc, self.solver = system, dc, solver
        self.angle = None
        self.power = None

    def rhs(self, supply=None, demand=None) -> np.ndarray:
        """pf.rhs (dcPowerFlow.jl:99-102); supply / demand may be [R][n] blocks."""
        s = self.system
        sup = s.supply[0] if supply is None else np.asarray(supply)
        dem = s.pd if demand is None else np.asarray(demand)
        return sup - dem - s.gs - self.dc.shift_power


def dc_power_flow(system: PowerSystem, ctx: Context | None = None, device: int = 0) -> DcPowerFlow:
    if not np.any((system.gen_status == 1) & (system.gen_bus == system.slack)):
        raise ValueError("the slack bus has no in-service generator (the reference would move the slack bus)")
    dc = dc_model(system)
    return DcPowerFlow(system, dc, LinearSolver(dc.nodal, skip=system.slack, ctx=ctx, device=device))


def _add_slack_angle(system: PowerSystem, angle: np.ndarray) -> np.ndarray:
    angle[..., system.slack] = 0.0
    if system.va[system.slack] != 0.0:
        angle += system.va[system.slack]
    return angle


def solve_dc(a: DcPowerFlow) -> np.ndarray:
    a.angle = _add_slack_angle(a.system, a.solver.solve(a.rhs()))
    return a.angle


def dc_batch(a: DcPowerFlow, supply=None, demand=None) -> np.ndarray:
    """Angles [R][n] for R injection scenarios (rows of `supply` and / or `demand`)."""
    return _add_slack_angle(a.system, a.solver.solve(np.atleast_2d(a.rhs(supply, demand))))


def power_dc(a: DcPowerFlow) -> dict:
    s, dc, th = a.system, a.dc, a.angle
    supply, _, first = s.supply
    slack = s.slack
    inj = supply - s.pd
    col = dc.nodal.getcol(slack)
    p_slack = float(col.data @ th[col.indices]) + s.gs[slack] + dc.shift_power[slack]
    inj[slack] = p_slack
    sup = supply.copy()
    sup[slack] = s.pd[slack] + p_slack
    gen = np.where(s.gen_status == 1, s.gen_p, 0.0)
    g0 = first[slack]
    others = (s.gen_status == 1) & (s.gen_bus == slack)
    others[g0] = False
    gen[g0] = p_slack + s.pd[slack] - s.gen_p[others].sum()
    frm = dc.admittance * (th[s.frm] - th[s.to] - s.shift)
    a.power = {"injection": inj, "supply": sup, "generator": gen, "from": frm, "to": -frm}
    return a.power
