"""Oracle restatement of the Newton-Raphson AC power flow (TEST INFRASTRUCTURE).

Reference lines followed (all under /root/reference/src):
  initializeACPowerFlow / changeSlackBus!   powerFlow/acPowerFlow.jl:1312-1358
  newtonJacobian                           powerFlow/acPowerFlow.jl:89-175
  mismatch!                                powerFlow/acPowerFlow.jl:645-685  (+ backend/equations.jl:63-103,126)
  solve! (Jacobian fill, update)           powerFlow/acPowerFlow.jl:793-911  (+ backend/equations.jl:105-143)
  powerFlow! loop                          powerFlow/acPowerFlow.jl:1389-1433
  factorization/solution!                  backend/utility.jl:470-586 -> SuiteSparse (third party); here SuperLU.

The scalar loops are written exactly in the reference's order so summation order (and therefore the
last bits of every value) matches the reference's arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import sin, cos
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .system import System
from .model import AcModel, ac_model


@dataclass
class NewtonRaphson:
    sys: System
    mdl: AcModel
    bus_type: np.ndarray     # after the PV->PQ / slack fix-ups of initializeACPowerFlow
    slack: int
    vm: np.ndarray           # analysis.voltage.magnitude
    va: np.ndarray           # analysis.voltage.angle
    pq: np.ndarray           # 0-based position in the mismatch vector, -1 if not PQ
    pvpq: np.ndarray         # 0-based, -1 for the slack bus
    pcount: np.ndarray
    j_colptr: np.ndarray     # 0-based CSC pattern of the Jacobian
    j_rowval: np.ndarray
    j_nzval: np.ndarray
    mismatch: np.ndarray
    increment: np.ndarray
    iteration: int = 0
    lu_options: dict | None = None

    @property
    def dim(self) -> int:
        return len(self.mismatch)


def initialize(sys: System):
    """initializeACPowerFlow + changeSlackBus! (acPowerFlow.jl:1312-1358). Returns copies; `sys` untouched."""
    bus_type = sys.bus_type.copy()
    slack = sys.slack
    vm = sys.vm.copy()
    va = sys.va.copy()
    for i in range(sys.n):
        has_gen = len(sys.bus_gens[i]) > 0
        if not has_gen and bus_type[i] == 2:
            bus_type[i] = 1
        if has_gen and bus_type[i] != 1:
            vm[i] = sys.gen_vm[sys.bus_gens[i][0]]
    if len(sys.bus_gens[slack]) == 0:
        bus_type[slack] = 1
        for i in range(sys.n):
            if bus_type[i] == 2 and len(sys.bus_gens[i]) > 0:
                bus_type[i] = 3
                slack = i
                break
        if bus_type[slack] == 1:
            raise RuntimeError("The slack bus is missing.")
    return bus_type, slack, vm, va


def newton_jacobian(mdl: AcModel, bus_type: np.ndarray, slack: int):
    """Index maps and CSC pattern of the Jacobian (acPowerFlow.jl:89-175), 0-based."""
    n = mdl.n
    pq = np.full(n, -1, dtype=np.int64)
    pvpq = np.full(n, -1, dtype=np.int64)
    pvpq_num = 0
    pq_num = 0
    for i in range(n):
        if bus_type[i] == 1:
            pq[i] = pq_num + n - 1
            pq_num += 1
        if bus_type[i] != 3:
            pvpq[i] = pvpq_num
            pvpq_num += 1
    dim = n + pq_num - 1
    colcount = np.zeros(dim, dtype=np.int64)
    pcount = np.zeros(n, dtype=np.int64)
    qcount = np.zeros(n, dtype=np.int64)
    for i in range(n):
        if i == slack:
            continue
        for ptr in range(mdl.colptr[i], mdl.colptr[i + 1]):
            t = bus_type[mdl.rowval[ptr]]
            if t != 3:
                pcount[i] += 1
            if t == 1:
                qcount[i] += 1
        colcount[pvpq[i]] = pcount[i] + qcount[i]
        if bus_type[i] == 1:
            colcount[pq[i]] = pcount[i] + qcount[i]
    colptr = np.zeros(dim + 1, dtype=np.int64)
    colptr[1:] = np.cumsum(colcount)
    nnz = int(colptr[-1])
    rowval = np.zeros(nnz, dtype=np.int64)
    for i in range(n):
        if i == slack:
            continue
        is_pq = bus_type[i] == 1
        pa = colptr[pvpq[i]]
        qa = pa + pcount[i]
        pm = colptr[pq[i]] if is_pq else 0
        qm = pm + pcount[i] if is_pq else 0
        for ptr in range(mdl.colptr[i], mdl.colptr[i + 1]):
            row = mdl.rowval[ptr]
            t = bus_type[row]
            if t != 3:
                rowval[pa] = pvpq[row]
                pa += 1
                if is_pq:
                    rowval[pm] = pvpq[row]
                    pm += 1
            if t == 1:
                rowval[qa] = pq[row]
                qa += 1
                if is_pq:
                    rowval[qm] = pq[row]
                    qm += 1
    return pq, pvpq, pcount, colptr, rowval


def newton_raphson(sys: System, mdl: AcModel | None = None, lu_options: dict | None = None) -> NewtonRaphson:
    """newtonRaphson(system) (acPowerFlow.jl:39-87)."""
    if mdl is None:
        mdl = ac_model(sys)
    bus_type, slack, vm, va = initialize(sys)
    pq, pvpq, pcount, colptr, rowval = newton_jacobian(mdl, bus_type, slack)
    dim = len(colptr) - 1
    return NewtonRaphson(sys, mdl, bus_type, slack, vm, va, pq, pvpq, pcount, colptr, rowval,
                         np.zeros(len(rowval)), np.zeros(dim), np.zeros(dim), 0, lu_options)


def mismatch(a: NewtonRaphson):
    """mismatch!(analysis) (acPowerFlow.jl:645-685). Returns (stopP, stopQ)."""
    m, sys = a.mdl, a.sys
    V, T = a.vm, a.va
    stop_p = 0.0
    stop_q = 0.0
    for i in range(m.n):
        if i == a.slack:
            continue
        k = a.pvpq[i]
        q = a.pq[i]
        cur_p = 0.0
        cur_q = 0.0
        is_pq = a.bus_type[i] == 1
        for ptr in range(m.colptr[i], m.colptr[i + 1]):
            row = m.rowval[ptr]
            y = m.nzval_t[ptr]                      # Y[i,row] (equations.jl:63-68)
            G, B = y.real, y.imag
            d = T[i] - T[row]
            s, c = sin(d), cos(d)
            cur_p += V[row] * (G * c + B * s)       # PiQiSumPlus (equations.jl:78-87)
            if is_pq:
                cur_q += V[row] * (G * s - B * c)   # PiQiSumMinus (equations.jl:89-98)
        a.mismatch[k] = V[i] * cur_p - sys.supply_p[i] + sys.pd[i]
        stop_p = max(stop_p, abs(a.mismatch[k]))
        if is_pq:
            a.mismatch[q] = V[i] * cur_q - sys.supply_q[i] + sys.qd[i]
            stop_q = max(stop_q, abs(a.mismatch[q]))
    return stop_p, stop_q


def fill_jacobian(a: NewtonRaphson):
    """Jacobian fill of solve! (acPowerFlow.jl:813-888)."""
    m = a.mdl
    V, T = a.vm, a.va
    nz = a.j_nzval
    for i in range(m.n):
        if i == a.slack:
            continue
        is_pq = a.bus_type[i] == 1
        pa = a.j_colptr[a.pvpq[i]]
        qa = pa + a.pcount[i]
        pm = a.j_colptr[a.pq[i]] if is_pq else 0
        qm = pm + a.pcount[i] if is_pq else 0
        for j in range(m.colptr[i], m.colptr[i + 1]):
            row = m.rowval[j]
            t = a.bus_type[row]
            if t == 3:
                continue
            y = m.nzval[j]                          # Y[row,i]
            G, B = y.real, y.imag
            if row != i:
                d = T[row] - T[i]
                s, c = sin(d), cos(d)
                nz[pa] = V[row] * V[i] * (G * s - B * c)            # Piθj (equations.jl:109)
                pa += 1
                if t == 1:
                    nz[qa] = -V[row] * V[i] * (G * c + B * s)       # Qiθj (:134)
                    qa += 1
                if is_pq:
                    nz[pm] = V[row] * (G * c + B * s)               # PiVj (:117)
                    pm += 1
                if is_pq and t == 1:
                    nz[qm] = V[row] * (G * s - B * c)               # QiVj (:142)
                    qm += 1
            else:
                cur_t = 0.0
                cur_v = 0.0
                for ptr in range(m.colptr[i], m.colptr[i + 1]):
                    q = m.rowval[ptr]
                    yk = m.nzval_t[ptr]
                    Gk, Bk = yk.real, yk.imag
                    d = T[i] - T[q]
                    s, c = sin(d), cos(d)
                    cur_t += V[q] * (Gk * s - Bk * c)               # PiQiSumMinus
                    if is_pq:
                        cur_v += V[q] * (Gk * c + Bk * s)           # PiQiSumPlus
                nz[pa] = V[row] * (-cur_t) - B * (V[row] * V[row])        # Piθi (:105)
                pa += 1
                if is_pq:
                    nz[qa] = V[row] * cur_v - G * (V[row] * V[row])       # Qiθi (:130)
                    qa += 1
                    nz[pm] = cur_v + G * V[row]                     # PiVi (:113)
                    pm += 1
                    nz[qm] = cur_t - B * V[row]                     # QiVi (:138)
                    qm += 1


def jacobian_csc(a: NewtonRaphson) -> sp.csc_matrix:
    return sp.csc_matrix((a.j_nzval, a.j_rowval, a.j_colptr), shape=(a.dim, a.dim))


def solve(a: NewtonRaphson):
    """solve!(analysis) (acPowerFlow.jl:793-911): fill J, factor, solve, update state."""
    fill_jacobian(a)
    opts = a.lu_options or {}
    lu = spla.splu(jacobian_csc(a), **opts)         # stands in for UMFPACK/KLU (utility.jl:470-500)
    a.increment[:] = lu.solve(a.mismatch)           # solution! (utility.jl:576-582)
    for i in range(a.mdl.n):
        if a.bus_type[i] == 1:
            a.vm[i] = a.vm[i] - a.increment[a.pq[i]]
        if i != a.slack:
            a.va[i] = a.va[i] - a.increment[a.pvpq[i]]
    a.iteration += 1


def power_flow(a: NewtonRaphson, iteration: int = 20, tolerance: float = 1e-8, trace: list | None = None):
    """powerFlow!(analysis) loop (acPowerFlow.jl:1389-1433). Returns converged flag."""
    a.iteration = 0
    converged = False
    for _ in range(iteration + 1):
        dp, dq = mismatch(a)
        if trace is not None:
            trace.append((dp, dq))
        if dp < tolerance and dq < tolerance:
            converged = True
            break
        if a.iteration == iteration:
            break
        solve(a)
    return converged


def export_one_based(a: NewtonRaphson) -> dict:
    """The reference's own 1-based Int64 arrays (pq/pvpq use 0 for 'absent')."""
    return {
        "pq": (a.pq + 1).astype(np.int64),
        "pvpq": (a.pvpq + 1).astype(np.int64),
        "pcount": a.pcount.astype(np.int64),
        "j_colptr": (a.j_colptr + 1).astype(np.int64),
        "j_rowval": (a.j_rowval + 1).astype(np.int64),
    }


def reactive_limit(a: NewtonRaphson) -> np.ndarray:
    """reactiveLimit!(analysis) (acPowerFlow.jl:1081-1156): generators whose reactive output violates its capability
    are fixed at the limit and their bus becomes a demand (PQ) bus; a converted slack bus hands the role to the
    first generator (PV) bus. Mutates `a.sys` (types, slack, generator outputs, bus supply) like the reference
    mutates `system`; returns the violation flags (-1 below minimum, +1 above maximum)."""
    from .post import powers, generator_powers
    sys = a.sys
    sys.bus_type = a.bus_type.copy()          # the analysis' fix-ups live in system.bus.layout in the reference
    sys.slack = a.slack
    pw = powers(sys, a.mdl, a.vm, a.va)
    pg, qg = generator_powers(sys, pw["injection_active"], pw["injection_reactive"], a.slack)
    violate = np.zeros(sys.ngen, dtype=np.int64)
    sys.supply_p[:] = 0.0
    sys.supply_q[:] = 0.0
    for k in range(sys.ngen):
        if sys.gen_status[k] == 1:
            b = int(sys.gen_bus[k])
            sys.gen_p[k] = pg[k]
            sys.supply_p[b] += pg[k]
            sys.supply_q[b] += qg[k]
    for i in range(sys.ngen):
        if sys.gen_status[i] == 0:
            continue
        if sys.gen_qmin[i] < sys.gen_qmax[i]:
            j = int(sys.gen_bus[i])
            vmin = qg[i] < sys.gen_qmin[i]
            vmax = qg[i] > sys.gen_qmax[i]
            if sys.bus_type[j] != 1 and (vmin or vmax):
                if vmin:
                    violate[i] = -1
                    new_q = sys.gen_qmin[i]
                if vmax:
                    violate[i] = 1
                    new_q = sys.gen_qmax[i]
                sys.bus_type[j] = 1
                sys.supply_q[j] -= qg[i]
                sys.gen_q[i] = new_q
                sys.supply_q[j] += new_q
                if j == sys.slack:
                    for k in range(sys.n):
                        if sys.bus_type[k] == 2:
                            sys.slack = k
                            sys.bus_type[k] = 3
                            break
    if sys.bus_type[sys.slack] != 3:
        raise RuntimeError("The slack bus is missing.")
    return violate


def adjust_angle(a: NewtonRaphson, slack: int):
    """adjustAngle!(analysis; slack) (acPowerFlow.jl:1186-1196)."""
    a.va = a.va + (a.sys.va[slack] - a.va[slack])


# ----------------------------------------------------------------------------- fast Newton-Raphson
@dataclass
class FastNewtonRaphson:
    sys: System
    mdl: AcModel
    bus_type: np.ndarray
    slack: int
    vm: np.ndarray
    va: np.ndarray
    pq: np.ndarray              # 0-based position among PQ buses, -1 otherwise
    pvpq: np.ndarray            # 0-based position among non-slack buses, -1 for the slack
    active: sp.csc_matrix       # B'  ((n-1) x (n-1))
    reactive: sp.csc_matrix     # B'' (npq x npq)
    mism_p: np.ndarray
    mism_q: np.ndarray
    bx: bool
    iteration: int = 0
    lu_p: object = None
    lu_q: object = None


def fast_newton_raphson(sys: System, bx: bool, mdl: AcModel | None = None) -> FastNewtonRaphson:
    """fastNewtonRaphsonBX / XB (acPowerFlow.jl:215-339): the two constant Jacobians on the Ybus pattern
    (fastNewtonJacobian :341-412), filled branch by branch (fastNewtonJacobian! :414-451, jacobianCoefficient
    :453-480, Pijθij / Pijθi :482-505) plus the shunt susceptance of PQ buses (:329-335)."""
    if mdl is None:
        mdl = ac_model(sys)
    bus_type, slack, vm, va = initialize(sys)
    n = sys.n
    pq = np.full(n, -1, dtype=np.int64)
    pvpq = np.full(n, -1, dtype=np.int64)
    npq = nps = 0
    for i in range(n):
        if bus_type[i] == 1:
            pq[i] = npq; npq += 1
        if bus_type[i] != 3:
            pvpq[i] = nps; nps += 1
    # patterns: for every non-slack column bus the Ybus rows that are non-slack (P) / PQ with a PQ column (Q)
    rp, cp, rq, cq = [], [], [], []
    for i in range(n):
        if i == slack:
            continue
        for ptr in range(mdl.colptr[i], mdl.colptr[i + 1]):
            row = int(mdl.rowval[ptr])
            if bus_type[row] != 3:
                rp.append(pvpq[row]); cp.append(pvpq[i])
            if bus_type[i] == 1 and bus_type[row] == 1:
                rq.append(pq[row]); cq.append(pq[i])
    P = sp.lil_matrix((n - 1, n - 1))
    Q = sp.lil_matrix((npq, npq))
    for k in range(sys.nbr):
        if sys.status[k] != 1:
            continue
        i, j = int(sys.frm[k]), int(sys.to[k])
        bsi = 0.5 * sys.b[k]
        tinv = 1 / sys.tap[k]
        s, c = np.sin(sys.shift[k]), np.cos(sys.shift[k])
        if bx:
            bmk = -1 / sys.x[k]
            A, B = mdl.admittance[k].real, mdl.admittance[k].imag
        else:
            bmk = mdl.admittance[k].imag
            A, B = 0.0, -1 / sys.x[k]
        den = c * c + s * s
        pij, pji = (-A * s - B * c) / den, (A * s - B * c) / den
        qa, qb, qc = -bmk * tinv, (bmk + bsi) * tinv * tinv, bmk + bsi
        m_, n_ = pvpq[i], pvpq[j]
        if i != slack and j != slack:
            P[m_, n_] += pij
            P[n_, m_] += pji
        if i != slack:
            P[m_, m_] += B / den
        if j != slack:
            P[n_, n_] += B
        ri, rj = pq[i], pq[j]
        if ri >= 0 and rj >= 0:
            Q[ri, rj] += qa
            Q[rj, ri] += qa
        if bus_type[i] == 1:
            Q[ri, ri] += qb
        if bus_type[j] == 1:
            Q[rj, rj] += qc
    for i in range(n):
        if bus_type[i] == 1 and sys.bs[i] != 0:
            Q[pq[i], pq[i]] += sys.bs[i]
    # keep the structural pattern of the reference (explicit zeros where branches are out of service)
    Pp = sp.csc_matrix((np.zeros(len(rp)), (rp, cp)), shape=(n - 1, n - 1))
    Qp = sp.csc_matrix((np.zeros(len(rq)), (rq, cq)), shape=(npq, npq))
    return FastNewtonRaphson(sys, mdl, bus_type, slack, vm, va, pq, pvpq, (P.tocsc() + Pp).tocsc(), (Q.tocsc() + Qp).tocsc(),
                             np.zeros(n - 1), np.zeros(npq), bx)


def _fnr_sums(a, i):
    sys, mdl, V, T = a.sys, a.mdl, a.vm, a.va
    cur_p = cur_q = 0.0
    for ptr in range(mdl.colptr[i], mdl.colptr[i + 1]):
        j = int(mdl.rowval[ptr])
        y = mdl.nzval_t[ptr]
        th = T[i] - T[j]
        sn, cs = sin(th), cos(th)
        cur_p += V[j] * (y.real * cs + y.imag * sn)
        cur_q += V[j] * (y.real * sn - y.imag * cs)
    return cur_p, cur_q


def fnr_mismatch(a: FastNewtonRaphson):
    """mismatch! for the fast method (acPowerFlow.jl:686-727): power mismatches divided by the voltage magnitude."""
    sys = a.sys
    stop_p = stop_q = 0.0
    for i in range(sys.n):
        if i == a.slack:
            continue
        cur_p, cur_q = _fnr_sums(a, i)
        vinv = 1 / a.vm[i]
        a.mism_p[a.pvpq[i]] = cur_p - (sys.supply_p[i] - sys.pd[i]) * vinv
        stop_p = max(stop_p, abs(a.mism_p[a.pvpq[i]]))
        if a.bus_type[i] == 1:
            a.mism_q[a.pq[i]] = cur_q - (sys.supply_q[i] - sys.qd[i]) * vinv
            stop_q = max(stop_q, abs(a.mism_q[a.pq[i]]))
    return stop_p, stop_q


def fnr_solve(a: FastNewtonRaphson):
    """solve! for the fast method (acPowerFlow.jl:913-983): angle step, reactive mismatch at the new angles, then the
    magnitude step; both matrices are factored once."""
    sys = a.sys
    if a.lu_p is None:
        a.lu_p = spla.splu(a.active.tocsc())
        a.lu_q = spla.splu(a.reactive.tocsc())
    dth = a.lu_p.solve(a.mism_p)
    for i in range(sys.n):
        if i != a.slack:
            a.va[i] += dth[a.pvpq[i]]
    for i in range(sys.n):
        if a.bus_type[i] == 1:
            _, cur_q = _fnr_sums(a, i)
            a.mism_q[a.pq[i]] = cur_q - (sys.supply_q[i] - sys.qd[i]) / a.vm[i]
    dv = a.lu_q.solve(a.mism_q)
    for i in range(sys.n):
        if a.bus_type[i] == 1:
            a.vm[i] += dv[a.pq[i]]
    a.iteration += 1


def fnr_power_flow(a: FastNewtonRaphson, iteration: int = 20, tolerance: float = 1e-8) -> bool:
    a.iteration = 0
    for _ in range(iteration + 1):
        sp_, sq_ = fnr_mismatch(a)
        if sp_ < tolerance and sq_ < tolerance:
            return True
        if a.iteration == iteration:
            break
        fnr_solve(a)
    return False
