"""Oracle post-processing: powers and currents from a voltage profile (TEST INFRASTRUCTURE).

Produces the exact values the reference's measurement generators consume
(`power!` `src/postprocessing/acAnalysis.jl:30-170`, `current!` `:672-723`): bus injections from the
nodal matrix, branch from/to flows and currents from the per-branch Y-parameters.
"""
from __future__ import annotations

import numpy as np

from .system import System
from .model import AcModel


def powers(sys: System, mdl: AcModel, vm: np.ndarray, va: np.ndarray) -> dict:
    v = vm * np.exp(1j * va)
    n = sys.n
    inj = np.zeros(n, dtype=complex)
    for c in range(n):                       # I_i = sum_j Y[i,j] V_j using the transpose values
        acc = 0j
        for p in range(mdl.colptr[c], mdl.colptr[c + 1]):
            acc += mdl.nzval_t[p] * v[mdl.rowval[p]]
        inj[c] = v[c] * np.conj(acc)
    vi, vj = v[sys.frm], v[sys.to]
    i_from = mdl.y_ff * vi + mdl.y_ft * vj
    i_to = mdl.y_tf * vi + mdl.y_tt * vj
    s_from = vi * np.conj(i_from)
    s_to = vj * np.conj(i_to)
    on = sys.status == 1
    z = lambda a: np.where(on, a, 0.0)
    return {
        "injection_active": inj.real, "injection_reactive": inj.imag,
        "from_active": z(s_from.real), "from_reactive": z(s_from.imag),
        "to_active": z(s_to.real), "to_reactive": z(s_to.imag),
        "from_current_magnitude": z(np.abs(i_from)), "from_current_angle": z(np.angle(i_from)),
        "to_current_magnitude": z(np.abs(i_to)), "to_current_angle": z(np.angle(i_to)),
    }


def generator_powers(sys: System, inj_p: np.ndarray, inj_q: np.ndarray, slack: int):
    """generatorPower for every generator (`src/postprocessing/acAnalysis.jl:538-629`): a bus's reactive output is
    shared between its in-service generators in proportion to their capability ranges; the first generator of the
    slack bus takes the active balance."""
    pg = np.zeros(sys.ngen)
    qg = np.zeros(sys.ngen)
    eps = np.finfo(float).eps
    for idx in range(sys.ngen):
        if sys.gen_status[idx] != 1:
            continue
        b = int(sys.gen_bus[idx])
        gens = sys.bus_gens[b]
        service = len(gens)
        if service == 1:
            pg[idx] = sys.gen_p[idx]
            qg[idx] = inj_q[b] + sys.qd[b]
            if b == slack:
                pg[idx] = inj_p[b] + sys.pd[b]
            continue
        qmin_sum = qmax_sum = 0.0
        qgen_sum = inj_q[b] + sys.qd[b]
        qmin_inf = qmax_inf = 0.0
        qmin_new, qmax_new = sys.gen_qmin[idx], sys.gen_qmax[idx]
        for i in gens:
            if not np.isinf(sys.gen_qmin[i]):
                qmin_sum += sys.gen_qmin[i]
            if not np.isinf(sys.gen_qmax[i]):
                qmax_sum += sys.gen_qmax[i]
        for i in gens:
            if np.isinf(sys.gen_qmin[i]):
                qmin = -abs(qgen_sum) - abs(qmin_sum) - abs(qmax_sum)
                if sys.gen_qmin[i] == np.inf:
                    qmin = -qmin
                if i == idx:
                    qmin_new = qmin
                qmin_inf += qmin
            if np.isinf(sys.gen_qmax[i]):
                qmax = abs(qgen_sum) + abs(qmin_sum) + abs(qmax_sum)
                if sys.gen_qmax[i] == -np.inf:
                    qmax = -qmax
                if i == idx:
                    qmax_new = qmax
                qmax_inf += qmax
        qmin_sum += qmin_inf
        qmax_sum += qmax_inf
        if sys.base_mva * abs(qmin_sum - qmax_sum) > 10 * eps:
            qg[idx] = qmin_new + ((qgen_sum - qmin_sum) / (qmax_sum - qmin_sum)) * (qmax_new - qmin_new)
        else:
            qg[idx] = qmin_new + (qgen_sum - qmin_sum) / service
        if b == slack and gens[0] == idx:
            pg[idx] = inj_p[b] + sys.pd[b]
            for i in gens[1:]:
                pg[idx] -= sys.gen_p[i]
        else:
            pg[idx] = sys.gen_p[idx]
    return pg, qg
