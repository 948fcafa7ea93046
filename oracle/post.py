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
