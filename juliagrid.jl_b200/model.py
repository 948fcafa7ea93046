"""Host-side `acModel!` of the product: Ybus in the reference's own CSC layout (src/powerSystem/model.jl:23-78).

In the Julia drop-in this step stays in JuliaGrid itself (the C ABI takes `system.model.ac.nodalMatrix` as is); the
Python host mirror needs its own builder to feed the same arrays. Vectorised NumPy; the oracle holds an independent
loop-level restatement and the tests compare the two bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .cases import PowerSystem


@dataclass
class AcModel:
    """system.model.ac: `colptr`/`rowval` are 1-based Int64 exactly as Julia's SparseMatrixCSC stores them."""
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray          # nodalMatrix.nzval            Y[row, col]
    nzval_t: np.ndarray        # nodalMatrixTranspose.nzval   Y[col, row] (same pattern)
    admittance: np.ndarray
    y_ff: np.ndarray           # nodalFromFrom
    y_ft: np.ndarray           # nodalFromTo
    y_tf: np.ndarray           # nodalToFrom
    y_tt: np.ndarray           # nodalToTo

    def position(self, row: int, col: int) -> int:
        """0-based index into nzval of entry (row, col), both 0-based."""
        lo, hi = self.colptr[col] - 1, self.colptr[col + 1] - 1
        k = lo + int(np.searchsorted(self.rowval[lo:hi], row + 1))
        if k >= hi or self.rowval[k] != row + 1:
            raise KeyError((row, col))
        return int(k)


def ac_model(system: PowerSystem) -> AcModel:
    n, m = system.n, system.nbr
    on = system.status == 1
    adm = np.zeros(m, dtype=complex)
    adm[on] = 1.0 / (system.r[on] + 1j * system.x[on])
    tinv = np.ones(m)
    tinv[on] = 1.0 / system.tap[on]
    # cis(-shift) = cos(-shift) + i sin(-shift)
    tr = tinv * (np.cos(-system.shift) + 1j * np.sin(-system.shift))
    shunt = system.g + 1j * system.b
    y_tt = np.where(on, adm + 0.5 * shunt, 0)
    y_ff = np.where(on, tinv ** 2 * y_tt, 0)
    y_ft = np.where(on, -np.conj(tr) * adm, 0)
    y_tf = np.where(on, -tr * adm, 0)

    diag = (system.gs + 1j * system.bs).astype(complex)
    # diagonal accumulation in branch order, from-end then to-end (model.jl:66-67)
    ends = np.empty(2 * m, dtype=np.int64)
    vals = np.empty(2 * m, dtype=complex)
    ends[0::2], ends[1::2] = system.frm, system.to
    vals[0::2], vals[1::2] = y_ff, y_tt
    keep = np.repeat(on, 2)
    np.add.at(diag, ends[keep], vals[keep])

    # off-diagonal entries in insertion order: (row=from, col=to, Yft), (row=to, col=from, Ytf) per branch
    rows = np.empty(2 * m, dtype=np.int64)
    cols = np.empty(2 * m, dtype=np.int64)
    offv = np.empty(2 * m, dtype=complex)
    rows[0::2], cols[0::2], offv[0::2] = system.frm, system.to, y_ft
    rows[1::2], cols[1::2], offv[1::2] = system.to, system.frm, y_tf
    rows = np.concatenate([np.arange(n), rows])
    cols = np.concatenate([np.arange(n), cols])
    allv = np.concatenate([diag, offv])
    order = np.lexsort((np.arange(len(rows)), rows, cols))      # stable: column, row, insertion order
    rows, cols, allv = rows[order], cols[order], allv[order]
    first = np.ones(len(rows), dtype=bool)
    first[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    start = np.flatnonzero(first)
    nzval = allv[start].copy()
    if len(start) != len(rows):                                  # parallel branches: sum duplicates in order
        seg = np.cumsum(first) - 1
        for k in np.flatnonzero(~first):
            nzval[seg[k]] += allv[k]
    rowval = rows[start]
    colidx = cols[start]
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, colidx + 1, 1)
    colptr = np.cumsum(colptr)
    # transpose values on the same (structurally symmetric) pattern
    key = colidx * n + rowval
    tkey = rowval * n + colidx
    pos = np.searchsorted(key, tkey)
    if not np.array_equal(key[pos], tkey):
        raise ValueError("Ybus pattern is not structurally symmetric")
    nzval_t = nzval[pos]
    return AcModel(n, (colptr + 1).astype(np.int64), (rowval + 1).astype(np.int64), nzval, nzval_t, adm,
                   y_ff, y_ft, y_tf, y_tt)


def apply_branch_status(system: PowerSystem, k: int, status: int):
    """Host side of updateBranch!(…; status) (updateBranchMain! powerSystem/branch.jl:313-431 + acNodalUpdate!
    model.jl:81-110): the four Ybus entries of branch k change in place on the fixed pattern. Returns the 0-based nzval
    positions touched, their new Y and Y-transpose values, and the branch's new series admittance."""
    mdl = system.model
    i, j = int(system.frm[k]), int(system.to[k])
    pos = np.array([mdl.position(i, i), mdl.position(j, j), mdl.position(i, j), mdl.position(j, i)])
    if status == int(system.status[k]):
        return pos, mdl.nzval[pos].copy(), mdl.nzval_t[pos].copy(), mdl.admittance[k]
    if status == 0:
        dff, dft, dtf, dtt = -mdl.y_ff[k], -mdl.y_ft[k], -mdl.y_tf[k], -mdl.y_tt[k]
    else:
        one = system.copy()
        one.status[:] = 0
        one.status[k] = 1
        one.model = None
        tmp = ac_model(one)
        dff, dft, dtf, dtt = tmp.y_ff[k], tmp.y_ft[k], tmp.y_tf[k], tmp.y_tt[k]
        mdl.admittance[k] = tmp.admittance[k]
    # nodalMatrix: (i,i)+=ff (j,j)+=tt (i,j)+=ft (j,i)+=tf ; transpose: the (j,i) position holds Y[i,j] etc.
    mdl.nzval[pos[0]] += dff
    mdl.nzval[pos[1]] += dtt
    mdl.nzval[pos[2]] += dft
    mdl.nzval[pos[3]] += dtf
    mdl.nzval_t[pos[0]] += dff
    mdl.nzval_t[pos[1]] += dtt
    mdl.nzval_t[pos[3]] += dft
    mdl.nzval_t[pos[2]] += dtf
    if status == 0:
        mdl.y_ff[k] = mdl.y_ft[k] = mdl.y_tf[k] = mdl.y_tt[k] = 0
        mdl.admittance[k] = 0
    else:
        mdl.y_ff[k], mdl.y_ft[k], mdl.y_tf[k], mdl.y_tt[k] = dff, dft, dtf, dtt
    system.status[k] = status
    return pos, mdl.nzval[pos].copy(), mdl.nzval_t[pos].copy(), mdl.admittance[k]
