"""Oracle restatement of `acModel!` (TEST INFRASTRUCTURE).

Follows `src/powerSystem/model.jl:23-78` with the CSC builder of `src/backend/sparse.jl:2-101`
(count -> fill -> canonicalize: stable insertion sort by row within a column, duplicates summed in
that order).  Out-of-service branches keep their (zero) off-diagonal entries (`model.jl:70-71`).
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

from .system import System


@dataclass
class AcModel:
    n: int
    colptr: np.ndarray      # int64, 0-based, length n+1       (nodalMatrix.colptr - 1)
    rowval: np.ndarray      # int64, 0-based, sorted per column (nodalMatrix.rowval - 1)
    nzval: np.ndarray       # complex128  Y[row, col]           (nodalMatrix.nzval)
    nzval_t: np.ndarray     # complex128  Y[col, row], same pattern (nodalMatrixTranspose.nzval)
    admittance: np.ndarray  # per branch 1/(r+jx); zero when out of service
    y_ff: np.ndarray        # nodalFromFrom
    y_ft: np.ndarray        # nodalFromTo
    y_tf: np.ndarray        # nodalToFrom
    y_tt: np.ndarray        # nodalToTo

    def position(self, row: int, col: int) -> int:
        lo, hi = self.colptr[col], self.colptr[col + 1]
        k = lo + int(np.searchsorted(self.rowval[lo:hi], row))
        if k >= hi or self.rowval[k] != row:
            raise KeyError((row, col))
        return int(k)

    def dense(self) -> np.ndarray:
        Y = np.zeros((self.n, self.n), dtype=complex)
        for c in range(self.n):
            for p in range(self.colptr[c], self.colptr[c + 1]):
                Y[self.rowval[p], c] = self.nzval[p]
        return Y


def branch_parameters(sys: System):
    """Per-branch Y-parameters, `model.jl:53-64` (zero for out-of-service branches)."""
    m = sys.nbr
    adm = np.zeros(m, dtype=complex)
    y_tt = np.zeros(m, dtype=complex)
    y_ff = np.zeros(m, dtype=complex)
    y_ft = np.zeros(m, dtype=complex)
    y_tf = np.zeros(m, dtype=complex)
    for i in range(m):
        if sys.status[i] == 1:
            adm[i] = 1 / complex(sys.r[i], sys.x[i])
            tinv = 1 / sys.tap[i]
            tr = tinv * complex(np.cos(-sys.shift[i]), np.sin(-sys.shift[i]))
            shunt = complex(sys.g[i], sys.b[i])
            y_tt[i] = adm[i] + 0.5 * shunt
            y_ff[i] = tinv ** 2 * y_tt[i]
            y_ft[i] = -np.conj(tr) * adm[i]
            y_tf[i] = -tr * adm[i]
    return adm, y_ff, y_ft, y_tf, y_tt


def ac_model(sys: System) -> AcModel:
    n, m = sys.n, sys.nbr
    adm, y_ff, y_ft, y_tf, y_tt = branch_parameters(sys)

    # per-column insertion lists: diagonal first (model.jl:43-47), then branch entries in branch order
    cols_rows = [[i] for i in range(n)]
    cols_vals = [[complex(sys.gs[i], sys.bs[i])] for i in range(n)]
    for i in range(m):
        f, t = int(sys.frm[i]), int(sys.to[i])
        if sys.status[i] == 1:
            cols_vals[f][0] += y_ff[i]
            cols_vals[t][0] += y_tt[i]
        cols_rows[t].append(f)       # addEntry!(builder, from, to, nodalFromTo): row=from, col=to
        cols_vals[t].append(y_ft[i])
        cols_rows[f].append(t)       # addEntry!(builder, to, from, nodalToFrom): row=to, col=from
        cols_vals[f].append(y_tf[i])

    colptr = np.zeros(n + 1, dtype=np.int64)
    rowval, nzval = [], []
    for c in range(n):
        rows = np.array(cols_rows[c])
        order = np.argsort(rows, kind="stable")          # canonicalize! insertion sort (sparse.jl:56-67)
        prev = None
        for k in order:
            r = int(rows[k])
            if prev is not None and r == prev:
                nzval[-1] += cols_vals[c][k]              # merge duplicates (sparse.jl:71-83)
            else:
                rowval.append(r)
                nzval.append(cols_vals[c][k])
                prev = r
        colptr[c + 1] = len(rowval)

    rowval = np.array(rowval, dtype=np.int64)
    nzval = np.array(nzval, dtype=complex)
    mdl = AcModel(n, colptr, rowval, nzval, None, adm, y_ff, y_ft, y_tf, y_tt)
    # nodalMatrixTranspose = copy(transpose(Y)) (model.jl:75): same pattern (structural symmetry)
    nz_t = np.empty_like(nzval)
    for c in range(n):
        for p in range(colptr[c], colptr[c + 1]):
            nz_t[p] = nzval[mdl.position(c, int(rowval[p]))]
    mdl.nzval_t = nz_t
    return mdl


def apply_outage(sys: System, mdl: AcModel, k: int) -> AcModel:
    """Branch k out of service, pattern kept (updateBranchMain!/acNodalUpdate!, branch.jl:313-431,
    model.jl:81-110): the branch's Y-parameters are subtracted from the stored entries in place."""
    out = AcModel(mdl.n, mdl.colptr, mdl.rowval, mdl.nzval.copy(), mdl.nzval_t.copy(), mdl.admittance.copy(),
                  mdl.y_ff.copy(), mdl.y_ft.copy(), mdl.y_tf.copy(), mdl.y_tt.copy())
    i, j = int(sys.frm[k]), int(sys.to[k])
    for (r, c, v) in ((i, i, mdl.y_ff[k]), (j, j, mdl.y_tt[k]), (i, j, mdl.y_ft[k]), (j, i, mdl.y_tf[k])):
        out.nzval[out.position(r, c)] -= v
        out.nzval_t[out.position(c, r)] -= v
    out.admittance[k] = 0
    out.y_ff[k] = out.y_ft[k] = out.y_tf[k] = out.y_tt[k] = 0
    return out
