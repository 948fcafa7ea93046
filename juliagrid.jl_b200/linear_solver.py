"""Host handle of the constant-matrix solver of libjgb200.so (`jgb_lin_*`): one sparse symmetric factorisation on
the device, blocks of right-hand sides solved with it.

Stands where the reference calls `factorization / factorization! / solution!` (src/backend/utility.jl:470-586) from
its linear analyses; `dc_power_flow.py`, `dc_state_estimation.py` and `pmu_state_estimation.py` are the callers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from ._lib import Context, ptr, f64, i64


def _one_based(a: sp.csc_matrix):
    a = a.tocsc()
    a.sort_indices()
    return i64(a.indptr + 1), i64(a.indices + 1), f64(a.data)


class LinearSolver:
    def __init__(self, matrix: sp.csc_matrix, skip: int = -1, ctx: Context | None = None, device: int = 0):
        """matrix: symmetric, both triangles stored; skip: 0-based row/column replaced by the identity (-1: none)."""
        self.ctx = ctx or Context(device)
        self.n = matrix.shape[0]
        self.m = 0
        cp, rv, nz = _one_based(matrix)
        self._nnz = len(nz)
        self.ctx.check(self.ctx.lib.jgb_lin_setup(self.ctx.handle, self.n, ptr(cp, C.c_int64), ptr(rv, C.c_int64),
                                                  ptr(nz, C.c_double), skip + 1))

    def refactor(self, matrix: sp.csc_matrix):
        cp, rv, nz = _one_based(matrix)
        if len(nz) != self._nnz:
            raise ValueError("refactor: the pattern changed; build a new LinearSolver")
        self.ctx.check(self.ctx.lib.jgb_lin_refactor(self.ctx.handle, ptr(nz, C.c_double)))

    def set_projection(self, wh: sp.csc_matrix):
        """wh = precision * coefficient (m x n): right-hand sides become b = wh' z on the device."""
        if wh.shape[1] != self.n:
            raise ValueError("projection must have n columns")
        cp, rv, nz = _one_based(wh)
        self.m = wh.shape[0]
        self.ctx.check(self.ctx.lib.jgb_lin_projection(self.ctx.handle, self.m, ptr(cp, C.c_int64),
                                                       ptr(rv, C.c_int64), ptr(nz, C.c_double)))

    def solve(self, b) -> np.ndarray:
        """b: [n] or [R][n]; returns x of the same shape."""
        b2 = f64(np.atleast_2d(b))
        if b2.shape[1] != self.n:
            raise ValueError("right-hand side must have n entries")
        x = np.empty_like(b2)
        self.ctx.check(self.ctx.lib.jgb_lin_solve(self.ctx.handle, b2.shape[0], ptr(b2, C.c_double),
                                                  ptr(x, C.c_double)))
        return x.reshape(np.shape(b))

    def solve_projected(self, z, out=None) -> np.ndarray:
        """z: [m] or [R][m] measurement vectors; returns x [n] or [R][n]. `out` (optional, [R][n] float64, e.g. a view
        of pinned memory) receives the result without an extra allocation."""
        z2 = f64(np.atleast_2d(z))
        if z2.shape[1] != self.m:
            raise ValueError("measurement vector must have m entries")
        x = np.empty((z2.shape[0], self.n)) if out is None else out
        if x.shape != (z2.shape[0], self.n) or x.dtype != np.float64 or not x.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape [R][n]")
        self.ctx.check(self.ctx.lib.jgb_lin_solve_projected(self.ctx.handle, z2.shape[0], ptr(z2, C.c_double),
                                                            ptr(x, C.c_double)))
        return x[0] if np.ndim(z) == 1 else x

    def solve_dev(self, R: int, in_ptr: int, out_ptr: int, projected: bool):
        """Device pointers: in [R][n] (or [R][m] when projected), out [R][n]."""
        self.ctx.check(self.ctx.lib.jgb_lin_solve_dev(self.ctx.handle, R, C.c_void_p(in_ptr), C.c_void_p(out_ptr),
                                                      1 if projected else 0))

    def dims(self) -> dict:
        v = [C.c_int64(0) for _ in range(4)]
        self.ctx.check(self.ctx.lib.jgb_lin_dims(self.ctx.handle, *[C.byref(q) for q in v]))
        return dict(zip(("n", "m", "nnz_factor", "fronts"), (q.value for q in v)))
