"""ctypes front-end of the C restatement (oracle/csrc): same algorithm as oracle/nr.py at C speed, used as the
CPU baseline in bench.py and cross-checked against the NumPy oracle in tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import nr as _nr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class FastNR:
    """Newton-Raphson with the C assembly loops and SuperLU (`splu`) standing in for UMFPACK/KLU.
    lu_options=None -> SuperLU defaults (COLAMD + partial pivoting, closest to the reference's `LU`);
    NOPIVOT -> MMD_AT_PLUS_A, diag_pivot_thresh=0, SymmetricMode (the faster setting)."""
    NOPIVOT = dict(permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))

    def __init__(self, a: _nr.NewtonRaphson, lu_options=None):
        self.a = a
        m = a.mdl
        self.n = m.n
        self.colptr = np.ascontiguousarray(m.colptr, dtype=np.int64)
        self.rowval = np.ascontiguousarray(m.rowval, dtype=np.int64)
        self.set_y(m.nzval, m.nzval_t)
        self.type = np.ascontiguousarray(a.bus_type, dtype=np.int8)
        self.pq = np.ascontiguousarray(a.pq, dtype=np.int64)
        self.pvpq = np.ascontiguousarray(a.pvpq, dtype=np.int64)
        self.pcount = np.ascontiguousarray(a.pcount, dtype=np.int64)
        self.jcolptr = np.ascontiguousarray(a.j_colptr, dtype=np.int64)
        self.jrowval = np.ascontiguousarray(a.j_rowval, dtype=np.int32)
        self.jcolptr32 = self.jcolptr.astype(np.int32)
        s = a.sys
        self.sp, self.sq = np.ascontiguousarray(s.supply_p), np.ascontiguousarray(s.supply_q)
        self.pd, self.qd = np.ascontiguousarray(s.pd), np.ascontiguousarray(s.qd)
        self.vm0, self.va0 = a.vm.copy(), a.va.copy()
        self.vm, self.va = a.vm.copy(), a.va.copy()
        self.mism = np.zeros(a.dim)
        self.jnz = np.zeros(len(a.j_rowval))
        self.stop = np.zeros(2)
        self.lu_options = lu_options or {}
        self.iteration = 0

    def set_y(self, nzval, nzval_t):
        self.y = np.ascontiguousarray(nzval, dtype=np.complex128).view(np.float64).copy()
        self.yt = np.ascontiguousarray(nzval_t, dtype=np.complex128).view(np.float64).copy()

    def reset(self):
        self.vm[:] = self.vm0
        self.va[:] = self.va0

    def mismatch(self):
        L = lib()
        L.onr_mismatch(C.c_int64(self.n), _p(self.colptr, C.c_int64), _p(self.rowval, C.c_int64),
                       _p(self.yt, C.c_double), _p(self.type, C.c_int8), C.c_int64(self.a.slack),
                       _p(self.pq, C.c_int64), _p(self.pvpq, C.c_int64), _p(self.vm, C.c_double),
                       _p(self.va, C.c_double), _p(self.sp, C.c_double), _p(self.sq, C.c_double),
                       _p(self.pd, C.c_double), _p(self.qd, C.c_double), _p(self.mism, C.c_double),
                       _p(self.stop, C.c_double))
        return float(self.stop[0]), float(self.stop[1])

    def jacobian(self):
        L = lib()
        L.onr_jacobian(C.c_int64(self.n), _p(self.colptr, C.c_int64), _p(self.rowval, C.c_int64),
                       _p(self.y, C.c_double), _p(self.yt, C.c_double), _p(self.type, C.c_int8),
                       C.c_int64(self.a.slack), _p(self.pq, C.c_int64), _p(self.pvpq, C.c_int64),
                       _p(self.pcount, C.c_int64), _p(self.jcolptr, C.c_int64), _p(self.vm, C.c_double),
                       _p(self.va, C.c_double), _p(self.jnz, C.c_double))

    def solve(self):
        self.jacobian()
        J = sp.csc_matrix((self.jnz, self.jrowval, self.jcolptr32), shape=(len(self.mism),) * 2)
        inc = spla.splu(J, **self.lu_options).solve(self.mism)
        lib().onr_update(C.c_int64(self.n), _p(self.type, C.c_int8), C.c_int64(self.a.slack), _p(self.pq, C.c_int64),
                         _p(self.pvpq, C.c_int64), _p(np.ascontiguousarray(inc), C.c_double), _p(self.vm, C.c_double),
                         _p(self.va, C.c_double))
        self.iteration += 1

    def power_flow(self, iteration=20, tolerance=1e-8):
        self.iteration = 0
        for _ in range(iteration + 1):
            dp, dq = self.mismatch()
            if dp < tolerance and dq < tolerance:
                return True
            if self.iteration == iteration:
                return False
            self.solve()
        return False
