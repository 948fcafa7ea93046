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


class Refactor:
    """KLU-style numeric refactorisation on a fixed pattern (oracle/csrc/oracle_lu.c): the first call factors with
    SuperLU (ordering, pivot order, patterns of L and U); every later call reuses all of that and only recomputes the
    values, like the reference's `lu!` / `klu!` after its first `factorization` (backend/utility.jl:470-500)."""

    def __init__(self, lu_options=None):
        self.lu_options = lu_options or {}
        self.ready = False
        self.symbolic_calls = 0
        self.numeric_calls = 0

    def _symbolic(self, J):
        n = J.shape[0]
        lu = spla.splu(J, **self.lu_options)                 # ordering + pivot order on the real values (klu_analyze + klu_factor)
        # SciPy drops the entries of L and U that are exactly zero in THIS factorisation (a flat start zeroes many
        # Jacobian entries), so the structural patterns of Pr J Pc = L U are computed symbolically (olu_symbolic) for the
        # ordering and pivot order SuperLU chose.
        row1 = np.asarray(lu.perm_r, dtype=np.int64)          # row i of J -> row perm_r[i]
        col1 = np.asarray(lu.perm_c, dtype=np.int64)          # column j of J -> column perm_c[j]
        coo = J.tocoo()
        B = sp.csc_matrix((np.ones(coo.nnz), (row1[coo.row], col1[coo.col])), shape=(n, n))
        B.sort_indices()
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        Bp, Bi = i32(B.indptr), i32(B.indices)
        cap = int(1.5 * (lu.L.nnz + lu.U.nnz)) + 4 * n
        while True:
            Lp, Up = np.zeros(n + 1, dtype=np.int32), np.zeros(n + 1, dtype=np.int32)
            Li, Ui = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
            mark, heap = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
            rc = lib().olu_symbolic(C.c_int32(n), _p(Bp, C.c_int32), _p(Bi, C.c_int32), _p(Lp, C.c_int32),
                                    _p(Li, C.c_int32), C.c_int64(cap), _p(Up, C.c_int32), _p(Ui, C.c_int32),
                                    C.c_int64(cap), _p(mark, C.c_int32), _p(heap, C.c_int32))
            if rc == 0:
                break
            cap *= 2
        self.n = n
        self.Lp, self.Li, self.Lx = Lp, np.ascontiguousarray(Li[:Lp[n]]), np.zeros(int(Lp[n]))
        self.Up, self.Ui, self.Ux = Up, np.ascontiguousarray(Ui[:Up[n]]), np.zeros(int(Up[n]))
        row2 = col2 = np.arange(n)
        self.row_new = i32(row2[row1])
        col_src = np.empty(n, dtype=np.int32)
        col_src[col2[col1]] = np.arange(n, dtype=np.int32)
        self.col_src = col_src
        self.work = np.zeros(n)
        self.x = np.zeros(n)
        self.ready = True
        self.symbolic_calls += 1
        self.nnz_lu = int(Lp[n] + Up[n] - n)

    def factor(self, J):
        """J: csc_matrix with int32 indices on the pattern of the first call."""
        first = not self.ready
        if first:
            self._symbolic(J)
        rc = lib().olu_refactor(C.c_int32(self.n), _p(J.indptr, C.c_int32), _p(J.indices, C.c_int32),
                                _p(J.data, C.c_double), _p(self.row_new, C.c_int32), _p(self.col_src, C.c_int32),
                                _p(self.Lp, C.c_int32), _p(self.Li, C.c_int32), _p(self.Lx, C.c_double),
                                _p(self.Up, C.c_int32), _p(self.Ui, C.c_int32), _p(self.Ux, C.c_double),
                                _p(self.work, C.c_double))
        self.numeric_calls += 0 if first else 1
        if rc != 0:
            raise np.linalg.LinAlgError(f"zero pivot in column {rc - 1} of the refactorisation")

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        lib().olu_solve(C.c_int32(self.n), _p(self.row_new, C.c_int32), _p(self.col_src, C.c_int32),
                        _p(self.Lp, C.c_int32), _p(self.Li, C.c_int32), _p(self.Lx, C.c_double),
                        _p(self.Up, C.c_int32), _p(self.Ui, C.c_int32), _p(self.Ux, C.c_double), _p(b, C.c_double),
                        _p(self.x, C.c_double), _p(self.work, C.c_double))
        return self.x.copy()


class FastNR:
    """Newton-Raphson with the C assembly loops and SuperLU (`splu`) standing in for UMFPACK/KLU.
    lu_options=None -> SuperLU defaults (COLAMD + partial pivoting, closest to the reference's `LU`);
    NOPIVOT -> MMD_AT_PLUS_A, diag_pivot_thresh=0, SymmetricMode (the faster setting)."""
    NOPIVOT = dict(permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))

    def __init__(self, a: _nr.NewtonRaphson, lu_options=None, refactor=False):
        # refactor=True: symbolic analysis + pivot order once (first solve of the object), numeric refactorisation
        # afterwards — the reference's `factorization` / `factorization!` split; False: a fresh SuperLU call per solve
        self.refactor = Refactor(lu_options) if refactor else None
        self.t_asm = self.t_fac = self.t_sol = 0.0
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
        import time
        t0 = time.perf_counter()
        r = self._mismatch()
        self.t_asm += time.perf_counter() - t0
        return r

    def _mismatch(self):
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
        import time
        t0 = time.perf_counter()
        self.jacobian()
        t1 = time.perf_counter()
        J = sp.csc_matrix((self.jnz, self.jrowval, self.jcolptr32), shape=(len(self.mism),) * 2)
        if self.refactor is not None:
            self.refactor.factor(J)
            t2 = time.perf_counter()
            inc = self.refactor.solve(self.mism)
        else:
            lu = spla.splu(J, **self.lu_options)
            t2 = time.perf_counter()
            inc = lu.solve(self.mism)
        t3 = time.perf_counter()
        self.t_asm += t1 - t0
        self.t_fac += t2 - t1
        self.t_sol += t3 - t2
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


class FastWLS:
    """Gauss-Newton WLS with the C normalEquation! loops (codes 1, 6-11, 16, 17; diagonal precision), SciPy SpGEMM for
    H'WH like the reference's two stdlib SpGEMMs, and SuperLU with the symmetric no-pivot settings for the gain."""

    def __init__(self, g, lu_options=None, refactor=False):
        from . import wls as _w
        self.g = g
        self.refactor = Refactor(lu_options or FastNR.NOPIVOT) if refactor else None
        self.t_rows = self.t_gain = self.t_fac = self.t_sol = 0.0
        sysm, m = g.sys, g.mdl
        n = sysm.n
        self.n = n
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.ycolptr, self.yrowval = i64(m.colptr), i64(m.rowval)
        self.y = np.ascontiguousarray(m.nzval, dtype=np.complex128).view(np.float64).copy()
        self.yt = np.ascontiguousarray(m.nzval_t, dtype=np.complex128).view(np.float64).copy()
        self.ydiag = i64([m.position(i, i) for i in range(n)])
        self.frm, self.to = i64(sysm.frm), i64(sysm.to)
        self.bg, self.bb = f64(m.admittance.real), f64(m.admittance.imag)
        self.bgsi, self.bbsi = f64(0.5 * sysm.g), f64(0.5 * sysm.b)
        self.btinv, self.bphi = f64(1 / sysm.tap), f64(sysm.shift)
        self.type, self.index = np.ascontiguousarray(g.type, dtype=np.int8), i64(g.index)
        # per-row H positions in the semantic slot order of the kernel / C loop
        H = sp.csc_matrix((np.arange(1, len(g.h_rowval) + 1), g.h_rowval, g.h_colptr), shape=(g.m, 2 * n)).tocsr()
        slotptr, slotpos = [0], []
        for r in range(g.m):
            code, k = int(g.type[r]), int(g.index[r])
            cols = {}
            for q in range(H.indptr[r], H.indptr[r + 1]):
                cols[int(H.indices[q])] = int(H.data[q]) - 1
            if code == 1:
                slotpos.append(cols[k + n])
            elif code in (6, 9):
                for p in range(m.colptr[k], m.colptr[k + 1]):
                    j = int(m.rowval[p])
                    slotpos += [cols[j], cols[j + n]]
            elif code in (16, 17):
                slotpos += [cols[k], cols[k + n]]
            elif code != 0:
                i, j = int(sysm.frm[k]), int(sysm.to[k])
                slotpos += [cols[i], cols[i + n], cols[j], cols[j + n]]
            slotptr.append(len(slotpos))
        self.slotptr, self.slotpos = i64(slotptr), i64(slotpos)
        self.wdiag = f64(g.w.diagonal())
        self.mean = f64(g.mean)
        self.vm0, self.va0 = g.vm.copy(), g.va.copy()
        self.vm, self.va = g.vm.copy(), g.va.copy()
        self.res = np.zeros(g.m)
        self.hnz = g.h_nzval.copy()
        self.hrow, self.hcolptr = g.h_rowval.astype(np.int32), g.h_colptr.astype(np.int32)
        self.lu_options = lu_options or FastNR.NOPIVOT
        self.W = sp.diags(self.wdiag).tocsc()
        self.iteration = 0
        self.objective = 0.0

    def reset(self):
        self.vm[:] = self.vm0
        self.va[:] = self.va0

    def increment(self):
        import time
        t0 = time.perf_counter()
        L = lib()
        L.owls_normal_equation.restype = C.c_double
        P = lambda a, t=C.c_int64: _p(a, t)
        D = lambda a: _p(a, C.c_double)
        self.objective = L.owls_normal_equation(
            C.c_int64(self.n), C.c_int64(self.g.m), P(self.ycolptr), P(self.yrowval), D(self.y), D(self.yt),
            P(self.ydiag), P(self.frm), P(self.to), D(self.bg), D(self.bb), D(self.bgsi), D(self.bbsi), D(self.btinv),
            D(self.bphi), _p(self.type, C.c_int8), P(self.index), P(self.slotptr), P(self.slotpos), D(self.wdiag),
            D(self.mean), D(self.vm), D(self.va), D(self.res), D(self.hnz))
        if self.objective != self.objective:
            raise ValueError("measurement code outside the C oracle's set")
        n, sl = self.n, self.g.slack
        lo, hi = self.hcolptr[sl], self.hcolptr[sl + 1]
        saved = self.hnz[lo:hi].copy()
        self.hnz[lo:hi] = 0.0
        H = sp.csc_matrix((self.hnz, self.hrow, self.hcolptr), shape=(self.g.m, 2 * n))
        t1 = time.perf_counter()
        temp = (H.T @ self.W).tocsc()
        gain = (temp @ H).tocsc()
        gain.sort_indices()
        # gain[slack, slack] = 1 (acStateEstimation.jl:889-891). Julia's SpGEMM keeps the structural pattern of H'WH;
        # SciPy's drops every entry whose sum happens to be zero, so its pattern moves with the values. The fixed
        # pattern (|H|'|H| plus the slack diagonal) is built once and each iteration's values are placed into it.
        nv = 2 * n
        if getattr(self, "_gpat", None) is None:
            Ho = sp.csc_matrix((np.ones(len(self.hrow)), self.hrow, self.hcolptr), shape=(self.g.m, nv))
            P = (Ho.T @ Ho + sp.csc_matrix(([1.0], ([sl], [sl])), shape=(nv, nv))).tocsc()
            P.sort_indices()
            self._gpat = (P.indptr.astype(np.int32), P.indices.astype(np.int32))
            cols = np.repeat(np.arange(nv, dtype=np.int64), np.diff(P.indptr))
            self._gkeys = cols * nv + P.indices.astype(np.int64)
            self._gslack = int(np.searchsorted(self._gkeys, sl * nv + sl))
        gcols = np.repeat(np.arange(nv, dtype=np.int64), np.diff(gain.indptr))
        pos = np.searchsorted(self._gkeys, gcols * nv + gain.indices.astype(np.int64))
        data = np.zeros(len(self._gkeys))
        data[pos] = gain.data
        data[self._gslack] = 1.0
        G = sp.csc_matrix((data, self._gpat[1], self._gpat[0]), shape=gain.shape)
        rhs = temp @ self.res
        t2 = time.perf_counter()
        if self.refactor is not None:
            self.refactor.factor(G)
            t3 = time.perf_counter()
            inc = self.refactor.solve(rhs)
        else:
            lu = spla.splu(G, **self.lu_options)
            t3 = time.perf_counter()
            inc = lu.solve(rhs)
        t4 = time.perf_counter()
        self.t_rows += t1 - t0
        self.t_gain += t2 - t1
        self.t_fac += t3 - t2
        self.t_sol += t4 - t3
        inc[sl] = 0.0
        self.hnz[lo:hi] = saved
        self.inc = inc
        return float(np.max(np.abs(inc)))

    def state_estimation(self, iteration=40, tolerance=1e-8):
        self.iteration = 0
        for _ in range(iteration + 1):
            if self.increment() < tolerance:
                return True
            if self.iteration == iteration:
                return False
            self.va += self.inc[:self.n]
            self.vm += self.inc[self.n:]
            self.iteration += 1
        return False
