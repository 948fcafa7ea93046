"""ctypes binding of libjgb200.so (include/jgb200.h). The product has no CPU fallback: a missing library or a
missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JGB200_LIB") or os.path.join(_HERE, "libjgb200.so")

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_i8p = C.POINTER(C.c_int8)
c_f64p = C.POINTER(C.c_double)


class JgbError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"jgb200 error {rc}: {msg}")
        self.rc = rc


PROTOTYPES = {
    "jgb_abi_version": (C.c_int32, []),
    "jgb_create": (C.c_void_p, [C.c_int32, C.c_void_p, c_i32p]),
    "jgb_destroy": (None, [C.c_void_p]),
    "jgb_last_error": (C.c_char_p, [C.c_void_p]),
    "jgb_synchronize": (C.c_int32, [C.c_void_p]),
    "jgb_nr_setup": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_f64p, c_f64p, c_i8p, C.c_int64]),
    "jgb_nr_dims": (C.c_int32, [C.c_void_p, c_i64p, c_i64p]),
    "jgb_nr_pattern": (C.c_int32, [C.c_void_p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p]),
    "jgb_nr_set_injection": (C.c_int32, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "jgb_nr_set_state": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_nr_get_state": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_nr_update_y": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_f64p, c_f64p]),
    "jgb_nr_mismatch": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_nr_solve": (C.c_int32, [C.c_void_p]),
    "jgb_nr_get_vectors": (C.c_int32, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_i64p]),
    "jgb_nr_run": (C.c_int32, [C.c_void_p, C.c_int64, C.c_double, c_i64p, c_f64p, c_f64p]),
    "jgb_nr_set_branches": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_f64p, c_f64p, c_f64p, c_f64p, c_i8p]),
    "jgb_nr_power": (C.c_int32, [C.c_void_p] + [c_f64p] * 10),
    "jgb_nr_batch": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_f64p, C.c_int64, C.c_double, c_f64p,
                                 c_f64p, c_i32p, c_i8p, c_i64p]),
    "jgb_nr_batch_dev": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]),
    "jgb_comm_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "jgb_comm_init": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "jgb_allgather_states": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 8),
    "jgb_comm_wait": (C.c_int32, [C.c_void_p, C.c_int32]),
    "jgb_comm_size": (C.c_int32, [C.c_void_p, c_i32p, c_i32p]),
    "jgb_stat": (C.c_double, [C.c_void_p, C.c_char_p]),
    "jgb_profile": (C.c_int32, [C.c_void_p, C.c_int32]),
    "jgb_selfcheck_symbolic": (C.c_int32, [C.c_int64, c_i64p, c_i64p, c_f64p, c_i64p, c_f64p, c_f64p, c_f64p]),
    "jgb_selfcheck_tasks": (C.c_int32, [C.c_int64, c_i64p, c_i64p, c_f64p, c_i64p, c_f64p, c_f64p, c_f64p]),
    "jgb_selfcheck_tree": (C.c_int32, [C.c_int64, c_i64p, c_i64p, c_i64p, C.c_int32, C.c_int64, c_i64p, c_i32p, c_i32p,
                                       c_i32p, c_i32p]),
}

WLS_PROTOTYPES = {
    "jgb_wls_setup": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, c_i64p, c_i64p, c_i8p, c_i64p, c_i64p,
                                  c_i64p, c_i64p, c_f64p, c_i64p, c_i64p, c_f64p, c_f64p, C.c_int64, c_i64p, c_i64p,
                                  c_f64p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "jgb_wls_dims": (C.c_int32, [C.c_void_p, c_i64p, c_i64p]),
    "jgb_wls_gain_pattern": (C.c_int32, [C.c_void_p, c_i64p, c_i64p]),
    "jgb_wls_set_mean": (C.c_int32, [C.c_void_p, c_f64p]),
    "jgb_wls_set_state": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_wls_get_state": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_wls_increment": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_wls_solve": (C.c_int32, [C.c_void_p]),
    "jgb_wls_get_vectors": (C.c_int32, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p, c_i64p]),
    "jgb_wls_run": (C.c_int32, [C.c_void_p, C.c_int64, C.c_double, c_i64p, c_f64p, c_f64p]),
    "jgb_wls_batch": (C.c_int32, [C.c_void_p, C.c_int64, c_f64p, C.c_int64, C.c_double, c_f64p, c_f64p, c_i32p,
                                  c_i8p, c_f64p, c_i64p]),
    "jgb_wls_residual_test": (C.c_int32, [C.c_void_p, C.c_double, c_f64p, c_i64p, c_f64p]),
    "jgb_wls_remove_row": (C.c_int32, [C.c_void_p, C.c_int64]),
    "jgb_wls_update_rows": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_f64p, c_f64p, c_f64p, c_i8p, c_i64p]),
    "jgb_wls_update_y": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_f64p, c_f64p]),
    "jgb_wls_update_branch": (C.c_int32, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                          c_f64p]),
    "jgb_wls_batch_dev": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libjgb200.so and declare every prototype of include/jgb200.h. Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; "
                          f"g.build()'` (make -C juliagrid.jl_b200/csrc). jgb200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for table in (PROTOTYPES, WLS_PROTOTYPES, LIN_PROTOTYPES):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)   # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(PROTOTYPES) + list(WLS_PROTOTYPES) + list(LIN_PROTOTYPES)


LIN_PROTOTYPES = {
    "jgb_lin_setup": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_f64p, C.c_int64]),
    "jgb_lin_refactor": (C.c_int32, [C.c_void_p, c_f64p]),
    "jgb_lin_projection": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_f64p]),
    "jgb_lin_solve": (C.c_int32, [C.c_void_p, C.c_int64, c_f64p, c_f64p]),
    "jgb_lin_solve_projected": (C.c_int32, [C.c_void_p, C.c_int64, c_f64p, c_f64p]),
    "jgb_lin_solve_dev": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    "jgb_lin_dims": (C.c_int32, [C.c_void_p, c_i64p, c_i64p, c_i64p, c_i64p]),
    "jgb_fnr_setup": (C.c_int32, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_f64p, c_i8p, C.c_int64, c_i64p, c_i64p, c_f64p,
                                  c_i64p, c_i64p, c_f64p]),
    "jgb_fnr_set_injection": (C.c_int32, [C.c_void_p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "jgb_fnr_set_state": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_fnr_get_state": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_fnr_mismatch": (C.c_int32, [C.c_void_p, c_f64p, c_f64p]),
    "jgb_fnr_solve": (C.c_int32, [C.c_void_p]),
    "jgb_fnr_run": (C.c_int32, [C.c_void_p, C.c_int64, C.c_double, c_i64p, c_f64p, c_f64p]),
    "jgb_fnr_batch": (C.c_int32, [C.c_void_p, C.c_int64, c_f64p, c_f64p, C.c_int64, C.c_double, c_f64p, c_f64p, c_i32p,
                                  c_i8p, c_i64p]),
}


def ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def i8(a):
    return np.ascontiguousarray(a, dtype=np.int8)


def cplx(a):
    """ComplexF64 vector -> interleaved (re, im) Float64 view, as `Ptr{Float64}` on a Julia Vector{ComplexF64}."""
    return np.ascontiguousarray(a, dtype=np.complex128).view(np.float64)


class Context:
    """One jgb_ctx: one GPU, one stream. `stream` is a raw cudaStream_t handle (int) or None."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load()
        rc = C.c_int32(0)
        self.handle = self.lib.jgb_create(device, C.c_void_p(stream) if stream else None, C.byref(rc))
        if not self.handle:
            raise JgbError(rc.value, self.lib.jgb_last_error(None).decode())
        self.device = device

    def check(self, rc: int) -> int:
        if rc < 0:
            raise JgbError(rc, self.lib.jgb_last_error(self.handle).decode())
        return rc

    def stat(self, key: str) -> float:
        return float(self.lib.jgb_stat(self.handle, key.encode()))

    def synchronize(self):
        self.check(self.lib.jgb_synchronize(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.jgb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
