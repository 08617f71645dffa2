"""ORACLE (test infrastructure, never a product path): ctypes access to oracle/c/liboracle_step.so,
the OpenMP C restatement of the reference's split step (see oracle/c/oracle_step.c for the
reference file:line it follows).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs may import this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(HERE, "c")
LIB = os.path.join(CDIR, "liboracle_step.so")

MODEL_ID = {"fhn": 0, "tp06": 1, "torord": 2}
SCHEME_ID = {"forward_explicit_euler": 0, "generalized_rush_larsen": 1}


class _Pde(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("indptr", C.c_void_p), ("indices", C.c_void_p), ("A", C.c_void_p), ("B", C.c_void_p),
        ("dinv", C.c_void_p), ("theta", C.c_double), ("rtol", C.c_double), ("atol", C.c_double), ("max_it", C.c_int),
        ("n_stim", C.c_int), ("stim_load", C.c_void_p), ("stim_t0", C.c_void_p), ("stim_t1", C.c_void_p),
        ("stim_amp", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    srcs = [os.path.join(CDIR, f) for f in ("oracle_step.c", "fhn.c", "tp06.c", "torord.c", "Makefile")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make", "-C", CDIR, "-B", "liboracle_step.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_threads.argtypes = [C.c_int]
        _lib.oracle_ode_step.restype = C.c_int
        _lib.oracle_ode_step.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int64, C.c_double, C.c_double]
        _lib.oracle_pde_step.restype = C.c_int
        _lib.oracle_pde_step.argtypes = [C.POINTER(_Pde), C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_double)]
        _lib.oracle_split_steps.restype = C.c_int64
        _lib.oracle_split_steps.argtypes = [C.POINTER(_Pde), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_double]
    return _lib


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def ode_step(tag: str, scheme: str, states: np.ndarray, params: np.ndarray, t: float, dt: float) -> np.ndarray:
    """In-place states[:] = fun(states, t, params, dt); states (ns, N) C-contiguous, params (np,) or (np, N)."""
    assert states.flags.c_contiguous and states.dtype == np.float64
    p = np.ascontiguousarray(params, dtype=np.float64)
    rc = lib().oracle_ode_step(MODEL_ID[tag], SCHEME_ID[scheme], states.shape[1], states.shape[1], states.ctypes.data,
                               p.ctypes.data, 1 if p.ndim == 2 else 0, p.shape[1] if p.ndim == 2 else 0, t, dt)
    if rc:
        raise RuntimeError("oracle_ode_step failed")
    return states


class SplitProblem:
    """Holds the arrays of one split problem (CSR A, B, Jacobi diagonal, stimuli) for the C driver."""

    def __init__(self, mass, stiff, C_m: float, theta: float, dt: float, stimuli=(), rtol=1e-5, atol=1e-50, max_it=10000):
        A = (C_m * mass + dt * theta * stiff).tocsr()
        B = (C_m * mass - dt * (1.0 - theta) * stiff).tocsr()
        A.sort_indices()
        B.sort_indices()
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        self.n = A.shape[0]
        self.indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.A = np.ascontiguousarray(A.data, dtype=np.float64)
        self.B = np.ascontiguousarray(B.data, dtype=np.float64)
        self.dinv = np.ascontiguousarray(1.0 / A.diagonal())
        self.loads = np.ascontiguousarray([s[0] for s in stimuli], dtype=np.float64).reshape(len(stimuli), self.n)
        self.t0 = np.ascontiguousarray([s[1] for s in stimuli], dtype=np.float64)
        self.t1 = np.ascontiguousarray([s[2] for s in stimuli], dtype=np.float64)
        self.amp = np.ascontiguousarray([s[3] for s in stimuli], dtype=np.float64)
        self.dt = dt
        self.c = _Pde(self.n, self.indptr.ctypes.data, self.indices.ctypes.data, self.A.ctypes.data, self.B.ctypes.data,
                      self.dinv.ctypes.data, theta, rtol, atol, max_it, len(stimuli), self.loads.ctypes.data,
                      self.t0.ctypes.data, self.t1.ctypes.data, self.amp.ctypes.data)

    def pde_step(self, t0: float, t1: float, v_prev: np.ndarray):
        assert abs((t1 - t0) - self.dt) < 1e-12
        x = np.empty(self.n)
        work = np.empty(5 * self.n)
        rn = C.c_double()
        vp = np.ascontiguousarray(v_prev, dtype=np.float64)
        its = lib().oracle_pde_step(C.byref(self.c), t0, t1, vp.ctypes.data, x.ctypes.data, work.ctypes.data, C.byref(rn))
        return x, its, rn.value

    def split_steps(self, tag: str, scheme: str, v_index: int, states: np.ndarray, params: np.ndarray, t0: float,
                    nsteps: int, theta_split: float = 1.0):
        """Advance in place; returns (v, total CG iterations)."""
        assert states.flags.c_contiguous and states.shape[1] == self.n
        p = np.ascontiguousarray(params, dtype=np.float64)
        v = np.empty(self.n)
        tot = lib().oracle_split_steps(C.byref(self.c), MODEL_ID[tag], SCHEME_ID[scheme], v_index, states.shape[1],
                                       states.ctypes.data, p.ctypes.data, v.ctypes.data, t0, self.dt, nsteps, theta_split)
        if tot < 0:
            raise RuntimeError("oracle_split_steps failed")
        return v, int(tot)
