"""ctypes binding of libmono_b200.so (include/mono_abi.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, the product
path fails loudly here.  The oracle under /oracle is test infrastructure and is never imported from
this package.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmono_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol include/mono_abi.h declares
# (tests/test_abi_symbols.py checks header <-> this table <-> the .so).
SIGNATURES = {
    "mono_abi_version": (C.c_int, []),
    "mono_last_error": (C.c_char_p, [C.c_void_p]),
    "mono_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mono_ctx_destroy": (C.c_int, [C.c_void_p]),
    "mono_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "mono_host_free": (C.c_int, [C.c_void_p]),
    "mono_sync": (C.c_int, [C.c_void_p]),
    "mono_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), c_int64_p]),
    "mono_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mono_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mono_set_halo": (C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "mono_halo_refresh_nccl": (C.c_int, [C.c_void_p]),
    "mono_ode_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "mono_ode_set_states": (C.c_int, [C.c_void_p, c_double_p, C.c_int64]),
    "mono_ode_get_states": (C.c_int, [C.c_void_p, c_double_p, C.c_int64]),
    "mono_ode_set_state_row": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mono_ode_get_state_row": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mono_ode_set_params": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int64, c_double_p, C.c_int]),
    "mono_ode_set_region_params": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_int32_p]),
    "mono_ode_step": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "mono_ode_to_dolfin": (C.c_int, [C.c_void_p]),
    "mono_ode_from_dolfin": (C.c_int, [C.c_void_p]),
    "mono_ode_to_pde": (C.c_int, [C.c_void_p]),
    "mono_pde_to_ode": (C.c_int, [C.c_void_p]),
    "mono_get_v_ode": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_set_v_ode": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_pde_set_matrices": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_int64_p, c_int32_p, c_double_p, c_double_p]),
    "mono_pde_config": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mono_pde_set_ksp_type": (C.c_int, [C.c_void_p, C.c_int]),
    "mono_pde_set_chebyshev": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "mono_pde_set_dt": (C.c_int, [C.c_void_p, C.c_double]),
    "mono_stim_add": (C.c_int, [C.c_void_p, C.c_int64, c_int32_p, c_double_p, C.c_double, C.c_double, C.c_double]),
    "mono_stim_set_amplitude": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "mono_stim_set_window": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "mono_pde_step": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "mono_pde_assign_previous": (C.c_int, [C.c_void_p]),
    "mono_get_v": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_set_v": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_get_v_prev": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_set_v_prev": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_ksp_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), c_double_p, C.POINTER(C.c_int)]),
    "mono_ksp_total_iterations": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "mono_split_step": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "mono_split_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_double]),
    "mono_probe_add": (C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_double_p]),
    "mono_probe_values": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_probe_activation": (C.c_int, [C.c_void_p, C.c_double]),
    "mono_probe_activation_times": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_observe_config": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "mono_activation_map": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_v_minmax": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "mono_get_v_strided": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, c_double_p]),
    "mono_timer_start": (C.c_int, [C.c_void_p, C.c_int]),
    "mono_timer_stop": (C.c_int, [C.c_void_p, C.c_int]),
    "mono_timer_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "mono_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "mono_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "mono_l2_flush": (C.c_int, [C.c_void_p]),
    "mono_stage_times_ms": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, C.c_int]),
    "mono_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "mono_bench_dfma": (C.c_int, [C.c_void_p, c_double_p]),
    "mono_bench_grid_sync": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "mono_debug_timeline": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
    "mono_launch_count": (C.c_int, [C.c_void_p, c_int64_p]),
    "mono_pde_dictionary_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), c_double_p, C.POINTER(C.c_int)]),
    "mono_csr_row_patterns": (C.c_int, [C.c_int64, c_int64_p, c_int32_p, c_double_p, c_double_p, C.c_int, C.POINTER(C.c_uint8),
                                        c_int32_p, c_int64_p, c_int64_p]),
    "mono_fem_assemble_p1": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, c_int64_p, c_double_p, C.c_int, C.c_int, c_double_p,
                                       c_int64_p, c_int32_p, c_double_p, c_double_p]),
}

_lib = None


class MonoError(RuntimeError):
    pass


def _point_at_bundled_nccl() -> None:
    """The library dlopens NCCL lazily (mono_comm_init).  When the process has not loaded torch's bundled
    libnccl.so.2 yet, tell it where that copy lives so both end up on the same NCCL."""
    if os.environ.get("MONO_NCCL_LIB"):
        return
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["MONO_NCCL_LIB"] = cand
            return


def load_library():
    """dlopen the C ABI and attach signatures.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MonoError(
            f"{LIB_PATH} not found: build it with `python fenicsx-beat_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    _point_at_bundled_nccl()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class _PinnedBlock:
    """Owner of one cudaMallocHost allocation; freed when the last NumPy view dies."""

    def __init__(self, lib, ptr, nbytes):
        self.lib, self.ptr, self.nbytes = lib, ptr, nbytes

    def __del__(self):  # pragma: no cover
        try:
            self.lib.mono_host_free(self.ptr)
        except Exception:
            pass


def pinned_zeros(shape, dtype=np.float64) -> np.ndarray:
    """NumPy array in page-locked host memory (host mirrors of device vectors).  Falls back to pageable
    memory when no CUDA device is usable (host-only unit tests of the set-up code): this only changes
    the transfer rate, never where the computation runs."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    try:
        lib = load_library()
        ptr = C.c_void_p()
        if lib.mono_host_alloc(nbytes, C.byref(ptr)) != 0 or not ptr:
            raise MonoError("no pinned memory")
    except (MonoError, OSError):
        return np.zeros(shape, dtype=dtype)
    block = _PinnedBlock(lib, ptr, nbytes)
    buf = (C.c_char * max(nbytes, 8)).from_address(ptr.value)
    buf._block = block  # keeps the allocation alive as long as any view of buf exists
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    arr[...] = 0
    return arr


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(c_int32_p)


def _i64p(a: np.ndarray):
    return a.ctypes.data_as(c_int64_p)


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def fem_assemble_p1(tdim: int, n_owned: int, cells: np.ndarray, x: np.ndarray, M) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """mono_fem_assemble_p1 (host threads, no GPU): CSR (indptr, indices, mass, stiff) of the owned rows.
    cells (ncell, tdim+1) local vertex ids, x (n_local, >= tdim), M scalar | (tdim, tdim) | (ncell, tdim, tdim)."""
    lib = load_library()
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    Mv = np.asarray(M, dtype=np.float64)  # (ascontiguousarray would turn a scalar into shape (1,))
    if cells.ndim != 2 or cells.shape[1] != tdim + 1 or x.ndim != 2 or x.shape[1] < tdim:
        raise ValueError("cells must be (ncell, tdim+1) and x (n_local, >= tdim)")
    if Mv.size == 1 and Mv.ndim < 2:
        kind = 0
    elif Mv.shape == (tdim, tdim):
        kind = 1
    elif Mv.shape == (cells.shape[0], tdim, tdim):
        kind = 2
    else:
        raise ValueError(f"conductivity of shape {Mv.shape} fits neither a scalar, a tensor nor one tensor per cell")
    Mv = np.ascontiguousarray(Mv).reshape(-1)
    indptr = np.zeros(n_owned + 1, dtype=np.int64)
    head = (int(tdim), int(x.shape[0]), int(n_owned), int(cells.shape[0]), _i64p(cells), _dp(x), int(x.shape[1]), kind, _dp(Mv), _i64p(indptr))
    rc = lib.mono_fem_assemble_p1(*head, None, None, None)
    if rc == 0:
        nnz = int(indptr[-1])
        indices = np.empty(nnz, dtype=np.int32)
        mass, stiff = np.empty(nnz), np.empty(nnz)
        rc = lib.mono_fem_assemble_p1(*head, _i32p(indices), _dp(mass), _dp(stiff))
    if rc != 0:
        raise MonoError(f"mono_fem_assemble_p1 failed ({rc}): {lib.mono_last_error(None).decode()}")
    return indptr, indices, mass, stiff


def csr_row_patterns(indptr, indices, mass, stiff, max_patterns: int = 64):
    """mono_csr_row_patterns: (pattern_of_row uint8 [255 = not in the dictionary], representative_row, rows_per_pattern)."""
    lib = load_library()
    ip = np.ascontiguousarray(indptr, dtype=np.int64)
    ix = np.ascontiguousarray(indices, dtype=np.int32)
    m, k = _f64(mass, ix.shape), _f64(stiff, ix.shape)
    n = ip.size - 1
    pat = np.empty(n, dtype=np.uint8)
    npat = np.zeros(1, dtype=np.int32)
    rep, cnt = np.zeros(max_patterns, dtype=np.int64), np.zeros(max_patterns, dtype=np.int64)
    rc = lib.mono_csr_row_patterns(n, _i64p(ip), _i32p(ix), _dp(m), _dp(k), int(max_patterns), pat.ctypes.data_as(C.POINTER(C.c_uint8)),
                                   _i32p(npat), _i64p(rep), _i64p(cnt))
    if rc != 0:
        raise MonoError(f"mono_csr_row_patterns failed ({rc}): {lib.mono_last_error(None).decode()}")
    return pat, rep[: npat[0]], cnt[: npat[0]]


class Context:
    """Owner of one mono_ctx (one GPU, one rank).  Thin, one method per ABI entry point."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mono_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise MonoError(f"mono_ctx_create failed ({rc}): {self.lib.mono_last_error(None).decode()}")
        self.h = h
        self.device = device
        self.n_local = 0
        self.ns = 0

    def _ck(self, rc: int):
        if rc < 0:
            raise MonoError(f"libmono_b200 error {rc}: {self.lib.mono_last_error(self.h).decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.mono_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- generic --------------------------------------------------------------------------------
    def sync(self):
        self._ck(self.lib.mono_sync(self.h))

    def device_info(self) -> dict:
        n_sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._ck(self.lib.mono_device_info(self.h, C.byref(n_sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"n_sm": n_sm.value, "cc": (ma.value, mi.value), "mem_bytes": mem.value}

    # ---- comm -----------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.mono_comm_unique_id(buf)
        if rc != 0:
            raise MonoError(f"mono_comm_unique_id failed: {self.lib.mono_last_error(None).decode()}")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(self.lib.mono_comm_init(self.h, nranks, rank, buf))

    def halo_refresh_nccl(self):
        self._ck(self.lib.mono_halo_refresh_nccl(self.h))

    def set_halo(self, nbr_ranks, send_ptr, send_idx, recv_ptr):
        nbr = np.ascontiguousarray(nbr_ranks, dtype=np.int32)
        sp = np.ascontiguousarray(send_ptr, dtype=np.int32)
        si = np.ascontiguousarray(send_idx, dtype=np.int32)
        rp = np.ascontiguousarray(recv_ptr, dtype=np.int32)
        self._ck(self.lib.mono_set_halo(self.h, len(nbr), _i32p(nbr), _i32p(sp), _i32p(si), _i32p(rp)))

    # ---- ODE ------------------------------------------------------------------------------------
    def ode_create(self, model_id: int, scheme_id: int, num_points: int, v_index: int, num_states: int):
        self._ck(self.lib.mono_ode_create(self.h, model_id, scheme_id, num_points, v_index))
        self.ns = num_states
        self.npts = num_points

    def ode_set_states(self, states: np.ndarray):
        s = _f64(states, (self.ns, self.npts))
        self._ck(self.lib.mono_ode_set_states(self.h, _dp(s), s.shape[1]))

    def ode_get_states(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.ns, self.npts), dtype=np.float64)
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.shape == (self.ns, self.npts)
        self._ck(self.lib.mono_ode_get_states(self.h, _dp(out), out.shape[1]))
        return out

    def ode_set_state_row(self, row: int, values):
        v = _f64(values, (self.npts,))
        self._ck(self.lib.mono_ode_set_state_row(self.h, row, _dp(v)))

    def ode_get_state_row(self, row: int) -> np.ndarray:
        out = np.empty(self.npts, dtype=np.float64)
        self._ck(self.lib.mono_ode_get_state_row(self.h, row, _dp(out)))
        return out

    def ode_set_params(self, params: np.ndarray, derived: np.ndarray | None = None):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            d = _f64(derived if derived is not None else np.zeros(0))
            self._ck(self.lib.mono_ode_set_params(self.h, _dp(p), p.shape[0], 0, 0, _dp(d) if d.size else None, d.size))
        elif p.ndim == 2:
            if p.shape[1] != self.npts:
                raise ValueError(f"per-node parameters need shape (num_parameters, {self.npts}), got {p.shape}")
            self._ck(self.lib.mono_ode_set_params(self.h, _dp(p), p.shape[0], 1, p.shape[1], None, 0))
        else:
            raise ValueError("parameters must be 1-D or 2-D")

    def ode_set_region_params(self, params: np.ndarray, derived: np.ndarray, region_of_node: np.ndarray | None):
        """params (n_regions, np), derived (n_regions, nd), region_of_node int32 (num_points,) or None to keep the map."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        d = np.ascontiguousarray(derived, dtype=np.float64).reshape(p.shape[0], -1)
        reg = None if region_of_node is None else np.ascontiguousarray(region_of_node, dtype=np.int32)
        self._ck(self.lib.mono_ode_set_region_params(self.h, p.shape[0], _dp(p), p.shape[1], _dp(d) if d.size else None, d.shape[1],
                                                     _i32p(reg) if reg is not None else None))

    def ode_step(self, t0: float, dt: float):
        self._ck(self.lib.mono_ode_step(self.h, t0, dt))

    def ode_to_dolfin(self):
        self._ck(self.lib.mono_ode_to_dolfin(self.h))

    def ode_from_dolfin(self):
        self._ck(self.lib.mono_ode_from_dolfin(self.h))

    def ode_to_pde(self):
        self._ck(self.lib.mono_ode_to_pde(self.h))

    def pde_to_ode(self):
        self._ck(self.lib.mono_pde_to_ode(self.h))

    def get_v_ode(self, out: np.ndarray) -> np.ndarray:
        self._ck(self.lib.mono_get_v_ode(self.h, _dp(out)))
        return out

    def set_v_ode(self, v):
        self._ck(self.lib.mono_set_v_ode(self.h, _dp(_f64(v, (self.npts,)))))

    # ---- PDE ------------------------------------------------------------------------------------
    def pde_set_matrices(self, n_owned: int, n_ghost: int, indptr, indices, mass, stiff):
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        m = _f64(mass)
        k = _f64(stiff)
        if ip.shape != (n_owned + 1,) or ix.shape != m.shape or m.shape != k.shape or ix.size != ip[-1]:
            raise ValueError("inconsistent CSR arrays")
        self._ck(self.lib.mono_pde_set_matrices(self.h, n_owned, n_ghost, _i64p(ip), _i32p(ix), _dp(m), _dp(k)))
        self.n_owned, self.n_ghost, self.n_local = n_owned, n_ghost, n_owned + n_ghost

    def pde_config(self, C_m, theta, rtol, atol, max_it, pc_type, norm_type, x0_mode):
        self._ck(self.lib.mono_pde_config(self.h, C_m, theta, rtol, atol, max_it, pc_type, norm_type, x0_mode))

    def pde_dictionary_info(self) -> dict:
        n, cover, active = C.c_int(0), C.c_double(0.0), C.c_int(0)
        self._ck(self.lib.mono_pde_dictionary_info(self.h, C.byref(n), C.byref(cover), C.byref(active)))
        return {"patterns": n.value, "rows_covered": cover.value, "active": bool(active.value)}

    def pde_set_chebyshev(self, steps: int, kappa: float):
        self._ck(self.lib.mono_pde_set_chebyshev(self.h, int(steps), float(kappa)))

    def pde_set_ksp_type(self, ksp_type: int):
        self._ck(self.lib.mono_pde_set_ksp_type(self.h, ksp_type))

    def pde_set_dt(self, dt: float):
        self._ck(self.lib.mono_pde_set_dt(self.h, dt))

    def stim_add(self, idx, val, t_start: float, t_end: float, amplitude: float) -> int:
        i = np.ascontiguousarray(idx, dtype=np.int32)
        v = _f64(val, i.shape)
        return self._ck(self.lib.mono_stim_add(self.h, i.size, _i32p(i), _dp(v), t_start, t_end, amplitude))

    def stim_set_amplitude(self, sid: int, amp: float):
        self._ck(self.lib.mono_stim_set_amplitude(self.h, sid, amp))

    def stim_set_window(self, sid: int, t_start: float, t_end: float):
        self._ck(self.lib.mono_stim_set_window(self.h, sid, t_start, t_end))

    def pde_step(self, t0: float, t1: float):
        self._ck(self.lib.mono_pde_step(self.h, t0, t1))

    def pde_assign_previous(self):
        self._ck(self.lib.mono_pde_assign_previous(self.h))

    def get_v(self, out: np.ndarray) -> np.ndarray:
        self._ck(self.lib.mono_get_v(self.h, _dp(out)))
        return out

    def set_v(self, v):
        self._ck(self.lib.mono_set_v(self.h, _dp(_f64(v, (self.n_local,)))))

    def get_v_prev(self, out: np.ndarray) -> np.ndarray:
        self._ck(self.lib.mono_get_v_prev(self.h, _dp(out)))
        return out

    def set_v_prev(self, v):
        self._ck(self.lib.mono_set_v_prev(self.h, _dp(_f64(v, (self.n_local,)))))

    def ksp_info(self):
        its, rn, reason = C.c_int(), C.c_double(), C.c_int()
        self._ck(self.lib.mono_ksp_info(self.h, C.byref(its), C.byref(rn), C.byref(reason)))
        return its.value, rn.value, reason.value

    def ksp_totals(self):
        tot, solves = C.c_int64(), C.c_int64()
        self._ck(self.lib.mono_ksp_total_iterations(self.h, C.byref(tot), C.byref(solves)))
        return tot.value, solves.value

    # ---- fused ----------------------------------------------------------------------------------
    def split_step(self, t0: float, t1: float, theta: float):
        self._ck(self.lib.mono_split_step(self.h, t0, t1, theta))

    def split_solve(self, t0: float, dt: float, nsteps: int, theta: float):
        self._ck(self.lib.mono_split_solve(self.h, t0, dt, nsteps, theta))

    # ---- observers ------------------------------------------------------------------------------
    def probe_add(self, nodes, weights) -> int:
        n = np.ascontiguousarray(nodes, dtype=np.int32)
        w = _f64(weights, n.shape)
        return self._ck(self.lib.mono_probe_add(self.h, n.size, _i32p(n), _dp(w)))

    def probe_values(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float64)
        self._ck(self.lib.mono_probe_values(self.h, _dp(out)))
        return out

    def probe_activation(self, threshold: float):
        self._ck(self.lib.mono_probe_activation(self.h, threshold))

    def probe_activation_times(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float64)
        self._ck(self.lib.mono_probe_activation_times(self.h, _dp(out)))
        return out

    def observe_config(self, activation_map: bool, threshold: float = 0.0, minmax: bool = False):
        self._ck(self.lib.mono_observe_config(self.h, 1 if activation_map else 0, float(threshold), 1 if minmax else 0))

    def activation_map(self) -> np.ndarray:
        out = np.empty(self.n_owned, dtype=np.float64)
        self._ck(self.lib.mono_activation_map(self.h, _dp(out)))
        return out

    def v_minmax(self) -> tuple[float, float]:
        lo, hi = C.c_double(), C.c_double()
        self._ck(self.lib.mono_v_minmax(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def get_v_strided(self, offset: int, stride: int, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.float64)
        self._ck(self.lib.mono_get_v_strided(self.h, int(offset), int(stride), int(count), _dp(out)))
        return out

    # ---- measurement ----------------------------------------------------------------------------
    def timer_start(self, slot: int = 0):
        self._ck(self.lib.mono_timer_start(self.h, slot))

    def timer_stop(self, slot: int = 0):
        self._ck(self.lib.mono_timer_stop(self.h, slot))

    def timer_elapsed_ms(self, slot: int = 0) -> float:
        ms = C.c_float()
        self._ck(self.lib.mono_timer_elapsed_ms(self.h, slot, C.byref(ms)))
        return ms.value

    def event_record(self, idx: int):
        self._ck(self.lib.mono_event_record(self.h, idx))

    def event_elapsed_ms(self, idx0: int, idx1: int) -> float:
        ms = C.c_float()
        self._ck(self.lib.mono_event_elapsed_ms(self.h, idx0, idx1, C.byref(ms)))
        return ms.value

    def l2_flush(self):
        self._ck(self.lib.mono_l2_flush(self.h))

    def stage_timing(self, enable: bool):
        self._ck(self.lib.mono_stage_timing(self.h, 1 if enable else 0))

    def stage_times_ms(self, reset: bool = False):
        ms = np.zeros(2, dtype=np.float64)
        steps = C.c_int64()
        self._ck(self.lib.mono_stage_times_ms(self.h, _dp(ms), C.byref(steps), 1 if reset else 0))
        return {"ode_ms": float(ms[0]), "pde_ms": float(ms[1]), "steps": steps.value}

    def bench_dfma(self) -> float:
        tf = C.c_double()
        self._ck(self.lib.mono_bench_dfma(self.h, C.byref(tf)))
        return tf.value

    def bench_grid_sync(self, n: int = 1000) -> float:
        us = C.c_float()
        self._ck(self.lib.mono_bench_grid_sync(self.h, n, C.byref(us)))
        return us.value

    def debug_timeline(self, enable: bool = True, read: bool = True):
        buf = (C.c_uint64 * 64)()
        self._ck(self.lib.mono_debug_timeline(self.h, 1 if enable else 0, buf if read else None))
        n = int(buf[0])
        return [int(buf[1 + k]) for k in range(min(n, 63))]

    def launch_count(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.mono_launch_count(self.h, C.byref(n)))
        return n.value
