"""Argument validation of the C ABI on a live context: bad set-up data is refused with MONO_E_INVALID and a message,
never dereferenced or turned into a device fault.  (Runs last: nothing here is a parity test.)"""

import numpy as np
import pytest

from beat_b200._lib import MonoError

pytestmark = pytest.mark.gpu


def _tridiag(n):
    indptr = np.zeros(n + 1, dtype=np.int64)
    cols, mass, stiff = [], [], []
    for r in range(n):
        for c in (r - 1, r, r + 1):
            if 0 <= c < n:
                cols.append(c)
                mass.append(4.0 if c == r else 1.0)
                stiff.append(2.0 if c == r else -1.0)
        indptr[r + 1] = len(cols)
    return indptr, np.array(cols, dtype=np.int32), np.array(mass), np.array(stiff)


def test_csr_is_validated(ctx_factory):
    n = 40
    indptr, cols, mass, stiff = _tridiag(n)
    bad = cols.copy()
    bad[5] = n + 3
    with pytest.raises(MonoError, match="column index out of range"):
        ctx_factory().pde_set_matrices(n, 0, indptr, bad, mass, stiff)
    nodiag = cols.copy()
    nodiag[indptr[7]:indptr[8]] = [5, 6, 8]  # row 7 without (7, 7)
    with pytest.raises(MonoError, match="diagonal"):
        ctx_factory().pde_set_matrices(n, 0, indptr, nodiag, mass, stiff)
    down = indptr.copy()
    down[10] = down[9] - 1
    with pytest.raises(MonoError, match="non-decreasing"):
        ctx_factory().pde_set_matrices(n, 0, down, cols, mass, stiff)
    ctx = ctx_factory()
    ctx.pde_set_matrices(n, 0, indptr, cols, mass, stiff)  # the good one is accepted ...
    with pytest.raises(MonoError, match="already set"):
        ctx.pde_set_matrices(n, 0, indptr, cols, mass, stiff)  # ... once


def test_stimulus_and_probe_indices_are_validated(ctx_factory):
    n = 40
    ctx = ctx_factory()
    ctx.pde_set_matrices(n, 0, *_tridiag(n))
    with pytest.raises(MonoError, match="not an owned dof"):
        ctx.stim_add(np.array([3, n], dtype=np.int32), np.array([1.0, 1.0]), 0.0, 1.0, 1.0)
    with pytest.raises(MonoError, match="out of range"):
        ctx.probe_add(np.array([n], dtype=np.int32), np.array([1.0]))
    with pytest.raises(MonoError, match="theta"):
        ctx.pde_config(1.0, 1.5, 1e-5, 1e-50, 100, 1, 0, 0)


@pytest.mark.parametrize("x0", [0, 1])
def test_stencil_dictionary_is_bit_identical(ctx_factory, monkeypatch, x0):
    """Stencil dictionary (default; MONO_PDE_DICT=0 disables): dictionary rows take their matrix entries from shared memory; same entries,
    same order, so the solve must reproduce the SELL-streaming kernel bit for bit (iterate, iteration count, norm)."""
    from beat_b200 import fem

    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.array([20.0, 7.0, 3.0])], [40, 14, 6])
    indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, np.diag([0.1334, 0.0176, 0.0176]))
    n = indptr.size - 1
    rng = np.random.default_rng(11)
    v_prev = -85.0 + 120.0 * rng.random(n)
    stim_idx = np.arange(0, 50, dtype=np.int32)
    monkeypatch.setenv("MONO_PDE_STREAM", "1")  # force the streaming mode on this small mesh
    out = {}
    for label in ("sell", "dict"):
        if label == "dict":
            monkeypatch.delenv("MONO_PDE_DICT", raising=False)  # the default
        else:
            monkeypatch.setenv("MONO_PDE_DICT", "0")
        ctx = ctx_factory()
        ctx.pde_set_matrices(n, 0, indptr, indices, mass, stiff)
        ctx.pde_config(1.0, 0.5, 1e-10, 1e-50, 200, 1, 0, x0)
        ctx.pde_set_ksp_type(0)
        ctx.stim_add(stim_idx, np.full(50, 0.01), 0.0, 2.0, 50.0)
        ctx.set_v_prev(v_prev)
        ctx.pde_step(0.0, 0.05)
        x1 = ctx.get_v(np.empty(n))
        its1, rnorm1, reason1 = ctx.ksp_info()
        ctx.set_v_prev(x1)
        ctx.pde_step(0.05, 0.15)  # dt changes: A, B and the dictionary values are rebuilt
        out[label] = (x1, ctx.get_v(np.empty(n)), its1, rnorm1, reason1, ctx.ksp_info(), ctx.pde_dictionary_info())
    assert out["sell"][6] == {"patterns": 0, "rows_covered": 0.0, "active": False}
    info = out["dict"][6]
    assert info["patterns"] == 27 and info["rows_covered"] == 1.0 and info["active"]
    assert out["sell"][4] > 0 and out["sell"][2] > 3
    for a, b in zip(out["sell"][:6], out["dict"][:6]):
        assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("ksp", [0, 1])
def test_storage_modes_agree(ctx_factory, monkeypatch, ksp):
    """The same solve in every storage mode of the persistent kernel - A rows and vectors in shared memory (default on
    this small mesh), vectors only (MONO_PDE_NO_MATSMEM), streaming with TMA-staged slices (MONO_PDE_STREAM), streaming
    with direct loads (+ MONO_PDE_NO_STAGING): same iteration count, iterates equal to rounding."""
    from beat_b200 import fem

    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.array([20.0, 7.0, 3.0])], [40, 14, 6])
    indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, np.diag([0.1334, 0.0176, 0.0176]))
    n = indptr.size - 1
    v_prev = -85.0 + 120.0 * np.random.default_rng(3).random(n)
    results = {}
    for label, env in (("matsmem", {}), ("resident", {"MONO_PDE_NO_MATSMEM": "1"}), ("staged", {"MONO_PDE_STREAM": "1"}),
                       ("direct", {"MONO_PDE_STREAM": "1", "MONO_PDE_NO_STAGING": "1"})):
        for k in ("MONO_PDE_NO_MATSMEM", "MONO_PDE_STREAM", "MONO_PDE_NO_STAGING", "MONO_PDE_DICT"):
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv("MONO_PDE_DICT", "0")  # (the dictionary has its own bit-identity test above)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = ctx_factory()
        ctx.pde_set_matrices(n, 0, indptr, indices, mass, stiff)
        ctx.pde_config(1.0, 0.5, 1e-10, 1e-50, 200, 1, 0, 0)
        ctx.pde_set_ksp_type(ksp)
        ctx.set_v_prev(v_prev)
        ctx.pde_step(0.0, 0.05)
        results[label] = (ctx.get_v(np.empty(n)), ctx.ksp_info())
        ctx.close()
    x0, (its0, _, reason0) = results["matsmem"]
    assert reason0 > 0 and its0 > 3
    for label, (x, (its, _, reason)) in results.items():
        assert (its, reason) == (its0, reason0), label
        assert np.abs(x - x0).max() <= 1e-12 * np.abs(x0).max(), label
