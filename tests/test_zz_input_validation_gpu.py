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
