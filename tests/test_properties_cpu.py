"""The full-size property checks (tests/_properties.py) accept an exact solve and reject a wrong one."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from _properties import pde_step_defects
from beat_b200 import fem


def test_pde_defects_on_an_exact_solution():
    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.array([20.0, 7.0, 3.0])], [40, 14, 6])
    indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, np.diag([0.1334, 0.0176, 0.0176]))
    n = indptr.size - 1
    Mm = sp.csr_matrix((mass, indices, indptr), shape=(n, n))
    K = sp.csr_matrix((stiff, indices, indptr), shape=(n, n))
    rng = np.random.default_rng(3)
    v_prev = -85.0 + 120.0 * rng.random(n)
    theta, dt, C_m = 0.5, 0.05, 0.01
    source = np.zeros(n)
    source[:50] = 0.5
    A = (C_m * Mm + theta * dt * K).tocsc()
    b = (C_m * Mm - (1 - theta) * dt * K) @ v_prev + dt * source
    x = sla.spsolve(A, b)
    res, cons = pde_step_defects(indptr, indices, mass, stiff, C_m, theta, dt, v_prev, source, x)
    assert res <= 1e-13 and cons <= 1e-12
    bad = x.copy()
    bad[n // 2] += 1e-3
    res_bad, cons_bad = pde_step_defects(indptr, indices, mass, stiff, C_m, theta, dt, v_prev, source, bad)
    assert res_bad > 1e-9 and cons_bad > 1e-12
