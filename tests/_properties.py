"""TEST INFRASTRUCTURE: size-independent checks of a theta-rule diffusion step, for meshes the oracle cannot solve in
seconds.  With A = C_m M + theta dt K and b = (C_m M - (1 - theta) dt K) v_ + dt s (monodomain_model.py:83-96):

* residual      ||A x - b|| / ||b||      what "solved" means, evaluated on the host with SciPy products only;
* conservation  1^T C_m M x = 1^T C_m M v_ + dt 1^T s    because the stiffness matrix has zero column sums (pure
                Neumann problem): diffusion moves charge, only the stimulus adds any.
"""
import numpy as np
import scipy.sparse as sp


def pde_step_defects(indptr, indices, mass, stiff, C_m, theta, dt, v_prev, source, x):
    n = len(indptr) - 1
    Mm = sp.csr_matrix((mass, indices, indptr), shape=(n, n))
    K = sp.csr_matrix((stiff, indices, indptr), shape=(n, n))
    Mv, Kv = Mm @ v_prev, K @ v_prev
    Mx, Kx = Mm @ x, K @ x
    b = C_m * Mv - (1.0 - theta) * dt * Kv + dt * source
    r = C_m * Mx + theta * dt * Kx - b
    residual = float(np.linalg.norm(r) / np.linalg.norm(b))
    total_new, total_old = C_m * Mx.sum(), C_m * Mv.sum() + dt * source.sum()
    conservation = float(abs(total_new - total_old) / max(abs(total_old), 1e-300))
    return residual, conservation
