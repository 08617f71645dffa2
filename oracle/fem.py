"""ORACLE (test infrastructure, never a product path): P1 meshes and assembly in NumPy/SciPy.

Restates what dolfinx/FFCx do for the reference on the meshes its tests and demos use.  dolfinx is not
installable here, so the mesh splits are restated from knowledge of dolfinx:
  * create_unit_interval(n): n segments                        (tests/test_stimulation.py:14)
  * create_unit_square(nx, ny, triangle): DiagonalType.right -> each square (v0 v1 / v2 v3) becomes
    triangles (v0, v1, v3) and (v0, v2, v3)                     (README.md:50, tests/test_monodomain.py:49)
  * create_box(..., tetrahedron): each hexahedron becomes the 6 Kuhn tetrahedra that share the
    (0,0,0)-(1,1,1) diagonal                                    (src/beat/geometry.py:133-139)
Discrete operators follow the weak form in src/beat/monodomain_model.py:83-96:
    Mass_ij = int phi_i phi_j dx      K_ij = int (M grad phi_j) . grad phi_i dx
and the stimulus term of src/beat/base_model.py:247-248:  s_i = int I_s phi_i dz(marker).
"""

from __future__ import annotations

import itertools
import math

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------- meshes
def interval_mesh(n: int, a: float = 0.0, b: float = 1.0):
    pts = np.linspace(a, b, n + 1).reshape(-1, 1)
    cells = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1)
    return pts, cells


def rectangle_mesh(nx: int, ny: int, p0=(0.0, 0.0), p1=(1.0, 1.0)):
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # node id = iy*(nx+1) + ix
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v2 + 1
    cells = np.concatenate([np.stack([v0, v1, v3], axis=1), np.stack([v0, v2, v3], axis=1)], axis=0)
    return pts, cells


def box_mesh(n, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0)):
    nx, ny, nz = n
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # node id = (iz*(ny+1) + iy)*(nx+1) + ix
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = ((iz * (ny + 1) + iy) * (nx + 1) + ix).ravel()
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    tets = []
    for perm in itertools.permutations((sx, sy, sz)):  # 6 monotone paths v0 -> v7
        a = v0 + perm[0]
        b = a + perm[1]
        c = b + perm[2]
        tets.append(np.stack([v0, a, b, c], axis=1))
    return pts, np.concatenate(tets, axis=0)


# ------------------------------------------------------------------------------------- geometry
def _cell_geometry(pts: np.ndarray, cells: np.ndarray):
    """volume (ncell,) and P1 gradients (ncell, d+1, gdim)."""
    d = cells.shape[1] - 1
    x0 = pts[cells[:, 0]]
    J = np.stack([pts[cells[:, a]] - x0 for a in range(1, d + 1)], axis=2)  # (ncell, gdim, d)
    assert J.shape[1] == d, "embedded simplices are not needed by the reference tests"
    det = np.linalg.det(J)
    vol = np.abs(det) / math.factorial(d)
    Jinv = np.linalg.inv(J)  # rows are gradients of phi_1..phi_d
    grads = np.empty((cells.shape[0], d + 1, d))
    grads[:, 1:, :] = Jinv
    grads[:, 0, :] = -Jinv.sum(axis=1)
    return vol, grads


def assemble_p1(pts: np.ndarray, cells: np.ndarray, M):
    """Returns (Mass, K) as CSR with identical sparsity.  M: scalar, (d,d) constant tensor, or (ncell,d,d) cell-wise
    tensor (fibre fields, src/beat/conductivities.py:101-104 with a Function f0)."""
    n = pts.shape[0]
    d = cells.shape[1] - 1
    vol, grads = _cell_geometry(pts, cells)
    Mt = np.eye(d) * float(M) if np.ndim(M) == 0 else np.asarray(M, dtype=float)
    # K_e[a,b] = vol * (M grad phi_b) . grad phi_a
    Mg = np.einsum("eij,ebj->ebi", Mt, grads) if Mt.ndim == 3 else np.einsum("ij,ebj->ebi", Mt, grads)
    Ke = np.einsum("eai,ebi->eab", grads, Mg) * vol[:, None, None]
    ref = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    Me = vol[:, None, None] * ref[None]
    rows = np.repeat(cells, d + 1, axis=1).ravel()
    cols = np.tile(cells, (1, d + 1)).ravel()
    mass = sp.coo_matrix((Me.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    stiff = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    # identical sparsity (explicit zeros of K are kept by construction: same (row, col) set)
    mass.sort_indices()
    stiff.sort_indices()
    assert np.array_equal(mass.indptr, stiff.indptr) and np.array_equal(mass.indices, stiff.indices)
    return mass, stiff


def cells_all_vertices(pts: np.ndarray, cells: np.ndarray, predicate) -> np.ndarray:
    """dolfinx.mesh.locate_entities semantics: an entity is selected when the predicate holds at ALL of
    its vertices (demos/niederer_benchmark.py:148-156)."""
    ok = predicate(pts.T)
    return np.nonzero(ok[cells].all(axis=1))[0]


def load_vector_cells(pts: np.ndarray, cells: np.ndarray, cell_ids=None) -> np.ndarray:
    """s_i = int phi_i dx over the given cells (all cells when None): constant-in-space stimulus."""
    n = pts.shape[0]
    d = cells.shape[1] - 1
    vol, _ = _cell_geometry(pts, cells)
    if cell_ids is None:
        cell_ids = np.arange(cells.shape[0])
    out = np.zeros(n)
    np.add.at(out, cells[cell_ids].ravel(), np.repeat(vol[cell_ids] / (d + 1), d + 1))
    return out


# Quadrature exact to the stated degree on the reference simplex (points in barycentric coordinates)
def _simplex_rule(d: int, degree: int):
    if d == 1:
        m = degree // 2 + 1
        x, w = np.polynomial.legendre.leggauss(m)
        x = 0.5 * (x + 1.0)
        return np.stack([1.0 - x, x], axis=1), 0.5 * w
    # collapsed (Duffy) Gauss-Jacobi-free construction: tensor Gauss-Legendre on the cube mapped to the
    # simplex, exact for polynomials of the requested degree when m = degree//2 + 1 + (d-1)//1 points
    m = degree // 2 + d
    x, w = np.polynomial.legendre.leggauss(m)
    x = 0.5 * (x + 1.0)
    w = 0.5 * w
    if d == 2:
        U, V = np.meshgrid(x, x, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        l1 = U.ravel()
        l2 = (V * (1.0 - U)).ravel()
        wt = (WU * WV * (1.0 - U)).ravel()
        bary = np.stack([1.0 - l1 - l2, l1, l2], axis=1)
        return bary, wt  # weights sum to 1/2
    if d == 3:
        U, V, W = np.meshgrid(x, x, x, indexing="ij")
        WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
        l1 = U.ravel()
        l2 = (V * (1.0 - U)).ravel()
        l3 = (W * (1.0 - U) * (1.0 - V)).ravel()
        wt = (WU * WV * WW * (1.0 - U) ** 2 * (1.0 - V)).ravel()
        bary = np.stack([1.0 - l1 - l2 - l3, l1, l2, l3], axis=1)
        return bary, wt  # weights sum to 1/6
    raise ValueError(d)


def load_vector_function(pts: np.ndarray, cells: np.ndarray, g, degree: int = 6) -> np.ndarray:
    """s_i = int g(x) phi_i dx by quadrature (spatially varying source, e.g. the MMS tests
    tests/test_monodomain.py:13-36); g takes an array of shape (gdim, npoints)."""
    n = pts.shape[0]
    d = cells.shape[1] - 1
    vol, _ = _cell_geometry(pts, cells)
    bary, wt = _simplex_rule(d, degree)
    wt = wt * math.factorial(d)  # normalise to the cell volume
    out = np.zeros(n)
    X = pts[cells]  # (ncell, d+1, gdim)
    for q in range(bary.shape[0]):
        xq = np.einsum("a,eag->eg", bary[q], X)
        gq = np.asarray(g(xq.T), dtype=float) * wt[q] * vol
        for a in range(d + 1):
            np.add.at(out, cells[:, a], gq * bary[q, a])
    return out


def l2_error(pts: np.ndarray, cells: np.ndarray, uh: np.ndarray, exact, degree: int = 8) -> float:
    """|| u_h - u ||_L2 with u_h P1 (the norm the reference's MMS tests assert on)."""
    d = cells.shape[1] - 1
    vol, _ = _cell_geometry(pts, cells)
    bary, wt = _simplex_rule(d, degree)
    wt = wt * math.factorial(d)
    X = pts[cells]
    U = uh[cells]
    tot = 0.0
    for q in range(bary.shape[0]):
        xq = np.einsum("a,eag->eg", bary[q], X)
        uq = U @ bary[q]
        tot += float(np.sum((uq - exact(xq.T)) ** 2 * wt[q] * vol))
    return math.sqrt(tot)


def boundary_facets(cells: np.ndarray) -> np.ndarray:
    """Facets (sorted vertex tuples) that belong to exactly one cell."""
    d = cells.shape[1] - 1
    facets = np.concatenate([np.delete(cells, a, axis=1) for a in range(d + 1)], axis=0)
    facets = np.sort(facets, axis=1)
    uniq, counts = np.unique(facets, axis=0, return_counts=True)
    return uniq[counts == 1]


def load_vector_facets(pts: np.ndarray, facets: np.ndarray) -> np.ndarray:
    """s_i = int phi_i ds over the given facets (constant surface stimulus, demos/lv_endocardial.py:260-270)."""
    n = pts.shape[0]
    k = facets.shape[1]  # vertices per facet
    if k == 1:
        meas = np.ones(facets.shape[0])
    elif k == 2:
        meas = np.linalg.norm(pts[facets[:, 1]] - pts[facets[:, 0]], axis=1)
    else:
        a = pts[facets[:, 1]] - pts[facets[:, 0]]
        b = pts[facets[:, 2]] - pts[facets[:, 0]]
        meas = 0.5 * np.linalg.norm(np.cross(a, b), axis=1)
    out = np.zeros(n)
    np.add.at(out, facets.ravel(), np.repeat(meas / k, k))
    return out
