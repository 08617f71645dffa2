"""mono_fem_assemble_p1 (host threads inside the C-ABI library, no GPU): the owned rows of the P1 mass / stiffness
matrices against the NumPy/SciPy restatement in beat_b200.fem and against the oracle's assembly, in 1-3 dimensions,
for scalar / tensor / cell-wise conductivities and for every rank of a partition."""

import numpy as np
import pytest

from beat_b200 import fem
from beat_b200._lib import MonoError, fem_assemble_p1, load_library


def _meshes(comm):
    yield fem.create_interval(comm, 37, (0.0, 2.0))
    yield fem.create_rectangle(comm, ((0, 0), (1.5, 1.0)), (7, 5))
    yield fem._create_box_generic(comm, ((0, 0, 0), (1, 2, 1.5)), (5, 4, 6))
    yield fem._create_lv_ellipsoid_generic(comm, 2, 6, 9)


@pytest.mark.parametrize("size", [1, 3])
def test_library_assembler_matches_numpy_restatement(size, monkeypatch):
    rng = np.random.default_rng(0)
    for threads in ("1", "5"):
        monkeypatch.setenv("MONO_HOST_THREADS", threads)
        for rank in range(size):
            for mesh in _meshes(fem.Comm(rank, size)):
                d = mesh.topology.dim
                A = rng.standard_normal((mesh.num_cells, d, d))
                for M in (0.7, np.eye(d) * 0.3 + 0.05, A @ A.transpose(0, 2, 1) + np.eye(d)):
                    ref = fem._assemble_p1_numpy(mesh, M)
                    out = fem_assemble_p1(d, mesh.index_map.size_local, mesh.cells, mesh.geometry.x, M)
                    assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])
                    for a, b in zip(ref[2:], out[2:]):
                        assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


def test_library_assembler_is_deterministic_and_matches_oracle(monkeypatch):
    from oracle import fem as ofem

    mesh = fem._create_lv_ellipsoid_generic(fem.COMM_SELF, 3, 8, 12)
    rng = np.random.default_rng(1)
    A = rng.standard_normal((mesh.num_cells, 3, 3))
    M = A @ A.transpose(0, 2, 1) + np.eye(3)
    monkeypatch.setenv("MONO_HOST_THREADS", "7")
    runs = [fem_assemble_p1(3, mesh.index_map.size_local, mesh.cells, mesh.geometry.x, M) for _ in range(3)]
    for r in runs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(runs[0], r))  # bit-identical whatever the thread timing
    import scipy.sparse as sp

    n = mesh.num_local_vertices
    indptr, indices, mass, stiff = runs[0]
    omass, ostiff = ofem.assemble_p1(mesh.geometry.x, mesh.cells, M)
    for got, want in ((mass, omass), (stiff, ostiff)):
        G = sp.csr_matrix((got, indices, indptr), shape=(n, n))
        assert abs(G - sp.csr_matrix(want)).max() <= 1e-13 * abs(sp.csr_matrix(want)).max()
    assert np.allclose(sp.csr_matrix((stiff, indices, indptr), shape=(n, n)) @ np.ones(n), 0.0, atol=1e-12)  # constants in the kernel


def test_library_assembler_reports_bad_input():
    lib = load_library()
    x = np.array([[0.0, 0, 0], [1.0, 0, 0], [0.0, 1, 0], [0.0, 0, 1]])
    good = np.array([[0, 1, 2, 3]])
    with pytest.raises(MonoError, match="outside"):
        fem_assemble_p1(3, 4, np.array([[0, 1, 2, 4]]), x, 1.0)
    with pytest.raises(MonoError, match="degenerate"):
        fem_assemble_p1(3, 4, np.array([[0, 1, 2, 2]]), x, 1.0)
    with pytest.raises(ValueError):
        fem_assemble_p1(3, 4, good, x, np.ones((2, 2)))
    with pytest.raises(ValueError):
        fem_assemble_p1(3, 4, good[:, :3], x, 1.0)
    indptr, indices, mass, stiff = fem_assemble_p1(3, 4, good, x, 2.0)
    assert indptr.tolist() == [0, 4, 8, 12, 16] and indices.tolist() == [0, 1, 2, 3] * 4
    assert np.isclose(mass.sum(), 1.0 / 6.0) and np.isclose(stiff[0], 2.0 * 3.0 / 6.0)  # |K| = 1/6, grad(phi_0) = -(1,1,1)
    # an indptr that does not belong to the mesh is refused instead of overrunning the arrays
    import ctypes as C

    bad_ptr = np.array([0, 1, 2, 3, 4], dtype=np.int64)
    buf_i, buf_m, buf_k = np.zeros(16, np.int32), np.zeros(16), np.zeros(16)
    cells = np.ascontiguousarray(good, dtype=np.int64)
    M = np.array([1.0])
    rc = lib.mono_fem_assemble_p1(3, 4, 4, 1, cells.ctypes.data_as(C.POINTER(C.c_int64)), x.ctypes.data_as(C.POINTER(C.c_double)), 3, 0,
                                  M.ctypes.data_as(C.POINTER(C.c_double)), bad_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                  buf_i.ctypes.data_as(C.POINTER(C.c_int32)), buf_m.ctypes.data_as(C.POINTER(C.c_double)),
                                  buf_k.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == -1 and b"indptr" in lib.mono_last_error(None)


@pytest.mark.parametrize("dim", [2, 3])
def test_library_assembler_on_delaunay_meshes(dim):
    """Genuinely unstructured connectivity (SciPy Delaunay of random points, irregular vertex degrees up to ~40): owned
    rows only, cell-wise tensor - against the dense accumulation of the element matrices."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(dim)
    pts = rng.random((300 if dim == 3 else 500, dim))
    tri = Delaunay(pts)
    cells = tri.simplices.astype(np.int64)
    e = pts[cells[:, 1:]] - pts[cells[:, :1]]
    vol = np.abs(np.linalg.det(e))
    cells = cells[vol > 1e-9]  # drop slivers on the hull
    x = np.zeros((pts.shape[0], 3))
    x[:, :dim] = pts
    A = rng.standard_normal((cells.shape[0], dim, dim))
    M = A @ A.transpose(0, 2, 1) + np.eye(dim)
    n_owned = pts.shape[0] * 2 // 3
    indptr, indices, mass, stiff = fem_assemble_p1(dim, n_owned, cells, x, M)
    # dense reference
    n = pts.shape[0]
    Md, Kd, pattern = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n), dtype=bool)
    ref = (1.0 + np.eye(dim + 1)) / ((dim + 1) * (dim + 2))
    fact = 2.0 if dim == 2 else 6.0
    for c, verts in enumerate(cells):
        P_ = np.hstack([np.ones((dim + 1, 1)), pts[verts]])
        G = np.linalg.inv(P_)[1:].T  # rows: gradients of the barycentric coordinates
        v = abs(np.linalg.det(P_)) / fact
        Md[np.ix_(verts, verts)] += v * ref
        Kd[np.ix_(verts, verts)] += v * (G @ M[c] @ G.T)
        pattern[np.ix_(verts, verts)] = True
    import scipy.sparse as sp

    got_m = sp.csr_matrix((mass, indices, indptr), shape=(n_owned, n)).toarray()
    got_k = sp.csr_matrix((stiff, indices, indptr), shape=(n_owned, n)).toarray()
    assert np.array_equal(np.diff(indptr), pattern[:n_owned].sum(axis=1))
    assert all(np.all(np.diff(indices[indptr[r]: indptr[r + 1]]) > 0) for r in range(n_owned))  # sorted, no duplicates
    assert np.abs(got_m - Md[:n_owned]).max() <= 1e-13 * np.abs(Md).max()
    assert np.abs(got_k - Kd[:n_owned]).max() <= 1e-11 * np.abs(Kd).max()
