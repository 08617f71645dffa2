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
