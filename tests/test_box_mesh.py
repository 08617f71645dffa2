"""BoxMesh (index arithmetic, no cell array, stencil assembly) == the generic builder (global cell array, np.unique,
sort-based assembly) on small boxes, for 1, 2 and 3 ranks: numbering, halo lists, cells, matrices, stimulus."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))
from beat_b200 import fem  # noqa: E402

M = np.array([[9.5e-4, 1e-5, 0.0], [1e-5, 1.26e-4, 2e-5], [0.0, 2e-5, 3.0e-4]])


@pytest.mark.parametrize("n", [(4, 3, 2), (7, 2, 3), (3, 3, 3)])
@pytest.mark.parametrize("size", [1, 2, 3])
def test_box_mesh_matches_generic_builder(n, size):
    pts = [np.zeros(3), np.array([2.0, 1.5, 0.5])]
    for rank in range(size):
        comm = fem.Comm(rank, size)
        a, b = fem.create_box(comm, pts, n), fem._create_box_generic(comm, pts, n)
        ia, ib = a.index_map, b.index_map
        assert ia.size_local == ib.size_local and ia.size_global == ib.size_global
        for name in ("ghosts", "owners", "local_to_global", "nbr_ranks", "send_ptr", "send_idx", "recv_ptr"):
            assert np.array_equal(getattr(ia, name), getattr(ib, name)), name
        assert np.allclose(a.geometry.x, b.geometry.x, atol=1e-15)
        assert np.array_equal(a.cells, b.cells)
        assert np.array_equal(a.cells_of(np.array([0, 5, a.num_cells - 1])), b.cells[[0, 5, b.num_cells - 1]])
        for Mv in (M, 0.37):
            pa = fem.assemble_p1_local(a, Mv)
            pb = fem.assemble_p1_local(b, Mv)
            assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
            assert np.abs(pa[2] - pb[2]).max() <= 1e-15 * np.abs(pb[2]).max()
            assert np.abs(pa[3] - pb[3]).max() <= 1e-13 * np.abs(pb[3]).max()
        marker = lambda x: (x[0] <= 1.0 + 1e-10) & (x[1] <= 1.0 + 1e-10)  # noqa: E731
        ca, cb = fem.locate_entities(a, 3, marker), fem.locate_entities(b, 3, marker)
        assert np.array_equal(ca, cb)
        ta, tb = fem.meshtags(a, 3, ca, 1), fem.meshtags(b, 3, cb, 1)
        la = fem.load_vector(a, fem.Measure("dx", domain=a, subdomain_data=ta), 1)
        lb = fem.load_vector(b, fem.Measure("dx", domain=b, subdomain_data=tb), 1)
        assert np.allclose(la, lb, rtol=0, atol=1e-16)


def test_lv_ellipsoid_mesh_and_partition():
    """Synthetic LV shell: positive cell volumes summing to the shell volume independent of the partition, every dof
    owned exactly once, periodic neighbours, halo lists consistent (the i-th dof sent = the i-th ghost held), ENDO facets
    of the owned part tile the endocardium."""
    from beat_b200 import geometry

    n = (3, 10, 12)
    full = geometry.get_lv_ellipsoid_geometry(fem.COMM_SELF, *n)
    x = full.mesh.geometry.x[full.mesh.cells]
    vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])) / 6
    assert (vol > 0).all()
    ip, ix, ms, st = fem.assemble_p1_local(full.mesh, 1.0)
    assert np.isclose(ms.sum(), vol.sum())            # sum of the mass matrix = volume
    assert np.abs(np.add.reduceat(st, ip[:-1])).max() < 1e-9 * np.abs(st).max()  # stiffness rows sum to zero
    endo_area = fem.load_vector(full.mesh, fem.Measure("ds", domain=full.mesh, subdomain_data=full.ffun), 1).sum()
    for size in (2, 3):
        geos = [geometry.get_lv_ellipsoid_geometry(fem.Comm(r, size), *n) for r in range(size)]
        maps = [g.mesh.index_map for g in geos]
        owned = np.concatenate([m.local_to_global[: m.size_local] for m in maps])
        assert np.array_equal(np.sort(owned), np.arange(full.mesh.index_map.size_global))
        area = 0.0
        for r, (g, m) in enumerate(zip(geos, maps)):
            assert set(m.nbr_ranks.tolist()) == {(r - 1) % size, (r + 1) % size}
            for k, q in enumerate(m.nbr_ranks):
                mq = maps[q]
                kq = list(mq.nbr_ranks).index(r)
                sent = m.local_to_global[m.send_idx[m.send_ptr[k]: m.send_ptr[k + 1]]]
                held = mq.ghosts[mq.recv_ptr[kq]: mq.recv_ptr[kq + 1]]
                assert np.array_equal(sent, held)
            area += fem.load_vector(g.mesh, fem.Measure("ds", domain=g.mesh, subdomain_data=g.ffun), 1).sum()
        assert np.isclose(area, endo_area)


@pytest.mark.parametrize("n", [(3, 6, 12), (2, 5, 9)])
@pytest.mark.parametrize("size", [1, 2, 3, 4])
def test_shell_mesh_matches_generic_builder(n, size):
    """ShellMesh (periodic index arithmetic + stencil assembly with varying element matrices) == the generic builder
    (cell array, np.unique, scipy accumulation) for the LV shell: numbering, halo lists, cells, matrices with a
    cell-wise tensor, facet loads."""
    rng = np.random.default_rng(0)
    for rank in range(size):
        comm = fem.Comm(rank, size)
        a, b = fem.create_lv_ellipsoid(comm, *n), fem._create_lv_ellipsoid_generic(comm, *n)
        ia, ib = a.index_map, b.index_map
        assert ia.size_local == ib.size_local and ia.size_global == ib.size_global
        for name in ("ghosts", "owners", "local_to_global", "nbr_ranks", "send_ptr", "send_idx", "recv_ptr"):
            assert np.array_equal(getattr(ia, name), getattr(ib, name)), (name, rank, size)
        assert np.allclose(a.geometry.x, b.geometry.x, atol=1e-15)
        assert np.array_equal(a.cells, b.cells)
        A = rng.standard_normal((a.num_cells, 3, 3))
        Mcell = A @ A.transpose(0, 2, 1) + 0.1 * np.eye(3)  # SPD per cell
        for Mv in (Mcell, np.diag([3.0, 1.0, 0.5]), 0.7):
            pb = fem._assemble_p1_numpy(b, Mv)
            for pa in (fem.assemble_p1_local(a, Mv), fem.assemble_p1_structured(a, np.asarray(Mv))):  # library threads / NumPy stencils
                assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
                assert np.abs(pa[2] - pb[2]).max() <= 1e-14 * np.abs(pb[2]).max()
                assert np.abs(pa[3] - pb[3]).max() <= 1e-12 * np.abs(pb[3]).max()
        fa, fb = a.boundary_facets(), b.boundary_facets()
        assert np.array_equal(fa, fb)
        marker = lambda x: x[0] > 2.0  # noqa: E731
        assert np.array_equal(fem.locate_entities(a, 3, marker), fem.locate_entities(b, 3, marker))


@pytest.mark.parametrize("size", [2, 3, 8])
def test_halo_lists_satisfy_the_abi_preconditions(size):
    """What mono_set_halo validates (csrc/halo.cu): pointers start at 0, are non-decreasing, cover exactly the ghost block;
    send indices are owned dofs; neighbours are other ranks; everything int32 - for every mesh builder."""
    for rank in range(size):
        comm = fem.Comm(rank, size)
        box = [np.zeros(3), np.array([20.0, 7.0, 3.0])]
        meshes = {"box": fem.create_box(comm, box, [40, 14, 6]), "box_generic": fem._create_box_generic(comm, box, [16, 6, 4]),
                  "shell": fem.create_lv_ellipsoid(comm, 2, 6, 24), "rectangle": fem.create_rectangle(comm, ((0, 0), (1, 1)), (16, 8)),
                  "interval": fem.create_interval(comm, 64, (0.0, 1.0))}
        for what, mesh in meshes.items():
            im = mesh.index_map
            sp, rp, nbr, si = im.send_ptr, im.recv_ptr, im.nbr_ranks, im.send_idx
            assert len(sp) == len(nbr) + 1 == len(rp), what
            assert sp[0] == 0 and rp[0] == 0 and (np.diff(sp) >= 0).all() and (np.diff(rp) >= 0).all(), what
            assert rp[-1] == im.num_ghosts and len(si) == sp[-1], what
            assert (si >= 0).all() and (si < im.size_local).all(), what
            assert all(0 <= q < size and q != rank for q in nbr), what
            assert all(a.dtype == np.int32 for a in (sp, rp, nbr, si)), what


@pytest.mark.parametrize("size", [1, 3])
def test_point_probe_by_index_arithmetic_matches_the_cell_search(size):
    """fem.point_probe on a BoxMesh looks only at the cubes around the point; same cell, same weights as the search over
    all cells of the generic mesh - for the Niederer probe points, points on faces / edges, random and outside points."""
    rng = np.random.default_rng(0)
    pts = [(0, 0, 0), (0, 7, 0), (20, 0, 0), (20, 7, 0), (0, 0, 3), (0, 7, 3), (20, 0, 3), (20, 7, 3), (10, 3.5, 1.5), (9.99999, 3.5, 1.5),
           (10.0, 3.4, 1.5), (5.0, 0.0, 1.25), (25, 1, 1), (-0.1, 0, 0)] + [tuple(rng.random(3) * [20, 7, 3]) for _ in range(40)]
    box = [np.zeros(3), np.array([20.0, 7.0, 3.0])]
    for rank in range(size):
        comm = fem.Comm(rank, size)
        a, b = fem.create_box(comm, box, [40, 14, 6]), fem._create_box_generic(comm, box, [40, 14, 6])
        for pt in pts:
            ha, hb = fem.point_probe(a, pt), fem.point_probe(b, pt)
            assert (ha is None) == (hb is None), (pt, rank)
            if ha is not None:
                assert np.array_equal(ha[0], hb[0]) and np.allclose(ha[1], hb[1], atol=1e-12), (pt, rank)
                assert np.allclose(a.geometry.x[ha[0]].T @ ha[1], pt, atol=1e-9)
