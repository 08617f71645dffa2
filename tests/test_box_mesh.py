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
