"""mono_csr_row_patterns (host threads, no GPU): the stencil dictionary of a CSR pair against a plain Python restatement."""

import numpy as np
import pytest

from beat_b200 import fem
from beat_b200._lib import MonoError, csr_row_patterns


def _python_patterns(indptr, indices, mass, stiff, keep):
    seen, ids, first, count = {}, [], [], []
    for r in range(len(indptr) - 1):
        a, b = indptr[r], indptr[r + 1]
        key = (tuple((indices[a:b] - r).tolist()), mass[a:b].tobytes(), stiff[a:b].tobytes())
        p = seen.setdefault(key, len(seen))
        if p == len(first):
            first.append(r)
            count.append(0)
        count[p] += 1
        ids.append(p)
    order = sorted(range(len(first)), key=lambda p: (-count[p], first[p]))[:keep]
    final = {p: k for k, p in enumerate(order)}
    return (np.array([final.get(p, 255) for p in ids], dtype=np.uint8), np.array([first[p] for p in order]), np.array([count[p] for p in order]))


@pytest.mark.parametrize("size,keep", [(1, 64), (3, 64), (3, 5)])
def test_box_stencils(size, keep, monkeypatch):
    M = np.diag([0.1334, 0.0176, 0.0176])
    for rank in range(size):
        mesh = fem.create_box(fem.Comm(rank, size), [np.zeros(3), np.array([20.0, 7.0, 3.0])], [40, 14, 6])
        csr = fem.assemble_p1_local(mesh, M)
        want = _python_patterns(*csr, keep)
        for threads in ("1", "6"):
            monkeypatch.setenv("MONO_HOST_THREADS", threads)
            got = csr_row_patterns(*csr, max_patterns=keep)
            for g, w in zip(got, want):
                assert np.array_equal(g, w)
        if size == 1:
            assert len(want[1]) == 27 and (want[0] != 255).all()  # interior, 6 faces, 12 edges, 8 corners


def test_mapped_mesh_has_no_dictionary_and_bad_input_is_refused():
    mesh = fem.create_lv_ellipsoid(fem.COMM_SELF, 2, 6, 9)
    csr = fem.assemble_p1_local(mesh, 1.0)
    pat, rep, cnt = csr_row_patterns(*csr, max_patterns=32)
    want = _python_patterns(*csr, 32)
    assert np.array_equal(pat, want[0]) and np.array_equal(rep, want[1]) and np.array_equal(cnt, want[2])
    assert (pat != 255).mean() < 0.5  # geometry varies from row to row
    with pytest.raises(MonoError, match="invalid argument"):
        csr_row_patterns(*csr, max_patterns=256)
    bad = csr[0].copy()
    bad[3] = bad[2] - 1
    with pytest.raises(MonoError, match="non-decreasing"):
        csr_row_patterns(bad, *csr[1:])


@pytest.mark.parametrize("size", [2, 4])
def test_dictionary_of_a_rank_of_a_partitioned_slab(size):
    """What pde_build_dictionary keeps on one rank of an x-slab partition (restated: the 32 most frequent stencils, minus the
    ones whose representative row reaches into the ghost block, minus every row with a ghost column): the rows outside the
    dictionary are the rows of the planes next to a cut (plus at most a few once-only corner stencils) - the side table of the
    streaming kernel - and everything else is one of the box's regular stencils."""
    M = np.diag([0.1334, 0.0176, 0.0176])
    n = (24, 6, 5)
    for rank in range(size):
        mesh = fem.create_box(fem.Comm(rank, size), [np.zeros(3), np.array([12.0, 3.0, 2.5])], list(n))
        indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, M)
        n_owned = indptr.size - 1
        pat, rep, cnt = csr_row_patterns(indptr, indices, mass, stiff, max_patterns=32)
        has_ghost = np.array([(indices[indptr[r]:indptr[r + 1]] >= n_owned).any() for r in range(n_owned)])
        keep = [p for p, r in enumerate(rep) if not has_ghost[r]]
        in_dict = np.isin(pat, keep) & ~has_ghost
        # the rows outside: those with a ghost column = owned vertices of the x-planes next to a cut
        x = mesh.geometry.x[:n_owned, 0]
        h = 12.0 / n[0]
        lo, hi = mesh.lo * h, (mesh.hi - 1) * h
        on_cut = np.zeros(n_owned, dtype=bool)
        if mesh.has_left:
            on_cut |= np.isclose(x, lo)
        if mesh.has_right:
            on_cut |= np.isclose(x, hi)
        assert np.array_equal(has_ghost, on_cut)
        # (a stencil that occurs once - a corner of the box - can lose its slot among the 32 to the equally rare stencils of
        # the cut planes: such rows join the side table too, they are a handful)
        assert (~in_dict)[on_cut].all() and (~in_dict & ~on_cut).sum() <= 8, (rank, (~in_dict).sum(), on_cut.sum())
        assert in_dict.mean() > 0.6
        # and no dictionary stencil points outside the owned block
        for p in keep:
            r = rep[p]
            assert (indices[indptr[r]:indptr[r + 1]] < n_owned).all()
