"""The NumPy half of beat_b200/dolfinx_adapter.py (the dolfinx-facing half cannot run here): ghost grouping and send lists
against the package's own partitioner, and the structured-numbering detection that makes the stencil dictionary apply to a
mesh whose dofs were reordered (dolfinx applies reverse Cuthill-McKee even to create_box meshes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))
from beat_b200 import dolfinx_adapter as da  # noqa: E402
from beat_b200 import fem  # noqa: E402
from beat_b200._lib import csr_row_patterns  # noqa: E402


def test_ghost_grouping_and_send_lists_reproduce_the_partitioner():
    size = 3
    meshes = [fem.create_box(fem.Comm(r, size), [np.zeros(3), np.array([6.0, 2.0, 1.0])], [12, 4, 2]) for r in range(size)]
    rng = np.random.default_rng(5)
    for r, mesh in enumerate(meshes):
        im = mesh.index_map
        shuffle = rng.permutation(im.num_ghosts)  # dolfinx promises no particular ghost order
        perm, nbr, recv_ptr = da.group_ghosts_by_owner(im.ghosts[shuffle], im.owners[shuffle])
        assert np.array_equal(im.ghosts[shuffle][perm], im.ghosts) and np.array_equal(nbr, im.nbr_ranks)
        assert np.array_equal(recv_ptr, im.recv_ptr)
        # what the neighbours hold of this rank, as (local, global) pairs in arbitrary order
        loc, glob = [], []
        for q in im.nbr_ranks:
            imq = meshes[q].index_map
            g = imq.ghosts[imq.owners == r]
            g = g[rng.permutation(g.size)]
            l2g_owned = im.local_to_global[: im.size_local]
            loc.append(np.searchsorted(l2g_owned, g))  # owned dofs are sorted by global index in this partitioner
            assert np.array_equal(l2g_owned[loc[-1]], g)
            glob.append(g)
        send_ptr, send_idx = da.send_lists(loc, glob)
        assert np.array_equal(send_ptr, im.send_ptr) and np.array_equal(send_idx, im.send_idx)
        # permuting ghost columns: the matrix is the same operator on the re-grouped ghost block
        indptr, indices, mass, _ = fem.assemble_p1_local(mesh, 1.0)
        inv = np.empty_like(shuffle)
        inv[shuffle] = np.arange(shuffle.size)
        shuffled_cols = indices.copy()
        gh = indices >= im.size_local
        shuffled_cols[gh] = im.size_local + inv[indices[gh] - im.size_local]  # columns as the shuffled ghost order names them
        assert np.array_equal(da.permute_ghost_columns(im.size_local, shuffled_cols, perm), indices)


def test_lexicographic_order_restores_the_dictionary_on_reordered_dofs():
    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.array([2.0, 1.5, 1.0])], [8, 6, 4])
    indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, np.diag([1.0, 0.3, 0.2]))
    n = indptr.size - 1
    scr = np.random.default_rng(2).permutation(n)  # what a bandwidth-reducing reordering does to the structured numbering
    inv = np.empty(n, dtype=np.int64)
    inv[scr] = np.arange(n)  # new row k = old row scr[k]; old row r becomes row inv[r]
    rows = np.repeat(np.arange(n), np.diff(indptr))
    coo_r, coo_c = inv[rows], inv[indices]
    order = np.lexsort((coo_c, coo_r))  # (explicit zeros of the stiffness matrix stay: one sparsity for both matrices)

    class _M:
        pass

    As, Ks = _M(), _M()
    As.indptr = Ks.indptr = np.concatenate([[0], np.cumsum(np.bincount(coo_r, minlength=n))]).astype(np.int64)
    As.indices = Ks.indices = coo_c[order].astype(np.int32)
    As.data, Ks.data = mass[order], stiff[order]
    xs = mesh.geometry.x[scr]
    pat, rep, cnt = csr_row_patterns(As.indptr, As.indices, As.data, Ks.data)
    assert cnt.max() <= 2  # scrambled: (almost) no two rows share their column offsets
    perm, shape = da.lexicographic_order(xs)
    assert shape == (9, 7, 5)
    ip, ix, (m2, k2) = da.permute_csr(As.indptr, As.indices, [As.data, Ks.data], perm, n)
    assert np.array_equal(ip, indptr) and np.array_equal(ix, indices) and np.allclose(m2, mass) and np.allclose(k2, stiff)
    pat2, rep2, cnt2 = csr_row_patterns(ip, ix, m2, k2)
    assert cnt2.sum() == n and len(cnt2) == 27  # every row is one of the 27 box stencils again
    assert da.lexicographic_order(xs[:-1]) is None  # not a full grid


# ---- the collective half: index_map_arrays over real processes (gloo), with an mpi4py-shaped communicator -------------
def _adapter_worker(rank, world, store_path, result_path):
    """An interval mesh: its x-partition owns CONTIGUOUS global ranges, as dolfinx IndexMaps do (the adapter turns a global
    index into a local one with `global - local_range[0]`)."""
    import sys as _sys

    for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200")):
        if p not in _sys.path:
            _sys.path.insert(0, p)
    import torch.distributed as dist

    from beat_b200 import dolfinx_adapter as da2
    from beat_b200 import fem as fem2

    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world)

    class Comm:  # the mpi4py surface the adapter uses
        size = world

        def __init__(self):
            self.rank = rank

        def alltoall(self, send):  # (gloo has no object all-to-all: gather every rank's row, pick our column)
            table = [None] * world
            dist.all_gather_object(table, [np.asarray(s) for s in send])
            return [table[q][rank] for q in range(world)]

    mesh = fem2.create_interval(fem2.Comm(rank, world), 17)
    im = mesh.index_map
    owned_g = im.local_to_global[: im.size_local]
    assert np.array_equal(owned_g, np.arange(owned_g[0], owned_g[0] + im.size_local))
    shuffle = np.random.default_rng(100 + rank).permutation(im.num_ghosts)  # dolfinx promises no ghost order

    class IMap:  # the dolfinx IndexMap attributes the adapter reads
        ghosts, owners = im.ghosts[shuffle], im.owners[shuffle]
        size_local = im.size_local
        local_range = (int(owned_g[0]), int(owned_g[-1]) + 1)

    perm, nbr, send_ptr, send_idx, recv_ptr = da2.index_map_arrays(IMap, Comm())
    ok = (np.array_equal(IMap.ghosts[perm], im.ghosts) and np.array_equal(nbr, im.nbr_ranks) and np.array_equal(recv_ptr, im.recv_ptr)
          and np.array_equal(send_ptr, im.send_ptr) and np.array_equal(send_idx, im.send_idx))
    np.save(f"{result_path}.{rank}.npy", np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_index_map_arrays_over_three_gloo_ranks():
    """dolfinx_adapter.index_map_arrays (collective) reproduces, from shuffled ghosts and an all-to-all of the wanted global
    indices, the halo arrays the package's own partitioner builds."""
    import tempfile

    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as tmp:
        store, res = os.path.join(tmp, "store"), os.path.join(tmp, "res")
        mp.spawn(_adapter_worker, args=(3, store, res), nprocs=3, join=True)
        for rank in range(3):
            assert bool(np.load(f"{res}.{rank}.npy")[0]), rank
