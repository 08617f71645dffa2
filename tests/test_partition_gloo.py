"""Host side of the multi-rank path on CPU: two `gloo` processes build their x-slab of a box, and the halo
lists / local matrices / ghost layout they would hand to mono_set_halo and mono_pde_set_matrices are checked by
running the SAME protocol the device uses - "the i-th value sent to a neighbour lands in the i-th ghost that
neighbour holds from the sender" for the CG vector, an all-reduce for the dot products - in NumPy over gloo:
  * send/recv counts and global ids agree between neighbours,
  * the distributed SpMV of owned rows equals the serial one,
  * a distributed Jacobi-PCG (halo per iteration + all-reduced dots) reproduces the serial solve.
"""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _halo_exchange(dist, torch, imap, vec):
    """ghosts of `vec` (owned+ghost) <- owners' values, through the halo lists (what halo.cu sets up)."""
    n_owned = imap.size_local
    reqs, recv_bufs = [], []
    for k, q in enumerate(imap.nbr_ranks):
        send = torch.from_numpy(np.ascontiguousarray(vec[imap.send_idx[imap.send_ptr[k]:imap.send_ptr[k + 1]]]))
        recv = torch.empty(int(imap.recv_ptr[k + 1] - imap.recv_ptr[k]), dtype=torch.float64)
        recv_bufs.append((k, recv))
        reqs.append(dist.isend(send, int(q)))
        reqs.append(dist.irecv(recv, int(q)))
    for r in reqs:
        r.wait()
    for k, recv in recv_bufs:
        vec[n_owned + imap.recv_ptr[k]: n_owned + imap.recv_ptr[k + 1]] = recv.numpy()


def _worker(rank, world, store_path, n, result_path):
    for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist

    from beat_b200 import fem

    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world)
    box = ([np.zeros(3), np.array([2.0, 1.0, 0.5])], n)
    M = np.diag([9.5e-4, 1.26e-4, 1.26e-4])
    mesh = fem.create_box(fem.Comm(rank, world), *box)
    imap = mesh.index_map
    n_owned, n_local = imap.size_local, imap.size_local + imap.num_ghosts
    indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, M)
    A = sp.csr_matrix((0.01 * mass + 0.005 * stiff, indices, indptr), shape=(n_owned, n_local))

    full = fem.create_box(fem.COMM_SELF, *box)
    ip, ix, ms, st = fem.assemble_p1_local(full, M)
    ng = full.index_map.size_local
    Ag = sp.csr_matrix((0.01 * ms + 0.005 * st, ix, ip), shape=(ng, ng))
    l2g = imap.local_to_global
    errs = {}

    # (1) the neighbour holds exactly the ghosts we send, in the order we send them
    counts = torch.zeros(world, world, dtype=torch.int64)
    for k, q in enumerate(imap.nbr_ranks):
        counts[rank, q] = int(imap.send_ptr[k + 1] - imap.send_ptr[k])
    recvc = torch.zeros(world, world, dtype=torch.int64)
    for k, q in enumerate(imap.nbr_ranks):
        recvc[q, rank] = int(imap.recv_ptr[k + 1] - imap.recv_ptr[k])
    dist.all_reduce(counts)
    dist.all_reduce(recvc)
    errs["counts_match"] = bool(torch.equal(counts, recvc)) and int(counts.sum()) > 0
    gid = l2g.astype(np.float64).copy()
    gid[n_owned:] = -1.0
    _halo_exchange(dist, torch, imap, gid)
    errs["ghost_ids_match"] = bool(np.array_equal(gid, l2g.astype(np.float64)))
    errs["owned_cover"] = int(n_owned)

    # (2) distributed SpMV == serial SpMV on the owned rows
    rng = np.random.default_rng(5)
    xg = rng.standard_normal(ng)
    x = np.zeros(n_local)
    x[:n_owned] = xg[l2g[:n_owned]]
    _halo_exchange(dist, torch, imap, x)
    errs["spmv"] = float(np.abs(A @ x - (Ag @ xg)[l2g[:n_owned]]).max())

    # (3) distributed Jacobi-PCG with all-reduced dots (the reductions the device does over peer memory)
    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t)

    bg = Ag @ xg
    b = bg[l2g[:n_owned]]
    dinv = 1.0 / A[np.arange(n_owned), np.arange(n_owned)].A1
    sol = np.zeros(n_local)
    r = b.copy()
    z = dinv * r
    p = np.zeros(n_local)
    p[:n_owned] = z
    rz = allsum(r @ z)
    bnorm = np.sqrt(allsum((dinv * b) @ (dinv * b)))
    for it in range(500):
        _halo_exchange(dist, torch, imap, p)
        q = A @ p
        alpha = rz / allsum(p[:n_owned] @ q)
        sol[:n_owned] += alpha * p[:n_owned]
        r -= alpha * q
        z = dinv * r
        rz_new = allsum(r @ z)
        if np.sqrt(allsum(z @ z)) <= 1e-12 * bnorm:
            break
        p[:n_owned] = z + (rz_new / rz) * p[:n_owned]
        rz = rz_new
    errs["pcg"] = float(np.abs(sol[:n_owned] - xg[l2g[:n_owned]]).max())
    errs["pcg_its"] = it + 1
    np.save(f"{result_path}.{rank}.npy", np.array([errs], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [(8, 3, 2), (5, 4, 3)])
def test_two_rank_partition_protocol(n):
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as tmp:
        store, res = os.path.join(tmp, "store"), os.path.join(tmp, "res")
        mp.spawn(_worker, args=(2, store, list(n), res), nprocs=2, join=True)
        owned = 0
        for rank in range(2):
            e = np.load(f"{res}.{rank}.npy", allow_pickle=True)[0]
            assert e["counts_match"] and e["ghost_ids_match"], e
            assert e["spmv"] <= 1e-15, e
            assert e["pcg"] <= 1e-9 and e["pcg_its"] < 200, e
            owned += e["owned_cover"]
        assert owned == (n[0] + 1) * (n[1] + 1) * (n[2] + 1)  # every dof owned exactly once
