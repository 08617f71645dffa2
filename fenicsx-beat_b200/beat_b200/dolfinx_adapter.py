"""Glue between REAL dolfinx objects and the C ABI (include/mono_abi.h): what a maintainer of the reference would call from
`BaseModel._setup` (src/beat/base_model.py:100-128) and `DolfinODESolver.__post_init__` (src/beat/odesolver.py:148-153)
instead of creating PETSc objects.  INTEGRATION.md walks through the same calls.

STATUS: the functions that touch dolfinx / ufl (`csr_from_forms`, `stimulus_load`, `build_context`) are UNTESTED here - those
packages cannot be installed in this environment; they are written against the dolfinx 0.9 Python API
and import it lazily.  The array logic they rely on - grouping ghosts by owner, the send lists that mirror the neighbours'
ghost blocks, the structured-numbering detection - is plain NumPy and IS tested (tests/test_dolfinx_adapter.py), and so is
the collective `index_map_arrays`, over three gloo processes with an mpi4py-shaped communicator and IndexMap stand-in.
"""

from __future__ import annotations

import numpy as np

from . import _lib

MONO_PC = {"none": 0, "jacobi": 1, "chebyshev": 2}


# ------------------------------------------------------------------------------------------ plain NumPy (tested)
def group_ghosts_by_owner(ghost_globals: np.ndarray, ghost_owners: np.ndarray):
    """mono_set_halo wants the ghost block grouped by owner rank (ascending) and, inside a group, sorted by global index;
    dolfinx does not promise either.  Returns (perm, nbr_ranks, recv_ptr): new ghost k = old ghost perm[k]."""
    ghost_globals = np.asarray(ghost_globals, dtype=np.int64)
    ghost_owners = np.asarray(ghost_owners, dtype=np.int32)
    perm = np.lexsort((ghost_globals, ghost_owners))
    nbr = np.unique(ghost_owners)
    recv_ptr = np.searchsorted(ghost_owners[perm], np.append(nbr, np.iinfo(np.int32).max)).astype(np.int32)
    recv_ptr[-1] = ghost_globals.size
    return perm, nbr.astype(np.int32), recv_ptr


def permute_ghost_columns(n_owned: int, indices: np.ndarray, perm: np.ndarray) -> np.ndarray:
    """Column indices of a local CSR after the ghost block was reordered with `perm` (new ghost k = old ghost perm[k])."""
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    out = np.asarray(indices, dtype=np.int32).copy()
    g = out >= n_owned
    out[g] = n_owned + inv[out[g] - n_owned]
    return out


def send_lists(shared_local: list[np.ndarray], shared_global: list[np.ndarray]):
    """Per neighbour: the owned dofs it holds as ghosts, sorted by GLOBAL index - the order in which that neighbour's ghost
    block (sorted by global index, see group_ghosts_by_owner) expects them.  Returns (send_ptr, send_idx)."""
    ptr, idx = [0], []
    for loc, glob in zip(shared_local, shared_global):
        order = np.argsort(np.asarray(glob, dtype=np.int64), kind="stable")
        idx.append(np.asarray(loc, dtype=np.int32)[order])
        ptr.append(ptr[-1] + len(loc))
    return np.asarray(ptr, dtype=np.int32), (np.concatenate(idx) if idx else np.zeros(0, np.int32))


def lexicographic_order(x_owned: np.ndarray, decimals: int = 9):
    """Detects dofs that sit on a tensor-product grid and returns the permutation that numbers them lexicographically, first
    axis fastest (x, then y, then z), or None when the points are not a full grid.  dolfinx reorders dofs (reverse Cuthill-McKee) even on `create_box` meshes; the stencil dictionary and the
    shared-memory rings of the streaming PDE kernel need rows whose column OFFSETS repeat, i.e. a
    structured numbering.  An adapter may apply this permutation to rows/columns at set-up and to V at the get/set boundary.
    Returns (perm, shape): new row k = old row perm[k]; shape = grid points per axis."""
    x = np.round(np.asarray(x_owned, dtype=np.float64), decimals)
    axes = [np.unique(x[:, k]) for k in range(x.shape[1])]
    shape = tuple(len(a) for a in axes)
    if int(np.prod(shape)) != x.shape[0]:
        return None
    ijk = [np.searchsorted(a, x[:, k]) for k, a in enumerate(axes)]
    key = np.zeros(x.shape[0], dtype=np.int64)
    stride = 1
    for k in range(x.shape[1]):
        key += ijk[k].astype(np.int64) * stride
        stride *= shape[k]
    if np.unique(key).size != x.shape[0]:
        return None
    return np.argsort(key, kind="stable"), shape


def permute_csr(indptr, indices, data_list, perm_rows: np.ndarray, n_owned: int):
    """CSR of the owned rows renumbered with new row k = old row perm_rows[k] (owned columns renumbered alike, ghost
    columns untouched), columns sorted inside every row."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    inv = np.empty(n_owned, dtype=np.int64)
    inv[perm_rows] = np.arange(n_owned)
    counts = np.diff(indptr)[perm_rows]
    new_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    take = np.concatenate([np.arange(indptr[r], indptr[r + 1]) for r in perm_rows]) if n_owned else np.zeros(0, np.int64)
    cols = indices[take]
    own = cols < n_owned
    cols[own] = inv[cols[own]]
    rows = np.repeat(np.arange(n_owned), counts)
    order = np.lexsort((cols, rows))
    return new_ptr, cols[order].astype(np.int32), [np.asarray(d)[take][order] for d in data_list]


# ------------------------------------------------------------------------------------------ dolfinx-facing (untested here)
def csr_from_forms(V, M, dx):
    """Mass and stiffness of the owned rows as one CSR pair (what mono_pde_set_matrices takes): the two bilinear forms of
    src/beat/monodomain_model.py:83-96 without their constant factors, assembled once by dolfinx."""
    import dolfinx
    import ufl

    imap = V.dofmap.index_map
    n_owned = imap.size_local
    v, w = ufl.TrialFunction(V), ufl.TestFunction(V)
    mass = dolfinx.fem.assemble_matrix(dolfinx.fem.form(v * w * dx))
    stiff = dolfinx.fem.assemble_matrix(dolfinx.fem.form(ufl.inner(M * ufl.grad(v), ufl.grad(w)) * dx))
    mass.scatter_reverse()
    stiff.scatter_reverse()
    indptr = np.asarray(mass.indptr[: n_owned + 1], dtype=np.int64)
    nnz = int(indptr[-1])
    indices = np.asarray(mass.indices[:nnz], dtype=np.int32)
    assert np.array_equal(indices, np.asarray(stiff.indices[:nnz], dtype=np.int32)), "mass and stiffness must share one sparsity"
    return indptr, indices, np.asarray(mass.data[:nnz], dtype=np.float64), np.asarray(stiff.data[:nnz], dtype=np.float64)


def stimulus_load(V, stimulus) -> np.ndarray:
    """int phi_i dz(marker) over the marked entities for the owned dofs (stimulation.py:264-272: the spatial factor of I_s)."""
    import dolfinx
    import ufl

    w = ufl.TestFunction(V)
    vec = dolfinx.fem.assemble_vector(dolfinx.fem.form(w * stimulus.dz))
    vec.scatter_reverse(dolfinx.la.InsertMode.add)
    return np.asarray(vec.array[: V.dofmap.index_map.size_local], dtype=np.float64)


def index_map_arrays(imap, comm):
    """(perm, nbr_ranks, send_ptr, send_idx, recv_ptr) for mono_set_halo from a dolfinx IndexMap.  Collective."""
    ghosts, owners = np.asarray(imap.ghosts, dtype=np.int64), np.asarray(imap.owners, dtype=np.int32)
    perm, nbr, recv_ptr = group_ghosts_by_owner(ghosts, owners)
    lo, _ = imap.local_range
    # every rank tells each neighbour which of ITS dofs it holds as ghosts (global indices); the owner turns them local
    wanted = {int(q): ghosts[perm][recv_ptr[k]: recv_ptr[k + 1]] for k, q in enumerate(nbr)}
    incoming = comm.alltoall([wanted.get(q, np.zeros(0, np.int64)) for q in range(comm.size)])
    dest = [q for q in range(comm.size) if len(incoming[q])]
    assert dest == [int(q) for q in nbr], "the halo pattern of a symmetric operator has the same neighbours both ways"
    send_ptr, send_idx = send_lists([np.asarray(incoming[q]) - lo for q in dest], [incoming[q] for q in dest])
    return perm, nbr, send_ptr, send_idx, recv_ptr


def build_context(V, M, dx, C_m: float, theta: float, stimuli=(), petsc_options: dict | None = None, device: int | None = None):
    """One call that does what section 1 of INTEGRATION.md spells out: returns (ctx, ghost_perm, stimulus_ids)."""
    mesh = V.mesh
    comm = mesh.comm
    ctx = _lib.Context(device if device is not None else comm.rank)  # one rank per GPU of the node
    imap = V.dofmap.index_map
    n_owned, n_ghost = imap.size_local, imap.num_ghosts
    indptr, indices, mass, stiff = csr_from_forms(V, M, dx)
    perm = np.arange(n_ghost)
    halo = None
    if comm.size > 1:
        perm, nbr, send_ptr, send_idx, recv_ptr = index_map_arrays(imap, comm)
        indices = permute_ghost_columns(n_owned, indices, perm)
        halo = (nbr, send_ptr, send_idx, recv_ptr)
    ctx.pde_set_matrices(n_owned, n_ghost, indptr, indices, mass, stiff)
    opts = petsc_options or {}
    ctx.pde_config(float(C_m), float(theta), float(opts.get("ksp_rtol", 1e-5)), float(opts.get("ksp_atol", 1e-50)),
                   int(opts.get("ksp_max_it", 10000)), MONO_PC.get(str(opts.get("pc_type", "jacobi")), 1), 0, 0)
    if halo is not None:
        uid = comm.bcast(ctx.comm_unique_id() if comm.rank == 0 else None, root=0)
        ctx.comm_init(comm.size, comm.rank, uid)
        ctx.set_halo(*halo)
    ids = []
    for s, (t0, t1, amp) in stimuli:  # (beat.Stimulus, (start, end, amplitude)) pairs
        load = stimulus_load(V, s)
        idx = np.nonzero(load)[0].astype(np.int32)
        ids.append(ctx.stim_add(idx, load[idx], t0, t1, amp))
    return ctx, perm, ids
