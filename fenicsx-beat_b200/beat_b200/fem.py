"""Light host-side stand-ins for the slice of dolfinx / ufl that the monodomain step touches.

In the reference, dolfinx builds the mesh, the P1 dof map, the partition and the constant matrices once
at set-up (north_star: "host code stays Python").  dolfinx cannot be installed in this environment
(SURVEY.md section 0), so this module provides the same *objects with the same attribute names* for the
structured meshes the reference's tests, README and Niederer demo use:

    create_unit_interval / create_unit_square / create_rectangle / create_box      (dolfinx.mesh.*)
    locate_entities / locate_entities_boundary / meshtags                          (dolfinx.mesh.*)
    functionspace / Function / Constant                                             (dolfinx.fem.*)
    Measure, conditional/And/ge/le/..., SpatialCoordinate-free separable sources    (ufl.*)

Partitioning follows dolfinx's IndexMap layout: owned dofs first, ghosts after, grouped by owner
(tests/test_odesolver.py:63 of the reference reads size_local + num_ghosts).  A rank's local cells are
all cells that touch an owned vertex, so owned matrix rows assemble completely without communication.

Mesh splits restated from dolfinx (not verifiable here): unit square = "right" diagonal, box = six Kuhn
tetrahedra per hexahedron sharing the (0,0,0)-(1,1,1) diagonal (src/beat/geometry.py:133-139).
"""

from __future__ import annotations

import itertools
import math
import os
from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np


# ----------------------------------------------------------------------------------------- comm
@dataclass(frozen=True)
class Comm:
    """Rank / size carrier (stands in for mpi4py's communicator; collectives go through NCCL)."""

    rank: int = 0
    size: int = 1

    def Get_rank(self) -> int:
        return self.rank

    def Get_size(self) -> int:
        return self.size


def comm_world() -> Comm:
    return Comm(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


COMM_SELF = Comm(0, 1)


# ------------------------------------------------------------------------------------- index map
@dataclass
class IndexMap:
    size_local: int
    ghosts: np.ndarray  # global ids of ghosts, grouped by owner
    owners: np.ndarray  # owner rank of each ghost
    local_to_global: np.ndarray  # owned then ghosts
    size_global: int
    # halo pattern (neighbour ranks ascending)
    nbr_ranks: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    send_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    send_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    recv_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))

    @property
    def num_ghosts(self) -> int:
        return int(self.ghosts.shape[0])


class _Geometry:
    def __init__(self, x: np.ndarray):
        self.x = x  # (n_local, 3) like dolfinx (always 3 columns)


class _Topology:
    def __init__(self, dim: int, index_map: IndexMap):
        self.dim = dim
        self._vertex_map = index_map

    def index_map(self, dim: int) -> IndexMap:
        if dim != 0:
            raise NotImplementedError("only the vertex index map is provided")
        return self._vertex_map


class Mesh:
    """Simplicial P1 mesh, local part of one rank."""

    def __init__(self, comm: Comm, points: np.ndarray, cells: np.ndarray, index_map: IndexMap, tdim: int,
                 cell_global: np.ndarray | None = None, info: dict | None = None):
        self.comm = comm
        pts3 = np.zeros((points.shape[0], 3))
        pts3[:, : points.shape[1]] = points
        self.geometry = _Geometry(pts3)
        self.gdim = points.shape[1]
        self._cells = None if cells is None else np.ascontiguousarray(cells, dtype=np.int64)  # local vertex ids, (ncell, tdim+1)
        self.topology = _Topology(tdim, index_map)
        self.index_map = index_map
        self.cell_global = cell_global
        self.info = info or {}
        self._device_ctx = None
        self._boundary_facets = None

    # one device context per mesh (= per rank/GPU): PDE and ODE stages of a simulation share it
    def device_context(self):
        if self._device_ctx is None:
            from ._lib import Context

            dev = int(os.environ.get("LOCAL_RANK", "0")) if self.comm.size > 1 else int(os.environ.get("MONO_DEVICE", "0"))
            self._device_ctx = Context(dev)
        return self._device_ctx

    def new_device_context(self):
        """Drop the cached context (a new simulation on the same mesh needs a fresh one)."""
        self._device_ctx = None
        return self.device_context()

    @property
    def num_local_vertices(self) -> int:
        return self.geometry.x.shape[0]

    @property
    def cells(self) -> np.ndarray:
        return self._cells

    @property
    def num_cells(self) -> int:
        return self.cells.shape[0]

    def cells_of(self, ids: np.ndarray) -> np.ndarray:
        """Local vertex ids of the given cells."""
        return self.cells[ids]

    def boundary_facets(self) -> np.ndarray:
        """Exterior facets of the GLOBAL mesh among the local cells, as sorted local vertex tuples."""
        if self._boundary_facets is None:
            d = self.topology.dim
            fac = np.concatenate([np.delete(self.cells, a, axis=1) for a in range(d + 1)], axis=0)
            fac = np.sort(fac, axis=1)
            on_ids = self.info.get("on_boundary_ids")
            if on_ids is not None:
                # structured generators know the exterior from the vertex indices: a cell face that lies in an exterior
                # surface belongs to exactly one cell, so no counting (and no sort of all faces) is needed
                cand = fac[np.asarray(on_ids(self.index_map.local_to_global[fac]), dtype=bool)]
                order = np.lexsort(tuple(cand[:, k] for k in range(cand.shape[1] - 1, -1, -1)))
                cand = cand[order]
            else:
                uniq, counts = np.unique(fac, axis=0, return_counts=True)
                cand = uniq[counts == 1]
            # a facet seen once locally may be interior to the global mesh (its other cell lives on another
            # rank); such a facet has only ghost/boundary vertices.  Structured generators pass a predicate.
            on_bnd = self.info.get("on_boundary")
            if on_bnd is not None:
                keep = np.ones(cand.shape[0], dtype=bool)
                x = self.geometry.x
                keep &= on_bnd(x[cand].transpose(2, 0, 1))
                cand = cand[keep]
            self._boundary_facets = cand
        return self._boundary_facets


def _partition_1d(n_planes: int, size: int) -> np.ndarray:
    """Plane ranges [start[r], start[r+1]) of a balanced 1-D block partition."""
    base, rem = divmod(n_planes, size)
    counts = np.array([base + (1 if r < rem else 0) for r in range(size)])
    return np.concatenate([[0], np.cumsum(counts)])


def _build_local(comm: Comm, cells_g: np.ndarray, coords_of: Callable[[np.ndarray], np.ndarray], owner_of: Callable[[np.ndarray], np.ndarray],
                 n_global: int, tdim: int, info: dict) -> Mesh:
    """cells_g: global vertex ids of every cell touching a vertex owned by this rank."""
    rank = comm.rank
    gids = np.unique(cells_g)
    own = owner_of(gids)
    owned = gids[own == rank]
    ghost_mask = own != rank
    ghosts = gids[ghost_mask]
    gown = own[ghost_mask]
    order = np.lexsort((ghosts, gown))
    ghosts, gown = ghosts[order], gown[order]
    l2g = np.concatenate([owned, ghosts])
    # global -> local through a sorted lookup
    sorter = np.argsort(l2g)
    cells_l = sorter[np.searchsorted(l2g, cells_g, sorter=sorter)]
    n_owned = owned.shape[0]

    # halo pattern, computed without communication (see module docstring)
    nbr = np.unique(gown).astype(np.int32)
    recv_ptr = np.concatenate([[0], np.cumsum([np.count_nonzero(gown == q) for q in nbr])]).astype(np.int32)
    send_lists = []
    if nbr.size:
        cell_owner = own[np.searchsorted(gids, cells_g)]  # owner of each cell vertex
        for q in nbr:
            touches_q = (cell_owner == q).any(axis=1)
            mine = cells_g[touches_q][cell_owner[touches_q] == rank]
            send_g = np.unique(mine)
            send_lists.append(sorter[np.searchsorted(l2g, send_g, sorter=sorter)].astype(np.int32))
    send_ptr = np.concatenate([[0], np.cumsum([len(s) for s in send_lists])]).astype(np.int32)
    send_idx = np.concatenate(send_lists).astype(np.int32) if send_lists else np.zeros(0, np.int32)
    imap = IndexMap(n_owned, ghosts, gown.astype(np.int32), l2g, n_global, nbr, send_ptr, send_idx, recv_ptr)
    return Mesh(comm, coords_of(l2g), cells_l, imap, tdim, info=info)


def create_interval(comm: Comm, n: int, points=(0.0, 1.0)) -> Mesh:
    a, b = points
    starts = _partition_1d(n + 1, comm.size)
    lo, hi = starts[comm.rank], starts[comm.rank + 1]
    c0, c1 = max(lo - 1, 0), min(hi, n)
    ids = np.arange(c0, c1)
    cells = np.stack([ids, ids + 1], axis=1)
    h = (b - a) / n
    return _build_local(
        comm, cells, lambda g: (a + h * g).reshape(-1, 1), lambda g: np.searchsorted(starts, g, side="right") - 1, n + 1, 1,
        {"kind": "interval", "n": (n,), "p0": (a,), "p1": (b,),
         "on_boundary": lambda x: np.all(np.isclose(x[0], a) | np.isclose(x[0], b), axis=1)},
    )


def create_unit_interval(comm: Comm, n: int) -> Mesh:
    return create_interval(comm, n, (0.0, 1.0))


def create_rectangle(comm: Comm, points, n, cell_type=None, dtype=np.float64) -> Mesh:
    (x0, y0), (x1, y1) = (tuple(points[0])[:2], tuple(points[1])[:2])
    nx, ny = int(n[0]), int(n[1])
    starts = _partition_1d(nx + 1, comm.size)
    lo, hi = starts[comm.rank], starts[comm.rank + 1]
    c0, c1 = max(lo - 1, 0), min(hi, nx)
    ix, iy = np.meshgrid(np.arange(c0, c1), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2 = v0 + 1, v0 + (nx + 1)
    v3 = v2 + 1
    cells = np.concatenate([np.stack([v0, v1, v3], axis=1), np.stack([v0, v2, v3], axis=1)], axis=0)
    hx, hy = (x1 - x0) / nx, (y1 - y0) / ny

    def coords(g):
        return np.stack([x0 + hx * (g % (nx + 1)), y0 + hy * (g // (nx + 1))], axis=1)

    def on_boundary(x):  # x: (3, nfacets, nverts)
        return (
            np.all(np.isclose(x[0], x0), axis=1) | np.all(np.isclose(x[0], x1), axis=1)
            | np.all(np.isclose(x[1], y0), axis=1) | np.all(np.isclose(x[1], y1), axis=1)
        )

    return _build_local(
        comm, cells, coords, lambda g: np.searchsorted(starts, g % (nx + 1), side="right") - 1, (nx + 1) * (ny + 1), 2,
        {"kind": "rectangle", "n": (nx, ny), "p0": (x0, y0), "p1": (x1, y1), "on_boundary": on_boundary},
    )


def create_unit_square(comm: Comm, nx: int, ny: int, cell_type=None) -> Mesh:
    return create_rectangle(comm, [np.array([0.0, 0.0]), np.array([1.0, 1.0])], [nx, ny])


_KUHN_PERMS = tuple(itertools.permutations((0, 1, 2)))  # axis order of the three steps from corner 000 to corner 111


class BoxMesh(Mesh):
    """Kuhn-split box (dolfinx.mesh.create_box with tetrahedra, src/beat/geometry.py:133-139), x-slab partition.

    Everything is index arithmetic on the (nx+1, ny+1, nz+1) vertex grid, so a 30M-dof slab needs no cell array:
    cells are generated on demand (``cells`` materialises all of them, ``cells_of`` only the ones asked for) and the
    P1 matrices are assembled from the six element matrices of the reference cube (``assemble_p1_box``).
    Numbering (identical to the generic ``_build_local``): owned vertices sorted by global id, then the ghost
    planes grouped by owner; cell c = perm * ncube + cube, cubes ordered z, y, x (x fastest, local cube range only).
    """

    def __init__(self, comm: Comm, p0, p1, n):
        nx, ny, nz = (int(v) for v in n)
        self.n = (nx, ny, nz)
        self.p0, self.p1 = tuple(float(v) for v in p0), tuple(float(v) for v in p1)
        self.h = tuple((self.p1[k] - self.p0[k]) / m for k, m in enumerate(self.n))
        starts = _partition_1d(nx + 1, comm.size)
        self.starts = starts
        lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
        self.lo, self.hi = lo, hi
        self.c0, self.c1 = max(lo - 1, 0), min(hi, nx)  # local cubes: every cube that touches an owned vertex
        self.nxo = hi - lo
        plane = (ny + 1) * (nz + 1)
        self.plane = plane
        self.has_left, self.has_right = lo > 0, hi < nx + 1
        n_owned = plane * self.nxo
        self.sx, self.sy, self.sz = 1, nx + 1, (nx + 1) * (ny + 1)

        kk, jj = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), indexing="ij")
        pl = (kk * (ny + 1) + jj).ravel().astype(np.int64)  # plane-local index -> (k, j)
        base = pl * (nx + 1)  # gid of (i=0, j, k)
        gk, gj, gi = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(lo, hi), indexing="ij")
        owned_g = ((gk * (ny + 1) + gj) * (nx + 1) + gi).ravel().astype(np.int64)
        ghosts, owners, nbr, send_lists = [], [], [], []
        if self.has_left:
            ghosts.append(base + (lo - 1))
            owners.append(np.full(plane, comm.rank - 1, dtype=np.int32))
            nbr.append(comm.rank - 1)
            send_lists.append((pl * self.nxo + 0).astype(np.int32))               # owned plane i = lo
        if self.has_right:
            ghosts.append(base + hi)
            owners.append(np.full(plane, comm.rank + 1, dtype=np.int32))
            nbr.append(comm.rank + 1)
            send_lists.append((pl * self.nxo + (self.nxo - 1)).astype(np.int32))  # owned plane i = hi-1
        ghosts_g = np.concatenate(ghosts) if ghosts else np.zeros(0, np.int64)
        gown = np.concatenate(owners) if owners else np.zeros(0, np.int32)
        l2g = np.concatenate([owned_g, ghosts_g])
        recv_ptr = np.arange(len(nbr) + 1, dtype=np.int32) * plane
        send_ptr = np.arange(len(nbr) + 1, dtype=np.int32) * plane
        send_idx = np.concatenate(send_lists).astype(np.int32) if send_lists else np.zeros(0, np.int32)
        imap = IndexMap(n_owned, ghosts_g, gown, l2g, (nx + 1) * plane, np.asarray(nbr, dtype=np.int32), send_ptr, send_idx, recv_ptr)

        def on_boundary(x):
            m = np.zeros(x.shape[1], dtype=bool)
            for k in range(3):
                m |= np.all(np.isclose(x[k], self.p0[k]), axis=1) | np.all(np.isclose(x[k], self.p1[k]), axis=1)
            return m

        super().__init__(comm, self._coords(l2g), None, imap, 3,
                         info={"kind": "box", "n": self.n, "p0": self.p0, "p1": self.p1, "on_boundary": on_boundary})

    def _coords(self, g):
        nx, ny, _ = self.n
        gx = g % (nx + 1)
        gy = (g // (nx + 1)) % (ny + 1)
        gz = g // self.sz
        return np.stack([self.p0[0] + self.h[0] * gx, self.p0[1] + self.h[1] * gy, self.p0[2] + self.h[2] * gz], axis=1)

    # ---- local numbering ------------------------------------------------------------------------------
    def local_index(self, i, j, k):
        """Local vertex id of grid point (i, j, k), i in [lo-1, hi] (arrays broadcast)."""
        pl = k * (self.n[1] + 1) + j
        n_owned = self.index_map.size_local
        left = n_owned + pl
        right = n_owned + (self.plane if self.has_left else 0) + pl
        owned = pl * self.nxo + (i - self.lo)
        return np.where(i < self.lo, left, np.where(i >= self.hi, right, owned))

    @property
    def num_cubes(self) -> int:
        return (self.c1 - self.c0) * self.n[1] * self.n[2]

    @property
    def num_cells(self) -> int:
        return 6 * self.num_cubes

    def _cube_ijk(self, cube):
        w = self.c1 - self.c0
        return self.c0 + cube % w, (cube // w) % self.n[1], cube // (w * self.n[1])

    def cells_of(self, ids: np.ndarray) -> np.ndarray:
        ids = np.asarray(ids, dtype=np.int64)
        ncube = self.num_cubes
        perm = ids // ncube
        i, j, k = self._cube_ijk(ids % ncube)
        steps = np.asarray(_KUHN_PERMS, dtype=np.int64)[perm]  # (m, 3) axis of step 1, 2, 3
        out = np.empty((ids.shape[0], 4), dtype=np.int64)
        off = np.zeros((ids.shape[0], 3), dtype=np.int64)
        out[:, 0] = self.local_index(i, j, k)
        rows = np.arange(ids.shape[0])
        for a in range(3):
            off[rows, steps[:, a]] += 1
            out[:, a + 1] = self.local_index(i + off[:, 0], j + off[:, 1], k + off[:, 2])
        return out

    @property
    def cells(self) -> np.ndarray:
        """All local cells, cell id = perm * num_cubes + cube.  The 8 corners of every cube are indexed once and the six
        Kuhn tets pick their 4 from them (same rows as cells_of(arange), a third of the index arithmetic)."""
        if self._cells is None:
            ncube = self.num_cubes
            i, j, k = self._cube_ijk(np.arange(ncube, dtype=np.int64))
            corner = {(dx_, dy_, dz_): self.local_index(i + dx_, j + dy_, k + dz_) for dx_ in (0, 1) for dy_ in (0, 1) for dz_ in (0, 1)}
            out = np.empty((6, ncube, 4), dtype=np.int64)
            for p, perm in enumerate(_KUHN_PERMS):
                off = [0, 0, 0]
                out[p, :, 0] = corner[0, 0, 0]
                for a in range(3):
                    off[perm[a]] += 1
                    out[p, :, a + 1] = corner[tuple(off)]
            self._cells = out.reshape(-1, 4)
        return self._cells

    def locate_cells(self, ok: np.ndarray) -> np.ndarray:
        """Cells whose 4 vertices all satisfy the vertex predicate ``ok`` (n_local bools).  Every Kuhn tet contains
        the 000 and 111 corners of its cube, so only cubes with both corners marked are expanded."""
        cube = np.arange(self.num_cubes, dtype=np.int64)
        i, j, k = self._cube_ijk(cube)
        cand = cube[ok[self.local_index(i, j, k)] & ok[self.local_index(i + 1, j + 1, k + 1)]]
        ids = (np.arange(6, dtype=np.int64)[:, None] * self.num_cubes + cand[None, :]).ravel()
        if ids.size == 0:
            return ids.astype(np.int32)
        keep = ok[self.cells_of(ids)].all(axis=1)
        return np.sort(ids[keep]).astype(np.int32 if self.num_cells < 2**31 else np.int64)


class ShellMesh(Mesh):
    """Kuhn-split structured mesh that is PERIODIC in its fastest index and geometrically mapped (the synthetic LV shell:
    x = phi ring, y = mu, z = transmural).  Like BoxMesh everything is index arithmetic - vertex id
    g = (k (ny+1) + j) vx + (i mod vx) with vx = nx vertices per ring - so a rank's part needs neither a global cell
    array nor np.unique; the partition cuts the ring into sectors (neighbours (r-1) % size and (r+1) % size).
    Unwrapped local column index i runs over [lo-1, hi] (ghost, owned..., ghost); numbering follows the same rules as
    the generic ``_build_local`` (owned by global id, ghosts grouped by owner then by global id), which the tests
    compare it with."""

    def __init__(self, comm: Comm, n, coords_ijk, info: dict | None = None):
        nx, ny, nz = (int(v) for v in n)
        self.n = (nx, ny, nz)
        self.vx = vx = nx
        self.coords_ijk = coords_ijk
        size, rank = comm.size, comm.rank
        starts = _partition_1d(vx, size)
        self.lo, self.hi = lo, hi = int(starts[rank]), int(starts[rank + 1])
        self.nxo = nxo = hi - lo
        if size > 1 and nxo < 2:
            raise ValueError("every rank needs at least two columns of the ring")
        self.plane = plane = (ny + 1) * (nz + 1)
        self.periodic_single = size == 1
        self.c0, self.c1 = (0, nx) if size == 1 else (lo - 1, hi)  # unwrapped cube columns
        n_owned = plane * nxo
        pl = np.arange(plane, dtype=np.int64)
        gk, gj, gi = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(lo, hi), indexing="ij")
        owned_g = ((gk * (ny + 1) + gj) * vx + gi).ravel().astype(np.int64)
        self._ghost_cols: dict[int, tuple[int, int, int]] = {}  # unwrapped column -> (group offset, columns in group, position)
        ghosts_g, gown, nbr, send_idx, send_ptr, recv_ptr = [], [], [], [], [0], [0]
        if size > 1:
            qL, qR = (rank - 1) % size, (rank + 1) % size
            groups: dict[int, list[int]] = {}
            groups.setdefault(qL, []).append(lo - 1)
            groups.setdefault(qR, []).append(hi)
            sends: dict[int, list[int]] = {}
            sends.setdefault(qL, []).append(lo)       # my first column is qL's right ghost
            sends.setdefault(qR, []).append(hi - 1)   # my last column is qR's left ghost
            off = 0
            for q in sorted(groups):
                cols = sorted(set(groups[q]), key=lambda c: c % vx)  # within an owner: by global id = by global column per row
                m = len(cols)
                for pos, c in enumerate(cols):
                    self._ghost_cols[c] = (off, m, pos)
                gcols = np.array([c % vx for c in cols], dtype=np.int64)
                ghosts_g.append((pl[:, None] * vx + gcols[None, :]).ravel())
                gown.append(np.full(plane * m, q, dtype=np.int32))
                off += plane * m
                recv_ptr.append(off)
                nbr.append(q)
                scols = np.array(sorted(set(sends[q]), key=lambda c: c % vx), dtype=np.int64)
                send_idx.append((pl[:, None] * nxo + (scols[None, :] - lo)).ravel().astype(np.int32))
                send_ptr.append(send_ptr[-1] + plane * len(scols))
        ghosts = np.concatenate(ghosts_g) if ghosts_g else np.zeros(0, np.int64)
        owners = np.concatenate(gown) if gown else np.zeros(0, np.int32)
        l2g = np.concatenate([owned_g, ghosts])
        imap = IndexMap(n_owned, ghosts, owners, l2g, vx * plane, np.asarray(nbr, dtype=np.int32), np.asarray(send_ptr, dtype=np.int32),
                        np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32), np.asarray(recv_ptr, dtype=np.int32))
        gi_, gj_, gk_ = l2g % vx, (l2g // vx) % (ny + 1), l2g // (vx * (ny + 1))
        inf = {"kind": "shell", "n": self.n}
        inf.update(info or {})
        super().__init__(comm, coords_ijk(gi_, gj_, gk_), None, imap, 3, info=inf)

    def local_index(self, i, j, k):
        """Local vertex id of grid point (i, j, k); i is the UNWRAPPED column (single rank: any integer, taken mod vx;
        several ranks: lo-1 .. hi)."""
        i = np.asarray(i)
        pl = np.asarray(k) * (self.n[1] + 1) + np.asarray(j)
        if self.periodic_single:
            return pl * self.vx + (i % self.vx)
        n_owned = self.index_map.size_local
        out = pl * self.nxo + (i - self.lo)
        for col, (off, m, pos) in self._ghost_cols.items():
            out = np.where(i == col, n_owned + off + pl * m + pos, out)
        return out

    @property
    def num_cubes(self) -> int:
        return (self.c1 - self.c0) * self.n[1] * self.n[2]

    @property
    def num_cells(self) -> int:
        return 6 * self.num_cubes

    def _cube_ijk(self, cube):
        w = self.c1 - self.c0
        return self.c0 + cube % w, (cube // w) % self.n[1], cube // (w * self.n[1])

    cells_of = BoxMesh.cells_of
    cells = BoxMesh.cells
    locate_cells = BoxMesh.locate_cells

    def boundary_facets(self) -> np.ndarray:
        """Exterior facets among the local cells, generated directly: the ring has no boundary in x; on the surfaces
        j = 0, j = ny, k = 0, k = nz every cube face is cut by the Kuhn split along its (0,0)-(1,1) diagonal into
        two triangles.  Same rows, same order as the generic search over all cell faces."""
        if self._boundary_facets is None:
            nx, ny, nz = self.n
            ii = np.arange(self.c0, self.c1)
            tris = []

            def quad_faces(v00, v10, v01, v11):
                tris.append(np.stack([v00, v10, v11], axis=-1).reshape(-1, 3))
                tris.append(np.stack([v00, v01, v11], axis=-1).reshape(-1, 3))

            for k in (0, nz):  # faces spanned by (x, y)
                I, J = np.meshgrid(ii, np.arange(ny), indexing="ij")
                L = lambda di, dj: self.local_index(I + di, J + dj, np.full_like(I, k))  # noqa: E731
                quad_faces(L(0, 0), L(1, 0), L(0, 1), L(1, 1))
            for j in (0, ny):  # faces spanned by (x, z)
                I, K = np.meshgrid(ii, np.arange(nz), indexing="ij")
                L = lambda di, dk: self.local_index(I + di, np.full_like(I, j), K + dk)  # noqa: E731
                quad_faces(L(0, 0), L(1, 0), L(0, 1), L(1, 1))
            fac = np.sort(np.concatenate(tris, axis=0), axis=1)
            order = np.lexsort((fac[:, 2], fac[:, 1], fac[:, 0]))
            self._boundary_facets = fac[order]
        return self._boundary_facets


def assemble_p1_structured(mesh, M):
    """CSR (indptr, indices, mass, stiff) of the owned rows of a ShellMesh: the stencil accumulation of assemble_p1_box with
    element matrices that VARY from cube to cube (mapped geometry, cell-wise tensor M of shape (ncell, 3, 3) in the
    mesh's cell order, or a constant).  Per tet type the volumes, gradients and element matrices of all cubes are
    computed at once from the vertex grid and added with slice additions - no global sort, no cell loop."""
    nx, ny, nz = mesh.n
    lo, hi, nxo = mesh.lo, mesh.hi, mesh.nxo
    c0, c1 = mesh.c0, mesh.c1
    ncx = c1 - c0
    ncube = mesh.num_cubes
    Mv = np.asarray(M.value if isinstance(M, Constant) else M, dtype=np.float64)
    # vertex coordinates on the unwrapped local grid [k, j, column c0 .. c1]
    kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(c0, c1 + 1), indexing="ij")
    Xg = mesh.geometry.x[mesh.local_index(ii, jj, kk)]  # (nz+1, ny+1, ncx+1, 3)
    corner = np.zeros((6, 4, 3), dtype=np.int64)
    for p, perm in enumerate(_KUHN_PERMS):
        for a in range(3):
            corner[p, a + 1] = corner[p, a]
            corner[p, a + 1, perm[a]] += 1
    offs = sorted({tuple(corner[p, b] - corner[p, a]) for p in range(6) for a in range(4) for b in range(4)},
                  key=lambda d: d[0] + d[1] * mesh.vx + d[2] * mesh.vx * (ny + 1))
    slot = {d: s for s, d in enumerate(offs)}
    ns = len(offs)
    single = mesh.periodic_single
    ncol = nx + 1 if single else nxo  # single rank: one extra (seam) column, folded onto column 0 afterwards
    row0 = 0 if single else lo        # unwrapped column of row-array column 0
    shape = (nz + 1, ny + 1, ncol)
    val_m = np.zeros((ns,) + shape)
    val_k = np.zeros((ns,) + shape)
    ref = (1.0 + np.eye(4)) / 20.0
    for p in range(6):
        verts = [Xg[corner[p, a, 2]: corner[p, a, 2] + nz, corner[p, a, 1]: corner[p, a, 1] + ny, corner[p, a, 0]: corner[p, a, 0] + ncx]
                 for a in range(4)]
        xcell = np.stack(verts, axis=3).reshape(-1, 4, 3)  # cube order z, y, x (x fastest) = the mesh's cell order
        vol, g = _simplex_measure_and_gradients(xcell)
        gt = g.transpose(0, 2, 1)
        if Mv.ndim == 0:
            Ke = float(Mv) * (g @ gt)
        elif Mv.ndim == 2:
            Ke = (g @ Mv) @ gt
        else:
            Ke = (g @ Mv[p * ncube: (p + 1) * ncube]) @ gt  # batched 4x3 . 3x3 . 3x4
        Ke *= vol[:, None, None]
        Ke = Ke.reshape(nz, ny, ncx, 4, 4)
        Me = (vol[:, None, None] * ref[None]).reshape(nz, ny, ncx, 4, 4)
        for a in range(4):
            ax, ay, az = corner[p, a]
            i0, i1 = (c0, c1) if single else (max(c0, lo - ax), min(c1, hi - ax))  # cubes whose vertex a is an owned row
            if i1 <= i0:
                continue
            rs = (slice(az, az + nz), slice(ay, ay + ny), slice(i0 + ax - row0, i1 + ax - row0))
            cs = (slice(None), slice(None), slice(i0 - c0, i1 - c0))
            for b in range(4):
                s_ = slot[tuple(corner[p, b] - corner[p, a])]
                val_m[s_][rs] += Me[cs + (a, b)]
                val_k[s_][rs] += Ke[cs + (a, b)]
    if single:  # the seam column nx is column 0
        val_m[..., 0] += val_m[..., nx]
        val_k[..., 0] += val_k[..., nx]
        val_m, val_k = val_m[..., :nx], val_k[..., :nx]
        shape = (nz + 1, ny + 1, nx)
    n_owned = mesh.index_map.size_local
    kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(lo, hi), indexing="ij", sparse=True)
    cols = np.empty((n_owned, ns), dtype=np.int32)
    valid = np.empty((n_owned, ns), dtype=bool)
    for s_, (dx_, dy_, dz_) in enumerate(offs):
        j2, k2 = jj + dy_, kk + dz_
        ok = np.broadcast_to((j2 >= 0) & (j2 <= ny) & (k2 >= 0) & (k2 <= nz), shape)  # the ring has no end in x
        c = mesh.local_index(ii + dx_, np.clip(j2, 0, ny), np.clip(k2, 0, nz))
        cols[:, s_] = np.broadcast_to(c, shape).reshape(-1)
        valid[:, s_] = ok.reshape(-1)
    mass = np.ascontiguousarray(val_m.reshape(ns, -1).T)
    stiff = np.ascontiguousarray(val_k.reshape(ns, -1).T)
    del val_m, val_k
    # rows whose neighbours wrap around the seam or live in a ghost column: local column order differs from the slot order
    pl = np.arange(mesh.plane, dtype=np.int64) * nxo
    rows = np.unique(np.concatenate([pl, pl + (nxo - 1)]))
    key = np.where(valid[rows], cols[rows].astype(np.int64), np.int64(2**40))
    order = np.argsort(key, axis=1, kind="stable")
    for arr in (cols, valid, mass, stiff):
        arr[rows] = np.take_along_axis(arr[rows], order, axis=1)
    counts = valid.sum(axis=1)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return indptr, cols[valid], mass[valid], stiff[valid]


def create_box(comm: Comm, points, n, cell_type=None, dtype=np.float64) -> Mesh:
    return BoxMesh(comm, points[0], points[1], n)


def _create_box_generic(comm: Comm, points, n) -> Mesh:
    """The same mesh through the generic builder (global cell array + np.unique): kept as the cross-check of BoxMesh."""
    p0 = tuple(float(v) for v in points[0])
    p1 = tuple(float(v) for v in points[1])
    nx, ny, nz = (int(v) for v in n)
    starts = _partition_1d(nx + 1, comm.size)
    lo, hi = starts[comm.rank], starts[comm.rank + 1]
    c0, c1 = max(lo - 1, 0), min(hi, nx)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(c0, c1), indexing="ij")
    v0 = ((iz * (ny + 1) + iy) * (nx + 1) + ix).ravel().astype(np.int64)
    strides = (1, nx + 1, (nx + 1) * (ny + 1))
    sz = strides[2]
    tets = []
    for perm in _KUHN_PERMS:
        a = v0 + strides[perm[0]]
        b = a + strides[perm[1]]
        tets.append(np.stack([v0, a, b, b + strides[perm[2]]], axis=1))
    cells = np.concatenate(tets, axis=0)
    h = [(p1[k] - p0[k]) / m for k, m in enumerate((nx, ny, nz))]

    def coords(g):
        gx = g % (nx + 1)
        gy = (g // (nx + 1)) % (ny + 1)
        gz = g // sz
        return np.stack([p0[0] + h[0] * gx, p0[1] + h[1] * gy, p0[2] + h[2] * gz], axis=1)

    def on_boundary(x):
        m = np.zeros(x.shape[1], dtype=bool)
        for k in range(3):
            m |= np.all(np.isclose(x[k], p0[k]), axis=1) | np.all(np.isclose(x[k], p1[k]), axis=1)
        return m

    return _build_local(
        comm, cells, coords, lambda g: np.searchsorted(starts, g % (nx + 1), side="right") - 1,
        (nx + 1) * (ny + 1) * (nz + 1), 3,
        {"kind": "box", "n": (nx, ny, nz), "p0": p0, "p1": p1, "on_boundary": on_boundary},
    )


def assemble_p1_box(mesh: "BoxMesh", Mv: np.ndarray):
    """CSR (indptr, indices, mass, stiff) of the owned rows of a BoxMesh for a CONSTANT conductivity tensor, without
    a cell array: the six Kuhn tets of every cube have the same element matrices, so entry (row, row + offset) is a
    sum of at most 24 constants, accumulated with 96 slice additions over the owned vertex grid."""
    nx, ny, nz = mesh.n
    lo, hi, nxo = mesh.lo, mesh.hi, mesh.nxo
    h = np.asarray(mesh.h)
    # element matrices of the 6 tets of the reference cube, and the cube-corner offset of each tet vertex
    corner = np.zeros((6, 4, 3), dtype=np.int64)
    for p, perm in enumerate(_KUHN_PERMS):
        for a in range(3):
            corner[p, a + 1] = corner[p, a]
            corner[p, a + 1, perm[a]] += 1
    vol, g = _simplex_measure_and_gradients(corner * h[None, None, :])
    if Mv.ndim == 0:
        Ke = float(Mv) * np.einsum("eai,ebi->eab", g, g)
    else:
        Ke = np.einsum("eai,ij,ebj->eab", g, Mv, g)
    Ke *= vol[:, None, None]
    Me = vol[:, None, None] * ((1.0 + np.eye(4)) / 20.0)[None]
    # stencil slots = distinct vertex offsets inside a tet, ordered by global-id distance (= local column order)
    offs = sorted({tuple(corner[p, b] - corner[p, a]) for p in range(6) for a in range(4) for b in range(4)},
                  key=lambda d: d[0] + d[1] * (nx + 1) + d[2] * (nx + 1) * (ny + 1))
    slot = {d: s for s, d in enumerate(offs)}
    ns = len(offs)
    shape = (nz + 1, ny + 1, nxo)
    val_m = np.zeros((ns,) + shape)
    val_k = np.zeros((ns,) + shape)
    for p in range(6):
        for a in range(4):
            ax, ay, az = corner[p, a]
            i0, i1 = max(mesh.c0, lo - ax), min(mesh.c1, hi - ax)  # cubes whose vertex a is an owned row
            if i1 <= i0:
                continue
            rs = (slice(az, az + nz), slice(ay, ay + ny), slice(i0 + ax - lo, i1 + ax - lo))
            for b in range(4):
                s_ = slot[tuple(corner[p, b] - corner[p, a])]
                val_m[s_][rs] += Me[p, a, b]
                val_k[s_][rs] += Ke[p, a, b]
    # columns: neighbour (i+dx, j+dy, k+dz) exists iff it lies in the global grid
    kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(lo, hi), indexing="ij", sparse=True)
    n_owned = mesh.index_map.size_local
    cols = np.empty((n_owned, ns), dtype=np.int32)
    valid = np.empty((n_owned, ns), dtype=bool)
    for s_, (dx_, dy_, dz_) in enumerate(offs):
        i2, j2, k2 = ii + dx_, jj + dy_, kk + dz_
        ok = (i2 >= 0) & (i2 <= nx) & (j2 >= 0) & (j2 <= ny) & (k2 >= 0) & (k2 <= nz)
        ok = np.broadcast_to(ok, shape)
        c = mesh.local_index(np.clip(i2, lo - 1, hi), np.clip(j2, 0, ny), np.clip(k2, 0, nz))
        cols[:, s_] = np.broadcast_to(c, shape).reshape(-1)
        valid[:, s_] = ok.reshape(-1)
    mass = np.ascontiguousarray(val_m.reshape(ns, -1).T)
    stiff = np.ascontiguousarray(val_k.reshape(ns, -1).T)
    del val_m, val_k
    # rows next to a ghost plane: their ghost columns carry the largest local ids, so re-sort those rows only
    pl = np.arange(mesh.plane, dtype=np.int64) * nxo
    fix = []
    if mesh.has_left:
        fix.append(pl)
    if mesh.has_right:
        fix.append(pl + (nxo - 1))
    if fix:
        rows = np.unique(np.concatenate(fix))
        key = np.where(valid[rows], cols[rows].astype(np.int64), np.int64(2**40))
        order = np.argsort(key, axis=1, kind="stable")
        for arr in (cols, valid, mass, stiff):
            arr[rows] = np.take_along_axis(arr[rows], order, axis=1)
    counts = valid.sum(axis=1)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return indptr, cols[valid], mass[valid], stiff[valid]


def _lv_maps(n_r: int, n_mu: int, n_phi: int, r_short_endo: float, r_short_epi: float, r_long_endo: float, r_long_epi: float,
             base: float, apex_cut: float):
    mu0, mu1_endo, mu1_epi = -math.pi + apex_cut, -math.acos(base / r_long_endo), -math.acos(base / r_long_epi)

    def decode(g):
        g = np.asarray(g)
        return g // ((n_mu + 1) * n_phi), (g // n_phi) % (n_mu + 1), g % n_phi  # (i_r, i_mu, i_phi)

    def coords_ijk(c, b, a):  # x = phi index, y = mu index, z = transmural index
        lam = np.asarray(a) / n_r
        rs = r_short_endo + lam * (r_short_epi - r_short_endo)
        rl = r_long_endo + lam * (r_long_epi - r_long_endo)
        mu = mu0 + (np.asarray(b) / n_mu) * ((mu1_endo + lam * (mu1_epi - mu1_endo)) - mu0)
        phi = 2.0 * math.pi * (np.asarray(c) % n_phi) / n_phi
        return np.stack([rl * np.cos(mu), rs * np.sin(mu) * np.cos(phi), rs * np.sin(mu) * np.sin(phi)], axis=-1)

    def on_boundary_ids(fg):  # facet (nf, 3) global vertex ids -> exterior?
        a, b, _ = decode(fg)
        return (np.all(a == 0, axis=1) | np.all(a == n_r, axis=1) | np.all(b == 0, axis=1) | np.all(b == n_mu, axis=1))

    return decode, coords_ijk, on_boundary_ids


def create_lv_ellipsoid(comm: Comm, n_r: int, n_mu: int, n_phi: int, r_short_endo: float = 2.5, r_short_epi: float = 3.5,
                        r_long_endo: float = 9.0, r_long_epi: float = 9.7, base: float = 0.0, apex_cut: float = 0.25) -> Mesh:
    """Synthetic truncated prolate-ellipsoid shell in the spirit of demos/lv_endocardial.py:35-58 (radii from there; the
    reference builds it with cardiac_geometries/gmsh, which are not available here): a structured (transmural, mu,
    phi) grid, periodic in phi, Kuhn-split into tetrahedra.  x = r_long cos(mu) is the long axis, the apex (mu = -pi)
    is cut off at mu = -pi + apex_cut to avoid the degenerate pole.  Partition: phi sectors (periodic: the first and
    the last rank are neighbours).  Vertex id = (i_r (n_mu+1) + i_mu) n_phi + i_phi; mesh.info carries the index maps
    needed for markers (transmural layer of a vertex, exterior surfaces).  Built by index arithmetic (ShellMesh)."""
    decode, coords_ijk, on_boundary_ids = _lv_maps(n_r, n_mu, n_phi, r_short_endo, r_short_epi, r_long_endo, r_long_epi, base, apex_cut)
    return ShellMesh(comm, (n_phi, n_mu, n_r), coords_ijk,
                     info={"kind": "lv_ellipsoid", "n": (n_r, n_mu, n_phi), "on_boundary_ids": on_boundary_ids, "decode": decode})


def _create_lv_ellipsoid_generic(comm: Comm, n_r: int, n_mu: int, n_phi: int, r_short_endo: float = 2.5, r_short_epi: float = 3.5,
                                 r_long_endo: float = 9.0, r_long_epi: float = 9.7, base: float = 0.0, apex_cut: float = 0.25) -> Mesh:
    """The same mesh through the generic builder (cell array + np.unique): kept as the cross-check of ShellMesh."""
    decode, coords_ijk, on_boundary_ids = _lv_maps(n_r, n_mu, n_phi, r_short_endo, r_short_epi, r_long_endo, r_long_epi, base, apex_cut)
    starts = _partition_1d(n_phi, comm.size)
    lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
    cols = np.arange(lo - 1, hi) if comm.size > 1 else np.arange(n_phi)  # unwrapped cube columns, as ShellMesh orders them
    ir, im, ip = np.meshgrid(np.arange(n_r), np.arange(n_mu), cols, indexing="ij")
    ir, im, ip = ir.ravel(), im.ravel(), ip.ravel()

    def vid(c, b, a):
        return (a * (n_mu + 1) + b) * n_phi + (c % n_phi)

    tets = []
    for perm in _KUHN_PERMS:  # steps along (phi, mu, transmural) in the order the permutation says
        off = [0, 0, 0]
        verts = [vid(ip, im, ir)]
        for ax in perm:
            off[ax] += 1
            verts.append(vid(ip + off[0], im + off[1], ir + off[2]))
        tets.append(np.stack(verts, axis=1))
    cells = np.concatenate(tets, axis=0).astype(np.int64)

    def coords(g):
        a, b, c = decode(g)
        return coords_ijk(c, b, a)

    return _build_local(comm, cells, coords, lambda g: np.searchsorted(starts, np.asarray(g) % n_phi, side="right") - 1,
                        (n_r + 1) * (n_mu + 1) * n_phi, 3,
                        {"kind": "lv_ellipsoid", "n": (n_r, n_mu, n_phi), "on_boundary_ids": on_boundary_ids, "decode": decode})


# ---------------------------------------------------------------------------- entities and tags
@dataclass
class MeshTags:
    mesh: Mesh
    dim: int
    indices: np.ndarray  # cell ids (dim == tdim) or rows of `facets` (dim == tdim-1) or vertex ids (dim == 0)
    values: np.ndarray
    facets: np.ndarray | None = None  # (nfacets, tdim) local vertex ids, for facet tags

    def find(self, value: int) -> np.ndarray:
        return self.indices[self.values == value]


def locate_entities(mesh: Mesh, dim: int, marker: Callable[[np.ndarray], np.ndarray]) -> np.ndarray:
    """Entities whose vertices ALL satisfy ``marker(x)`` (x has shape (3, npoints)), dolfinx semantics."""
    ok = np.asarray(marker(mesh.geometry.x.T), dtype=bool)
    if dim == mesh.topology.dim:
        if isinstance(mesh, (BoxMesh, ShellMesh)):
            return mesh.locate_cells(ok)
        return np.nonzero(ok[mesh.cells].all(axis=1))[0].astype(np.int32)
    if dim == 0:
        return np.nonzero(ok)[0].astype(np.int32)
    raise NotImplementedError("interior sub-entities: use locate_entities_boundary for facets")


class _FacetSelection(np.ndarray):
    """int32 ids into mesh.boundary_facets() (keeps the facet table attached for meshtags)."""


def locate_entities_boundary(mesh: Mesh, dim: int, marker: Callable[[np.ndarray], np.ndarray]) -> np.ndarray:
    if dim != mesh.topology.dim - 1:
        raise NotImplementedError("only boundary facets are provided")
    fac = mesh.boundary_facets()
    ok = np.asarray(marker(mesh.geometry.x.T), dtype=bool)
    return np.nonzero(ok[fac].all(axis=1))[0].astype(np.int32)


def meshtags(mesh: Mesh, dim: int, entities: np.ndarray, values) -> MeshTags:
    entities = np.asarray(entities, dtype=np.int32)
    values = np.broadcast_to(np.asarray(values, dtype=np.int32), entities.shape).copy()
    facets = mesh.boundary_facets() if dim == mesh.topology.dim - 1 and dim != 0 else None
    if dim == 0 and mesh.topology.dim == 1:
        facets = mesh.boundary_facets()
    return MeshTags(mesh, dim, entities, values, facets)


class Measure:
    """ufl.Measure("dx" | "ds", domain=mesh, subdomain_data=tags); calling it selects a marker."""

    def __init__(self, kind: str, domain: Mesh | None = None, subdomain_data: MeshTags | None = None, marker: int | None = None,
                 metadata: dict | None = None):
        if kind not in ("dx", "ds"):
            raise ValueError("measure must be 'dx' or 'ds'")
        self.kind, self.domain, self.subdomain_data, self.marker = kind, domain, subdomain_data, marker
        self.metadata = metadata

    def __call__(self, marker: int | None = None, **kw) -> "Measure":
        return Measure(self.kind, kw.get("domain", self.domain), self.subdomain_data, marker, self.metadata)


def dx(domain: Mesh | None = None, **kw) -> Measure:
    return Measure("dx", domain, **kw)


# ------------------------------------------------------------------------ functions, constants
class Constant:
    """dolfinx.fem.Constant: a mutable scalar / small array with ``.value``."""

    def __init__(self, mesh: Mesh | None, value):
        self.mesh = mesh
        self._value = np.array(value, dtype=np.float64)

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        self._value[...] = v

    def __float__(self) -> float:
        return float(self._value)


class Vector:
    """``Function.x``: host mirror of a device vector with explicit coherence.

    ``.array`` returns the (writable) host array, downloading first if the device copy is newer; because
    the caller may write into it, the next device operation re-uploads it.  ``.array_ro`` gives a
    read-only view without that upload.

    Two refinements keep the per-step host traffic at one copy each way:
      * uploads from the page-locked mirror are asynchronous; the mirror is not handed out again before the
        stream has consumed it (``_upload_guard``);
      * after a fused split step the reference's post-condition v_ode == v == v_ holds on the device, so a
        vector whose *twin* has already been downloaded refreshes its mirror with a host copy instead of a
        second device-to-host transfer (``twin_of``).
    """

    def __init__(self, n: int):
        from ._lib import pinned_zeros

        self._host = pinned_zeros(n)
        self._download: Callable[[np.ndarray], None] | None = None
        self._upload: Callable[[np.ndarray], None] | None = None
        self._sync: Callable[[], None] | None = None
        self.device_newer = False
        self.host_dirty = False
        self._upload_in_flight = False
        self._epoch = 0          # bumps whenever the device copy changes
        self._twin: "Vector | None" = None
        self._twin_epoch = -1

    def bind(self, download, upload, push_now: bool = True, sync=None):
        self._download, self._upload, self._sync = download, upload, sync
        self.host_dirty = push_now
        self.device_newer = False

    def _upload_guard(self):
        if self._upload_in_flight:
            if self._sync is not None:
                self._sync()
            self._upload_in_flight = False

    def _pull(self):
        self._upload_guard()
        if self.device_newer and self._download is not None:
            tw = self._twin
            if tw is not None and tw._epoch == self._twin_epoch and not tw.device_newer and not tw.host_dirty:
                np.copyto(self._host, tw._host)  # same device content, already on the host
            else:
                self._download(self._host)
            self.device_newer = False

    @property
    def array(self) -> np.ndarray:
        self._pull()
        if self._upload is not None:
            self.host_dirty = True
        return self._host

    @property
    def array_wo(self) -> np.ndarray:
        """The host array for a caller that OVERWRITES all of it: no download of the device copy first (``.array`` has to
        assume a partial update and refreshes the mirror before handing it out); the next device operation uploads it."""
        self._upload_guard()
        self.device_newer = False
        if self._upload is not None:
            self.host_dirty = True
        return self._host

    @property
    def array_ro(self) -> np.ndarray:
        self._pull()
        v = self._host.view()
        v.flags.writeable = False
        return v

    def flush_to_device(self):
        if self.host_dirty and self._upload is not None:
            self._upload(self._host)
            self._upload_in_flight = self._sync is not None
            self._epoch += 1
            self._twin = None
        self.host_dirty = False

    def mark_device_newer(self, twin_of: "Vector | None" = None):
        self.device_newer = True
        self.host_dirty = False
        self._epoch += 1
        self._twin = twin_of
        self._twin_epoch = twin_of._epoch if twin_of is not None else -1

    def scatter_forward(self):
        """Ghost refresh.  The device step already refreshes ghosts of the solution (inside the PDE kernel)."""
        return None


class _Element:
    def __init__(self, family: str, degree: int):
        self.family_name, self._degree = family, degree

    def degree(self) -> int:
        return self._degree


class _DofMap:
    def __init__(self, index_map: IndexMap):
        self.index_map = index_map
        self.index_map_bs = 1


class FunctionSpace:
    def __init__(self, mesh: Mesh, family: str = "Lagrange", degree: int = 1):
        if family not in ("P", "Lagrange", "CG") or degree != 1:
            raise NotImplementedError(
                "the device path implements P1 Lagrange spaces (identity ODE<->PDE projection, "
                "src/beat/utils.py:52-54); other spaces are a 'next' row of SURVEY.md section 8f"
            )
        self.mesh = mesh
        self.dofmap = _DofMap(mesh.index_map)
        self._element = _Element("Lagrange", 1)

    def ufl_element(self):
        return self._element

    def tabulate_dof_coordinates(self) -> np.ndarray:
        return self.mesh.geometry.x

    @property
    def num_local_dofs(self) -> int:
        return self.mesh.index_map.size_local + self.mesh.index_map.num_ghosts


def functionspace(mesh: Mesh, element=("Lagrange", 1)) -> FunctionSpace:
    family, degree = element[0], element[1]
    return FunctionSpace(mesh, family, degree)


class Function:
    def __init__(self, V: FunctionSpace, name: str = "f"):
        self.function_space = V
        self.name = name
        self.x = Vector(V.num_local_dofs)

    def ufl_element(self):
        return self.function_space.ufl_element()

    def interpolate(self, f: Callable[[np.ndarray], np.ndarray]):
        self.x.array[:] = f(self.function_space.tabulate_dof_coordinates().T)


# ----------------------------------------------------------------------- source-term expressions
class TimeWindow:
    """conditional(And(ge(time, start), le(time, end)), amplitude, 0): the reference's stimulus
    (src/beat/stimulation.py:270).  Evaluated ON THE DEVICE: the window test is part of the RHS kernel."""

    def __init__(self, time: Constant, start: float, end: float, amplitude):
        self.time, self.start, self.end = time, float(start), float(end)
        self.amplitude = amplitude  # float or Constant; Stimulus.assign overwrites it

    def amplitude_now(self) -> float:
        return float(self.amplitude)


class TimeFunction:
    """I_s = h(t): amplitude evaluated on the host at the theta-point each step, applied on the device."""

    def __init__(self, time: Constant, h: Callable[[float], float]):
        self.time, self.h = time, h
        self.amplitude = 1.0

    def amplitude_now(self) -> float:
        return float(self.amplitude) * float(self.h(float(self.time.value)))


class Separable(TimeFunction):
    """I_s(x, t) = g(x) * h(t) (the manufactured sources of tests/test_monodomain.py:13-36): g enters the
    load vector by quadrature at set-up, h(t) is the per-step amplitude."""

    def __init__(self, time: Constant, g: Callable[[np.ndarray], np.ndarray], h: Callable[[float], float], degree: int = 6):
        super().__init__(time, h)
        self.g, self.degree = g, degree


class WindowedField(TimeWindow):
    """conditional(And(<x in a region>, ge(time, start), le(time, end)), amplitude, 0): a TimeWindow whose amplitude is
    multiplied by a spatial indicator ``g(x)`` (array (gdim, n) -> 0/1).  The window test runs on the device like
    TimeWindow's; ``g`` enters the load vector at set-up with the quadrature UFL would estimate for such a conditional
    (degree of the true/false values + the P1 test function = 1: the centroid rule)."""

    def __init__(self, time: Constant, start: float, end: float, amplitude, g: Callable[[np.ndarray], np.ndarray], degree: int = 1):
        super().__init__(time, start, end, amplitude)
        self.g, self.degree = g, degree

    def evaluate(self, x: np.ndarray, t: float | None = None) -> np.ndarray:
        """Value at the points x (gdim, n) at time t (default: the current value of ``time``)."""
        t = float(self.time.value) if t is None else float(t)
        on = self.start <= t <= self.end
        return float(self.amplitude) * np.asarray(self.g(x), dtype=np.float64) if on else np.zeros(np.asarray(x).shape[1])


class ExprSum:
    """A sum of source terms (what ``a + b`` of two UFL expressions is): the model turns every term into its own stimulus."""

    def __init__(self, terms):
        self.terms = list(terms)

    def __add__(self, other):
        return ExprSum(self.terms + (other.terms if isinstance(other, ExprSum) else [other]))

    def evaluate(self, x: np.ndarray, t: float | None = None) -> np.ndarray:
        out = np.zeros(np.asarray(x).shape[1])
        for e in self.terms:
            out = out + e.evaluate(x, t)
        return out


class _Cmp:
    def __init__(self, op: str, a, b):
        self.op, self.a, self.b = op, a, b


class _And:
    def __init__(self, *terms):
        self.terms = terms


def ge(a, b):
    return _Cmp("ge", a, b)


def le(a, b):
    return _Cmp("le", a, b)


def And(*terms):
    return _And(*terms)


def conditional(cond, true_value, false_value):
    """Recognises the reference's window pattern and returns a device-evaluated TimeWindow."""
    if isinstance(cond, _And) and len(cond.terms) == 2 and float(false_value) == 0.0:
        lo = [t for t in cond.terms if isinstance(t, _Cmp) and t.op == "ge" and isinstance(t.a, Constant)]
        hi = [t for t in cond.terms if isinstance(t, _Cmp) and t.op == "le" and isinstance(t.a, Constant)]
        if len(lo) == 1 and len(hi) == 1 and lo[0].a is hi[0].a:
            return TimeWindow(lo[0].a, float(lo[0].b), float(hi[0].b), true_value)
    raise NotImplementedError(
        "only conditional(And(ge(time, a), le(time, b)), amplitude, 0) is recognised; use TimeFunction / "
        "Separable for other time dependences"
    )


# -------------------------------------------------------------------------------- P1 assembly
def _simplex_measure_and_gradients(x: np.ndarray):
    """x: (ncell, d+1, d).  Closed-form volume and P1 gradients (ncell, d+1, d) per dimension."""
    d = x.shape[2]
    e = x[:, 1:, :] - x[:, :1, :]  # edge vectors from vertex 0, (ncell, d, d)
    if d == 1:
        det = e[:, 0, 0]
        g = np.empty((x.shape[0], 2, 1))
        g[:, 1, 0] = 1.0 / det
        g[:, 0, 0] = -g[:, 1, 0]
        return np.abs(det), g
    if d == 2:
        a, b = e[:, 0], e[:, 1]
        det = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        g = np.empty((x.shape[0], 3, 2))
        g[:, 1, 0], g[:, 1, 1] = b[:, 1] / det, -b[:, 0] / det
        g[:, 2, 0], g[:, 2, 1] = -a[:, 1] / det, a[:, 0] / det
        g[:, 0] = -(g[:, 1] + g[:, 2])
        return np.abs(det) / 2.0, g
    a, b, c = e[:, 0], e[:, 1], e[:, 2]
    bc, ca, ab = np.cross(b, c), np.cross(c, a), np.cross(a, b)
    det = np.einsum("ij,ij->i", a, bc)
    g = np.empty((x.shape[0], 4, 3))
    g[:, 1], g[:, 2], g[:, 3] = bc / det[:, None], ca / det[:, None], ab / det[:, None]
    g[:, 0] = -(g[:, 1] + g[:, 2] + g[:, 3])
    return np.abs(det) / 6.0, g


def assemble_p1_local(mesh: Mesh, M) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """CSR (indptr int64, indices int32, mass, stiff) of the OWNED rows over local columns.

    M: scalar, (d,d) tensor, or per-cell (ncell,d,d) tensor.  Both matrices share one sparsity with sorted columns.
    Structured meshes take the stencil routes (index arithmetic, no cell array)."""
    d = mesh.topology.dim
    Mv = np.asarray(M.value if isinstance(M, Constant) else M, dtype=np.float64)
    if isinstance(mesh, BoxMesh) and Mv.ndim in (0, 2) and not os.environ.get("MONO_GENERIC_ASSEMBLY"):
        return assemble_p1_box(mesh, Mv)
    if isinstance(mesh, ShellMesh) and os.environ.get("MONO_STENCIL_ASSEMBLY"):
        return assemble_p1_structured(mesh, Mv)  # NumPy stencil route, ~7x slower than the library on 8 cores
    # everything else: the library's multi-threaded host assembler (rows gathered from their incident cells)
    from ._lib import fem_assemble_p1

    return fem_assemble_p1(d, mesh.index_map.size_local, mesh.cells, mesh.geometry.x, Mv)


def _assemble_p1_numpy(mesh: Mesh, M):
    """NumPy/SciPy restatement of the library's assembler (element matrices of all cells at once, duplicates summed by
    scipy's coo -> csr).  Kept as the cross-check of mono_fem_assemble_p1 in the tests."""
    d = mesh.topology.dim
    Mv = np.asarray(M.value if isinstance(M, Constant) else M, dtype=np.float64)
    cells = mesh.cells
    n_owned = mesh.index_map.size_local
    n_local = mesh.num_local_vertices
    x = mesh.geometry.x[:, :d][cells]
    vol, g = _simplex_measure_and_gradients(x)
    if Mv.ndim == 0:
        Ke = float(Mv) * np.einsum("eai,ebi->eab", g, g)
    elif Mv.ndim == 2:
        Ke = np.einsum("eai,ij,ebj->eab", g, Mv, g)
    else:
        Ke = np.einsum("eai,eij,ebj->eab", g, Mv, g)
    Ke *= vol[:, None, None]
    ref = (1.0 + np.eye(d + 1)) / ((d + 1) * (d + 2))
    Me = vol[:, None, None] * ref[None]
    rows = np.broadcast_to(cells[:, :, None], Ke.shape).ravel()
    cols = np.broadcast_to(cells[:, None, :], Ke.shape).ravel()
    keep = rows < n_owned
    rows, cols = rows[keep], cols[keep]
    import scipy.sparse as sp

    mass_m = sp.coo_matrix((Me.ravel()[keep], (rows, cols)), shape=(n_owned, n_local)).tocsr()
    stiff_m = sp.coo_matrix((Ke.ravel()[keep], (rows, cols)), shape=(n_owned, n_local)).tocsr()
    mass_m.sort_indices()
    stiff_m.sort_indices()
    if not (np.array_equal(mass_m.indptr, stiff_m.indptr) and np.array_equal(mass_m.indices, stiff_m.indices)):
        raise AssertionError("mass and stiffness patterns differ")
    return mass_m.indptr.astype(np.int64), mass_m.indices.astype(np.int32), mass_m.data, stiff_m.data


def _gauss_simplex(d: int, degree: int):
    if degree <= 1:  # the one-point (centroid) rule, what FFCx picks for an integrand of estimated degree <= 1
        return np.full((1, d + 1), 1.0 / (d + 1)), np.ones(1)
    m = degree // 2 + d
    xg, wg = np.polynomial.legendre.leggauss(m)
    xg, wg = 0.5 * (xg + 1.0), 0.5 * wg
    if d == 1:
        return np.stack([1.0 - xg, xg], axis=1), wg
    if d == 2:
        u, v = np.meshgrid(xg, xg, indexing="ij")
        wu, wv = np.meshgrid(wg, wg, indexing="ij")
        l1, l2 = u.ravel(), (v * (1 - u)).ravel()
        return np.stack([1 - l1 - l2, l1, l2], axis=1), (wu * wv * (1 - u)).ravel() * 2.0
    u, v, w = np.meshgrid(xg, xg, xg, indexing="ij")
    wu, wv, ww = np.meshgrid(wg, wg, wg, indexing="ij")
    l1, l2, l3 = u.ravel(), (v * (1 - u)).ravel(), (w * (1 - u) * (1 - v)).ravel()
    return np.stack([1 - l1 - l2 - l3, l1, l2, l3], axis=1), (wu * wv * ww * (1 - u) ** 2 * (1 - v)).ravel() * 6.0


def load_vector(mesh: Mesh, measure: Measure | None, marker: int | None, g: Callable | None = None, degree: int = 6) -> np.ndarray:
    """s_i = int g(x) phi_i dz(marker) for OWNED dofs i (g == None: g = 1).  Length size_local."""
    d = mesh.topology.dim
    n_owned = mesh.index_map.size_local
    out = np.zeros(mesh.num_local_vertices)
    kind = measure.kind if measure is not None else "dx"
    tags = measure.subdomain_data if measure is not None else None
    if marker is None and measure is not None:
        marker = measure.marker
    if kind == "dx":
        if marker is None or tags is None:
            ids = np.arange(mesh.num_cells)
        else:
            if tags.dim != d:
                raise ValueError("dx measure needs cell tags")
            ids = tags.find(marker)
        ents = mesh.cells_of(ids)
    else:
        if tags is None or marker is None:
            ents = mesh.boundary_facets()
        else:
            ents = tags.facets[tags.find(marker)]
    if ents.shape[0] == 0:
        return out[:n_owned]
    k = ents.shape[1]
    x = mesh.geometry.x[:, : mesh.gdim][ents]
    if k == 1:
        meas = np.ones(ents.shape[0])
    elif k == 2:
        meas = np.linalg.norm(x[:, 1] - x[:, 0], axis=1)
    elif k == 3:
        a, b = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
        if a.shape[1] == 2:
            meas = 0.5 * np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])
        else:
            meas = 0.5 * np.linalg.norm(np.cross(a, b), axis=1)
    else:
        meas, _ = _simplex_measure_and_gradients(x)
    if g is None:
        np.add.at(out, ents.ravel(), np.repeat(meas / k, k))
    else:
        bary, wq = _gauss_simplex(k - 1, degree) if k > 1 else (np.ones((1, 1)), np.ones(1))
        x3 = mesh.geometry.x[ents]
        for q in range(bary.shape[0]):
            xq = np.einsum("a,eag->eg", bary[q], x3)
            gq = np.asarray(g(xq.T), dtype=np.float64) * wq[q] * meas
            for a in range(k):
                np.add.at(out, ents[:, a], gq * bary[q, a])
    return out[:n_owned]


def point_probe(mesh: Mesh, point: Sequence[float]):
    """(local vertex ids, barycentric weights) of the cell containing ``point`` (P1 evaluation, what
    scifem.evaluate_function does in demos/niederer_benchmark.py:285); None if no local cell has it."""
    d = mesh.topology.dim
    p = np.asarray(point, dtype=np.float64)[:d]
    X = mesh.geometry.x[:, :d]
    if isinstance(mesh, BoxMesh):
        # candidates by index arithmetic: the Kuhn tets of the (at most 8) cubes around the point that are local -
        # a 10^7-dof rank must not materialise its 6 x 10^7 cells for nine probes
        rel = [(p[k] - mesh.p0[k]) / mesh.h[k] for k in range(3)]
        if any(r < -1e-9 or r > mesh.n[k] + 1e-9 for k, r in enumerate(rel)):
            return None
        spans = []
        for k, r in enumerate(rel):
            base = int(np.floor(r + 1e-9))
            lo_c, hi_c = (mesh.c0, mesh.c1 - 1) if k == 0 else (0, mesh.n[k] - 1)
            spans.append([c for c in {base - 1, base} if lo_c <= c <= hi_c and abs(r - np.clip(r, c, c + 1)) <= 1e-9])
        w_, ncube = mesh.c1 - mesh.c0, mesh.num_cubes
        cubes = [(k_ * mesh.n[1] + j_) * w_ + (i_ - mesh.c0) for i_ in spans[0] for j_ in spans[1] for k_ in spans[2]]
        cand_ids = np.sort(np.array([q * ncube + c for c in cubes for q in range(6)], dtype=np.int64))
        cand_cells = mesh.cells_of(cand_ids) if cand_ids.size else np.zeros((0, 4), dtype=np.int64)
    else:
        cells = mesh.cells
        cxa = X[cells]  # (ncell, d+1, d)
        lo, hi = cxa.min(axis=1), cxa.max(axis=1)
        cand_ids = np.nonzero(np.all((p >= lo - 1e-12) & (p <= hi + 1e-12), axis=1))[0]
        cand_cells = cells[cand_ids]
    best = None
    for verts in cand_cells:  # (in cell order: ties go to the lowest cell id on both routes)
        cx = X[verts]
        T = (cx[1:] - cx[0]).T
        lam = np.linalg.solve(T, p - cx[0])
        w = np.concatenate([[1.0 - lam.sum()], lam])
        if w.min() >= -1e-10 and (best is None or w.min() > best[1].min()):
            best = (verts.astype(np.int32), w)
    if best is None:
        return None
    nodes, w = best
    keep = np.abs(w) > 1e-14
    return nodes[keep], w[keep]


_ = math
