"""Slab geometries (src/beat/geometry.py:9-218): structured meshes + constant fibre fields."""

from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import fem


class Geometry(NamedTuple):
    mesh: fem.Mesh
    ffun: fem.MeshTags | None = None
    markers: dict | None = None
    f0: fem.Constant | None = None
    s0: fem.Constant | None = None
    n0: fem.Constant | None = None


def get_2D_slab_microstructure(mesh, transverse: bool = False):
    if transverse:
        return fem.Constant(mesh, (0.0, 1.0)), fem.Constant(mesh, (1.0, 0.0))
    return fem.Constant(mesh, (1.0, 0.0)), fem.Constant(mesh, (0.0, 1.0))


def get_3D_slab_microstructure(mesh, transverse: bool = False):
    if transverse:
        return fem.Constant(mesh, (0.0, 0.0, 1.0)), fem.Constant(mesh, (1.0, 0.0, 0.0)), fem.Constant(mesh, (0.0, 1.0, 0.0))
    return fem.Constant(mesh, (1.0, 0.0, 0.0)), fem.Constant(mesh, (0.0, 1.0, 0.0)), fem.Constant(mesh, (0.0, 0.0, 1.0))


def get_2D_slab_mesh(comm, dx: float, Lx: float, Ly: float):
    nx, ny = int(np.rint(Lx / dx)), int(np.rint(Ly / dx))
    return fem.create_rectangle(comm, [np.array([0.0, 0.0]), np.array([Lx, Ly])], [nx, ny])


def get_3D_slab_mesh(comm, dx: float, Lx: float, Ly: float, Lz: float):
    nx, ny, nz = int(np.rint(Lx / dx)), int(np.rint(Ly / dx)), int(np.rint(Lz / dx))  # geometry.py:130-132
    return fem.create_box(comm, [np.array([0.0, 0.0, 0.0]), np.array([Lx, Ly, Lz])], [nx, ny, nz])


def get_2D_slab_geometry(comm, Lx: float, Ly: float, dx: float, transverse: bool = False) -> Geometry:
    mesh = get_2D_slab_mesh(comm, dx, Lx, Ly)
    f0, s0 = get_2D_slab_microstructure(mesh, transverse)
    return Geometry(mesh=mesh, f0=f0, s0=s0)


def get_3D_slab_geometry(comm, Lx: float, Ly: float, Lz: float, dx: float, transverse: bool = False) -> Geometry:
    mesh = get_3D_slab_mesh(comm, dx, Lx, Ly, Lz)
    f0, s0, n0 = get_3D_slab_microstructure(mesh, transverse)
    return Geometry(mesh=mesh, f0=f0, s0=s0, n0=n0)
