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


def get_lv_ellipsoid_geometry(comm, n_r: int = 4, n_mu: int = 24, n_phi: int = 32, fiber_angle_endo: float = 60.0,
                              fiber_angle_epi: float = -60.0, **radii) -> Geometry:
    """Synthetic LV shell with the markers and the fibre field the reference's demos/lv_endocardial.py takes from
    cardiac_geometries (:35-75): facet tags ENDO / EPI / BASE / APEX (the cut), a vertex function ``endo_epi`` in info
    (1 endo / 2 mid / 3 epi third of the wall, for DolfinMultiODESolver), and a CELL-WISE fibre direction whose helix
    angle turns linearly from fiber_angle_endo to fiber_angle_epi across the wall (f0: (ncell, 3) array)."""
    mesh = fem.create_lv_ellipsoid(comm, n_r, n_mu, n_phi, **radii)
    decode = mesh.info["decode"]
    l2g = mesh.index_map.local_to_global
    a, b, _ = decode(l2g)
    markers = {"ENDO": (1, 2), "EPI": (2, 2), "BASE": (3, 2), "APEX": (4, 2)}
    fac = mesh.boundary_facets()
    fa, fb = a[fac], b[fac]
    val = np.zeros(fac.shape[0], dtype=np.int32)
    val[np.all(fb == 0, axis=1)] = 4
    val[np.all(fb == n_mu, axis=1)] = 3
    val[np.all(fa == n_r, axis=1)] = 2
    val[np.all(fa == 0, axis=1)] = 1
    ffun = fem.meshtags(mesh, 2, np.arange(fac.shape[0], dtype=np.int32), val)
    layer = np.minimum(3 * a // (n_r + 1), 2) + 1  # vertex layer 1..3
    mesh.info["endo_epi"] = layer.astype(np.float64)
    # fibres at the cell centroids: circumferential / longitudinal frame of the local ellipsoid coordinates
    # (component arrays throughout: row-wise reductions over (ncell, 3) temporaries cost several times more in NumPy)
    cells, xv = mesh.cells, mesh.geometry.x
    corners = [np.ascontiguousarray(cells[:, q]) for q in range(4)]
    X, Y, Z = (0.25 * (xv[corners[0], d] + xv[corners[1], d] + xv[corners[2], d] + xv[corners[3], d]) for d in range(3))
    lam = (a[corners[0]] + a[corners[1]] + a[corners[2]] + a[corners[3]]) * (0.25 / n_r)
    inv = 1.0 / np.sqrt(Y * Y + Z * Z)
    ephi_y, ephi_z = -Z * inv, Y * inv  # e_phi = (0, -z, y) / |.|
    # surface normal of the ellipsoid through the point (gradient of x^2/rl^2 + (y^2+z^2)/rs^2), then e_mu = n x e_phi
    rs = radii.get("r_short_endo", 2.5) + lam * (radii.get("r_short_epi", 3.5) - radii.get("r_short_endo", 2.5))
    rl = radii.get("r_long_endo", 9.0) + lam * (radii.get("r_long_epi", 9.7) - radii.get("r_long_endo", 9.0))
    nx_, ny_, nz_ = X / rl**2, Y / rs**2, Z / rs**2
    inv = 1.0 / np.sqrt(nx_ * nx_ + ny_ * ny_ + nz_ * nz_)
    nx_, ny_, nz_ = nx_ * inv, ny_ * inv, nz_ * inv
    emu = (ny_ * ephi_z - nz_ * ephi_y, -nx_ * ephi_z, nx_ * ephi_y)
    alpha = np.deg2rad(fiber_angle_endo + lam * (fiber_angle_epi - fiber_angle_endo))
    ca, sa = np.cos(alpha), np.sin(alpha)
    fx, fy, fz = sa * emu[0], ca * ephi_y + sa * emu[1], ca * ephi_z + sa * emu[2]
    inv = 1.0 / np.sqrt(fx * fx + fy * fy + fz * fz)
    f0 = np.stack([fx * inv, fy * inv, fz * inv], axis=1)
    nrm = np.stack([nx_, ny_, nz_], axis=1)
    return Geometry(mesh=mesh, ffun=ffun, markers=markers, f0=f0, n0=nrm)
