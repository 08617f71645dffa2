"""Shared problem builders for the tests (oracle side: meshes, matrices, Niederer constants)."""
import functools
import importlib

import numpy as np

from oracle import fem

# Niederer set, mm / ms units (src/beat/conductivities.py:31-37,82-93)


def niederer_conductivities():
    """sigma = g_i*g_e/(g_i+g_e) [S/m] / chi [1/m]  -> uA/mV (= S*mm /mm... in the mm/ms/mV/uA system)."""
    chi_per_m = 1400.0 * 100.0
    sl = 0.17 * 0.62 / (0.17 + 0.62)  # S/m
    st = 0.019 * 0.24 / (0.019 + 0.24)
    # S/m / (1/m) = S = A/V ; in uA/mV: 1 A/V = 1e6 uA / 1e3 mV = 1e3 uA/mV
    return sl / chi_per_m * 1e3, st / chi_per_m * 1e3


@functools.lru_cache(maxsize=4)
def niederer_slab(dx: float, L=(20.0, 7.0, 3.0)):
    n = tuple(int(np.rint(l / dx)) for l in L)
    pts, cells = fem.box_mesh(n, (0, 0, 0), L)
    sl, st = niederer_conductivities()
    M = np.diag([sl, st, st])
    mass, stiff = fem.assemble_p1(pts, cells, M)
    tol = 1e-10
    stim_cells = fem.cells_all_vertices(pts, cells, lambda x: (x[0] <= 1.5 + tol) & (x[1] <= 1.5 + tol) & (x[2] <= 1.5 + tol))
    load = fem.load_vector_cells(pts, cells, stim_cells)
    amp = 50000.0 / 1400.0 * 1e-2  # 50000 uA/cm^3 / (1400 /cm) = 35.714 uA/cm^2 = 0.35714 uA/mm^2
    return dict(pts=pts, cells=cells, mass=mass, stiff=stiff, stim_load=load, stim_amp=amp, C_m=0.01, n=n)


def oracle_model(tag):
    return importlib.import_module(f"oracle.models.{tag}")


def perturbed_states(om, n, rng, vname):
    y0 = om.init_state_values()
    s = np.repeat(y0[:, None], n, axis=1) * (1 + 1e-3 * rng.uniform(-1, 1, (len(y0), n)))
    s[om.state_index(vname)] = rng.uniform(-90.0, 40.0, n)
    return np.ascontiguousarray(s)


V_NAME = {"fhn": "v", "tp06": "V", "torord": "v"}
MODEL_ID = {"fhn": 0, "tp06": 1, "torord": 2}
SCHEME_ID = {"forward_explicit_euler": 0, "generalized_rush_larsen": 1}
