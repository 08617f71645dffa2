"""ORACLE (test infrastructure, never a product path): the Niederer benchmark run of
/root/reference/demos/niederer_benchmark.py:44-289 on the CPU, using the oracle's assembly (oracle/fem.py) and
the OpenMP C restatement of the split step (oracle/c/oracle_step.c).

Activation time of a point = start time t of the first step after which v(p) > 0 (:281-289); the centre point
P9 is evaluated by P1 interpolation inside its cell (:285, scifem.evaluate_function).  The published table
(:315-325) is the only reference output that exercises the gotranx-generated TP06 step, so this run is what
pins the cell-model part of the oracle (to the table's resolution: one dt).
"""

from __future__ import annotations

import numpy as np

from . import cport, fem

POINTS = {"P1": (0, 0, 0), "P2": (0, 7, 0), "P3": (20, 0, 0), "P4": (20, 7, 0), "P5": (0, 0, 3), "P6": (0, 7, 3),
          "P7": (20, 0, 3), "P8": (20, 7, 3), "P9": (10, 3.5, 1.5)}  # niederer_benchmark.py:233-243
PUBLISHED = {  # niederer_benchmark.py:317-325, (dx, dt) -> P1..P9 in ms
    (0.5, 0.05): [1.25, 51.1, 34.9, 58.9, 14.1, 49.5, 34, 56.65, 26.05],
    (0.5, 0.01): [1.22, 50.85, 33.96, 58.05, 13.98, 49.36, 33.07, 55.91, 25.64],
    (0.5, 0.005): [1.215, 50.775, 33.825, 57.96, 13.97, 49.345, 32.945, 55.825, 25.595],
    (0.2, 0.05): [1.25, 29.7, 32.9, 40.2, 9.55, 30, 32.95, 39.9, 18.9],
    (0.2, 0.01): [1.24, 29.09, 31.25, 38.66, 9.34, 29.4, 31.29, 38.42, 18.14],
    (0.2, 0.005): [1.235, 29.015, 31.05, 38.475, 9.315, 29.32, 31.08, 38.235, 18.045],
    (0.1, 0.05): [1.25, 26.85, 33.3, 40.35, 8.4, 27.5, 33.85, 40.55, 18.95],
    (0.1, 0.01): [1.23, 25.64, 31.46, 38.08, 8.03, 26.24, 31.94, 38.21, 17.95],
    (0.1, 0.005): [1.225, 25.5, 31.26, 37.81, 7.99, 26.09, 31.72, 37.93, 17.835],
}


def conductivities():
    """Niederer set -> (s_l, s_t) in uA/mV (conductivities.py:31-37,82-93): harmonic mean of intra/extra, over chi."""
    chi_per_m = 1400.0 * 100.0
    sl = 0.17 * 0.62 / (0.17 + 0.62)
    st = 0.019 * 0.24 / (0.019 + 0.24)
    return sl / chi_per_m * 1e3, st / chi_per_m * 1e3


def probe_weights(pts, cells, point):
    """(nodes, barycentric weights) of `point` in the mesh (first cell that contains it)."""
    p = np.asarray(point, dtype=float)
    d = np.abs(pts - p).max(axis=1)
    k = int(np.argmin(d))
    if d[k] < 1e-12:
        return np.array([k]), np.array([1.0])
    x = pts[cells]  # (nc, 4, 3)
    T = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
    lam = np.linalg.solve(T, (np.broadcast_to(p, (x.shape[0], 3)) - x[:, 0, :])[..., None])[..., 0]
    bary = np.concatenate([1.0 - lam.sum(axis=1, keepdims=True), lam], axis=1)
    inside = np.nonzero((bary >= -1e-12).all(axis=1))[0]
    c = int(inside[0])
    return cells[c], bary[c]


def run(dx: float, dt: float, T: float = 100.0, tp06=None, rtol: float = 1e-5, progress=None):
    """Returns {point: activation time} (-1 when not activated by T)."""
    n = tuple(int(np.rint(l / dx)) for l in (20.0, 7.0, 3.0))  # geometry.py:130-132
    pts, cells = fem.box_mesh(n, (0, 0, 0), (20.0, 7.0, 3.0))
    sl, st = conductivities()
    mass, stiff = fem.assemble_p1(pts, cells, np.diag([sl, st, st]))
    tol = 1e-10
    stim_cells = fem.cells_all_vertices(pts, cells, lambda x: (x[0] <= 1.5 + tol) & (x[1] <= 1.5 + tol) & (x[2] <= 1.5 + tol))
    load = fem.load_vector_cells(pts, cells, stim_cells)
    amp = 50000.0 / 1400.0 * 1e-2
    sp = cport.SplitProblem(mass, stiff, 0.01, 0.5, dt, [(load, 0.0, 2.0, amp)], rtol=rtol)
    params = tp06.init_parameter_values(stim_amplitude=0.0)
    y0 = tp06.init_state_values()
    nn = pts.shape[0]
    states = np.ascontiguousarray(np.repeat(y0[:, None], nn, axis=1))
    vidx = tp06.state_index("V")
    probes = {k: probe_weights(pts, cells, p) for k, p in POINTS.items()}
    act = {k: -1.0 for k in POINTS}
    t = 0.0
    nsteps = int(round(T / dt))
    for k in range(nsteps):
        v, _ = sp.split_steps("tp06", "generalized_rush_larsen", vidx, states, params, t, 1)
        for name, (nodes, w) in probes.items():
            if act[name] < 0 and float(v[nodes] @ w) > 0.0:
                act[name] = t
        t += dt
        if progress and k % progress == 0:
            print(k, t, act, flush=True)
        if all(a >= 0 for a in act.values()):
            break
    return act
