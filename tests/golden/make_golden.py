"""Generates the committed golden vectors of the split-step path from the CPU oracle.

    python tests/golden/make_golden.py

The reference itself (dolfinx / PETSc / gotranx) cannot be imported in this image, so the vectors come from
the oracle restatement (oracle/), which is pinned on the reference's own known-answer tests
(tests/test_oracle_known_answers.py) and on the published Niederer activation times
(tests/test_oracle_niederer.py).  Files (all float64, seeds fixed):
  ode_<model>_<scheme>.npz   states_in (ns, 16), params, t, dt, states_out            one cell-model step
  pde_slab.npz               CSR mass/stiff of a 4x3x2 Kuhn box, v_prev, stimulus, theta, dt, v_out (LU solve)
  split_slab.npz             12 Godunov and 12 Strang split steps (TP06 GRL1 + CN diffusion + stimulus) on that box
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _problems as P  # noqa: E402
from oracle import fem  # noqa: E402
from oracle import monodomain as om  # noqa: E402


def ode_vectors():
    for tag in ("fhn", "tp06", "torord"):
        m = P.oracle_model(tag)
        rng = np.random.default_rng(20240607)
        s = P.perturbed_states(m, 16, rng, P.V_NAME[tag])
        prm = m.init_parameter_values()
        for scheme in ("forward_explicit_euler", "generalized_rush_larsen"):
            with np.errstate(all="ignore"):
                out = getattr(m, scheme)(s, 3.25, 0.01, prm)
            np.savez_compressed(os.path.join(HERE, f"ode_{tag}_{scheme}.npz"), states_in=s, params=prm, t=3.25, dt=0.01,
                                states_out=out)


def slab():
    pts, cells = fem.box_mesh((4, 3, 2), (0, 0, 0), (2.0, 1.5, 1.0))
    sl, st = P.niederer_conductivities()
    mass, stiff = fem.assemble_p1(pts, cells, np.diag([sl, st, st]))
    stim_cells = fem.cells_all_vertices(pts, cells, lambda x: (x[0] <= 1.0 + 1e-10) & (x[1] <= 1.0 + 1e-10))
    load = fem.load_vector_cells(pts, cells, stim_cells)
    return pts, cells, mass, stiff, load


def pde_vector():
    pts, cells, mass, stiff, load = slab()
    rng = np.random.default_rng(7)
    v_prev = -85.0 + 120.0 * rng.random(pts.shape[0])
    model = om.MonodomainModel(mass, stiff, [om.Stimulus.window(load, 0.0, 2.0, 0.35714285714285715)], C_m=0.01, theta=0.5, solver="lu")
    model.v_[:] = v_prev
    model.step((0.5, 0.55))
    np.savez_compressed(os.path.join(HERE, "pde_slab.npz"), indptr=mass.indptr.astype(np.int64), indices=mass.indices.astype(np.int32),
                        mass=mass.data, stiff=stiff.data, load=load, amp=0.35714285714285715, t_start=0.0, t_end=2.0, C_m=0.01,
                        theta=0.5, t0=0.5, t1=0.55, v_prev=v_prev, v_out=model.state.copy())


def split_vector():
    pts, cells, mass, stiff, load = slab()
    m = P.oracle_model("tp06")
    prm = m.init_parameter_values(stim_amplitude=0.0)
    y0 = m.init_state_values()
    out = {}
    for name, theta in (("godunov", 1.0), ("strang", 0.5)):
        pde = om.MonodomainModel(mass, stiff, [om.Stimulus.window(load, 0.0, 2.0, 0.35714285714285715)], C_m=0.01, theta=0.5, solver="lu")
        ode = om.ODESolver(v_pde=pde.state, init_states=y0, parameters=prm, fun=m.generalized_rush_larsen, num_states=len(y0),
                           v_index=m.state_index("V"))
        sol = om.SplittingSolver(pde, ode, theta=theta)
        t, dt = 0.0, 0.05
        for _ in range(12):
            sol.step((t, t + dt))
            t += dt
        out[f"v_{name}"] = pde.state.copy()
        out[f"states_{name}"] = ode.values.copy()
    np.savez_compressed(os.path.join(HERE, "split_slab.npz"), params=prm, y0=y0, dt=0.05, nsteps=12, **out)


if __name__ == "__main__":
    ode_vectors()
    pde_vector()
    split_vector()
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
