"""The Niederer-benchmark set-up of demos/niederer_benchmark.py:44-230, written against this package
(BASELINE.json config 2: slab 20x7x3 mm, TP06 generalized Rush-Larsen, Crank-Nicolson diffusion,
Godunov splitting, S1 stimulus 50 000 uA/cm^3 for 2 ms in the [0,1.5]^3 mm corner)."""

from __future__ import annotations

import numpy as np

from . import conductivities, fem, geometry, stimulation
from .models import tp06
from .monodomain_model import MonodomainModel
from .monodomain_solver import MonodomainSplittingSolver
from .odesolver import DolfinODESolver
from .telemetry import NullMonitor

IC = {  # demos/niederer_benchmark.py:45-65
    "V": -85.23, "Xr1": 0.00621, "Xr2": 0.4712, "Xs": 0.0095, "m": 0.00172, "h": 0.7444, "j": 0.7045,
    "d": 3.373e-05, "f": 0.7888, "f2": 0.9755, "fCass": 0.9953, "s": 0.999998, "r": 2.42e-08,
    "Ca_i": 0.000126, "R_prime": 0.9073, "Ca_SR": 3.64, "Ca_ss": 0.00036, "Na_i": 8.604, "K_i": 136.89,
}
POINTS = {  # demos/niederer_benchmark.py:233-243
    "P1": (0, 0, 0), "P2": (0, 7, 0), "P3": (20, 0, 0), "P4": (20, 7, 0), "P5": (0, 0, 3), "P6": (0, 7, 3),
    "P7": (20, 0, 3), "P8": (20, 7, 3), "P9": (10, 3.5, 1.5),
}


def setup(dx: float = 0.2, comm=None, rtol: float | None = None, initial_guess_previous: bool = False, monitor=None,
          L=(20.0, 7.0, 3.0), probes: bool = True, scheme: str = "generalized_rush_larsen", ksp_type: str = "cg",
          pc_type: str | None = None):
    """Returns (solver, info).  rtol None = PETSc's default 1e-5 as the demo runs it (:182-188)."""
    comm = comm or fem.COMM_SELF
    monitor = monitor or NullMonitor()
    geo = geometry.get_3D_slab_geometry(comm=comm, Lx=L[0], Ly=L[1], Lz=L[2], dx=dx)
    mesh = geo.mesh
    cond = conductivities.default_conductivities("Niederer")
    time = fem.Constant(mesh, 0.0)
    tol, Ls = 1.0e-10, 1.5
    cells = fem.locate_entities(mesh, mesh.topology.dim, lambda x: (x[0] <= Ls + tol) & (x[1] <= Ls + tol) & (x[2] <= Ls + tol))
    tags = fem.meshtags(mesh, mesh.topology.dim, cells, np.full(len(cells), 1, dtype=np.int32))
    I_s = stimulation.define_stimulus(mesh=mesh, chi=cond["chi"], time=time, subdomain_data=tags, marker=1, mesh_unit="mm",
                                      amplitude=50_000.0)
    M = conductivities.define_conductivity_tensor(f0=geo.f0, **cond)
    opts = {"ksp_type": ksp_type, "pc_type": "hypre", "pc_hypre_type": "boomeramg"}  # as the demo asks (:182-188)
    if pc_type is not None:
        opts["pc_type"] = pc_type
    if rtol is not None:
        opts["ksp_rtol"] = rtol
    params = {"petsc_options": opts, "initial_guess_previous": initial_guess_previous}
    C_m = 1.0 * 1e-2  # 1 uF/cm^2 in uF/mm^2 (:136,201)
    pde = MonodomainModel(time=time, mesh=mesh, M=M, I_s=I_s, params=params, C_m=C_m, dx=I_s.dZ, monitor=monitor)
    fun = getattr(tp06, scheme)
    init_states = tp06.init_state_values(**IC)
    parameters = tp06.init_parameter_values(stim_amplitude=0.0)
    ode = DolfinODESolver(v_ode=fem.Function(fem.functionspace(mesh, ("Lagrange", 1))), v_pde=pde.state, fun=fun,
                          init_states=init_states, parameters=parameters, num_states=len(init_states),
                          v_index=tp06.state_index("V"), monitor=monitor)
    solver = MonodomainSplittingSolver(pde=pde, ode=ode, monitor=monitor)
    probe_ids = {}
    if probes:
        for name, p in POINTS.items():
            scaled = tuple(c * l / l0 for c, l, l0 in zip(p, L, (20.0, 7.0, 3.0)))
            probe_ids[name] = pde.add_probe(scaled)
        pde.track_activation(0.0)
    n_global = mesh.index_map.size_global
    return solver, {"mesh": mesh, "n_global": n_global, "n_owned": mesh.index_map.size_local, "probe_ids": probe_ids,
                    "num_states": len(init_states)}
