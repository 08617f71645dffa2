"""The set-up of demos/lv_endocardial.py:35-294 on the synthetic LV shell of this package (BASELINE.json config 5):
Bishop conductivities with a transmurally rotating cell-wise fibre field, endocardial SURFACE stimulus
(2000 uA/cm^2 / chi for 1 ms on the ENDO facets, :260-270), three transmural layers with their own parameter sets
(DolfinMultiODESolver, :203-257), Godunov splitting.  The demo uses ToR-ORd endo/mid/epi variants; the layers here
carry TP06 (the BASELINE config names TP06) with g_Ks / g_to scaled per layer as in the epi / M / endo TP06 variants."""

from __future__ import annotations

from . import conductivities, fem, geometry, stimulation
from .models import tp06
from .monodomain_model import MonodomainModel
from .monodomain_solver import MonodomainSplittingSolver
from .odesolver import DolfinMultiODESolver
from .telemetry import NullMonitor

# ten Tusscher & Panfilov 2006, Table 1: the cell types differ in g_Ks and g_to (epi values are the .ode defaults)
LAYER_PARAMETERS = {1: dict(g_Ks=0.392, g_to=0.073), 2: dict(g_Ks=0.098, g_to=0.294), 3: dict()}  # endo, mid, epi


def setup(n=(4, 24, 32), comm=None, rtol: float | None = None, ksp_type: str = "auto", pc_type: str | None = None, monitor=None,
          initial_guess_previous: bool = False):
    comm = comm or fem.COMM_SELF
    monitor = monitor or NullMonitor()
    geo = geometry.get_lv_ellipsoid_geometry(comm, *n)
    mesh = geo.mesh
    cond = conductivities.default_conductivities("Bishop")
    M = conductivities.define_conductivity_tensor(f0=geo.f0, **cond)  # (ncell, 3, 3)
    time = fem.Constant(mesh, 0.0)
    I_s = stimulation.define_stimulus(mesh=mesh, chi=cond["chi"], time=time, subdomain_data=geo.ffun, marker=geo.markers["ENDO"][0],
                                      amplitude=2000.0, mesh_unit="cm", start=0.0, duration=1.0)
    opts = {"ksp_type": ksp_type, "pc_type": pc_type or "hypre"}
    if rtol is not None:
        opts["ksp_rtol"] = rtol
    pde = MonodomainModel(time=time, mesh=mesh, M=M, I_s=I_s, C_m=1.0, dx=None, monitor=monitor,
                          params={"petsc_options": opts, "initial_guess_previous": initial_guess_previous})
    V = fem.functionspace(mesh, ("P", 1))
    markers = fem.Function(V)
    markers.x.array[:] = mesh.info["endo_epi"]
    init = {k: tp06.init_state_values() for k in LAYER_PARAMETERS}
    params = {k: tp06.init_parameter_values(stim_amplitude=0.0, **kw) for k, kw in LAYER_PARAMETERS.items()}
    fun = {k: tp06.generalized_rush_larsen for k in LAYER_PARAMETERS}
    vi = tp06.state_index("V")
    ode = DolfinMultiODESolver(v_ode=fem.Function(V), v_pde=pde.state, markers=markers, init_states=init, parameters=params, fun=fun,
                               num_states={k: len(init[k]) for k in init}, v_index={k: vi for k in init}, monitor=monitor)
    solver = MonodomainSplittingSolver(pde=pde, ode=ode, monitor=monitor)
    info = {"mesh": mesh, "geo": geo, "M": M, "n_global": mesh.index_map.size_global, "n_owned": mesh.index_map.size_local,
            "num_states": len(init[1]), "layer_parameters": params, "stim_amplitude": I_s.expr.amplitude_now()}
    return solver, info
