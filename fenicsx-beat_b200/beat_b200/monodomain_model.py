"""MonodomainModel on the device: the reference's BaseModel + MonodomainModel
(src/beat/base_model.py:48-297, src/beat/monodomain_model.py:17-98) with the same constructor, attributes
and step/solve semantics, talking to libmono_b200 through ctypes.

Discrete problem (derived from monodomain_model.py:83-96, a, L = ufl.system(G)):
    A = C_m*Mass + dt*theta*K            b = (C_m*Mass - dt*(1-theta)*K) v_ + dt * sum_k I_k(t0+theta*dt) s_k
Mass/K are assembled once on the host (fem.assemble_p1_local) and handed to the device as CSR; the
re-assembly the reference does every step (_update_rhs) becomes one SpMV in the persistent PDE kernel.
"""

from __future__ import annotations

import logging
import math
from enum import Enum, auto
from typing import Any, NamedTuple, Sequence

import numpy as np

from . import fem
from ._lib import Context
from .stimulation import Stimulus
from .telemetry import BaseMonitor, NullMonitor

logger = logging.getLogger(__name__)

PC = {"none": 0, "jacobi": 1, "chebyshev": 2}
NORM = {"preconditioned": 0, "unpreconditioned": 1, "natural": 2, "default": 0}


class Status(str, Enum):
    OK = auto()
    NOT_CONVERGING = auto()


class Results(NamedTuple):
    state: fem.Function
    status: Status


def _transform_I_s(I_s, dZ: fem.Measure) -> list[Stimulus]:  # base_model.py:33-45
    if I_s is None:
        return []
    if isinstance(I_s, fem.ExprSum):  # a + b of source expressions (stimulation.generate_random_activation)
        return [Stimulus(expr=e, dZ=dZ) for e in I_s.terms]
    if isinstance(I_s, Stimulus):
        if isinstance(I_s.expr, fem.ExprSum):
            return [Stimulus(expr=e, dZ=I_s.dZ, marker=I_s.marker) for e in I_s.expr.terms]
        return [I_s]
    if isinstance(I_s, (fem.TimeWindow, fem.TimeFunction, fem.Constant, float, int)):
        return [Stimulus(expr=I_s, dZ=dZ)]
    out: list[Stimulus] = []
    for s in I_s:
        out.extend(_transform_I_s(s, dZ))
    return out


class DeviceKSP:
    """What monitor.record_ksp() receives: the PETSc KSP getters the reference's monitor calls
    (src/beat/telemetry.py:67-76), answered from the device CG's result block."""

    def __init__(self, ctx: Context):
        self._ctx = ctx

    def _info(self):
        return self._ctx.ksp_info()

    def getIterationNumber(self) -> int:
        return self._info()[0]

    def getResidualNorm(self) -> float:
        return self._info()[1]

    def getConvergedReason(self) -> int:
        return self._info()[2]


class MonodomainModel:
    r"""Solve  C_m dV/dt - div(M grad V) - I_stim = 0  with the theta-rule, on one B200 per rank.

    Parameters as in the reference (monodomain_model.py:27-40, base_model.py:73-82).  ``M`` is a float, a
    :class:`fem.Constant`, a (d,d) array or a per-cell (ncell,d,d) array.  Solver selection follows
    ``params["petsc_options"]``: ksp_type "cg" | "pipecg" | "auto" -> device CG with ksp_rtol / ksp_atol / ksp_max_it /
    ksp_norm_type; pc_type "jacobi" | "none" | "chebyshev" (Chebyshev polynomial in the Jacobi-scaled operator,
    pc_chebyshev_steps / pc_chebyshev_kappa, pipecg only) are native; "hypre"/"gamg"/... map to Jacobi, or with
    ksp_type "auto" to whatever is fastest for the mesh size (the system is mass-dominated);
    ksp_type "preonly" (the reference's LU/MUMPS default) is reproduced by iterating the same CG to rtol 1e-12.
    """

    def __init__(self, time: fem.Constant, mesh: fem.Mesh, M, I_s=None, params: dict | None = None, C_m: float = 1.0,
                 dx: fem.Measure | None = None, monitor: BaseMonitor | None = None, **kwargs: Any) -> None:
        if kwargs:
            logger.warning("Unused keyword arguments: %s", ", ".join(f"{k}={v}" for k, v in kwargs.items()))
        self._mesh = mesh
        self.time = time
        self.dx = dx or fem.Measure("dx", domain=mesh)
        self.monitor = monitor or NullMonitor()
        self._M = M
        self.C_m = fem.Constant(mesh, C_m)
        self.parameters = type(self).default_parameters()
        if params is not None:
            self.parameters.update(params)
        self._I_s = _transform_I_s(I_s, dZ=self.dx)
        self._setup_state_space()
        self._timestep = fem.Constant(mesh, self.parameters["default_timestep"])
        self._setup_device()

    # ---- reference surface ---------------------------------------------------------------------
    @staticmethod
    def default_parameters(solver_type: str = "direct") -> dict[str, Any]:  # base_model.py:136-168
        if solver_type == "iterative":
            petsc_options = {"ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg"}
        else:
            petsc_options = {"ksp_type": "preonly", "pc_type": "lu", "pc_factor_mat_solver_type": "mumps"}
        return {
            "theta": 0.5,
            "degree": 1,
            "family": "Lagrange",
            "default_timestep": 1.0,
            "jit_options": {},
            "form_compiler_options": {},
            "petsc_options": petsc_options,
            "log_timings": False,
            "timing_log_frequency": 1,
            "use_custom_preconditioner": True,
            # device-only knob: start CG from v_ instead of zero (fewer iterations, same fixed point)
            "initial_guess_previous": False,
        }

    def _setup_state_space(self) -> None:  # monodomain_model.py:42-53
        self.V = fem.functionspace(self._mesh, (self.parameters["family"], self.parameters["degree"]))
        self.v_ = fem.Function(self.V, name="v_")
        self._state = fem.Function(self.V, name="v")

    @property
    def state(self) -> fem.Function:
        return self._state

    def _solver_settings(self):
        opts = self.parameters["petsc_options"] or {}
        ksp = str(opts.get("ksp_type", "preonly"))
        pc = str(opts.get("pc_type", "lu"))
        if ksp == "preonly":
            rtol, atol, max_it, pc_id = 1e-12, 1e-50, 10000, PC["jacobi"]
        elif ksp in ("cg", "pipecg", "auto"):
            rtol = float(opts.get("ksp_rtol", 1e-5))  # PETSc defaults
            atol = float(opts.get("ksp_atol", 1e-50))
            max_it = int(opts.get("ksp_max_it", 10000))
            pc_id = PC.get(pc, PC["jacobi"])
        else:
            raise NotImplementedError(f"ksp_type={ksp!r}: the device solvers are 'cg', 'pipecg', 'auto' (and 'preonly' = tight cg)")
        norm = NORM[str(opts.get("ksp_norm_type", "default"))]
        x0 = 1 if self.parameters.get("initial_guess_previous") or opts.get("ksp_initial_guess_nonzero") else 0
        if ksp == "auto":
            # same iterates in exact arithmetic; which driver is faster depends on where the CG vectors live:
            # one row per thread (everything in shared memory, latency bound) -> one reduction per iteration wins;
            # larger meshes stream from HBM -> KSPCG moves ~20 % fewer bytes per iteration
            n_sm = self._ctx.device_info()["n_sm"]
            per_rank = self._mesh.index_map.size_global / max(self._mesh.comm.size, 1)  # the same number on every rank
            ksp = "pipecg" if per_rank <= (n_sm - 1) * 512 else "cg"
            # Polynomial (Chebyshev-Jacobi) preconditioning trades extra dataflow-synchronised SpMVs (~1.6 us each) for fewer
            # reductions.  One GPU: a reduction costs ~2.1 us, so it does not pay (measured: 72 vs 63 us/step) and stays
            # opt-in.  Several GPUs: every reduction also crosses NVLink (+2.4 us) -> 3 steps win (65 vs 74 us/step at 2 GPUs).
            if ksp == "pipecg" and pc not in PC and self._mesh.comm.size > 1:
                pc_id = PC["chebyshev"]
        if pc_id == PC["chebyshev"] and ksp != "pipecg":
            raise NotImplementedError("pc_type 'chebyshev' is implemented in the pipelined driver: use ksp_type 'pipecg' or 'auto'")
        self._ksp_type = 1 if ksp == "pipecg" else 0  # MONO_KSP_PIPECG / MONO_KSP_CG
        self.ksp_type_used = ksp
        self.pc_type_used = {v: k for k, v in PC.items()}[pc_id]
        import os

        self._cheb = (int(opts.get("pc_chebyshev_steps", os.environ.get("MONO_CHEB_STEPS", 3))),
                      float(opts.get("pc_chebyshev_kappa", os.environ.get("MONO_CHEB_KAPPA", 4.0))))
        return rtol, atol, max_it, pc_id, norm, x0

    def _setup_device(self) -> None:
        mesh = self._mesh
        self._ctx = ctx = mesh.device_context()
        if getattr(ctx, "n_local", 0):
            self._ctx = ctx = mesh.new_device_context()
        imap = mesh.index_map
        indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, self._M)
        ctx.pde_set_matrices(imap.size_local, imap.num_ghosts, indptr, indices, mass, stiff)
        self._nnz_per_row = len(indices) / max(imap.size_local, 1)
        rtol, atol, max_it, pc_id, norm, x0 = self._solver_settings()
        ctx.pde_set_chebyshev(*self._cheb)  # before pde_config / set_halo: the number of exchange buffers depends on it
        ctx.pde_config(float(self.C_m), float(self.parameters["theta"]), rtol, atol, max_it, pc_id, norm, x0)
        ctx.pde_set_ksp_type(self._ksp_type)
        if mesh.comm.size > 1:
            from .dist import init_comm

            init_comm(ctx, mesh.comm)
            ctx.set_halo(imap.nbr_ranks, imap.send_ptr, imap.send_idx, imap.recv_ptr)
        self._stim_ids: list[int] = []
        self._stim_amp: list[float] = []
        for s in self._I_s:
            expr = s.expr
            g = expr.g if isinstance(expr, (fem.Separable, fem.WindowedField)) else None
            deg = expr.degree if isinstance(expr, (fem.Separable, fem.WindowedField)) else 0
            meas = s.dZ if s.dZ is not None else self.dx
            load = fem.load_vector(mesh, meas, s.marker, g, deg)
            idx = np.nonzero(load)[0].astype(np.int32)
            if isinstance(expr, fem.TimeWindow):
                t0, t1, amp = expr.start, expr.end, expr.amplitude_now()
            else:
                t0, t1, amp = -math.inf, math.inf, 0.0
            self._stim_ids.append(ctx.stim_add(idx, load[idx], t0, t1, amp))
            self._stim_amp.append(amp)
        # host mirrors of the two PDE vectors
        self._state.x.bind(ctx.get_v, ctx.set_v, push_now=False, sync=ctx.sync)
        self.v_.x.bind(ctx.get_v_prev, ctx.set_v_prev, push_now=False, sync=ctx.sync)
        self._state._owner = self
        self.ksp = DeviceKSP(ctx)

    def _push_stimulus_amplitudes(self) -> None:
        """Host-side part of the source term: current amplitudes (Constants the user re-assigns every
        step as in demos/pace_train.py:216-219, Stimulus.assign, or h(t) of a TimeFunction)."""
        for k, s in enumerate(self._I_s):
            expr = s.expr
            if isinstance(expr, (fem.TimeWindow, fem.TimeFunction)):
                amp = expr.amplitude_now()
            else:
                amp = float(expr)
            if amp != self._stim_amp[k]:
                self._ctx.stim_set_amplitude(self._stim_ids[k], amp)
                self._stim_amp[k] = amp

    def has_host_evaluated_sources(self) -> bool:
        """True when some source amplitude is a function of time the HOST evaluates each step (TimeFunction / Separable):
        such a source cannot be left alone on the device for several steps (MonodomainSplittingSolver.solve_on_device)."""
        return any(isinstance(s.expr, fem.TimeFunction) for s in self._I_s)

    def _flush_host(self) -> None:
        self._state.x.flush_to_device()
        self.v_.x.flush_to_device()

    def assign_previous(self) -> None:  # monodomain_model.py:59-60
        self._flush_host()
        self._ctx.pde_assign_previous()
        self.v_.x.mark_device_newer()

    def step(self, interval) -> None:  # base_model.py:208-245
        t0, t1 = interval
        dt = t1 - t0
        theta = self.parameters["theta"]
        t = t0 + theta * dt
        with self.monitor.track_time("pde_total_step"):
            with self.monitor.track_time("pde_set_time"):
                self.time.value = t
            if not abs(dt - float(self._timestep)) < 1.0e-12:
                self._timestep.value = dt  # the device rebuilds A, B inside mono_pde_step (K3)
            self._push_stimulus_amplitudes()
            self._flush_host()
            with self.monitor.track_time("pde_linear_solve"):
                self._ctx.pde_step(t0, t1)
            self._state.x.mark_device_newer()
            self.monitor.record_ksp(self.ksp)
        self.monitor.advance_step(t0, t1)

    def solve(self, interval, dt: float | None = None) -> Results:  # base_model.py:250-297
        T0, T = interval
        if dt is None:
            dt = T - T0
        t0 = T0
        t1 = T0 + dt
        while True:
            self.step((t0, t1))
            if (t1 + dt) > (T + 1e-12):
                break
            self.assign_previous()
            t0 = t1
            t1 = t0 + dt
        return Results(state=self.state, status=Status.OK)

    # ---- device extras ---------------------------------------------
    def add_probe(self, point: Sequence[float]) -> int | None:
        """Register a P1 point evaluation on the device; returns the probe id (None if the point is not
        in this rank's cells)."""
        hit = fem.point_probe(self._mesh, point)
        if hit is None:
            return None
        nodes, w = hit
        if not hasattr(self, "_n_probes"):
            self._n_probes = 0
        self._n_probes += 1
        return self._ctx.probe_add(nodes, w)

    def track_activation(self, threshold: float = 0.0) -> None:
        self._ctx.probe_activation(threshold)

    def observe(self, activation_map: bool = True, threshold: float = 0.0, minmax: bool = True) -> None:
        """Whole-field observers evaluated on the device after every split step (same launch as the probes): per-node
        activation times and min / max of v - what the demos compute from ``state.x.array`` on the host every step
        (demos/niederer_benchmark.py:271-287), without moving the field."""
        self._ctx.observe_config(activation_map, threshold, minmax)

    def activation_map(self) -> np.ndarray:
        """Start time of the first split step after which v > threshold, per owned dof (-1: not yet)."""
        return self._ctx.activation_map()

    def v_minmax(self) -> tuple[float, float]:
        """(min, max) of v over this rank's owned dofs after the last split step."""
        return self._ctx.v_minmax()

    def state_snapshot(self, stride: int = 1, offset: int = 0, count: int | None = None) -> np.ndarray:
        """v[offset::stride] of the owned dofs gathered on the device (coarse frames without moving the whole field)."""
        n = self._mesh.index_map.size_local
        if count is None:
            count = max(0, (n - offset + stride - 1) // stride)
        self._flush_host()
        return self._ctx.get_v_strided(offset, stride, count)

    def probe_values(self) -> np.ndarray:
        return self._ctx.probe_values(getattr(self, "_n_probes", 0))

    def activation_times(self) -> np.ndarray:
        return self._ctx.probe_activation_times(getattr(self, "_n_probes", 0))
