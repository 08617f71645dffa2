"""Host-side pieces of the drop-in that need no GPU: monitors (reference tests/test_telemetry.py), define_stimulus unit
rules (tests/test_stimulation.py:111-304), conductivities (src/beat/conductivities.py:63-118)."""
import json
import logging
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))
from beat_b200 import conductivities, fem, stimulation, telemetry  # noqa: E402


class _Ksp:
    def __init__(self, its, rn, reason):
        self._v = (its, rn, reason)

    def getIterationNumber(self):
        return self._v[0]

    def getResidualNorm(self):
        return self._v[1]

    def getConvergedReason(self):
        return self._v[2]


def test_null_monitor():  # tests/test_telemetry.py:11-23
    m = telemetry.NullMonitor()
    with m.track_time("anything"):
        pass
    m.record_ksp(_Ksp(3, 1e-6, 2))
    m.advance_step(0.0, 0.1)


def test_performance_monitor_tracking_and_ksp(tmp_path, caplog):  # tests/test_telemetry.py:26-117
    m = telemetry.PerformanceMonitor(log_frequency=1)
    with m.track_time("pde_step"):
        pass
    with m.track_time("pde_step"):
        pass
    assert m.timings["pde_step"] >= 0.0
    m.record_ksp(_Ksp(5, 1e-7, 2))
    m.record_ksp(_Ksp(9, 1e-8, -3))
    assert m.ksp_total_iterations == 14 and m.ksp_max_iterations == 9 and m.ksp_last_iterations == 9
    with caplog.at_level(logging.INFO):
        m.advance_step(0.0, 0.1)
    assert m.step_counter == 1
    out = tmp_path / "summary.json"
    m.save_summary(out)
    data = json.loads(out.read_text())
    assert data["ksp"]["total_iterations"] == 14 if "ksp" in data else "ksp_total_iterations" in json.dumps(data)


def test_define_stimulus_units_and_window():  # tests/test_stimulation.py:253-304 and the unit rules :111-250
    mesh = fem.create_unit_square(fem.COMM_SELF, 2, 2)
    cells = fem.locate_entities(mesh, 2, lambda x: np.full(x.shape[1], True))
    tags = fem.meshtags(mesh, 2, cells, np.full(len(cells), 1, dtype=np.int32))
    time = fem.Constant(mesh, 0.0)
    start, duration, amplitude, chi = 1.0, 2.0, 3.0, 2.0
    stim = stimulation.define_stimulus(mesh=mesh, chi=chi, time=time, amplitude=amplitude, start=start, duration=duration,
                                       mesh_unit="cm", marker=1, subdomain_data=tags)
    assert stim.marker == 1
    load = fem.load_vector(mesh, stim.dZ, stim.marker)  # int phi_i over the marked cells: sums to the area (1)
    assert np.isclose(load.sum(), 1.0)

    def total(t):
        time.value = t
        e = stim.expr
        on = e.start <= float(time.value) <= e.end
        return load.sum() * (e.amplitude_now() if on else 0.0)

    assert np.isclose(total(0.0), 0.0)
    assert np.isclose(total(start), amplitude / chi)
    assert np.isclose(total(start + duration / 2), amplitude / chi)
    assert np.isclose(total(start + duration + 1e-6), 0.0)
    # mm mesh: uA/cm^2 -> uA/mm^2 for a cell stimulus in 2-D (effective dimension 3): factor 1e-2 (niederer_benchmark.py)
    from beat_b200.units import PerLength

    chi_cm = conductivities.default_conductivities("Niederer")["chi"]
    assert isinstance(chi_cm, PerLength) and chi_cm.unit == "cm" and float(chi_cm) == 1400.0 and float(chi_cm.to("mm")) == 140.0
    mm = stimulation.define_stimulus(mesh=mesh, chi=chi_cm, time=time, amplitude=50000.0, mesh_unit="mm", marker=1, subdomain_data=tags)
    assert np.isclose(mm.expr.amplitude_now(), 50000.0 / 1400.0 * 1e-2)
    # a plain number is 1/mesh_unit (src/beat/stimulation.py:186-207): (A uA/cm^3) / (chi /mm) -> uA/mm^2 = A/chi * 0.1^3
    mmf = stimulation.define_stimulus(mesh=mesh, chi=140.0, time=time, amplitude=50000.0, mesh_unit="mm", marker=1, subdomain_data=tags)
    assert np.isclose(mmf.expr.amplitude_now(), 50000.0 / 140.0 * 1e-3) and np.isclose(mmf.expr.amplitude_now(), mm.expr.amplitude_now())
    mf = stimulation.define_stimulus(mesh=mesh, chi=1.4e5, time=time, amplitude=50000.0, mesh_unit="m", marker=1, subdomain_data=tags)
    assert np.isclose(mf.expr.amplitude_now(), 50000.0 / 1.4e5 * 1e6)  # uA/m^2
    # facet (ds) stimulus in 2-D: effective dimension 2, uA/cm^2 / chi -> uA/mesh_unit
    facets = fem.locate_entities_boundary(mesh, 1, lambda x: np.isclose(x[0], 0.0))
    ftags = fem.meshtags(mesh, 1, facets, np.full(len(facets), 2, dtype=np.int32))
    fs = stimulation.define_stimulus(mesh=mesh, chi=chi_cm, time=time, amplitude=2000.0, mesh_unit="mm", marker=2, subdomain_data=ftags)
    assert fs.dZ.kind == "ds" and np.isclose(fs.expr.amplitude_now(), 2000.0 / 1400.0 * 0.1)
    with pytest.raises(ValueError):
        stimulation.define_stimulus(mesh=mesh, chi=1.0, time=time, mesh_unit="inch", marker=1, subdomain_data=tags)
    stim.assign(7.0)  # Stimulus.assign, stimulation.py:23-24
    assert stim.expr.amplitude_now() == 7.0


def test_generate_random_activation():  # tests/test_stimulation.py:305-380 of the reference, evaluated at the cell centroids (DG0)
    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.ones(3)], [4, 4, 4])
    t = fem.Constant(mesh, 0.0)
    points = np.array([[0.5, 0.5, 0.5], [1.0, 1.0, 1.0]])
    expr = stimulation.generate_random_activation(mesh=mesh, time=t, points=points, delays=np.array([1.0, 3.0]), stim_start=0.0,
                                                  stim_duration=1.0, stim_amplitude=5.0, tol=0.2)
    cent = mesh.geometry.x[mesh.cells].mean(axis=1).T

    def at(tv):
        t.value = tv
        return expr.evaluate(cent)

    assert np.allclose(at(0.5), 0.0)
    assert at(1.5).max() == pytest.approx(5.0) and at(1.5).min() == pytest.approx(0.0)
    assert np.allclose(at(2.5), 0.0)
    assert at(3.5).max() == pytest.approx(5.0)
    assert np.allclose(at(4.5), 0.0)
    first = expr.terms[0]
    assert (first.start, first.end) == (1.0, 2.0)
    # its load vector: int 1[x near p] phi_i dx with the centroid rule = (volume of the marked cells) / 4 per vertex
    load = fem.load_vector(mesh, fem.dx(domain=mesh), None, first.g, first.degree)
    marked = first.g(cent) > 0
    assert marked.sum() > 0 and np.isclose(load.sum(), marked.sum() * (1.0 / 64) / 6)
    with pytest.raises(AssertionError):
        stimulation.generate_random_activation(mesh=mesh, time=t, points=np.zeros((2, 3)), delays=np.zeros(3))
    assert len(stimulation.generate_random_activation(mesh=mesh, time=t, points=np.zeros((0, 3)), delays=np.zeros(0)).terms) == 0


def test_conductivities():  # src/beat/conductivities.py:29-118
    c = conductivities.default_conductivities("Niederer")
    s_l, s_t = conductivities.get_harmonic_mean_conductivity(**c)
    assert np.isclose(s_l, 0.17 * 0.62 / (0.17 + 0.62) / 1.4e5 * 1e3)
    assert np.isclose(s_t, 0.019 * 0.24 / (0.019 + 0.24) / 1.4e5 * 1e3)
    M = conductivities.define_conductivity_tensor(f0=np.array([1.0, 0.0, 0.0]), **c)
    assert np.allclose(M, np.diag([s_l, s_t, s_t]))
    f = np.array([1.0, 1.0, 0.0]) / np.sqrt(2)
    M2 = conductivities.conductivity_tensor(s_l, s_t, f)
    assert np.allclose(M2 @ f, s_l * f) and np.allclose(M2, M2.T)
    with pytest.raises(ValueError):
        conductivities.default_conductivities("nobody")


def test_single_cell_cache_and_type_check(tmp_path):
    """get_steady_state (reference src/beat/single_cell.py:86-156): Python callables are refused (no CPU fallback), the
    cache key depends on every input, and a cached result is returned without touching the device."""
    from beat_b200 import single_cell
    from beat_b200.models import tp06

    y0, p = tp06.init_state_values(), tp06.init_parameter_values()
    with pytest.raises(TypeError, match="device model handle"):
        single_cell.get_steady_state(lambda **kw: kw["states"], y0, p, tmp_path)
    fun = tp06.generalized_rush_larsen
    key = single_cell.compute_hash(fun, y0, p, 200, 1000, 0.05)
    assert key != single_cell.compute_hash(fun, y0, tp06.init_parameter_values(g_Ks=0.1), 200, 1000, 0.05)
    assert key != single_cell.compute_hash(tp06.forward_explicit_euler, y0, p, 200, 1000, 0.05)
    assert key != single_cell.compute_hash(fun, y0, p, 100, 1000, 0.05)
    np.save(tmp_path / f"steady_states_{key}.npy", np.arange(len(y0), dtype=float))
    assert np.array_equal(single_cell.get_steady_state(fun, y0, p, tmp_path), np.arange(len(y0)))


def test_performance_monitor_log_line_summary_and_device_labels(tmp_path, caplog):
    """The tokens the reference's tests look for (tests/test_telemetry.py:72-146) plus the device ledger: a label timed
    with CUDA events replaces the host interval of the same name."""
    m = telemetry.PerformanceMonitor(log_frequency=2)
    m.timings["step_time"] = 0.5
    with caplog.at_level(logging.INFO):
        m.advance_step(0.0, 0.1)
        assert len(caplog.records) == 0
        m.advance_step(0.1, 0.2)
    assert len(caplog.records) == 1
    msg = caplog.records[0].message
    assert "PDE step timing step=2" in msg and "step_time=" in msg and "ksp_iterations=0" in msg
    caplog.clear()
    m.timings.update({"slow_op": 10.0, "fast_op": 1.0})
    with m.track_time("pde_step"):
        pass
    m.record_device_stages({"pde_step": 0.25})
    m.record_device_stages({"pde_step": 0.25})
    assert m.timings["pde_step"] == 0.5  # host interval dropped, device totals accumulate
    with m.track_time("pde_step"):
        pass
    assert m.timings["pde_step"] == 0.5
    with caplog.at_level(logging.INFO):
        m.display_summary()
    text = caplog.records[0].message
    assert "PERFORMANCE SUMMARY" in text and "Total Steps:           2" in text and "pde_step [device]" in text
    assert text.find("slow_op") < text.find("fast_op")
    out = tmp_path / "sub" / "summary.json"
    m.save_summary(out)
    data = json.loads(out.read_text())
    assert data["total_steps"] == 2 and data["timings"]["slow_op"] == 10.0 and data["device_timed"] == ["pde_step"]
    other = telemetry.PerformanceMonitor(comm=fem.Comm(1, 2))
    other.save_summary(tmp_path / "rank1.json")
    assert not (tmp_path / "rank1.json").exists()


@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_splitting_solver_protocol_sequence(theta):
    """A non-device ODE backend goes through the reference's hand-off sequence (monodomain_solver.py:33-37,66-113):
    same calls, same order, same time arguments, one monitor label each."""
    from beat_b200.monodomain_solver import MonodomainSplittingSolver

    calls = []

    class Ode:
        def step(self, t0, dt):
            calls.append(("ode.step", round(t0, 12), round(dt, 12)))

        def to_dolfin(self):
            calls.append("to_dolfin")

        def from_dolfin(self):
            calls.append("from_dolfin")

        def ode_to_pde(self):
            calls.append("ode_to_pde")

        def pde_to_ode(self):
            calls.append("pde_to_ode")

    class Pde:
        _ctx = object()

        def assign_previous(self):
            calls.append("assign_previous")

        def step(self, interval):
            calls.append(("pde.step", interval))

    mon = telemetry.PerformanceMonitor(log_frequency=0)
    solver = MonodomainSplittingSolver(pde=Pde(), ode=Ode(), theta=theta, monitor=mon)
    assert calls == ["to_dolfin", "ode_to_pde", "assign_previous"]
    calls.clear()
    solver.solve((0.0, 0.2), 0.1)
    half = [("ode.step", 0.0, round(theta * 0.1, 12)), "to_dolfin", "ode_to_pde", "assign_previous", ("pde.step", (0.0, 0.1)), "pde_to_ode",
            "from_dolfin"]
    tail = ["assign_previous"] if theta == 1.0 else [("ode.step", 0.05, 0.05), "to_dolfin", "ode_to_pde", "assign_previous"]
    assert calls[: len(half) + len(tail)] == half + tail
    assert len(calls) == 2 * (len(half) + len(tail)) and mon.step_counter == 2
    labels = set(mon.timings)
    assert {"total_step", "ode_step", "pde_step", "ode_from_dolfin"} <= labels
    assert ("pde_assign_previous_after" in labels) == (theta == 1.0) and ("corrective_ode_step" in labels) == (theta != 1.0)


def test_petsc_options_to_device_solver_settings():
    """MonodomainModel._solver_settings: how the reference's petsc_options (base_model.py:136-168, demos) select the device
    solver - checked on a stub, no GPU: preonly/lu = tight CG, unknown preconditioners (hypre, ...) = Jacobi, 'auto' picks
    the driver by rows per rank and the polynomial preconditioner only for multi-rank resident meshes."""
    from types import SimpleNamespace

    from beat_b200.monodomain_model import NORM, PC, MonodomainModel

    def settings(opts, n_global=58_176, ranks=1, **params):
        stub = SimpleNamespace(parameters={"petsc_options": opts, **params},
                               _ctx=SimpleNamespace(device_info=lambda: {"n_sm": 148}),
                               _mesh=SimpleNamespace(index_map=SimpleNamespace(size_global=n_global), comm=SimpleNamespace(size=ranks)))
        out = MonodomainModel._solver_settings(stub)
        return out, stub

    (rtol, atol, max_it, pc, norm, x0), s = settings({"ksp_type": "preonly", "pc_type": "lu", "pc_factor_mat_solver_type": "mumps"})
    assert (rtol, pc, norm, x0) == (1e-12, PC["jacobi"], NORM["default"], 0) and s.ksp_type_used == "preonly" and s._ksp_type == 0
    (rtol, atol, max_it, pc, norm, x0), s = settings({"ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg"})
    assert (rtol, atol, max_it, pc) == (1e-5, 1e-50, 10000, PC["jacobi"]) and s.pc_type_used == "jacobi"
    (rtol, _, max_it, pc, norm, x0), s = settings({"ksp_type": "pipecg", "pc_type": "none", "ksp_rtol": 1e-9, "ksp_max_it": 50,
                                                   "ksp_norm_type": "unpreconditioned", "ksp_initial_guess_nonzero": True})
    assert (rtol, max_it, pc, norm, x0) == (1e-9, 50, PC["none"], NORM["unpreconditioned"], 1) and s._ksp_type == 1
    assert settings({"ksp_type": "cg"}, initial_guess_previous=True)[0][5] == 1
    # auto: one row per thread on 147 worker CTAs -> pipelined driver; beyond -> KSPCG
    assert settings({"ksp_type": "auto"}, n_global=147 * 512)[1].ksp_type_used == "pipecg"
    assert settings({"ksp_type": "auto"}, n_global=147 * 512 + 1)[1].ksp_type_used == "cg"
    assert settings({"ksp_type": "auto"}, n_global=8 * 58_176, ranks=8)[1].pc_type_used == "chebyshev"       # multi-rank, resident
    assert settings({"ksp_type": "auto", "pc_type": "jacobi"}, n_global=8 * 58_176, ranks=8)[1].pc_type_used == "jacobi"
    assert settings({"ksp_type": "auto"}, n_global=58_176, ranks=1)[1].pc_type_used == "jacobi"
    assert settings({"ksp_type": "auto"}, n_global=27_000_000, ranks=8)[1].pc_type_used == "jacobi"          # streaming: KSPCG
    (_, _, _, pc, _, _), s = settings({"ksp_type": "pipecg", "pc_type": "chebyshev", "pc_chebyshev_steps": 2, "pc_chebyshev_kappa": 6})
    assert pc == PC["chebyshev"] and s._cheb == (2, 6.0)
    with pytest.raises(NotImplementedError, match="pipelined"):
        settings({"ksp_type": "cg", "pc_type": "chebyshev"})
    with pytest.raises(NotImplementedError, match="gmres"):
        settings({"ksp_type": "gmres"})


def test_vector_mirror_coherence_protocol():
    """fem.Vector: the host mirror of a device vector (what dolfinx's Function.x.array is to the reference).  A fake device
    counts transfers: lazy download, one upload per modification, read-only views do not dirty, an asynchronous upload is
    waited for before the buffer is handed out again, and twins (v_ode / v / v_ after a fused step) share one download."""
    log = []
    dev = {"a": np.zeros(4), "b": np.zeros(4)}

    def binder(key):
        def download(out):
            log.append(("down", key))
            out[:] = dev[key]

        def upload(src):
            log.append(("up", key))
            dev[key] = np.array(src)

        return download, upload

    a, b = fem.Vector(4), fem.Vector(4)
    a.bind(*binder("a"), push_now=False, sync=lambda: log.append(("sync", "a")))
    b.bind(*binder("b"), push_now=False, sync=lambda: log.append(("sync", "b")))
    assert a.array_ro.sum() == 0 and log == []          # nothing newer on the device, nothing to fetch
    a.flush_to_device()
    assert log == []                                     # a read-only view did not dirty the mirror
    a.array[:] = [1, 2, 3, 4]
    a.flush_to_device()
    a.flush_to_device()
    assert log == [("up", "a")] and dev["a"].tolist() == [1, 2, 3, 4]
    _ = a.array_ro                                       # the upload may still be in flight: wait before exposing the buffer
    assert log[-1] == ("sync", "a")
    _ = a.array_ro
    assert log.count(("sync", "a")) == 1
    with pytest.raises(ValueError):
        a.array_ro[0] = 7.0
    # the device changes both vectors to the same content (fused step): b is a's twin
    dev["a"] = dev["b"] = np.array([9.0, 8.0, 7.0, 6.0])
    a.mark_device_newer()
    b.mark_device_newer(twin_of=a)
    log.clear()
    assert a.array_ro.tolist() == [9, 8, 7, 6] and log == [("down", "a")]
    assert b.array_ro.tolist() == [9, 8, 7, 6] and log == [("down", "a")]   # copied on the host, no second transfer
    # ... but not when the twin moved on in between
    dev["a"], dev["b"] = np.full(4, 1.0), np.full(4, 2.0)
    a.mark_device_newer()
    b.mark_device_newer(twin_of=a)
    a.mark_device_newer()                                # a changed again: b must fetch its own content
    log.clear()
    assert b.array_ro.tolist() == [2, 2, 2, 2] and log == [("down", "b")]


def test_vector_write_only_access_skips_the_download():
    """fem.Vector.array_wo: for a caller that overwrites the whole array (bench.py's e2e loop: the host owns V between
    steps) - no device-to-host transfer first, the content is uploaded by the next device operation, and an upload still
    in flight is waited for before the buffer is handed out."""
    log = []
    dev = {"a": np.full(4, 5.0)}
    a = fem.Vector(4)
    a.bind(lambda out: (log.append("down"), out.__setitem__(slice(None), dev["a"])), lambda src: (log.append("up"), dev.__setitem__("a", np.array(src))),
           push_now=False, sync=lambda: log.append("sync"))
    a.mark_device_newer()
    w = a.array_wo
    assert log == [] and not a.device_newer and a.host_dirty     # nothing fetched; the host copy is now the truth
    w[:] = [1.0, 2.0, 3.0, 4.0]
    a.flush_to_device()
    assert log == ["up"] and dev["a"].tolist() == [1, 2, 3, 4]
    _ = a.array_wo                                                # asynchronous upload: wait before exposing the buffer again
    assert log == ["up", "sync"]
    a.mark_device_newer()
    assert a.array.tolist() == [1, 2, 3, 4] and log[-1] == "down"  # the ordinary accessor still fetches


def test_bench_workload_description_is_the_same_in_both_arms():
    """bench.py: the `config` object depends only on (workload, scaling, world) - the driver compares the two arms' configs -
    and the CPU arm's sample is a bounded x-block of the same slab."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c1 = bench.workload_config("niederer_dx0.025", "strong", 8)
    assert c1 == bench.workload_config("niederer_dx0.025", "strong", 8) and c1["nodes"] == 801 * 281 * 121 and "model" not in c1
    assert bench.workload_config("niederer_dx0.2", "weak", 2)["nodes"] == 201 * 36 * 16
    assert bench.slab_nodes(0.2, 20.0) == 58176
    for dx in (0.5, 0.2, 0.1, 0.05, 0.025, 0.016):
        L = bench.cpu_sample_length(dx)
        assert 0 < L <= 20.0 and bench.slab_nodes(dx, L) <= 2.0e6
        assert abs(L / dx - round(L / dx)) < 1e-9                 # a whole number of elements
    assert bench.cpu_sample_length(0.2) == 20.0                   # small slabs are timed whole
    assert bench.cpu_sample_length(0.025) >= 1.25 - 1e-9          # ... large ones on a block of >= 16 elements incl. the stimulus corner


def test_mixed_model_multi_ode_solver_protocol_with_a_stand_in_device_stage():
    """DolfinMultiODESolver with DIFFERENT cell models per region (src/beat/odesolver.py:228-354 allows any callable per marker)
    dispatches to MixedModelODESolver: one device ODE stage per region, membrane potential exchanged through the host mirrors.
    The composite's bookkeeping (region index sets, broadcast of initial states, the five hand-offs, the reference's accessors)
    is checked here with a NumPy stand-in for the device stage (the oracle's models); on a GPU the stage is ODESystemSolver."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _problems as P
    from beat_b200 import odesolver
    from beat_b200.models import fhn, tp06

    oracle = {"fhn": P.oracle_model("fhn"), "tp06": P.oracle_model("tp06")}

    class StandIn:  # what MixedModelODESolver needs from ODESystemSolver
        def __init__(self, fun, states, parameters, monitor=None, v_index=0):
            self.fun, self.states, self.parameters = fun, np.array(states), parameters
            self.step_fn = getattr(oracle[fun.model_tag], "generalized_rush_larsen" if fun.scheme_id == 1 else "forward_explicit_euler")

        def step(self, t0, dt):
            self.states[:] = self.step_fn(self.states, t0, dt, np.asarray(self.parameters))

        def state_row(self, i):
            return self.states[i].copy()

        def set_state_row(self, i, v):
            self.states[i] = v

    mesh = fem.create_unit_square(fem.COMM_SELF, 5, 4)
    V = fem.functionspace(mesh, ("P", 1))
    markers = fem.Function(V)
    markers.x.array[:] = np.where(mesh.geometry.x[:, 0] < 0.45, 7, 3)
    marr = np.asarray(markers.x.array_ro)
    v_ode, v_pde = fem.Function(V), fem.Function(V)
    y = {7: fhn.init_state_values(), 3: tp06.init_state_values(V=-80.0)}
    p = {7: fhn.init_parameter_values(), 3: tp06.init_parameter_values()}
    fun = {7: fhn.forward_explicit_euler, 3: tp06.generalized_rush_larsen}
    vi = {7: fhn.state_index("v"), 3: tp06.state_index("V")}
    ode = odesolver.DolfinMultiODESolver(v_ode=v_ode, v_pde=v_pde, markers=markers, init_states=y, parameters=p, fun=fun,
                                         num_states={7: 2, 3: 19}, v_index=vi, system_solver=StandIn)
    assert isinstance(ode, odesolver.MixedModelODESolver)
    assert ode.num_points(7) == int((marr == 7).sum()) and ode.shape(3) == (19, int((marr == 3).sum())) and ode.num_parameters(3) == len(p[3])
    assert np.allclose(ode.values(3), np.asarray(y[3])[:, None]) and np.allclose(ode.values(7), np.asarray(y[7])[:, None])
    want = {k: np.repeat(np.asarray(y[k])[:, None], ode.num_points(k), axis=1) for k in (7, 3)}
    t = 0.0
    for _ in range(3):
        ode.step(t, 0.02)
        want[7] = oracle["fhn"].forward_explicit_euler(want[7], t, 0.02, p[7])
        want[3] = oracle["tp06"].generalized_rush_larsen(want[3], t, 0.02, p[3])
        t += 0.02
    for k in (7, 3):
        assert np.array_equal(ode.values(k), want[k])
    ode.to_dolfin()
    assert np.array_equal(v_ode.x.array_ro[marr == 7], want[7][vi[7]]) and np.array_equal(v_ode.x.array_ro[marr == 3], want[3][vi[3]])
    ode.ode_to_pde()
    assert np.array_equal(v_pde.x.array_ro, v_ode.x.array_ro)
    ramp = np.arange(marr.size, dtype=float) + 0.25
    v_pde.x.array[:] = ramp
    ode.pde_to_ode()
    ode.from_dolfin()
    assert np.array_equal(ode.values(3)[vi[3]], ramp[marr == 3])
    assert np.array_equal(ode.values(7)[vi[7]], ramp[marr == 7])
    assert np.array_equal(ode.values(3)[0], want[3][0])  # the other states are untouched
    with pytest.raises(RuntimeError, match="equal size"):
        ode.full_values
    # the splitting solver drives it through the protocol sequence (not the fused device step)
    calls = []

    class Pde:
        state = v_pde
        monitor = telemetry.NullMonitor()
        parameters = {"theta": 0.5}

        def assign_previous(self):
            calls.append("assign_previous")

        def step(self, interval):
            calls.append("pde_step")

    from beat_b200.monodomain_solver import MonodomainSplittingSolver

    solver = MonodomainSplittingSolver.__new__(MonodomainSplittingSolver)
    solver.pde, solver.ode, solver.theta, solver.monitor = Pde(), ode, 1.0, telemetry.NullMonitor()
    solver.__post_init__()
    assert solver._fused is False
    solver.step((0.0, 0.01))
    assert calls == ["assign_previous", "assign_previous", "pde_step", "assign_previous"]
    # same models in every region still take the single-launch path (refused here only because there is no GPU context)
    same = {7: tp06.generalized_rush_larsen, 3: tp06.generalized_rush_larsen}
    assert odesolver.DolfinMultiODESolver.__new__(odesolver.DolfinMultiODESolver, v_ode=v_ode, v_pde=v_pde, markers=markers,
                                                  init_states=y, parameters=p, fun=same, num_states={7: 19, 3: 19},
                                                  v_index={7: 17, 3: 17}).__class__ is odesolver.DolfinMultiODESolver
