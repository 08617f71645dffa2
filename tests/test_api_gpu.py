"""The reference's own known-answer tests for this path, run through the drop-in API on the device
(reference files cited per test).  pytest -m gpu."""
import numpy as np
import pytest

import _problems as P

pytestmark = pytest.mark.gpu


def _beat():
    import beat_b200 as beat

    return beat


def test_dolfin_ode_solver_data_movement():
    """tests/test_odesolver.py:52-117 of the reference (shapes, step result, to_dolfin / ode_to_pde /
    pde_to_ode / from_dolfin semantics) with FitzHugh-Nagumo forward Euler as the cell model."""
    beat = _beat()
    fem = beat.fem
    om = P.oracle_model("fhn")
    mesh = fem.create_unit_square(fem.COMM_SELF, 5, 5)
    time = fem.Constant(mesh, 0.0)
    pde = beat.MonodomainModel(time=time, mesh=mesh, M=1.0)
    v_pde = pde.state
    v_ode = fem.Function(fem.functionspace(mesh, ("P", 1)))
    imap = v_ode.function_space.dofmap.index_map
    n_ode = imap.size_local + imap.num_ghosts
    s0, v0 = 2.0, 1.0
    params = beat.models.fhn.init_parameter_values()
    ode = beat.odesolver.DolfinODESolver(v_ode=v_ode, v_pde=v_pde, init_states=np.array([s0, v0]), parameters=params,
                                         fun=beat.models.fhn.forward_explicit_euler, num_states=2, v_index=1)
    assert ode.full_values.shape == (2, n_ode) and ode.values.shape == (2, n_ode)
    assert np.allclose(ode.values[0], s0) and np.allclose(ode.values[1], v0)
    dt = 0.1
    want = om.forward_explicit_euler(np.array([[s0], [v0]]), 0.0, dt, params)
    ode.step(0.0, dt)
    assert np.allclose(ode.values[0], want[0], rtol=1e-14) and np.allclose(ode.values[1], want[1], rtol=1e-14)
    assert np.allclose(v_ode.x.array, 0.0)  # the dolfin function is not updated by step
    ode.to_dolfin()
    assert np.allclose(v_ode.x.array, want[1])
    assert np.allclose(v_pde.x.array, 0.0)
    ode.ode_to_pde()
    assert np.allclose(v_pde.x.array, want[1])
    v_pde.x.array[:] = 1.0
    ode.pde_to_ode()
    assert np.allclose(v_ode.x.array, 1.0)
    ode.from_dolfin()
    assert np.allclose(ode.values[1], 1.0)
    assert np.allclose(ode.values[0], want[0])
    states = ode.states_to_dolfin()
    assert len(states) == 2 and np.allclose(states[1].x.array, 1.0) and np.allclose(states[0].x.array, want[0])


def test_rejects_python_callables():
    beat = _beat()
    fem = beat.fem
    mesh = fem.create_unit_interval(fem.COMM_SELF, 4)
    pde = beat.MonodomainModel(time=fem.Constant(mesh, 0.0), mesh=mesh, M=1.0)
    with pytest.raises(TypeError):
        beat.odesolver.DolfinODESolver(v_ode=fem.Function(pde.V), v_pde=pde.state, init_states=np.zeros(2), parameters=np.zeros(2),
                                       fun=lambda **kw: None, num_states=2)
    with pytest.raises(RuntimeError):
        beat.models.tp06.generalized_rush_larsen(states=None, t=0, parameters=None, dt=0.1)


def test_single_stimulation():
    """tests/test_stimulation.py:12-46 of the reference: M = 0 on the unit interval, exact integrals."""
    beat = _beat()
    fem = beat.fem
    mesh = fem.create_unit_interval(fem.COMM_SELF, 10)
    value, end, start, dt = 2.0, 1.0, 0.5, 0.01
    time = fem.Constant(mesh, 0.0)
    expr = fem.conditional(fem.And(fem.ge(time, start), fem.le(time, end)), value, 0.0)
    I_s = beat.stimulation.Stimulus(dZ=fem.dx(domain=mesh), expr=expr)
    pde = beat.MonodomainModel(time=time, mesh=mesh, M=fem.Constant(mesh, 0.0), I_s=I_s)
    pde.step((0.0, 0.4))
    assert np.allclose(pde.state.x.array, 0.0)
    t0 = 0.9
    pde.solve((0.4, t0), dt=dt)
    assert np.allclose(pde.state.x.array, value * (t0 - start))
    pde.solve((t0, end + dt), dt=dt)
    assert np.allclose(pde.state.x.array, (end - start - dt) * value)
    pde.solve((end + dt, 2 * end), dt=dt)
    assert np.allclose(pde.state.x.array, (end - start - dt) * value)


def test_double_stimulation():
    """tests/test_stimulation.py:49-107 of the reference."""
    beat = _beat()
    fem = beat.fem
    mesh = fem.create_unit_interval(fem.COMM_SELF, 10)
    dt, value1, value2, start1, end1, start2, end2 = 0.01, 2.0, 3.0, 0.5, 1.0, 0.9, 1.5
    time = fem.Constant(mesh, 0.0)
    e1 = fem.conditional(fem.And(fem.ge(time, start1), fem.le(time, end1)), value1, 0.0)
    e2 = fem.conditional(fem.And(fem.ge(time, start2), fem.le(time, end2)), value2, 0.0)
    dx = fem.dx(domain=mesh)
    pde = beat.MonodomainModel(time=time, mesh=mesh, M=fem.Constant(mesh, 0.0),
                               I_s=[beat.stimulation.Stimulus(dZ=dx, expr=e1), beat.stimulation.Stimulus(dZ=dx, expr=e2)])
    pde.step((0.0, 0.4))
    assert np.allclose(pde.state.x.array, 0.0)
    t0 = 0.9
    pde.solve((0.4, t0), dt=dt)
    assert np.allclose(pde.state.x.array, value1 * (t0 - start1))
    pde.solve((t0, end1 + dt), dt=dt)
    assert np.allclose(pde.state.x.array, (end1 - start1 - dt) * value1 + (end1 + dt - start2) * value2)
    pde.solve((end1 + dt, end2 + dt), dt=dt)
    want = (end1 - start1 - dt) * value1 + (end2 - start2 - dt) * value2
    assert np.allclose(pde.state.x.array, want)
    pde.solve((end2 + dt, 2 * end2), dt=dt)
    assert np.allclose(pde.state.x.array, want)


@pytest.mark.parametrize("M,err", [(0.0, 1e-4), (1.0, 2e-4), (2.0, 2e-4)])
def test_monodomain_analytic(M, err):
    """tests/test_monodomain.py:38-64 of the reference: PDE-only MMS, N=15, theta=0.5, dt=1e-3, 10 steps."""
    from oracle import fem as ofem

    beat = _beat()
    fem = beat.fem
    N, dt = 15, 0.001
    T = 10 * dt
    mesh = fem.create_unit_square(fem.COMM_SELF, N, N)
    time = fem.Constant(mesh, 0.0)
    g = lambda x: np.cos(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1])  # noqa: E731
    h = lambda t: np.cos(t) + M * 8 * np.pi**2 * np.sin(t)  # noqa: E731
    model = beat.MonodomainModel(time=time, mesh=mesh, M=M, I_s=fem.Separable(time, g, h, degree=8), params=dict(theta=0.5))
    res = model.solve((0, T), dt=dt)
    pts, cells = ofem.rectangle_mesh(N, N)
    assert np.allclose(pts, mesh.geometry.x[:, :2])
    e = ofem.l2_error(pts, cells, res.state.x.array, lambda x: g(x) * np.sin(T))
    assert e < err


def test_niederer_api_matches_oracle_and_probes():
    """The packaged Niederer set-up (dx = 0.5) against the oracle's restatement, through the public API,
    including the device-side probes / activation tracker."""
    from beat_b200 import niederer
    from oracle import monodomain as om_mono

    solver, info = niederer.setup(dx=0.5, rtol=1e-12)
    prob = P.niederer_slab(0.5)
    om = P.oracle_model("tp06")
    params = om.init_parameter_values(stim_amplitude=0.0)
    y0 = om.init_state_values()
    stim = om_mono.Stimulus.window(prob["stim_load"], 0.0, 2.0, prob["stim_amp"])
    pde = om_mono.MonodomainModel(prob["mass"], prob["stiff"], [stim], C_m=prob["C_m"], theta=0.5, solver="lu")
    ode = om_mono.ODESolver(v_pde=pde.state, init_states=y0, parameters=params, fun=om.generalized_rush_larsen, num_states=19,
                            v_index=om.state_index("V"))
    ref = om_mono.SplittingSolver(pde, ode)
    dt, t = 0.05, 0.0
    act = {k: -1.0 for k in niederer.POINTS}
    for _ in range(60):
        solver.step((t, t + dt))
        ref.step((t, t + dt))
        if pde.state[0] > 0 and act["P1"] < 0:
            act["P1"] = t
        t += dt
    v = solver.pde.state.x.array
    assert np.abs(v - pde.state).max() <= 1e-8 * np.abs(pde.state).max()
    assert np.abs(solver.ode.values - ode.values).max() <= 1e-7 * np.abs(ode.values).max()
    got = solver.pde.activation_times()
    assert got[info["probe_ids"]["P1"]] == pytest.approx(act["P1"], abs=1e-12)
    assert act["P1"] > 0
    assert np.allclose(solver.pde.probe_values()[info["probe_ids"]["P1"]], v[0])


def test_random_activation_through_the_model():
    """stimulation.generate_random_activation (src/beat/stimulation.py:279-363) as I_s of the PDE model with M = 0: the
    charge int v dx grows by dt * amplitude * |marked cells| for every step whose theta-point lies in a point's window."""
    beat = _beat()
    fem = beat.fem
    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.ones(3)], [4, 4, 4])
    time = fem.Constant(mesh, 0.0)
    points, delays = np.array([[0.5, 0.5, 0.5], [1.0, 1.0, 1.0]]), np.array([1.0, 3.0])
    I_s = beat.stimulation.generate_random_activation(mesh=mesh, time=time, points=points, delays=delays, stim_start=0.0,
                                                      stim_duration=1.0, stim_amplitude=5.0, tol=0.2)
    pde = beat.MonodomainModel(time=time, mesh=mesh, M=0.0, I_s=I_s, params={"theta": 0.5})
    cent = mesh.geometry.x[mesh.cells].mean(axis=1).T
    vol = [float((term.g(cent) > 0).sum()) / (64 * 6) for term in I_s.terms]
    ones = fem.load_vector(mesh, fem.dx(domain=mesh), None)  # Mass * 1
    dt, t, want = 0.1, 0.0, 0.0
    for _ in range(50):
        pde.step((t, t + dt))
        pde.assign_previous()
        tm = t + 0.5 * dt
        for term, vo in zip(I_s.terms, vol):
            if term.start <= tm <= term.end:
                want += dt * 5.0 * vo
        t += dt
        assert abs(float(ones @ pde.state.x.array_ro) - want) <= 1e-9 * max(want, 1.0)
    assert want > 0


def test_whole_field_observers_match_host_tracking():
    """the per-node activation map, min / max of v and a strided snapshot computed on the device equal what the
    demos compute from state.x.array on the host after every step (demos/niederer_benchmark.py:271-287)."""
    from beat_b200 import niederer

    solver, info = niederer.setup(dx=0.5, rtol=1e-10)
    pde = solver.pde
    pde.observe(activation_map=True, threshold=0.0, minmax=True)
    n = info["n_owned"]
    act = np.full(n, -1.0)
    dt, t = 0.05, 0.0
    for k in range(80):
        solver.step((t, t + dt))
        v = np.array(pde.state.x.array_ro)[:n]
        newly = (v > 0.0) & (act < 0.0)
        act[newly] = t
        if k % 20 == 19:
            lo, hi = pde.v_minmax()
            assert lo == v.min() and hi == v.max()
        t += dt
    got = pde.activation_map()
    assert (act >= 0).sum() > 10 and (act < 0).sum() > 10  # the front is somewhere inside the slab
    assert np.array_equal(got, act)
    assert np.array_equal(pde.state_snapshot(stride=7, offset=3), v[3::7])
    assert np.array_equal(pde.state_snapshot(stride=1), v)


def test_ode_system_solver():
    """src/beat/odesolver.py:46-79: the plain driver, states held by reference, shared and per-node parameters."""
    beat = _beat()
    om = P.oracle_model("tp06")
    rng = np.random.default_rng(11)
    n = 777
    states = P.perturbed_states(om, n, rng, "V")
    params = beat.models.tp06.init_parameter_values()
    ode = beat.odesolver.ODESystemSolver(fun=beat.models.tp06.generalized_rush_larsen, states=states.copy(), parameters=params)
    assert ode.num_points == n and ode.num_states == 19
    want = states.copy()
    t = 0.0
    for _ in range(3):
        ode.step(t, 0.02)
        want = om.generalized_rush_larsen(want, t, 0.02, params)
        t += 0.02
    assert np.abs(ode.states - want).max() <= 1e-11 * np.abs(want).max()
    params[beat.models.tp06.parameter_index("g_Na")] *= 0.5  # mutated in place by the user (pace_train.py:224)
    ode.step(t, 0.02)
    want = om.generalized_rush_larsen(want, t, 0.02, params)
    assert np.abs(ode.states - want).max() <= 1e-11 * np.abs(want).max()
    with pytest.raises(TypeError):
        beat.odesolver.ODESystemSolver(fun=lambda **kw: None, states=states, parameters=params)


def test_dolfin_multi_ode_solver():
    """src/beat/odesolver.py:228-354: two regions with different parameter sets and initial states."""
    beat = _beat()
    fem = beat.fem
    om = P.oracle_model("tp06")
    mesh = fem.create_unit_square(fem.COMM_SELF, 6, 6)
    pde = beat.MonodomainModel(time=fem.Constant(mesh, 0.0), mesh=mesh, M=1.0)
    V = fem.functionspace(mesh, ("P", 1))
    markers = fem.Function(V)
    x = mesh.geometry.x
    markers.x.array[:] = np.where(x[:, 0] < 0.5, 1, 2)
    m = beat.models.tp06
    p = {1: m.init_parameter_values(), 2: m.init_parameter_values(g_Ks=0.098, g_to=0.073)}
    y = {1: m.init_state_values(), 2: m.init_state_values(V=-80.0)}
    fun = {k: m.generalized_rush_larsen for k in (1, 2)}
    ode = beat.odesolver.DolfinMultiODESolver(v_ode=fem.Function(V), v_pde=pde.state, markers=markers, init_states=y, parameters=p,
                                              fun=fun, num_states={1: 19, 2: 19}, v_index={1: m.state_index("V"), 2: m.state_index("V")})
    marr = markers.x.array_ro
    assert ode.num_points(1) == int((marr == 1).sum()) and ode.shape(2) == (19, int((marr == 2).sum()))
    t = 0.0
    want = {k: np.repeat(y[k][:, None], ode.num_points(k), axis=1) for k in (1, 2)}
    for _ in range(4):
        ode.step(t, 0.05)
        for k in (1, 2):
            want[k] = om.generalized_rush_larsen(want[k], t, 0.05, p[k])
        t += 0.05
    for k in (1, 2):
        assert np.abs(ode.values(k) - want[k]).max() <= 1e-11 * np.abs(want[k]).max()
    ode.to_dolfin()
    assert np.allclose(ode.v_ode.x.array_ro[marr == 2], want[2][m.state_index("V")])
    assert ode.full_values.shape == (19, marr.size)
    with pytest.raises(RuntimeError):
        bad = fem.Function(fem.functionspace(fem.create_unit_square(fem.COMM_SELF, 3, 3), ("P", 1)))
        beat.odesolver.DolfinMultiODESolver(v_ode=fem.Function(V), v_pde=pde.state, markers=bad, init_states=y, parameters=p, fun=fun,
                                            num_states={1: 19, 2: 19}, v_index={1: 0, 2: 0})


def test_readme_unit_square_fhn():
    """BASELINE config 1 (README.md:35-205): unit square 32x32, FitzHugh-Nagumo forward Euler, M = 0.001, dt = 0.01,
    stimulus 600 on the lower-left quarter for t <= 0.5; 400 steps through the public API against the oracle."""
    from oracle import fem as ofem
    from oracle import monodomain as om_mono

    beat = _beat()
    fem = beat.fem
    om = P.oracle_model("fhn")
    N, dt, nsteps = 32, 0.01, 400
    mesh = fem.create_unit_square(fem.COMM_SELF, N, N)
    time = fem.Constant(mesh, 0.0)
    cells = fem.locate_entities(mesh, 2, lambda x: (x[0] <= 0.5 + 1e-12) & (x[1] <= 0.5 + 1e-12))
    tags = fem.meshtags(mesh, 2, cells, 1)
    I_s = beat.Stimulus(expr=fem.TimeWindow(time, 0.0, 0.5, 600.0), dZ=fem.Measure("dx", domain=mesh, subdomain_data=tags), marker=1)
    pde = beat.MonodomainModel(time=time, mesh=mesh, M=0.001, I_s=I_s, dx=I_s.dZ,
                               params={"petsc_options": {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12}})
    m = beat.models.fhn
    prm = m.init_parameter_values(stim_amplitude=0.0)
    ode = beat.odesolver.DolfinODESolver(v_ode=fem.Function(pde.V), v_pde=pde.state, fun=m.forward_explicit_euler,
                                         init_states=m.init_state_values(), parameters=prm, num_states=2, v_index=m.state_index("v"))
    solver = beat.MonodomainSplittingSolver(pde=pde, ode=ode)
    solver.solve((0.0, nsteps * dt), dt)

    pts, ocells = ofem.rectangle_mesh(N, N)
    mass, stiff = ofem.assemble_p1(pts, ocells, 0.001 * np.eye(2))
    sc = ofem.cells_all_vertices(pts, ocells, lambda x: (x[0] <= 0.5 + 1e-12) & (x[1] <= 0.5 + 1e-12))
    load = ofem.load_vector_cells(pts, ocells, sc)
    opde = om_mono.MonodomainModel(mass, stiff, [om_mono.Stimulus.window(load, 0.0, 0.5, 600.0)], C_m=1.0, theta=0.5, solver="lu")
    oode = om_mono.ODESolver(v_pde=opde.state, init_states=om.init_state_values(), parameters=prm, fun=om.forward_explicit_euler,
                             num_states=2, v_index=om.state_index("v"))
    ref = om_mono.SplittingSolver(opde, oode)
    ref.solve((0.0, nsteps * dt), dt)
    assert np.allclose(pts, mesh.geometry.x[:, :2])
    v = solver.pde.state.x.array_ro
    assert v.max() > opde.state.min() + 10.0  # the stimulus did something
    assert np.abs(v - opde.state).max() <= 1e-8 * np.abs(opde.state).max()
    assert np.abs(solver.ode.values - oode.values).max() <= 1e-8 * np.abs(oode.values).max()


def test_niederer_full_config_activation_times():
    """BASELINE config 2 as the demo runs it (dx = 0.2 mm, dt = 0.01 ms, PETSc-default rtol): activation times from the
    device-side probes agree with the oracle's run of the same algorithm within one dt (north_star) - and, like the
    oracle, with the published row (demos/niederer_benchmark.py:321) to 0.07 ms (0.2 %; the residual at small dt is the
    reference's solver tolerance, tests/test_oracle_niederer.py)."""
    from beat_b200 import niederer
    from oracle import niederer as onied

    dx, dt, T = 0.2, 0.01, 45.0
    solver, info = niederer.setup(dx=dx, ksp_type="cg")
    nsteps = int(round(T / dt))
    solver.solve_on_device(0.0, dt, nsteps)
    got = solver.pde.activation_times()
    want = onied.run(dx, dt, T=T, tp06=P.oracle_model("tp06"))
    pub = dict(zip(onied.POINTS, onied.PUBLISHED[(dx, dt)]))
    for name, pid in info["probe_ids"].items():
        assert want[name] >= 0 and got[pid] >= 0, (name, want[name], got[pid])
        assert abs(got[pid] - want[name]) <= dt + 1e-9, (name, got[pid], want[name])
        assert abs(got[pid] - pub[name]) <= 0.07, (name, got[pid], pub[name])  # ms; the oracle's own distance to the table at small dt


# ---- the reference's splitting tests with its own two-state linear cell model, on the device -------------------
def _g2(x):
    return np.cos(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1])


def _simple_split_solver(N, theta=1.0, ksp_rtol=1e-10):
    beat = _beat()
    fem = beat.fem
    mesh = fem.create_unit_square(fem.COMM_SELF, N, N)
    time = fem.Constant(mesh, 0.0)
    I_s = fem.Separable(time, _g2, lambda t: 8 * np.pi**2 * np.sin(t), degree=6)  # ac_func, test_monodomain_solver.py:21-22
    pde = beat.MonodomainModel(time=time, mesh=mesh, M=1.0, I_s=I_s,
                               params={"petsc_options": {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": ksp_rtol}})
    m = beat.models.simple
    init = np.zeros((2, mesh.geometry.x.shape[0]))
    init[1] = -_g2(mesh.geometry.x.T)  # s_exact at t = 0
    ode = beat.odesolver.DolfinODESolver(v_ode=fem.Function(pde.V), v_pde=pde.state, fun=m.forward_explicit_euler, init_states=init,
                                         parameters=m.init_parameter_values(), num_states=2, v_index=0)
    return beat.MonodomainSplittingSolver(pde=pde, ode=ode, theta=theta), mesh


def _l2_error_vs_exact(solver, mesh, N):
    from oracle import fem as ofem

    pts, cells = ofem.rectangle_mesh(N, N)
    assert np.allclose(pts, mesh.geometry.x[:, :2])
    t_eval = float(solver.pde.time.value)  # the reference evaluates v_exact at time.value of the last PDE step
    return ofem.l2_error(pts, cells, solver.pde.state.x.array_ro, lambda x: _g2(x) * np.sin(t_eval))


def test_monodomain_splitting_analytic():
    """tests/test_monodomain_solver.py:41-87 (P1 ODE space), the reference's bound: L2 error < 0.002."""
    solver, mesh = _simple_split_solver(50)
    solver.solve((0.0, 1.0), dt=0.01)
    assert _l2_error_vs_exact(solver, mesh, 50) < 0.002


def test_solve_on_device_with_host_evaluated_source_equals_solve():
    """A source whose amplitude the host evaluates each step (fem.Separable h(t)) must not be frozen at its first value by
    solve_on_device (it falls back to the per-step loop); and the monitors count every step (telemetry.py:86-92)."""
    beat = _beat()
    a, mesh = _simple_split_solver(12)
    b, _ = _simple_split_solver(12)

    class Counter(beat.NullMonitor):
        n = 0

        def advance_step(self, t0, t1):
            self.n += 1

    b.monitor = Counter()
    assert a.pde.has_host_evaluated_sources()
    a.solve((0.0, 0.3), dt=0.01)
    b.solve_on_device(0.0, 0.01, 30)
    assert b.monitor.n == 30
    assert np.array_equal(a.pde.state.x.array_ro, b.pde.state.x.array_ro)
    assert np.array_equal(a.ode.values, b.ode.values)
    assert float(a.pde.time.value) == float(b.pde.time.value)


def test_monodomain_splitting_spatial_convergence():
    """tests/test_monodomain_solver.py:98-149: dt = 1e-3, T = 1, N = 8, 16, 32 -> mean rate > 1.85."""
    errors = []
    for N in (8, 16, 32):
        solver, mesh = _simple_split_solver(N)
        solver.solve((0.0, 1.0), dt=0.001)
        errors.append(_l2_error_vs_exact(solver, mesh, N))
    rates = [np.log(e1 / e2) / np.log(2) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert sum(rates) / len(rates) > 1.85, (rates, errors)


def test_monodomain_splitting_temporal_convergence():
    """tests/test_monodomain_solver.py:161-216 (theta_split = 1): N = 150, dt = 1/8, 1/16, 1/32 -> mean rate > 1.0."""
    errors = []
    for dt in (1.0 / 8, 1.0 / 16, 1.0 / 32):
        solver, mesh = _simple_split_solver(150)
        solver.solve((0.0, 1.0), dt=dt)
        errors.append(_l2_error_vs_exact(solver, mesh, 150))
    rates = [np.log(e1 / e2) / np.log(2) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert sum(rates) / len(rates) > 1.0, (rates, errors)


def test_simple_ode_odesystemsolver_rate():
    """tests/test_odesolver.py:20-49: forward Euler on v' = -s, s' = v from (1, 0): samples at t = 0.1 .. 1.0 against
    (cos t, sin t), dt = 0.1, 0.01, 0.001 -> error decades per dt decade = 1 +- 0.01."""
    beat = _beat()
    m = beat.models.simple
    x = np.arange(0.1, 1.0 + 0.1, 0.1)
    sol = np.vstack((np.cos(x), np.sin(x))).T
    errors = []
    for dt in (0.1, 0.01, 0.001):
        ode = beat.odesolver.ODESystemSolver(fun=m.forward_explicit_euler, states=np.array([[1.0], [0.0]]), parameters=m.init_parameter_values())
        y = np.zeros((len(x), 2))
        j, t = 0, 0.0
        for _ in range(int(round(1.0 / dt))):
            ode.step(t, dt)
            t += dt
            if j < len(x) and np.isclose(t, x[j]):
                y[j, :] = ode.states[:, 0]
                j += 1
        assert j == len(x)
        errors.append(np.linalg.norm(sol - y))
    rates = [np.log(e1 / e2) / np.log(10) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert np.allclose(rates, 1, atol=0.01), rates


def test_lv_ellipsoid_matches_oracle():
    """BASELINE config 5 in small: synthetic LV shell (unstructured-style connectivity: periodic in phi, cell-wise
    conductivity tensor), ENDO surface stimulus, three transmural parameter sets (DolfinMultiODESolver) - through the
    public API against the oracle's assembly / LU solve / NumPy cell model."""
    from beat_b200 import lv_ellipsoid
    from oracle import fem as ofem
    from oracle import monodomain as om_mono

    solver, info = lv_ellipsoid.setup(n=(3, 12, 16), rtol=1e-12, ksp_type="cg")
    mesh, geo = info["mesh"], info["geo"]
    pts, cells = mesh.geometry.x, mesh.cells
    mass, stiff = ofem.assemble_p1(pts, cells, info["M"])
    endo = geo.ffun.facets[geo.ffun.find(geo.markers["ENDO"][0])]
    load = ofem.load_vector_facets(pts, endo)
    assert load.sum() > 0
    om = P.oracle_model("tp06")
    layer = mesh.info["endo_epi"].astype(int)
    prm = np.stack([info["layer_parameters"][k] for k in layer], axis=1)  # (np, N)
    y0 = np.repeat(om.init_state_values()[:, None], pts.shape[0], axis=1)
    opde = om_mono.MonodomainModel(mass, stiff, [om_mono.Stimulus.window(load, 0.0, 1.0, info["stim_amplitude"])], C_m=1.0, theta=0.5,
                                   solver="lu")
    oode = om_mono.ODESolver(v_pde=opde.state, init_states=y0, parameters=prm, fun=om.generalized_rush_larsen, num_states=19,
                             v_index=om.state_index("V"))
    ref = om_mono.SplittingSolver(opde, oode)
    t, dt = 0.0, 0.05
    for _ in range(30):
        solver.step((t, t + dt))
        ref.step((t, t + dt))
        t += dt
    v = solver.pde.state.x.array_ro
    assert v.max() > -82.0 and v.min() < -85.0  # the surface stimulus acts on the endocardium only (rest: -85.23 mV)
    assert np.abs(v - opde.state).max() <= 1e-8 * np.abs(opde.state).max()
    assert np.abs(solver.ode.full_values - oode.values).max() <= 1e-7 * np.abs(oode.values).max()
