"""Pins the CPU oracle against the reference's OWN known-answer tests for the split-step path
(SURVEY.md section 8c).  Each test restates one reference test (file:line cited) with the oracle's classes
in place of dolfinx/PETSc.  Runs on CPU (`-m "not gpu"`).

The cell-model arithmetic (gotranx boundary) has no reference test: that part of the oracle stays
"parity unpinned" except for the Niederer activation-time table checked in test_oracle_niederer.py.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem
from oracle import monodomain as om


def _ode_fe(states, t, dt, parameters):  # tests/test_odesolver.py:11-17 (harmonic oscillator, forward Euler)
    v, s = states
    out = np.zeros_like(states)
    out[0] = v - s * dt
    out[1] = s + v * dt
    return out


def test_ode_forward_euler_convergence_rate():
    """tests/test_odesolver.py:20-49, literally: v'=-s, s'=v from (1, 0); samples at t = 0.1 .. 1.0; the norm
    of the sampled error drops by 10 per decade of dt (rate 1 +- 0.01)."""
    t_bound, t0 = 1.0, 0.0
    x = np.arange(0.1, t_bound + 0.1, 0.1)
    y = np.zeros((len(x), 2))
    sol = np.vstack((np.cos(x), np.sin(x))).T
    errors = []
    for dt in [0.1, 0.01, 0.001, 0.0001]:
        states = np.zeros((2, 1))
        states.T[:] = [1, 0]
        j, t = 0, 0.0
        for _ in range(int((t_bound - t0) / dt)):
            states[:] = _ode_fe(states, t, dt, None)  # ODESystemSolver.step, odesolver.py:67-79
            t += dt
            if j < len(x) and np.isclose(t, x[j]):
                y[j, :] = states[:, 0]
                j += 1
        errors.append(np.linalg.norm(sol - y))
    rates = [np.log(e1 / e2) / np.log(10) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert np.allclose(rates, 1, atol=0.01), rates


def test_dolfin_ode_solver_data_movement():
    """tests/test_odesolver.py:52-117: shapes, one-step closed form, to_dolfin/ode_to_pde/pde_to_ode/from_dolfin."""
    pts, cells = fem.rectangle_mesh(5, 5)
    n = pts.shape[0]
    v_pde = np.zeros(n)
    v0, s0, dt = 1.0, 2.0, 0.1
    ode = om.ODESolver(v_pde=v_pde, init_states=np.array([v0, s0]), parameters=None,
                       fun=lambda states, t, parameters, dt: _ode_fe(states, t, dt, parameters), num_states=2, v_index=0)
    assert ode.values.shape == (2, n)
    assert np.allclose(ode.values[0], v0) and np.allclose(ode.values[1], s0)
    ode.step(0.0, dt)
    assert np.allclose(ode.values[0], v0 - s0 * dt)  # :89
    assert np.allclose(ode.values[1], s0 + v0 * dt)  # :90
    assert np.allclose(ode.v_ode, 0.0)  # :92-93 the dolfin function is untouched by step
    ode.to_dolfin()
    assert np.allclose(ode.v_ode, v0 - s0 * dt)  # :96
    assert np.allclose(v_pde, 0.0)
    ode.ode_to_pde()
    assert np.allclose(v_pde, v0 - s0 * dt)  # :101
    v_pde[:] = 1.0
    ode.pde_to_ode()
    assert np.allclose(ode.v_ode, 1.0)  # :106
    ode.from_dolfin()
    assert np.allclose(ode.values[0], 1.0)  # :110
    assert np.allclose(ode.values[1], s0 + v0 * dt)


def _separable(pts, cells, g, h, degree=8):
    load = fem.load_vector_function(pts, cells, g, degree)
    return om.Stimulus(load, h)


G = staticmethod(lambda x: np.cos(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1]))


def _g(x):
    return np.cos(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1])


@pytest.mark.parametrize("M,err", [(0.0, 1e-4), (1.0, 2e-4), (2.0, 2e-4)])
def test_monodomain_analytic(M, err):
    """tests/test_monodomain.py:38-64: PDE-only MMS, N=15, theta=0.5, dt=1e-3, 10 steps, direct solve."""
    N, dt = 15, 0.001
    T = 10 * dt
    pts, cells = fem.rectangle_mesh(N, N)
    mass, stiff = fem.assemble_p1(pts, cells, M * np.eye(2))
    stim = _separable(pts, cells, _g, lambda t: np.cos(t) + M * 8 * np.pi**2 * np.sin(t))
    model = om.MonodomainModel(mass, stiff, [stim], C_m=1.0, theta=0.5, solver="lu")
    state = model.solve((0, T), dt=dt)
    e = fem.l2_error(pts, cells, state, lambda x: _g(x) * np.sin(T))
    assert e < err, e


def test_monodomain_spatial_convergence():
    """tests/test_monodomain.py:67-104: rates >= 2.0 for N = 4, 8, 16, 32."""
    dt = 0.001
    T = 10 * dt
    errors = []
    for N in (4, 8, 16, 32):
        pts, cells = fem.rectangle_mesh(N, N)
        mass, stiff = fem.assemble_p1(pts, cells, np.eye(2))
        stim = _separable(pts, cells, _g, lambda t: np.cos(t) + 8 * np.pi**2 * np.sin(t))
        model = om.MonodomainModel(mass, stiff, [stim], theta=0.5, solver="lu")
        state = model.solve((0, T), dt=dt)
        errors.append(fem.l2_error(pts, cells, state, lambda x: _g(x) * np.sin(T)))
    rates = [np.log(e1 / e2) / np.log(2) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert all(r >= 2.0 for r in rates), rates


def test_monodomain_temporal_convergence():
    """tests/test_monodomain.py:107-147: theta=0.5 temporal rate >= 2 at N=100, dt = 1 .. 1/8, T=1.
    (M=0 there: the source is cos(t) * g, exact solution sin(t) * g projected.)"""
    N, T = 60, 1.0
    pts, cells = fem.rectangle_mesh(N, N)
    mass, stiff = fem.assemble_p1(pts, cells, 0.0 * np.eye(2))
    stim = _separable(pts, cells, _g, lambda t: np.cos(t))
    # exact discrete-in-space solution: C_m Mass v' = cos(t) * load -> v = sin(t) * Mass^-1 load
    import scipy.sparse.linalg as spla

    vg = spla.spsolve(mass.tocsc(), stim.load)
    errors = []
    for dt in (1.0, 0.5, 0.25, 0.125):
        model = om.MonodomainModel(mass, stiff, [stim], theta=0.5, solver="lu")
        state = model.solve((0, T), dt=dt)
        errors.append(np.sqrt((state - np.sin(T) * vg) @ (mass @ (state - np.sin(T) * vg))))
    rates = [np.log(e1 / e2) / np.log(2) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert all(r >= 1.95 for r in rates), rates


def test_monodomain_splitting_analytic():
    """tests/test_monodomain_solver.py:41-87 (P1 ODE space): Godunov split, FE ODE v'=-s, s'=v,
    I_s = 8 pi^2 g sin(t), M=1, N=50, dt=0.01, T=1 -> L2 error < 0.002."""
    N, dt, T = 50, 0.01, 1.0
    pts, cells = fem.rectangle_mesh(N, N)
    mass, stiff = fem.assemble_p1(pts, cells, np.eye(2))
    stim = _separable(pts, cells, _g, lambda t: 8 * np.pi**2 * np.sin(t), degree=6)
    pde = om.MonodomainModel(mass, stiff, [stim], theta=0.5, solver="lu")
    init = np.zeros((2, pts.shape[0]))
    init[1] = -_g(pts.T) * np.cos(0.0)  # s_exact at t = 0
    ode = om.ODESolver(v_pde=pde.state, init_states=init, parameters=None,
                       fun=lambda states, t, parameters, dt: _ode_fe(states, t, dt, parameters), num_states=2, v_index=0)
    solver = om.SplittingSolver(pde, ode)
    solver.solve((0.0, T), dt=dt)
    # the reference evaluates v_exact at time.value = t of the last PDE step (theta point): :79-85
    e = fem.l2_error(pts, cells, pde.state, lambda x: _g(x) * np.sin(pde.time))
    assert e < 0.002, e


def test_monodomain_splitting_spatial_convergence():
    """tests/test_monodomain_solver.py:98-149 (P1 ODE space): dt=1e-3, T=1, N = 8, 16, 32, theta_split = 1;
    mean spatial rate > 1.85; v_exact evaluated at time.value of the last PDE step as the reference does."""
    dt, T = 0.001, 1.0
    errors = []
    for N in (8, 16, 32):
        pts, cells = fem.rectangle_mesh(N, N)
        mass, stiff = fem.assemble_p1(pts, cells, np.eye(2))
        stim = _separable(pts, cells, _g, lambda t: 8 * np.pi**2 * np.sin(t), degree=6)
        pde = om.MonodomainModel(mass, stiff, [stim], theta=0.5, solver="lu")
        init = np.zeros((2, pts.shape[0]))
        init[1] = -_g(pts.T)
        ode = om.ODESolver(v_pde=pde.state, init_states=init, parameters=None,
                           fun=lambda states, t, parameters, dt: _ode_fe(states, t, dt, parameters), num_states=2, v_index=0)
        om.SplittingSolver(pde, ode, theta=1.0).solve((0.0, T), dt=dt)
        errors.append(fem.l2_error(pts, cells, pde.state, lambda x: _g(x) * np.sin(pde.time)))
    rates = [np.log(e1 / e2) / np.log(2) for e1, e2 in zip(errors[:-1], errors[1:])]
    assert sum(rates) / len(rates) > 1.85, (rates, errors)


def _interval_problem():
    pts, cells = fem.interval_mesh(10)
    mass, stiff = fem.assemble_p1(pts, cells, np.zeros((1, 1)))
    load = fem.load_vector_cells(pts, cells)
    return pts, mass, stiff, load


def test_single_stimulation():
    """tests/test_stimulation.py:12-46: M = 0 on the unit interval; exact integrals of the stimulus window.
    As in the reference there is no assign_previous between solve() calls: solve() leaves v_ one step behind
    (base_model.py:250-297), which is what the "- dt" in the expected values encodes."""
    pts, mass, stiff, load = _interval_problem()
    value, end, start, dt = 2.0, 1.0, 0.5, 0.01
    pde = om.MonodomainModel(mass, stiff, [om.Stimulus.window(load, start, end, value)], theta=0.5, solver="lu")
    pde.step((0.0, 0.4))
    assert np.allclose(pde.state, 0.0)
    t0 = 0.9
    pde.solve((0.4, t0), dt=dt)
    assert np.allclose(pde.state, value * (t0 - start))
    pde.solve((t0, end + dt), dt=dt)
    assert np.allclose(pde.state, (end - start - dt) * value)
    pde.solve((end + dt, 2 * end), dt=dt)
    assert np.allclose(pde.state, (end - start - dt) * value)


def test_double_stimulation():
    """tests/test_stimulation.py:49-107."""
    pts, mass, stiff, load = _interval_problem()
    dt, value1, value2, start1, end1, start2, end2 = 0.01, 2.0, 3.0, 0.5, 1.0, 0.9, 1.5
    pde = om.MonodomainModel(mass, stiff, [om.Stimulus.window(load, start1, end1, value1),
                                           om.Stimulus.window(load, start2, end2, value2)], theta=0.5, solver="lu")
    pde.step((0.0, 0.4))
    assert np.allclose(pde.state, 0.0)
    t0 = 0.9
    pde.solve((0.4, t0), dt=dt)
    assert np.allclose(pde.state, value1 * (t0 - start1))
    pde.solve((t0, end1 + dt), dt=dt)
    assert np.allclose(pde.state, (end1 - start1 - dt) * value1 + (end1 + dt - start2) * value2)
    pde.solve((end1 + dt, end2 + dt), dt=dt)
    want = (end1 - start1 - dt) * value1 + (end2 - start2 - dt) * value2
    assert np.allclose(pde.state, want)
    pde.solve((end2 + dt, 2 * end2), dt=dt)
    assert np.allclose(pde.state, want)


def test_p1_matrices_basic_identities():
    """Sanity of the oracle's own assembly: sum(Mass) = |Omega|, K 1 = 0, symmetry; box mesh has the 15-entry
    interior stencil of a 6-tet-per-hex split."""
    pts, cells = fem.box_mesh((4, 3, 2), (0, 0, 0), (2.0, 1.5, 1.0))
    mass, stiff = fem.assemble_p1(pts, cells, np.diag([3.0, 2.0, 1.0]))
    assert np.isclose(mass.sum(), 3.0)
    assert np.abs(stiff @ np.ones(pts.shape[0])).max() < 1e-12
    assert abs(mass - mass.T).max() < 1e-15 and abs(stiff - stiff.T).max() < 1e-13
    pts, cells = fem.box_mesh((6, 6, 6))
    mass, _ = fem.assemble_p1(pts, cells, np.eye(3))
    nnz_row = np.diff(sp.csr_matrix(mass).indptr)
    assert nnz_row.max() == 15
