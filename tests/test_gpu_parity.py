"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run on a B200: pytest -m gpu.

Tolerances (BASELINE.json north_star): one ODE step <= 1e-12 relative in fp64; one PDE step <= 1e-8
relative at matched (tight) solver tolerance.
"""
import numpy as np
import pytest

import _problems as P
from oracle import monodomain as om_mono

pytestmark = pytest.mark.gpu


def rel_err(a, b, y_in):
    """Element-wise error of one step relative to the size of the quantities the update adds up:
    max(|y_new|, |y_old|), floored at 1e-6 of the state's largest magnitude.  (An explicit update
    y + dt*f can cancel - forward Euler on the TP06 m gate at rest has dt/tau ~ 50 - so |y_new| alone is
    not the scale of the arithmetic.)"""
    scale = np.maximum(np.maximum(np.abs(b), np.abs(y_in)), 1e-6 * np.abs(b).max(axis=1, keepdims=True) + 1e-300)
    return float((np.abs(a - b) / scale).max())


@pytest.mark.parametrize("tag", ["fhn", "tp06", "torord"])
@pytest.mark.parametrize("scheme", ["forward_explicit_euler", "generalized_rush_larsen"])
@pytest.mark.parametrize("n", [1, 31, 20000])
def test_ode_single_step(ctx_factory, tag, scheme, n):
    import importlib

    om = P.oracle_model(tag)
    hm = importlib.import_module(f"beat_b200.models.{tag}")
    rng = np.random.default_rng(1234)
    states = P.perturbed_states(om, n, rng, P.V_NAME[tag])
    params = om.init_parameter_values()
    dev = getattr(hm, scheme)
    ctx = ctx_factory()
    ctx.ode_create(P.MODEL_ID[tag], P.SCHEME_ID[scheme], n, om.state_index(P.V_NAME[tag]), states.shape[0])
    ctx.ode_set_states(states)
    ctx.ode_set_params(params, dev.derived(params))
    for t0 in (0.0, 10.5):
        ctx.ode_set_states(states)
        ctx.ode_step(t0, 0.01)
        got = ctx.ode_get_states()
        with np.errstate(all="ignore"):
            want = getattr(om, scheme)(states, t0, 0.01, params)
        assert np.isfinite(got).all()
        err = rel_err(got, want, states)
        # Rush-Larsen increments f*(exp(lin*dt) - 1)/lin (gotranx writes exp(x) - 1, not expm1) cancel when |lin*dt| is
        # small: two correctly rounded exp() that differ by one ulp then differ by ~ulp*|f/lin| in the state.  On 20 000
        # random states per model that reaches ~2e-12 of the state's size (FitzHugh-Nagumo's cubic, the ToR-ORd membrane
        # potential; more while a stimulus current is on); forward Euler and the golden vectors stay below 1e-12.
        # The bar cannot be 1e-12 for this scheme in float64 at all: the ORACLE's own float64 evaluation is 3.7e-11 (TP06,
        # model stimulus on) / 2.9e-10 (ToR-ORd) / 4.2e-12 (FHN) away from an extended-precision evaluation of the same
        # formulas, on the membrane potential only (tests/test_grl1_rounding_bound.py, tools/grl1_rounding_bound.py).
        tol = 2e-11 if scheme == "generalized_rush_larsen" else 1e-12
        assert err <= tol, f"{tag}/{scheme} n={n} t={t0}: rel err {err:.3e}"
    ctx.close()


@pytest.mark.parametrize("tag", ["tp06"])
def test_ode_per_node_parameters(ctx_factory, tag):
    om = P.oracle_model(tag)
    rng = np.random.default_rng(7)
    n = 4097
    states = P.perturbed_states(om, n, rng, P.V_NAME[tag])
    p0 = om.init_parameter_values()
    params = np.repeat(p0[:, None], n, axis=1) * (1 + 0.05 * rng.uniform(-1, 1, (len(p0), n)))
    ctx = ctx_factory()
    ctx.ode_create(P.MODEL_ID[tag], 1, n, om.state_index("V"), states.shape[0])
    ctx.ode_set_states(states)
    ctx.ode_set_params(params)
    ctx.ode_step(3.0, 0.01)
    got = ctx.ode_get_states()
    want = om.generalized_rush_larsen(states, 3.0, 0.01, params)
    assert rel_err(got, want, states) <= 1e-12
    ctx.close()


def _pde_ctx(ctx_factory, prob, theta=0.5, rtol=1e-13, x0=0, pc=1, norm=0, max_it=1000, ksp=0):
    ctx = ctx_factory()
    mass, stiff = prob["mass"], prob["stiff"]
    n = mass.shape[0]
    ctx.pde_set_matrices(n, 0, mass.indptr, mass.indices, mass.data, stiff.data)
    ctx.pde_config(prob["C_m"], theta, rtol, 1e-50, max_it, pc, norm, x0)
    ctx.pde_set_ksp_type(ksp)
    return ctx


KSP = {"cg": 0, "pipecg": 1}


@pytest.mark.parametrize("ksp", ["cg", "pipecg"])
@pytest.mark.parametrize("x0", [0, 1])
def test_pde_single_step_tight(ctx_factory, x0, ksp):
    prob = P.niederer_slab(0.5)
    n = prob["mass"].shape[0]
    rng = np.random.default_rng(3)
    v_prev = -85.0 + 120.0 * rng.random(n)
    ctx = _pde_ctx(ctx_factory, prob, x0=x0, ksp=KSP[ksp], rtol=1e-13 if ksp == "cg" else 1e-12)
    idx = np.nonzero(prob["stim_load"])[0]
    ctx.stim_add(idx, prob["stim_load"][idx], 0.0, 2.0, prob["stim_amp"])
    ctx.set_v_prev(v_prev)
    ctx.pde_step(0.5, 0.55)
    got = ctx.get_v(np.empty(n))
    its, rnorm, reason = ctx.ksp_info()
    ref = om_mono.MonodomainModel(prob["mass"], prob["stiff"], [om_mono.Stimulus.window(prob["stim_load"], 0.0, 2.0, prob["stim_amp"])],
                                  C_m=prob["C_m"], theta=0.5, solver="lu")
    ref.v_[:] = v_prev
    ref.step((0.5, 0.55))
    err = np.abs(got - ref.state).max() / np.abs(ref.state).max()
    assert reason > 0, (its, rnorm, reason)
    assert err <= 1e-8, f"PDE step rel err {err:.3e} after {its} iterations"
    assert err <= 1e-10  # what a tight solve actually delivers
    ctx.close()


@pytest.mark.parametrize("ksp", ["cg", "pipecg"])
@pytest.mark.parametrize("norm", [0, 1, 2])
def test_pde_matches_petsc_style_cg_iteration_count(ctx_factory, ksp, norm):
    """Same algorithm, same tolerance -> same iteration count and the same iterate as the oracle's KSPCG
    restatement (pipelined CG produces the same iterates up to rounding)."""
    prob = P.niederer_slab(0.5)
    n = prob["mass"].shape[0]
    rng = np.random.default_rng(5)
    v_prev = -85.0 + 120.0 * rng.random(n)
    ctx = _pde_ctx(ctx_factory, prob, rtol=1e-5, ksp=KSP[ksp], norm=norm)
    ctx.set_v_prev(v_prev)
    ctx.pde_step(0.0, 0.05)
    got = ctx.get_v(np.empty(n))
    its, rnorm, reason = ctx.ksp_info()
    ref = om_mono.MonodomainModel(prob["mass"], prob["stiff"], [], C_m=prob["C_m"], theta=0.5, solver="cg-jacobi", rtol=1e-5,
                                  norm=["preconditioned", "unpreconditioned", "natural"][norm])
    ref.v_[:] = v_prev
    ref.step((0.0, 0.05))
    assert its == ref.ksp["iterations"], (its, ref.ksp)
    assert reason == ref.ksp["reason"]
    assert abs(rnorm - ref.ksp["residual_norm"]) <= 1e-6 * ref.ksp["residual_norm"]
    assert np.abs(got - ref.state).max() <= 1e-9 * np.abs(ref.state).max()
    ctx.close()


@pytest.mark.parametrize("ksp,stream", [("cg", False), ("cg", True), ("pipecg", False), ("pipecg", True)])
def test_pde_two_rows_per_thread(ctx_factory, ksp, stream, monkeypatch):
    """132k-row slab: every thread of the persistent kernel owns two rows (shared-memory resident and
    streaming variants of the pipelined solver), stimulus active, against the oracle's tight PCG."""
    if stream:
        monkeypatch.setenv("MONO_PDE_STREAM", "1")
    prob = P.niederer_slab(0.15)
    n = prob["mass"].shape[0]
    assert n > 148 * 512
    rng = np.random.default_rng(17)
    v_prev = -85.0 + 120.0 * rng.random(n)
    ctx = _pde_ctx(ctx_factory, prob, rtol=1e-12, ksp=KSP[ksp])
    idx = np.nonzero(prob["stim_load"])[0]
    ctx.stim_add(idx, prob["stim_load"][idx], 0.0, 2.0, prob["stim_amp"])
    ctx.set_v_prev(v_prev)
    ref = om_mono.MonodomainModel(prob["mass"], prob["stiff"], [om_mono.Stimulus.window(prob["stim_load"], 0.0, 2.0, prob["stim_amp"])],
                                  C_m=prob["C_m"], theta=0.5, solver="cg-jacobi", rtol=1e-13)
    ref.v_[:] = v_prev
    for t0, t1 in ((0.5, 0.51), (2.5, 2.51)):  # stimulus on, then off (the dense source vector is cleared)
        ctx.pde_step(t0, t1)
        ref.step((t0, t1))
        got = ctx.get_v(np.empty(n))
        its, rnorm, reason = ctx.ksp_info()
        assert reason > 0, (its, rnorm, reason)
        assert np.abs(got - ref.state).max() <= 1e-9 * np.abs(ref.state).max()
    ctx.close()


def test_pde_dt_change_rebuilds_matrices(ctx_factory):
    prob = P.niederer_slab(1.0)
    n = prob["mass"].shape[0]
    rng = np.random.default_rng(11)
    v0 = rng.random(n)
    ctx = _pde_ctx(ctx_factory, prob)
    ref = om_mono.MonodomainModel(prob["mass"], prob["stiff"], [], C_m=prob["C_m"], theta=0.5, solver="lu")
    ctx.set_v_prev(v0)
    ref.v_[:] = v0
    t = 0.0
    for dt in (0.05, 0.05, 0.2, 0.01):
        ctx.pde_step(t, t + dt)
        ref.step((t, t + dt))
        ctx.pde_assign_previous()
        ref.assign_previous()
        t += dt
    got = ctx.get_v(np.empty(n))
    assert np.abs(got - ref.state).max() <= 1e-9 * np.abs(ref.state).max()
    ctx.close()


@pytest.mark.parametrize("ksp", ["cg", "pipecg"])
@pytest.mark.parametrize("theta_split", [1.0, 0.5])
def test_split_steps_niederer_small(ctx_factory, theta_split, ksp):
    """A few fused split steps (TP06 GRL1 + CN diffusion + S1 stimulus) against the oracle's literal
    restatement of MonodomainSplittingSolver.step."""
    import beat_b200.models.tp06 as hm

    prob = P.niederer_slab(0.5)
    n = prob["mass"].shape[0]
    om = P.oracle_model("tp06")
    params = om.init_parameter_values(stim_amplitude=0.0)
    y0 = om.init_state_values()
    stim = om_mono.Stimulus.window(prob["stim_load"], 0.0, 2.0, prob["stim_amp"])
    pde = om_mono.MonodomainModel(prob["mass"], prob["stiff"], [stim], C_m=prob["C_m"], theta=0.5, solver="lu")
    ode = om_mono.ODESolver(v_pde=pde.state, init_states=y0, parameters=params, fun=om.generalized_rush_larsen,
                            num_states=len(y0), v_index=om.state_index("V"))
    ref = om_mono.SplittingSolver(pde, ode, theta=theta_split)

    ctx = _pde_ctx(ctx_factory, prob, rtol=1e-12, ksp=KSP[ksp])
    idx = np.nonzero(prob["stim_load"])[0]
    ctx.stim_add(idx, prob["stim_load"][idx], 0.0, 2.0, prob["stim_amp"])
    ctx.ode_create(1, 1, n, om.state_index("V"), len(y0))
    ctx.ode_set_states(np.repeat(y0[:, None], n, axis=1))
    ctx.ode_set_params(params, hm.generalized_rush_larsen.derived(params))
    # MonodomainSplittingSolver.__post_init__ (monodomain_solver.py:33-37)
    ctx.ode_to_dolfin()
    ctx.ode_to_pde()
    ctx.pde_assign_previous()

    dt, nsteps = 0.05, 40
    t = 0.0
    for k in range(nsteps):
        ref.step((t, t + dt))
        ctx.split_step(t, t + dt, theta_split)
        t += dt
    got_v = ctx.get_v(np.empty(n))
    got_states = ctx.ode_get_states()
    assert np.abs(got_v - pde.state).max() <= 1e-8 * np.abs(pde.state).max()
    scale = np.abs(ode.values).max(axis=1, keepdims=True)
    assert (np.abs(got_states - ode.values) / scale).max() <= 1e-8
    # post-condition of a split step: states[V] == v == v_
    assert np.array_equal(got_states[om.state_index("V")], got_v)
    assert np.array_equal(ctx.get_v_prev(np.empty(n)), got_v)
    assert pde.state.max() > 0.0, "stimulated corner must have depolarised"
    ctx.close()


# ---------------------------------------------------------------------------------------- Chebyshev preconditioner
def _cheb_pcg_numpy(A, b, x0, steps, kappa, rtol, norm):
    """Plain restatement of KSPCG with M^-1 = p(D^-1 A) D^-1 (Chebyshev iteration on [g/kappa, g], g = Gershgorin
    bound; the device's mono_pde_set_chebyshev) and PETSc's default convergence test.  Returns (x, its, rnorm)."""
    d = A.diagonal()
    dinv = 1.0 / d
    g = (abs(A).sum(axis=1).A1 / d).max() * (1.0 + 1e-12)
    lo = g / kappa
    theta, delta = 0.5 * (g + lo), 0.5 * (g - lo)
    sigma = theta / delta

    def M(w):
        gg = dinv * w
        y = gg / theta
        dv = y.copy()
        rho = 1.0 / sigma
        for _ in range(1, steps):
            rho_n = 1.0 / (2.0 * sigma - rho)
            dv = rho_n * rho * dv + (2.0 * rho_n / delta) * (gg - dinv * (A @ y))
            y = y + dv
            rho = rho_n
        return y

    def nrm(r, z):
        return np.sqrt({0: z @ z, 1: r @ r, 2: abs(r @ z)}[norm])

    x = x0.copy()
    r = b - A @ x
    z = M(r)
    ttol = rtol * nrm(b, M(b))
    p = z.copy()
    rz = r @ z
    its = 0
    rn = nrm(r, z)
    while rn > ttol and its < 1000:
        q = A @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = M(r)
        its += 1
        rn = nrm(r, z)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, its, rn


@pytest.mark.parametrize("steps", [2, 3, 4])
@pytest.mark.parametrize("x0", [0, 1])
@pytest.mark.parametrize("norm", [0, 1, 2])
def test_pde_chebyshev_preconditioner(ctx_factory, steps, x0, norm):
    """pc_type chebyshev in the pipelined driver: (a) tight solve == LU oracle; (b) at the PETSc-default rtol the
    iteration count, residual norm and iterate equal the plain NumPy restatement of the same preconditioned CG."""
    prob = P.niederer_slab(0.5)
    n = prob["mass"].shape[0]
    rng = np.random.default_rng(9)
    v_prev = -85.0 + 120.0 * rng.random(n)
    dt, kappa = 0.05, 4.0
    A = (prob["C_m"] * prob["mass"] + 0.5 * dt * prob["stiff"]).tocsr()
    B = (prob["C_m"] * prob["mass"] - 0.5 * dt * prob["stiff"]).tocsr()
    idx = np.nonzero(prob["stim_load"])[0]
    b = B @ v_prev + dt * prob["stim_amp"] * prob["stim_load"]
    for rtol in (1e-12, 1e-5):
        ctx = ctx_factory()
        ctx.pde_set_matrices(n, 0, prob["mass"].indptr, prob["mass"].indices, prob["mass"].data, prob["stiff"].data)
        ctx.pde_set_chebyshev(steps, kappa)
        ctx.pde_config(prob["C_m"], 0.5, rtol, 1e-50, 1000, 2, norm, x0)
        ctx.pde_set_ksp_type(KSP["pipecg"])
        ctx.stim_add(idx, prob["stim_load"][idx], 0.0, 2.0, prob["stim_amp"])
        ctx.set_v_prev(v_prev)
        ctx.pde_step(0.5, 0.5 + dt)
        got = ctx.get_v(np.empty(n))
        its, rnorm, reason = ctx.ksp_info()
        assert reason > 0, (its, rnorm, reason)
        if rtol < 1e-10:
            import scipy.sparse.linalg as sla

            exact = sla.spsolve(A.tocsc(), b)
            assert np.abs(got - exact).max() <= 1e-9 * np.abs(exact).max()
        else:
            xr, its_r, rn_r = _cheb_pcg_numpy(A, b, v_prev.copy() if x0 else np.zeros(n), steps, kappa, rtol, norm)
            assert abs(its - its_r) <= 1, (its, its_r)
            if its == its_r:
                assert abs(rnorm - rn_r) <= 1e-5 * rn_r
                assert np.abs(got - xr).max() <= 1e-9 * np.abs(xr).max()
        ctx.close()


def test_pde_chebyshev_needs_pipecg(ctx_factory):
    prob = P.niederer_slab(0.5)
    n = prob["mass"].shape[0]
    ctx = _pde_ctx(ctx_factory, prob, pc=2, ksp=KSP["cg"])
    ctx.set_v_prev(np.full(n, -80.0))
    with pytest.raises(Exception, match="pipecg"):
        ctx.pde_step(0.0, 0.05)
    ctx.close()
