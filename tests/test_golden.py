"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the NumPy oracle).

CPU part (`-m "not gpu"`): the two independent restatements of the oracle - the NumPy modules under
oracle/models + oracle/monodomain.py and the OpenMP C port oracle/c/oracle_step.c - both reproduce the
fixtures, so a silent change to either shows up here.  GPU part: the CUDA path, through the C ABI,
reproduces the same fixtures (ODE step <= 1e-12, PDE step <= 1e-8, 12 split steps <= 1e-8).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import _problems as P
from oracle import cport
from oracle import monodomain as om_mono

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["fhn", "tp06", "torord"]
SCHEMES = ["forward_explicit_euler", "generalized_rush_larsen"]


def load(name):
    return np.load(os.path.join(GOLD, name))


def step_scale(out, inp):
    return np.maximum(np.maximum(np.abs(out), np.abs(inp)), 1e-6 * np.abs(out).max(axis=1, keepdims=True) + 1e-300)


def slab_matrices(g):
    n = len(g["indptr"]) - 1
    mass = sp.csr_matrix((g["mass"], g["indices"], g["indptr"]), shape=(n, n))
    stiff = sp.csr_matrix((g["stiff"], g["indices"], g["indptr"]), shape=(n, n))
    return n, mass, stiff


# ------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("tag", MODELS)
@pytest.mark.parametrize("scheme", SCHEMES)
def test_numpy_oracle_reproduces_ode_golden(tag, scheme):
    g = load(f"ode_{tag}_{scheme}.npz")
    om = P.oracle_model(tag)
    with np.errstate(all="ignore"):
        out = getattr(om, scheme)(g["states_in"], float(g["t"]), float(g["dt"]), g["params"])
    assert np.array_equal(out, g["states_out"])  # same code, same inputs: bit-exact


@pytest.mark.parametrize("tag", MODELS)
@pytest.mark.parametrize("scheme", SCHEMES)
def test_c_port_reproduces_ode_golden(tag, scheme):
    g = load(f"ode_{tag}_{scheme}.npz")
    s = np.ascontiguousarray(g["states_in"].copy())
    cport.ode_step(tag, scheme, s, g["params"], float(g["t"]), float(g["dt"]))
    err = (np.abs(s - g["states_out"]) / step_scale(g["states_out"], g["states_in"])).max()
    # libm vs NumPy exp/log differ by an ulp; forward Euler on the stiff TP06 m gate (dt/tau ~ 50) amplifies that
    assert err <= (2e-12 if scheme == "forward_explicit_euler" else 2e-13), err


def test_oracles_reproduce_pde_golden():
    g = load("pde_slab.npz")
    n, mass, stiff = slab_matrices(g)
    stim = (g["load"], float(g["t_start"]), float(g["t_end"]), float(g["amp"]))
    model = om_mono.MonodomainModel(mass, stiff, [om_mono.Stimulus.window(*stim)], C_m=float(g["C_m"]), theta=float(g["theta"]),
                                    solver="lu")
    model.v_[:] = g["v_prev"]
    model.step((float(g["t0"]), float(g["t1"])))
    assert np.abs(model.state - g["v_out"]).max() <= 1e-12 * np.abs(g["v_out"]).max()
    spc = cport.SplitProblem(mass, stiff, float(g["C_m"]), float(g["theta"]), float(g["t1"]) - float(g["t0"]), [stim], rtol=1e-14)
    x, its, _ = spc.pde_step(float(g["t0"]), float(g["t1"]), g["v_prev"])
    assert its > 0
    assert np.abs(x - g["v_out"]).max() <= 1e-10 * np.abs(g["v_out"]).max()


@pytest.mark.parametrize("name,theta", [("godunov", 1.0), ("strang", 0.5)])
def test_c_port_reproduces_split_golden(name, theta):
    gp, gs = load("pde_slab.npz"), load("split_slab.npz")
    n, mass, stiff = slab_matrices(gp)
    stim = (gp["load"], float(gp["t_start"]), float(gp["t_end"]), float(gp["amp"]))
    spc = cport.SplitProblem(mass, stiff, float(gp["C_m"]), float(gp["theta"]), float(gs["dt"]), [stim], rtol=1e-14)
    om = P.oracle_model("tp06")
    states = np.ascontiguousarray(np.repeat(gs["y0"][:, None], n, axis=1))
    v, _ = spc.split_steps("tp06", "generalized_rush_larsen", om.state_index("V"), states, gs["params"], 0.0, int(gs["nsteps"]), theta)
    assert np.abs(v - gs[f"v_{name}"]).max() <= 1e-9 * np.abs(gs[f"v_{name}"]).max()
    ref = gs[f"states_{name}"]
    assert (np.abs(states - ref) / (np.abs(ref).max(axis=1, keepdims=True))).max() <= 1e-9


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", MODELS)
@pytest.mark.parametrize("scheme", SCHEMES)
def test_cuda_reproduces_ode_golden(ctx_factory, tag, scheme):
    import importlib

    g = load(f"ode_{tag}_{scheme}.npz")
    om = P.oracle_model(tag)
    hm = importlib.import_module(f"beat_b200.models.{tag}")
    s = np.ascontiguousarray(g["states_in"])
    ctx = ctx_factory()
    ctx.ode_create(P.MODEL_ID[tag], P.SCHEME_ID[scheme], s.shape[1], om.state_index(P.V_NAME[tag]), s.shape[0])
    ctx.ode_set_states(s)
    ctx.ode_set_params(g["params"], getattr(hm, scheme).derived(g["params"]))
    ctx.ode_step(float(g["t"]), float(g["dt"]))
    got = ctx.ode_get_states()
    err = (np.abs(got - g["states_out"]) / step_scale(g["states_out"], g["states_in"])).max()
    assert err <= 1e-12, err
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ksp", [0, 1])
def test_cuda_reproduces_pde_golden(ctx_factory, ksp):
    g = load("pde_slab.npz")
    n = len(g["indptr"]) - 1
    ctx = ctx_factory()
    ctx.pde_set_matrices(n, 0, g["indptr"], g["indices"], g["mass"], g["stiff"])
    ctx.pde_config(float(g["C_m"]), float(g["theta"]), 1e-13, 1e-50, 1000, 1, 0, 0)
    ctx.pde_set_ksp_type(ksp)
    idx = np.nonzero(g["load"])[0].astype(np.int32)
    ctx.stim_add(idx, g["load"][idx], float(g["t_start"]), float(g["t_end"]), float(g["amp"]))
    ctx.set_v_prev(g["v_prev"])
    ctx.pde_step(float(g["t0"]), float(g["t1"]))
    got = ctx.get_v(np.empty(n))
    its, rnorm, reason = ctx.ksp_info()
    assert reason > 0, (its, rnorm, reason)
    assert np.abs(got - g["v_out"]).max() <= 1e-8 * np.abs(g["v_out"]).max()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,theta", [("godunov", 1.0), ("strang", 0.5)])
def test_cuda_reproduces_split_golden(ctx_factory, name, theta):
    from beat_b200.models import tp06

    gp, gs = load("pde_slab.npz"), load("split_slab.npz")
    n = len(gp["indptr"]) - 1
    om = P.oracle_model("tp06")
    ctx = ctx_factory()
    ctx.pde_set_matrices(n, 0, gp["indptr"], gp["indices"], gp["mass"], gp["stiff"])
    ctx.pde_config(float(gp["C_m"]), float(gp["theta"]), 1e-13, 1e-50, 1000, 1, 0, 0)
    idx = np.nonzero(gp["load"])[0].astype(np.int32)
    ctx.stim_add(idx, gp["load"][idx], float(gp["t_start"]), float(gp["t_end"]), float(gp["amp"]))
    ns = len(gs["y0"])
    ctx.ode_create(P.MODEL_ID["tp06"], 1, n, om.state_index("V"), ns)
    states = np.ascontiguousarray(np.repeat(gs["y0"][:, None], n, axis=1))
    ctx.ode_set_states(states)
    ctx.ode_set_params(gs["params"], tp06.generalized_rush_larsen.derived(gs["params"]))
    v0 = states[om.state_index("V")].copy()
    ctx.set_v(v0)
    ctx.set_v_prev(v0)
    ctx.set_v_ode(v0)
    t, dt = 0.0, float(gs["dt"])
    for _ in range(int(gs["nsteps"])):
        ctx.split_step(t, t + dt, theta)
        t += dt
    got_v = ctx.get_v(np.empty(n))
    got_s = ctx.ode_get_states()
    assert np.abs(got_v - gs[f"v_{name}"]).max() <= 1e-8 * np.abs(gs[f"v_{name}"]).max()
    ref = gs[f"states_{name}"]
    assert (np.abs(got_s - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 1e-8
    ctx.close()
