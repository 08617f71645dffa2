"""Parity at BASELINE.json's larger sizes, where the oracle cannot follow in seconds, through size-independent
properties (tests/_properties.py): the diffusion step of the 3.4 M-dof Niederer slab (dx = 0.05 mm) must satisfy its
own linear system and conserve charge, evaluated on the host with SciPy products only; the cell-model kernel at 1e7
(TP06) / 1e6 (ToR-ORd) nodes must be pointwise - identical nodes give bit-identical results wherever they sit - and
agree with the oracle on the distinct ones; a Delaunay mesh with rows twice as wide as a structured mesh has (the shape of
a real LV mesh) against a sparse LU.  Runs last (several GB of host memory, ~1 min)."""

import importlib

import numpy as np
import pytest

import _problems as P
from _properties import pde_step_defects

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def slab_005():
    from beat_b200 import fem

    sl, st = P.niederer_conductivities()
    mesh = fem.create_box(fem.COMM_SELF, [np.zeros(3), np.array([20.0, 7.0, 3.0])], [400, 140, 60])
    indptr, indices, mass, stiff = fem.assemble_p1_local(mesh, np.diag([sl, st, st]))
    n = indptr.size - 1
    assert n == 401 * 141 * 61
    rng = np.random.default_rng(5)
    x = mesh.geometry.x
    v_prev = -85.0 + 120.0 * np.exp(-((x[:, 0] - 6.0) ** 2 + (x[:, 1] - 3.0) ** 2) / 4.0) + 0.5 * rng.random(n)
    corner = np.nonzero((x[:, 0] <= 1.5) & (x[:, 1] <= 1.5) & (x[:, 2] <= 1.5))[0].astype(np.int32)
    load = np.full(corner.size, 1.25e-4)  # ~ the lumped load of a 0.05 mm vertex patch
    return dict(indptr=indptr, indices=indices, mass=mass, stiff=stiff, n=n, v_prev=v_prev, stim_idx=corner, stim_val=load)


def _solve(ctx_factory, s, ksp, x0, steps=((0.0, 0.01),)):
    C_m, theta, amp = 0.01, 0.5, 0.357
    ctx = ctx_factory()
    ctx.pde_set_matrices(s["n"], 0, s["indptr"], s["indices"], s["mass"], s["stiff"])
    ctx.pde_config(C_m, theta, 1e-11, 1e-50, 500, 1, 0, x0)
    ctx.pde_set_ksp_type(ksp)
    ctx.stim_add(s["stim_idx"], s["stim_val"], 0.0, 2.0, amp)
    ctx.set_v_prev(s["v_prev"])
    out = []
    for t0, t1 in steps:
        ctx.pde_step(t0, t1)
        out.append((ctx.get_v(np.empty(s["n"])), ctx.ksp_info()))
    info = ctx.pde_dictionary_info()
    ctx.close()
    return out, info, (C_m, theta, amp)


@pytest.mark.parametrize("ksp,x0", [(0, 0), (1, 0), (0, 1)])
def test_diffusion_step_3p4M_dofs_solves_its_system_and_conserves_charge(ctx_factory, slab_005, ksp, x0):
    s = slab_005
    dt = 0.01
    (res,), _, (C_m, theta, amp) = _solve(ctx_factory, s, ksp, x0)
    x, (its, rnorm, reason) = res
    assert reason > 0 and 3 <= its <= 200, (its, rnorm, reason)
    source = np.zeros(s["n"])
    source[s["stim_idx"]] = amp * s["stim_val"]
    residual, conservation = pde_step_defects(s["indptr"], s["indices"], s["mass"], s["stiff"], C_m, theta, dt, s["v_prev"], source, x)
    assert residual <= 1e-9, residual        # (rtol 1e-11 on the preconditioned norm)
    assert conservation <= 1e-9, conservation


@pytest.mark.parametrize("tag,n_nodes", [("tp06", 10_000_000), ("torord", 1_000_000)])
def test_cell_model_kernel_is_pointwise_at_full_size(ctx_factory, tag, n_nodes):
    om = P.oracle_model(tag)
    hm = importlib.import_module(f"beat_b200.models.{tag}")
    block = 1000
    rng = np.random.default_rng(77)
    distinct = P.perturbed_states(om, block, rng, P.V_NAME[tag])
    states = np.ascontiguousarray(np.tile(distinct, (1, n_nodes // block)))
    params = om.init_parameter_values()
    ctx = ctx_factory()
    ctx.ode_create(P.MODEL_ID[tag], 1, n_nodes, om.state_index(P.V_NAME[tag]), states.shape[0])
    ctx.ode_set_states(states)
    ctx.ode_set_params(params, hm.generalized_rush_larsen.derived(params))
    ctx.ode_step(1.0, 0.01)
    got = ctx.ode_get_states(states)  # (overwrites the input buffer: one 1.5 GB array is enough)
    ctx.close()
    first = got[:, :block].copy()
    assert np.isfinite(first).all()
    assert (got.reshape(got.shape[0], -1, block) == first[:, None, :]).all()  # same node data -> same bits, anywhere
    with np.errstate(all="ignore"):
        want = om.generalized_rush_larsen(distinct, 1.0, 0.01, params)
    scale = np.maximum(np.maximum(np.abs(want), np.abs(distinct)), 1e-6 * np.abs(want).max(axis=1, keepdims=True) + 1e-300)
    assert float((np.abs(first - want) / scale).max()) <= 2e-11


def test_stencil_dictionary_at_3p4M_dofs_is_bit_identical(ctx_factory, slab_005, monkeypatch):
    """The dictionary kernel (default) against the SELL kernel (MONO_PDE_DICT=0) at full size: many slices per warp,
    two steps with a dt change in between."""
    steps = ((0.0, 0.01), (0.01, 0.03))
    monkeypatch.setenv("MONO_PDE_DICT", "0")
    ref, info0, _ = _solve(ctx_factory, slab_005, 0, 0, steps)
    monkeypatch.delenv("MONO_PDE_DICT", raising=False)
    got, info1, _ = _solve(ctx_factory, slab_005, 0, 0, steps)
    assert not info0["active"] and info1["active"] and info1["patterns"] == 27 and info1["rows_covered"] == 1.0
    for (xa, ka), (xb, kb) in zip(ref, got):
        assert ka == kb
        assert np.array_equal(xa, xb)


def _delaunay_problem():
    """A genuinely unstructured tetrahedral mesh (SciPy Delaunay of random points): vertex degrees far above the 15 of a
    Kuhn-split box, so SELL slices are wider than one 16-entry batch - the shape of a real LV mesh (BASELINE config 5)."""
    from scipy.spatial import Delaunay

    from beat_b200._lib import fem_assemble_p1

    rng = np.random.default_rng(42)
    pts = rng.random((6000, 3)) * [4.0, 2.0, 1.0]
    cells = Delaunay(pts).simplices.astype(np.int64)
    e = pts[cells[:, 1:]] - pts[cells[:, :1]]
    vol = np.abs(np.linalg.det(e)) / 6.0
    longest = np.max(np.linalg.norm(pts[cells][:, :, None, :] - pts[cells][:, None, :, :], axis=-1), axis=(1, 2))
    cells = cells[vol / longest**3 > 5e-3]  # no slivers
    used = np.unique(cells)
    renum = np.full(pts.shape[0], -1, dtype=np.int64)
    renum[used] = np.arange(used.size)
    pts, cells = pts[used], renum[cells]
    A = rng.standard_normal((cells.shape[0], 3, 3))
    M = 2e-3 * (A @ A.transpose(0, 2, 1) + np.eye(3))  # cell-wise anisotropic conductivity
    indptr, indices, mass, stiff = fem_assemble_p1(3, pts.shape[0], cells, pts, M)
    return dict(indptr=indptr, indices=indices, mass=mass, stiff=stiff, n=pts.shape[0], pts=pts)


@pytest.mark.parametrize("ksp,stream", [(0, False), (1, False), (0, True), (1, True)])
def test_unstructured_mesh_with_wide_rows(ctx_factory, monkeypatch, ksp, stream):
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla

    s = _delaunay_problem()
    n = s["n"]
    assert np.diff(s["indptr"]).max() > 20  # wider than one batch of the row kernels
    if stream:
        monkeypatch.setenv("MONO_PDE_STREAM", "1")
    else:
        monkeypatch.delenv("MONO_PDE_STREAM", raising=False)
    C_m, theta, dt, amp = 0.01, 0.5, 0.02, 0.4
    rng = np.random.default_rng(8)
    v_prev = -85.0 + 100.0 * np.exp(-((s["pts"][:, 0] - 1.0) ** 2) / 0.3) + rng.random(n)
    stim_idx = np.nonzero(s["pts"][:, 0] < 0.4)[0].astype(np.int32)
    stim_val = np.full(stim_idx.size, 3e-4)
    ctx = ctx_factory()
    ctx.pde_set_matrices(n, 0, s["indptr"], s["indices"], s["mass"], s["stiff"])
    ctx.pde_config(C_m, theta, 1e-11, 1e-50, 2000, 1, 0, 0)  # (condition number 26: ~65 iterations; error <= 26 * 1e-11)
    ctx.pde_set_ksp_type(ksp)
    ctx.stim_add(stim_idx, stim_val, 0.0, 1.0, amp)
    ctx.set_v_prev(v_prev)
    ctx.pde_step(0.0, dt)
    x = ctx.get_v(np.empty(n))
    its, rnorm, reason = ctx.ksp_info()
    ctx.close()
    assert reason > 0 and its >= 3, (its, rnorm, reason)
    Mm = sp.csr_matrix((s["mass"], s["indices"], s["indptr"]), shape=(n, n))
    K = sp.csr_matrix((s["stiff"], s["indices"], s["indptr"]), shape=(n, n))
    source = np.zeros(n)
    source[stim_idx] = amp * stim_val
    b = (C_m * Mm - (1 - theta) * dt * K) @ v_prev + dt * source
    want = sla.spsolve((C_m * Mm + theta * dt * K).tocsc(), b)
    assert np.abs(x - want).max() <= 1e-8 * np.abs(want).max()
