"""The device code generator (fenicsx-beat_b200/codegen) against the independent SymPy-generated oracle models, on the
CPU: the scheduled program the CUDA emitter prints is interpreted with NumPy (tests/_ir_eval.py) and must reproduce the
oracle step to rounding, for every model and scheme, shared and per-node parameters.  Also: the committed generated
files are what the generator produces from the current sources (needs /root/reference for the .ode inputs)."""
import os
import sys

import numpy as np
import pytest

import _ir_eval
import _problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))
from codegen import generate, program  # noqa: E402

ODES = "/root/reference/odes"
needs_reference = pytest.mark.skipif(not os.path.isdir(ODES), reason="the .ode sources live under /root/reference (not on the GPU box)")


def _model(tag):
    if generate.MODELS[tag][1] is None or os.path.isdir(ODES):
        return generate.load_model(tag, ODES)[0]
    pytest.skip("needs /root/reference/odes")


@pytest.mark.parametrize("tag", ["fhn", "tp06", "torord"])
@pytest.mark.parametrize("scheme", ["forward_explicit_euler", "generalized_rush_larsen"])
def test_program_matches_oracle_model(tag, scheme):
    model = _model(tag)
    om = P.oracle_model(tag)
    assert list(model.states) == list(om.state) and list(model.parameters) == list(om.parameter)  # same index order
    prog = program.build_program(model, scheme)
    rng = np.random.default_rng(42)
    states = P.perturbed_states(om, 2000, rng, P.V_NAME[tag])
    p0 = om.init_parameter_values()
    for params in (p0, np.repeat(p0[:, None], 2000, axis=1) * (1 + 0.02 * rng.uniform(-1, 1, (len(p0), 2000)))):
        for t in (0.0, 10.5):
            got = _ir_eval.evaluate(prog, states, t, 0.01, params)
            with np.errstate(all="ignore"):
                want = getattr(om, scheme)(states, t, 0.01, params)
            scale = np.maximum(np.maximum(np.abs(want), np.abs(states)), 1e-6 * np.abs(want).max(axis=1, keepdims=True) + 1e-300)
            err = (np.abs(got - want) / scale).max()
            # forward Euler: a few ulp (reciprocal pass); Rush-Larsen: exp(x) - 1 cancellation, see tests/test_gpu_parity.py
            assert err <= (1e-12 if scheme == "forward_explicit_euler" else 2e-11), (tag, scheme, t, err)


def test_rush_larsen_uses_the_total_derivative():
    """TP06: V and the concentrations have no explicit self-dependence in their written right-hand sides; the chain rule
    through the intermediates makes all 19 states Rush-Larsen (what reproduces the published Niederer table)."""
    prog = program.build_program(_model("tp06"), "generalized_rush_larsen")
    assert prog.fe_states == [] and len(prog.rl_states) == 19
    names = {n for n, _ in prog.body}
    assert "dV_dt_linearized" in names and any(n.startswith("di_") and n.endswith("_dV") for n in names)
    opaque = program.build_program(_model("fhn"), "forward_explicit_euler")
    assert opaque.rl_states == []


@needs_reference
def test_committed_generated_files_are_current():
    assert generate.main(["--check", "--odes", ODES]) == 0
