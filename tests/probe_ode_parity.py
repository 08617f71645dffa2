"""TEST INFRASTRUCTURE (uses the oracle): per-state single-step parity of the device GRL1 kernels against the NumPy oracle on
20 000 random states, printing the worst states.  Run from the repository root: python tests/probe_ode_parity.py [t0]"""
import sys, importlib
sys.path.insert(0, "fenicsx-beat_b200"); sys.path.insert(0, ".")
import numpy as np
import _problems as P
from beat_b200._lib import Context
for tag in ("tp06", "torord", "fhn"):
    om = P.oracle_model(tag); hm = importlib.import_module(f"beat_b200.models.{tag}")
    rng = np.random.default_rng(1234); n = 20000
    states = P.perturbed_states(om, n, rng, P.V_NAME[tag]); params = om.init_parameter_values()
    dev = hm.generalized_rush_larsen
    ctx = Context(0); ctx.ode_create(P.MODEL_ID[tag], 1, n, om.state_index(P.V_NAME[tag]), states.shape[0])
    t0 = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
    ctx.ode_set_states(states); ctx.ode_set_params(params, dev.derived(params)); ctx.ode_step(t0, 0.01)
    got = ctx.ode_get_states()
    with np.errstate(all="ignore"):
        want = om.generalized_rush_larsen(states, t0, 0.01, params)
    scale = np.maximum(np.maximum(np.abs(want), np.abs(states)), 1e-6 * np.abs(want).max(axis=1, keepdims=True) + 1e-300)
    e = np.abs(got - want) / scale
    names = {v: k for k, v in om.state.items()}
    worst = np.argsort(-e.max(axis=1))[:5]
    print(tag, "max", e.max(), [(names[i], float(e[i].max())) for i in worst])
    i, j = np.unravel_index(e.argmax(), e.shape)
    print("   worst state", names[i], "y_old", states[i, j], "y_new", want[i, j], "dev", got[i, j], "V", states[om.state_index(P.V_NAME[tag]), j])
    ctx.close()
