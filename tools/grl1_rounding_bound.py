"""How much of the device-vs-oracle difference of one generalized Rush-Larsen step is the fp64 rounding of the FORMULA
itself?  Evaluates the NumPy oracle's step (oracle/models/*.py, the same expressions the device code is generated from) on
the 20 000-state probe of tests/test_gpu_parity.py twice - in float64 and in x87 extended precision (np.longdouble, 64-bit
mantissa: 2048 times finer) - and prints, per model, the worst per-state distance of the float64 result from the
extended one in the metric of the parity test.  CPU only; test infrastructure (uses the oracle).

    python tools/grl1_rounding_bound.py            # prints one JSON line per (model, t0)

Reading: a float64 evaluation of these formulas is only defined up to this distance (a different but equally valid
rounding of exp / division / the order of a sum moves the result by as much); a device result may differ from the
float64 oracle by about twice that and still be a correctly rounded evaluation of the same expressions."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import _problems as P  # noqa: E402


def rel(a, b, y_in):
    scale = np.maximum(np.maximum(np.abs(b), np.abs(y_in)), 1e-6 * np.abs(b).max(axis=1, keepdims=True) + 1e-300)
    return np.abs(a - b) / scale


class _NumpyKeepingDtype:
    """The generated oracle modules allocate their result with dtype=float64; for the extended-precision pass the same
    code must keep the dtype of its input."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def zeros_like(a, dtype=None):
        return np.zeros_like(a)


def main():
    assert np.finfo(np.longdouble).nmant >= 63, "needs x87 extended precision"
    for tag in ("tp06", "torord", "fhn"):
        om = importlib.import_module(f"oracle.models.{tag}")
        rng = np.random.default_rng(1234)
        states = P.perturbed_states(om, 20000, rng, P.V_NAME[tag])
        params = om.init_parameter_values()
        names = {v: k for k, v in om.state.items()}
        for t0 in (0.0, 10.5):
            with np.errstate(all="ignore"):
                y64 = om.generalized_rush_larsen(states, t0, 0.01, params)
                om.numpy = _NumpyKeepingDtype()
                yld = om.generalized_rush_larsen(states.astype(np.longdouble), np.longdouble(t0), np.longdouble(0.01), params.astype(np.longdouble))
                om.numpy = np
            assert yld.dtype == np.longdouble, yld.dtype
            e = rel(y64.astype(np.longdouble), yld, states)
            per_state = np.asarray(e.max(axis=1), dtype=np.float64)
            worst = np.argsort(-per_state)[:4]
            print(json.dumps({"model": tag, "t0": t0, "max_rel_err_of_float64_evaluation": float(per_state.max()),
                              "worst_states": {names[int(i)]: float(per_state[i]) for i in worst},
                              "states_above_1e-12": int((per_state > 1e-12).sum())}))


if __name__ == "__main__":
    main()
