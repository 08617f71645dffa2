"""GENERATED host-side description of the cell model 'simple_oscillator' (see codegen/generate.py).

The step functions are *device handles*: they name a CUDA kernel, they are not callable on the CPU.
"""
import numpy as np

from ..device_model import DeviceODE

MODEL_ID = 3
MODEL_TAG = 'simple'
state = {'v': 0, 's': 1}
parameter = {'omega': 0}
_state_defaults = [0.0, 0.0]
_parameter_defaults = [1.0]


def state_index(name: str) -> int:
    return state[name]


def parameter_index(name: str) -> int:
    return parameter[name]


def init_state_values(**values):
    out = np.array(_state_defaults, dtype=np.float64)
    for k, v in values.items():
        out[state[k]] = v
    return out


def init_parameter_values(**values):
    out = np.array(_parameter_defaults, dtype=np.float64)
    for k, v in values.items():
        out[parameter[k]] = v
    return out


def _derived_fe(p):
    """Parameter-only intermediates, evaluated once per parameter set and passed to the kernel."""
    v__u0 = (-p[0])
    return np.array([v__u0], dtype=np.float64)


def _derived_grl1(p):
    """Parameter-only intermediates, evaluated once per parameter set and passed to the kernel."""
    v__u0 = (-p[0])
    return np.array([v__u0], dtype=np.float64)


def _ipow(x, n):
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


forward_explicit_euler = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=0, scheme='forward_explicit_euler', num_states=2, num_parameters=1, derived=_derived_fe, op_counts={'add': 2, 'mul': 4, 'div': 0, 'exp': 0, 'log': 0, 'sqrt': 0, 'pow': 0, 'floor': 0, 'abs': 0, 'cmp': 0, 'select': 0, 'neg': 0})
generalized_rush_larsen = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=1, scheme='generalized_rush_larsen', num_states=2, num_parameters=1, derived=_derived_grl1, op_counts={'add': 2, 'mul': 4, 'div': 0, 'exp': 0, 'log': 0, 'sqrt': 0, 'pow': 0, 'floor': 0, 'abs': 0, 'cmp': 0, 'select': 0, 'neg': 0})
