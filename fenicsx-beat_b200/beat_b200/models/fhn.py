"""GENERATED host-side description of the cell model 'fitzhugh_nagumo' (see codegen/generate.py).

The step functions are *device handles*: they name a CUDA kernel, they are not callable on the CPU.
"""
import numpy as np

from ..device_model import DeviceODE

MODEL_ID = 0
MODEL_TAG = 'fhn'
state = {'s': 0, 'v': 1}
parameter = {'c_1': 0, 'c_2': 1, 'c_3': 2, 'a': 3, 'b': 4, 'v_amp': 5, 'v_rest': 6, 'v_peak': 7, 'stim_amplitude': 8, 'stim_duration': 9, 'stim_start': 10}
_state_defaults = [0.0, -85.0]
_parameter_defaults = [0.26, 0.1, 1.0, 0.13, 0.013, 125.0, -85.0, 40.0, 100.0, 1.0, 0.0]


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
    v_v_th = ((p[5] * p[3]) + p[6])
    v__u0 = (p[10] + p[9])
    v__u1 = (p[1] / p[5])
    v__u2 = (p[0] / _ipow(p[5], 2))
    v__u3 = (-p[2])
    return np.array([v_v_th, v__u0, v__u1, v__u2, v__u3], dtype=np.float64)


def _derived_grl1(p):
    """Parameter-only intermediates, evaluated once per parameter set and passed to the kernel."""
    v_v_th = ((p[5] * p[3]) + p[6])
    _t0 = (-p[2])
    v_ds_dt_linearized = (p[4] * _t0)
    v__u0 = (p[10] + p[9])
    v__u1 = (p[1] / p[5])
    v__u2 = (p[0] / _ipow(p[5], 2))
    v__u3 = _t0
    v__u4 = abs(v_ds_dt_linearized)
    v__r0 = (1.0 / v_ds_dt_linearized)
    return np.array([v_v_th, v_ds_dt_linearized, v__u0, v__u1, v__u2, v__u3, v__u4, v__r0], dtype=np.float64)


def _ipow(x, n):
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


forward_explicit_euler = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=0, scheme='forward_explicit_euler', num_states=2, num_parameters=11, derived=_derived_fe, op_counts={'add': 8, 'mul': 9, 'div': 0, 'exp': 0, 'log': 0, 'sqrt': 0, 'pow': 0, 'floor': 0, 'abs': 0, 'cmp': 3, 'select': 1, 'neg': 2})
generalized_rush_larsen = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=1, scheme='generalized_rush_larsen', num_states=2, num_parameters=11, derived=_derived_grl1, op_counts={'add': 13, 'mul': 16, 'div': 1, 'exp': 2, 'log': 0, 'sqrt': 0, 'pow': 0, 'floor': 0, 'abs': 1, 'cmp': 5, 'select': 3, 'neg': 2})
