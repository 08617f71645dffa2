"""GENERATED host-side description of the cell model 'tentusscher_panfilov_2006_epi_cell' (see codegen/generate.py).

The step functions are *device handles*: they name a CUDA kernel, they are not callable on the CPU.
"""
import math

import numpy as np

from ..device_model import DeviceODE

MODEL_ID = 1
MODEL_TAG = 'tp06'
state = {'Xr1': 0, 'Xr2': 1, 'Xs': 2, 'm': 3, 'h': 4, 'j': 5, 'd': 6, 'f': 7, 'f2': 8, 'fCass': 9, 's': 10, 'r': 11, 'R_prime': 12, 'Ca_i': 13, 'Ca_SR': 14, 'Ca_ss': 15, 'Na_i': 16, 'V': 17, 'K_i': 18}
parameter = {'P_kna': 0, 'g_K1': 1, 'g_Kr': 2, 'g_Ks': 3, 'g_Na': 4, 'g_bna': 5, 'g_CaL': 6, 'g_bca': 7, 'g_to': 8, 'P_NaK': 9, 'K_mk': 10, 'K_mNa': 11, 'K_NaCa': 12, 'K_sat': 13, 'alpha': 14, 'gamma': 15, 'Km_Ca': 16, 'Km_Nai': 17, 'g_pCa': 18, 'K_pCa': 19, 'g_pK': 20, 'Ca_o': 21, 'k1_prime': 22, 'k2_prime': 23, 'k3': 24, 'k4': 25, 'EC': 26, 'max_sr': 27, 'min_sr': 28, 'V_rel': 29, 'V_xfer': 30, 'K_up': 31, 'V_leak': 32, 'Vmax_up': 33, 'Buf_c': 34, 'K_buf_c': 35, 'Buf_sr': 36, 'K_buf_sr': 37, 'Buf_ss': 38, 'K_buf_ss': 39, 'V_sr': 40, 'V_ss': 41, 'Na_o': 42, 'R': 43, 'T': 44, 'F': 45, 'Cm': 46, 'V_c': 47, 'stim_start': 48, 'stim_period': 49, 'stim_duration': 50, 'stim_amplitude': 51, 'K_o': 52}
_state_defaults = [0.00621, 0.4712, 0.0095, 0.00172, 0.7444, 0.7045, 3.373e-05, 0.7888, 0.9755, 0.9953, 0.999998, 2.42e-08, 0.9073, 0.000126, 3.64, 0.00036, 8.604, -85.23, 136.89]
_parameter_defaults = [0.03, 5.405, 0.153, 0.392, 14.838, 0.00029, 0.0398, 0.000592, 0.294, 2.724, 1.0, 40.0, 1000.0, 0.1, 2.5, 0.35, 1.38, 87.5, 0.1238, 0.0005, 0.0146, 2.0, 0.15, 0.045, 0.06, 0.005, 1.5, 2.5, 1.0, 0.102, 0.0038, 0.00025, 0.00036, 0.006375, 0.2, 0.001, 10.0, 0.3, 0.4, 0.00025, 1094.0, 54.68, 140.0, 8.314, 310.0, 96.485, 185.0, 16404.0, 10.0, 1000.0, 1.0, -52.0, 5.4]


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
    _t0 = (p[43] * p[44])
    v__u0 = (_t0 / p[45])
    v__u1 = (p[52] + (p[0] * p[42]))
    v__u2 = (((0.5 * p[43]) * p[44]) / p[45])
    v__u3 = math.sqrt((p[52] / 5.4))
    _t1 = math.sqrt((p[52] / 5.4))
    v__u4 = (p[2] * _t1)
    v__u5 = _ipow(p[45], 2)
    v__u6 = _t0
    v__u7 = ((p[9] * p[52]) / (p[52] + p[10]))
    v__u8 = (p[15] - 1.0)
    v__u9 = _ipow(p[42], 3)
    _t2 = _ipow(p[42], 3)
    v__u10 = ((_ipow(p[17], 3) + _t2) * (p[16] + p[21]))
    v__u11 = _ipow(p[31], 2)
    v__u12 = (p[27] - p[28])
    v__u13 = ((2.0 * p[47]) * p[45])
    v__u14 = (p[34] * p[35])
    v__u15 = (p[36] * p[37])
    v__u16 = (p[38] * p[39])
    v__u17 = ((2.0 * p[41]) * p[45])
    v__u18 = (p[48] + p[50])
    v__u19 = (p[47] * p[45])
    v__r0 = (1.0 / v__u6)
    v__r1 = (1.0 / v__u13)
    v__r2 = (1.0 / p[47])
    v__r3 = (1.0 / v__u17)
    v__r4 = (1.0 / p[41])
    v__r5 = (1.0 / p[49])
    v__r6 = (1.0 / v__u19)
    return np.array([v__u0, v__u1, v__u2, v__u3, v__u4, v__u5, v__u6, v__u7, v__u8, v__u9, v__u10, v__u11, v__u12, v__u13, v__u14, v__u15, v__u16, v__u17, v__u18, v__u19, v__r0, v__r1, v__r2, v__r3, v__r4, v__r5, v__r6], dtype=np.float64)


def _derived_grl1(p):
    """Parameter-only intermediates, evaluated once per parameter set and passed to the kernel."""
    v_di_b_Na_dV = p[5]
    v_di_b_Ca_dV = p[7]
    v_di_leak_dCa_i = (-p[32])
    v_di_leak_dCa_SR = p[32]
    v_di_xfer_dCa_i = (-p[30])
    v_di_xfer_dCa_ss = p[30]
    _t0 = (p[43] * p[44])
    v__u0 = (_t0 / p[45])
    v__u1 = (p[52] + (p[0] * p[42]))
    v__u2 = (((0.5 * p[43]) * p[44]) / p[45])
    v__u3 = math.sqrt((p[52] / 5.4))
    _t1 = math.sqrt((p[52] / 5.4))
    v__u4 = (p[2] * _t1)
    v__u5 = _ipow(p[45], 2)
    v__u6 = _t0
    v__u7 = ((2.0 * p[45]) / _t0)
    v__u8 = ((p[9] * p[52]) / (p[52] + p[10]))
    v__u9 = (((-0.1) * p[45]) / _t0)
    v__u10 = ((-p[45]) / _t0)
    v__u11 = (p[15] - 1.0)
    v__u12 = _ipow(p[42], 3)
    _t2 = _ipow(p[42], 3)
    v__u13 = ((_ipow(p[17], 3) + _t2) * (p[16] + p[21]))
    v__u14 = ((p[15] * p[45]) / _t0)
    _t3 = (p[15] - 1.0)
    v__u15 = ((_t3 * p[45]) / _t0)
    v__u16 = _ipow(p[31], 2)
    v__u17 = (p[27] - p[28])
    v__u18 = ((2.0 * p[47]) * p[45])
    v__u19 = (p[34] * p[35])
    v__u20 = (p[36] * p[37])
    v__u21 = (p[38] * p[39])
    v__u22 = ((2.0 * p[41]) * p[45])
    v__u23 = ((v_di_xfer_dCa_ss * p[47]) / p[41])
    v__u24 = (p[48] + p[50])
    v__u25 = (p[47] * p[45])
    v__r0 = (1.0 / v__u6)
    v__r1 = (1.0 / v__u18)
    v__r2 = (1.0 / p[47])
    v__r3 = (1.0 / v__u22)
    v__r4 = (1.0 / p[41])
    v__r5 = (1.0 / p[49])
    v__r6 = (1.0 / v__u25)
    return np.array([v_di_b_Na_dV, v_di_b_Ca_dV, v_di_leak_dCa_i, v_di_leak_dCa_SR, v_di_xfer_dCa_i, v_di_xfer_dCa_ss, v__u0, v__u1, v__u2, v__u3, v__u4, v__u5, v__u6, v__u7, v__u8, v__u9, v__u10, v__u11, v__u12, v__u13, v__u14, v__u15, v__u16, v__u17, v__u18, v__u19, v__u20, v__u21, v__u22, v__u23, v__u24, v__u25, v__r0, v__r1, v__r2, v__r3, v__r4, v__r5, v__r6], dtype=np.float64)


def _ipow(x, n):
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


forward_explicit_euler = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=0, scheme='forward_explicit_euler', num_states=19, num_parameters=53, derived=_derived_fe, op_counts={'add': 177, 'mul': 199, 'div': 70, 'exp': 51, 'log': 4, 'sqrt': 1, 'pow': 0, 'floor': 1, 'abs': 0, 'cmp': 4, 'select': 5, 'neg': 11})
generalized_rush_larsen = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=1, scheme='generalized_rush_larsen', num_states=19, num_parameters=53, derived=_derived_grl1, op_counts={'add': 247, 'mul': 413, 'div': 93, 'exp': 70, 'log': 4, 'sqrt': 1, 'pow': 0, 'floor': 1, 'abs': 7, 'cmp': 11, 'select': 12, 'neg': 51})
