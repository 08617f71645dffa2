"""GENERATED host-side description of the cell model 'ToRORd_dynCl_endo' (see codegen/generate.py).

The step functions are *device handles*: they name a CUDA kernel, they are not callable on the CPU.
"""
import math

import numpy as np

from ..device_model import DeviceODE

MODEL_ID = 2
MODEL_TAG = 'torord'
state = {'C1': 0, 'C2': 1, 'C3': 2, 'I_': 3, 'O_': 4, 'CaMKt': 5, 'Jrel_np': 6, 'Jrel_p': 7, 'a': 8, 'ap': 9, 'iF': 10, 'iFp': 11, 'iS': 12, 'iSp': 13, 'cai': 14, 'cajsr': 15, 'cansr': 16, 'cass': 17, 'cli': 18, 'clss': 19, 'ki': 20, 'kss': 21, 'nai': 22, 'nass': 23, 'd': 24, 'fcaf': 25, 'fcafp': 26, 'fcas': 27, 'ff_': 28, 'ffp': 29, 'fs': 30, 'jca': 31, 'nca_i': 32, 'nca_ss': 33, 'h': 34, 'hp': 35, 'j': 36, 'jp': 37, 'm': 38, 'hL': 39, 'hLp': 40, 'mL': 41, 'v': 42, 'xs1': 43, 'xs2': 44}
parameter = {'A_atp': 0, 'K_atp': 1, 'K_o_n': 2, 'fkatp': 3, 'gkatp': 4, 'Aff': 5, 'ICaL_fractionSS': 6, 'Kmn': 7, 'PCa_b': 8, 'dielConstant': 9, 'k2n': 10, 'offset': 11, 'tjca': 12, 'vShift': 13, 'BSLmax': 14, 'BSRmax': 15, 'KmBSL': 16, 'KmBSR': 17, 'cmdnmax_b': 18, 'csqnmax': 19, 'kmcmdn': 20, 'kmcsqn': 21, 'kmtrpn': 22, 'trpnmax': 23, 'CaMKo': 24, 'KmCaM': 25, 'KmCaMK': 26, 'aCaMK': 27, 'bCaMK': 28, 'EKshift': 29, 'Gto_b': 30, 'F': 31, 'R': 32, 'T': 33, 'zca': 34, 'zcl': 35, 'zk': 36, 'zna': 37, 'Fjunc': 38, 'GClCa': 39, 'GClb': 40, 'KdClCa': 41, 'GK1_b': 42, 'GKb_b': 43, 'GKr_b': 44, 'alpha_1': 45, 'beta_1': 46, 'GKs_b': 47, 'GNa': 48, 'GNaL_b': 49, 'thL': 50, 'Gncx_b': 51, 'INaCa_fractionSS': 52, 'KmCaAct': 53, 'kasymm': 54, 'kcaoff': 55, 'kcaon': 56, 'kna1': 57, 'kna2': 58, 'kna3': 59, 'qca': 60, 'qna': 61, 'wca': 62, 'wna': 63, 'wnaca': 64, 'GpCa': 65, 'KmCap': 66, 'H': 67, 'Khp': 68, 'Kki': 69, 'Kko': 70, 'Kmgatp': 71, 'Knai0': 72, 'Knao0': 73, 'Knap': 74, 'Kxkur': 75, 'MgADP': 76, 'MgATP': 77, 'Pnak_b': 78, 'delta': 79, 'eP': 80, 'k1m': 81, 'k1p': 82, 'k2m': 83, 'k2p': 84, 'k3m': 85, 'k3p': 86, 'k4m': 87, 'k4p': 88, 'Jrel_b': 89, 'bt': 90, 'cajsr_half': 91, 'Jup_b': 92, 'L': 93, 'rad_': 94, 'PCab': 95, 'PKNa': 96, 'PNab': 97, 'cao': 98, 'clo': 99, 'ko': 100, 'nao': 101, 'celltype': 102, 'i_Stim_Amplitude': 103, 'i_Stim_End': 104, 'i_Stim_Period': 105, 'i_Stim_PulseDuration': 106, 'i_Stim_Start': 107, 'tauCa': 108, 'tauCl': 109, 'tauK': 110, 'tauNa': 111}
_state_defaults = [0.9982511, 0.000793602, 0.0006532143, 9.804083e-06, 0.0002922449, 0.01095026, 1.808248e-22, 4.358608e-21, 0.0008899259, 0.0004534165, 0.9996716, 0.9996716, 0.5988908, 0.6620692, 7.453481e-05, 1.525693, 1.528001, 6.497341e-05, 29.20698, 29.20696, 147.7115, 147.7114, 12.39736, 12.3977, 1.588841e-31, 1.0, 1.0, 0.9999014, 1.0, 1.0, 0.9401791, 0.9999846, 0.0008326009, 0.0004899378, 0.8473267, 0.7018454, 0.8471657, 0.8469014, 0.0006517154, 0.5566017, 0.3115491, 0.0001351203, -89.74808, 0.243959, 0.0001586167]
_parameter_defaults = [2.0, 0.25, 5.0, 0.0, 4.3195, 0.6, 0.8, 0.002, 8.3757e-05, 74.0, 500.0, 0.0, 72.5, 0.0, 1.124, 0.047, 0.0087, 0.00087, 0.05, 10.0, 0.00238, 0.8, 0.0005, 0.07, 0.05, 0.0015, 0.15, 0.05, 0.00068, 0.0, 0.16, 96485.0, 8314.0, 310.0, 2.0, -1.0, 1.0, 1.0, 1.0, 0.2843, 0.00198, 0.1, 0.6992, 0.0189, 0.0321, 0.154375, 0.1911, 0.0011, 11.7802, 0.0279, 200.0, 0.0034, 0.35, 0.00015, 12.5, 5000.0, 1500000.0, 15.0, 5.0, 88.12, 0.167, 0.5224, 60000.0, 60000.0, 5000.0, 0.0005, 0.0005, 1e-07, 1.698e-07, 0.5, 0.3582, 1.698e-07, 9.073, 27.78, 224.0, 292.0, 0.05, 9.8, 15.4509, -0.155, 4.2, 182.4, 949.5, 39.4, 687.2, 79300.0, 1899.0, 40.0, 639.0, 1.5378, 4.75, 1.7, 1.0, 0.01, 0.0011, 5.9194e-08, 0.01833, 1.9239e-09, 1.8, 150.0, 5.0, 140.0, 0.0, -53.0, 1e+17, 1000.0, 1.0, 0.0, 0.2, 2.0, 2.0, 2.0]


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
    _t0 = (6.28 * p[94])
    v_Ageo = ((p[93] * _t0) + (p[94] * _t0))
    v_Acap = (2.0 * v_Ageo)
    v_vcell = (p[93] * (p[94] * (3140.0 * p[94])))
    v_vjsr = (0.0048 * v_vcell)
    v_vmyo = (0.68 * v_vcell)
    v_vnsr = (0.0552 * v_vcell)
    v_vss = (0.02 * v_vcell)
    v_Afs = (1.0 - p[5])
    _t1 = (p[102] == 1.0)
    _t2 = (p[102] == 2.0)
    v_PCa = ((1.2 * p[8]) if _t1 else ((2.0 * p[8]) if _t2 else p[8]))
    v_PCaK = (0.0003574 * v_PCa)
    v_PCap = (1.1 * v_PCa)
    v_PCaKp = (0.0003574 * v_PCap)
    v_constA = (1820000.0 / math.pow((p[33] * p[9]), 1.5))
    v_Io = ((0.5 * ((4.0 * p[98]) + (p[99] + (p[100] + p[101])))) / 1000.0)
    _t3 = (-v_constA)
    _t4 = math.sqrt(v_Io)
    _t5 = ((_t4 / (_t4 + 1.0)) - (0.3 * v_Io))
    v_gamma_ko = math.exp((_t3 * _t5))
    _t6 = (_t3 * 4.0)
    v_gamma_cao = math.exp((_t6 * _t5))
    v_PCaNa = (0.00125 * v_PCa)
    v_PCaNap = (0.00125 * v_PCap)
    v_gamma_nao = math.exp((_t3 * _t5))
    v_Gto = ((2.0 * p[30]) if (_t1 or _t2) else p[30])
    v_cmdnmax = ((1.3 * p[18]) if _t1 else p[18])
    v_a2 = p[84]
    _t7 = (1.0 + (p[77] / p[71]))
    v_a4 = (((p[77] * p[88]) / p[71]) / _t7)
    v_b1 = (p[76] * p[81])
    v_Pnak = ((0.9 * p[78]) if _t1 else ((0.7 * p[78]) if _t2 else p[78]))
    v_k2_i = p[55]
    v_k5_i = p[55]
    v_h10_i = (((p[101] / p[57]) * (1.0 + (p[101] / p[58]))) + (p[54] + 1.0))
    v_h12_i = (1.0 / v_h10_i)
    v_k1_i = (p[56] * (p[98] * v_h12_i))
    _t8 = (p[101] * p[101])
    v_h11_i = (_t8 / (p[58] * (v_h10_i * p[57])))
    v_k2_ss = p[55]
    v_k5_ss = p[55]
    v_h10_ss = (((p[101] / p[57]) * (1.0 + (p[101] / p[58]))) + (p[54] + 1.0))
    v_h12_ss = (1.0 / v_h10_ss)
    v_k1_ss = (p[56] * (p[98] * v_h12_ss))
    v_h11_ss = (_t8 / (p[58] * (v_h10_ss * p[57])))
    v_Gncx = ((1.1 * p[51]) if _t1 else ((1.4 * p[51]) if _t2 else p[51]))
    v_GK1 = ((1.2 * p[42]) if _t1 else ((1.3 * p[42]) if _t2 else p[42]))
    v_GKb = ((0.6 * p[43]) if _t1 else p[43])
    v_GKr = ((1.3 * p[44]) if _t1 else ((0.8 * p[44]) if _t2 else p[44]))
    v_GKs = ((1.4 * p[47]) if _t1 else p[47])
    v_GNaL = ((0.6 * p[49]) if _t1 else p[49])
    v_thLp = (3.0 * p[50])
    v_akik = math.pow((p[100] / p[2]), 0.24)
    v_bkik = (1.0 / (_ipow((p[0] / p[1]), 2) + 1.0))
    v_upScale = (1.3 if _t1 else 1.0)
    v_a_rel = (0.5 * p[90])
    v_btp = (1.25 * p[90])
    v_a_relp = (0.5 * v_btp)
    v__u0 = _t3
    v__u1 = (p[32] * p[33])
    v__u2 = ((-v_gamma_ko) * p[100])
    v__u3 = (1.0 - p[6])
    v__u4 = _t6
    v__u5 = ((-p[98]) * v_gamma_cao)
    v__u6 = ((-v_gamma_nao) * p[101])
    v__u7 = (p[11] + 0.6)
    _t9 = (p[32] * p[33])
    v__u8 = (_t9 / (p[31] * p[36]))
    v__u9 = (p[22] * p[23])
    v__u10 = (v_cmdnmax * p[20])
    v__u11 = (p[19] * p[21])
    v__u12 = (p[14] * p[16])
    v__u13 = (p[15] * p[17])
    v__u14 = (1.0 - p[79])
    _t10 = (p[100] / p[70])
    v__u15 = _ipow((1.0 + _t10), 2)
    v__u16 = ((p[67] / p[68]) + 1.0)
    v__u17 = _t7
    v__u18 = (p[86] * _ipow(_t10, 2))
    v__u19 = (p[101] / p[59])
    v__u20 = (v_Gncx * (1.0 - p[52]))
    v__u21 = (v_Gncx * p[52])
    v__u22 = (_t9 / (p[31] * p[35]))
    v__u23 = ((p[96] * p[101]) + p[100])
    v__u24 = (_t9 / (p[31] * p[37]))
    _t11 = math.sqrt((p[100] / 5.0))
    v__u25 = (v_GK1 * _t11)
    v__u26 = (v_GKr * _t11)
    v__u27 = (p[95] * 4.0)
    v__u28 = (p[38] * p[39])
    v__u29 = (p[39] * (1.0 - p[38]))
    v__u30 = (v_bkik * (v_akik * (p[3] * p[4])))
    v__u31 = (-p[105])
    v__u32 = (v_upScale * 0.005425)
    v__u33 = ((v_upScale * 2.75) * 0.005425)
    v__u34 = (-v_a_rel)
    v__u35 = (-v_a_relp)
    _t12 = (2.0 * p[31])
    v__u36 = (_t12 * v_vmyo)
    v__u37 = (_t12 * v_vss)
    v__u38 = (p[31] * v_vmyo)
    v__u39 = (p[31] * v_vss)
    v__r0 = (1.0 / v__u1)
    v__r1 = (1.0 / p[69])
    v__r2 = (1.0 / p[74])
    v__r3 = (1.0 / p[75])
    v__r4 = (1.0 / v__u17)
    v__r5 = (1.0 / p[59])
    v__r6 = (1.0 / p[57])
    v__r7 = (1.0 / p[58])
    v__r8 = (1.0 / p[105])
    v__r9 = (1.0 / p[108])
    v__r10 = (1.0 / p[111])
    v__r11 = (1.0 / p[110])
    v__r12 = (1.0 / v__u36)
    v__r13 = (1.0 / v_vmyo)
    v__r14 = (1.0 / v_vnsr)
    v__r15 = (1.0 / v__u37)
    v__r16 = (1.0 / v_vss)
    v__r17 = (1.0 / v__u38)
    v__r18 = (1.0 / v__u39)
    v__r19 = (1.0 / p[12])
    v__r20 = (1.0 / p[50])
    v__r21 = (1.0 / v_thLp)
    return np.array([v_Ageo, v_Acap, v_vcell, v_vjsr, v_vmyo, v_vnsr, v_vss, v_Afs, v_PCa, v_PCaK, v_PCap, v_PCaKp, v_constA, v_Io, v_gamma_ko, v_gamma_cao, v_PCaNa, v_PCaNap, v_gamma_nao, v_Gto, v_cmdnmax, v_a2, v_a4, v_b1, v_Pnak, v_k2_i, v_k5_i, v_h10_i, v_h12_i, v_k1_i, v_h11_i, v_k2_ss, v_k5_ss, v_h10_ss, v_h12_ss, v_k1_ss, v_h11_ss, v_Gncx, v_GK1, v_GKb, v_GKr, v_GKs, v_GNaL, v_thLp, v_akik, v_bkik, v_upScale, v_a_rel, v_btp, v_a_relp, v__u0, v__u1, v__u2, v__u3, v__u4, v__u5, v__u6, v__u7, v__u8, v__u9, v__u10, v__u11, v__u12, v__u13, v__u14, v__u15, v__u16, v__u17, v__u18, v__u19, v__u20, v__u21, v__u22, v__u23, v__u24, v__u25, v__u26, v__u27, v__u28, v__u29, v__u30, v__u31, v__u32, v__u33, v__u34, v__u35, v__u36, v__u37, v__u38, v__u39, v__r0, v__r1, v__r2, v__r3, v__r4, v__r5, v__r6, v__r7, v__r8, v__r9, v__r10, v__r11, v__r12, v__r13, v__r14, v__r15, v__r16, v__r17, v__r18, v__r19, v__r20, v__r21], dtype=np.float64)


def _derived_grl1(p):
    """Parameter-only intermediates, evaluated once per parameter set and passed to the kernel."""
    _t0 = (6.28 * p[94])
    v_Ageo = ((p[93] * _t0) + (p[94] * _t0))
    v_Acap = (2.0 * v_Ageo)
    v_vcell = (p[93] * (p[94] * (3140.0 * p[94])))
    v_vjsr = (0.0048 * v_vcell)
    v_vmyo = (0.68 * v_vcell)
    v_vnsr = (0.0552 * v_vcell)
    v_vss = (0.02 * v_vcell)
    v_Afs = (1.0 - p[5])
    _t1 = (p[102] == 1.0)
    _t2 = (p[102] == 2.0)
    v_PCa = ((1.2 * p[8]) if _t1 else ((2.0 * p[8]) if _t2 else p[8]))
    v_PCaK = (0.0003574 * v_PCa)
    v_PCap = (1.1 * v_PCa)
    v_PCaKp = (0.0003574 * v_PCap)
    v_constA = (1820000.0 / math.pow((p[33] * p[9]), 1.5))
    v_Io = ((0.5 * ((4.0 * p[98]) + (p[99] + (p[100] + p[101])))) / 1000.0)
    _t3 = (-v_constA)
    _t4 = math.sqrt(v_Io)
    _t5 = ((_t4 / (_t4 + 1.0)) - (0.3 * v_Io))
    v_gamma_ko = math.exp((_t3 * _t5))
    _t6 = (p[32] * p[33])
    v_dvffrt_dv = ((p[31] * p[31]) / _t6)
    v_dvfrt_dv = (p[31] / _t6)
    _t7 = (_t3 * 4.0)
    v_gamma_cao = math.exp((_t7 * _t5))
    v_PCaNa = (0.00125 * v_PCa)
    v_PCaNap = (0.00125 * v_PCap)
    v_gamma_nao = math.exp((_t3 * _t5))
    v_Gto = ((2.0 * p[30]) if (_t1 or _t2) else p[30])
    v_cmdnmax = ((1.3 * p[18]) if _t1 else p[18])
    v_a2 = p[84]
    _t8 = (1.0 + (p[77] / p[71]))
    v_a4 = (((p[77] * p[88]) / p[71]) / _t8)
    v_b1 = (p[76] * p[81])
    v_Pnak = ((0.9 * p[78]) if _t1 else ((0.7 * p[78]) if _t2 else p[78]))
    v_k2_i = p[55]
    v_k5_i = p[55]
    v_h10_i = (((p[101] / p[57]) * (1.0 + (p[101] / p[58]))) + (p[54] + 1.0))
    v_h12_i = (1.0 / v_h10_i)
    v_k1_i = (p[56] * (p[98] * v_h12_i))
    _t9 = (p[101] * p[101])
    v_h11_i = (_t9 / (p[58] * (v_h10_i * p[57])))
    v_k2_ss = p[55]
    v_k5_ss = p[55]
    v_h10_ss = (((p[101] / p[57]) * (1.0 + (p[101] / p[58]))) + (p[54] + 1.0))
    v_h12_ss = (1.0 / v_h10_ss)
    v_k1_ss = (p[56] * (p[98] * v_h12_ss))
    v_h11_ss = (_t9 / (p[58] * (v_h10_ss * p[57])))
    v_Gncx = ((1.1 * p[51]) if _t1 else ((1.4 * p[51]) if _t2 else p[51]))
    v_GK1 = ((1.2 * p[42]) if _t1 else ((1.3 * p[42]) if _t2 else p[42]))
    v_GKb = ((0.6 * p[43]) if _t1 else p[43])
    v_GKr = ((1.3 * p[44]) if _t1 else ((0.8 * p[44]) if _t2 else p[44]))
    v_GKs = ((1.4 * p[47]) if _t1 else p[47])
    v_GNaL = ((0.6 * p[49]) if _t1 else p[49])
    v_thLp = (3.0 * p[50])
    v_dIClb_dv = p[40]
    v_akik = math.pow((p[100] / p[2]), 0.24)
    v_bkik = (1.0 / (_ipow((p[0] / p[1]), 2) + 1.0))
    v_dI_katp_I_katp_dv = (v_bkik * (v_akik * (p[3] * p[4])))
    v_dJdiff_dcai = ((-1.0) / p[108])
    v_dJdiff_dcass = (1.0 / p[108])
    v_dJdiffCl_dcli = ((-1.0) / p[111])
    v_dJdiffCl_dclss = (1.0 / p[111])
    v_dJdiffK_dki = ((-1.0) / p[110])
    v_dJdiffK_dkss = (1.0 / p[110])
    v_dJdiffNa_dnai = ((-1.0) / p[111])
    v_dJdiffNa_dnass = (1.0 / p[111])
    v_upScale = (1.3 if _t1 else 1.0)
    v_dJup_dcansr = (p[92] * (-0.0003255))
    v_a_rel = (0.5 * p[90])
    v_btp = (1.25 * p[90])
    v_a_relp = (0.5 * v_btp)
    v_dcansr_dt_linearized = (v_dJup_dcansr - ((0.016666666666666666 * v_vjsr) / v_vnsr))
    v_djca_dt_linearized = ((-1.0) / p[12])
    v_dhL_dt_linearized = ((-1.0) / p[50])
    v_dhLp_dt_linearized = ((-1.0) / v_thLp)
    v__u0 = _t3
    v__u1 = _t6
    v__u2 = ((-v_gamma_ko) * p[100])
    v__u3 = (-p[24])
    v__u4 = (1.0 - p[6])
    v__u5 = _t7
    v__u6 = ((-p[98]) * v_gamma_cao)
    v__u7 = (4.0 * v_dvffrt_dv)
    v__u8 = (2.0 * v_dvfrt_dv)
    v__u9 = ((-v_gamma_nao) * p[101])
    v__u10 = (p[11] + 0.6)
    v__u11 = (_t6 / (p[31] * p[36]))
    v__u12 = (p[22] * p[23])
    v__u13 = (v_cmdnmax * p[20])
    v__u14 = (p[19] * p[21])
    v__u15 = (p[14] * p[16])
    v__u16 = (p[15] * p[17])
    v__u17 = ((p[79] * v_dvfrt_dv) / 3.0)
    v__u18 = (1.0 / p[69])
    v__u19 = (1.0 - p[79])
    _t10 = (1.0 - p[79])
    v__u20 = ((v_dvfrt_dv * _t10) / 3.0)
    _t11 = (p[100] / p[70])
    v__u21 = _ipow((1.0 + _t11), 2)
    v__u22 = ((p[67] / p[68]) + 1.0)
    v__u23 = (1.0 / p[75])
    v__u24 = (1.0 / p[74])
    v__u25 = _t8
    v__u26 = (p[86] * _ipow(_t11, 2))
    v__u27 = (p[61] * v_dvfrt_dv)
    v__u28 = (p[101] / p[59])
    v__u29 = (1.0 / p[59])
    v__u30 = (p[60] * v_dvfrt_dv)
    v__u31 = (1.0 / p[57])
    v__u32 = (1.0 / p[58])
    v__u33 = (v_Gncx * (1.0 - p[52]))
    v__u34 = (v_Gncx * p[52])
    v__u35 = (_t6 / (p[31] * p[35]))
    v__u36 = ((p[96] * p[101]) + p[100])
    v__u37 = (_t6 / (p[31] * p[37]))
    _t12 = math.sqrt((p[100] / 5.0))
    v__u38 = (v_GK1 * _t12)
    v__u39 = (v_GKr * _t12)
    v__u40 = (p[95] * 4.0)
    _t13 = (p[95] * 4.0)
    v__u41 = (v_dvffrt_dv * _t13)
    v__u42 = (p[38] * p[39])
    v__u43 = (p[39] * (1.0 - p[38]))
    v__u44 = (p[97] * v_dvffrt_dv)
    v__u45 = (v_bkik * (v_akik * (p[3] * p[4])))
    v__u46 = (-p[105])
    v__u47 = (v_upScale * 0.005425)
    v__u48 = ((v_upScale * 2.75) * 0.005425)
    v__u49 = (-v_a_rel)
    v__u50 = (-v_a_relp)
    v__u51 = (-p[28])
    _t14 = (2.0 * p[31])
    v__u52 = (_t14 * v_vmyo)
    v__u53 = ((v_dJdiff_dcai * v_vss) / v_vmyo)
    v__u54 = (_t14 * v_vss)
    v__u55 = (-v_dJdiff_dcass)
    v__u56 = (p[31] * v_vmyo)
    v__u57 = ((v_dJdiffCl_dcli * v_vss) / v_vmyo)
    v__u58 = (p[31] * v_vss)
    v__u59 = (-v_dJdiffCl_dclss)
    v__u60 = ((v_dJdiffK_dki * v_vss) / v_vmyo)
    v__u61 = (-v_dJdiffK_dkss)
    v__u62 = ((v_dJdiffNa_dnai * v_vss) / v_vmyo)
    v__u63 = (-v_dJdiffNa_dnass)
    v__u64 = abs(v_dcansr_dt_linearized)
    v__r0 = (1.0 / v__u1)
    v__r1 = (1.0 / p[69])
    v__r2 = (1.0 / p[74])
    v__r3 = (1.0 / p[75])
    v__r4 = (1.0 / v__u25)
    v__r5 = (1.0 / p[59])
    v__r6 = (1.0 / p[57])
    v__r7 = (1.0 / p[58])
    v__r8 = (1.0 / p[105])
    v__r9 = (1.0 / p[108])
    v__r10 = (1.0 / p[111])
    v__r11 = (1.0 / p[110])
    v__r12 = (1.0 / v__u52)
    v__r13 = (1.0 / v_vmyo)
    v__r14 = (1.0 / v_vnsr)
    v__r15 = (1.0 / v__u54)
    v__r16 = (1.0 / v_vss)
    v__r17 = (1.0 / v__u56)
    v__r18 = (1.0 / v__u58)
    v__r19 = (1.0 / p[12])
    v__r20 = (1.0 / p[50])
    v__r21 = (1.0 / v_thLp)
    v__r22 = (1.0 / v_dcansr_dt_linearized)
    v__r23 = (1.0 / v_djca_dt_linearized)
    v__r24 = (1.0 / v_dhL_dt_linearized)
    v__r25 = (1.0 / v_dhLp_dt_linearized)
    return np.array([v_Ageo, v_Acap, v_vcell, v_vjsr, v_vmyo, v_vnsr, v_vss, v_Afs, v_PCa, v_PCaK, v_PCap, v_PCaKp, v_constA, v_Io, v_gamma_ko, v_dvffrt_dv, v_dvfrt_dv, v_gamma_cao, v_PCaNa, v_PCaNap, v_gamma_nao, v_Gto, v_cmdnmax, v_a2, v_a4, v_b1, v_Pnak, v_k2_i, v_k5_i, v_h10_i, v_h12_i, v_k1_i, v_h11_i, v_k2_ss, v_k5_ss, v_h10_ss, v_h12_ss, v_k1_ss, v_h11_ss, v_Gncx, v_GK1, v_GKb, v_GKr, v_GKs, v_GNaL, v_thLp, v_dIClb_dv, v_akik, v_bkik, v_dI_katp_I_katp_dv, v_dJdiff_dcai, v_dJdiff_dcass, v_dJdiffCl_dcli, v_dJdiffCl_dclss, v_dJdiffK_dki, v_dJdiffK_dkss, v_dJdiffNa_dnai, v_dJdiffNa_dnass, v_upScale, v_dJup_dcansr, v_a_rel, v_btp, v_a_relp, v_dcansr_dt_linearized, v_djca_dt_linearized, v_dhL_dt_linearized, v_dhLp_dt_linearized, v__u0, v__u1, v__u2, v__u3, v__u4, v__u5, v__u6, v__u7, v__u8, v__u9, v__u10, v__u11, v__u12, v__u13, v__u14, v__u15, v__u16, v__u17, v__u18, v__u19, v__u20, v__u21, v__u22, v__u23, v__u24, v__u25, v__u26, v__u27, v__u28, v__u29, v__u30, v__u31, v__u32, v__u33, v__u34, v__u35, v__u36, v__u37, v__u38, v__u39, v__u40, v__u41, v__u42, v__u43, v__u44, v__u45, v__u46, v__u47, v__u48, v__u49, v__u50, v__u51, v__u52, v__u53, v__u54, v__u55, v__u56, v__u57, v__u58, v__u59, v__u60, v__u61, v__u62, v__u63, v__u64, v__r0, v__r1, v__r2, v__r3, v__r4, v__r5, v__r6, v__r7, v__r8, v__r9, v__r10, v__r11, v__r12, v__r13, v__r14, v__r15, v__r16, v__r17, v__r18, v__r19, v__r20, v__r21, v__r22, v__r23, v__r24, v__r25], dtype=np.float64)


def _ipow(x, n):
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


forward_explicit_euler = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=0, scheme='forward_explicit_euler', num_states=45, num_parameters=112, derived=_derived_fe, op_counts={'add': 404, 'mul': 620, 'div': 119, 'exp': 75, 'log': 5, 'sqrt': 2, 'pow': 1, 'floor': 1, 'abs': 0, 'cmp': 9, 'select': 11, 'neg': 77})
generalized_rush_larsen = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id=1, scheme='generalized_rush_larsen', num_states=45, num_parameters=112, derived=_derived_grl1, op_counts={'add': 817, 'mul': 1608, 'div': 167, 'exp': 120, 'log': 5, 'sqrt': 2, 'pow': 1, 'floor': 1, 'abs': 18, 'cmp': 28, 'select': 30, 'neg': 179})
