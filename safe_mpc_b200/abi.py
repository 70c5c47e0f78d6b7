"""ctypes mirror of include/safe_mpc_b200.h (struct smpc_problem and constants).

The same structure layout is used by the test oracle (oracle/oracle.h: orc_problem), so the host layer
fills it once and hands it to either library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

NQ = 5
NX = 2 * NQ
NU = NQ
NPAIR = 6
MAX_POINTS = 8
HID = 256
NN_NPARAM = HID * NX + HID + 2 * (HID * HID + HID) + HID + 1
REC = 192
MAX_N = 128
QP_NR = 22
QP_NC = 2 * QP_NR + 2
# stage-record offsets (include/safe_mpc_b200.h SMPC_REC_*)
REC_U, REC_X, REC_G, REC_HQQ, REC_TAU, REC_JTAU, REC_DIST, REC_JDIST, REC_NN, REC_JNN, REC_B = 0, 5, 15, 30, 45, 50, 125, 131, 161, 162, 172
REC_HU, REC_HV, REC_HQ, REC_NNROW, REC_SOFT, REC_NTAU, REC_NDIST = 182, 183, 184, 185, 186, 187, 188
OUT_CONVERGED, OUT_COLLIDED, OUT_ABORTED = 1, 2, 4

HOST, DEVICE = 0, 1

CTRL = {'naive': 0, 'zerovel': 1, 'st': 2, 'stwa': 3, 'htwa': 4, 'receding': 5, 'real_receding': 6,
        'constraint_everywhere': 7, 'backup': 8, 'parallel': 9}
NN_NONE, NN_TERMINAL, NN_RECEDING, NN_EVERYWHERE, NN_PARALLEL = 0, 1, 2, 3, 4
NN_PRECISION = {'strict': 0, 'tf32x3': 1}
PRECISION = {'f64': 0, 'f32': 1}
COST_ZERO, COST_EXT, COST_NLS = 0, 1, 2
STATE_FAILS, STATE_R, STATE_STATUS, STATE_QP_ITER, STATE_QP_STATUS = 0, 1, 2, 3, 4


class Problem(C.Structure):
    _fields_ = [
        ('nq', C.c_int32), ('N', C.c_int32), ('n_pairs', C.c_int32), ('n_points', C.c_int32),
        ('controller', C.c_int32), ('nn_rows', C.c_int32), ('nn_terminal_soft', C.c_int32),
        ('stage0_collision_rows', C.c_int32), ('cost_type', C.c_int32), ('abort_flag', C.c_int32),
        ('qp_iter_max', C.c_int32), ('lm_scale_dt', C.c_int32), ('qp_cond_pred_corr', C.c_int32),
        ('nn_precision', C.c_int32), ('qp_keep_slots', C.c_int32), ('precision', C.c_int32),
        ('dt', C.c_double), ('q_weight', C.c_double), ('r_weight', C.c_double), ('lm', C.c_double),
        ('alpha', C.c_double), ('eps', C.c_double), ('slack_penalty_e', C.c_double),
        ('tol_x', C.c_double), ('tol_tau', C.c_double), ('tol_obs', C.c_double), ('tol_safe', C.c_double),
        ('tol_conv', C.c_double),
        ('qp_mu0', C.c_double), ('qp_tol_stat', C.c_double), ('qp_tol_eq', C.c_double),
        ('qp_tol_ineq', C.c_double), ('qp_tol_comp', C.c_double), ('qp_alpha_min', C.c_double),
        ('qp_reg_prim', C.c_double),
        ('gravity', C.c_double * 3), ('qp_maxiter_accept', C.c_double), ('reserved_d', C.c_double * 7),
        ('joint_R', (C.c_double * 9) * NQ), ('joint_p', (C.c_double * 3) * NQ),
        ('joint_axis', (C.c_double * 3) * NQ), ('inertial', (C.c_double * 10) * NQ),
        ('x_min', C.c_double * NX), ('x_max', C.c_double * NX),
        ('lbx', C.c_double * NX), ('ubx', C.c_double * NX),
        ('lbx_e', C.c_double * NX), ('ubx_e', C.c_double * NX),
        ('tau_min', C.c_double * NU), ('tau_max', C.c_double * NU),
        ('ee_ref', C.c_double * 3),
        ('point_body', C.c_int32 * MAX_POINTS), ('point_local', (C.c_double * 3) * MAX_POINTS),
        ('pair_pa', C.c_int32 * NPAIR), ('pair_pb', C.c_int32 * NPAIR),
        ('pair_C', (C.c_double * 3) * NPAIR), ('pair_D', (C.c_double * 3) * NPAIR),
        ('pair_lo_ocp', C.c_double * NPAIR), ('pair_lo_chk', C.c_double * NPAIR),
        ('pair_hi', C.c_double),
        ('nn_mean', C.c_double * NQ), ('nn_std', C.c_double * NQ),
        ('nn_weights', C.POINTER(C.c_float)),
    ]


def _set(dst, src):
    """Copy a numpy array into a (possibly nested) ctypes array field."""
    base = dst._type_
    while hasattr(base, '_length_'):
        base = base._type_
    dtype = {C.c_double: np.float64, C.c_int32: np.int32, C.c_float: np.float32}[base]
    flat = np.ascontiguousarray(src, dtype=dtype).ravel()
    n = C.sizeof(dst) // flat.itemsize
    if flat.size != n:
        raise ValueError(f'size mismatch: field holds {n} values, got {flat.size}')
    C.memmove(C.addressof(dst), flat.ctypes.data, C.sizeof(dst))


def pack_nn_weights(ws, bs) -> np.ndarray:
    """W1 b1 W2 b2 W3 b3 W4 b4, row-major float32 (layout documented in the header)."""
    parts = []
    for w, b in zip(ws, bs):
        parts += [np.ascontiguousarray(w, dtype=np.float32).ravel(), np.ascontiguousarray(b, dtype=np.float32).ravel()]
    flat = np.concatenate(parts)
    if flat.size != NN_NPARAM:
        raise ValueError(f'network has {flat.size} parameters, this build expects {NN_NPARAM} '
                         f'({NX}->{HID}->{HID}->{HID}->1)')
    return flat


def fill(prob: Problem, **fields):
    for key, val in fields.items():
        cur = getattr(prob, key)
        if isinstance(cur, C.Array):
            _set(cur, val)
        else:
            setattr(prob, key, val)
    return prob
