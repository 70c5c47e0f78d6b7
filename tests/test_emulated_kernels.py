"""Kernel-logic tests without a GPU: the engine's device sources (safe_mpc_b200/csrc/dev_model.cuh, qp_warp.cuh)
compiled for the host by tests/emu and compared with the oracle.
This is a test harness, not a fallback: the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import Oracle
from safe_mpc_b200 import abi
from tests.common import make_problem, start_states, rollout_guess

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emu():
    so = os.path.join(HERE, 'emu', 'libsmpc_emu.so')
    srcs = [os.path.join(HERE, 'emu', 'emu.cpp')] + [os.path.join(HERE, '..', 'safe_mpc_b200', 'csrc', f) for f in ('dev_model.cuh', 'qp_warp.cuh')]
    if not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(['g++', '-O2', '-std=c++17', '-DEMU_QP', '-fPIC', '-shared', '-o', so, srcs[0]], check=True)
    return C.CDLL(so)


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_analytic_linearisation_matches_ad_oracle(emu):
    prob, params, md = make_problem('htwa')
    B, N = 6, prob.N
    o = Oracle(prob, B, 2)
    x0 = start_states(B, seed=6)
    xg, ug = rollout_guess(x0, N, params.dt, seed=3)
    xg += 0.05 * np.random.default_rng(7).uniform(-1, 1, xg.shape)
    o.set_guess(xg, ug); o.rti_solve(x0)
    lin = o.get_lin()
    n = B * (N + 1)
    k = np.tile(np.arange(N + 1), B).astype(np.int32)
    x = xg.reshape(n, 10).copy()
    u = np.concatenate([ug, np.zeros((B, 1, 5))], axis=1).reshape(n, 5).copy()
    xn = np.concatenate([xg[:, 1:], np.zeros((B, 1, 10))], axis=1).reshape(n, 10).copy()
    gate = np.ones(n, dtype=np.int32)
    nn11 = np.zeros((n, 11))
    c, g = o.nn_constraint(x)
    nn11[:, 0] = c; nn11[:, 1:] = g
    rec = np.zeros((n, abi.REC))
    emu.emu_linearize(C.byref(prob), n, _p(k), _p(x), _p(u), _p(xn), _p(gate), _p(nn11), _p(rec))
    rec = rec.reshape(B, N + 1, abi.REC)
    scale = np.maximum(1.0, np.abs(lin))
    assert (np.abs(rec - lin) / scale).max() < 1e-11


@pytest.mark.parametrize('controller,cost', [('naive', 'ext'), ('st', 'ext'), ('htwa', 'ext'), ('receding', 'ext'),
                                             ('real_receding', 'ext'), ('zerovel', 'ext'), ('backup', 'zero')])
@pytest.mark.parametrize('lazy', [0, 1])
def test_qp_kernel_source_matches_oracle(emu, controller, cost, lazy):
    N, B = 10, 3
    prob, params, md = make_problem(controller, cost=cost, N=N)
    o = Oracle(prob, B, 1)
    x0 = start_states(B, seed=11, vel=0.5)
    xg, ug = rollout_guess(x0, N, params.dt, seed=12)
    o.set_guess(xg, ug)
    rset = 3 if controller in ('receding', 'real_receding') else N
    o.set_state(abi.STATE_R, np.full(B, rset, dtype=np.int32))
    o.rti_solve(x0 + 1e-3)
    lin = o.get_lin(); dz, pi, lam, t = o.get_qp()
    for b in range(B):
        z = np.zeros((N + 1, 15)); pi_e = np.zeros((N, 10)); lam_e = np.zeros((N + 1, 44)); t_e = np.zeros((N + 1, 44))
        it = C.c_int(); st = C.c_int(); res = np.zeros(5)
        x0b = (x0[b] + 1e-3).copy()
        emu.emu_qp_solve(C.byref(prob), _p(np.ascontiguousarray(lin[b])), _p(x0b), C.c_int(rset), C.c_int(lazy), _p(z), _p(pi_e), _p(lam_e), _p(t_e),
                         C.byref(it), C.byref(st), _p(res))
        _, _, oit, ost = o.qp_info(b)
        assert (it.value, st.value) == (oit, ost)
        zz = z.copy(); zz[N, :10] = z[N, 5:15]; zz[N, 10:] = 0
        assert np.abs(zz - dz[b]).max() < 1e-9
        assert np.abs(pi_e - pi[b]).max() < 1e-8
        scale = np.maximum(1.0, np.abs(lam[b][:, :44]))
        assert (np.abs(lam_e - lam[b][:, :44]) / scale).max() < 1e-6
