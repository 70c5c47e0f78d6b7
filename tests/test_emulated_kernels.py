"""Kernel-logic tests without a GPU: the engine's device sources (safe_mpc_b200/csrc/dev_model.cuh, qp_split.cuh)
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
    from tests.emu import load
    return load()


def _emu_old():
    so = os.path.join(HERE, 'emu', 'libsmpc_emu.so')
    srcs = [os.path.join(HERE, 'emu', 'emu.cpp')] + [os.path.join(HERE, '..', 'safe_mpc_b200', 'csrc', f) for f in ('dev_model.cuh', 'qp_split.cuh')]
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


def _emu_solve(emu, prob, lin, x0, r, act=None, order=0):
    B, N = lin.shape[0], prob.N
    z = np.zeros((B, N + 1, 15)); pi = np.zeros((B, N, 10)); lam = np.zeros((B, N + 1, abi.QP_NC)); t = np.zeros((B, N + 1, abi.QP_NC))
    xt = np.full((B, N + 1, 10), np.nan); ut = np.full((B, N, 5), np.nan)
    status = np.full(B, -1, dtype=np.int32); it = np.full(B, -1, dtype=np.int32); qst = np.full(B, -1, dtype=np.int32)
    res = np.zeros((B, 5)); redo = C.c_int()
    lin = np.ascontiguousarray(lin); x0 = np.ascontiguousarray(x0); r = np.ascontiguousarray(r, dtype=np.int32)
    emu.emu_qp_solve(C.byref(prob), C.c_int(B), _p(lin), _p(x0), _p(r), _p(act) if act is not None else None, C.c_int(order),
                     _p(z), _p(pi), _p(lam), _p(t), _p(xt), _p(ut), _p(status), _p(it), _p(qst), _p(res), C.byref(redo))
    return dict(z=z, pi=pi, lam=lam, t=t, xt=xt, ut=ut, status=status, iter=it, qp_status=qst, res=res, redo=redo.value)


@pytest.mark.parametrize('controller,cost', [('naive', 'ext'), ('st', 'ext'), ('htwa', 'ext'), ('receding', 'ext'),
                                             ('real_receding', 'ext'), ('zerovel', 'ext'), ('backup', 'zero'),
                                             ('constraint_everywhere', 'ext'), ('naive', 'nls')])
@pytest.mark.parametrize('order', [0, 1])
def test_qp_kernel_source_matches_oracle(emu, controller, cost, order):
    """The split IPM of safe_mpc_b200/csrc/qp_split.cuh (the source the CUDA kernels instantiate), run on the host in
    two different work-item orders, against the oracle's QP solution: same iteration count and status, same iterate."""
    N, B = 10, 35                      # 35 problems: two tiles, the second one padded
    prob, params, md = make_problem(controller, cost=cost, N=N)
    o = Oracle(prob, B, 2)
    x0 = start_states(B, seed=11, vel=0.5)
    xg, ug = rollout_guess(x0, N, params.dt, seed=12)
    o.set_guess(xg, ug)
    rset = 3 if controller in ('receding', 'real_receding') else N
    r = np.full(B, rset, dtype=np.int32)
    o.set_state(abi.STATE_R, r)
    st_o = o.rti_solve(x0 + 1e-3)
    lin = o.get_lin(); dz, pi, lam, t = o.get_qp()
    xt_o, ut_o = o.get_temp()
    e = _emu_solve(emu, prob, lin, x0 + 1e-3, r, order=order)
    oit = np.array([o.qp_info(b)[2] for b in range(B)]); ost = np.array([o.qp_info(b)[3] for b in range(B)])
    assert (e['iter'] == oit).all() and (e['qp_status'] == ost).all()
    assert (e['status'] == st_o).all()
    assert (np.abs(e['z'] - dz) / np.maximum(1.0, np.abs(dz))).max() < 1e-8
    assert (np.abs(e['pi'] - pi) / np.maximum(1.0, np.abs(pi))).max() < 1e-7
    assert (np.abs(e['lam'] - lam) / np.maximum(1.0, np.abs(lam))).max() < 1e-6
    assert (np.abs(e['t'] - t) / np.maximum(1.0, np.abs(t))).max() < 1e-6
    assert np.abs(e['xt'] - xt_o).max() < 1e-8 and (np.abs(e['ut'] - ut_o) / np.maximum(1.0, np.abs(ut_o))).max() < 1e-8


def test_qp_kernel_source_mask_and_infeasible(emu):
    """Masked problems keep their outputs untouched; an infeasible QP reports acados status 4 and a zero step."""
    N, B = 8, 6
    prob, params, md = make_problem('naive', N=N)
    o = Oracle(prob, B, 2)
    x0 = start_states(B, seed=3)
    xg, ug = rollout_guess(x0, N, params.dt, seed=4)
    xg[1, 3, 1] = md.x_max[1] + 2.0            # far outside the state box with x0 pinned -> inconsistent rows
    o.set_guess(xg, ug)
    act = np.ones(B, dtype=np.uint8); act[4] = 0
    st_o = o.rti_solve(x0, active=act)
    lin = o.get_lin()
    e = _emu_solve(emu, prob, lin, x0, np.full(B, N, dtype=np.int32), act=act)
    assert e['status'][4] == -1 and np.isnan(e['xt'][4]).all()
    on = act.astype(bool)
    assert (e['status'][on] == st_o[on]).all()
    xt_o, ut_o = o.get_temp()
    assert np.abs(e['xt'][on] - xt_o[on]).max() < 1e-9


@pytest.mark.parametrize('controller', ['st', 'receding', 'zerovel'])
def test_run_ahead_and_compaction_do_not_change_results(emu, controller):
    """The host loop running ahead of the counters (step<2> / red queued unconditionally, iterations queued past the end of the
    solve) and the compaction of the slots between iterations (qp_split.cuh: qs_compact_*) give bit-identical results: problems
    with very different iteration counts share tiles, so the compaction moves most of them at least once."""
    N, B = 12, 100                     # 4 tiles, two groups
    prob, params, md = make_problem(controller, N=N)
    o = Oracle(prob, B, 2)
    x0 = start_states(B, seed=21, vel=0.6)
    x0[::3, 5:] *= 4.0                 # every third problem starts fast: more iterations, some infeasible
    xg, ug = rollout_guess(x0, N, params.dt, seed=22)
    o.set_guess(xg, ug)
    r = np.full(B, 5 if controller == 'receding' else N, dtype=np.int32)
    o.set_state(abi.STATE_R, r)
    o.rti_solve(x0)
    lin = o.get_lin()
    act = np.ones(B, dtype=np.uint8); act[7::11] = 0
    emu.emu_set_options(C.c_int(0), C.c_int(0))
    base = _emu_solve(emu, prob, lin, x0, r, act=act)
    assert len(set(base['iter'][act == 1].tolist())) >= 4, 'the case should spread the iteration counts'
    for depth, compact in ((3, 0), (0, 1), (2, 1)):
        emu.emu_set_options(C.c_int(depth), C.c_int(compact))
        try:
            e = _emu_solve(emu, prob, lin, x0, r, act=act)
            nc, nm = C.c_int(), C.c_int()
            emu.emu_last_compactions(C.byref(nc), C.byref(nm))
        finally:
            emu.emu_set_options(C.c_int(0), C.c_int(0))
        assert (nc.value > 0 and nm.value > 0) == bool(compact)
        for key in ('status', 'iter', 'qp_status', 'res'):
            assert np.array_equal(e[key], base[key]), (depth, compact, key)
        on = act == 1
        assert np.array_equal(e['xt'][on], base['xt'][on]) and np.array_equal(e['ut'][on], base['ut'][on]), (depth, compact)
        assert np.isnan(e['xt'][~on]).all()                       # masked problems stay untouched


@pytest.mark.parametrize('controller,cost', [('st', 'ext'), ('htwa', 'ext'), ('naive', 'ext'), ('receding', 'ext'), ('constraint_everywhere', 'ext'),
                                             ('backup', 'zero'), ('zerovel', 'ext')])
def test_fp32_storage_flavour_against_the_fp64_oracle(controller, cost):
    """smpc_problem_t::precision = SMPC_PREC_F32 (qp_split.cuh with QS_REAL = float: stage records, search directions, condensed matrices
    and Riccati factors stored in fp32, iterate and arithmetic fp64; tolerances 1e-5 / 1e-6 and the stall test of qs_ctl), compiled for the
    host: same accept / fail status as the fp64 oracle with its fp64 tolerances, trajectories within the 1e-3 relative of BASELINE.json's
    fp32 mode, and no run-away iteration (the fp32 search direction bottoms the residuals out; the stall test ends the solve there)."""
    from tests.emu import kernel_source_oracle
    N, B = 45, 48
    prob64, params, md = make_problem(controller, cost=cost, N=N)
    prob32, _, _ = make_problem(controller, cost=cost, N=N, precision='f32')
    assert prob32.precision == 1 and prob32.qp_tol_stat == 1e-5 and prob64.qp_tol_stat == 1e-6
    x0 = start_states(B, seed=31, vel=0.5)
    xg, ug = rollout_guess(x0, N, params.dt, seed=32)
    o = Oracle(prob64, B, 0)
    e = kernel_source_oracle(prob32, B, 0, f32=True)
    for h in (o, e):
        h.set_guess(xg, ug)
        if controller == 'receding':
            h.set_state(abi.STATE_R, np.full(B, 7, dtype=np.int32))
    st_o, st_e = o.rti_solve(x0 + 1e-3), e.rti_solve(x0 + 1e-3)
    assert np.array_equal(st_o, st_e)
    ok = st_o == 0
    assert ok.any()
    (xo, uo), (xe, ue) = o.get_temp(), e.get_temp()
    assert np.abs(xo - xe)[ok].max() <= 1e-3 * max(1.0, np.abs(xo).max())
    assert np.abs(uo - ue)[ok].max() <= 1e-3 * max(1.0, np.abs(uo).max())
    it_o, it_e = o.get_state(abi.STATE_QP_ITER), e.get_state(abi.STATE_QP_ITER)
    assert it_e[ok].max() <= it_o[ok].max() + 6, (it_e.max(), it_o.max())
