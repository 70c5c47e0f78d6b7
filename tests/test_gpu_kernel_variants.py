"""The QP solver picks between kernel forms at run time (csrc/qp.cu): cooperative / thread-per-stage prep by the active fraction,
two-warp / one-warp factorising sweep, lane-per-problem / warp-per-problem (tail) Riccati sweeps by the active count.  Small test
batches would only ever see the tail forms, so every form is forced here through the development switches and checked against the
oracle; all forms share expressions and summation orders and must agree bit for bit."""
import os

import numpy as np
import pytest

from tests.common import make_problem, start_states, rollout_guess

pytestmark = pytest.mark.gpu

B, N = 160, 30


def _solve(controller, env):
    from safe_mpc_b200.engine import Engine
    env = {'SMPC_QP_SOLO': '0', **env}     # (a batch this small would be served by the solo kernel alone: off unless the test asks for it)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        prob, params, md = make_problem(controller, N=N)
        eng = Engine(prob, B, 0)            # the switches are read when the handle is created
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    x0 = start_states(B, seed=21)
    xg, ug = rollout_guess(x0, N, params.dt, seed=22, scale=1.0)
    eng.set_guess(xg, ug)
    st = eng.rti_solve(x0)
    xt, ut = eng.get_temp()
    from safe_mpc_b200 import abi
    it = eng.get_state(abi.STATE_QP_ITER)
    eng.close()
    return st, xt, ut, it, (prob, params, x0, xg, ug)


@pytest.mark.parametrize('controller', ['st', 'receding', 'htwa'])
def test_riccati_forms_agree_bitwise_and_with_the_oracle(controller):
    from oracle.oracle import Oracle
    tail = _solve(controller, {'SMPC_QP_TAIL': '100000'})                       # warp-per-problem sweeps from the first iteration
    bulk2 = _solve(controller, {'SMPC_QP_TAIL': '0'})                           # lane-per-problem sweeps, two-warp factorisation
    bulk1 = _solve(controller, {'SMPC_QP_TAIL': '0', 'SMPC_QP_RIC1': 'single'}) # one-warp factorisation
    for other in (bulk2, bulk1):
        assert np.array_equal(tail[0], other[0]) and np.array_equal(tail[3], other[3])
        assert np.array_equal(tail[1], other[1]) and np.array_equal(tail[2], other[2])
    prob, params, x0, xg, ug = tail[4]
    orc = Oracle(prob, B, 0)
    orc.set_guess(xg, ug)
    st_o = orc.rti_solve(x0)
    xt_o, ut_o = orc.get_temp()
    assert (st_o == tail[0]).all()
    ok = st_o == 0
    assert ok.any()
    assert np.abs(tail[1][ok] - xt_o[ok]).max() <= 1e-6 * max(1.0, np.abs(xt_o[ok]).max())
    assert np.abs(tail[2][ok] - ut_o[ok]).max() <= 1e-6 * max(1.0, np.abs(ut_o[ok]).max())


@pytest.mark.parametrize('controller', ['st', 'receding'])
def test_prep_forms_agree_bitwise(controller):
    """cooperative (TMA-staged, four warps) and thread-per-stage prep: same arithmetic per term and the same summation orders, so a
    problem's result does not depend on which form the host picked for a launch (and therefore not on the batch it is solved in)"""
    coop = _solve(controller, {'SMPC_QP_TAIL': '100000'})                  # cooperative form in every iteration
    thread = _solve(controller, {'SMPC_QP_TAIL': '0', 'SMPC_QP_PREP': 'thread'})
    mixed = _solve(controller, {})                                         # the default policy
    for other in (thread, mixed):
        for i in range(4):
            assert np.array_equal(coop[i], other[i])


@pytest.mark.parametrize('controller', ['st', 'receding'])
def test_host_loop_depth_and_compaction_agree_bitwise(controller):
    """The host queues IPM iterations ahead of the counters it has read (SMPC_QP_DEPTH) and packs the problems still iterating into
    the leading slots between iterations (qs_compact_*): neither may change a bit of any result.  The batch mixes slow and fast
    problems in every tile and is large enough (40 tiles) for the compaction to act."""
    from safe_mpc_b200.engine import Engine
    from safe_mpc_b200 import abi
    Bc, Nc = 1280, 16

    def solve(env):
        env = {'SMPC_QP_SOLO': '0', **env}
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            prob, params, md = make_problem(controller, N=Nc)
            eng = Engine(prob, Bc, 0)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        x0 = start_states(Bc, seed=23, vel=0.5)
        x0[::3, 5:] *= 4.0                                    # every third problem starts fast: many more iterations, some infeasible
        xg, ug = rollout_guess(x0, Nc, params.dt, seed=24, scale=1.0)
        act = np.ones(Bc, dtype=np.uint8); act[5::17] = 0
        eng.set_guess(xg, ug)
        st = eng.rti_solve(x0, act)
        xt, ut = eng.get_temp()
        it = eng.get_state(abi.STATE_QP_ITER)
        prof = None
        eng.close()
        return st, xt, ut, it

    base = solve({'SMPC_QP_DEPTH': '0', 'SMPC_QP_COMPACT': '0'})       # one host round trip per iteration, no compaction
    assert len(set(base[3].tolist())) >= 6
    # ... and neither may the solo kernel (one CTA per problem, whole iterations on the device) that takes over the tail of a solve
    # ... nor the form of the switch to the centering direction (one CTA per flagged problem when few are flagged, else a pass over the tiles)
    for env in ({'SMPC_QP_DEPTH': '3', 'SMPC_QP_COMPACT': '0'}, {'SMPC_QP_DEPTH': '0', 'SMPC_QP_COMPACT': '1'}, {},
                {'SMPC_QP_REDO_LIST': '0'}, {'SMPC_QP_REDO_LIST': '100000'}, {'SMPC_QP_REDO_LIST': '100000', 'SMPC_QP_COMPACT': '0'},
                {'SMPC_QP_SOLO': '384', 'SMPC_QP_SOLO_TAIL': '1'}, {'SMPC_QP_SOLO': '200', 'SMPC_QP_SOLO_TAIL': '1', 'SMPC_QP_COMPACT': '0'},
                {'SMPC_QP_SOLO': '100000'}):
        other = solve(env)
        for i in range(4):
            assert np.array_equal(base[i], other[i]), (env, i)


@pytest.mark.parametrize('controller', ['st', 'receding', 'htwa', 'naive'])
def test_solo_kernel_agrees_bitwise(controller):
    """qs_solo_kernel (one CTA per problem, all phases of every interior-point iteration in one launch): from the first iteration on
    (what a batch of this size gets by default), or taking over once at most 60 problems are left -- same bits as the multi-kernel path."""
    multi = _solve(controller, {'SMPC_QP_SOLO': '0'})
    for env in ({'SMPC_QP_SOLO': '384'}, {'SMPC_QP_SOLO': '60', 'SMPC_QP_SOLO_TAIL': '1'}, {'SMPC_QP_SOLO': '60', 'SMPC_QP_SOLO_TAIL': '1', 'SMPC_QP_TAIL': '0'}):
        solo = _solve(controller, env)
        for i in range(4):
            assert np.array_equal(multi[i], solo[i]), (env, i)
