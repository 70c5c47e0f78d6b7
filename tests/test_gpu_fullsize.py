"""BASELINE.json full size (configs[1]: ST controller, N = 45, batch 10 000): size-independent properties of one RTI solve.

The oracle needs minutes for 10 000 problems, so at this size the GPU result is checked through properties the domain
offers: the QP imposes x_0 and the (linear) double-integrator dynamics exactly, the state box and the linearised torque rows
hold on the full step, and the result of a problem does not depend on the batch it is solved in (the multi-GPU sharding
argument of DESIGN.md section 6: block sharding = solving sub-batches on their own)."""
import numpy as np
import pytest

from safe_mpc_b200 import abi
from tests.common import make_problem, start_states, rollout_guess

pytestmark = pytest.mark.gpu

B, N = 10000, 45


@pytest.fixture(scope='module')
def solved():
    from safe_mpc_b200.engine import Engine
    prob, params, md = make_problem('st', N=N, keep_slots=True)      # the stage records are read back below: no compaction of the slots
    x0 = start_states(B, seed=11)
    xg, ug = rollout_guess(x0, N, params.dt, seed=12, scale=1.0)
    eng = Engine(prob, B, 0)
    eng.set_guess(xg, ug)
    st = eng.rti_solve(x0)
    xt, ut = eng.get_temp()
    return dict(prob=prob, params=params, md=md, x0=x0, xg=xg, ug=ug, eng=eng, st=st, xt=xt, ut=ut)


def test_full_size_solution_properties(solved):
    s = solved
    st, xt, ut, x0, prob = s['st'], s['xt'], s['ut'], s['x0'], s['prob']
    ok = st == 0
    assert ok.mean() > 0.9, f'only {ok.mean():.3f} of the problems solved'
    assert set(np.unique(st)) <= {0, 1, 2, 3, 4}
    it = s['eng'].get_state(abi.STATE_QP_ITER)
    assert it.max() <= prob.qp_iter_max and it[ok].min() >= 1
    dt, nq = s['params'].dt, abi.NQ
    # x_0 is imposed (lbx_0 = ubx_0 = x0, controller.py:144-145) and the dynamics are linear: both hold on the full step
    assert np.abs(xt[ok, 0] - x0[ok]).max() < 1e-7
    q, v = xt[ok, :-1, :nq], xt[ok, :-1, nq:]
    qn = q + dt * v + 0.5 * dt * dt * ut[ok]
    vn = v + dt * ut[ok]
    assert np.abs(qn - xt[ok, 1:, :nq]).max() < 1e-7
    assert np.abs(vn - xt[ok, 1:, nq:]).max() < 1e-7
    # state box of the stages 1..N-1 (controller.py:49-51) and of the terminal stage
    lbx, ubx = np.array(prob.lbx), np.array(prob.ubx)
    assert (xt[ok, 1:N] >= lbx - 1e-7).all() and (xt[ok, 1:N] <= ubx + 1e-7).all()
    # linearised torque rows tau + J dz in [tau_min, tau_max] (env_model.py:263-271)
    lin = s['eng'].get_lin()[ok][:, :N]
    dz = np.concatenate([ut[ok] - s['ug'][ok], xt[ok, :-1] - s['xg'][ok, :-1]], axis=2)           # [du; dq; dv]
    tau = lin[:, :, abi.REC_TAU:abi.REC_TAU + 5] + np.einsum('bkrc,bkc->bkr', lin[:, :, abi.REC_JTAU:abi.REC_JTAU + 75].reshape(-1, N, 5, 15), dz)
    tmin, tmax = np.array(prob.tau_min), np.array(prob.tau_max)
    assert (tau >= tmin - 1e-6).all() and (tau <= tmax + 1e-6).all()
    # linearised capsule rows (squared distances stay above the OCP bound)
    dist = lin[:, 1:, abi.REC_DIST:abi.REC_DIST + 6] + np.einsum('bkrc,bkc->bkr', lin[:, 1:, abi.REC_JDIST:abi.REC_JDIST + 30].reshape(-1, N - 1, 6, 5), dz[:, 1:, 5:10])
    assert (dist >= np.array(prob.pair_lo_ocp) - 1e-6).all()


def test_result_does_not_depend_on_the_batch(solved):
    """two half batches solved on their own give bit-identical results (what the per-rank shards of bench.py --gpus N do) -- and
    they are solved with the default policy, i.e. WITH the compaction of the solver's slots that the full-batch fixture switches off"""
    from safe_mpc_b200.engine import Engine
    s = solved
    h = B // 2 + 16                                       # not a multiple of the tile or group size
    prob2, _, _ = make_problem('st', N=N)
    for lo, hi in ((0, h), (h, B)):
        eng = Engine(prob2, hi - lo, 0)
        eng.set_guess(s['xg'][lo:hi], s['ug'][lo:hi])
        st = eng.rti_solve(s['x0'][lo:hi])
        xt, ut = eng.get_temp()
        assert (st == s['st'][lo:hi]).all()
        assert np.array_equal(xt, s['xt'][lo:hi]) and np.array_equal(ut, s['ut'][lo:hi])
        eng.close()
