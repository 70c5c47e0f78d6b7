"""Tensor-core evaluation of the viability network (csrc/mlp_tc.cu, nn_precision='tf32x3') against the oracle.

The oracle accumulates in fp64 (strict mode); the tensor-core path is fp32-class like the reference's libtorch call
(safe_set.py:76-94), so the comparison carries an fp32 tolerance: 2e-5 relative to the scale of the network output /
of the gradient (measured error is ~1e-6, see profiles/)."""
import numpy as np
import pytest

from tests.common import make_problem, random_states, start_states, rollout_guess

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _engines(controller, N, B, kernel='single'):
    """kernel: 'single' = csrc/mlp_tc.cu (one CTA per 64-row tile, the default), 'pair' = csrc/mlp_tc2.cu (cta_group::2 CTA pairs)"""
    import os
    from safe_mpc_b200.engine import Engine
    from oracle.oracle import Oracle
    prob_tc, params, md = make_problem(controller, N=N, nn_precision='tf32x3')
    prob_st, _, _ = make_problem(controller, N=N)
    old = os.environ.get('SMPC_MLP_TC')
    os.environ['SMPC_MLP_TC'] = kernel                    # read when the handle is created
    try:
        eng = Engine(prob_tc, B, 0)
    finally:
        if old is None:
            os.environ.pop('SMPC_MLP_TC', None)
        else:
            os.environ['SMPC_MLP_TC'] = old
    return eng, Oracle(prob_st, B, 0), params, md


@pytest.mark.parametrize('kernel', ['single', 'pair'])
@pytest.mark.parametrize('n', [1, 63, 64, 65, 127, 129, 1000, 20000])
def test_nn_constraint_rows(n, kernel):
    eng, orc, params, md = _engines('st', 20, 64, kernel)
    x = random_states(md, n, seed=n)
    c_g, g_g = eng.nn_constraint(x)
    c_o, g_o = orc.nn_constraint(x)
    assert np.isfinite(c_g).all() and np.isfinite(g_g).all()
    sc = max(1.0, np.abs(c_o).max()); sg = max(1.0, np.abs(g_o).max())
    assert np.abs(c_g - c_o).max() <= TOL * sc, np.abs(c_g - c_o).max()
    assert np.abs(g_g - g_o).max() <= TOL * sg, np.abs(g_g - g_o).max()


@pytest.mark.parametrize('kernel', ['single', 'pair'])
@pytest.mark.parametrize('controller', ['st', 'htwa', 'receding', 'constraint_everywhere'])
def test_rti_solve_fp32_mode(controller, kernel):
    """one RTI iteration with the tensor-core network: stage records of the viability rows and the step agree with the
    strict oracle within the fp32-mode tolerance of BASELINE.json (1e-3 relative on the trajectories)"""
    from safe_mpc_b200 import abi
    B, N = 96, 20
    eng, orc, params, md = _engines(controller, N, B, kernel)
    x0 = start_states(B, seed=3)
    xg, ug = rollout_guess(x0, N, params.dt, seed=4)
    for e in (eng, orc):
        e.set_guess(xg, ug)
    st_g = eng.rti_solve(x0); st_o = orc.rti_solve(x0)
    lin_g, lin_o = eng.get_lin(), orc.get_lin()
    nn_g = lin_g[:, :, abi.REC_NN:abi.REC_NN + 11]; nn_o = lin_o[:, :, abi.REC_NN:abi.REC_NN + 11]
    assert np.abs(nn_g - nn_o).max() <= TOL * max(1.0, np.abs(nn_o[np.abs(nn_o) < 1e5]).max())
    xt_g, ut_g = eng.get_temp(); xt_o, ut_o = orc.get_temp()
    ok = (st_o == 0)
    assert (st_g == st_o).mean() > 0.98
    both = ok & (st_g == 0)
    assert both.any()
    assert np.abs(xt_g[both] - xt_o[both]).max() <= 1e-3 * max(1.0, np.abs(xt_o[both]).max())
    assert np.abs(ut_g[both] - ut_o[both]).max() <= 1e-3 * max(1.0, np.abs(ut_o[both]).max())
