"""smpc_set_ee_trajectory on the GPU: the stage parameters p[0:3] = cost.traj[:, current_step + i] of the reference
(controller.py:153-156) and the tracking costs (cost_definition.py:102-288) -- engine against oracle through the C ABI, both
linearisation kernels, and the host classes (NaiveController + Tracking8EXT) end to end."""
import os

import numpy as np
import pytest

from safe_mpc_b200 import abi
from tests.common import make_problem, params_model, start_states, rollout_guess

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _close(a, b, rtol, what):
    err = np.abs(a - b).max() / max(1.0, np.abs(b).max())
    assert err <= rtol, f'{what}: {err:.2e}'


@pytest.mark.parametrize('cost,lin_form', [('ext', 'thread'), ('nls', 'thread'), ('ext', 'coop')])
def test_trajectory_enters_the_solve_like_in_the_oracle(cost, lin_form):
    from safe_mpc_b200.engine import Engine
    from oracle.oracle import Oracle
    B, N = 96, 20
    prob, params, md = make_problem('naive', cost=cost, N=N)
    x0 = start_states(B, seed=71, vel=0.3)
    xg, ug = rollout_guess(x0, N, params.dt, seed=72, scale=1.0)
    rng = np.random.default_rng(73)
    traj = np.asarray(params.ee_ref) + 0.05 * np.cumsum(rng.uniform(-1, 1, (N + 6, 3)), axis=0) / np.sqrt(N)
    old = os.environ.get('SMPC_LIN')
    os.environ['SMPC_LIN'] = lin_form
    try:
        eng, orc = Engine(prob, B, 0), Oracle(prob, B, 0)
        for e in (eng, orc):
            e.set_guess(xg, ug); e.reset_controller(); e.set_ee_trajectory(traj)
        st_g, st_o = eng.rti_solve(x0), orc.rti_solve(x0)
        np.testing.assert_array_equal(st_g, st_o)
        lin_g, lin_o = eng.get_lin(), orc.get_lin()
        assert (np.abs(lin_g - lin_o) / np.maximum(1.0, np.abs(lin_o))).max() < 1e-11
        # closed loop: the step counter moves the window along the path (controller.py:283)
        x_g, x_o = x0.copy(), x0.copy()
        for step in range(5):
            u_g, ab_g = eng.controller_step(x_g); u_o, ab_o = orc.controller_step(x_o)
            np.testing.assert_array_equal(ab_g, ab_o)
            _close(u_g, u_o, RTOL, f'u step {step}')
            x_g, _ = eng.plant_step(x_g, u_g); x_o, _ = orc.plant_step(x_o, u_o)
            _close(x_g, x_o, RTOL, f'x step {step}')
        lin_g, lin_o = eng.get_lin(), orc.get_lin()
        assert (np.abs(lin_g - lin_o) / np.maximum(1.0, np.abs(lin_o))).max() < 1e-6      # (guesses agree to 1e-6 by now)
        # back to the constant reference
        eng.set_ee_trajectory(None); eng.set_guess(xg, ug); eng.reset_controller(); eng.rti_solve(x0)
        base = Engine(prob, B, 0); base.set_guess(xg, ug); base.rti_solve(x0)
        assert np.array_equal(eng.get_lin(), base.get_lin())
        base.close(); eng.close()
    finally:
        os.environ.pop('SMPC_LIN', None) if old is None else os.environ.__setitem__('SMPC_LIN', old)


def test_linearisation_kernel_forms_agree_bitwise():
    """thread-per-stage and cooperative linearisation kernel (csrc/kernels.cu) run the same functions on the same operands"""
    from safe_mpc_b200.engine import Engine
    B, N = 200, 30
    x0 = start_states(B, seed=75, vel=0.4)
    out = []
    for form, controller in (('thread', 'receding'), ('coop', 'receding')):
        os.environ['SMPC_LIN'] = form
        try:
            prob, params, md = make_problem(controller, N=N)
            xg, ug = rollout_guess(x0, N, params.dt, seed=76, scale=1.0)
            eng = Engine(prob, B, 0)
            eng.set_guess(xg, ug)
            eng.set_state(abi.STATE_R, np.full(B, 9, dtype=np.int32))
            act = np.ones(B, dtype=np.uint8); act[3::7] = 0
            eng.rti_solve(x0, act)
            out.append(eng.get_lin()[act.astype(bool)])
            eng.close()
        finally:
            os.environ.pop('SMPC_LIN', None)
    assert np.array_equal(out[0], out[1])


def test_tracking_cost_through_the_controller_classes():
    """mpc.py:40-52 with `cost_controller = Tracking8EXT(model, Q, R)`: the controller hands the path to the engine at build time and the
    solve of step j is linearised around traj[:, j + i]"""
    from safe_mpc_b200.env_model import AdamModel
    from safe_mpc_b200.utils import get_controller
    from safe_mpc_b200.cost_definition import Tracking8EXT
    B = 32
    params, md = params_model()
    params.N = 20
    model = AdamModel(params, batch=B)
    controller = get_controller('naive', model)
    cost = Tracking8EXT(model, params.Q_weight, params.R_weight)
    cost.set_solver_cost(controller)
    controller.build_controller()
    assert params.track_traj and cost.traj.shape == (3, params.n_steps_tracking + 1 + 20)
    x0 = start_states(B, seed=77, vel=0.1)
    xg, ug = rollout_guess(x0, 20, params.dt, seed=78, scale=0.5)
    controller.setGuess(xg, ug)
    # the same solves by hand: a plain engine handle that is given the rows of the reference path explicitly
    from safe_mpc_b200.engine import Engine
    from safe_mpc_b200.problem import build_problem
    prob, keep = build_problem(params, 'naive', cost='ext', N=20, model=model.data)
    plain = Engine(prob, B, 0)
    plain.set_guess(xg, ug); plain.reset_controller()
    plain.set_ee_trajectory(np.ascontiguousarray(cost.traj.T))
    const = Engine(prob, B, 0)                                      # ... and one that keeps the constant ee_ref
    const.set_guess(xg, ug); const.reset_controller()
    x = x0.copy()
    for j in range(3):
        u, ab = controller.step(x)
        u2, ab2 = plain.controller_step(x)
        u3, _ = const.controller_step(x)
        assert np.array_equal(u, u2) and np.array_equal(ab, ab2)
        assert np.array_equal(controller.ocp_solver.get_lin(), plain.get_lin())
        assert np.abs(u - u3).max() > 1e-6                          # the path is not the constant reference
        x, _ = model.integrate(x, u)
    assert np.array_equal(controller.ocp_solver.get_state(abi.STATE_STATUS), plain.get_state(abi.STATE_STATUS))
    plain.close(); const.close()


def test_caller_stream():
    """smpc_set_stream: the handle launches on the caller's stream; same results, and back to a private stream with NULL"""
    import torch
    from safe_mpc_b200.engine import Engine
    B, N = 640, 20
    prob, params, md = make_problem('st', N=N)
    x0 = start_states(B, seed=81, vel=0.3)
    xg, ug = rollout_guess(x0, N, params.dt, seed=82, scale=1.0)
    eng = Engine(prob, B, 0)
    eng.set_guess(xg, ug); st0 = eng.rti_solve(x0); xt0, ut0 = eng.get_temp()
    own = eng.stream()
    s = torch.cuda.Stream()
    eng.set_stream(s)
    assert eng.stream() == s.cuda_stream != own
    with torch.cuda.stream(s):
        xd = torch.tensor(x0, device='cuda')
        eng.set_guess(torch.tensor(xg, device='cuda'), torch.tensor(ug, device='cuda'))
        st1 = eng.rti_solve(xd)
        xt1, ut1 = eng.get_temp(like=xd)
    s.synchronize()
    assert np.array_equal(st0, st1.cpu().numpy() if hasattr(st1, 'cpu') else st1)
    assert np.array_equal(xt0, xt1.cpu().numpy()) and np.array_equal(ut0, ut1.cpu().numpy())
    eng.set_stream(None)
    assert eng.stream() not in (0, s.cuda_stream)
    eng.set_guess(xg, ug); st2 = eng.rti_solve(x0); xt2, _ = eng.get_temp()
    assert np.array_equal(st0, st2) and np.array_equal(xt0, xt2)
    eng.close()
