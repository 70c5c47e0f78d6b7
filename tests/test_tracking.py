"""Stage parameters of the cost (reference controller.py:153-156: p[0:3] = cost.traj[:, current_step + i]) and the tracking costs that
make them vary (cost_definition.py:102-288).  CPU: the path generators against the reference's own functions (tests/golden/
ref_tracking.npz, made by make_ref_tracking.py), the oracle's handling of the trajectory and of the step counter, and the engine's
linearisation source (host emulation) against the oracle's AD linearisation on a varying reference."""
import ctypes as C
import os
import types

import numpy as np
import pytest

from safe_mpc_b200 import abi, cost_definition as cd
from tests.common import make_problem, params_model, start_states, rollout_guess

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'ref_tracking.npz')


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize('tag,vel_const', [('const', True), ('accel', False)])
def test_path_generators_match_the_reference_functions(tag, vel_const):
    g = np.load(GOLD)
    params, md = params_model()
    cfg = cd._tracking_settings(params)
    n = int(g[f'n_track_{tag}'])
    p = types.SimpleNamespace(N=45, dt=params.dt)
    cd._Tracking8._configure(p, cfg)
    cd._TrackingMovingCircle._configure(p, cfg)
    p.n_steps = p.n_steps_tracking = n
    p.vel_const = vel_const
    p.circle_traj_vel = float(g[f'circle_vel_{tag}'])
    eight = cd.generate_8shape_trajectory(p)
    circle = cd.generate_moving_circle_trajectory(p)
    assert eight.shape == g[f'eight_{tag}'].shape == (3, n + 1 + 45)
    # (the reference differentiates the lemniscate with sympy, here the derivative is written out: same numbers up to rounding)
    assert np.abs(eight - g[f'eight_{tag}']).max() < 1e-12
    assert np.abs(circle - g[f'circle_{tag}']).max() < 1e-12


def test_reach_costs_keep_a_constant_reference():
    params, md = params_model()
    m = types.SimpleNamespace(params=params)
    c = cd.ReachTargetEXT(m)
    assert c.traj.shape == (3, params.n_steps + 1 + params.N) and not c.tracking
    assert np.array_equal(c.traj[:, 17], np.asarray(params.ee_ref))
    t = cd.Tracking8NLS(m)
    assert t.tracking and t.kind == 'nls' and params.track_traj and params.n_steps == params.n_steps_tracking
    assert t.traj.shape == (3, params.n_steps_tracking + 1 + params.N)


@pytest.mark.parametrize('cost', ['ext', 'nls'])
def test_oracle_and_engine_source_linearise_around_the_trajectory(cost):
    from oracle.oracle import Oracle
    from tests.emu import load
    emu = load()
    prob, params, md = make_problem('naive', cost=cost, N=20)
    B, N = 5, prob.N
    x0 = start_states(B, seed=4)
    xg, ug = rollout_guess(x0, N, params.dt, seed=5)
    rng = np.random.default_rng(11)
    traj = np.asarray(params.ee_ref) + 0.1 * rng.uniform(-1, 1, (N + 8, 3))
    base = Oracle(prob, B, 2); base.set_guess(xg, ug); base.rti_solve(x0)
    same = Oracle(prob, B, 2); same.set_guess(xg, ug)
    same.set_ee_trajectory(np.tile(np.asarray(params.ee_ref), (N + 8, 1)))
    same.rti_solve(x0)
    assert np.array_equal(base.get_lin(), same.get_lin())                  # a constant path = the constant ee_ref
    o = Oracle(prob, B, 2); o.set_guess(xg, ug); o.set_ee_trajectory(traj)
    for step in range(3):
        if step:
            o.controller_step(x0)                                          # advances current_step (controller.py:283)
            o.set_guess(xg, ug)
        o.rti_solve(x0)
        lin = o.get_lin()
        n = B * (N + 1)
        k = np.tile(np.arange(N + 1), B).astype(np.int32)
        x = xg.reshape(n, 10).copy()
        u = np.concatenate([ug, np.zeros((B, 1, 5))], axis=1).reshape(n, 5).copy()
        xn = np.concatenate([xg[:, 1:], np.zeros((B, 1, 10))], axis=1).reshape(n, 10).copy()
        gate = np.ones(n, dtype=np.int32); nn11 = np.zeros((n, 11))
        ee = np.ascontiguousarray(traj[np.minimum(step + k, len(traj) - 1)])
        rec = np.zeros((n, abi.REC))
        emu.emu_linearize_ref(C.byref(prob), n, _p(k), _p(x), _p(u), _p(xn), _p(gate), _p(nn11), _p(ee), _p(rec))
        rec = rec.reshape(B, N + 1, abi.REC)
        assert (np.abs(rec - lin) / np.maximum(1.0, np.abs(lin))).max() < 1e-11
        if step == 0:
            assert np.abs(lin - base.get_lin()).max() > 1e-3               # the reference does enter the records
    o.reset_controller()                                                    # current_step = 0 again (controller.py:238)
    o.set_guess(xg, ug); o.rti_solve(x0)
    first = Oracle(prob, B, 2); first.set_guess(xg, ug); first.set_ee_trajectory(traj); first.rti_solve(x0)
    assert np.array_equal(o.get_lin(), first.get_lin())
    o.set_ee_trajectory(None); o.set_guess(xg, ug); o.rti_solve(x0)
    assert np.array_equal(o.get_lin(), base.get_lin())
