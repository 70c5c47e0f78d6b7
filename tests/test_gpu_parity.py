"""Parity of the CUDA engine (through the C ABI) with the oracle on the same seeded inputs.
Tolerance: 1e-6 relative on state/control trajectories (BASELINE.json north_star, fp64 mode); integer outputs
(status, fails, receding index, outcome codes) must be identical."""
import numpy as np
import pytest

from safe_mpc_b200 import abi
from tests.common import make_problem, random_states, start_states, constant_guess, rollout_guess, kkt_residuals

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def _pair(controller, cost='ext', N=None, B=32, **kw):
    from safe_mpc_b200.engine import Engine
    from oracle.oracle import Oracle
    prob, params, md = make_problem(controller, cost=cost, N=N, **kw)
    return Engine(prob, B, 0), Oracle(prob, B, 0), prob, params, md


def _close(a, b, rtol=RTOL, what=''):
    a = np.asarray(a); b = np.asarray(b)
    scale = max(1.0, float(np.abs(b).max()))
    err = float(np.abs(a - b).max()) / scale
    assert err <= rtol, f'{what}: max rel err {err:.3e} > {rtol}'


def test_model_functions():
    eng, orc, prob, params, md = _pair('htwa', B=8)
    x = random_states(md, 200, seed=21, vel_scale=0.6)
    u = np.random.default_rng(22).uniform(-8, 8, (200, 5))
    _close(eng.tau(x, u), orc.tau(x, u), 1e-11, 'tau')
    ee_g, d_g = eng.kinematics(x); ee_o, d_o = orc.kinematics(x)
    _close(ee_g, ee_o, 1e-12, 'ee'); _close(d_g, d_o, 1e-12, 'dist')
    c_g, g_g = eng.nn_constraint(x); c_o, g_o = orc.nn_constraint(x)
    _close(c_g, c_o, 1e-10, 'nn value'); _close(g_g, g_o, 1e-9, 'nn gradient')


@pytest.mark.parametrize('controller,cost,N', [('naive', 'ext', 45), ('naive', 'nls', 20), ('zerovel', 'ext', 30), ('st', 'ext', 45),
                                               ('htwa', 'ext', 45), ('receding', 'ext', 25), ('real_receding', 'ext', 20),
                                               ('constraint_everywhere', 'ext', 20), ('backup', 'zero', 45)])
def test_rti_solve(controller, cost, N):
    B = 32
    eng, orc, prob, params, md = _pair(controller, cost, N, B)
    x0 = start_states(B, seed=31, vel=0.5)
    xg, ug = rollout_guess(x0, N, params.dt, seed=32)
    r0 = np.full(B, min(7, N - 1), dtype=np.int32)
    for e in (eng, orc):
        e.set_guess(xg, ug)
        if controller in ('receding', 'real_receding'):
            e.set_state(abi.STATE_R, r0)
    st_g = eng.rti_solve(x0 + 1e-3); st_o = orc.rti_solve(x0 + 1e-3)
    np.testing.assert_array_equal(st_g, st_o)
    np.testing.assert_array_equal(eng.get_state(abi.STATE_QP_STATUS), orc.get_state(abi.STATE_QP_STATUS))
    lin_g, lin_o = eng.get_lin(), orc.get_lin()
    _close(lin_g, lin_o, 1e-11, 'stage records')
    xt_g, ut_g = eng.get_temp(); xt_o, ut_o = orc.get_temp()
    _close(xt_g, xt_o, RTOL, 'x_temp'); _close(ut_g, ut_o, RTOL, 'u_temp')
    it_g, it_o = eng.get_state(abi.STATE_QP_ITER), orc.get_state(abi.STATE_QP_ITER)
    assert np.abs(it_g - it_o).max() <= 1          # same algorithm; a residual at a threshold may flip one iteration
    # the GPU solution is itself a KKT point of the QP rebuilt from its own records
    dz, pi, lam, t = eng.get_qp()
    if controller != 'real_receding':
        for b in range(3):
            if eng.get_state(abi.STATE_QP_STATUS)[b] == 0:
                r = kkt_residuals(prob, lin_g[b], x0[b] + 1e-3, dz[b], pi[b], lam[b], t[b])
                assert r['stat'] < 5e-6 and r['eq'] < 1e-7 and r['ineq'] < 1e-7 and r['comp'] < 1e-6, r


def test_rti_solve_active_mask_and_failure():
    B = 16
    eng, orc, prob, params, md = _pair('zerovel', 'ext', 4, B)     # 4 steps cannot brake 3 rad/s
    x0 = start_states(B, seed=33, vel=0.0)
    x0[::2, 5:] = 3.0
    xg, ug = constant_guess(x0, prob.N)
    act = (np.arange(B) % 4 != 1).astype(np.uint8)
    for e in (eng, orc):
        e.set_guess(xg, ug)
    st_g = eng.rti_solve(x0, act); st_o = orc.rti_solve(x0, act)
    np.testing.assert_array_equal(st_g, st_o)
    xt_g, ut_g = eng.get_temp(); xt_o, ut_o = orc.get_temp()
    _close(xt_g, xt_o, RTOL, 'x_temp'); _close(ut_g, ut_o, RTOL, 'u_temp')
    assert np.abs(ut_g[act == 0]).max() == 0.0


def _well_posed(orc, B, margin=1e-1):
    """Problems whose last QP the oracle solved with its stationarity residual at least `margin` x below the exit
    tolerance.  The IPM stops at |res_stat| <= 1e-6 while the Levenberg-Marquardt curvature is lm*dt = 2.5e-3, so a
    solve that exits with a residual near the tolerance determines the step only to ~1e-6 / 2.5e-3 = 4e-4: two
    implementations that differ in rounding agree to 1e-6 only while the residual is far below the tolerance
    (it is 1e-8 .. 1e-10 on every well-conditioned QP; it approaches 1e-6 right before a QP becomes infeasible)."""
    ok = np.zeros(B, dtype=bool)
    for b in range(B):
        res, mu, it, st = orc.qp_info(b)
        ok[b] = (st == 0) and (res[0] < margin * orc.prob.qp_tol_stat)
    return ok


@pytest.mark.parametrize('controller', ['naive', 'st', 'htwa', 'receding', 'real_receding', 'constraint_everywhere', 'zerovel'])
def test_controller_steps_closed_loop(controller):
    """controller.step + plant, 12 steps: per-step control, guess, fails / r / abort identical.  Each implementation
    runs its own closed loop; a problem is compared until its first ill-conditioned QP (see _well_posed)."""
    B, N = 24, 20
    eng, orc, prob, params, md = _pair(controller, 'ext', N, B, noise=0.0)
    x0 = start_states(B, seed=41, vel=0.3)
    xg, ug = rollout_guess(x0, N, params.dt, seed=42, scale=1.0)
    rng = np.random.default_rng(43)
    pin = np.tile(md.inertial, (B, 1, 1)) * (1 + 0.05 * rng.uniform(-1, 1, (B, 5, 10)))
    for e in (eng, orc):
        e.set_guess(xg, ug); e.reset_controller(); e.set_plant_inertial(pin)
    x_g, x_o = x0.copy(), x0.copy()
    keep = np.ones(B, dtype=bool)
    for step in range(12):
        u_g, ab_g = eng.controller_step(x_g); u_o, ab_o = orc.controller_step(x_o)
        keep &= _well_posed(orc, B)
        m = keep
        np.testing.assert_array_equal(ab_g[m], ab_o[m], err_msg=f'abort flags, step {step}')
        np.testing.assert_array_equal(eng.get_state(abi.STATE_FAILS)[m], orc.get_state(abi.STATE_FAILS)[m], err_msg=f'fails, step {step}')
        np.testing.assert_array_equal(eng.get_state(abi.STATE_R)[m], orc.get_state(abi.STATE_R)[m], err_msg=f'r, step {step}')
        _close(u_g[m], u_o[m], RTOL, f'u step {step}')
        xgg, ugg = eng.get_guess(); xgo, ugo = orc.get_guess()
        _close(xgg[m], xgo[m], RTOL, f'x_guess step {step}'); _close(ugg[m], ugo[m], RTOL, f'u_guess step {step}')
        x_g, _ = eng.plant_step(x_g, u_g); x_o, _ = orc.plant_step(x_o, u_o)
        _close(x_g[m], x_o[m], RTOL, f'x step {step}')
    print(f'\n{controller}: {int(keep.sum())} of {B} problems compared over all 12 steps ({B - int(keep.sum())} dropped at an ill-conditioned QP)')
    assert keep.sum() >= B // 2, f'only {keep.sum()} of {B} problems stayed well-posed'
    _close(eng.get_x_viable()[keep], orc.get_x_viable()[keep], RTOL, 'x_viable')


def test_plant_step_with_noise_and_saturation():
    B = 64
    eng, orc, prob, params, md = _pair('naive', B=B)
    rng = np.random.default_rng(51)
    x = random_states(md, B, seed=52)
    u = rng.uniform(-60, 60, (B, 5))
    pin = np.tile(md.inertial, (B, 1, 1)) * (1 + 0.2 * rng.uniform(-1, 1, (B, 5, 10)))
    noise = rng.normal(0, 2.0, (B, 5))
    for e in (eng, orc):
        e.set_plant_inertial(pin); e.set_torque_noise(noise)
    xn_g, a_g = eng.plant_step(x, u); xn_o, a_o = orc.plant_step(x, u)
    _close(xn_g, xn_o, 1e-10, 'x_next'); _close(a_g, a_o, 1e-9, 'applied acceleration')


@pytest.mark.parametrize('controller,noise', [('naive', 0.0), ('htwa', 0.0), ('receding', 5.0), ('st', 5.0)])
def test_closed_loop_sim(controller, noise):
    """The whole mpc.py loop (abort handling included) for 40 steps: identical outcome codes, logs within 1e-6."""
    from safe_mpc_b200.engine import Engine, Sim
    from oracle.oracle import Oracle, OracleSim
    B, N, Nb, steps = 48, 12, 12, 40
    prob, params, md = make_problem(controller, N=N, noise=noise)
    bprob, _, _ = make_problem('backup', cost='zero', N=Nb, noise=noise)
    x0 = start_states(B, seed=61, vel=0.4, spread=0.3)
    xg, ug = rollout_guess(x0, N, params.dt, seed=62, scale=1.0)
    rng = np.random.default_rng(63)
    pin = np.tile(md.inertial, (B, 1, 1)) * (1 + noise / 100 * rng.uniform(-1, 1, (B, 5, 10)))
    tn = rng.normal(0, 0.3, (B, 5))
    res = []
    for E, S in ((Engine, Sim), (Oracle, OracleSim)):
        main, bk = E(prob, B, 0), E(bprob, B, 0)
        main.set_guess(xg, ug); main.reset_controller(); main.set_plant_inertial(pin); main.set_torque_noise(tn)
        sim = S(main, bk, steps)
        sim.reset(x0)
        sim.run(steps)
        res.append((sim.outcome(), sim.log(), sim.counters(), sim.x_viable()))
    (o_g, (x_g, u_g), c_g, xv_g), (o_o, (x_o, u_o), c_o, xv_o) = res
    np.testing.assert_array_equal(o_g, o_o)
    np.testing.assert_array_equal(np.isnan(x_g), np.isnan(x_o))
    np.testing.assert_array_equal(np.isnan(u_g), np.isnan(u_o))
    _close(np.nan_to_num(x_g), np.nan_to_num(x_o), RTOL, 'x log')
    _close(np.nan_to_num(u_g), np.nan_to_num(u_o), 1e-5, 'u log')
    for key in ('rti_solves', 'backup_solves', 'plant_steps'):
        assert c_g[key] == c_o[key], (key, c_g, c_o)
    np.testing.assert_array_equal(np.isnan(xv_g), np.isnan(xv_o))


def test_device_pointer_path_matches_host_path():
    import torch
    B, N = 16, 15
    eng, orc, prob, params, md = _pair('htwa', 'ext', N, B)
    x0 = start_states(B, seed=71)
    xg, ug = rollout_guess(x0, N, params.dt, seed=72)
    eng.set_guess(torch.tensor(xg, device='cuda'), torch.tensor(ug, device='cuda'))
    u_d, ab_d = eng.controller_step(torch.tensor(x0, device='cuda'))
    eng.sync()
    orc.set_guess(xg, ug)
    u_o, ab_o = orc.controller_step(x0)
    assert u_d.is_cuda
    _close(u_d.cpu().numpy(), u_o, RTOL, 'u (device path)')
    np.testing.assert_array_equal(ab_d.cpu().numpy(), ab_o)


@pytest.mark.parametrize('alpha', [10.0, 60.0])
def test_parallel_controller_steps(alpha):
    """ParallelController.step (controller.py:614-640): N solves per step, one per candidate node; the node kept, the receding index,
    fails / abort and the control agree with the oracle.  alpha = 60 tightens the viability row so that candidates fail."""
    B, N = 20, 8
    eng, orc, prob, params, md = _pair('parallel', 'ext', N, B, alpha=alpha)
    x0 = start_states(B, seed=51, vel=0.6)
    xg, ug = rollout_guess(x0, N, params.dt, seed=52, scale=1.0)
    for e in (eng, orc):
        e.set_guess(xg, ug); e.reset_controller()
    x_g, x_o = x0.copy(), x0.copy()
    keep = np.ones(B, dtype=bool)
    seen_r = set()
    for step in range(8):
        u_g, ab_g = eng.controller_step(x_g); u_o, ab_o = orc.controller_step(x_o)
        keep &= _well_posed(orc, B, margin=0.05)
        m = keep
        np.testing.assert_array_equal(ab_g[m], ab_o[m], err_msg=f'abort flags, step {step}')
        np.testing.assert_array_equal(eng.get_state(abi.STATE_FAILS)[m], orc.get_state(abi.STATE_FAILS)[m], err_msg=f'fails, step {step}')
        np.testing.assert_array_equal(eng.get_state(abi.STATE_R)[m], orc.get_state(abi.STATE_R)[m], err_msg=f'r, step {step}')
        seen_r |= set(orc.get_state(abi.STATE_R)[m].tolist())
        _close(u_g[m], u_o[m], RTOL, f'u step {step}')
        xgg, ugg = eng.get_guess(); xgo, ugo = orc.get_guess()
        _close(xgg[m], xgo[m], RTOL, f'x_guess step {step}'); _close(ugg[m], ugo[m], RTOL, f'u_guess step {step}')
        x_g, _ = eng.plant_step(x_g, u_g); x_o, _ = orc.plant_step(x_o, u_o)
        _close(x_g[m], x_o[m], RTOL, f'x step {step}')
    assert keep.sum() >= B // 2, f'only {keep.sum()} of {B} problems stayed well-posed'
    print('receding indices seen:', sorted(seen_r))


@pytest.mark.parametrize('B', [1, 33, 97])
def test_ragged_batches(B):
    """Batches that do not fill a tile of 32 problems (1), spill one problem into a second tile (33), and give the three tile
    groups unequal shares (97 = 4 tiles): solve, controller step and closed loop against the oracle; padding lanes stay silent."""
    N = 10
    eng, orc, prob, params, md = _pair('st', 'ext', N, B)
    x0 = start_states(B, seed=71, vel=0.4)
    xg, ug = rollout_guess(x0, N, params.dt, seed=72)
    for e in (eng, orc):
        e.set_guess(xg, ug)
    st_g, st_o = eng.rti_solve(x0), orc.rti_solve(x0)
    np.testing.assert_array_equal(st_g, st_o)
    xt_g, ut_g = eng.get_temp(); xt_o, ut_o = orc.get_temp()
    ok = st_o == 0
    _close(xt_g[ok], xt_o[ok], RTOL, 'x_temp'); _close(ut_g[ok], ut_o[ok], RTOL, 'u_temp')
    for e in (eng, orc):
        e.set_guess(xg, ug); e.reset_controller()
    x_g, x_o = x0.copy(), x0.copy()
    for _ in range(3):
        u_g, ab_g = eng.controller_step(x_g); u_o, ab_o = orc.controller_step(x_o)
        np.testing.assert_array_equal(ab_g, ab_o)
        _close(u_g, u_o, RTOL, 'u')
        x_g, _ = eng.plant_step(x_g, u_g); x_o, _ = orc.plant_step(x_o, u_o)
    _close(x_g, x_o, RTOL, 'x after 3 steps')


def test_empty_active_mask_and_horizon_range():
    """No problem active: nothing moves and nothing is reported; the shortest (N = 2) and the longest (N = 128) horizon of the
    boundary solve like the oracle, horizons outside the range are refused."""
    B = 40
    eng, orc, prob, params, md = _pair('naive', 'ext', 6, B)
    x0 = start_states(B, seed=73, vel=0.3)
    xg, ug = rollout_guess(x0, 6, params.dt, seed=74)
    eng.set_guess(xg, ug); orc.set_guess(xg, ug)
    eng.rti_solve(x0); orc.rti_solve(x0)
    xt0, ut0 = (a.copy() for a in eng.get_temp())
    none = np.zeros(B, dtype=np.uint8)
    st_g, st_o = eng.rti_solve(x0 + 0.01, none), orc.rti_solve(x0 + 0.01, none)
    np.testing.assert_array_equal(st_g, st_o)
    xt1, ut1 = eng.get_temp()
    assert np.array_equal(xt0, xt1) and np.array_equal(ut0, ut1)
    for bad in (1, abi.MAX_N + 1):                                 # the horizon range of the boundary
        with pytest.raises(ValueError, match='horizon'):
            make_problem('naive', cost='ext', N=bad)
    for N1, B1 in ((2, B), (abi.MAX_N, 8)):                        # shortest and longest horizon
        eng1, orc1, prob1, params1, md1 = _pair('naive', 'ext', N1, B1)
        xg1, ug1 = rollout_guess(x0[:B1], N1, params1.dt, seed=75)
        eng1.set_guess(xg1, ug1); orc1.set_guess(xg1, ug1)
        st1 = orc1.rti_solve(x0[:B1])
        np.testing.assert_array_equal(eng1.rti_solve(x0[:B1]), st1)
        a, b = eng1.get_temp(), orc1.get_temp()
        ok = st1 == 0
        _close(a[0][ok], b[0][ok], RTOL, f'x_temp N={N1}'); _close(a[1][ok], b[1][ok], RTOL, f'u_temp N={N1}')
