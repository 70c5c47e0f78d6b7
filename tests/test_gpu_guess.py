"""Warm-start generation (reference scripts/guess_acados.py:98-159, controller.py:255-272,369-373):
the batched SQP-to-convergence loop against the same loop on the oracle, the batched checkGuess against plain-numpy
restatements of the reference's predicates, and the acceptance / fallback logic of the generator."""
import numpy as np
import pytest

from safe_mpc_b200 import abi
from safe_mpc_b200.parser import Parameters, default_args
from safe_mpc_b200.env_model import AdamModel
from safe_mpc_b200.utils import get_controller
from safe_mpc_b200.cost_definition import ReachTargetEXT
from safe_mpc_b200.problem import build_problem
from safe_mpc_b200.guess import HaltonInitialStates, halton_initial_states, generate_guesses

pytestmark = pytest.mark.gpu


def _controller(name, B, N, lm=None):
    args = default_args(controller=name, horizon=N)
    params = Parameters(args, 'z1', rti=False)
    params.N = N
    if lm is not None:
        params.levenberg_marquardt = lm
    model = AdamModel(params, batch=B)
    c = get_controller(name, model)
    ReachTargetEXT(model, params.Q_weight, params.R_weight).set_solver_cost(c)
    c.build_controller()
    return c, model, params


def _oracle_sqp(prob, B, N, x0, max_iter, tol):
    """controller.solve_sqp, statement by statement, on the oracle's RTI solve."""
    from oracle.oracle import Oracle
    orc = Oracle(prob, B, 0)
    xg = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); ug = np.zeros((B, N, abi.NU))
    orc.set_guess(xg, ug)
    todo = np.ones(B, dtype=bool); status = np.full(B, 2, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
    for _ in range(max_iter):
        if not todo.any():
            break
        st = orc.rti_solve(x0, todo.astype(np.uint8))
        xt, ut = orc.get_temp()
        iters[todo] += 1
        bad = todo & (st != 0); good = todo & (st == 0)
        status[bad] = st[bad]
        step = np.maximum(np.abs(xt - xg).reshape(B, -1).max(axis=1), np.abs(ut - ug).reshape(B, -1).max(axis=1))
        xg[good], ug[good] = xt[good], ut[good]
        conv = good & (step < tol)
        status[conv] = 0
        todo &= ~(bad | conv)
        orc.set_guess(xg, ug)
    return status, iters, xg, ug


@pytest.mark.parametrize('name,lm,max_iter', [('naive', 1e-3, 40), ('st', None, 12)])
def test_sqp_loop_matches_the_oracle_loop(name, lm, max_iter):
    """With the reference's Levenberg-Marquardt term (0.5) the SQP iteration is heavily damped and ends on the iteration limit
    (status 2, which guess_acados.py:118 accepts); with a small term the naive OCP converges quadratically from the trivial guess."""
    B, N = 8, 12
    c, model, params = _controller(name, B, N, lm)
    x0 = halton_initial_states(model, B)
    c.setGuess(np.repeat(x0[:, None, :], N + 1, axis=1), np.zeros((B, N, abi.NU)))
    st = c.solve_sqp(x0, max_iter=max_iter, tol=1e-7, globalization='FIXED_STEP')
    xt, ut = c.x_temp, c.u_temp
    prob, keep = build_problem(params, name, cost='ext', N=N, model=model.data)
    st_o, it_o, x_o, u_o = _oracle_sqp(prob, B, N, x0, max_iter, 1e-7)
    assert np.array_equal(st, st_o)
    conv = st == 0
    run = st != 4
    assert np.abs(xt[run] - x_o[run]).max() < 1e-5 and np.abs(ut[run] - u_o[run]).max() < 1e-4
    if lm is None:
        assert (st == 2).all() and (c.sqp_iter == max_iter).all() and np.array_equal(it_o, c.sqp_iter)
        return
    assert conv.sum() >= B // 2 and c.sqp_iter[conv].max() < 25
    assert np.abs(c.sqp_iter[conv] - it_o[conv]).max() <= 1       # the step test sits at 1e-7: one iteration of slack
    # frozen problems keep their last QP solution in the engine; after a full step it is the last iterate
    xf, uf = c._sqp_result
    assert np.array_equal(xt[conv], xf[conv]) and np.array_equal(ut[conv], uf[conv])
    # a converged trajectory is a fixed point of the RTI solve, and it passes the reference's acceptance test
    assert c.checkGuess()[conv].all()
    c.setGuess(xt, ut)
    c.solve(x0, conv.astype(np.uint8))
    assert np.abs(c.x_temp[conv] - xt[conv]).max() < 1e-6


def test_merit_backtracking_line_search():
    """The opt-in line search of solve_sqp: step lengths on the alpha_reduction ladder down to alpha_min (parser.py:136-137),
    iterates = guess + alpha * QP step, still on the (linear) dynamics."""
    B, N = 8, 12
    c, model, params = _controller('st', B, N, 1e-2)
    assert params.globalization == 'MERIT_BACKTRACKING' and (params.alpha_reduction, params.alpha_min) == (0.3, 1e-2)
    x0 = halton_initial_states(model, B)
    xg0, ug0 = np.repeat(x0[:, None, :], N + 1, axis=1), np.zeros((B, N, abi.NU))
    c.setGuess(xg0, ug0)
    alphas = []
    xp, up = xg0, ug0
    for _ in range(4):
        st = c.solve_sqp(x0, max_iter=1, tol=1e-7, globalization='MERIT_BACKTRACKING')     # one line-searched SQP iteration per call
        x, u = c._sqp_result
        a = c.sqp_alpha[:, None, None]
        run = st != 4
        assert np.array_equal(x, c.x_guess) and np.array_equal(u, c.u_guess)
        assert np.allclose(x[run], (xp + a * (c.x_temp - xp))[run], atol=1e-12) and np.allclose(u[run], (up + a * (c.u_temp - up))[run], atol=1e-12)
        alphas.append(c.sqp_alpha.copy())
        xp, up = x, u
    alphas = np.array(alphas)
    assert set(np.unique(np.round(alphas, 6))) <= {1.0, 0.3, 0.09, 0.027, 0.01}        # alpha_reduction ladder down to alpha_min
    assert (alphas < 1.0).any()                                    # the search does shorten steps on this problem
    mu = np.full(B, 1.0)
    assert np.isfinite(c.merit(x, u, mu)).all() and (c.merit(x, u, 10 * mu) >= c.merit(x, u, mu)).all()
    assert model.checkDynamicsConstraints(x, u)[model.checkTorqueConstraints(x, u)].all()
    with pytest.raises(ValueError, match='globalization'):
        c.solve_sqp(x0, max_iter=1, globalization='nope')


def test_check_guess_agrees_with_plain_numpy_predicates():
    B, N = 8, 12
    c, model, params = _controller('st', B, N)
    x0 = halton_initial_states(model, B)
    c.setGuess(np.repeat(x0[:, None, :], N + 1, axis=1), np.zeros((B, N, abi.NU)))
    c.solve_sqp(x0, max_iter=30, tol=1e-7, globalization='FIXED_STEP')
    x, u = c.x_temp.copy(), c.u_temp.copy()
    # spoil some trajectories: a state beyond its bound (1), a dynamics defect (2), a torque far out of range (3)
    x[1, 5, 0] = model.x_max[0] + 10 * params.tol_x
    x[2, 7, 6] += 1e-3
    u[3, 2, :] = 500.0
    dt = params.dt
    tau = np.stack([model.tau_fun(x[:, k], u[:, k]) for k in range(N)], axis=1)
    tau_ok = ((tau >= model.tau_min - params.tol_tau) & (tau <= model.tau_max + params.tol_tau)).all(axis=(1, 2))
    bounds_ok = ((x >= model.x_min - params.tol_x) & (x <= model.x_max + params.tol_x)).all(axis=(1, 2))
    assert np.array_equal(model.checkTorqueConstraints(x, u), tau_ok) and not tau_ok[3]
    assert np.array_equal(model.checkRunningConstraints(x, u), bounds_ok & tau_ok & model.checkCollision(x[:, 0]))
    # env_model.py:226-234 with the double-integrator update of env_model.py:222-223 (valid where the torque stays in range)
    xs = np.zeros_like(x); xs[:, 0] = x[:, 0]
    for k in range(N):
        xs[:, k + 1, :5] = xs[:, k, :5] + dt * xs[:, k, 5:] + 0.5 * dt * dt * u[:, k]
        xs[:, k + 1, 5:] = xs[:, k, 5:] + dt * u[:, k]
    dyn_ok = np.linalg.norm((x - xs).reshape(B, -1), axis=1) < params.tol_dyn * np.sqrt(N + 1)
    got = model.checkDynamicsConstraints(x, u)
    sel = tau_ok                                                  # (saturated rows take the forward-dynamics branch)
    assert np.array_equal(got[sel], dyn_ok[sel]) and not got[2]
    assert not bounds_ok[1] and not model.checkRunningConstraints(x, u)[1]


def test_generator_accepts_in_sequence_order_and_falls_back():
    B, N, count = 8, 12, 10
    net, model, params = _controller('st', B, N)
    naive, _, _ = _controller('naive', B, N)
    zerovel, _, _ = _controller('zerovel', B, N)
    out, stats = generate_guesses(net, naive, zerovel, count, sqp_iter=30, tol=1e-7)
    assert stats['succ'] >= count and stats['rounds'] >= 2
    for k in ('net', 'naive', 'zerovel'):
        assert out[k]['xg'].shape == (count, N + 1, 10) and out[k]['ug'].shape == (count, N, 5)
        assert np.allclose(out[k]['xg'][:, 0], out['net']['xg'][:, 0], atol=1e-9)   # same initial conditions in every file
    # accepted initial conditions are a subsequence of the Halton sequence, in order (guess_acados.py:98-125)
    seq = HaltonInitialStates(model).draw(stats['rounds'] * B)
    pos = [int(np.where((np.abs(seq - x0) < 1e-9).all(axis=1))[0][0]) for x0 in out['net']['xg'][:, 0]]
    assert pos == sorted(pos) and len(set(pos)) == count
    # zero-velocity guesses: either their own solution (terminal velocity 0) or the network controller's trajectory
    own = np.abs(out['zerovel']['xg'][:, -1, 5:]).max(axis=1) < 1e-5
    same = np.array([np.array_equal(a, b) for a, b in zip(out['zerovel']['xg'], out['net']['xg'])])
    assert (own | same).all()
    assert stats['zerovel_own'] >= own[~same].sum()


def test_guess_script_writes_the_reference_files_and_mpc_reads_them(tmp_path, monkeypatch):
    import importlib.util
    import os
    import pickle
    here = os.path.dirname(__file__)

    def load(name):
        spec = importlib.util.spec_from_file_location(name + '_script', os.path.join(here, '..', 'scripts', name + '.py'))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        orig = mod.Parameters.__init__

        def init(self, *a, **k):
            orig(self, *a, **k)
            self.DATA_DIR, self.n_steps, self.test_num, self.nlp_max_iter = str(tmp_path) + '/', 10, 6, min(self.nlp_max_iter, 30)
        monkeypatch.setattr(mod.Parameters, '__init__', init)
        return mod
    load('guess_acados').main(['-c', 'st', '--horizon', '12', '--batch', '6'])
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith('_guess.pkl'))
    # guess_acados.py:235-244: the naive and zero-velocity files and one file per network controller
    assert 'z1_naive_12hor_10sm_use_netNone__q_collision_margins_0.0_0.0_guess.pkl' in files
    assert 'z1_zerovel_12hor_10sm_use_netNone__q_collision_margins_0.0_0.0_guess.pkl' in files
    # one file per name of get_ocp_acados' dict that guess_acados.py:242 lists: st, htwa, receding, real_receding, constraint_everywhere
    assert 'z1_st_12hor_10sm_use_netTrue__q_collision_margins_0.0_0.0_guess.pkl' in files and len(files) == 7
    assert not any('_parallel_' in f or '_stwa_' in f for f in files)
    d = pickle.load(open(tmp_path / 'z1_st_12hor_10sm_use_netTrue__q_collision_margins_0.0_0.0_guess.pkl', 'rb'))
    assert set(d) == {'xg', 'ug'} and d['xg'].shape == (6, 13, 10) and d['ug'].shape == (6, 12, 5)
    # every network guess is generated with HTWAController (utils.py:46-62): its terminal node is inside the viable set, which is what
    # STWA / HTWA / receding latch as x_viable in setGuess (controller.py:390-393)
    c, model, params = _controller('htwa', 6, 12)
    assert c.checkSafeConstraints(d['xg'][:, -1]).all()
    # the closed-loop script starts from that file (mpc.py:79-84)
    load('mpc').main(['-c', 'st', '--horizon', '12', '--back_hor', '12', '--batch', '6'])
    out = [f for f in os.listdir(tmp_path) if f.endswith('_mpc.pkl')]
    assert len(out) == 1
    x = pickle.load(open(tmp_path / out[0], 'rb'))['x']
    assert np.allclose(x[:, 0], d['xg'][:, 0])


def test_initialize_keeps_the_solution_where_it_passes():
    """NaiveController.initialize (controller.py:260-272) for a batch: trivial guess, one solve, the solution becomes the
    guess exactly where the status is 0 and checkGuess holds."""
    B, N = 8, 12
    c, model, params = _controller('naive', B, N)
    x0 = halton_initial_states(model, B)
    ok = c.initialize(x0)
    assert ok.shape == (B,) and set(np.unique(ok)) <= {0, 1}
    st = c.last_status if hasattr(c, 'last_status') else None
    chk = c.checkGuess()
    xg, ug = c.getGuess()
    xt, ut = c.x_temp, c.u_temp
    triv = np.repeat(x0[:, None, :], N + 1, axis=1)
    for b in range(B):
        if ok[b]:
            assert chk[b] and np.array_equal(xg[b], xt[b]) and np.array_equal(ug[b], ut[b])
        else:
            assert np.array_equal(xg[b], triv[b]) and np.abs(ug[b]).max() == 0.0
    assert np.array_equal(ok.astype(bool), chk & (np.asarray(st) == 0)) if st is not None else True
    # an explicit u0 is broadcast over the horizon (controller.py:263-265)
    c.initialize(x0, u0=np.full(5, 0.1))
