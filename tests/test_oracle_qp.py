"""Pins the oracle's QP solve: the returned primal/dual point must satisfy the KKT conditions of the QP rebuilt
independently (numpy) from the stage records, for every controller family, incl. soft rows and failure paths."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from safe_mpc_b200 import abi
from tests.common import make_problem, start_states, constant_guess, rollout_guess, kkt_residuals


@pytest.mark.parametrize('controller,cost', [('naive', 'ext'), ('naive', 'nls'), ('zerovel', 'ext'), ('st', 'ext'),
                                             ('htwa', 'ext'), ('receding', 'ext'), ('constraint_everywhere', 'ext'),
                                             ('backup', 'zero')])
def test_qp_solution_satisfies_kkt(controller, cost):
    prob, params, md = make_problem(controller, cost=cost, N=20)
    B = 6
    o = Oracle(prob, B, 2)
    x0 = start_states(B, seed=11, vel=0.5)
    xg, ug = rollout_guess(x0, prob.N, params.dt, seed=12)
    o.set_guess(xg, ug)
    if controller == 'receding':
        o.set_state(abi.STATE_R, np.full(B, 9, dtype=np.int32))
    st = o.rti_solve(x0 + 1e-3)          # x0 differs from the guess' first node
    lin = o.get_lin()
    dz, pi, lam, t = o.get_qp()
    xt, ut = o.get_temp()
    qst = o.get_state(abi.STATE_QP_STATUS)
    n_ok = 0
    for b in range(B):
        if qst[b] != 0:
            continue
        n_ok += 1
        r = kkt_residuals(prob, lin[b], x0[b] + 1e-3, dz[b], pi[b], lam[b], t[b])
        assert r['stat'] < 5e-6 and r['eq'] < 1e-7 and r['ineq'] < 1e-7 and r['comp'] < 1e-6, (b, r)
        # full step
        np.testing.assert_allclose(xt[b, :-1], xg[b, :-1] + dz[b, :-1, 5:], atol=1e-14)
        np.testing.assert_allclose(xt[b, -1], xg[b, -1] + dz[b, -1, :10], atol=1e-14)
        np.testing.assert_allclose(ut[b], ug[b] + dz[b, :-1, :5], atol=1e-14)
        np.testing.assert_allclose(xt[b, 0], x0[b] + 1e-3, atol=1e-7)   # lbx_0 = ubx_0 = x0
        assert st[b] == 0
    assert n_ok >= B - 1
    if controller == 'receding':
        # only stage r and the terminal stage carry an active viability row (controller.py:452-469)
        for k in range(1, prob.N):
            assert lin[0, k, abi.REC_NNROW] == 1.0
            assert (lin[0, k, abi.REC_NN] == 5e5) == (k != 9)
        assert lin[0, prob.N, abi.REC_SOFT] == params.ws_t


def test_infeasible_qp_reports_failure_and_keeps_guess():
    prob, params, md = make_problem('zerovel', N=4)   # 4 steps cannot brake 3 rad/s within torque limits
    o = Oracle(prob, 2, 1)
    x0 = start_states(2, seed=13, vel=0.0)
    x0[:, 5:] = 3.0
    xg, ug = constant_guess(x0, prob.N)
    o.set_guess(xg, ug)
    st = o.rti_solve(x0)
    xt, ut = o.get_temp()
    qst = o.get_state(abi.STATE_QP_STATUS)
    for b in range(2):
        if qst[b] in (2, 3):
            assert st[b] == 4
            np.testing.assert_array_equal(xt[b], xg[b])
        else:   # max-iter is tolerated by SQP_RTI (status 0, step taken)
            assert st[b] == 0


def test_active_mask_skips_problems():
    prob, params, md = make_problem('naive', N=10)
    o = Oracle(prob, 4, 2)
    x0 = start_states(4, seed=14)
    xg, ug = constant_guess(x0, prob.N)
    o.set_guess(xg, ug)
    o.rti_solve(x0, active=np.array([1, 0, 1, 0], dtype=np.uint8))
    xt, ut = o.get_temp()
    assert np.abs(ut[0]).max() > 0 and np.abs(ut[2]).max() > 0
    assert np.abs(ut[1]).max() == 0 and np.abs(ut[3]).max() == 0


@pytest.mark.parametrize('controller', ['naive', 'htwa'])
def test_qp_solution_matches_an_independent_dense_solver(controller):
    """a7 from another side: the stage QPs assembled into one dense QP (numpy) and handed to scipy's SLSQP, a solver that
    shares nothing with the Riccati interior-point method.  The QP is strictly convex (Levenberg-Marquardt term), so the
    minimiser is unique: both must reach the same objective value and the same point.  Distances are loose where the
    curvature is only lm * dt = 2.5e-3 (the IPM stops at a stationarity residual of 1e-6)."""
    from scipy.optimize import minimize
    from tests.common import stage_qp, dyn_mats
    N = 4
    prob, params, md = make_problem(controller, cost='ext', N=N)
    B = 3
    o = Oracle(prob, B, 1)
    x0 = start_states(B, seed=21, vel=0.5)
    xg, ug = rollout_guess(x0, N, params.dt, seed=22)
    o.set_guess(xg, ug)
    o.rti_solve(x0 + 1e-3)
    lin = o.get_lin()
    dz = o.get_qp()[0]
    qst = o.get_state(abi.STATE_QP_STATUS)
    A, Bm = dyn_mats(prob.dt)
    lbx, ubx = np.array(prob.lbx[:]), np.array(prob.ubx[:])
    lbx_e, ubx_e = np.array(prob.lbx_e[:]), np.array(prob.ubx_e[:])
    nz_tot = N * 15 + 10
    off = [15 * k for k in range(N + 1)]
    checked = 0
    for b in range(B):
        if qst[b] != 0:
            continue
        H = np.zeros((nz_tot, nz_tot)); g = np.zeros(nz_tot)
        Aeq, beq, Ain, bin_ = [], [], [], []                      # Aeq z = beq,  Ain z >= bin
        for k in range(N + 1):
            rec = lin[b, k]
            xk = rec[abi.REC_X:abi.REC_X + 10]
            lo, hi = (x0[b] + 1e-3 - xk,) * 2 if k == 0 else ((lbx_e - xk, ubx_e - xk) if k == N else (lbx - xk, ubx - xk))
            Hk, gk, rows, rlo, rhi, ids, soft = stage_qp(prob, rec, k, lo, hi)
            assert soft is None
            nz = len(gk); nu = nz - 10
            sl = slice(off[k], off[k] + nz)
            H[sl, sl] = Hk; g[sl] = gk
            for a, l_, h_, rid in zip(rows, rlo, rhi, ids):
                full = np.zeros(nz_tot); full[sl] = a
                if k == 0 and rid < 10:
                    Aeq.append(full); beq.append(l_)
                    continue
                Ain.append(full); bin_.append(l_)
                if h_ < 1e5:
                    Ain.append(-full); bin_.append(-h_)
            if k < N:
                e = np.zeros((10, nz_tot))
                e[:, off[k]:off[k] + 5] = Bm; e[:, off[k] + 5:off[k] + 15] = A
                nxt = off[k + 1] + (0 if k + 1 == N else 5)
                e[:, nxt:nxt + 10] -= np.eye(10)
                Aeq.extend(e); beq.extend(-rec[abi.REC_B:abi.REC_B + 10])
        Aeq, beq, Ain, bin_ = map(np.array, (Aeq, beq, Ain, bin_))
        z_o = np.concatenate([dz[b, k, :15] for k in range(N)] + [dz[b, N, :10]])
        f = lambda z: 0.5 * z @ H @ z + g @ z
        res = minimize(f, np.zeros(nz_tot), jac=lambda z: H @ z + g, method='SLSQP',
                       constraints=[{'type': 'eq', 'fun': lambda z: Aeq @ z - beq, 'jac': lambda z: Aeq},
                                    {'type': 'ineq', 'fun': lambda z: Ain @ z - bin_, 'jac': lambda z: Ain}],
                       options={'ftol': 1e-15, 'maxiter': 1000})
        assert res.success, res.message
        assert np.abs(Aeq @ z_o - beq).max() < 1e-7 and (Ain @ z_o - bin_).min() > -1e-7          # the oracle's point is feasible
        assert abs(f(z_o) - res.fun) <= 1e-6 * max(1.0, abs(res.fun)), (f(z_o), res.fun)
        assert np.abs(z_o - res.x).max() <= 2e-3 * max(1.0, np.abs(res.x).max())
        checked += 1
    assert checked >= 2
