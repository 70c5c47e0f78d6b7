"""Warm-start generation for the closed loop (the role of the reference's scripts/guess_acados.py:14-244).

The reference generates its ``*_guess.pkl`` files by solving each initial condition to convergence with acados' SQP from the
trivial guess and keeps the ones whose status is 0 or 2 and whose solution passes ``checkGuess``; the naive and zero-velocity
controllers are solved from the same initial conditions and fall back to the network controller's trajectory where they fail
(guess_acados.py:113-150).  Here the same procedure runs for a whole batch of initial conditions at a time:
``generate_guesses`` (the loop of guess_acados.py:98-159), on ``controller.solve_sqp`` (RTI iterations of the engine to
convergence, with an l1-merit backtracking line search) and the batched ``checkGuess``.  ``sqp_guess`` -- a fixed number of full-step iterations, no
acceptance test -- is what bench.py and scripts/mpc.py use when no guess file exists.
"""
from __future__ import annotations

import numpy as np

from . import abi


class HaltonInitialStates:
    """guess_acados.py:79,100-109: one Halton(nq, scramble=False) sequence scaled to the joint box, zero velocity, a point is
    kept when it is collision free.  The sequence continues across calls, like the reference's sampler in its while-loop.
    ``shipped_ic`` reproduces the script's TEST_NOISE branch (every IC = the shipped configuration)."""

    def __init__(self, model, shipped_ic=False):
        from scipy.stats import qmc
        self.model, self.shipped_ic = model, shipped_ic
        self.sampler = qmc.Halton(model.nq, scramble=False)
        self.skipped = 0
        self._left = np.zeros((0, model.nx))

    def draw(self, count):
        from scipy.stats import qmc
        model, nq = self.model, self.model.nq
        out = [self._left]
        have = len(self._left)
        while have < count:
            block = qmc.scale(self.sampler.random(max(64, count)), model.x_min[:nq], model.x_max[:nq])
            if self.shipped_ic:
                block[:] = np.array([-0.3, 0.8, -1.65, 0.658, 0.0])[:nq]
            x = np.zeros((len(block), model.nx))
            x[:, :nq] = block
            ok = np.ones(len(x), dtype=bool)
            for s in range(0, len(x), model.batch):                  # the model handle evaluates `batch` rows per call
                chunk = x[s:s + model.batch]
                pad = np.vstack([chunk, np.repeat(chunk[-1:], model.batch - len(chunk), axis=0)]) if len(chunk) < model.batch else chunk
                ok[s:s + len(chunk)] = model.checkCollision(pad)[:len(chunk)]
            self.skipped += int((~ok).sum())
            out.append(x[ok])
            have += int(ok.sum())
        allx = np.concatenate(out)
        self._left = allx[count:]                                    # points drawn past `count` stay in sequence order
        return allx[:count]


def halton_initial_states(model, count, shipped_ic=False):
    """The first ``count`` collision-free initial conditions of the reference's sampler."""
    return HaltonInitialStates(model, shipped_ic).draw(count)


def generate_guesses(ctrl_net, ctrl_naive, ctrl_zerovel, count, shipped_ic=False, max_rounds=50, sqp_iter=None, tol=1e-6, globalization=None):
    """guess_acados.py:98-159 for batches of initial conditions.  Every round draws one batch of collision-free Halton points,
    solves the network controller to convergence from the trivial guess and accepts a problem when its status is 0 or 2 and
    ``checkGuess`` holds; for the accepted ones the naive and the zero-velocity controller are solved from the same trivial
    guess and keep their own solution when it passes the same test, the network controller's otherwise.  Accepted initial
    conditions keep the order of the Halton sequence, as in the reference.
    -> dict name -> {'xg': [count, N+1, nx], 'ug': [count, N, nu]} for 'net', 'naive', 'zerovel', and a statistics dict."""
    model = ctrl_net.model
    B, N = ctrl_net.B, ctrl_net.N
    sampler = HaltonInitialStates(model, shipped_ic)
    acc = {k: {'xg': [], 'ug': []} for k in ('net', 'naive', 'zerovel')}
    stats = {'rounds': 0, 'fails': 0, 'succ': 0, 'naive_own': 0, 'zerovel_own': 0}
    others = [c for c in (('naive', ctrl_naive), ('zerovel', ctrl_zerovel)) if c[1] is not None]
    for _ in range(max_rounds):
        if stats['succ'] >= count:
            break
        stats['rounds'] += 1
        x0 = sampler.draw(B)
        xg0 = np.repeat(x0[:, None, :], N + 1, axis=1)
        ug0 = np.zeros((B, N, abi.NU))
        ctrl_net.setGuess(xg0, ug0)
        st = ctrl_net.solve_sqp(x0, max_iter=sqp_iter, tol=tol, globalization=globalization)
        x_net, u_net = ctrl_net._sqp_result
        ok = ((st == 0) | (st == 2)) & ctrl_net.checkGuess(x_net, u_net)          # guess_acados.py:118
        stats['fails'] += int((~ok).sum())
        stats['succ'] += int(ok.sum())
        acc['net']['xg'].append(x_net[ok]); acc['net']['ug'].append(u_net[ok])
        for name, c in others:                                                    # guess_acados.py:132-150
            c.setGuess(xg0, ug0)
            s2 = c.solve_sqp(x0, max_iter=sqp_iter, tol=tol, active=ok, globalization=globalization)
            x_c, u_c = c._sqp_result
            own = ok & ((s2 == 0) | (s2 == 2)) & c.checkGuess(x_c, u_c)
            xo, uo = np.where(own[:, None, None], x_c, x_net), np.where(own[:, None, None], u_c, u_net)
            stats[name + '_own'] += int(own.sum())
            acc[name]['xg'].append(xo[ok]); acc[name]['ug'].append(uo[ok])
    out = {}
    for k, v in acc.items():
        if v['xg']:
            out[k] = {'xg': np.concatenate(v['xg'])[:count], 'ug': np.concatenate(v['ug'])[:count]}
    stats['skipped'] = sampler.skipped
    return out, stats


def sqp_guess(controller, x0, iters=5):
    """Full-step SQP iterations from the constant guess (x0 repeated, u = 0): each iteration is one batched RTI solve whose
    result becomes the next guess.  -> (x_guess [B, N+1, nx], u_guess [B, N, nu], last status [B])."""
    B, N = controller.B, controller.N
    xg = np.repeat(np.asarray(x0, dtype=np.float64)[:, None, :], N + 1, axis=1).copy()
    ug = np.zeros((B, N, abi.NU))
    controller.setGuess(xg, ug)
    status = np.zeros(B, dtype=np.int32)
    for _ in range(iters):
        status = controller.solve(x0)
        xt, ut = controller.x_temp, controller.u_temp
        ok = (status == 0) | (status == 2)                             # guess_acados.py:118 accepts 0 and 2
        xg[ok], ug[ok] = xt[ok], ut[ok]
        controller.setGuess(xg, ug)
    controller.reset_controller()
    return xg, ug, status
