"""Warm-start generation for the closed loop (the role of the reference's scripts/guess_acados.py:14-244).

The reference generates its ``*_guess.pkl`` files by solving each initial condition to convergence with acados' SQP and
keeps the ones that succeed.  SURVEY.md section 8(f) lists the batched SQP-to-convergence generator as the next row after
the hot path; what is here is the part the hot path needs to run at all when no guess file exists: the reference's
initial-condition sampler and a fixed number of full-step SQP iterations of the engine itself.
"""
from __future__ import annotations

import numpy as np

from . import abi


def halton_initial_states(model, count, shipped_ic=False):
    """guess_acados.py:79,100-109: Halton(nq, scramble=False) scaled to the joint box, zero velocity, kept when collision
    free.  ``shipped_ic`` reproduces the script's TEST_NOISE branch (every IC = the shipped configuration)."""
    from scipy.stats import qmc
    nq = model.nq
    sampler = qmc.Halton(nq, scramble=False)
    out = []
    while len(out) < count:
        block = qmc.scale(sampler.random(max(64, count)), model.x_min[:nq], model.x_max[:nq])
        if shipped_ic:
            block[:] = np.array([-0.3, 0.8, -1.65, 0.658, 0.0])[:nq]
        x = np.zeros((len(block), model.nx))
        x[:, :nq] = block
        ok = np.ones(len(x), dtype=bool)
        for s in range(0, len(x), model.batch):                      # the model handle evaluates `batch` rows per call
            chunk = x[s:s + model.batch]
            pad = np.vstack([chunk, np.repeat(chunk[-1:], model.batch - len(chunk), axis=0)]) if len(chunk) < model.batch else chunk
            ok[s:s + len(chunk)] = model.checkCollision(pad)[:len(chunk)]
        out.extend(x[ok])
    return np.array(out[:count])


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
