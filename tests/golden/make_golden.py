#!/usr/bin/env python
"""Generates tests/golden/rti_<controller>.npz: seeded inputs and the ORACLE's outputs of one RTI iteration and of three
closed-loop controller steps.  The reference itself cannot run offline (acados / CasADi / adam / l4casadi are absent, see
DESIGN.md), so these vectors pin the oracle restatement against regressions and give the CUDA path a fixture that does
not need the oracle library at test time; they are to be re-validated against real acados wherever it is installed.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle                                   # noqa: E402
from safe_mpc_b200 import abi                                      # noqa: E402
from tests.common import make_problem, start_states, rollout_guess  # noqa: E402

CASES = {'naive': dict(cost='ext'), 'st': dict(cost='ext'), 'htwa': dict(cost='ext'), 'receding': dict(cost='ext'),
         'zerovel': dict(cost='nls'), 'backup': dict(cost='zero')}
B, N = 4, 8


def main():
    for name, kw in CASES.items():
        prob, params, md = make_problem(name, cost=kw['cost'], N=N)
        o = Oracle(prob, B, 1)
        x0 = start_states(B, seed=101, vel=0.4)
        xg, ug = rollout_guess(x0, N, params.dt, seed=102, scale=1.5)
        o.set_guess(xg, ug); o.reset_controller()
        status = o.rti_solve(x0)
        xt, ut = o.get_temp()
        it = o.get_state(abi.STATE_QP_ITER)
        us, xs, fails, rs = [], [], [], []
        x = x0.copy()
        o.set_guess(xg, ug); o.reset_controller()
        for _ in range(3):
            u, ab = o.controller_step(x)
            x, _ = o.plant_step(x, u)
            us.append(u); xs.append(x); fails.append(o.get_state(abi.STATE_FAILS)); rs.append(o.get_state(abi.STATE_R))
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f'rti_{name}.npz')
        np.savez_compressed(out, x0=x0, xg=xg, ug=ug, x_temp=xt, u_temp=ut, status=status, qp_iter=it, u_steps=np.array(us),
                            x_steps=np.array(xs), fails=np.array(fails), r=np.array(rs), cost=kw['cost'], N=N)
        print(name, 'status', status.tolist(), 'iters', it.tolist(), '->', out)


def main_rk4():
    """tests/golden/rk4_sens.npz: the torque-input RK4 step with sensitivities (extension row (f)4) of the oracle's forward-mode AD."""
    from tests.common import random_states
    prob, params, md = make_problem('st')
    o = Oracle(prob, 4, 1)
    x = random_states(md, 48, seed=201, vel_scale=0.6)
    tau = np.random.default_rng(202).uniform(-8, 8, (48, 5))
    dts = np.array([params.dt, 0.02])
    res = [o.rk4_sens(x, tau, float(dt)) for dt in dts]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'rk4_sens.npz')
    np.savez_compressed(out, x=x, tau=tau, dt=dts, x_next=np.array([r[0] for r in res]), A=np.array([r[1] for r in res]),
                        B=np.array([r[2] for r in res]))
    print('rk4_sens ->', out)


if __name__ == '__main__':
    if 'rk4' in sys.argv[1:]:
        main_rk4()
    else:
        main()
        main_rk4()
