"""dev: which outputs of the two prep forms differ?  (IPM cut after 1 / 2 iterations)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.common import make_problem, start_states, rollout_guess
from safe_mpc_b200.engine import Engine
B, N = 64, 20
for itmax in (1, 2, 3):
    out = {}
    for tag, env in (('coop', {'SMPC_QP_TAIL': '100000', 'SMPC_QP_PREP': 'coop'}), ('thread', {'SMPC_QP_TAIL': '100000', 'SMPC_QP_PREP': 'thread'})):
        os.environ.update(env)
        prob, params, md = make_problem('st', N=N)
        prob.qp_iter_max = itmax
        x0 = start_states(B, seed=3); xg, ug = rollout_guess(x0, N, params.dt, seed=4, scale=1.0)
        e = Engine(prob, B, 0); e.set_guess(xg, ug); e.rti_solve(x0)
        out[tag] = e.get_qp()
        e.close()
    names = ['dz', 'pi', 'lam', 't']
    for n, a, b in zip(names, out['coop'], out['thread']):
        d = np.abs(a - b)
        print('iter_max', itmax, n, 'max abs diff', d.max(), 'n differing', int((d > 0).sum()), 'first idx', np.argwhere(d > 0)[:3].tolist())
