"""dev: parity with the oracle and agreement of the kernel forms at the long horizons of BASELINE.json configs[4] (N up to 80) and at SMPC_MAX_N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.common import make_problem, start_states, rollout_guess
from safe_mpc_b200.engine import Engine
from safe_mpc_b200 import abi
from oracle.oracle import Oracle
for ctrl, N in (('st', 60), ('receding', 80), ('htwa', 128)):
    prob, params, md = make_problem(ctrl, N=N)
    B = 96
    x0 = start_states(B, seed=N)
    xg, ug = rollout_guess(x0, N, params.dt, seed=N + 1, scale=0.5)
    res = {}
    os.environ['SMPC_QP_PREP'] = 'thread'                 # one prep form, so that only the Riccati forms differ between the two runs
    for tag, env in (('tail', {'SMPC_QP_TAIL': '100000'}), ('bulk', {'SMPC_QP_TAIL': '0'})):
        os.environ.update(env)
        e = Engine(prob, B, 0)
        e.set_guess(xg, ug); st = e.rti_solve(x0); xt, ut = e.get_temp()
        res[tag] = (np.asarray(st).copy(), xt.copy(), ut.copy(), e.get_state(abi.STATE_QP_ITER).copy())
        e.close()
    o = Oracle(prob, B, 0); o.set_guess(xg, ug); st_o = o.rti_solve(x0); xt_o, ut_o = o.get_temp()
    ok = (st_o == 0)
    same = all(np.array_equal(res['tail'][i], res['bulk'][i]) for i in range(4))
    err = np.abs(res['tail'][1][ok] - xt_o[ok]).max() / max(1.0, np.abs(xt_o[ok]).max())
    print(ctrl, 'N', N, 'forms bitwise equal:', same, ' status equal to oracle:', bool((res['tail'][0] == st_o).all()), ' solved', int(ok.sum()), 'of', B,
          ' max rel err vs oracle %.2e' % err, ' ipm it mean %.1f' % res['tail'][3].mean())
