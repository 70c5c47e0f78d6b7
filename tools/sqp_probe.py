"""dev: convergence of controller.solve_sqp from the trivial guess on Halton initial conditions (full steps vs l1-merit backtracking)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from safe_mpc_b200 import abi
from safe_mpc_b200.parser import Parameters, default_args
from safe_mpc_b200.env_model import AdamModel
from safe_mpc_b200.utils import get_controller
from safe_mpc_b200.cost_definition import ReachTargetEXT
from safe_mpc_b200.guess import halton_initial_states
np.set_printoptions(linewidth=250, precision=3)
for name, N, lm in (('st', 12, 1e-2), ('st', 45, 1e-2), ('st', 45, 0.5)):
    B = 8
    args = default_args(controller=name, horizon=N)
    params = Parameters(args, 'z1', rti=False); params.N = N; params.levenberg_marquardt = lm
    model = AdamModel(params, batch=B)
    c = get_controller(name, model)
    ReachTargetEXT(model, params.Q_weight, params.R_weight).set_solver_cost(c)
    c.build_controller()
    x0 = halton_initial_states(model, B)
    xg0 = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); ug0 = np.zeros((B, N, abi.NU))
    mu = np.full(B, 10.0)
    for glob in ('FIXED_STEP', 'MERIT_BACKTRACKING'):
        c.setGuess(xg0, ug0)
        print(name, N, lm, glob, 'merit0', c.merit(xg0, ug0, mu))
        for it in range(30):
            st = c.solve_sqp(x0, max_iter=1, tol=1e-7, globalization=glob)
            x, u = c._sqp_result
            if it in (0, 1, 2, 4, 8, 12, 20, 29):
                print('  it', it, 'st', st.tolist(), 'alpha', getattr(c, 'sqp_alpha', None), 'merit', c.merit(x, u, mu))
        st = c.solve_sqp(x0, max_iter=100, tol=1e-6, globalization=glob)
        x, u = c._sqp_result
        print('  after +100: status', st.tolist(), 'iters', c.sqp_iter.tolist(), 'checkGuess', c.checkGuess(x, u).tolist())
