"""dev: convergence of controller.solve_sqp from the trivial guess on Halton initial conditions"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from safe_mpc_b200 import abi
from safe_mpc_b200.parser import Parameters, default_args
from safe_mpc_b200.env_model import AdamModel
from safe_mpc_b200.utils import get_controller
from safe_mpc_b200.cost_definition import ReachTargetEXT
from safe_mpc_b200.guess import halton_initial_states
for name, N, lm in (('st', 12, 1e-3), ('st', 12, 1e-2), ('naive', 12, 1e-3), ('st', 45, 1e-2)):
    B = 16
    args = default_args(controller=name, horizon=N)
    params = Parameters(args, 'z1', rti=False); params.N = N; params.levenberg_marquardt = lm
    model = AdamModel(params, batch=B)
    c = get_controller(name, model)
    ReachTargetEXT(model, params.Q_weight, params.R_weight).set_solver_cost(c)
    c.build_controller()
    x0 = halton_initial_states(model, B)
    xg = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); ug = np.zeros((B, N, abi.NU))
    c.setGuess(xg, ug)
    hist = []
    for it in range(60):
        st = c.solve(x0)
        xt, ut = c.x_temp, c.u_temp
        step = np.maximum(np.abs(xt - xg).reshape(B, -1).max(axis=1), np.abs(ut - ug).reshape(B, -1).max(axis=1))
        hist.append(step)
        ok = st == 0
        xg[ok], ug[ok] = xt[ok], ut[ok]
        c.setGuess(xg, ug)
        if it in (0, 1, 2, 5, 10, 20, 40, 59):
            print(name, N, lm, 'it', it, 'status', st.tolist(), 'step', np.array2string(step, precision=1))
    print(name, N, 'checkGuess', c.checkGuess().tolist())
