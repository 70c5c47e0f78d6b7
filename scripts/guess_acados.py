#!/usr/bin/env python
"""Warm-start file generator -- writes the reference's ``*_guess.pkl`` ({'xg', 'ug'}, guess_acados.py:235-244) from the
reference's Halton initial conditions and full-step SQP iterations of the engine (safe_mpc_b200/guess.py).

    python scripts/guess_acados.py -c st --horizon 45 --alpha 10 [--batch 100]
"""
import os
import pickle
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from safe_mpc_b200.parser import Parameters, parse_args            # noqa: E402
from safe_mpc_b200.env_model import AdamModel                      # noqa: E402
from safe_mpc_b200.utils import get_controller                     # noqa: E402
from safe_mpc_b200.cost_definition import ReachTargetEXT           # noqa: E402
from safe_mpc_b200.guess import halton_initial_states, sqp_guess   # noqa: E402


def main(argv=None):
    args = parse_args(argv)
    params = Parameters(args, args['system'], rti=True)
    params.q_margin = args['joint_bounds_margin']
    params.collision_margin = args['collision_margin']
    params.alpha = args['alpha']
    params.N = args['horizon']
    batch = args['batch'] or params.test_num
    model = AdamModel(params, batch=batch)
    cont_name = args['controller']
    controller = get_controller(cont_name, model)
    ReachTargetEXT(model, params.Q_weight, params.R_weight).set_solver_cost(controller)
    controller.build_controller(args['build'])
    x_init = halton_initial_states(model, batch)
    xg, ug, st = sqp_guess(controller, x_init, iters=20)
    use_net = True if cont_name not in ('naive', 'zerovel') else None
    path = (f'{params.DATA_DIR}{args["system"]}_{cont_name}_{params.N}hor_{int(params.alpha)}sm_use_net{use_net}__q_collision_margins_'
            f'{params.q_margin}_{params.collision_margin}_guess.pkl')
    os.makedirs(params.DATA_DIR, exist_ok=True)
    with open(path, 'wb') as f:
        pickle.dump({'xg': xg, 'ug': ug}, f)
    print(f'{int((st == 0).sum())}/{batch} solves ended with status 0; saved {path}')


if __name__ == '__main__':
    main()
