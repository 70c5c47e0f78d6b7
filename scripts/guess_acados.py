#!/usr/bin/env python
"""Warm-start file generator -- the reference's scripts/guess_acados.py for a batch of initial conditions at a time.

Writes the reference's ``*_guess.pkl`` files ({'xg', 'ug'}, guess_acados.py:235-244, same file names): Halton initial
conditions, the controller named by ``-c`` solved to convergence from the trivial guess and accepted on status 0 / 2 +
``checkGuess``; the naive and zero-velocity files hold their own solutions where those pass the same test and the network
controller's trajectory otherwise (guess_acados.py:132-150).  All of it is safe_mpc_b200.guess.generate_guesses.

    python scripts/guess_acados.py -c st --horizon 45 --alpha 10 [--batch 100]
"""
import os
import pickle
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from safe_mpc_b200.parser import Parameters, parse_args            # noqa: E402
from safe_mpc_b200.env_model import AdamModel                      # noqa: E402
from safe_mpc_b200.utils import get_ocp_acados                     # noqa: E402
from safe_mpc_b200.cost_definition import ReachTargetEXT, ReachTargetNLS   # noqa: E402
from safe_mpc_b200.guess import generate_guesses                   # noqa: E402

# guess_acados.py:240-243: the network guess is stored under every name of get_ocp_acados' dict that is in this list
NET_FILE_NAMES = ['st', 'stwa', 'htwa', 'receding', 'real_receding', 'receding_parallel', 'parallel2', 'constraint_everywhere']


def make_controller(name, args, batch):
    """guess_acados.py:14-71: Parameters(rti=False) -> acados 'SQP' with nlp_max_iter and MERIT_BACKTRACKING (parser.py:139); the
    controller class of get_ocp_acados (every network controller -> HTWAController, utils.py:46-62); cost: ReachTargetEXT for the
    network controller (guess_acados.py:31,43), ReachTargetNLS for naive / zerovel (guess_acados.py:33,62-68)."""
    params = Parameters(args, args['system'], rti=False)
    params.q_margin = args['joint_bounds_margin']
    params.collision_margin = args['collision_margin']
    params.alpha = args['alpha']
    params.N = args['horizon']
    model = AdamModel(params, batch=batch)
    controller, names = get_ocp_acados(name, model)
    cost = ReachTargetNLS if name in ('naive', 'zerovel') else ReachTargetEXT
    cost(model, params.Q_weight, params.R_weight).set_solver_cost(controller)
    controller.build_controller(args['build'])
    return controller, params, names


def main(argv=None):
    args = parse_args(argv)
    cont_name = args['controller']
    probe = Parameters(args, args['system'], rti=False)
    count = probe.test_num
    batch = args['batch'] or count
    ctrl, params, names = make_controller(cont_name, args, batch)
    naive = zerovel = None
    if cont_name not in ('naive', 'zerovel'):
        naive, _, _ = make_controller('naive', args, batch)
        zerovel, _, _ = make_controller('zerovel', args, batch)
    t0 = time.time()
    out, stats = generate_guesses(ctrl, naive, zerovel, count, sqp_iter=min(params.nlp_max_iter, args.get('sqp_iter') or 100),
                                  globalization=params.globalization)
    print(f'{stats["succ"]} accepted, {stats["fails"]} failed, {stats["skipped"]} initial conditions in collision, '
          f'{stats["rounds"]} batches of {batch}, {time.time() - t0:.1f} s')

    def path(name, use_net):
        return (f'{params.DATA_DIR}{args["system"]}_{name}_{params.N}hor_{int(params.alpha)}sm_use_net{use_net}__q_collision_margins_'
                f'{params.q_margin}_{params.collision_margin}_guess.pkl')

    os.makedirs(params.DATA_DIR, exist_ok=True)
    files = []
    if cont_name in ('naive', 'zerovel'):
        files.append((path(cont_name, None), out['net']))
    else:
        files += [(path('naive', None), out['naive']), (path('zerovel', None), out['zerovel'])]
        files += [(path(c, True), out['net']) for c in names if c in NET_FILE_NAMES]
    for p, d in files:
        with open(p, 'wb') as f:
            pickle.dump({'xg': d['xg'], 'ug': d['ug']}, f)
        print('saved', p, d['xg'].shape)


if __name__ == '__main__':
    main()
