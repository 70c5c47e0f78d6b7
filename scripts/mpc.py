#!/usr/bin/env python
"""Closed-loop RTI benchmark -- the batched equivalent of the reference's scripts/mpc.py (same CLI, same result file).

    python scripts/mpc.py -c st --horizon 45 --alpha 10 [--noise 5 --control_noise 1 --batch 1000]

All ``test_num`` (or ``--batch``) tests run concurrently on the GPU: controller.step, safe-abort handling with the backup
OCP, plant step on the per-test perturbed model, bounds / collision checks and the outcome bookkeeping of
mpc.py:102-291 are one ``Sim`` object (smpc_sim_*).  The guesses are read from the reference's ``*_guess.pkl`` when it
exists, otherwise generated on the fly (safe_mpc_b200/guess.py).  Exit code = number of collisions (mpc.py:317).
"""
import os
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from safe_mpc_b200.parser import Parameters, parse_args            # noqa: E402
from safe_mpc_b200.env_model import AdamModel                      # noqa: E402
from safe_mpc_b200.utils import get_controller                     # noqa: E402
from safe_mpc_b200.controller import SafeBackupController          # noqa: E402
from safe_mpc_b200.cost_definition import ReachTargetEXT, ZeroCost  # noqa: E402
from safe_mpc_b200.engine import Sim                               # noqa: E402
from safe_mpc_b200.guess import halton_initial_states, sqp_guess   # noqa: E402
from safe_mpc_b200 import abi                                      # noqa: E402


def main(argv=None):
    args = parse_args(argv)
    model_name = args['system']
    params = Parameters(args, model_name, rti=True)
    params.q_margin = args['joint_bounds_margin']
    params.collision_margin = args['collision_margin']
    params.act = args['activation']
    params.alpha = args['alpha']
    horizon = args['horizon']
    params.N = horizon
    batch = args['batch'] or params.test_num
    model = AdamModel(params, batch=batch)
    cont_name = args['controller']
    controller = get_controller(cont_name, model)
    ReachTargetEXT(model, params.Q_weight, params.R_weight).set_solver_cost(controller)      # mpc.py:48-51
    controller.build_controller(args['build'])

    param_backup = Parameters(args, model_name, rti=True)                                       # mpc.py:54-66
    param_backup.q_margin = args['joint_bounds_margin']
    param_backup.collision_margin = args['collision_margin']
    param_backup.N = args['back_hor']
    model_backup = AdamModel(param_backup, batch=batch)
    safe_ocp = SafeBackupController(model_backup)
    ZeroCost(model_backup).set_solver_cost(safe_ocp)
    safe_ocp.build_controller(build=args['build'], name=cont_name)

    use_net = True if cont_name not in ('naive', 'zerovel') else None                           # reset_controller, controller.py:234-238
    base = (f'{params.DATA_DIR}{model_name}_{cont_name}_{horizon}hor_{int(params.alpha)}sm_use_net{use_net}__q_collision_margins_'
            f'{params.q_margin}_{params.collision_margin}')
    guess_file = base + '_guess.pkl'
    if os.path.isfile(guess_file):                                                              # mpc.py:79-84
        data = pickle.load(open(guess_file, 'rb'))
        x_guess, u_guess = np.asarray(data['xg'])[:batch], np.asarray(data['ug'])[:batch]
        if len(x_guess) < batch:
            raise SystemExit(f'{guess_file} holds {len(x_guess)} guesses, batch is {batch}')
    elif args.get('generate_guess'):
        x_init = halton_initial_states(model, batch)
        x_guess, u_guess, st = sqp_guess(controller, x_init, iters=10)
        print(f'WARNING: {guess_file} does not exist; --generate-guess: {batch} warm starts from 10 full-step SQP iterations, NOT checked by '
              f'checkGuess ({int((st == 0).sum())} converged RTI solves at the last iteration).  Run scripts/guess_acados.py for the '
              "reference's accepted guesses.", file=sys.stderr)
    else:                                                                                       # the reference fails on the missing pickle (mpc.py:80)
        raise FileNotFoundError(f'{guess_file} does not exist: run scripts/guess_acados.py -c {cont_name} --horizon {horizon} first, '
                                'or pass --generate-guess')
    x_init = x_guess[:, 0, :].copy()
    controller.setGuess(x_guess, u_guess)
    controller.reset_controller()

    # per-test perturbed plants (mpc.py:106-107) and the per-test torque-noise draw (mpc.py:126-127)
    # mpc.py:106-107 loads z1_randomizednoise<noise>_<i>.urdf of scripts/generate_urdf_noise.py for test i; when those files are absent the
    # same perturbation is drawn here (randomize_model's draw order on default_rng(0): the numbers of the first noise level of the script)
    names = [f'noise{args["noise"]}_{i}' for i in range(batch)]
    have_files = all(os.path.isfile(params.robot_urdf[:-5] + f'_randomized{n}.urdf') for n in names)
    print('perturbed plants:', 'read from robots/*_randomizednoise*.urdf' if have_files else f'drawn (uniform +-{args["noise"]} %, default_rng(0))')
    for m, c in ((model, controller), (model_backup, safe_ocp)):
        if have_files:
            m.update_randomized_dynamics(controller_name=names)
        else:
            m.update_randomized_dynamics(noise_percent=args['noise'], seed=0)
        m.reset_seed()
        c.ocp_solver.set_plant_inertial(m.plant_inertial)
        c.ocp_solver.set_torque_noise(m.torque_noise)

    print(f'robot: {"synthetic Z1-like chain" if params.synthetic_robot else params.robot_urdf}; viability network: '
          f'{"random-init stand-in (synthetic)" if getattr(params, "synthetic_net", False) else params.net_path if use_net else "none"}')
    sim = Sim(controller.ocp_solver, safe_ocp.ocp_solver, params.n_steps)
    sim.reset(x_init)
    t0 = time.perf_counter()
    stats = []                                                                                  # mpc.py:239: controller.getTime() of every step
    for _ in range(params.n_steps):
        sim.step()
        stats.append(sim.step_times())
    outcome = sim.outcome()
    dt = time.perf_counter() - t0
    x_log, u_log = sim.log()
    cnt = sim.counters()

    conv = (outcome & abi.OUT_CONVERGED) != 0
    coll = (outcome & abi.OUT_COLLIDED) != 0
    abrt = (outcome & abi.OUT_ABORTED) != 0
    conv_idx = np.where(conv)[0].tolist()                                                       # mpc.py:273-291
    collisions_idx = np.where(coll)[0].tolist()
    viable_idx = [i for i in np.where(abrt)[0].tolist() if i not in conv_idx and i not in collisions_idx]
    unconv_idx = [i for i in range(batch) if i not in conv_idx and i not in collisions_idx and i not in viable_idx]
    print('Completed task: ', len(conv_idx))
    print('Collisions: ', len(collisions_idx))
    print('Viable states: ', len(viable_idx))
    print('Not converged: ', len(unconv_idx))
    solves = cnt['rti_solves'] + cnt['backup_solves']
    print(f'{solves} RTI solves in {dt:.2f} s = {solves / dt:.0f} RTI iterations/s '
          f'({cnt["ipm_iterations"] / max(1, solves):.1f} IPM iterations per solve)')

    # mpc.py:300-303: 99 % quantile of the computation time per field -- here of one BATCHED controller step (all problems of the
    # batch at once, CUDA events), and the same divided by the number of problems that solved in that step
    times = np.array([t for t, n in stats])
    per_solve = np.array([t / max(1, n) for t, n in stats])
    print('99% quantile of the computation time (batched step | per solve):')
    for field, t, tp in zip(controller.time_fields, np.quantile(times, 0.99, axis=0), np.quantile(per_solve, 0.99, axis=0)):
        print(f'{field:<20} -> {t:.6f} | {tp:.3e}')

    xv = sim.x_viable()
    out = {'x': x_log, 'u': u_log, 'r': np.full((batch, params.n_steps, 1), np.nan), 'conv_idx': conv_idx,
           'collisions_idx': collisions_idx, 'unconv_idx': unconv_idx, 'viable_idx': viable_idx,
           'x_viable': xv[~np.isnan(xv).any(axis=1)]}
    os.makedirs(params.DATA_DIR, exist_ok=True)
    res_file = (f'{params.DATA_DIR}{model_name}_{cont_name}_use_net{use_net}_{horizon}hor_{int(params.alpha)}sm_noise_{args["noise"]}'
                f'_control_noise{args["control_noise"]}_q_collision_margins_{params.q_margin}_{params.collision_margin}_mpc.pkl')
    with open(res_file, 'wb') as f:                                                             # mpc.py:307-315
        pickle.dump(out, f)
    print('saved', res_file)
    return len(collisions_idx)


if __name__ == '__main__':
    sys.exit(min(main(), 255))
