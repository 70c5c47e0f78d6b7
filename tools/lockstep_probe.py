#!/usr/bin/env python
"""dev (GPU): the closed loop on the engine and on the oracle side by side, one step at a time; for every problem whose outcome differs,
the first step at which the controllers disagree and what the solvers reported there.

    python tools/lockstep_probe.py <controller> <halton|shipped|stress> [noise] [B] [steps] [N]
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import bench
from safe_mpc_b200 import abi
from safe_mpc_b200.engine import Engine, Sim
from oracle.oracle import Oracle, OracleSim
from tests.common import make_problem, cfg0_initial_states, cfg0_plants, sqp_warm_start, outcome_sets

ctrl = sys.argv[1] if len(sys.argv) > 1 else 'receding'
flav = sys.argv[2] if len(sys.argv) > 2 else 'halton'
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
B = int(sys.argv[4]) if len(sys.argv) > 4 else 100
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 800
N = int(sys.argv[6]) if len(sys.argv) > 6 else 45
if flav == 'stress':
    params, md, x0, pin = bench.workload(ctrl, N, noise, 3, 0, B)
    x0[:, 5:] *= 3.0
    tn = np.zeros((B, abi.NU))
    prob, _, _ = make_problem(ctrl, N=N, noise=noise)
    bprob, _, _ = make_problem('backup', cost='zero', N=params.back_hor, noise=noise)
    iters = 3
else:
    prob, params, md = make_problem(ctrl, N=N, noise=noise)
    bprob, _, _ = make_problem('backup', cost='zero', N=params.back_hor, noise=noise)
    e0 = Engine(prob, B, 0)
    x0 = cfg0_initial_states(e0, md, params, B, flav)
    e0.close()
    pin, tn = cfg0_plants(md, params, B, noise, 0.0)
    iters = 10
e0 = Engine(prob, B, 0)
xg, ug = sqp_warm_start(e0, x0, N, iters)
e0.close()
sides = {}
for name, E, S in (('gpu', Engine, Sim), ('orc', Oracle, OracleSim)):
    main, bk = E(prob, B, 0), E(bprob, B, 0)
    main.set_guess(xg, ug); main.reset_controller()
    for h in (main, bk):
        h.set_plant_inertial(pin); h.set_torque_noise(tn)
    sim = S(main, bk, steps); sim.reset(x0)
    sides[name] = (main, bk, sim)
F = ('STATUS', 'QP_STATUS', 'QP_ITER', 'FAILS', 'R')
first = np.full(B, -1)
info = {}
for j in range(steps):
    rec = {}
    for name, (main, bk, sim) in sides.items():
        sim.step()
        c = sim.counters()
        rec[name] = dict(m={f: main.get_state(getattr(abi, 'STATE_' + f)).copy() for f in F}, b={f: bk.get_state(getattr(abi, 'STATE_' + f)).copy() for f in F[:3]},
                         mres=main.get_qp_residuals().copy(), bres=bk.get_qp_residuals().copy(), nb=c['backup_solves'])
    g, o = rec['gpu'], rec['orc']
    diff = np.zeros(B, dtype=bool)
    for f in ('STATUS', 'FAILS', 'R'):
        diff |= g['m'][f] != o['m'][f]
    if g['nb'] != o['nb'] or True:
        diff |= g['b']['STATUS'] != o['b']['STATUS']
    new = diff & (first < 0)
    for b in np.where(new)[0]:
        first[b] = j
        info[b] = {s: dict(main={f: int(rec[s]['m'][f][b]) for f in F}, main_res=['%.1e' % v for v in rec[s]['mres'][b]],
                           backup={f: int(rec[s]['b'][f][b]) for f in F[:3]}, backup_res=['%.1e' % v for v in rec[s]['bres'][b]]) for s in ('gpu', 'orc')}
og, oo = sides['gpu'][2].outcome(), sides['orc'][2].outcome()
print('gpu', outcome_sets(og), sides['gpu'][2].counters())
print('orc', outcome_sets(oo), sides['orc'][2].counters())
print('identical outcome codes', int((og == oo).sum()), 'of', B)
for b in np.where(og != oo)[0]:
    print(f'problem {b}: outcome gpu {og[b]} orc {oo[b]}; first controller difference at step {first[b]}')
    if b in info:
        for s in ('gpu', 'orc'):
            print('   ', s, info[b][s])
print('problems with a controller difference but the same outcome:', [int(b) for b in np.where((first >= 0) & (og == oo))[0]][:40])
