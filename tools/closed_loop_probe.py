#!/usr/bin/env python
"""dev (CPU only): BASELINE configs[0] closed loops on the oracle and on the oracle with the engine's kernel-source QP
(tests/emu): outcome codes, first diverging step, and whether the oracle's sensitivity probe flags the problem.

    python scripts/closed_loop_probe.py <controller> <flavour: shipped|halton> <noise> [B] [steps] [N] [sqp_iters]
"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from tests.common import make_problem, cfg0_initial_states, cfg0_plants, sqp_warm_start, run_closed_loop, outcome_sets
from tests import emu
from oracle.oracle import Oracle, OracleSim

ctrl = sys.argv[1] if len(sys.argv) > 1 else 'st'
flav = sys.argv[2] if len(sys.argv) > 2 else 'halton'
noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
B = int(sys.argv[4]) if len(sys.argv) > 4 else 100
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 800
N = int(sys.argv[6]) if len(sys.argv) > 6 else 45
sqp = int(sys.argv[7]) if len(sys.argv) > 7 else 10
cn = 1.0 if noise > 0 else 0.0
prob, params, md = make_problem(ctrl, N=N, noise=noise, control_noise=cn)
bprob, _, _ = make_problem('backup', cost='zero', N=params.back_hor, noise=noise, control_noise=cn)
o = Oracle(prob, B, 0)
x0 = cfg0_initial_states(o, md, params, B, flav)
pin, tn = cfg0_plants(md, params, B, noise, cn)
t0 = time.time()
xg, ug = sqp_warm_start(o, x0, N, sqp)
print(f'warm start {time.time() - t0:.1f} s', flush=True)
o.close()
t0 = time.time()
ro = run_closed_loop(Oracle, OracleSim, prob, bprob, x0, xg, ug, pin, tn, steps, probe_eps=1e-11)
print('oracle', outcome_sets(ro['outcome']), ro['counters'], f'{time.time() - t0:.1f} s', 'flagged problems', int((ro['flips'] > 0).sum()), flush=True)
t0 = time.time()
re_ = run_closed_loop(Oracle, OracleSim, prob, bprob, x0, xg, ug, pin, tn, steps, hook=emu.load().emu_qp_solve1)
print('kernel-source', outcome_sets(re_['outcome']), re_['counters'], f'{time.time() - t0:.1f} s', flush=True)
same = ro['outcome'] == re_['outcome']
print('identical outcome codes:', int(same.sum()), 'of', B)
xo, xe = np.nan_to_num(ro['x']), np.nan_to_num(re_['x'])
err = np.abs(xo - xe).max(axis=2)
nanmis = (np.isnan(ro['x']) != np.isnan(re_['x'])).any(axis=2)
for b in range(B):
    bad = np.where((err[b] > 1e-6 * max(1.0, np.abs(xo[b]).max())) | nanmis[b])[0]
    if len(bad) or not same[b]:
        print(f'  problem {b}: outcome {ro["outcome"][b]} / {re_["outcome"][b]}, first differing step {bad[0] if len(bad) else None}, flips {ro["flips"][b]}')
print('max |x_oracle - x_kernel| over problems that agree:', float(err[same & ~nanmis.any(axis=1)].max()) if (same & ~nanmis.any(axis=1)).any() else None)
