"""Linearisation kernel alone: time per launch (CUDA events through smpc_get_times) and agreement of the forms (GPU)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from safe_mpc_b200.engine import Engine
from safe_mpc_b200.problem import build_problem

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
params, md, x0, pin = bench.workload('st', 45, 0.0, 0, 0, B)
out = {}
for form in ('coop', 'thread'):
    if form == 'thread': os.environ['SMPC_LIN'] = 'thread'
    else: os.environ.pop('SMPC_LIN', None)
    os.environ['SMPC_QP_COMPACT'] = '0'
    prob, keep = build_problem(params, 'naive', cost='ext', model=md)
    eng = Engine(prob, B, 0)
    xg, ug = bench.rollout(x0, 45, params.dt) if hasattr(bench, 'rollout') else (None, None)
    if xg is None:
        from tests.common import rollout_guess
        xg, ug = rollout_guess(x0, 45, params.dt, seed=3)
    eng.set_guess(xg, ug)
    ts = []
    for i in range(5):
        eng.rti_solve(x0); eng.sync(); ts.append(eng.times()['time_lin'] * 1e3)
    lin = eng.get_lin()
    out[form] = lin
    print(form, 'time_lin ms', [round(t, 3) for t in ts], flush=True)
    eng.close()
d = np.abs(out['coop'] - out['thread'])
print('max abs diff coop vs thread', d.max(), 'differing entries', int((out['coop'] != out['thread']).sum()), 'of', d.size)
