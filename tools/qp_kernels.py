"""dev: per-kernel times of one batched QP solve (CUDA events through smpc_set_profiling) + the solve time with tile groups."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from safe_mpc_b200.engine import Engine
from safe_mpc_b200 import abi
ctrl = sys.argv[1] if len(sys.argv) > 1 else 'st'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
params, md, x0, pin = bench.workload(ctrl, 45, 0.0, 0, 0, B)
main, bk, prob = bench.make_handles(Engine, params, md, ctrl, B, 0)
bench.warm_guess(main, x0, 45, 5)
x = x0.copy()
for _ in range(steps):
    u, ab = main.controller_step(x); x, _ = main.plant_step(x, u)
xs = torch.tensor(x, device='cuda')
tq = []
for _ in range(4):
    main.rti_solve(xs); main.sync(); tq.append(main.times())
os.environ['SMPC_QP_GROUPS'] = '1'
probe, _bk, _ = bench.make_handles(Engine, params, md, ctrl, B, 0)
probe.set_guess(*main.get_guess())
probe.rti_solve(xs); probe.sync()
probe.set_profiling(True)
probe.rti_solve(xs); probe.sync()
kern, span, itmax = probe.profile()
t1 = probe.times()
it = main.get_state(abi.STATE_QP_ITER)
print('LIB', os.environ.get('SMPC_LIB', 'default'), 'ctrl', ctrl, 'B', B)
print('solve ms (groups) %.2f  (1 group) %.2f  lin %.2f  iters mean %.1f max %d' % (np.median([t['time_qp'] for t in tq]) * 1e3, t1['time_qp'] * 1e3,
      np.median([t['time_lin'] for t in tq]) * 1e3, it.mean(), it.max()))
print('  '.join('%s=%.2f/%d' % (k, v[0], v[1]) for k, v in kern.items()))
