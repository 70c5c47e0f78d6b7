"""dev: accuracy and throughput of the two viability-network kernels (strict fp64-accumulate vs tcgen05 tf32x3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.common import make_problem, random_states
from safe_mpc_b200.engine import Engine
from oracle.oracle import Oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 460000
prob_tc, params, md = make_problem('st', N=20, nn_precision='tf32x3')
prob_st, _, _ = make_problem('st', N=20)
e_tc, e_st = Engine(prob_tc, 64, 0), Engine(prob_st, 64, 0)
x = random_states(md, n, seed=1)
xd = torch.tensor(x, device='cuda')
res = {}
for name, e in (('strict', e_st), ('tf32x3', e_tc)):
    l0 = e.launch_count()
    c, g = e.nn_constraint(xd); e.sync()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); c, g = e.nn_constraint(xd); e.sync(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    res[name] = (c.cpu().numpy(), g.cpu().numpy())
    ms = min(ts) * 1e3
    print(f'{name}: {ms:.3f} ms for {n} rows  ({n * 535040 / ms / 1e9:.2f} algorithmic TFLOP/s), launches {e.launch_count() - l0}')
m = min(n, 4000)
orc = Oracle(prob_st, 64, 0)
c_o, g_o = orc.nn_constraint(x[:m])
for name in res:
    c, g = res[name]
    print(f'{name} vs oracle: max |dc| {np.abs(c[:m] - c_o).max():.3e} (|c| max {np.abs(c_o).max():.3f})  max |dgrad| {np.abs(g[:m] - g_o).max():.3e} (|g| max {np.abs(g_o).max():.3f})')
