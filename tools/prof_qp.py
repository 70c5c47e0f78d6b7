import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench
from safe_mpc_b200.engine import Engine
ctrl=sys.argv[1] if len(sys.argv)>1 else 'st'
B=int(sys.argv[2]) if len(sys.argv)>2 else 10000
params, md, x0, pin = bench.workload(ctrl, 45, 0.0, 0, 0, B)
main, bk, prob = bench.make_handles(Engine, params, md, ctrl, B, 0)
bench.warm_guess(main, x0, 45, 3)
for i in range(2):
    st = main.rti_solve(x0)
print('times', main.times(), 'iters', main.get_state(3).mean())
