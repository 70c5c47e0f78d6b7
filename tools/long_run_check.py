"""dev: a long closed loop (aborts, backup solves, terminations) on the GPU against the oracle: outcome codes must be identical."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from safe_mpc_b200.engine import Engine, Sim
from safe_mpc_b200 import distributed as D
from oracle.oracle import Oracle, OracleSim
ctrl = sys.argv[1] if len(sys.argv) > 1 else 'receding'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
N = 20
params, md, x0, pin = bench.workload(ctrl, N, 5.0, 3, 0, B)
x0[:, 5:] *= 3.0                                          # faster initial velocities: more aborts
res = {}
for name, E, S in (('gpu', Engine, Sim), ('oracle', Oracle, OracleSim)):
    main, bk, prob = bench.make_handles(E, params, md, ctrl, B, 0)
    main.set_plant_inertial(pin)
    bench.warm_guess(main, x0, N, 3)
    sim = S(main, bk, steps); sim.reset(x0)
    t0 = time.perf_counter(); sim.run(steps); 
    out = np.asarray(sim.outcome()); dt = time.perf_counter() - t0
    res[name] = out
    print(name, D.outcome_counts(out), sim.counters(), f'{dt:.1f} s')
same = (res['gpu'] == res['oracle'])
print('identical outcome codes:', int(same.sum()), 'of', B)
