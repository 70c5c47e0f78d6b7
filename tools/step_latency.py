"""dev: per-step latency of the closed loop of bench.py next to the deepest IPM iteration count of the step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from safe_mpc_b200.engine import Engine, Sim
from safe_mpc_b200 import abi
B, steps = 10000, 26
params, md, x0, pin = bench.workload('st', 45, 0.0, 0, 0, B)
main, bk, prob = bench.make_handles(Engine, params, md, 'st', B, 0)
main.set_plant_inertial(pin)
bench.warm_guess(main, x0, 45, 5)
sim = Sim(main, bk, steps); sim.reset(x0)
for i in range(steps):
    main.sync(); t0 = time.perf_counter(); sim.step(); main.sync(); dt = time.perf_counter() - t0
    it = main.get_state(abi.STATE_QP_ITER); st = main.get_state(abi.STATE_STATUS)
    print(f'step {i:2d} {dt * 1e3:7.1f} ms  ipm mean {it.mean():5.1f} max {it.max():3d}  p99.9 {np.percentile(it, 99.9):.0f}  status!=0: {(st != 0).sum()}  n(it>40) {(it > 40).sum()}')
