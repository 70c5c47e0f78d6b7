"""Kernel timeline of ONE closed-loop step of the headline workload (SMPC_QP_TRACE=1: one CUDA-event pair per kernel, active / centering
counts per iteration on stderr).  usage: SMPC_QP_TRACE=1 python tools/trace_step.py [controller] [B] [steps before the traced one]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
from safe_mpc_b200.engine import Engine, Sim
ctrl = sys.argv[1] if len(sys.argv) > 1 else 'st'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
n0 = int(sys.argv[3]) if len(sys.argv) > 3 else 8
params, md, x0, pin = bench.workload(ctrl, 45, 0.0, 0, 0, B)
main, bk, prob = bench.make_handles(Engine, params, md, ctrl, B, 0)
main.set_plant_inertial(pin)
bench.warm_guess(main, x0, 45, 5)
sim = Sim(main, bk, n0 + 1)
sim.reset(x0)
sim.run(n0); main.sync()
sys.stderr.write('==== TRACED STEP ====\n'); sys.stderr.flush()
sim.step(); main.sync()
print('times', main.times())
