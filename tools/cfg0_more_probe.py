"""configs[0] closed loop (100 problems x 800 steps, N = 45) GPU against oracle for the controllers the -m gpu test does not cover yet."""
import os, sys, traceback
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tests.test_gpu_baseline_configs import test_cfg0_closed_loop_outcomes_identical as t
for c in sys.argv[1:] or ['zerovel', 'stwa', 'real_receding', 'constraint_everywhere', 'parallel']:
    for fl, noise in (('halton', 0.0), ('shipped', 5.0)):
        try:
            t(c, fl, noise, 800); print('PASS', c, fl, noise, flush=True)
        except Exception as e:
            print('FAIL', c, fl, noise, repr(e)[:300], flush=True); traceback.print_exc()
