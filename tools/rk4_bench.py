"""Timing of the torque-input RK4 sensitivity kernel (smpc_rk4_sens) on one GPU: 460 000 rows = the (problem, stage) pairs of cfg[1],
device-resident buffers, CUDA events on torch's stream around the stream-ordered call.  Prints one JSON line (not the bench.py headline)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from safe_mpc_b200.engine import Engine          # noqa: E402
from tests.common import make_problem, random_states   # noqa: E402

prob, params, md = make_problem('st')
eng = Engine(prob, 8, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 460000
x = torch.from_numpy(random_states(md, n, seed=1, vel_scale=0.6)).cuda()
tau = torch.from_numpy(np.random.default_rng(2).uniform(-8, 8, (n, 5))).cuda()
for sens in (True, False):
    for _ in range(3):
        out = eng.rk4_sens(x, tau, params.dt, sens=sens)
    eng.sync(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()                                   # (Engine._call orders the engine's stream after / before torch's current stream)
        out = eng.rk4_sens(x, tau, params.dt, sens=sens)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(json.dumps({'kernel': 'rk4_sens_kernel', 'sensitivities': sens, 'rows': n, 'ms': ms, 'rows_per_s': n / (ms * 1e-3),
                      'bytes_written': n * 8 * (10 + (150 if sens else 0)), 'hbm_gbs_written': n * 8 * (10 + (150 if sens else 0)) / (ms * 1e-3) / 1e9}))
