"""Which kernel form of the fp32-storage flavour departs from the thread-per-stage / one-warp baseline, and by how much (GPU)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_mpc_b200 import abi
from tests.common import make_problem, start_states, rollout_guess


def solve(controller, env, precision='f32', Bc=1280, Nc=16):
    from safe_mpc_b200.engine import Engine
    env = {'SMPC_QP_SOLO': '0', **env}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        prob, params, md = make_problem(controller, N=Nc, precision=precision)
        eng = Engine(prob, Bc, 0)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    x0 = start_states(Bc, seed=23, vel=0.5)
    x0[::3, 5:] *= 4.0
    xg, ug = rollout_guess(x0, Nc, params.dt, seed=24, scale=1.0)
    eng.set_guess(xg, ug)
    st = eng.rti_solve(x0)
    xt, ut = eng.get_temp()
    it = eng.get_state(abi.STATE_QP_ITER)
    eng.close()
    return st, xt, ut, it


for controller in ('st',):
    base = solve(controller, {'SMPC_QP_TAIL': '0', 'SMPC_QP_PREP': 'thread', 'SMPC_QP_RIC1': 'single', 'SMPC_QP_COMPACT': '0'})
    for env in ({'SMPC_QP_TAIL': '0', 'SMPC_QP_PREP': 'thread', 'SMPC_QP_COMPACT': '0'},               # two-warp ric1
                {'SMPC_QP_TAIL': '0', 'SMPC_QP_RIC1': 'single', 'SMPC_QP_COMPACT': '0'},              # cooperative prep
                {'SMPC_QP_TAIL': '100000', 'SMPC_QP_PREP': 'thread', 'SMPC_QP_COMPACT': '0'},         # warp-per-problem sweeps
                {'SMPC_QP_TAIL': '100000', 'SMPC_QP_TAIL_WHICH': '1', 'SMPC_QP_PREP': 'thread', 'SMPC_QP_RIC1': 'single', 'SMPC_QP_COMPACT': '0'},
                {'SMPC_QP_TAIL': '100000', 'SMPC_QP_TAIL_WHICH': '2', 'SMPC_QP_PREP': 'thread', 'SMPC_QP_RIC1': 'single', 'SMPC_QP_COMPACT': '0'},
                {'SMPC_QP_TAIL': '0', 'SMPC_QP_PREP': 'thread', 'SMPC_QP_RIC1': 'single', 'SMPC_QP_COMPACT': '1'},
                {'SMPC_QP_SOLO': '100000', 'SMPC_QP_COMPACT': '0'}):
        o = solve(controller, env)
        nd = [int((base[i] != o[i]).sum()) for i in range(4)]
        dx = np.abs(base[1] - o[1]).max()
        print(controller, env, 'differing entries st/xt/ut/it', nd, 'max |dxt|', dx, 'iters differ', int((base[3] != o[3]).sum()), flush=True)
