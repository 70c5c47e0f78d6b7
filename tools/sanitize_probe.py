"""Small solves through every QP path (solo kernel, multi-kernel with warp-per-problem sweeps, bulk forms) and both linearisation kernels,
for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import make_problem, start_states, rollout_guess


def run(env, controller='st', B=70, N=12, precision=None, nn_precision=None):
    from safe_mpc_b200.engine import Engine
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        prob, params, md = make_problem(controller, N=N, precision=precision, nn_precision=nn_precision)
        eng = Engine(prob, B, 0)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    x0 = start_states(B, seed=5, vel=0.4)
    xg, ug = rollout_guess(x0, N, params.dt, seed=6, scale=1.0)
    eng.set_guess(xg, ug)
    traj = np.asarray(params.ee_ref) + 0.01 * np.arange(3 * (N + 3)).reshape(N + 3, 3) / (N + 3)
    eng.set_ee_trajectory(traj)
    st = eng.rti_solve(x0)
    u, ab = eng.controller_step(x0)
    eng.close()
    print(env, controller, precision, 'status', np.bincount(st, minlength=5).tolist(), flush=True)


run({'SMPC_QP_SOLO': '384'})
run({'SMPC_QP_SOLO': '0', 'SMPC_QP_TAIL': '100000'})
run({'SMPC_QP_SOLO': '0', 'SMPC_QP_TAIL': '0'})
run({'SMPC_QP_SOLO': '0', 'SMPC_QP_TAIL': '0', 'SMPC_LIN': 'coop'}, controller='receding')
run({'SMPC_QP_SOLO': '384'}, precision='f32')
run({'SMPC_QP_SOLO': '20', 'SMPC_QP_SOLO_TAIL': '1', 'SMPC_QP_TAIL': '40'}, B=130)
run({'SMPC_QP_SOLO': '384'}, controller='constraint_everywhere', nn_precision='tf32x3')
run({'SMPC_QP_SOLO': '384', 'SMPC_MLP_TC': 'pair'}, controller='receding', nn_precision='tf32x3')
# round 2, later additions: slot compaction with the lane = move kernel (40 tiles, problems of very different iteration counts), the
# centering switch per flagged problem (list form forced / never), cooperative prep with rolled row loops and bounds in shared memory
run({'SMPC_QP_SOLO': '0', 'SMPC_QP_REDO_LIST': '100000'}, B=1280, N=10)
run({'SMPC_QP_SOLO': '0', 'SMPC_QP_REDO_LIST': '0', 'SMPC_QP_TAIL': '0'}, controller='real_receding', B=1280, N=10)
