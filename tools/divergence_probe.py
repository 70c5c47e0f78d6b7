"""dev: where do the GPU and the oracle closed loops of scripts/long_run_check.py part ways?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from safe_mpc_b200.engine import Engine, Sim
from safe_mpc_b200 import abi
from oracle.oracle import Oracle, OracleSim
ctrl, B, steps, N = 'receding', 256, 150, 20
params, md, x0, pin = bench.workload(ctrl, N, 5.0, 3, 0, B)
x0[:, 5:] *= 3.0
sims = {}
for name, E, S in (('gpu', Engine, Sim), ('oracle', Oracle, OracleSim)):
    main, bk, prob = bench.make_handles(E, params, md, ctrl, B, 0)
    main.set_plant_inertial(pin)
    bench.warm_guess(main, x0, N, 3)
    sim = S(main, bk, steps); sim.reset(x0)
    sims[name] = (main, bk, sim)
first = np.full(B, -1)
info = {}
for j in range(steps):
    st = {}
    for name, (main, bk, sim) in sims.items():
        sim.step()
        x, u = sim.log()
        st[name] = (np.array(x[:, j + 1]), np.array(u[:, j]), main.get_state(abi.STATE_STATUS).copy(), main.get_state(abi.STATE_FAILS).copy(),
                    main.get_state(abi.STATE_R).copy(), main.get_state(abi.STATE_QP_ITER).copy(), bk.get_state(abi.STATE_STATUS).copy(),
                    bk.get_state(abi.STATE_QP_ITER).copy(), bk.get_state(abi.STATE_QP_STATUS).copy())
    dx = np.abs(np.nan_to_num(st['gpu'][0]) - np.nan_to_num(st['oracle'][0])).max(axis=1)
    nanmis = (np.isnan(st['gpu'][0]).any(axis=1) != np.isnan(st['oracle'][0]).any(axis=1))
    for b in np.flatnonzero(((dx > 1e-6) | nanmis) & (first < 0)):
        first[b] = j
        info[b] = (j, dx[b], [(k, int(st[k][2][b]), int(st[k][3][b]), int(st[k][4][b]), int(st[k][5][b]), 'bk status/iter/qpstatus', int(st[k][6][b]), int(st[k][7][b]), int(st[k][8][b])) for k in st], np.abs(np.nan_to_num(st['gpu'][1][b]) - np.nan_to_num(st['oracle'][1][b])).max())
    if j and j % 50 == 0:
        print('step', j, 'diverged so far', int((first >= 0).sum()))
print('diverged problems', int((first >= 0).sum()))
for b, v in list(info.items())[:12]:
    print('problem', b, 'first divergence at step', v[0], 'dx', f'{v[1]:.2e}', 'du', f'{v[3]:.2e}', '(impl, status, fails, r, qp_iter, backup status):', v[2])
