"""dev: the backup OCP (SafeBackupController) solved from hard viable states on the GPU and by the oracle: status / IPM iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.common import make_problem, random_states
from safe_mpc_b200.engine import Engine
from safe_mpc_b200 import abi
from oracle.oracle import Oracle
B, N = 512, 45
prob, params, md = make_problem('backup', cost='zero', N=N)
x = random_states(md, B, seed=77, vel_scale=float(sys.argv[1]) if len(sys.argv) > 1 else 0.6, shrink=0.7)
xg = np.repeat(x[:, None, :], N + 1, axis=1).copy(); ug = np.zeros((B, N, 5))
out = {}
for name, E in (('gpu', Engine), ('oracle', Oracle)):
    e = E(prob, B, 0)
    e.set_guess(xg, ug)
    st = e.rti_solve(x)
    out[name] = (np.asarray(st).copy(), e.get_state(abi.STATE_QP_ITER).copy(), e.get_state(abi.STATE_QP_STATUS).copy(), e.get_temp()[0].copy())
g, o = out['gpu'], out['oracle']
print('status equal', int((g[0] == o[0]).sum()), 'of', B, ' qp_status equal', int((g[2] == o[2]).sum()), ' qp_iter equal', int((g[1] == o[1]).sum()))
print('gpu    status counts', {int(k): int((g[0] == k).sum()) for k in np.unique(g[0])}, 'qp_status', {int(k): int((g[2] == k).sum()) for k in np.unique(g[2])}, 'iter mean', g[1].mean())
print('oracle status counts', {int(k): int((o[0] == k).sum()) for k in np.unique(o[0])}, 'qp_status', {int(k): int((o[2] == k).sum()) for k in np.unique(o[2])}, 'iter mean', o[1].mean())
bad = np.flatnonzero(g[2] != o[2])
for b in bad[:10]:
    print('problem', b, 'gpu (st, it, qst)', int(g[0][b]), int(g[1][b]), int(g[2][b]), ' oracle', int(o[0][b]), int(o[1][b]), int(o[2][b]))
ok = (g[0] == 0) & (o[0] == 0) & (g[2] == 0) & (o[2] == 0)
print('both converged:', int(ok.sum()), 'max |dx_temp|', float(np.abs(g[3][ok] - o[3][ok]).max()) if ok.any() else None)
