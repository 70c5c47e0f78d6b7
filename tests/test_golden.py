"""Golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle): the oracle must keep
reproducing them (CPU), and the CUDA path must match them through the C ABI (gpu)."""
import glob
import os

import numpy as np
import pytest

from safe_mpc_b200 import abi
from tests.common import make_problem

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'rti_*.npz')))


def _case(path):
    g = np.load(path)
    name = os.path.basename(path)[4:-4]
    prob, params, md = make_problem(name, cost=str(g['cost']), N=int(g['N']))
    return name, g, prob


def _run(e, g):
    x0, xg, ug = g['x0'], g['xg'], g['ug']
    e.set_guess(xg, ug); e.reset_controller()
    status = e.rti_solve(x0)
    xt, ut = e.get_temp()
    it = e.get_state(abi.STATE_QP_ITER)
    e.set_guess(xg, ug); e.reset_controller()
    x = x0.copy()
    us, xs, fails, rs = [], [], [], []
    for _ in range(len(g['u_steps'])):
        u, ab = e.controller_step(x)
        x, _ = e.plant_step(x, u)
        us.append(u); xs.append(x); fails.append(e.get_state(abi.STATE_FAILS)); rs.append(e.get_state(abi.STATE_R))
    return status, xt, ut, it, np.array(us), np.array(xs), np.array(fails), np.array(rs)


def _check(got, g, rtol):
    status, xt, ut, it, us, xs, fails, rs = got
    assert np.array_equal(status, g['status']) and np.array_equal(it, g['qp_iter'])
    assert np.array_equal(fails, g['fails']) and np.array_equal(rs, g['r'])
    for a, b in ((xt, g['x_temp']), (ut, g['u_temp']), (us, g['u_steps']), (xs, g['x_steps'])):
        assert np.abs(a - b).max() <= rtol * max(1.0, np.abs(b).max())


def test_golden_files_exist():
    assert len(FILES) >= 6


@pytest.mark.parametrize('path', FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(path):
    from oracle.oracle import Oracle
    name, g, prob = _case(path)
    _check(_run(Oracle(prob, len(g['x0']), 1), g), g, 1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize('path', FILES, ids=[os.path.basename(f) for f in FILES])
def test_engine_matches_golden(path):
    from safe_mpc_b200.engine import Engine
    name, g, prob = _case(path)
    _check(_run(Engine(prob, len(g['x0']), 0), g), g, 1e-6)      # 1e-6 relative: north_star fp64 tolerance
