"""CUDA torque-input RK4 step with sensitivities (smpc_rk4_sens) against the oracle's forward-mode AD
through the C ABI: 1e-10 relative on x_next, A = d x_next / d x and B = d x_next / d tau; host and device buffers; a row count that
is not a multiple of the CTA size; the semigroup-free size-independent property A B consistency at a large row count."""
import numpy as np
import pytest

from tests.common import make_problem, random_states

pytestmark = pytest.mark.gpu


def _pair(B=8):
    from safe_mpc_b200.engine import Engine
    from oracle.oracle import Oracle
    prob, params, md = make_problem('st')
    return Engine(prob, B, 0), Oracle(prob, B, 0), params, md


def _rel(a, b):
    return float((np.abs(np.asarray(a) - b) / np.maximum(1.0, np.abs(b))).max())


def test_rk4_sens_matches_oracle():
    eng, orc, params, md = _pair()
    n = 4133
    x = random_states(md, n, seed=41, vel_scale=0.6)
    tau = np.random.default_rng(42).uniform(-8, 8, (n, 5))
    for dt in (params.dt, 0.02):
        xg, Ag, Bg = eng.rk4_sens(x, tau, dt)
        xo, Ao, Bo = orc.rk4_sens(x, tau, dt)
        assert _rel(xg, xo) < 1e-10 and _rel(Ag, Ao) < 1e-10 and _rel(Bg, Bo) < 1e-10
    # the value-only kernel (no tangent passes) gives the same x_next
    assert _rel(eng.rk4_sens(x, tau, 0.02, sens=False), xg) < 1e-13


def test_rk4_sens_device_buffers_and_errors():
    import torch
    eng, orc, params, md = _pair()
    n = 777
    x = random_states(md, n, seed=43, vel_scale=0.6)
    tau = np.random.default_rng(44).uniform(-8, 8, (n, 5))
    xh, Ah, Bh = eng.rk4_sens(x, tau, params.dt)
    xd, Ad, Bd = eng.rk4_sens(torch.from_numpy(x).cuda(), torch.from_numpy(tau).cuda(), params.dt)
    eng.sync()
    np.testing.assert_array_equal(xd.cpu().numpy(), xh)
    np.testing.assert_array_equal(Ad.cpu().numpy(), Ah)
    np.testing.assert_array_equal(Bd.cpu().numpy(), Bh)
    with pytest.raises(RuntimeError):
        eng.rk4_sens(x, tau, 0.0)


def test_rk4_sens_full_size_property():
    """460 000 rows (cfg[1]: 10 000 problems x 46 stages): the sensitivities predict the effect of a small perturbation of (x, tau)
    on x_next to second order -- checked on every row without the oracle."""
    eng, orc, params, md = _pair()
    n = 460000
    rng = np.random.default_rng(45)
    x = random_states(md, n, seed=46, vel_scale=0.6)
    tau = rng.uniform(-8, 8, (n, 5))
    dx = 1e-6 * rng.uniform(-1, 1, (n, 10)); dtau = 1e-6 * rng.uniform(-1, 1, (n, 5))
    dt = params.dt
    xn, A, B = eng.rk4_sens(x, tau, dt)
    xp = eng.rk4_sens(x + dx, tau + dtau, dt, sens=False)
    pred = xn + np.einsum('nij,nj->ni', A, dx) + np.einsum('nij,nj->ni', B, dtau)
    assert np.isfinite(xn).all()
    assert np.abs(xp - pred).max() < 2e-8          # second-order remainder: 6e-10 on a 40 000-row sample (host build of the same source)
    # a sample of the rows against the oracle
    idx = rng.choice(n, 2000, replace=False)
    xo, Ao, Bo = orc.rk4_sens(x[idx], tau[idx], dt)
    assert _rel(xn[idx], xo) < 1e-10 and _rel(A[idx], Ao) < 1e-10 and _rel(B[idx], Bo) < 1e-10


def test_rk4_sens_matches_the_golden_vectors():
    """tests/golden/rk4_sens.npz (oracle forward-mode AD, tests/golden/make_golden.py rk4)"""
    import os
    eng, orc, params, md = _pair()
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rk4_sens.npz'))
    for i, dt in enumerate(g['dt']):
        xn, A, B = eng.rk4_sens(g['x'], g['tau'], float(dt))
        assert _rel(xn, g['x_next'][i]) < 1e-10 and _rel(A, g['A'][i]) < 1e-10 and _rel(B, g['B'][i]) < 1e-10
