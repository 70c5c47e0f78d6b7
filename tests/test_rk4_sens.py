"""Torque-input RK4 step with sensitivities (an extension: the reference's OCP dynamics are the
double integrator, env_model.py:58-71, so there is no reference output to hold this to).  CPU tests: the oracle's forward-mode AD
against an independent numpy composition (its own mass matrix / bias, numpy solve, numpy RK4) and central differences; the
engine's device source (dev_model.cuh: fd_sens / rk4_sens, analytic inverse-dynamics tangents) compiled for the host against
the oracle.  Tolerances: 1e-10 relative between the two implementations, 1e-6 against finite differences."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.common import make_problem, random_states


@pytest.fixture(scope='module')
def setup():
    prob, params, md = make_problem('st')
    return prob, params, md, Oracle(prob, 4, 2)


def _inputs(md, n, seed):
    x = random_states(md, n, seed=seed, vel_scale=0.6)
    tau = np.random.default_rng(seed + 1).uniform(-8, 8, (n, 5))
    return x, tau


def _rk4_numpy(o, x, tau, dt):
    def f(xx):
        M, h = o.mass_bias(0, xx)
        return np.concatenate([xx[5:], np.linalg.solve(M, tau - h)])
    k1 = f(x); k2 = f(x + 0.5 * dt * k1); k3 = f(x + 0.5 * dt * k2); k4 = f(x + dt * k3)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def test_oracle_value_matches_numpy_composition(setup):
    prob, params, md, o = setup
    x, tau = _inputs(md, 12, 3)
    for dt in (params.dt, 0.02):
        xn = o.rk4_sens(x, tau, dt, sens=False)
        ref = np.array([_rk4_numpy(o, x[i], tau[i], dt) for i in range(len(x))])
        assert np.abs(xn - ref).max() < 1e-11


def test_oracle_sensitivities_match_central_differences(setup):
    prob, params, md, o = setup
    x, tau = _inputs(md, 8, 5)
    dt, e = 0.01, 1e-5
    xn, A, B = o.rk4_sens(x, tau, dt)
    for j in range(10):
        d = np.zeros(10); d[j] = e
        fd = (o.rk4_sens(x + d, tau, dt, sens=False) - o.rk4_sens(x - d, tau, dt, sens=False)) / (2 * e)
        assert np.abs(fd - A[:, :, j]).max() < 1e-6 * max(1.0, np.abs(A).max())
    for j in range(5):
        d = np.zeros(5); d[j] = e
        fd = (o.rk4_sens(x, tau + d, dt, sens=False) - o.rk4_sens(x, tau - d, dt, sens=False)) / (2 * e)
        assert np.abs(fd - B[:, :, j]).max() < 1e-6 * max(1.0, np.abs(B).max())


def test_first_slope_inverts_the_torque_model(setup):
    """tau = tau_fun(x, u) (env_model.py:80-83) fed back gives the acceleration u: (x_next - x) / dt -> [v; u] as dt -> 0"""
    prob, params, md, o = setup
    x, _ = _inputs(md, 16, 7)
    u = np.random.default_rng(8).uniform(-3, 3, (16, 5))
    tau = o.tau(x, u)
    dt = 1e-6
    xn = o.rk4_sens(x, tau, dt, sens=False)
    slope = (xn - x) / dt
    assert np.abs(slope[:, :5] - x[:, 5:]).max() < 1e-4
    assert np.abs(slope[:, 5:] - u).max() < 1e-3


def test_device_source_matches_oracle_ad(setup):
    from tests.emu import load
    emu = load()
    prob, params, md, o = setup
    x, tau = _inputs(md, 96, 11)
    p = lambda a: C.c_void_p(a.ctypes.data)
    for dt in (params.dt, 0.02):
        xn, A, B = o.rk4_sens(x, tau, dt)
        xe = np.zeros_like(xn); Ae = np.zeros_like(A); Be = np.zeros_like(B)
        emu.emu_rk4_sens(C.byref(prob), len(x), p(x), p(tau), C.c_double(dt), p(xe), p(Ae), p(Be))
        for got, want in ((xe, xn), (Ae, A), (Be, B)):
            assert (np.abs(got - want) / np.maximum(1.0, np.abs(want))).max() < 1e-10
        xv = np.zeros_like(xn)                       # value-only form (no tangent passes)
        emu.emu_rk4_sens(C.byref(prob), len(x), p(x), p(tau), C.c_double(dt), p(xv), None, None)
        assert (np.abs(xv - xe) / np.maximum(1.0, np.abs(xe))).max() < 1e-13


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rk4_sens.npz'))


def test_oracle_and_device_source_reproduce_the_golden_vectors(setup):
    """tests/golden/rk4_sens.npz (tests/golden/make_golden.py rk4): regression pin of the oracle, fixture of the kernel source"""
    from tests.emu import load
    emu = load()
    prob, params, md, o = setup
    g = _golden()
    x, tau = g['x'], g['tau']
    p = lambda a: C.c_void_p(a.ctypes.data)
    for i, dt in enumerate(g['dt']):
        xn, A, B = o.rk4_sens(x, tau, float(dt))
        xe = np.zeros_like(xn); Ae = np.zeros_like(A); Be = np.zeros_like(B)
        emu.emu_rk4_sens(C.byref(prob), len(x), p(x), p(tau), C.c_double(float(dt)), p(xe), p(Ae), p(Be))
        for got, emu_got, want in ((xn, xe, g['x_next'][i]), (A, Ae, g['A'][i]), (B, Be, g['B'][i])):
            assert (np.abs(got - want) / np.maximum(1.0, np.abs(want))).max() < 1e-12
            assert (np.abs(emu_got - want) / np.maximum(1.0, np.abs(want))).max() < 1e-10
