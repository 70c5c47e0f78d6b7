"""Pins the oracle's model functions (rigid-body dynamics, kinematics, capsule distances, viability network)
against independent implementations: a numpy chain RNEA, torch fp64 autograd, finite differences, identities.
The reference holds no golden vectors for this path -- these cross-checks are the pin."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from safe_mpc_b200 import abi, robot_model
from safe_mpc_b200.problem import load_network
from tests.common import make_problem, random_states, start_states, constant_guess


@pytest.fixture(scope='module')
def orc():
    prob, params, md = make_problem('htwa')
    return Oracle(prob, 8, 2), prob, params, md


def test_rnea_matches_numpy_chain(orc):
    o, prob, params, md = orc
    x = random_states(md, 16, seed=1)
    u = np.random.default_rng(2).uniform(-5, 5, (16, 5))
    tau = o.tau(x, u)
    for i in range(16):
        ref = robot_model.rnea(md.chain, md.inertial, x[i, :5], x[i, 5:], u[i])
        np.testing.assert_allclose(tau[i], ref, rtol=1e-11, atol=1e-11)


def test_mass_matrix_identities(orc):
    o, prob, params, md = orc
    x = random_states(md, 4, seed=3)
    for i in range(4):
        M, h = o.mass_bias(0, x[i])
        np.testing.assert_allclose(M, M.T, atol=1e-12)
        assert np.linalg.eigvalsh(M).min() > 0
        u = np.random.default_rng(i).uniform(-3, 3, 5)
        np.testing.assert_allclose(o.tau(x[i:i + 1], u[None])[0], M @ u + h, rtol=1e-10, atol=1e-10)
    # gravity only at rest: potential-energy gradient by finite differences
    q = x[0, :5]
    def potential(qq):
        Rs, os_ = robot_model.fk_frames(md.chain, qq)
        return sum(md.inertial[b, 0] * robot_model.GRAVITY * (os_[b] + Rs[b] @ md.inertial[b, 1:4])[2] for b in range(5))
    g_fd = np.array([(potential(q + 1e-6 * e) - potential(q - 1e-6 * e)) / 2e-6 for e in np.eye(5)])
    x_rest = np.hstack([q, np.zeros(5)])
    np.testing.assert_allclose(o.tau(x_rest[None], np.zeros((1, 5)))[0], g_fd, rtol=1e-6, atol=1e-7)


def test_euler_lagrange_derivation_of_mass_matrix_and_coriolis(orc):
    """a2 from the energy side: M(q) is the Hessian of the kinetic energy T = 1/2 sum_b (m |v_cb|^2 + w_b' R I R' w_b) in the joint
    velocities and the velocity-dependent part of h(q, qd) follows from the Christoffel symbols of M (Euler-Lagrange) -- no
    Newton-Euler recursion anywhere in this derivation, only forward kinematics differentiated numerically."""
    o, prob, params, md = orc
    ch, inert, n = md.chain, md.inertial, 5

    def mass(q, h=1e-6):
        Rs0, _ = robot_model.fk_frames(ch, q)
        Jv, Jw = np.zeros((n, 3, n)), np.zeros((n, 3, n))
        for i in range(n):
            e = np.zeros(n); e[i] = h
            Rp, op = robot_model.fk_frames(ch, q + e); Rm, om = robot_model.fk_frames(ch, q - e)
            for b in range(n):
                c = inert[b, 1:4]
                Jv[b, :, i] = ((op[b] + Rp[b] @ c) - (om[b] + Rm[b] @ c)) / (2 * h)
                S = (Rp[b] - Rm[b]) / (2 * h) @ Rs0[b].T                 # skew of the angular velocity per unit joint rate
                Jw[b, :, i] = [S[2, 1], S[0, 2], S[1, 0]]
        M = np.zeros((n, n))
        for b in range(n):
            I = np.array([[inert[b, 4], inert[b, 7], inert[b, 9]], [inert[b, 7], inert[b, 5], inert[b, 8]], [inert[b, 9], inert[b, 8], inert[b, 6]]])
            M += inert[b, 0] * Jv[b].T @ Jv[b] + Jw[b].T @ (Rs0[b] @ I @ Rs0[b].T) @ Jw[b]
        return M

    x = random_states(md, 3, seed=11, vel_scale=0.8)
    for i in range(3):
        q, qd = x[i, :5], x[i, 5:]
        M_o, h_o = o.mass_bias(0, x[i])
        np.testing.assert_allclose(mass(q), M_o, rtol=1e-7, atol=1e-8)
        d = 1e-4
        dM = np.stack([(mass(q + d * e) - mass(q - d * e)) / (2 * d) for e in np.eye(n)], axis=2)      # dM[i, j, k] = dM_ij / dq_k
        cor = np.einsum('ijk,j,k->i', dM, qd, qd) - 0.5 * np.einsum('jki,j,k->i', dM, qd, qd)
        _, g_o = o.mass_bias(0, np.hstack([q, np.zeros(5)]))
        np.testing.assert_allclose(h_o - g_o, cor, rtol=1e-5, atol=1e-6 * max(1.0, np.abs(cor).max()))


def test_fk_and_distances(orc):
    o, prob, params, md = orc
    x = random_states(md, 8, seed=4)
    ee, dist = o.kinematics(x)
    for i in range(8):
        ref = robot_model.fk_point(md.chain, x[i, :5], md.point_body[0], md.point_local[0])
        np.testing.assert_allclose(ee[i], ref, atol=1e-12)
        for p, pr in enumerate(md.pairs):
            A = robot_model.fk_point(md.chain, x[i, :5], md.point_body[pr['pa']], md.point_local[pr['pa']])
            Bp = robot_model.fk_point(md.chain, x[i, :5], md.point_body[pr['pb']], md.point_local[pr['pb']])
            # brute-force minimum squared distance between the two segments
            s = np.linspace(0, 1, 401)
            P1 = A[None] + s[:, None] * (Bp - A)[None]
            P2 = pr['C'][None] + s[:, None] * (pr['D'] - pr['C'])[None]
            d2 = ((P1[:, None, :] - P2[None, :, :]) ** 2).sum(-1).min()
            assert dist[i, p] <= d2 + 1e-9
            assert dist[i, p] >= d2 - 2e-3     # grid resolution + the reference's 1e-5 regulariser


def test_mlp_matches_torch_fp64(orc):
    import torch
    o, prob, params, md = orc
    ws, bs, mean, std = load_network(params)
    x = random_states(md, 32, seed=5, vel_scale=0.6)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    v = xt[:, 5:] + torch.tensor([params.eps, 0, 0, 0, 0], dtype=torch.float64)
    nrm = v.norm(dim=1, keepdim=True)
    h = torch.cat([(xt[:, :5] - torch.tensor(mean)) / torch.tensor(std), v / nrm], dim=1)
    for l in range(3):
        h = torch.nn.functional.gelu(h @ torch.tensor(ws[l], dtype=torch.float64).T + torch.tensor(bs[l], dtype=torch.float64), approximate='tanh')
    y = h @ torch.tensor(ws[3], dtype=torch.float64).T + torch.tensor(bs[3], dtype=torch.float64)
    c = y[:, 0] * (100 - params.alpha) / 100 - nrm[:, 0]
    c.sum().backward()
    co, go = o.nn_constraint(x)
    np.testing.assert_allclose(co, c.detach().numpy(), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(go, xt.grad.numpy(), rtol=1e-9, atol=1e-10)
    # the reference evaluates the network in fp32 (libtorch behind L4CasADi): the fp64-accumulate mode stays
    # inside that rounding band
    o.set_mlp_fp32(1)
    c32 = o.nn_constraint(x, grad=False)
    o.set_mlp_fp32(0)
    assert np.abs(c32 - co).max() < 2e-5


def _records(o, prob, x0, xg, ug):
    o.set_guess(xg, ug)
    o.rti_solve(x0)
    return o.get_lin()


def test_stage_records_against_finite_differences(orc):
    """Jacobians / Hessian in the stage records vs central differences of the oracle's own value functions."""
    o, prob, params, md = orc
    B, N = o.B, prob.N
    x0 = start_states(B, seed=6)
    xg, ug = constant_guess(x0, N)
    rng = np.random.default_rng(7)
    xg += 0.05 * rng.uniform(-1, 1, xg.shape)
    ug += rng.uniform(-2, 2, ug.shape)
    lin = _records(o, prob, x0, xg, ug)
    eps = 1e-6
    for b in range(2):
        for k in (0, 7, N - 1, N):
            rec = lin[b, k]
            x = xg[b, k]
            np.testing.assert_array_equal(rec[abi.REC_X:abi.REC_X + 10], x)
            # capsule rows
            J = rec[abi.REC_JDIST:abi.REC_JDIST + 30].reshape(6, 5)
            Jfd = np.zeros((6, 5))
            for j in range(5):
                e = np.zeros(10); e[j] = eps
                Jfd[:, j] = (o.kinematics((x + e)[None])[1][0] - o.kinematics((x - e)[None])[1][0]) / (2 * eps)
            np.testing.assert_allclose(J, Jfd, rtol=2e-5, atol=2e-7)
            np.testing.assert_allclose(rec[abi.REC_DIST:abi.REC_DIST + 6], o.kinematics(x[None])[1][0], atol=1e-14)
            # cost: f(q) = s * Q * |p(q) - ref|^2
            s = 1.0 if k == N else params.dt
            def cost(xx):
                ee = o.kinematics(xx[None])[0][0]
                return s * params.Q_weight * ((ee - params.ee_ref) ** 2).sum()
            g_fd = np.zeros(5); H_fd = np.zeros((5, 5))
            h = 1e-4
            for i in range(5):
                ei = np.zeros(10); ei[i] = h
                g_fd[i] = (cost(x + ei) - cost(x - ei)) / (2 * h)
                for j in range(5):
                    ej = np.zeros(10); ej[j] = h
                    H_fd[i, j] = (cost(x + ei + ej) - cost(x + ei - ej) - cost(x - ei + ej) + cost(x - ei - ej)) / (4 * h * h)
            np.testing.assert_allclose(rec[abi.REC_G + 5:abi.REC_G + 10], g_fd, rtol=1e-6, atol=1e-8)
            tri = rec[abi.REC_HQQ:abi.REC_HQQ + 15]
            Hq = np.zeros((5, 5)); oidx = 0
            for i in range(5):
                for j in range(i + 1):
                    Hq[i, j] = Hq[j, i] = tri[oidx]; oidx += 1
            np.testing.assert_allclose(Hq, H_fd, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(H_fd).max()))
            if k < N:
                u = ug[b, k]
                np.testing.assert_allclose(rec[abi.REC_TAU:abi.REC_TAU + 5], o.tau(x[None], u[None])[0], atol=1e-13)
                Jt = rec[abi.REC_JTAU:abi.REC_JTAU + 75].reshape(5, 15)
                Jfd = np.zeros((5, 15))
                for j in range(15):
                    du = np.zeros(5); dx = np.zeros(10)
                    if j < 5: du[j] = eps
                    else: dx[j - 5] = eps
                    Jfd[:, j] = (o.tau((x + dx)[None], (u + du)[None])[0] - o.tau((x - dx)[None], (u - du)[None])[0]) / (2 * eps)
                np.testing.assert_allclose(Jt, Jfd, rtol=1e-6, atol=1e-7)
                M, _ = o.mass_bias(0, x)
                np.testing.assert_allclose(Jt[:, :5], M, atol=1e-12)
                # u-gradient and Hessian diagonal, LM scaled by dt
                np.testing.assert_allclose(rec[abi.REC_G:abi.REC_G + 5], params.dt * 2 * params.R_weight * u, rtol=1e-14)
                assert rec[abi.REC_HU] == pytest.approx(params.dt * 2 * params.R_weight + params.dt * params.levenberg_marquardt)
                xn = np.hstack([x[:5] + params.dt * x[5:] + 0.5 * params.dt ** 2 * u, x[5:] + params.dt * u])
                np.testing.assert_allclose(rec[abi.REC_B:abi.REC_B + 10], xn - xg[b, k + 1], atol=1e-14)
            else:
                assert rec[abi.REC_HV] == pytest.approx(params.levenberg_marquardt)
                # viability row of the terminal stage
                c, g = o.nn_constraint(x[None])
                assert rec[abi.REC_NNROW] == 1.0
                np.testing.assert_allclose(rec[abi.REC_NN], c[0], atol=1e-14)
                np.testing.assert_allclose(rec[abi.REC_JNN:abi.REC_JNN + 10], g[0], atol=1e-14)
                gfd = np.array([(o.nn_constraint((x + eps * e)[None], grad=False)[0] - o.nn_constraint((x - eps * e)[None], grad=False)[0]) / (2 * eps) for e in np.eye(10)])
                np.testing.assert_allclose(g[0], gfd, rtol=1e-5, atol=1e-6)


def test_plant_step_consistency(orc):
    """integrate(): with the nominal plant, no noise and unsaturated torques the applied acceleration is u."""
    o, prob, params, md = orc
    x = start_states(o.B, seed=8)
    u = np.random.default_rng(9).uniform(-1, 1, (o.B, 5))
    xn, a = o.plant_step(x, u)
    np.testing.assert_allclose(a, u, atol=1e-9)
    np.testing.assert_allclose(xn[:, 5:], x[:, 5:] + params.dt * u, atol=1e-11)
    # perturbed plant: M_p a + h_p = clip(tau_nominal + noise)
    rng = np.random.default_rng(10)
    pin = np.tile(md.inertial, (o.B, 1, 1)) * (1 + 0.1 * rng.uniform(-1, 1, (o.B, 5, 10)))
    noise = rng.normal(0, 1.0, (o.B, 5))
    o.set_plant_inertial(pin); o.set_torque_noise(noise)
    big_u = 40 * u
    xn, a = o.plant_step(x, big_u)
    tau = np.clip(o.tau(x, big_u) + noise, md.tau_min, md.tau_max)
    for b in range(o.B):
        M, h = o.mass_bias(b, x[b], nominal=False)
        np.testing.assert_allclose(M @ a[b] + h, tau[b], rtol=1e-10, atol=1e-10)
    o.set_plant_inertial(np.tile(md.inertial, (o.B, 1, 1))); o.set_torque_noise(np.zeros((o.B, 5)))


def test_flop_tally():
    """the oracle built with a counting AD scalar (oracle/dual.hpp, -DORC_COUNT_FLOPS) tallies the floating-point
    operations of one linearisation.  Forward-mode AD (15 directions for the torque rows, second-order duals for the cost) costs about
    five times the hand-derived analytic recursion the CUDA kernel executes (16 274 flop per stage, profiles/r02_linearize_flops.md)."""
    import ctypes as C
    import os
    import subprocess
    import oracle.oracle as oo
    d = os.path.dirname(os.path.abspath(oo.__file__))
    subprocess.run(['make', '-C', d, '-s', 'liboracle_count.so'], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(d, 'liboracle_count.so'))
    lib.orc_flops_get.restype = C.c_ulonglong
    from tests.common import make_problem, start_states, rollout_guess
    prob, params, md = make_problem('naive', N=45)
    saved = oo._LIB
    oo._LIB = lib
    try:
        o = oo.Oracle(prob, 1, 1)
        x0 = start_states(1, seed=3)
        xg, ug = rollout_guess(x0, 45, params.dt, seed=4)
        o.set_guess(xg, ug)
        lib.orc_flops_reset()
        o.rti_solve(x0)
        per_stage = lib.orc_flops_get() / 46.0
        o.close()
    finally:
        oo._LIB = saved
    assert abs(per_stage - 84382.5) < 1.0, per_stage
