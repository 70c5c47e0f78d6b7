"""Shared fixtures of the test-suite: problem construction, seeded inputs, a numpy rebuild of the stage QP."""
from __future__ import annotations

import functools

import numpy as np

from safe_mpc_b200 import abi
from safe_mpc_b200.parser import Parameters, default_args
from safe_mpc_b200.problem import ModelData, build_problem

Q0 = np.array([-0.3, 0.8, -1.65, 0.658, 0.0])     # the shipped initial configuration (guess_acados.py:101-103)


@functools.lru_cache(maxsize=None)
def params_model(noise=0.0, alpha=10.0, q_margin=0.0, collision_margin=0.0, control_noise=0.0):
    args = default_args(noise=noise, alpha=alpha, joint_bounds_margin=q_margin, collision_margin=collision_margin,
                        control_noise=control_noise)
    params = Parameters(args, 'z1', rti=True)
    params.alpha = alpha
    return params, ModelData(params)


def make_problem(controller='naive', cost='ext', N=None, nn_precision=None, keep_slots=False, precision=None, **kw):
    params, md = params_model(**kw)
    prob, keep = build_problem(params, controller, cost=cost, N=N, model=md, nn_precision=nn_precision, precision=precision)
    prob.qp_keep_slots = int(keep_slots)
    prob._keep = keep
    return prob, params, md


def random_states(md, n, seed=0, vel_scale=0.3, shrink=0.8):
    rng = np.random.default_rng(seed)
    mid = 0.5 * (md.x_min + md.x_max)
    half = 0.5 * (md.x_max - md.x_min)
    x = mid + shrink * half * rng.uniform(-1, 1, (n, abi.NX))
    x[:, abi.NQ:] *= vel_scale
    return x


def start_states(B, seed=0, spread=0.15, vel=0.2):
    """States around the shipped initial configuration."""
    rng = np.random.default_rng(seed)
    x = np.zeros((B, abi.NX))
    x[:, :abi.NQ] = Q0 + spread * rng.uniform(-1, 1, (B, abi.NQ))
    x[:, abi.NQ:] = vel * rng.uniform(-1, 1, (B, abi.NQ))
    return x


def constant_guess(x0, N):
    B = x0.shape[0]
    return np.repeat(x0[:, None, :], N + 1, axis=1).copy(), np.zeros((B, N, abi.NU))


def rollout_guess(x0, N, dt, seed=0, scale=2.0):
    """Dynamically consistent guess with small random controls."""
    rng = np.random.default_rng(seed)
    B = x0.shape[0]
    ug = scale * rng.uniform(-1, 1, (B, N, abi.NU))
    xg = np.zeros((B, N + 1, abi.NX))
    xg[:, 0] = x0
    nq = abi.NQ
    for k in range(N):
        xg[:, k + 1, :nq] = xg[:, k, :nq] + dt * xg[:, k, nq:] + 0.5 * dt * dt * ug[:, k]
        xg[:, k + 1, nq:] = xg[:, k, nq:] + dt * ug[:, k]
    return xg, ug


# --------------------------------------------------------------------------------------------------
# numpy rebuild of the QP of one problem from its stage records (independent of both implementations)
# --------------------------------------------------------------------------------------------------
def stage_qp(prob, rec, k, box_lo, box_hi):
    """-> H[nz,nz], g[nz], rows a[nr,nz], lo[nr], hi[nr], canonical row ids, soft penalty (or None)."""
    N = prob.N
    term = k == N
    nu = 0 if term else abi.NU
    nz = nu + abi.NX
    H = np.zeros((nz, nz)); g = np.zeros(nz)
    if not term:
        H[:nu, :nu] = np.eye(nu) * rec[abi.REC_HU]
        g[:nu] = rec[abi.REC_G:abi.REC_G + 5]
    tri = rec[abi.REC_HQQ:abi.REC_HQQ + 15]
    o = 0
    for i in range(5):
        for j in range(i + 1):
            H[nu + i, nu + j] = H[nu + j, nu + i] = tri[o]; o += 1
    H[nu:nu + 5, nu:nu + 5] += np.eye(5) * rec[abi.REC_HQ]
    H[nu + 5:, nu + 5:] = np.eye(5) * rec[abi.REC_HV]
    g[nu:] = rec[abi.REC_G + 5:abi.REC_G + 15]
    rows, lo, hi, ids = [], [], [], []
    for i in range(abi.NX):
        a = np.zeros(nz); a[nu + i] = 1.0
        rows.append(a); lo.append(box_lo[i]); hi.append(box_hi[i]); ids.append(i)
    tau_min = np.array(prob.tau_min[:]); tau_max = np.array(prob.tau_max[:])
    for i in range(int(rec[abi.REC_NTAU])):
        a = rec[abi.REC_JTAU + 15 * i:abi.REC_JTAU + 15 * (i + 1)].copy()
        rows.append(a); lo.append(tau_min[i] - rec[abi.REC_TAU + i]); hi.append(tau_max[i] - rec[abi.REC_TAU + i]); ids.append(10 + i)
    for p in range(int(rec[abi.REC_NDIST])):
        a = np.zeros(nz); a[nu:nu + 5] = rec[abi.REC_JDIST + 5 * p:abi.REC_JDIST + 5 * (p + 1)]
        rows.append(a); lo.append(prob.pair_lo_ocp[p] - rec[abi.REC_DIST + p]); hi.append(prob.pair_hi - rec[abi.REC_DIST + p]); ids.append(15 + p)
    soft = None
    if rec[abi.REC_NNROW] > 0.5:
        a = np.zeros(nz); a[nu:] = rec[abi.REC_JNN:abi.REC_JNN + 10]
        rows.append(a); lo.append(0.0 - rec[abi.REC_NN]); hi.append(1e6 - rec[abi.REC_NN]); ids.append(21)
        if rec[abi.REC_SOFT] >= 0:
            soft = float(rec[abi.REC_SOFT])
    return H, g, np.array(rows), np.array(lo), np.array(hi), ids, soft


def dyn_mats(dt):
    nq = abi.NQ
    A = np.eye(2 * nq); A[:nq, nq:] = dt * np.eye(nq)
    Bm = np.vstack([0.5 * dt * dt * np.eye(nq), dt * np.eye(nq)])
    return A, Bm


def kkt_residuals(prob, lin, x0, dz, pi, lam, t, boxes=None):
    """Max-norm KKT residuals of (dz, pi, lam, t) for the QP defined by the stage records `lin` of ONE problem.
    Soft rows: the slack values are implied (s = max(0, violation)); stationarity w.r.t. the slack is checked as
    lam_row <= penalty.  Returns dict(stat, eq, ineq, comp)."""
    N = prob.N
    A, Bm = dyn_mats(prob.dt)
    lbx, ubx = np.array(prob.lbx[:]), np.array(prob.ubx[:])
    lbx_e, ubx_e = np.array(prob.lbx_e[:]), np.array(prob.ubx_e[:])
    stat = eq = ineq = comp = 0.0
    for k in range(N + 1):
        rec = lin[k]
        xk = rec[abi.REC_X:abi.REC_X + 10]
        if boxes is not None:
            lo, hi = boxes[k]
        elif k == 0:
            lo = hi = x0 - xk
        elif k == N:
            lo, hi = lbx_e - xk, ubx_e - xk
        else:
            lo, hi = lbx - xk, ubx - xk
        H, g, rows, rlo, rhi, ids, soft = stage_qp(prob, rec, k, lo, hi)
        nu = 0 if k == N else abi.NU
        nz = nu + abi.NX
        z = dz[k][:nz]
        r = H @ z + g
        for a, rid in zip(rows, ids):
            r += a * (lam[k][abi.QP_NR + rid] - lam[k][rid])
        if k < N:
            r += np.concatenate([Bm.T @ pi[k], A.T @ pi[k]])
        if k > 0:
            r[nu:] -= pi[k - 1]
        stat = max(stat, np.abs(r).max())
        if k < N:
            xn = dz[k + 1][(0 if k + 1 == N else abi.NU):][:abi.NX]
            e = A @ z[nu:] + Bm @ z[:nu] + rec[abi.REC_B:abi.REC_B + 10] - xn
            eq = max(eq, np.abs(e).max())
        for a, l_, h_, rid in zip(rows, rlo, rhi, ids):
            v = a @ z
            ll, lu = lam[k][rid], lam[k][abi.QP_NR + rid]
            assert ll >= 0 and lu >= 0
            vl, vu = max(0.0, l_ - v), max(0.0, v - h_)
            if rid == 21 and soft is not None:
                assert ll <= soft * (1 + 1e-6) + 1e-6 and lu <= soft * (1 + 1e-6) + 1e-6
                # violated soft row: multiplier sits at the penalty
                if vl > 1e-6:
                    comp = max(comp, abs(soft - ll) * min(vl, 1.0))
                comp = max(comp, ll * max(0.0, v - l_) if vl == 0 else 0.0, lu * max(0.0, h_ - v))
            else:
                ineq = max(ineq, vl, vu)
                comp = max(comp, ll * abs(v - l_), lu * abs(h_ - v))
    return dict(stat=stat, eq=eq, ineq=ineq, comp=comp)


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: the reference's own closed-loop workload (config.yaml:1-7: test_num = 100 tests x n_steps = 800,
# N = 45), as scripts/guess_acados.py + scripts/mpc.py of the reference set it up
# --------------------------------------------------------------------------------------------------
def cfg0_initial_states(handle, md, params, B, flavour):
    """'shipped': every test starts from the shipped configuration (guess_acados.py:101-103, TEST_NOISE = True), the tests differ
    by their perturbed plant and torque-noise draw; 'halton': the collision-free points of Halton(nq, scramble=False) scaled to the
    joint box, zero velocity (guess_acados.py:79,100-109).  `handle`: anything with .kinematics (engine or oracle)."""
    nq = abi.NQ
    if flavour == 'shipped':
        x = np.zeros((B, abi.NX)); x[:, :nq] = Q0
        return x
    from scipy.stats import qmc
    sampler = qmc.Halton(nq, scramble=False)
    lo_chk = np.array([pr['lo_chk'] for pr in md.pairs])
    out = []
    while len(out) < B:
        q = qmc.scale(sampler.random(256), md.x_min[:nq], md.x_max[:nq])
        x = np.zeros((256, abi.NX)); x[:, :nq] = q
        d = handle.kinematics(x)[1]
        out += [xi for xi, ok in zip(x, (d >= lo_chk).all(axis=1)) if ok]
    return np.array(out[:B])


def cfg0_plants(md, params, B, noise, control_noise):
    """Per-test perturbed plants (mpc.py:106-107, utils.py:126-171 draw order, seed 0) and the per-test torque-noise vector
    (mpc.py:126-127: re-seeded with the test index every step, i.e. one fixed draw per test; env_model.py:196)."""
    from safe_mpc_b200.robot_model import randomized_link_inertials
    links = randomized_link_inertials(md.nominal_links, noise, noise, noise, B, seed=0)
    pin = np.stack([md.chain.lump(links['mass'][i], links['com'][i], links['inertia6'][i], links['rpy']) for i in range(B)])
    scale = md.tau_max * (control_noise / 100.0)
    tn = np.stack([np.random.default_rng(i).normal(np.zeros(abi.NU), scale, size=abi.NU) for i in range(B)])
    return pin, tn


def sqp_warm_start(handle, x0, N, iters):
    """Full-step SQP iterations from the trivial guess (x0 repeated, u = 0), accepted per problem on status 0 -> (xg, ug)."""
    B = x0.shape[0]
    xg, ug = constant_guess(x0, N)
    handle.set_guess(xg, ug)
    for _ in range(iters):
        st = handle.rti_solve(x0)
        xt, ut = handle.get_temp()
        ok = st == 0
        xg[ok], ug[ok] = xt[ok], ut[ok]
        handle.set_guess(xg, ug)
    handle.reset_controller()
    return xg, ug


def run_closed_loop(E, S, prob, bprob, x0, xg, ug, pin, tn, steps, arg=0, probe_eps=0.0, hook=None):
    """One closed loop of mpc.py:102-291 on implementation (E, S) -> dict(outcome, x, u, counters, x_viable[, flips])."""
    B = x0.shape[0]
    main, bk = E(prob, B, arg), E(bprob, B, arg)
    if hook is not None:
        main.set_qp_hook(hook); bk.set_qp_hook(hook)
    if probe_eps > 0.0:
        main.set_probe(probe_eps); bk.set_probe(probe_eps)
    main.set_guess(xg, ug); main.reset_controller()
    for h in (main, bk):
        h.set_plant_inertial(pin); h.set_torque_noise(tn)
    sim = S(main, bk, steps)
    sim.reset(x0)
    sim.run(steps)
    x, u = sim.log()
    out = dict(outcome=np.asarray(sim.outcome()), x=x, u=u, counters=sim.counters(), x_viable=sim.x_viable())
    if probe_eps > 0.0:
        out['flips'] = main.probe_flips() + bk.probe_flips()
    sim.close() if hasattr(sim, 'close') else None
    main.close(); bk.close()
    return out


def outcome_sets(outcome):
    """The four index sets of mpc.py:273-291 -> dict of counts (Completed task / Collisions / Viable states / Not converged)."""
    conv = (outcome & abi.OUT_CONVERGED) != 0
    coll = (outcome & abi.OUT_COLLIDED) != 0
    abrt = (outcome & abi.OUT_ABORTED) != 0
    viable = abrt & ~conv & ~coll
    return dict(completed=int(conv.sum()), collisions=int(coll.sum()), viable=int(viable.sum()),
                not_converged=int((~conv & ~coll & ~viable).sum()))
