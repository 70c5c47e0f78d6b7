"""OCP description -> ``struct smpc_problem``.

This is the engine-side equivalent of what the reference assembles with ``AdamModel.__init__``
(env_model.py:18-165: chain, limits, capsule kinematics, NL-constraint lists) and
``AbstractController.__init__`` + the per-controller ``additionalSetting`` (controller.py:12-125,
300-306,332-357,400-401,439-442,514-515,687-689,696-712): bounds, which stages carry which rows, which
rows are soft, the solver options.  Nothing is code-generated; the result is a flat struct of numbers.
"""
from __future__ import annotations

import os

import numpy as np

from . import abi, synthetic
from .robot_model import build_chain, capsule_local_points, point_on_link, GRAVITY

# CLI / class name -> (engine state machine, where the viability rows are, soft terminal row?)
CONTROLLERS = {
    'naive': (abi.CTRL['naive'], abi.NN_NONE, False),
    'zerovel': (abi.CTRL['zerovel'], abi.NN_NONE, False),
    'st': (abi.CTRL['st'], abi.NN_TERMINAL, True),
    'stwa': (abi.CTRL['stwa'], abi.NN_TERMINAL, True),
    'htwa': (abi.CTRL['htwa'], abi.NN_TERMINAL, False),
    'receding': (abi.CTRL['receding'], abi.NN_RECEDING, True),
    'real_receding': (abi.CTRL['real_receding'], abi.NN_TERMINAL, False),
    'constraint_everywhere': (abi.CTRL['constraint_everywhere'], abi.NN_EVERYWHERE, False),
    'backup': (abi.CTRL['backup'], abi.NN_NONE, False),
    'parallel': (abi.CTRL['parallel'], abi.NN_PARALLEL, False),     # controller.py:573-576: hard terminal + hard running rows
}

COSTS = {'zero': abi.COST_ZERO, 'ext': abi.COST_EXT, 'nls': abi.COST_NLS}


def resolve_network_path(params):
    """The reference opens ``params.net_path`` as it stands (safe_set.py:76), i.e. relative to the working directory of its scripts
    (``scripts/``: config.yaml:65 says '../nn_models/z1/...').  Tried in that order: the path as given, relative to ``scripts/`` of
    this repository, relative to the repository root."""
    path = params.net_path
    if os.path.isabs(path):
        return path
    cands = [os.path.abspath(path), os.path.normpath(os.path.join(params.ROOT_DIR, 'scripts', path)),
             os.path.normpath(os.path.join(params.ROOT_DIR, path))]
    for c in cands:
        if os.path.isfile(c):
            return c
    return cands[1]


def load_network(params):
    """Weights / normalisation of the viability network in the reference's file format (safe_set.py:76-85: ``torch.load`` of
    {'model', 'mean', 'std'}).  A missing file raises, as the reference's ``torch.load`` does; the synthetic stand-in of
    safe_mpc_b200/synthetic.py is only written on explicit request (``SMPC_SYNTHETIC_ASSETS=1`` or ``params.allow_synthetic``)."""
    path = resolve_network_path(params)
    if not os.path.isfile(path):
        if not (os.environ.get('SMPC_SYNTHETIC_ASSETS') == '1' or getattr(params, 'allow_synthetic', False)):
            raise FileNotFoundError(f'viability network {path} (config.yaml network_path = {params.net_path!r}) does not exist; set '
                                    'SMPC_SYNTHETIC_ASSETS=1 to generate the random-init stand-in of safe_mpc_b200/synthetic.py')
        synthetic.write_synthetic_net(path, nq=params.nq, hidden=int(params.net_size[1]))
    import torch
    data = torch.load(path, map_location='cpu', weights_only=True)
    sd = data['model']
    ws = [sd[f'linear_stack.{i}.weight'].numpy() for i in (0, 2, 4, 6)]
    # is this the random-init stand-in (same generator, seed 0)?  recorded so that scripts / bench can say so
    ref0 = synthetic.synthetic_net_arrays(params.nq, int(params.net_size[1]))[0][0]
    params.synthetic_net = bool(ws[0].shape == ref0.shape and np.array_equal(ws[0], ref0))
    bs = [sd[f'linear_stack.{i}.bias'].numpy() for i in (0, 2, 4, 6)]
    mean = np.broadcast_to(np.asarray(data['mean'], dtype=np.float64).ravel(), (params.nq,)) \
        if np.asarray(data['mean']).size in (1, params.nq) else np.asarray(data['mean'], dtype=np.float64).ravel()[:params.nq]
    std = np.broadcast_to(np.asarray(data['std'], dtype=np.float64).ravel(), (params.nq,)) \
        if np.asarray(data['std']).size in (1, params.nq) else np.asarray(data['std'], dtype=np.float64).ravel()[:params.nq]
    return ws, bs, np.array(mean, dtype=np.float64), np.array(std, dtype=np.float64)


class ModelData:
    """Numbers the reference's AdamModel derives from URDF + config (env_model.py:23-40,106-121,131-165,263-271)."""

    def __init__(self, params):
        self.params = params
        nq = params.nq
        if nq != abi.NQ:
            raise NotImplementedError(f'this build of the engine is compiled for n_dofs={abi.NQ} (config.yaml n_dofs={nq})')
        self.chain, self.nominal_links = build_chain(params.robot_descr, nq)
        c = self.chain
        self.inertial = c.lump(self.nominal_links['mass'], self.nominal_links['com'], self.nominal_links['inertia6'],
                               self.nominal_links['rpy'])
        # joint limits, widened by q_margin (env_model.py:106-121)
        self.tau_min, self.tau_max = -c.effort.copy(), c.effort.copy()
        x_min = np.hstack([c.lower, -c.velocity])
        x_max = np.hstack([c.upper, c.velocity])
        self.bounds_diff = np.abs(x_max - x_min)
        self.x_min = x_min - self.bounds_diff * (params.q_margin / 100)
        self.x_max = x_max + self.bounds_diff * (params.q_margin / 100)
        # moving points: 0 = end effector (env_model.py:91-95), then capsule end points (:131-147)
        bodies, locals_ = [], []
        b, p = point_on_link(c, params.frame_name, np.append(params.ee_pos, 1.0))
        bodies.append(b); locals_.append(p)
        self.capsule_points = {}
        for cap in params.robot_capsules:
            idx = []
            for lp in capsule_local_points(cap):
                b, p = point_on_link(c, cap['link_name'], lp)
                idx.append(len(bodies)); bodies.append(b); locals_.append(p)
            self.capsule_points[cap['name']] = idx
        used = {0}
        self.pairs = []
        for pair in params.collisions_pairs:
            if pair['type'] != 'capsule-capsule':
                raise NotImplementedError(f"collision pair type {pair['type']} is outside the engine's scope "
                                          '(config.yaml ships capsule-capsule pairs only)')
            mov, fix = pair['elements']
            if mov.get('type') != 'moving_capsule' or fix.get('type') != 'fixed_capsule':
                raise NotImplementedError('capsule pairs must be (robot capsule, obstacle capsule)')
            ia, ib = self.capsule_points[mov['name']]
            used.update((ia, ib))
            rsum = mov['radius'] + fix['radius']
            self.pairs.append(dict(pa=ia, pb=ib, C=np.asarray(fix['end_points'][0], dtype=np.float64),
                                   D=np.asarray(fix['end_points'][1], dtype=np.float64),
                                   lo_ocp=(rsum + params.collision_margin * 2) ** 2,
                                   lo_chk=rsum ** 2 - params.tol_obs, name=(mov['name'], fix['name'])))
        if len(self.pairs) != abi.NPAIR:
            raise NotImplementedError(f'this build expects {abi.NPAIR} capsule-capsule pairs, config has {len(self.pairs)}')
        # compact the point list to the ones in use
        order = sorted(used)
        if len(order) > abi.MAX_POINTS:
            raise NotImplementedError('too many moving points')
        remap = {old: new for new, old in enumerate(order)}
        self.point_body = np.full(abi.MAX_POINTS, -1, dtype=np.int32)
        self.point_local = np.zeros((abi.MAX_POINTS, 3))
        for old in order:
            self.point_body[remap[old]] = bodies[old]
            self.point_local[remap[old]] = locals_[old]
        for pr in self.pairs:
            pr['pa'], pr['pb'] = remap[pr['pa']], remap[pr['pb']]
        self.n_points = len(order)


def build_problem(params, controller: str, cost: str = 'ext', N: int | None = None, model: ModelData | None = None,
                  nn_precision: str | None = None, precision: str | None = None):
    """-> (abi.Problem, keepalive).  ``keepalive`` owns the weight buffer the struct points to.

    ``nn_precision``: 'strict' (fp32 weights, fp64 accumulation) or 'tf32x3' (fp32-class evaluation on the tensor cores,
    the precision of the reference's libtorch call); default: ``params.nn_precision`` if set, else 'strict'."""
    if controller not in CONTROLLERS:
        raise ValueError(f'Controller {controller} not available')
    ctrl, nn_rows, soft = CONTROLLERS[controller]
    md = model or ModelData(params)
    N = int(params.N if N is None else N)
    if not 2 <= N <= abi.MAX_N:
        raise ValueError(f'horizon {N} outside [2, {abi.MAX_N}]')
    p = abi.Problem()
    q_m = params.q_margin / 100
    lbx = md.x_min + q_m * md.bounds_diff          # controller.py:49-55: re-narrow by the same margin
    ubx = md.x_max - q_m * md.bounds_diff
    lbx_e, ubx_e = lbx.copy(), ubx.copy()
    nq = params.nq
    if controller == 'zerovel':                    # controller.py:300-306
        lbx_e[nq:] = 0.0; ubx_e[nq:] = 0.0
    if controller == 'backup':                     # controller.py:701-707 (uses the model bounds)
        lbx_e = np.hstack([md.x_min[:nq], np.zeros(nq)]); ubx_e = np.hstack([md.x_max[:nq], np.zeros(nq)])
    keep = None
    if nn_rows != abi.NN_NONE:
        if not params.use_net:
            raise NotImplementedError('use_net: false selects the analytic safe set (safe_set.py:106-155) as OCP rows; the engine evaluates it '
                                      '(safe_set.AnalyticSafeSet: values and check_constraint) but its QP carries the network row only')
        if str(params.act_fun).lower() != 'gelu':
            raise NotImplementedError(f"act_fun = {params.act_fun!r}: the engine implements the shipped network (GELU, tanh form; config.yaml act_fun: 'gelu') only")
        if int(params.n_dof_safe_set) != int(params.nq):
            raise NotImplementedError(f'n_dof_safe_set = {params.n_dof_safe_set} != n_dofs = {params.nq}: the engine feeds every joint to the network')
        ws, bs, mean, std = load_network(params)
        keep = abi.pack_nn_weights(ws, bs)
    else:
        mean, std = np.zeros(nq), np.ones(nq)
    precision = precision or getattr(params, 'precision', None) or 'f64'
    if precision not in abi.PRECISION:
        raise ValueError(f'unknown precision {precision} (f64, f32)')
    # HPIPM BALANCE tolerances (UNVERIFIED offline).  fp32 storage: the stored search direction carries a rounding of
    # 6e-8 that the condensation amplifies by lam / t, the residuals bottom out near 1e-5 -> one decade looser
    tols = dict(qp_tol_stat=1e-6, qp_tol_eq=1e-8, qp_tol_ineq=1e-8, qp_tol_comp=1e-8) if precision == 'f64' else \
        dict(qp_tol_stat=1e-5, qp_tol_eq=1e-6, qp_tol_ineq=1e-6, qp_tol_comp=1e-6)
    if controller == 'receding':
        penalty = params.ws_t                      # controller.py:461-462 (runtime cost_set wins over zl_e)
    else:
        penalty = params.ws_r                      # controller.py:351-354
    abi.fill(
        p, nq=nq, N=N, n_pairs=abi.NPAIR, n_points=md.n_points, controller=ctrl, nn_rows=nn_rows,
        nn_terminal_soft=int(soft), stage0_collision_rows=int(not params.noise > 0), cost_type=COSTS[cost],
        abort_flag=int(params.abort_flag), qp_iter_max=int(params.qp_max_iter), lm_scale_dt=1, qp_cond_pred_corr=1,
        # the SQP of the guess generator reads the QP multipliers back (merit line search): keep every problem in its slot
        qp_keep_slots=int(getattr(params, 'qp_keep_slots', params.solver_type == 'SQP')),
        nn_precision=abi.NN_PRECISION[nn_precision or getattr(params, 'nn_precision', None) or 'strict'],
        dt=params.dt, q_weight=params.Q_weight, r_weight=params.R_weight, lm=params.levenberg_marquardt,
        alpha=params.alpha, eps=params.eps, slack_penalty_e=penalty,
        tol_x=params.tol_x, tol_tau=params.tol_tau, tol_obs=params.tol_obs, tol_safe=params.tol_safe_set,
        tol_conv=params.tol_conv,
        qp_mu0=1e1, qp_alpha_min=1e-12, precision=abi.PRECISION[precision], **tols,
        qp_reg_prim=1e-15, qp_maxiter_accept=float(getattr(params, 'qp_maxiter_accept', 1e3)),
        gravity=[0.0, 0.0, -GRAVITY],
        joint_R=md.chain.joint_R, joint_p=md.chain.joint_p, joint_axis=md.chain.joint_axis, inertial=md.inertial,
        x_min=md.x_min, x_max=md.x_max, lbx=lbx, ubx=ubx, lbx_e=lbx_e, ubx_e=ubx_e,
        tau_min=md.tau_min, tau_max=md.tau_max, ee_ref=params.ee_ref,
        point_body=md.point_body, point_local=md.point_local,
        pair_pa=[pr['pa'] for pr in md.pairs], pair_pb=[pr['pb'] for pr in md.pairs],
        pair_C=[pr['C'] for pr in md.pairs], pair_D=[pr['D'] for pr in md.pairs],
        pair_lo_ocp=[pr['lo_ocp'] for pr in md.pairs], pair_lo_chk=[pr['lo_chk'] for pr in md.pairs],
        pair_hi=1e6, nn_mean=mean, nn_std=std,
    )
    if keep is not None:
        import ctypes as C
        p.nn_weights = keep.ctypes.data_as(C.POINTER(C.c_float))
    return p, keep
