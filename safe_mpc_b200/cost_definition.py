"""Cost objects of the reference (src/safe_mpc/cost_definition.py), reduced to what the engine needs: which of the
three cost kinds the stage records are linearised with.  The expressions themselves live in the CUDA linearisation
kernel (csrc/dev_model.cuh: EE position, Jacobian and the exact second-order term of the EXTERNAL cost)."""
from __future__ import annotations

import os

import numpy as np
import yaml


class _Cost:
    kind = 'ext'
    tracking = False      # True: self.traj is a time-indexed end-effector path the engine must be given (smpc_set_ee_trajectory)

    def __init__(self, model, Q_weight=None, R_weight=None):
        self.model = model
        model.params.track_traj = False                     # cost_definition.py:12
        # the weights are read from params by build_problem; explicit arguments override them like the reference's
        if Q_weight is not None:
            model.params.Q_weight = float(Q_weight)
        if R_weight is not None:
            model.params.R_weight = float(R_weight)

    def set_solver_cost(self, controller):
        """cost_definition.py:17-31: attach this cost to a controller before build_controller()."""
        controller.cost = self
        controller.cost_kind = self.kind

    @property
    def traj(self):
        """cost_definition.py:29-31,41,67,89: the reach costs track ee_ref at every step, [3, n_steps + 1 + N]"""
        p = self.model.params
        return np.tile(np.asarray(p.ee_ref, dtype=np.float64), (p.n_steps + 1 + p.N, 1)).T

    def get_reference_traj(self):
        return np.tile(np.asarray(self.model.params.ee_ref, dtype=np.float64), (self.model.params.N + 1, 1)).T

    def update_trajectory(self):
        """cost_definition.py:29-31 (called by resetHorizon, controller.py:214)"""


class ZeroCost(_Cost):
    """cost_definition.py:34-46 -- used by the backup OCP (mpc.py:63-64)."""
    kind = 'zero'

    def __init__(self, model):
        super().__init__(model)


class ReachTargetNLS(_Cost):
    """cost_definition.py:61-81 -- NONLINEAR_LS reach cost (Gauss-Newton Hessian), used for naive / zerovel guesses."""
    kind = 'nls'


class ReachTargetEXT(_Cost):
    """cost_definition.py:83-100 -- EXTERNAL reach cost (exact Hessian), the closed-loop cost of mpc.py:48-51."""
    kind = 'ext'


# ---------------------------------------------------------------------------------------------------------------------------------
# Tracking costs (cost_definition.py:102-288): the same NLS / EXT stage costs around a time-indexed end-effector reference.
# The reference hands cost.traj[:, current_step + i] to stage i of every solve (controller.py:153-156); here the whole path is given
# to the engine once (Engine.set_ee_trajectory) and indexed on the device with the per-problem step counter.
# ---------------------------------------------------------------------------------------------------------------------------------
def _tracking_settings(params):
    """The tracking keys are read from config.yaml by the cost object itself (cost_definition.py:105-114,206-216)."""
    with open(os.path.join(params.ROOT_DIR, 'config.yaml')) as fh:
        return yaml.load(fh, Loader=yaml.FullLoader)


def _lemniscate_speed(a, theta):
    """|d(x, y)/d theta| of x = a cos t / (1 + sin^2 t), y = a cos t sin t / (1 + sin^2 t) -- what the reference obtains from sympy
    (cost_definition.py:124-133,183-190): dx = -a sin t (3 - sin^2 t) / (1 + sin^2 t)^2, dy = a (1 - 3 sin^2 t) / (1 + sin^2 t)^2"""
    s = np.sin(theta)
    d2 = (1.0 + s * s) ** 2
    dx = -a * s * (3.0 - s * s) / d2
    dy = a * (1.0 - 3.0 * s * s) / d2
    return np.sqrt(dx ** 2 + dy ** 2)


def generate_8shape_trajectory(params):
    """cost_definition.py:181-209: figure-eight path in the plane, rotated and offset, one point per control step."""
    from .robot_model import rot_mat_x, rot_mat_y, rot_mat_z
    if params.vel_const is False:
        velocity = 0
        acc = params.vel_max_traj / (params.n_steps * params.acc_time)
    else:
        velocity = params.vel_max_traj
    traj = np.zeros((3, params.n_steps_tracking + 1 + params.N))
    theta = 0
    for i in range(traj.shape[1]):
        traj[:, i] = np.array([(params.dim_shape_8 * np.cos(theta)) / (1 + np.sin(theta) ** 2),
                               (params.dim_shape_8 * np.cos(theta) * np.sin(theta)) / (1 + np.sin(theta) ** 2), 0])
        theta = theta + (velocity / _lemniscate_speed(params.dim_shape_8, theta)) * params.dt
        if not params.vel_const and velocity <= params.vel_max_traj:
            velocity += acc
    rot = rot_mat_x(params.theta_rot_traj[0]) @ rot_mat_y(params.theta_rot_traj[1]) @ rot_mat_z(params.theta_rot_traj[2])
    return (rot[:3, :3] @ traj) + np.asarray(params.offset_traj).reshape(3, 1)


def generate_moving_circle_trajectory(params):
    """cost_definition.py:262-286: circle whose centre oscillates along y."""
    if params.vel_const is False:
        velocity = 0
        acc = params.circle_traj_vel / (params.n_steps * params.acc_time)
    else:
        velocity = params.circle_traj_vel
    n = params.n_steps_tracking + 1 + params.N
    traj = np.zeros((3, n)); circle = np.zeros((3, n)); linear_mov = np.zeros((3, n))
    theta = 0
    sign_vel = 1
    for i in range(n):
        circle[:, i] = params.circle_rad * np.array([-np.cos(theta), np.sin(theta), 0])
        linear_mov[:, i] = linear_mov[:, max(i - 1, 0)] - sign_vel * np.array([0, params.circle_center_vel * params.dt, 0])
        traj[:, i] = circle[:, i] + linear_mov[:, i] + np.array(params.circle_offset_traj)
        theta = theta + (velocity / (np.sqrt(params.circle_rad * (np.sin(theta) ** 2 + np.cos(theta) ** 2)))) * params.dt
        if sign_vel > 0 and traj[1, i] < -0.5:
            sign_vel = -1
        if sign_vel < 0 and traj[1, i] > 0.5:
            sign_vel = 1
        if not params.vel_const and velocity <= params.circle_traj_vel:
            velocity += acc
    return traj


class _Tracking(_Cost):
    tracking = True
    generator = None

    def __init__(self, model, Q_weight=None, R_weight=None):
        super().__init__(model, Q_weight, R_weight)
        self._configure(model.params, _tracking_settings(model.params))
        model.params.track_traj = True
        self._traj = type(self).generator(model.params)

    @property
    def traj(self):
        return self._traj

    def update_trajectory(self):
        self._traj = type(self).generator(self.model.params)


class _Tracking8(_Tracking):
    generator = staticmethod(generate_8shape_trajectory)

    @staticmethod
    def _configure(params, cfg):                            # cost_definition.py:107-114
        params.n_steps = int(cfg['n_steps_tracking'])
        params.n_steps_tracking = int(cfg['n_steps_tracking'])
        params.dim_shape_8 = float(cfg['dim_shape_8'])
        params.offset_traj = np.array(cfg['offset_traj'], dtype=np.float64)
        params.theta_rot_traj = np.array(cfg['theta_rot_traj'], dtype=np.float64)
        params.vel_max_traj = float(cfg['vel_max_traj'])
        params.vel_const = bool(cfg['vel_const'])
        params.acc_time = float(cfg['acc_time'])


class _TrackingMovingCircle(_Tracking):
    generator = staticmethod(generate_moving_circle_trajectory)

    @staticmethod
    def _configure(params, cfg):                            # cost_definition.py:208-216
        params.n_steps = int(cfg['n_steps_tracking'])
        params.n_steps_tracking = int(cfg['n_steps_tracking'])
        params.circle_rad = float(cfg['circle_rad'])
        params.circle_offset_traj = np.array(cfg['circle_offset_traj'], dtype=np.float64)
        params.circle_traj_vel = float(cfg['circle_traj_vel'])
        params.vel_const = bool(cfg['vel_const'])
        params.circle_center_vel = float(cfg['circle_center_vel'])
        params.acc_time = float(cfg['acc_time'])


class Tracking8NLS(_Tracking8):
    """cost_definition.py:160-167"""
    kind = 'nls'


class Tracking8EXT(_Tracking8):
    """cost_definition.py:169-176"""
    kind = 'ext'


class TrackingMovingCircleNLS(_TrackingMovingCircle):
    """cost_definition.py:244-251"""
    kind = 'nls'


class TrackingMovingCircleEXT(_TrackingMovingCircle):
    """cost_definition.py:253-260"""
    kind = 'ext'
