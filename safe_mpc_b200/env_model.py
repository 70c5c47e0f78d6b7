"""Batched mirror of the reference's ``AdamModel`` (src/safe_mpc/env_model.py): same attribute and method names, every
state / control argument carries a leading batch dimension, and every number is produced by the CUDA engine through
the C ABI (there is no CasADi / adam evaluation and no CPU fallback here).

The model owns a small engine handle (naive OCP, N = 2) that serves the model functions -- ``tau_fun``, ``ee_fun`` /
``jointToEE``, the capsule constraint values, ``integrate`` on the per-problem perturbed plant -- exactly the calls
scripts/mpc.py makes on ``controller.model`` between two controller steps.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .problem import ModelData, build_problem
from .robot_model import randomized_link_inertials


class AdamModel:
    def __init__(self, params, batch: int = 1, device: int = 0):
        self.params = params
        self.batch = int(batch)
        self.device = device
        self.data = ModelData(params)
        d = self.data
        self.nq, self.nx, self.nu = params.nq, 2 * params.nq, params.nq          # env_model.py:55-57
        self.x_min, self.x_max = d.x_min.copy(), d.x_max.copy()                # env_model.py:115-121 (already widened)
        self.tau_min, self.tau_max = d.tau_min.copy(), d.tau_max.copy()        # env_model.py:113-114
        self.bounds_diff = d.bounds_diff
        self.ee_ref = np.asarray(params.ee_ref, dtype=np.float64)
        self.plant_inertial = np.tile(d.inertial, (self.batch, 1, 1))          # nominal until update_randomized_dynamics
        self.torque_noise = np.zeros((self.batch, self.nu))
        self._engine = None

    # ---- engine handle that evaluates the model functions ------------------------------------------------------
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            prob, keep = build_problem(self.params, 'naive', cost='ext', N=2, model=self.data)
            self._engine = Engine(prob, self.batch, self.device)
            self._engine._keepalive = (prob, keep)
            self._engine.set_plant_inertial(self.plant_inertial)
            self._engine.set_torque_noise(self.torque_noise)
        return self._engine

    def _rows(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return np.ascontiguousarray(x)

    # ---- functions of env_model.py ------------------------------------------------------------------------------
    def tau_fun(self, x, u):
        """env_model.py:80-83: M(q) u + h(q, qdot), rows of x / u."""
        return self.engine().tau(self._rows(x), self._rows(u))

    def ee_fun(self, x):
        """env_model.py:91-95."""
        return self.engine().kinematics(self._rows(x))[0]

    def jointToEE(self, x):
        """env_model.py:163-165."""
        return self.ee_fun(x)

    def collision_constraints(self, x):
        """The six squared capsule distances of env_model.py:263-271, rows of x."""
        return self.engine().kinematics(self._rows(x))[1]

    def checkStateBounds(self, x):
        """env_model.py:175-177 -- one flag per row."""
        x = self._rows(x)
        return np.logical_and(x >= self.x_min - self.params.tol_x, x <= self.x_max + self.params.tol_x).all(axis=1)

    def checkTorqueBounds(self, tau):
        """env_model.py:179-181."""
        tau = self._rows(tau)
        return np.logical_and(tau >= self.tau_min - self.params.tol_tau, tau <= self.tau_max + self.params.tol_tau).all(axis=1)

    def checkCollision(self, x):
        """env_model.py:226-244 -- one flag per row (the reference returns after the first row of a trajectory; callers
        that want that quirk pass the first row only, as the engine's controller kernels do)."""
        d = self.collision_constraints(x)
        lo = np.array([p['lo_chk'] for p in self.data.pairs])
        return (d >= lo).all(axis=1)

    def checkStateConstraints(self, x):
        """env_model.py:170-173 for a trajectory [T, nx]: bounds on every row, collision on row 0 only."""
        x = self._rows(x)
        return bool(self.checkStateBounds(x).all() and self.checkCollision(x[:1])[0])

    def checkTorqueConstraints(self, x, u):
        """env_model.py:179-182 for trajectories x [B, N+1, nx], u [B, N, nu] -> one flag per problem."""
        x, u = np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64)
        ok = np.ones(x.shape[0], dtype=bool)
        for k in range(u.shape[1]):
            ok &= self.checkTorqueBounds(self.tau_fun(x[:, k], u[:, k]))
        return ok

    def checkRunningConstraints(self, x, u):
        """env_model.py:189-190 per problem: state bounds on every node, collision on node 0 only (the early return of
        env_model.py:237-244), torque bounds on every interval."""
        x = np.asarray(x, dtype=np.float64)
        ok = np.logical_and(x >= self.x_min - self.params.tol_x, x <= self.x_max + self.params.tol_x).all(axis=(1, 2))
        return ok & self.checkCollision(x[:, 0]) & self.checkTorqueConstraints(x, u)

    def integrate_controller_model(self, x, u):
        """env_model.py:212-224, rows of x / u: the nominal model, no noise; the torque is saturated and the acceleration
        recomputed only where it leaves the bounds (inside them the engine's M^-1 (tau - h) returns u to rounding)."""
        e = self.engine()
        e.set_plant_inertial(np.tile(self.data.inertial, (self.batch, 1, 1)))
        e.set_torque_noise(np.zeros((self.batch, self.nu)))
        try:
            x_next, u_app = e.plant_step(self._rows(x), self._rows(u))
        finally:
            e.set_plant_inertial(self.plant_inertial)
            e.set_torque_noise(self.torque_noise)
        return x_next, u_app

    def checkDynamicsConstraints(self, x, u):
        """env_model.py:226-234 per problem: roll the controls out with the controller model and compare the trajectories,
        ||x - x_sim||_F < tol_dyn sqrt(n + 1)."""
        x, u = np.asarray(x, dtype=np.float64), np.asarray(u, dtype=np.float64)
        n = u.shape[1]
        x_sim = np.zeros_like(x)
        x_sim[:, 0] = x[:, 0]
        for i in range(n):
            x_sim[:, i + 1], _ = self.integrate_controller_model(x_sim[:, i], u[:, i])
        return np.linalg.norm((x - x_sim).reshape(x.shape[0], -1), axis=1) < self.params.tol_dyn * np.sqrt(n + 1)

    def integrate(self, x, u):
        """env_model.py:192-206 for the whole batch: nominal torque + noise, clipped, forward dynamics of the perturbed
        plant, double-integrator update.  -> (x_next [B, nx], applied acceleration [B, nu])."""
        return self.engine().plant_step(self._rows(x), self._rows(u))

    def integrate_torque_rk4(self, x, tau, dt=None, sens=True):
        """Extension (no reference counterpart): rows of states / joint torques through one explicit RK4
        step of x' = [v; M(q)^-1 (tau - h(q, v))] on the controller model -> (x_next, A = d x_next / d x, B = d x_next / d tau),
        or x_next alone with sens=False (smpc_rk4_sens)."""
        return self.engine().rk4_sens(self._rows(x), self._rows(tau), float(self.params.dt if dt is None else dt), sens=sens)

    # ---- perturbed plants / noise (env_model.py:321-331, utils.py:126-171) -----------------------------------------
    def update_randomized_dynamics(self, inertial=None, noise_percent=None, seed=0, controller_name=None):
        """Per-problem plant parameters [B, nq, 10].  The reference reloads ``z1_randomized<name>.urdf`` per test; here the
        whole batch is set at once, either from explicit parameters or drawn like ``randomize_model`` (uniform +- percent)."""
        if inertial is None and controller_name is not None:
            # env_model.py:321-328: the reference reloads robots/.../<sys>_randomized<controller_name>.urdf; here one name per problem
            # ('noise<level>_<i>', mpc.py:106), the files of scripts/generate_urdf_noise.py
            from .urdf import URDF
            from .robot_model import nominal_link_inertials
            names = [controller_name] * self.batch if isinstance(controller_name, str) else list(controller_name)
            rows = []
            for nm in names:
                ln = nominal_link_inertials(URDF.from_xml_file(self.params.robot_urdf[:-5] + f'_randomized{nm}.urdf'))
                rows.append(self.data.chain.lump(ln['mass'], ln['com'], ln['inertia6'], ln['rpy']))
            inertial = np.stack(rows)
        if inertial is None:
            n = float(self.params.noise if noise_percent is None else noise_percent)
            links = randomized_link_inertials(self.data.nominal_links, n, n, n, self.batch, seed=seed)
            inertial = np.stack([self.data.chain.lump(links['mass'][i], links['com'][i], links['inertia6'][i], links['rpy'])
                                 for i in range(self.batch)])
        self.plant_inertial = np.ascontiguousarray(inertial, dtype=np.float64).reshape(self.batch, self.nq, 10)
        if self._engine is not None:
            self._engine.set_plant_inertial(self.plant_inertial)

    def reset_seed(self, seeds=None):
        """mpc.py:126-127 re-seeds with the test index every step, i.e. test i always sees the same torque-noise vector:
        one draw per problem, ``default_rng(problem index).normal(0, tau_max * control_noise / 100)`` (env_model.py:196,330-331)."""
        seeds = range(self.batch) if seeds is None else seeds
        scale = self.tau_max * (self.params.control_noise / 100.0)
        self.torque_noise = np.stack([np.random.default_rng(int(s)).normal(np.zeros(self.nu), scale, size=self.nu) for s in seeds])
        if self._engine is not None:
            self._engine.set_torque_noise(self.torque_noise)
