"""Batched mirror of the reference's controller classes (src/safe_mpc/controller.py): same class names, method names and
attributes, every per-problem quantity with a leading batch dimension ``B = model.batch``.

Where the reference object drives one ``AcadosOcpSolver`` (controller.py:247) and keeps its warm start, fail counter,
receding index and viable state in Python, these classes are thin handles on one engine instance
(``safe_mpc_b200.engine.Engine``, i.e. ``include/safe_mpc_b200.h``): the state machines themselves run on the GPU
(csrc/kernels.cu: ctrl_post1 / ctrl_post2, citing controller.py:274-284,375-388,448-498,524-565,651-660) so that B
problems step together without a host round trip per problem.  There is no CPU path.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .problem import build_problem


class AbstractController:
    engine_name = 'naive'          # key of problem.CONTROLLERS

    def __init__(self, model):
        self.model = model
        self.B = model.batch
        self.N = int(model.params.N)
        self.cost = None
        self.cost_kind = getattr(self, 'cost_kind_default', 'ext')
        self.build_flag = False
        self.ocp_solver = None     # the engine handle once built (stands where AcadosOcpSolver stands)
        self.time_fields = ['time_lin', 'time_sim', 'time_qp', 'time_qp_solver_call', 'time_glob', 'time_reg', 'time_tot']
        self._pending_guess = None
        self.reset_controller()

    # ---- build (controller.py:240-248) ----------------------------------------------------------------------------
    def set_cost(self, cost):
        cost.set_solver_cost(self)

    def build_controller(self, build=True, name=''):
        """Nothing is generated or compiled per controller: the CUDA library is built once (safe_mpc_b200.build) and the OCP is
        a struct of numbers.  ``build`` / ``name`` are accepted for source compatibility with mpc.py:52,72."""
        from .engine import Engine
        prob, keep = build_problem(self.model.params, self.engine_name, cost=self.cost_kind, N=self.N, model=self.model.data)
        self.ocp_solver = Engine(prob, self.B, self.model.device)
        self.ocp_solver._keepalive = (prob, keep)
        self.ocp_solver.set_plant_inertial(self.model.plant_inertial)
        self.ocp_solver.set_torque_noise(self.model.torque_noise)
        self._upload_trajectory()
        self.build_flag = True
        if self._pending_guess is not None:
            self.setGuess(*self._pending_guess)

    def _upload_trajectory(self):
        """The stage parameters p[0:3] of the reference: cost.traj[:, current_step + i] on stage i of every solve (controller.py:153-156).
        A tracking cost hands its whole path to the engine, which indexes it with the per-problem step counter; the reach costs keep
        the constant ee_ref of the problem struct."""
        if self.cost is not None and getattr(self.cost, 'tracking', False):
            self.ocp_solver.set_ee_trajectory(np.ascontiguousarray(self.cost.traj.T))

    def resetHorizon(self, N):
        """controller.py:203-214.  The horizon is a creation parameter of an engine handle: changing it re-creates the handle."""
        N = int(N)
        if N != self.N or not self.build_flag:
            self.N = N
            self.model.params.N = N
            if self.cost is not None:
                self.cost.update_trajectory()                # controller.py:214
            if self.build_flag:
                self.ocp_solver.close()
                self.build_controller()
        elif self.cost is not None and getattr(self.cost, 'tracking', False):
            self.cost.update_trajectory()
            self._upload_trajectory()

    def _solver(self):
        if not self.build_flag:
            raise ValueError('Controller not built, call build_controller() first')      # controller.py:137-138
        return self.ocp_solver

    # ---- warm start ----------------------------------------------------------------------------------------------------
    def setGuess(self, x_guess, u_guess):
        """controller.py:195-197 (STWA/HTWA also latch x_viable = x_guess[-1], :390-393 -- done by the engine)."""
        xg = np.ascontiguousarray(x_guess, dtype=np.float64).reshape(self.B, self.N + 1, abi.NX)
        ug = np.ascontiguousarray(u_guess, dtype=np.float64).reshape(self.B, self.N, abi.NU)
        if not self.build_flag:
            self._pending_guess = (xg, ug)
            return
        self.ocp_solver.set_guess(xg, ug)

    def getGuess(self):
        return self._solver().get_guess()

    @property
    def x_guess(self):
        return self._solver().get_guess()[0]

    @property
    def u_guess(self):
        return self._solver().get_guess()[1]

    @property
    def x_temp(self):
        return self._solver().get_temp()[0]

    @property
    def u_temp(self):
        return self._solver().get_temp()[1]

    # ---- per-problem state ----------------------------------------------------------------------------------------------
    @property
    def fails(self):
        return self._solver().get_state(abi.STATE_FAILS)

    @property
    def r(self):
        return self._solver().get_state(abi.STATE_R)

    @property
    def last_status(self):
        return self._solver().get_state(abi.STATE_STATUS)

    def getLastViableState(self):
        return self._solver().get_x_viable()

    @property
    def x_viable(self):
        return self.getLastViableState()

    def reset_controller(self):
        """controller.py:234-238 (fails = 0, current_step = 0; receding: r = N, :444-446)."""
        self.current_step = 0
        if self.build_flag:
            self.ocp_solver.reset_controller()

    # ---- the hot path -------------------------------------------------------------------------------------------------------
    def solve(self, x0, active=None):
        """AbstractController.solve (controller.py:136-167) for every problem: one RTI iteration at the stored guess.
        -> acados status per problem (0 ok, 1 NaN, 2 max-iter, 3 min-step, 4 QP failure)."""
        return self._solver().rti_solve(np.ascontiguousarray(x0, dtype=np.float64).reshape(self.B, abi.NX), active)

    def step(self, x, active=None):
        """controller.step(x) -> (u [B, nu], abort flag [B]) -- controller.py:274-284 and the per-class variants."""
        u, ab = self._solver().controller_step(np.ascontiguousarray(x, dtype=np.float64).reshape(self.B, abi.NX), active)
        self.current_step += 1
        return u, ab.astype(bool)

    def getTime(self):
        """controller.py:192-193: the acados time fields (seconds, CUDA events) of the last solve / step."""
        t = self._solver().times()
        return np.array([t[f] for f in self.time_fields])

    def checkSafeConstraints(self, x):
        """safe_set.check_constraint (safe_set.py:61-68): c(x) within [0 - tol, 1e6 + tol], rows of x."""
        c = self._solver().nn_constraint(np.atleast_2d(np.asarray(x, dtype=np.float64)), grad=False)
        tol = self.model.params.tol_safe_set
        return np.logical_and(c >= 0.0 - tol, c <= 1e6 + tol)


class NaiveController(AbstractController):               # controller.py:251-284
    engine_name = 'naive'

    def _all_nodes_collision_free(self, x):
        ok = np.ones(self.B, dtype=bool)
        for k in range(x.shape[1]):
            ok &= self.model.checkCollision(x[:, k])
        return ok

    def checkGuess(self, x=None, u=None):
        """controller.py:255-258, one flag per problem: running constraints, dynamics defect and collisions of (x_temp, u_temp)
        -- or of the trajectories passed in (the last iterate of ``solve_sqp``)."""
        x, u = (self.x_temp if x is None else x), (self.u_temp if u is None else u)
        return self.model.checkRunningConstraints(x, u) & self.model.checkDynamicsConstraints(x, u) & self._all_nodes_collision_free(x)

    def merit(self, x, u, mu):
        """l1 merit of a trajectory, per problem: the cost of cost_definition.py:91-96 (EE tracking + control effort, terminal
        EE tracking) + mu * (violation of the torque rows, the capsule rows and -- controllers with a terminal viability row --
        c(x_N) >= 0).  The dynamics and the state box are linear: a step between two points that satisfy them satisfies them."""
        p, m = self.model.params, self.model
        B, N = x.shape[0], u.shape[1]
        eng = m.engine()
        ee, dist = eng.kinematics(np.ascontiguousarray(x.reshape(-1, abi.NX)))
        e = ee.reshape(B, N + 1, 3) - np.asarray(p.ee_ref)
        cost = p.Q_weight * (e * e).sum(axis=(1, 2)) + p.R_weight * (u * u).sum(axis=(1, 2))
        tau = eng.tau(np.ascontiguousarray(x[:, :N].reshape(-1, abi.NX)), np.ascontiguousarray(u.reshape(-1, abi.NU))).reshape(B, N, abi.NU)
        viol = (np.maximum(m.tau_min - tau, 0.0) + np.maximum(tau - m.tau_max, 0.0)).sum(axis=(1, 2))
        lo = np.array([pr['lo_ocp'] for pr in m.data.pairs])
        viol += np.maximum(lo - dist.reshape(B, N + 1, -1), 0.0).sum(axis=(1, 2))
        if self.engine_name in ('st', 'stwa', 'htwa'):
            viol += np.maximum(-self._solver().nn_constraint(np.ascontiguousarray(x[:, -1]), grad=False), 0.0)
        return cost + mu * viol

    def solve_sqp(self, x0, max_iter=None, tol=1e-6, active=None, globalization=None):
        """The SQP solve of the guess generator (guess_acados.py builds its controllers with rti=False: acados 'SQP',
        nlp_max_iter): RTI iterations of the engine repeated per problem until the full step is below ``tol`` (infinity norm
        over the trajectory) -> status 0; a problem still moving after ``max_iter`` iterations -> status 2, as acados reports
        it; a QP failure keeps its status (1, 3, 4) and stops that problem.  Finished problems are frozen through the engine's
        ``active`` mask, so the batch shrinks as it converges.  ``globalization``: default = ``params.globalization``, i.e. what
        parser.py:139 selects -- 'FIXED_STEP' for rti=True, 'MERIT_BACKTRACKING' for the generator's Parameters(rti=False).
        'FIXED_STEP' takes full steps.  'MERIT_BACKTRACKING' shortens the step by ``alpha_reduction`` until the l1 merit
        (``merit``, penalty = the largest inequality multiplier of the QP, at least 1) decreases, down to ``alpha_min``, which
        is then taken regardless: the structure of acados' line search with a plain decrease test (acados' own merit weights
        and Armijo constant cannot be pinned offline).  -> status [B]; the last iterate of every problem (what acados' get(i, 'x')
        returns after its SQP solve) is ``self._sqp_result`` = getGuess(); (x_temp, u_temp) hold the last QP solution, the
        same thing when the last step was a full one."""
        p = self.model.params
        # parser.py:139: FIXED_STEP for rti=True, MERIT_BACKTRACKING for the SQP of the guess generator (rti=False)
        glob = getattr(p, 'globalization', 'FIXED_STEP') if globalization is None else globalization
        if glob not in ('FIXED_STEP', 'MERIT_BACKTRACKING'):
            raise ValueError(f'unknown globalization {glob}')
        max_iter = int(self.model.params.nlp_max_iter if max_iter is None else max_iter)
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(self.B, abi.NX)
        todo = np.ones(self.B, dtype=bool) if active is None else np.asarray(active, dtype=bool).copy()
        status = np.where(todo, 2, 0).astype(np.int32)
        xg, ug = (a.copy() for a in self.getGuess())
        self.sqp_iter = np.zeros(self.B, dtype=np.int32)
        for _ in range(max_iter):
            if not todo.any():
                break
            st = self.solve(x0, todo.astype(np.uint8))
            xt, ut = self.x_temp, self.u_temp
            self.sqp_iter[todo] += 1
            bad = todo & (st != 0)
            status[bad] = st[bad]
            good = todo & (st == 0)
            step = np.maximum(np.abs(xt - xg).reshape(self.B, -1).max(axis=1), np.abs(ut - ug).reshape(self.B, -1).max(axis=1))
            if glob == 'MERIT_BACKTRACKING' and good.any():
                lam = self._solver().get_qp()[2]
                mu = np.maximum(1.0, lam.reshape(self.B, -1).max(axis=1))
                phi0 = self.merit(xg, ug, mu)
                alpha = np.ones(self.B)
                pend = good & (step >= tol)
                while pend.any():
                    a = alpha[:, None, None]
                    phi = self.merit(xg + a * (xt - xg), ug + a * (ut - ug), mu)
                    pend &= ~(phi < phi0)
                    nxt = alpha * p.alpha_reduction
                    stop = pend & (nxt < p.alpha_min)
                    alpha[stop] = p.alpha_min
                    pend &= ~stop
                    alpha[pend] = nxt[pend]
                self.sqp_alpha = alpha
                a = alpha[:, None, None]
                xt, ut = xg + a * (xt - xg), ug + a * (ut - ug)
            xg[good], ug[good] = xt[good], ut[good]
            conv = good & (step < tol)
            status[conv] = 0
            todo &= ~(bad | conv)
            self.setGuess(xg, ug)
        self._sqp_result = (xg, ug)
        return status

    def initialize(self, x0, u0=None):
        """controller.py:260-272 for the batch: trivial guess, solve, keep the solution as the guess where it succeeded and
        passes checkGuess -> 1 / 0 per problem."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(self.B, abi.NX)
        xg = np.repeat(x0[:, None, :], self.N + 1, axis=1)
        ug = np.zeros((self.B, self.N, abi.NU)) if u0 is None else np.broadcast_to(np.asarray(u0, dtype=np.float64), (self.B, self.N, abi.NU)).copy()
        self.setGuess(xg, ug)
        status = self.solve(x0)
        ok = (status == 0) & self.checkGuess()
        xt, ut = self.x_temp, self.u_temp
        xg[ok], ug[ok] = xt[ok], ut[ok]
        self.setGuess(xg, ug)
        return ok.astype(np.int32)


class TerminalZeroVelocity(NaiveController):             # controller.py:295-317
    engine_name = 'zerovel'


class STController(NaiveController):                     # controller.py:319-361
    engine_name = 'st'


class STWAController(STController):                      # controller.py:364-393
    engine_name = 'stwa'

    def checkGuess(self, x=None, u=None):
        """controller.py:369-373: the checks of NaiveController plus the viability constraint at the terminal node."""
        x = self.x_temp if x is None else x
        return super().checkGuess(x, u) & self.checkSafeConstraints(x[:, -1])


class HTWAController(STWAController):                    # controller.py:396-401
    engine_name = 'htwa'


class RecedingController(STWAController):                # controller.py:404-502
    engine_name = 'receding'


class RealReceding(STWAController):                      # controller.py:504-565
    engine_name = 'real_receding'


class ParallelController(RecedingController):            # controller.py:567-644
    """One solve per candidate node n = N .. 1 in every step (sing_step, controller.py:596-612), the best node kept; the engine batches
    each candidate over the problems that have not reached n = N yet.  Like the reference, ``get_controller`` has no key for it
    (utils.py:64-75): construct it directly."""
    engine_name = 'parallel'


class ControllerSafeSetEverywhere(STWAController):       # controller.py:646-689
    engine_name = 'constraint_everywhere'


class SafeBackupController(AbstractController):          # controller.py:692-712
    engine_name = 'backup'
    cost_kind_default = 'zero'
