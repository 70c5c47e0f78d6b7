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
        self.build_flag = True
        if self._pending_guess is not None:
            self.setGuess(*self._pending_guess)

    def resetHorizon(self, N):
        """controller.py:203-214.  The horizon is a creation parameter of an engine handle: changing it re-creates the handle."""
        N = int(N)
        if N != self.N or not self.build_flag:
            self.N = N
            self.model.params.N = N
            if self.build_flag:
                self.ocp_solver.close()
                self.build_controller()

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


class TerminalZeroVelocity(NaiveController):             # controller.py:295-317
    engine_name = 'zerovel'


class STController(NaiveController):                     # controller.py:319-361
    engine_name = 'st'


class STWAController(STController):                      # controller.py:364-393
    engine_name = 'stwa'


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
