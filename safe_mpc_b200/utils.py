"""Helpers of the reference's src/safe_mpc/utils.py that the closed-loop scripts use."""
from __future__ import annotations

from .controller import (NaiveController, TerminalZeroVelocity, STController, HTWAController, RecedingController, RealReceding,
                         ControllerSafeSetEverywhere)


def get_ocp_acados(cont_name, model):
    """utils.py:46-62 -- the controller a GUESS is generated with: every network controller name maps to ``HTWAController`` (hard
    terminal viability row, ``checkGuess`` includes ``checkSafeConstraints(x_temp[-1])``), naive / zerovel to their own class.
    -> (controller, dict of the names), like the reference (guess_acados.py:42 iterates the dict to name the files)."""
    controllers = {'naive': NaiveController,
                   'zerovel': TerminalZeroVelocity,
                   'st': HTWAController,
                   'htwa': HTWAController,
                   'receding': HTWAController,
                   'real_receding': HTWAController,
                   'parallel': HTWAController,
                   'st_analytic': HTWAController,
                   'htwa_analytic': HTWAController,
                   'constraint_everywhere': HTWAController,
                   'receding_analytic': HTWAController,
                   'parallel_analytic': HTWAController}
    if cont_name in controllers:
        return controllers[cont_name](model), controllers
    raise ValueError(f'Controller {cont_name} not available')


def get_controller(cont_name, model):
    """utils.py:64-75 -- same keys as the reference (it has no 'stwa' / 'parallel' entries), same error."""
    controllers = {'naive': NaiveController,
                   'zerovel': TerminalZeroVelocity,
                   'st': STController,
                   'htwa': HTWAController,
                   'receding': RecedingController,
                   'real_receding': RealReceding,
                   'constraint_everywhere': ControllerSafeSetEverywhere}
    if cont_name in controllers:
        return controllers[cont_name](model)
    raise ValueError(f'Controller {cont_name} not available')
