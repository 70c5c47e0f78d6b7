"""Helpers of the reference's src/safe_mpc/utils.py that the closed-loop scripts use."""
from __future__ import annotations

from .controller import (NaiveController, TerminalZeroVelocity, STController, HTWAController, RecedingController, RealReceding,
                         ControllerSafeSetEverywhere)


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
