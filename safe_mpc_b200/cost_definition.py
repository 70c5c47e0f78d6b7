"""Cost objects of the reference (src/safe_mpc/cost_definition.py), reduced to what the engine needs: which of the
three cost kinds the stage records are linearised with.  The expressions themselves live in the CUDA linearisation
kernel (csrc/dev_model.cuh: EE position, Jacobian and the exact second-order term of the EXTERNAL cost)."""
from __future__ import annotations


class _Cost:
    kind = 'ext'

    def __init__(self, model, Q_weight=None, R_weight=None):
        self.model = model
        # the weights are read from params by build_problem; explicit arguments override them like the reference's
        if Q_weight is not None:
            model.params.Q_weight = float(Q_weight)
        if R_weight is not None:
            model.params.R_weight = float(R_weight)

    def set_solver_cost(self, controller):
        """cost_definition.py:17-31: attach this cost to a controller before build_controller()."""
        controller.cost = self
        controller.cost_kind = self.kind


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
