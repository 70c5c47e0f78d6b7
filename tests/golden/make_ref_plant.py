"""Golden of the REFERENCE's own plant step ``AdamModel.integrate`` (reference src/safe_mpc/env_model.py:192-206, row a11 of SURVEY.md
section 8), method body extracted with ``ast`` and executed UNMODIFIED, run in the build container.  The CasADi / adam functions it calls
(``tau_noisy_fun``, ``mass_noisy``, ``bias_noisy``, ``f_fun``) are stand-ins that return the oracle's torque, mass matrix and bias of the
perturbed plant and the double integrator of env_model.py:63-71; what the golden pins is what the method itself does with them: noise on
the torque, saturation at the torque limits, forward dynamics of the perturbed plant by a linear solve, integration, returned acceleration.

    python tests/golden/make_ref_plant.py         ->  tests/golden/ref_plant.npz
"""
import ast
import os
import sys
import types
from copy import deepcopy

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ENV = '/root/reference/src/safe_mpc/env_model.py'


def main():
    from tests.common import make_problem, random_states
    from oracle.oracle import Oracle
    cls = next(n for n in ast.parse(open(ENV).read()).body if isinstance(n, ast.ClassDef) and n.name == 'AdamModel')
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'integrate']
    mod = ast.Module(body=[ast.ClassDef(name='RefPlant', bases=[], keywords=[], body=keep, decorator_list=[])], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {'np': np, 'deepcopy': deepcopy}
    exec(compile(mod, ENV, 'exec'), ns)
    B = 48
    prob, params, md = make_problem('naive', N=10)
    orc = Oracle(prob, B, 1)
    rng = np.random.default_rng(5)
    pin = np.tile(md.inertial, (B, 1, 1)) * (1 + 0.15 * rng.uniform(-1, 1, (B, 5, 10)))
    orc.set_plant_inertial(pin)
    x = random_states(md, B, seed=8, vel_scale=0.8)
    u = rng.uniform(-60, 60, (B, 5))                        # large accelerations: part of the torques saturate
    tau_nom = orc.tau(x, u)
    xn = np.zeros((B, 10)); acc = np.zeros((B, 5)); sat = np.zeros(B, dtype=bool)
    for b in range(B):
        M, h = orc.mass_bias(b, x[b], nominal=False)
        m = ns['RefPlant']()
        m.nx, m.nq, m.nu = 10, 5, 5
        m.params = types.SimpleNamespace(dt=params.dt, control_noise=0.0)
        m.tau_min, m.tau_max = md.tau_min, md.tau_max
        m.rng = np.random.default_rng(b)
        m.tau_noisy_fun = lambda xx, uu, b=b: tau_nom[b].reshape(5, 1)           # built from the NOMINAL model upstream
        big = np.zeros((11, 11)); big[6:, 6:] = M
        hb = np.zeros((11, 1)); hb[6:, 0] = h              # column vectors, like the CasADi DM slices of the reference
        m.mass_noisy = lambda H, q, big=big: big
        m.bias_noisy = lambda H, q, vb, v, hb=hb: hb
        m.f_fun = lambda xx, uu, dt=params.dt: (lambda uv: np.hstack([xx[:5] + dt * xx[5:] + 0.5 * dt * dt * uv, xx[5:] + dt * uv]).reshape(10, 1))(np.asarray(uu).ravel())
        xn[b], a = m.integrate(x[b], u[b])
        acc[b] = np.asarray(a).ravel()
        sat[b] = bool(np.any((tau_nom[b] < md.tau_min) | (tau_nom[b] > md.tau_max)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_plant.npz')
    np.savez_compressed(path, x=x, u=u, pin=pin, xn=xn, acc=acc, sat=sat)
    print('wrote', path, 'saturated rows', int(sat.sum()), 'of', B)


if __name__ == '__main__':
    main()
