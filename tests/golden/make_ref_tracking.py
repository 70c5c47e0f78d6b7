"""Golden of the REFERENCE's own tracking paths: ``generate_8shape_trajectory`` and ``generate_moving_circle_trajectory`` (reference
src/safe_mpc/cost_definition.py:181-209,262-286), the arrays ``cost.traj`` from which AbstractController.solve takes the stage
parameters p[0:3] = traj[:, current_step + i] (controller.py:153-156).  cost_definition.py cannot be imported (casadi), so the two
function definitions are taken out of the file with ``ast`` and executed UNMODIFIED; their ``from .utils import rot_mat_x, ...`` is
served by a stand-in package holding the rotation helpers of the reference's utils.py (extracted the same way).  Parameters: the tracking
keys of the reference's own config.yaml, with vel_const true (as shipped) and false, N = 45, dt = 5 ms.

    python tests/golden/make_ref_tracking.py     ->  tests/golden/ref_tracking.npz
"""
import ast
import os
import sys
import types

import numpy as np
import sympy as sym
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = '/root/reference'


def extract(path, names, ns):
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in keep} == set(names)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, 'exec'), ns)
    return ns


def main():
    utils_ns = extract(f'{REF}/src/safe_mpc/utils.py', ['rot_mat_x', 'rot_mat_y', 'rot_mat_z'], {'np': np})
    pkg = types.ModuleType('refpkg'); pkg.__path__ = []
    mod = types.ModuleType('refpkg.utils')
    for k in ('rot_mat_x', 'rot_mat_y', 'rot_mat_z'):
        setattr(mod, k, utils_ns[k])
    sys.modules['refpkg'] = pkg; sys.modules['refpkg.utils'] = mod
    ns = {'np': np, 'sym': sym, '__package__': 'refpkg', '__name__': 'refpkg.cost_definition'}
    extract(f'{REF}/src/safe_mpc/cost_definition.py', ['generate_8shape_trajectory', 'generate_moving_circle_trajectory'], ns)
    cfg = yaml.load(open(f'{REF}/config.yaml'), Loader=yaml.FullLoader)
    out = {}
    for tag, vel_const, n_track, circle_vel in (('const', True, 600, 0.25), ('accel', False, 400, 0.25)):
        p = types.SimpleNamespace(
            N=45, dt=float(cfg['dt']), n_steps=n_track, n_steps_tracking=n_track, dim_shape_8=float(cfg['dim_shape_8']),
            offset_traj=np.array(cfg['offset_traj']), theta_rot_traj=np.array(cfg['theta_rot_traj']), vel_max_traj=float(cfg['vel_max_traj']),
            vel_const=vel_const, acc_time=float(cfg['acc_time']), circle_rad=float(cfg['circle_rad']),
            circle_offset_traj=np.array(cfg['circle_offset_traj']), circle_traj_vel=circle_vel, circle_center_vel=float(cfg['circle_center_vel']))
        out[f'eight_{tag}'] = ns['generate_8shape_trajectory'](p)
        out[f'circle_{tag}'] = ns['generate_moving_circle_trajectory'](p)
        out[f'n_track_{tag}'] = n_track
        out[f'circle_vel_{tag}'] = circle_vel
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ref_tracking.npz'), **out)
    for k, v in out.items():
        if hasattr(v, 'shape') and v.ndim == 2:
            print(k, v.shape, v[:, 0], v[:, -1])


if __name__ == '__main__':
    main()
