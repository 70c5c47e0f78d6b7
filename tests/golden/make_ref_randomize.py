"""Golden of the REFERENCE's own ``randomize_model`` (reference src/safe_mpc/utils.py:126-171, row a13 of SURVEY.md section 8) and of its
rotation helpers ``rot_mat_x/y/z`` (utils.py:77-93) and its capsule distance ``casadi_segment_dist`` (utils.py:94-113, row a4), run in the
build container.  utils.py cannot be imported (meshcat, pinocchio, acados,
casadi), so the function definitions are taken out of the file with ``ast`` and executed UNMODIFIED in a namespace that holds what they
reference (numpy, ElementTree, and the ``RandomGenerator`` class of the same file for ``rng1``).  Input robot: the synthetic Z1-like URDF of
this repo (the reference's robot file is absent); calls: as scripts/generate_urdf_noise.py:32-36 (``reset_rng1(seed)``, then
``test_num`` consecutive calls).

    python tests/golden/make_ref_randomize.py     ->  tests/golden/ref_randomize.npz
"""
import ast
import os
import shutil
import sys
import tempfile
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SRC = '/root/reference/src/safe_mpc/utils.py'


def reference_namespace(names):
    tree = ast.parse(open(SRC).read())
    ns = {'np': np, 'ET': ET}
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert {n.name for n in keep} == set(names)
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, 'exec'), ns)
    return ns


def main():
    ns = reference_namespace(['RandomGenerator', 'randomize_model', 'reset_rng1', 'rot_mat_x', 'rot_mat_y', 'rot_mat_z'])
    ns['rng1'] = ns['RandomGenerator']()
    from safe_mpc_b200 import urdf as U
    from safe_mpc_b200.robot_model import nominal_link_inertials
    src = os.path.join(ROOT, 'robots', 'z1_description', 'urdf', 'z1.urdf')
    tmp = tempfile.mkdtemp()
    base = os.path.join(tmp, 'z1.urdf')
    shutil.copy(src, base)
    out = {}
    for seed, pct, count in ((1, 5.0, 3), (2, 20.0, 2)):
        ns['reset_rng1'](seed)                               # generate_urdf_noise.py:36
        mass, com, inertia = [], [], []
        for t in range(count):
            ns['randomize_model'](base, noise_mass=pct, noise_inertia=pct, noise_cm_position=pct, controller_name=f'_{t}')
            nl = nominal_link_inertials(U.URDF.from_xml_file(base[:-5] + f'_randomized_{t}.urdf'))
            mass.append(nl['mass']); com.append(nl['com']); inertia.append(nl['inertia6'])
        out[f'mass_{seed}'] = np.array(mass); out[f'com_{seed}'] = np.array(com); out[f'inertia6_{seed}'] = np.array(inertia)
        out[f'pct_{seed}'] = pct
    th = np.array([0.3, -1.1, 2.0])
    out['theta'] = th
    out['rot_x'] = np.array([ns['rot_mat_x'](t) for t in th]); out['rot_y'] = np.array([ns['rot_mat_y'](t) for t in th])
    out['rot_z'] = np.array([ns['rot_mat_z'](t) for t in th])
    # ---- capsule distance (row a4): the reference's casadi_segment_dist (utils.py:94-113), body unmodified, on numbers.  It is written
    # against the casadi namespace `cs`; the three functions it uses (sum1, fmin, fmax) are given their numpy meaning.
    class _cs:
        sum1 = staticmethod(lambda v: np.sum(v, axis=0))
        fmin = staticmethod(np.minimum)
        fmax = staticmethod(np.maximum)
    ns2 = reference_namespace(['casadi_segment_dist'])
    ns2['cs'] = _cs
    from tests.common import params_model, random_states
    from safe_mpc_b200 import robot_model
    params, md = params_model()
    xs = random_states(md, 64, seed=9)
    d = np.zeros((len(xs), len(md.pairs)))
    for i, x in enumerate(xs):
        for p_, pr in enumerate(md.pairs):
            A = robot_model.fk_point(md.chain, x[:5], md.point_body[pr['pa']], md.point_local[pr['pa']])
            B = robot_model.fk_point(md.chain, x[:5], md.point_body[pr['pb']], md.point_local[pr['pb']])
            d[i, p_] = float(ns2['casadi_segment_dist'](A, B, np.asarray(pr['C'], dtype=float), np.asarray(pr['D'], dtype=float)))
    out['dist_x'] = xs; out['dist'] = d
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_randomize.npz')
    np.savez_compressed(path, **out)
    shutil.rmtree(tmp)
    print('wrote', path, {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
