"""Golden of the REFERENCE's own ``Parameters`` (reference src/safe_mpc/parser.py:60-315) evaluated on the REFERENCE's own config.yaml,
run in the build container (``/root/reference`` does not travel).  parser.py needs ``urdf_parser_py`` (absent) to read link names and joint origins of the robot file (absent as well); the
stdlib URDF reader of this repo (same attribute names) is put in ``sys.modules`` in its place and reads the synthetic Z1-like robot.  Everything else -- horizon, tolerances, weights, margins, obstacles, capsules, collision pairs --
is what the unmodified reference code computes from its shipped configuration.

    python tests/golden/make_ref_parameters.py    ->  tests/golden/ref_parameters.json
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def reference_parser():
    from safe_mpc_b200 import urdf as our_urdf

    class URDF:                                              # the robot file of the reference checkout is absent: the synthetic Z1-like robot of this repo
        @staticmethod
        def from_xml_file(path):
            return our_urdf.URDF.from_xml_file(os.path.join(ROOT, 'robots', os.path.relpath(path, os.path.join(REF, 'robots'))))
    m = types.ModuleType('urdf_parser_py'); mu = types.ModuleType('urdf_parser_py.urdf'); mu.URDF = URDF; m.urdf = mu
    sys.modules['urdf_parser_py'] = m; sys.modules['urdf_parser_py.urdf'] = mu
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_parser', os.path.join(REF, 'src', 'safe_mpc', 'parser.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def plain(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    if isinstance(v, dict):
        return {k: plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [plain(x) for x in v]
    return v


def main():
    ref = reference_parser()
    out = {}
    for tag, argv in (('default', []), ('margins', ['--joint_bounds_margin', '5', '--collision_margin', '0.02', '--noise', '10', '-c', 'receding', '--horizon', '35', '--alpha', '20'])):
        sys.argv = ['x'] + argv
        args = ref.parse_args()
        p = ref.Parameters(args, 'z1', rti=True, filename=os.path.join(REF, 'config.yaml'))
        keep = {}
        for k, v in vars(p).items():
            if k in ('robot_descr', 'links', 'joints', 'act_fun') or k.endswith('_DIR') or k in ('robot_urdf',):
                continue
            try:
                json.dumps(plain(v)); keep[k] = plain(v)
            except TypeError:
                pass
        out[tag] = {'args': plain(args), 'params': keep}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_parameters.json')
    json.dump(out, open(path, 'w'), indent=1, sort_keys=True)
    print('wrote', path, {t: len(o['params']) for t, o in out.items()})


if __name__ == '__main__':
    main()
