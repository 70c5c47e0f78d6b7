#!/usr/bin/env python
"""Perturbed plant models -- the reference's scripts/generate_urdf_noise.py on this repository's model layer.

For every noise level the reference calls ``randomize_model`` ``test_num`` times on one RNG stream and writes
``robots/<sys>_description/urdf/<sys>_randomizednoise<level>_<i>.urdf`` (utils.py:126-171, generate_urdf_noise.py:32-36); after a
level it re-seeds the stream with ``reset_rng1(i + 1)`` = ``test_num`` (generate_urdf_noise.py:36), so the first level draws from
``default_rng(0)`` and every later one from ``default_rng(test_num)``.  The same files are written here, from
``robot_model.randomized_link_inertials`` (the draw order of randomize_model, bit for bit: tests/test_ref_golden.py).
scripts/mpc.py --noise <level> reads them when they exist and draws the same numbers itself when they do not.

    python scripts/generate_urdf_noise.py [--noises 0.1 1.3 2.5 3.7 5.0 10.0 15.0 20.0 25.0 30.0]
"""
import os
import sys
import xml.etree.ElementTree as ET

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from safe_mpc_b200.parser import Parameters, build_arg_parser          # noqa: E402
from safe_mpc_b200.robot_model import nominal_link_inertials, randomized_link_inertials   # noqa: E402
from safe_mpc_b200.urdf import INERTIA_FIELDS                           # noqa: E402

NOISES = [0.1, 1.3, 2.5, 3.7, 5.0, 10.0, 15.0, 20.0, 25.0, 30.0]        # generate_urdf_noise.py:20


def level_seed(level_index, test_num):
    """default_rng(0) at import (utils.py:19,24), reset_rng1(test_num) after every level (generate_urdf_noise.py:36)."""
    return 0 if level_index == 0 else test_num


def write_perturbed_urdf(src_path, dst_path, mass, com, inertia6):
    """The nominal URDF with the <inertial> of every link that has one replaced (URDF order), as randomize_model writes it."""
    tree = ET.parse(src_path)
    k = 0
    for link in tree.getroot().findall('link'):
        inertial = link.find('inertial')
        if inertial is None:
            continue
        inertial.find('mass').set('value', str(float(mass[k])))
        ie = inertial.find('inertia')
        for name, v in zip(INERTIA_FIELDS, inertia6[k]):
            ie.set(name, str(float(v)))
        inertial.find('origin').set('xyz', ' '.join(str(float(v)) for v in com[k]))
        k += 1
    tree.write(dst_path, encoding='utf-8', xml_declaration=True)


def main(argv=None):
    ap = build_arg_parser()
    ap.add_argument('--noises', type=float, nargs='*', default=None, help='noise levels in percent (default: the list of the reference script)')
    args = vars(ap.parse_args(argv))
    params = Parameters(args, args['system'], rti=True)
    nominal = nominal_link_inertials(params.robot_descr)
    levels = NOISES if args['noises'] is None else args['noises']
    written = []
    for li, noise in enumerate(levels):
        links = randomized_link_inertials(nominal, noise, noise, noise, params.test_num, seed=level_seed(li, params.test_num))
        for i in range(params.test_num):
            dst = params.robot_urdf[:-5] + f'_randomizednoise{noise}_{i}.urdf'
            write_perturbed_urdf(params.robot_urdf, dst, links['mass'][i], links['com'][i], links['inertia6'][i])
            written.append(dst)
        print(f'noise {noise} %: {params.test_num} models, seed {level_seed(li, params.test_num)}')
    return written


if __name__ == '__main__':
    main()
