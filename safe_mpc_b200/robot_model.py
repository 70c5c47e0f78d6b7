"""URDF -> flat serial-chain description consumed by the CUDA engine (and by the test oracle).

Replaces what the reference obtains from adam's ``KinDynComputations(urdf, joint_names, root)``
(reference env_model.py:23-53): the first ``nq`` non-fixed URDF joints are actuated, every other joint is
locked at zero and the links rigidly attached to an actuated joint's child link are lumped into one body
(mass, centre of mass, inertia about the CoM, all in the body = joint-child frame).  Per-problem
perturbed inertial parameters (reference utils.py:126-171) go through the same lumping, vectorised
over the batch, instead of through ``z1_randomized*.urdf`` files.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .urdf import URDF, rpy_to_matrix

GRAVITY = 9.80665  # adam default


def _hom(R, p):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = p
    return T


def rot_mat_x(t):
    return np.array([[1, 0, 0, 0], [0, np.cos(t), -np.sin(t), 0], [0, np.sin(t), np.cos(t), 0], [0, 0, 0, 1.]])


def rot_mat_y(t):
    return np.array([[np.cos(t), 0, np.sin(t), 0], [0, 1, 0, 0], [-np.sin(t), 0, np.cos(t), 0], [0, 0, 0, 1.]])


def rot_mat_z(t):
    return np.array([[np.cos(t), -np.sin(t), 0, 0], [np.sin(t), np.cos(t), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.]])


@dataclass
class Chain:
    """Serial chain with ``nq`` actuated revolute joints."""
    nq: int
    joint_names: List[str]
    joint_R: np.ndarray        # [nq,3,3] rotation parent-body <- joint frame at q=0
    joint_p: np.ndarray        # [nq,3]   joint origin in the parent body frame
    joint_axis: np.ndarray     # [nq,3]
    link_body: Dict[str, int]  # link name -> body index (-1: rigidly attached to the base)
    link_T: Dict[str, np.ndarray]  # link name -> 4x4 transform body <- link
    inertial_links: List[str]  # links that carry an <inertial>, in URDF order
    lower: np.ndarray
    upper: np.ndarray
    velocity: np.ndarray
    effort: np.ndarray

    def lump(self, mass, com, inertia6, rpy=None):
        """Lump per-link inertial parameters into per-body ones.

        mass [..., L], com [..., L, 3], inertia6 [..., L, 6] (ixx,iyy,izz,ixy,iyz,ixz) with L the
        number of links in ``inertial_links``; leading dims are batch dims.  Returns [..., nq, 10]:
        m, c(3), Ixx,Iyy,Izz,Ixy,Iyz,Ixz about the body CoM, body frame.
        """
        mass = np.asarray(mass, dtype=np.float64)
        com = np.asarray(com, dtype=np.float64)
        inertia6 = np.asarray(inertia6, dtype=np.float64)
        lead = mass.shape[:-1]
        M = np.zeros(lead + (self.nq,))
        S = np.zeros(lead + (self.nq, 3))          # first moment about the body origin
        Io = np.zeros(lead + (self.nq, 3, 3))      # inertia about the body origin
        eye = np.eye(3)
        for li, name in enumerate(self.inertial_links):
            b = self.link_body[name]
            if b < 0:
                continue
            T = self.link_T[name]
            R = T[:3, :3]
            if rpy is not None:
                R = R @ rpy_to_matrix(rpy[li])
            m = mass[..., li]
            r = com[..., li, :] @ T[:3, :3].T + T[:3, 3]
            i6 = inertia6[..., li, :]
            I = np.zeros(lead + (3, 3))
            I[..., 0, 0], I[..., 1, 1], I[..., 2, 2] = i6[..., 0], i6[..., 1], i6[..., 2]
            I[..., 0, 1] = I[..., 1, 0] = i6[..., 3]
            I[..., 1, 2] = I[..., 2, 1] = i6[..., 4]
            I[..., 0, 2] = I[..., 2, 0] = i6[..., 5]
            Irot = R @ I @ R.T
            rr = np.einsum('...i,...j->...ij', r, r)
            r2 = np.sum(r * r, axis=-1)[..., None, None]
            M[..., b] += m
            S[..., b, :] += m[..., None] * r
            Io[..., b, :, :] += Irot + m[..., None, None] * (r2 * eye - rr)
        c = S / M[..., None]
        cc = np.einsum('...i,...j->...ij', c, c)
        c2 = np.sum(c * c, axis=-1)[..., None, None]
        Ic = Io - M[..., None, None] * (c2 * eye - cc)
        out = np.zeros(lead + (self.nq, 10))
        out[..., 0] = M
        out[..., 1:4] = c
        out[..., 4] = Ic[..., 0, 0]
        out[..., 5] = Ic[..., 1, 1]
        out[..., 6] = Ic[..., 2, 2]
        out[..., 7] = Ic[..., 0, 1]
        out[..., 8] = Ic[..., 1, 2]
        out[..., 9] = Ic[..., 0, 2]
        return out


def build_chain(robot: URDF, nq: int) -> Tuple[Chain, dict]:
    """Select the first ``nq`` non-fixed joints (reference env_model.py:23-32) and lump the rest."""
    actuated = [j for j in robot.joints if j.type != 'fixed'][:nq]
    if len(actuated) != nq:
        raise ValueError(f'URDF has only {len(actuated)} movable joints, need {nq}')
    for j in actuated:
        if j.type not in ('revolute', 'continuous'):
            raise NotImplementedError(f'joint {j.name}: only revolute joints are supported (got {j.type})')
    act_idx = {j.name: i for i, j in enumerate(actuated)}
    child_joint = {j.child: j for j in robot.joints}

    link_body: Dict[str, int] = {}
    link_T: Dict[str, np.ndarray] = {}

    def resolve(link: str):
        if link in link_body:
            return
        if link not in child_joint:                      # root
            link_body[link], link_T[link] = -1, np.eye(4)
            return
        j = child_joint[link]
        if j.name in act_idx:
            link_body[link], link_T[link] = act_idx[j.name], np.eye(4)
            return
        resolve(j.parent)
        # locked / fixed joint: q = 0 -> pure origin transform
        link_body[link] = link_body[j.parent]
        link_T[link] = link_T[j.parent] @ _hom(rpy_to_matrix(j.origin.rpy), j.origin.xyz)

    for l in robot.links:
        resolve(l.name)

    joint_R = np.zeros((nq, 3, 3))
    joint_p = np.zeros((nq, 3))
    joint_axis = np.zeros((nq, 3))
    for i, j in enumerate(actuated):
        if link_body[j.parent] != i - 1:
            raise NotImplementedError('only serial chains are supported: joint '
                                      f'{j.name} hangs off body {link_body[j.parent]}, expected {i - 1}')
        T = link_T[j.parent] @ _hom(rpy_to_matrix(j.origin.rpy), j.origin.xyz)
        joint_R[i], joint_p[i] = T[:3, :3], T[:3, 3]
        joint_axis[i] = j.axis / np.linalg.norm(j.axis)

    inertial_links = [l.name for l in robot.links if l.inertial is not None]
    chain = Chain(nq, [j.name for j in actuated], joint_R, joint_p, joint_axis, link_body, link_T, inertial_links,
                  np.array([j.limit.lower for j in actuated]), np.array([j.limit.upper for j in actuated]),
                  np.array([j.limit.velocity for j in actuated]), np.array([j.limit.effort for j in actuated]))
    nominal = nominal_link_inertials(robot)
    return chain, nominal


def nominal_link_inertials(robot: URDF) -> dict:
    links = [l for l in robot.links if l.inertial is not None]
    return {'mass': np.array([l.inertial.mass for l in links]),
            'com': np.array([l.inertial.origin.xyz for l in links]),
            'inertia6': np.array([l.inertial.inertia for l in links]),
            'rpy': np.array([l.inertial.origin.rpy for l in links])}


def randomized_link_inertials(nominal: dict, noise_mass, noise_inertia, noise_cm, count: int, seed: int = 0) -> dict:
    """``count`` consecutive calls of the reference's ``randomize_model`` (utils.py:126-171) on one RNG stream.

    Draw order per call: for every link with an <inertial> (URDF order): mass, then ixx,iyy,izz,ixy,iyz,ixz,
    then x,y,z of the CoM; each ``uniform(-n, n)`` with n = |value|*percent/100.  ``generate_urdf_noise.py:32-36``
    calls it ``test_num`` times per noise level on ``default_rng(seed)`` (seed 0 for the first level).
    """
    rng = np.random.default_rng(seed)
    L = nominal['mass'].shape[0]
    mass = np.zeros((count, L))
    com = np.zeros((count, L, 3))
    inertia6 = np.zeros((count, L, 6))
    for t in range(count):
        for l in range(L):
            m = float(nominal['mass'][l])
            n = m * noise_mass / 100
            mass[t, l] = m + rng.uniform(-n, n)
            for k in range(6):
                v = float(nominal['inertia6'][l, k])
                n = abs(v) * noise_inertia / 100
                inertia6[t, l, k] = v + rng.uniform(-n, n)
            for k in range(3):
                e = float(nominal['com'][l, k])
                n = abs(e * noise_cm / 100)
                com[t, l, k] = e + rng.uniform(-n, n)
    return {'mass': mass, 'com': com, 'inertia6': inertia6, 'rpy': nominal['rpy']}


def point_on_link(chain: Chain, link_name: str, local_h) -> Tuple[int, np.ndarray]:
    """(body index, point in body frame) of a homogeneous point given in ``link_name``'s frame."""
    if link_name not in chain.link_body:
        raise ValueError(f'unknown link {link_name}')
    p = chain.link_T[link_name] @ np.asarray(local_h, dtype=np.float64)
    return chain.link_body[link_name], p[:3]


def capsule_local_points(capsule) -> List[np.ndarray]:
    """End points of a robot capsule in its link frame (reference env_model.py:134-147)."""
    rot = np.eye(4)
    if capsule.get('rotation_offset') is not None:
        th = capsule['rotation_offset']
        rot = rot_mat_x(th[0]) @ rot_mat_y(th[1]) @ rot_mat_z(th[2])
    if capsule.get('spatial_offset') is not None:
        prism = np.eye(4)
        prism[:3, 3] = capsule['spatial_offset']
        rot = prism @ rot
    return [rot @ capsule['end_points'][0], rot @ capsule['end_points'][1]]


# ----------------------------------------------------------------------------------------------
# small numpy implementation of the chain algorithms -- used by the host-side AdamModel mirror for
# scalar convenience calls and by the tests as a third, independent implementation.
# ----------------------------------------------------------------------------------------------

def _rodrigues(axis, q):
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def fk_frames(chain: Chain, q):
    """World rotation/origin of every body frame."""
    R, o = np.eye(3), np.zeros(3)
    Rs, os_ = [], []
    for i in range(chain.nq):
        o = o + R @ chain.joint_p[i]
        R = R @ chain.joint_R[i] @ _rodrigues(chain.joint_axis[i], q[i])
        Rs.append(R)
        os_.append(o)
    return Rs, os_


def fk_point(chain: Chain, q, body: int, local):
    if body < 0:
        return np.asarray(local, dtype=np.float64)
    Rs, os_ = fk_frames(chain, q)
    return os_[body] + Rs[body] @ np.asarray(local)


def rnea(chain: Chain, inertial, q, v, a, gravity=GRAVITY):
    """Inverse dynamics tau = M(q) a + h(q, v) of the fixed-base chain (numpy, scalar)."""
    n = chain.nq
    w = np.zeros(3)
    wd = np.zeros(3)
    vd = np.array([0., 0., gravity])
    Rl, F, Nm = [], [], []
    for i in range(n):
        R = chain.joint_R[i] @ _rodrigues(chain.joint_axis[i], q[i])
        ax = chain.joint_axis[i]
        p = chain.joint_p[i]
        vd = R.T @ (vd + np.cross(wd, p) + np.cross(w, np.cross(w, p)))
        w_new = R.T @ w + ax * v[i]
        wd = R.T @ wd + ax * a[i] + np.cross(w_new, ax * v[i])
        w = w_new
        m, c = inertial[i, 0], inertial[i, 1:4]
        I = np.array([[inertial[i, 4], inertial[i, 7], inertial[i, 9]],
                      [inertial[i, 7], inertial[i, 5], inertial[i, 8]],
                      [inertial[i, 9], inertial[i, 8], inertial[i, 6]]])
        ac = vd + np.cross(wd, c) + np.cross(w, np.cross(w, c))
        Fi = m * ac
        Ni = I @ wd + np.cross(w, I @ w) + np.cross(c, Fi)
        Rl.append(R)
        F.append(Fi)
        Nm.append(Ni)
    tau = np.zeros(n)
    f = np.zeros(3)
    nn = np.zeros(3)
    for i in reversed(range(n)):
        f_i = F[i] + f
        n_i = Nm[i] + nn
        tau[i] = chain.joint_axis[i] @ n_i
        # express in the parent frame for the next (lower) body
        f = Rl[i] @ f_i
        nn = Rl[i] @ n_i + np.cross(chain.joint_p[i], f)
    return tau
