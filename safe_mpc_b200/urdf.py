"""Minimal URDF reader (stdlib xml only).

The reference uses ``urdf_parser_py.urdf.URDF`` (reference parser.py:5,80) which is not
installed here; this module exposes the small subset of that API the hot path needs
(``robot.links[i].name``, ``robot.joints[i].{name,type,parent,child,origin.xyz,origin.rpy,axis,limit}``,
``robot.get_root()``) plus the inertial blocks that adam reads from the URDF.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


def _floats(text: Optional[str], n: int = 3) -> np.ndarray:
    if text is None:
        return np.zeros(n)
    vals = [float(v) for v in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} floats, got {text!r}")
    return np.array(vals, dtype=np.float64)


@dataclass
class Pose:
    xyz: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rpy: np.ndarray = field(default_factory=lambda: np.zeros(3))


@dataclass
class Inertial:
    mass: float
    origin: Pose
    # ixx, iyy, izz, ixy, iyz, ixz -- the draw order of randomize_model (reference utils.py:128)
    inertia: np.ndarray


@dataclass
class Limit:
    lower: float = 0.0
    upper: float = 0.0
    effort: float = 0.0
    velocity: float = 0.0


@dataclass
class Link:
    name: str
    inertial: Optional[Inertial] = None


@dataclass
class Joint:
    name: str
    type: str
    parent: str
    child: str
    origin: Pose
    axis: np.ndarray
    limit: Optional[Limit] = None


INERTIA_FIELDS = ("ixx", "iyy", "izz", "ixy", "iyz", "ixz")


class URDF:
    def __init__(self, name: str, links: List[Link], joints: List[Joint]):
        self.name = name
        self.links = links
        self.joints = joints
        self.link_map = {l.name: l for l in links}
        self.joint_map = {j.name: j for j in joints}

    @classmethod
    def from_xml_file(cls, path: str) -> "URDF":
        return cls.from_xml_string(open(path, "r", encoding="utf-8").read())

    @classmethod
    def from_xml_string(cls, text: str) -> "URDF":
        root = ET.fromstring(text)
        links, joints = [], []
        for le in root.findall("link"):
            ine = le.find("inertial")
            inertial = None
            if ine is not None:
                oe = ine.find("origin")
                pose = Pose(_floats(oe.get("xyz")) if oe is not None else np.zeros(3),
                            _floats(oe.get("rpy")) if oe is not None and oe.get("rpy") else np.zeros(3))
                ie = ine.find("inertia")
                inertial = Inertial(float(ine.find("mass").get("value")), pose,
                                    np.array([float(ie.get(k)) for k in INERTIA_FIELDS]))
            links.append(Link(le.get("name"), inertial))
        for je in root.findall("joint"):
            oe = je.find("origin")
            pose = Pose(_floats(oe.get("xyz")) if oe is not None and oe.get("xyz") else np.zeros(3),
                        _floats(oe.get("rpy")) if oe is not None and oe.get("rpy") else np.zeros(3))
            ae = je.find("axis")
            axis = _floats(ae.get("xyz")) if ae is not None else np.array([1.0, 0.0, 0.0])
            lim = je.find("limit")
            limit = None
            if lim is not None:
                limit = Limit(float(lim.get("lower", 0.0)), float(lim.get("upper", 0.0)),
                              float(lim.get("effort", 0.0)), float(lim.get("velocity", 0.0)))
            joints.append(Joint(je.get("name"), je.get("type"), je.find("parent").get("link"),
                                je.find("child").get("link"), pose, axis, limit))
        return cls(root.get("name", "robot"), links, joints)

    def get_root(self) -> str:
        children = {j.child for j in self.joints}
        roots = [l.name for l in self.links if l.name not in children]
        if len(roots) != 1:
            raise ValueError(f"URDF must have exactly one root link, found {roots}")
        return roots[0]


def rpy_to_matrix(rpy) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) @ Ry(pitch) @ Rx(roll)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])
