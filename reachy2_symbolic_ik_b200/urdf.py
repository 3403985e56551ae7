"""URDF -> IK parameter dictionary (host side, runs once per ControlIK).

Same contract as the reference's ``get_ik_parameters_from_urdf`` (utils.py:661-694): for each
requested arm prefix it reads the origins of ``{p}_shoulder_base_joint``, ``{p}_elbow_base_joint``,
``{p}_wrist_base_joint`` and ``{p}_tip_joint``.  Keys and float round-trips are kept (the shoulder
roll goes through ``rpy -+ pi/2`` then ``np.degrees``, which is why ControlIK sees
-14.999999999999996 where bare SymbolicIK uses -15).
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import Any

import numpy as np


def parse_vector(vector_str: str) -> np.ndarray:
    return np.array([float(tok) for tok in vector_str.split()])


def get_ik_parameters_from_urdf(urdf_str: str, arm: list) -> dict:
    root = ET.fromstring(urdf_str)
    out: dict[str, Any] = {}
    wanted = {}
    for p in arm:
        wanted[f"{p}_shoulder_base_joint"] = (p, "shoulder")
        wanted[f"{p}_elbow_base_joint"] = (p, "elbow")
        wanted[f"{p}_wrist_base_joint"] = (p, "wrist")
        wanted[f"{p}_tip_joint"] = (p, "tip")
    for joint in root.findall("joint"):
        hit = wanted.get(joint.attrib["name"])
        if hit is None:
            continue
        p, what = hit
        origin = joint.find("origin").attrib
        xyz = parse_vector(origin["xyz"])
        if what == "shoulder":
            rpy = parse_vector(origin["rpy"])
            rpy[0] += -np.pi / 2 if p == "r" else np.pi / 2
            out[f"{p}_shoulder_position"] = xyz
            out[f"{p}_shoulder_orientation"] = np.degrees(rpy)
        elif what == "elbow":
            out[f"{p}_upper_arm_size"] = xyz[2]
            out[f"{p}_elbow_roll_offset"] = -xyz[0]
        elif what == "wrist":
            out[f"{p}_forearm_size"] = xyz[2]
            out[f"{p}_wrist_pitch_offset"] = -xyz[1]
        else:
            out[f"{p}_tip_position"] = xyz
    return out
