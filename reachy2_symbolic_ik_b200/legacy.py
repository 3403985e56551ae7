"""Host-side elbow-angle policies of the reference that only its example scripts still call
(``src/example/placo/movement_memory.py:51``, ``test_limits.py:31``): ``get_best_continuous_theta`` -- the first
version of the continuous-mode policy, ``utils.py:130-217`` -- and ``tend_to_preferred_theta`` (``utils.py:115-127``).

They are scalar decision rules around a ``get_elbow_position(theta)`` callback; with this package the callback is
``SymbolicIK.get_elbow_position`` (one launch of the scalar kernel per call), so nothing here solves IK on the host.
ControlIK itself uses the second version of the policy, which runs inside the continuous-mode kernels
(``csrc/r2ik_control.cuh``).  Results -- flag, theta and the debug text -- are pinned to the reference's by
``tests/golden/legacy_theta.npz``.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Sequence, Tuple

import numpy as np

TWO_PI = 2.0 * math.pi


def angle_diff(a: float, b: float) -> float:
    """Signed difference a - b wrapped to [-pi, pi) with Python's float modulo (``utils.py:486-490``)."""
    return ((a - b) + math.pi) % TWO_PI - math.pi


def is_valid_angle(angle: float, interval: Sequence[float]) -> bool:
    """Membership in a theta interval that may wrap; ends equal modulo 2 pi mean the full circle (``utils.py:468-474``)."""
    lo, hi = interval[0], interval[1]
    if lo % TWO_PI == hi % TWO_PI:
        return True
    inside_lo, inside_hi = bool(lo <= angle), bool(angle <= hi)
    return (inside_lo and inside_hi) if lo < hi else (inside_lo or inside_hi)


def is_elbow_ok(elbow_position, side: int, singularity_offset: float, singularity_limit_coeff: float,
                elbow_singularity_position) -> bool:
    """The predicate ``utils.py:443-465`` reduces to (its first test is overwritten before it is read): the elbow stays
    0.2 m outside the torso side plane and below the singularity-limit line in the (x, z) plane."""
    outside = bool(elbow_position[1] * side < -0.2)
    limit_z = (elbow_position[0] - elbow_singularity_position[0]) * singularity_limit_coeff \
        + elbow_singularity_position[2] - singularity_offset
    return outside and bool(elbow_position[2] < limit_z)


def _step(previous_theta: float, target: float, d_theta_max: float):
    """(arrived, theta): ``target`` if it is closer than d_theta_max, else previous_theta moved by d_theta_max toward it.
    The sign is the reference's quotient diff / |diff| (a float, printed in the debug text)."""
    diff = angle_diff(target, previous_theta)
    if abs(diff) < d_theta_max:
        return True, target, None
    sign = diff / np.abs(diff)
    return False, previous_theta + sign * d_theta_max, sign


def tend_to_preferred_theta(previous_theta: float, interval, get_joints: Any, d_theta_max: float,
                            goal_theta: float = -np.pi * 5 / 4) -> Tuple[bool, float]:
    """``utils.py:115-127``: one rate-limited step toward ``goal_theta`` (``interval`` / ``get_joints`` are unused there too)."""
    arrived, theta, _ = _step(previous_theta, goal_theta, d_theta_max)
    return arrived, theta


def get_best_continuous_theta(previous_theta: float, interval, get_elbow_position: Callable[[float], Any],
                              d_theta_max: float, preferred_theta: float, arm: str, singularity_offset: float,
                              singularity_limit_coeff: float, elbow_singularity_position) -> Tuple[bool, float, str]:
    """``utils.py:130-217``.  Aim at the middle of the interval (the preferred theta when the whole circle is possible):
    reachable and near -> take it; reachable but far -> one step toward it, valid only if that point is acceptable and
    inside the interval; middle not acceptable -> stay if the previous theta is acceptable, else one step toward the
    preferred theta.  Returns (flag, theta, debug text) like the reference."""
    side = -1 if arm == "l_arm" else 1

    def acceptable(theta: float) -> bool:
        return is_elbow_ok(get_elbow_position(theta), side, singularity_offset, singularity_limit_coeff,
                           elbow_singularity_position)

    lines = [f"{arm}", f"interval: {interval}"]
    if abs(abs(interval[0]) + abs(interval[1]) - TWO_PI) < 0.00001:
        lines.append("All the circle is possible.")
        middle = preferred_theta
    else:
        middle = (interval[0] + interval[1]) / 2
        if interval[0] > interval[1]:
            middle -= np.pi
    lines += [f"theta milieu {middle}", f"angle diff {angle_diff(middle, previous_theta)}"]

    if acceptable(middle):
        arrived, theta, sign = _step(previous_theta, middle, d_theta_max)
        if arrived:
            lines.append("theta milieu ok et proche")
            return True, middle, "\n".join(lines)
        lines.append(f"sign = {sign}")
        ok = acceptable(theta)
        ok = ok and is_valid_angle(theta, interval)
        lines += [f"previous_theta: {previous_theta}", f"theta milieu ok mais loin - et moi je suis {ok}"]
        return ok, theta, "\n".join(lines)
    if acceptable(previous_theta):
        lines.append("theta milieu pas ok mais moi ok - bouge pas ")
        return True, previous_theta, "\n".join(lines)
    arrived, theta, _ = _step(previous_theta, preferred_theta, d_theta_max)
    lines.append("theta milieu pas ok et moi pas ok - " + ("proche de theta pref" if arrived else "bouge vers theta pref"))
    return False, theta, "\n".join(lines)
