"""B200-native batched implementation of Reachy2's symbolic 7-DoF arm IK.

Drop-in for the hot path of pollen-robotics/reachy2_symbolic_ik: ``SymbolicIK`` and
``ControlIK`` keep the reference's Python API and gain batched entry points that run
hand-written sm_100a CUDA kernels through the C-ABI library ``libr2ik.so``
(``include/r2ik.h``).  There is no CPU fallback: importing the solver classes without the
built CUDA library raises.
"""
from __future__ import annotations

__version__ = "0.1.0"

__all__ = ["SymbolicIK", "ControlIK", "STATE_STRINGS"]


def __getattr__(name):  # lazy: `fk` / `params` stay importable without the native library
    if name == "SymbolicIK":
        from .symbolic_ik import SymbolicIK
        return SymbolicIK
    if name == "ControlIK":
        from .control_ik import ControlIK
        return ControlIK
    if name == "STATE_STRINGS":
        from .states import STATE_STRINGS
        return STATE_STRINGS
    raise AttributeError(name)
