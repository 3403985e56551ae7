"""``ControlIK`` -- drop-in for ``reachy2_symbolic_ik.control_ik.ControlIK`` on the CUDA library,
plus batched entry points (N poses in discrete mode, T trajectories x W waypoints in continuous
mode).  Reference: constructor ``control_ik.py:28-160``; ``symbolic_inverse_kinematics``
``:162-274``; continuous ``:276-407``; discrete ``:409-462``; ``safety_checks`` ``:464-497``.

All arithmetic (matrix -> euler front end with the identity snap, reachability, elbow-angle
policies, joints, safety chain) runs in the kernels; this module only keeps the controller
state dictionaries the reference keeps, and turns state codes back into its strings.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Any, Dict, Tuple

import numpy as np

from . import _abi, _native
from .fk import bundled_urdf_path
from .states import STATE_EMERGENCY, STATE_INVALID_ROTATION, STATE_STRINGS, emergency_text
from .symbolic_ik import SymbolicIK, _ptr
from .urdf import get_ik_parameters_from_urdf

DEBUG = False

_DEFAULT_CURRENT_JOINTS = [
    [0.0, 0.2617993877991494, -0.17453292519943295, 0.0, 0.0, 0.0, 0.0],
    [0.0, -0.2617993877991494, 0.17453292519943295, 0.0, 0.0, 0.0, 0.0],
]
_DEFAULT_CURRENT_POSE = [
    np.array([[1, 0, 0, 0], [0, 1, 0, -0.2], [0, 0, 1, -0.66], [0, 0, 0, 1]]),
    np.array([[1, 0, 0, 0], [0, 1, 0, 0.2], [0, 0, 1, -0.66], [0, 0, 0, 1]]),
]


class ControlIK:
    def __init__(
        self,
        current_joints: list = _DEFAULT_CURRENT_JOINTS,
        current_pose: list = _DEFAULT_CURRENT_POSE,
        logger: Any = None,
        urdf: str = "",
        urdf_path: str = "",
        reachy_model: str = "full_kit",
        is_dvt: bool = False,
        device: int | None = None,
    ) -> None:
        self.symbolic_ik_solver: Dict[str, SymbolicIK] = {}
        self.last_call_t: Dict[str, float] = {}
        self.call_timeout = 0.2
        self.nb_search_points = 20
        self.emergency_state = ""
        self.emergency_stop = False
        self.init = True
        self.logger = logger
        if is_dvt:
            self.singularity_offset = 0.03
            if self.logger is not None:
                self.logger.info("DVT mode activated", throttle_duration_sec=0.1)
            else:
                print("DVT mode activated")
        else:
            self.singularity_offset = -1.01
        self.singularity_limit_coeff = 1.0
        self.preferred_theta: Dict[str, float] = {}
        self.previous_theta: Dict[str, float] = {}
        self.previous_sol: Dict[str, np.ndarray] = {}
        self.previous_pose: Dict[str, np.ndarray] = {}
        self.orbita3D_max_angle = np.deg2rad(42.5)

        if urdf_path == "" and urdf == "":
            raise ValueError("No URDF provided")
        if urdf_path != "" and urdf == "":
            full = os.path.join(os.path.dirname(__file__), urdf_path)
            if not os.path.isfile(full) and os.path.basename(urdf_path) == "reachy2.urdf":
                # the reference ships the robot description next to its package
                # ("../config_files/reachy2.urdf"); here the arm chains are bundled instead
                full = bundled_urdf_path()
            if os.path.isfile(full) and os.path.getsize(full) > 0:
                with open(full, "r") as f:
                    urdf = f.read()
            if urdf == "":
                raise ValueError("Empty URDF file")
        if reachy_model == "full_kit" or reachy_model == "headless":
            arms = ["r", "l"]
        elif reachy_model == "starter_kit_right":
            arms = ["r"]
        elif reachy_model == "starter_kit_left":
            arms = ["l"]
        elif reachy_model == "mini":
            arms = []
        else:
            raise ValueError(f"Unknown Reachy model {reachy_model}")
        try:
            ik_parameters = get_ik_parameters_from_urdf(urdf, arms)
        except Exception as e:
            raise ValueError(f"Error while parsing URDF: {e}")

        self._torch = _native.require_cuda() if arms else None
        self._device = None
        for prefix in arms:
            arm = f"{prefix}_arm"
            if ik_parameters != {}:
                self.symbolic_ik_solver[arm] = SymbolicIK(
                    arm=arm, ik_parameters=ik_parameters, singularity_offset=self.singularity_offset,
                    singularity_limit_coeff=self.singularity_limit_coeff, device=device)
            else:
                self.symbolic_ik_solver[arm] = SymbolicIK(
                    arm=arm, wrist_limit=np.rad2deg(self.orbita3D_max_angle), singularity_offset=self.singularity_offset,
                    singularity_limit_coeff=self.singularity_limit_coeff, device=device)
            self._device = self.symbolic_ik_solver[arm]._device
            preferred_theta = -4 * np.pi / 6
            k = 0 if prefix == "r" else 1
            self.preferred_theta[arm] = preferred_theta if prefix == "r" else -np.pi - preferred_theta
            self.previous_sol[arm] = np.array(current_joints[k], dtype=np.float64)
            self.previous_pose[arm] = np.array(current_pose[k], dtype=np.float64)
            # The reference seeds previous_theta here from a call that receives the two-arm joint
            # list by mistake (control_ik.py:152-158); the value never reaches an
            # output (discrete mode ignores it, continuous mode re-initialises on its first call).
            self.previous_theta[arm] = self.preferred_theta[arm]
            self.last_call_t[arm] = 0.0

    # ------------------------------------------------------------------ parameters
    def _ctl_params(self, name: str, constrained_mode: str, preferred_theta: float, d_theta_max: float) -> _abi.CtlParams:
        side = 1 if name.startswith("r") else -1
        if constrained_mode == "unconstrained":
            low = False
        elif constrained_mode == "low_elbow":
            low = True
        else:
            # the reference leaves interval_limit unbound here (UnboundLocalError)
            raise ValueError(f"Unknown constrained_mode {constrained_mode}")
        p = _abi.CtlParams()
        p.interval_limit[:] = _native.interval_limit(side, low)
        p.preferred_theta = preferred_theta if side > 0 else -np.pi - preferred_theta
        p.preferred_theta_ctor = self.preferred_theta[name]
        p.d_theta_max = d_theta_max
        p.orbita3d_max_angle = float(self.orbita3D_max_angle)
        p.nb_search_points = int(self.nb_search_points)
        p.nb_search_points_continuous = 10
        return p

    def _dev(self, x, shape=None):
        torch = self._torch
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64) if not hasattr(x, "is_cuda") else x,
                            dtype=torch.float64).to(self._device)
        if shape is not None:
            t = t.reshape(shape)
        return t.contiguous()

    # ------------------------------------------------------------------ batched API
    def symbolic_inverse_kinematics_batch(self, name: str, M, control_type: str = "discrete", current_joints=None,
                                          constrained_mode: str = "unconstrained", current_pose=None,
                                          d_theta_max: float = 0.01, preferred_theta: float = -4 * np.pi / 6,
                                          previous_joints=None, states=None, out=None):
        """Batched ``symbolic_inverse_kinematics``.

        discrete:   M (N,4,4) -> joints (N,7), reachable (N,), state (N,) uint8, emergency bits (N,).
                    Every pose is solved against the same previous solution (``previous_joints``,
                    default ``self.previous_sol[name]``), like N independent reference calls.
        continuous: M (T,W,4,4) -> joints (T,W,7), reachable (T,W), state (T,W), states (T,) structured
                    array (``_abi.TRAJ_STATE_DTYPE``) that can be passed back to resume the trajectories.
                    ``states`` may also be a CUDA uint8 tensor (T,80): it is then updated in place and
                    returned as is (no host round trip).
        ``out``: the tuple a previous call returned for CUDA input of the same shape; its tensors are reused.
        """
        torch = self._torch
        solver = self.symbolic_ik_solver[name]
        par = self._ctl_params(name, constrained_mode, preferred_theta, d_theta_max)
        was_cuda = hasattr(M, "is_cuda") and M.is_cuda
        with torch.cuda.device(self._device):
            stream = C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
            Md = self._dev(M)
            if control_type == "discrete":
                n = Md.shape[0]
                Md = Md.reshape(n, 16)
                prev = self._dev(self.previous_sol[name] if previous_joints is None else previous_joints, (7,))
                cur = prev if current_joints is None else self._dev(current_joints, (7,))
                if out is not None and was_cuda:
                    joints, reach, state, emg = out[0], out[1].view(torch.uint8), out[2], out[3]
                else:
                    joints = torch.empty((n, 7), dtype=torch.float64, device=self._device)
                    reach = torch.empty(n, dtype=torch.uint8, device=self._device)
                    state = torch.empty(n, dtype=torch.uint8, device=self._device)
                    emg = torch.empty(n, dtype=torch.uint8, device=self._device)
                rc = solver._handle.lib.r2ik_ctl_discrete_f64(solver._handle.h, C.byref(par), _ptr(Md), C.c_int64(n),
                                                              _ptr(prev), _ptr(cur), _ptr(joints), _ptr(reach), _ptr(state),
                                                              _ptr(emg), stream)
                _native.check(rc, "r2ik_ctl_discrete_f64")
                res = (joints, reach.view(torch.bool), state, emg)
                return res if was_cuda else tuple(x.cpu().numpy() for x in res)
            if control_type == "continuous":
                T, W = Md.shape[0], Md.shape[1]
                Md = Md.reshape(T, W, 16)
                cj = torch.empty((T, 7), dtype=torch.float64, device=self._device)
                cj[:] = self._dev(self.previous_sol[name] if current_joints is None else current_joints)
                cp = torch.empty((T, 16), dtype=torch.float64, device=self._device)
                cp[:] = self._dev(self.previous_pose[name] if current_pose is None else current_pose).reshape(-1, 16)
                states_on_device = hasattr(states, "is_cuda") and states.is_cuda
                if states_on_device:
                    st = states
                else:
                    if states is None:
                        states = np.zeros(T, dtype=_abi.TRAJ_STATE_DTYPE)
                        states["init"] = 1
                    st = torch.from_numpy(np.ascontiguousarray(states).view(np.uint8).reshape(T, -1)).to(self._device)
                if out is not None and was_cuda:
                    joints, reach, state = out[0], out[1].view(torch.uint8), out[2]
                else:
                    joints = torch.empty((T, W, 7), dtype=torch.float64, device=self._device)
                    reach = torch.empty((T, W), dtype=torch.uint8, device=self._device)
                    state = torch.empty((T, W), dtype=torch.uint8, device=self._device)
                rc = solver._handle.lib.r2ik_ctl_continuous_f64(solver._handle.h, C.byref(par), _ptr(Md), C.c_int64(T),
                                                                C.c_int32(W), _ptr(cj), _ptr(cp), _ptr(st), _ptr(joints),
                                                                _ptr(reach), _ptr(state), stream)
                _native.check(rc, "r2ik_ctl_continuous_f64")
                st_out = st if states_on_device else st.cpu().numpy().reshape(-1).view(_abi.TRAJ_STATE_DTYPE).copy()
                res = (joints, reach.view(torch.bool), state)
                res = res if was_cuda else tuple(x.cpu().numpy() for x in res)
                return (*res, st_out)
            raise ValueError(f"Unknown type {control_type}")

    # ------------------------------------------------------------------ scalar API (reference signature)
    def symbolic_inverse_kinematics(  # noqa: C901
        self,
        name: str,
        M: np.ndarray,
        control_type: str,
        current_joints: list = [],
        constrained_mode: str = "unconstrained",
        current_pose: np.ndarray = np.array([]),
        d_theta_max: float = 0.01,
        preferred_theta: float = -4 * np.pi / 6,
    ) -> Tuple[np.ndarray, bool, str]:
        if control_type == "unfreeze":
            self.emergency_stop = False
            self.emergency_state = ""
            self.init = True
            if self.logger is not None:
                self.logger.info(f"{name} Unfreeze", throttle_duration_sec=1.0)
            else:
                print(f"{name} Unfreeze")
        if self.emergency_stop:
            if self.logger is not None:
                self.logger.info(f"{name} Emergency state: {self.emergency_state}", throttle_duration_sec=1.0)
            else:
                print(f"{name} Emergency state: {self.emergency_state}")
            return self.previous_sol[name], False, self.emergency_state
        M = np.asarray(M, dtype=np.float64)
        if len(current_pose) == 0:
            current_pose = self.previous_pose[name]
        if current_joints == []:
            current_joints = self.previous_sol[name].tolist()

        if control_type == "continuous" or control_type == "unfreeze":
            ik_joints, is_reachable, state = self._continuous_one(name, M, current_joints, current_pose, constrained_mode,
                                                                  preferred_theta, d_theta_max)
        elif control_type == "discrete":
            j, r, s, e = self.symbolic_inverse_kinematics_batch(
                name, M[None], "discrete", current_joints=current_joints, constrained_mode=constrained_mode,
                preferred_theta=preferred_theta)
            if int(s[0]) == STATE_INVALID_ROTATION:
                raise ValueError("Non-positive determinant (left-handed or null coordinate frame) in rotation matrix")
            ik_joints, is_reachable, state = j[0], bool(r[0]), STATE_STRINGS[int(s[0])]
            if int(e[0]):
                self.emergency_state += emergency_text(int(e[0]))
                self.emergency_stop = True
        else:
            raise ValueError(f"Unknown type {control_type}")
        self.previous_pose[name] = M
        return ik_joints, is_reachable, state

    def _continuous_one(self, name, M, current_joints, current_pose, constrained_mode, preferred_theta, d_theta_max):
        t = time.time()
        if abs(t - self.last_call_t[name]) > self.call_timeout:  # control_ik.py:296-304
            self.previous_sol[name] = np.array([])
            self.init = True
        self.last_call_t[name] = t
        st = np.zeros(1, dtype=_abi.TRAJ_STATE_DTYPE)
        st["init"] = int(self.init)
        st["previous_theta"] = self.previous_theta[name]
        if len(self.previous_sol[name]) != 0:
            st["has_previous_sol"] = 1
            st["previous_sol"][0] = self.previous_sol[name]
        # previous_sol must be defined for the parameter builder when it was just reset
        j, r, s, st = self.symbolic_inverse_kinematics_batch(
            name, M[None, None], "continuous", current_joints=np.asarray(current_joints, dtype=np.float64),
            constrained_mode=constrained_mode, current_pose=np.asarray(current_pose, dtype=np.float64),
            d_theta_max=d_theta_max, preferred_theta=preferred_theta, states=st)
        code = int(s[0, 0])
        if code == STATE_INVALID_ROTATION:
            raise ValueError("Non-positive determinant (left-handed or null coordinate frame) in rotation matrix")
        self.previous_theta[name] = float(st["previous_theta"][0])
        self.previous_sol[name] = np.array(st["previous_sol"][0])
        self.init = bool(st["init"][0])
        bits = int(st["emergency_bits"][0])
        if st["emergency_stop"][0]:
            self.emergency_stop = True
            self.emergency_state += emergency_text(bits)
            if bits & _abi.EMG_DISCONTINUITY:
                self.emergency_state += (f"\n EMERGENCY STOP: joints are not continuous \n previous_joints: "
                                         f"{self.previous_sol[name]} \n joints: (see device log)")
        state = self.emergency_state if code == STATE_EMERGENCY else STATE_STRINGS[code]
        return np.array(j[0, 0]), bool(r[0, 0]), state
