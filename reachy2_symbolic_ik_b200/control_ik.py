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
                                          previous_joints=None, states=None, out=None, phased: bool = True,
                                          exhaustive: bool = False):
        """Batched ``symbolic_inverse_kinematics``.

        discrete:   M (N,4,4) -> joints (N,7), reachable (N,), state (N,) uint8, emergency bits (N,).
                    Every pose is solved against the same previous solution (``previous_joints``,
                    default ``self.previous_sol[name]``), like N independent reference calls.
        continuous: M (T,W,4,4) -> joints (T,W,7), reachable (T,W), state (T,W), states (T,) structured
                    array (``_abi.TRAJ_STATE_DTYPE``) that can be passed back to resume the trajectories.
                    ``states`` may also be a CUDA uint8 tensor (T,80): it is then updated in place and
                    returned as is (no host round trip).
        ``out``: the tuple a previous call returned for CUDA input of the same shape; its tensors are reused.
        ``exhaustive`` (discrete): False = the elbow search finds the arg-min over the nb_search_points samples from
        the crossings of the two elbow tests (cost independent of K); True = every sample is visited by the
        warp-cooperative scan kernel.  Same outputs.
        ``phased`` (continuous): True = per-waypoint kernels + per-trajectory scans (needs T*W doubles of device
        scratch, allocated here); False = the single one-thread-per-trajectory kernel.  Same flags / states; joints equal to rounding.
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
                entry = solver._handle.lib.r2ik_ctl_discrete_scan_f64 if exhaustive else solver._handle.lib.r2ik_ctl_discrete_f64
                rc = entry(solver._handle.h, C.byref(par), _ptr(Md), C.c_int64(n),
                           _ptr(prev), _ptr(cur), _ptr(joints), _ptr(reach), _ptr(state), _ptr(emg), stream)
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
                if phased:
                    ws = self._scratch(T * W)
                    rc = solver._handle.lib.r2ik_ctl_continuous_phased_f64(
                        solver._handle.h, C.byref(par), _ptr(Md), C.c_int64(T), C.c_int32(W), _ptr(cj), _ptr(cp), _ptr(st),
                        _ptr(joints), _ptr(reach), _ptr(state), _ptr(ws), stream)
                    _native.check(rc, "r2ik_ctl_continuous_phased_f64")
                else:
                    rc = solver._handle.lib.r2ik_ctl_continuous_f64(solver._handle.h, C.byref(par), _ptr(Md), C.c_int64(T),
                                                                    C.c_int32(W), _ptr(cj), _ptr(cp), _ptr(st), _ptr(joints),
                                                                    _ptr(reach), _ptr(state), stream)
                    _native.check(rc, "r2ik_ctl_continuous_f64")
                st_out = st if states_on_device else st.cpu().numpy().reshape(-1).view(_abi.TRAJ_STATE_DTYPE).copy()
                res = (joints, reach.view(torch.bool), state)
                res = res if was_cuda else tuple(x.cpu().numpy() for x in res)
                return (*res, st_out)
            raise ValueError(f"Unknown type {control_type}")

    def _scratch(self, n: int):
        """Device scratch of n doubles for the phased continuous kernels (grown on demand, reused across calls)."""
        torch = self._torch
        buf = getattr(self, "_scratch_buf", None)
        if buf is None or buf.numel() < n or buf.device != self._device:
            buf = self._scratch_buf = torch.empty(n, dtype=torch.float64, device=self._device)
        return buf

    # ------------------------------------------------------------------ host-buffer pipelines
    def alloc_host_outputs(self, control_type: str, shape) -> tuple:
        """Pinned host output buffers for ``symbolic_inverse_kinematics_batch_host`` (reusable across calls).
        discrete: shape = N -> (joints (N,7), reachable (N,) u8, state (N,) u8, emergency (N,) u8);
        continuous: shape = (T, W) -> (joints (T,W,7), reachable (T,W) u8, state (T,W) u8, states (T,80) u8)."""
        torch = self._torch
        pin = lambda *sz, dt=torch.uint8: torch.empty(sz, dtype=dt).pin_memory()
        if control_type == "discrete":
            n = int(shape)
            return (pin(n, 7, dt=torch.float64), pin(n), pin(n), pin(n))
        if control_type == "continuous":
            T, W = shape
            return (pin(T, W, 7, dt=torch.float64), pin(T, W), pin(T, W), pin(T, _abi.TRAJ_STATE_DTYPE.itemsize))
        raise ValueError(f"Unknown type {control_type}")

    def symbolic_inverse_kinematics_batch_host(self, name: str, M_host, control_type: str = "discrete", out=None,
                                               chunk: int | None = None, n_streams: int = 3, current_joints=None,
                                               constrained_mode: str = "unconstrained", current_pose=None,
                                               d_theta_max: float = 0.01, preferred_theta: float = -4 * np.pi / 6,
                                               previous_joints=None, states=None):
        """Host-to-host ``symbolic_inverse_kinematics_batch``: ``M_host`` is a CPU tensor (ideally pinned),
        (N,4,4)/(N,16) for discrete mode or (T,W,4,4)/(T,W,16) for continuous mode; the results land in
        ``out`` (pinned CPU tensors from ``alloc_host_outputs``).  The batch is cut into chunks -- ``chunk``
        poses in discrete mode, ``chunk`` waypoints of every trajectory in continuous mode -- that flow
        H2D -> kernel -> D2H on CUDA streams, so both PCIe directions overlap the kernels.  Synchronous on
        return."""
        torch = self._torch
        solver = self.symbolic_ik_solver[name]
        par = self._ctl_params(name, constrained_mode, preferred_theta, d_theta_max)
        lib, h = solver._handle.lib, solver._handle.h
        if not hasattr(M_host, "is_cuda"):
            M_host = torch.from_numpy(np.ascontiguousarray(M_host, dtype=np.float64))
        dev = self._device
        f64, u8 = torch.float64, torch.uint8
        with torch.cuda.device(dev):
            cur_stream = torch.cuda.current_stream(dev)
            if control_type == "discrete":
                n = M_host.shape[0]
                P = M_host.reshape(n, 16)
                chunk = chunk or (1 << 17)
                if out is None:
                    out = self.alloc_host_outputs("discrete", n)
                prev = self._dev(self.previous_sol[name] if previous_joints is None else previous_joints, (7,))
                cur = prev if current_joints is None else self._dev(current_joints, (7,))
                pipe = self._pipeline(("discrete", chunk, n_streams), lambda: dict(
                    M=torch.empty((chunk, 16), dtype=f64, device=dev), joints=torch.empty((chunk, 7), dtype=f64, device=dev),
                    reach=torch.empty(chunk, dtype=u8, device=dev), state=torch.empty(chunk, dtype=u8, device=dev),
                    emg=torch.empty(chunk, dtype=u8, device=dev)), n_streams)
                for s in pipe["streams"]:
                    s.wait_stream(cur_stream)
                for ci, lo in enumerate(range(0, n, chunk)):
                    hi = min(n, lo + chunk)
                    m = hi - lo
                    s, b = pipe["streams"][ci % n_streams], pipe["bufs"][ci % n_streams]
                    with torch.cuda.stream(s):
                        b["M"][:m].copy_(P[lo:hi], non_blocking=True)
                        rc = lib.r2ik_ctl_discrete_f64(h, C.byref(par), _ptr(b["M"]), C.c_int64(m), _ptr(prev), _ptr(cur),
                                                       _ptr(b["joints"]), _ptr(b["reach"]), _ptr(b["state"]), _ptr(b["emg"]),
                                                       C.c_void_p(s.cuda_stream))
                        _native.check(rc, "r2ik_ctl_discrete_f64")
                        out[0][lo:hi].copy_(b["joints"][:m], non_blocking=True)
                        out[1][lo:hi].copy_(b["reach"][:m], non_blocking=True)
                        out[2][lo:hi].copy_(b["state"][:m], non_blocking=True)
                        out[3][lo:hi].copy_(b["emg"][:m], non_blocking=True)
            elif control_type == "continuous":
                # A trajectory is a recursion over its waypoints, so K3's run time is set by W, not by T: the
                # pipeline therefore cuts the WAYPOINT axis.  Chunk c = waypoints [w0, w1) of every trajectory
                # (a strided block of the (T, W, 16) host array, moved by 2-D copies); the controller states
                # stay on the device between chunks (R2ikTrajState is the resume point).  Copies in, kernels
                # and copies out run on three streams chained by events.
                T, W = M_host.shape[0], M_host.shape[1]
                P = M_host.reshape(T, W, 16)
                wc = chunk or max(1, min(W, (1 << 18) // max(T, 1)))          # waypoints per chunk
                if out is None:
                    out = self.alloc_host_outputs("continuous", (T, W))
                cj = torch.empty((T, 7), dtype=f64, device=dev)
                cj[:] = self._dev(self.previous_sol[name] if current_joints is None else current_joints)
                cp = torch.empty((T, 16), dtype=f64, device=dev)
                cp[:] = self._dev(self.previous_pose[name] if current_pose is None else current_pose).reshape(-1, 16)
                if states is None:
                    states = np.zeros(T, dtype=_abi.TRAJ_STATE_DTYPE)
                    states["init"] = 1
                st = torch.from_numpy(np.ascontiguousarray(states).view(np.uint8).reshape(T, -1)).to(dev)
                n_slots = max(2, n_streams)
                pipe = self._pipeline(("continuous", T, wc, n_slots), lambda: dict(
                    M=torch.empty((T, wc, 16), dtype=f64, device=dev), joints=torch.empty((T, wc, 7), dtype=f64, device=dev),
                    reach=torch.empty((T, wc), dtype=u8, device=dev), state=torch.empty((T, wc), dtype=u8, device=dev),
                    ws=torch.empty((T, wc), dtype=f64, device=dev)),
                    max(3, n_slots))
                s_in, s_k, s_out = pipe["streams"][:3]
                for s in (s_in, s_k, s_out):
                    s.wait_stream(cur_stream)
                ev = [dict(h2d=torch.cuda.Event(), k=None, d2h=None) for _ in range(n_slots)]

                def copy2d(dst_ptr, dpitch, src_ptr, spitch, width, stream):
                    rc = lib.r2ik_copy2d_async(C.c_void_p(dst_ptr), dpitch, C.c_void_p(src_ptr), spitch, width, T,
                                               C.c_void_p(stream.cuda_stream))
                    _native.check(rc, "r2ik_copy2d_async")

                for ci, w0 in enumerate(range(0, W, wc)):
                    w1 = min(W, w0 + wc)
                    m = w1 - w0
                    e, b = ev[ci % n_slots], pipe["bufs"][ci % n_slots]
                    if e["k"] is not None:
                        s_in.wait_event(e["k"])                   # the kernel that read this slot's poses is done
                    copy2d(b["M"].data_ptr(), m * 128, P.data_ptr() + w0 * 128, W * 128, m * 128, s_in)
                    e["h2d"].record(s_in)
                    s_k.wait_event(e["h2d"])
                    if e["d2h"] is not None:
                        s_k.wait_event(e["d2h"])                  # this slot's previous results have left
                    rc = lib.r2ik_ctl_continuous_phased_f64(h, C.byref(par), _ptr(b["M"]), C.c_int64(T), C.c_int32(m), _ptr(cj),
                                                            _ptr(cp), _ptr(st), _ptr(b["joints"]), _ptr(b["reach"]),
                                                            _ptr(b["state"]), _ptr(b["ws"]), C.c_void_p(s_k.cuda_stream))
                    _native.check(rc, "r2ik_ctl_continuous_phased_f64")
                    e["k"] = torch.cuda.Event()
                    e["k"].record(s_k)
                    s_out.wait_event(e["k"])
                    copy2d(out[0].data_ptr() + w0 * 56, W * 56, b["joints"].data_ptr(), m * 56, m * 56, s_out)
                    copy2d(out[1].data_ptr() + w0, W, b["reach"].data_ptr(), m, m, s_out)
                    copy2d(out[2].data_ptr() + w0, W, b["state"].data_ptr(), m, m, s_out)
                    e["d2h"] = torch.cuda.Event()
                    e["d2h"].record(s_out)
                s_out.wait_stream(s_k)
                with torch.cuda.stream(s_out):
                    out[3].copy_(st, non_blocking=True)
            else:
                raise ValueError(f"Unknown type {control_type}")
            for s in pipe["streams"]:
                s.synchronize()
        return out

    def _pipeline(self, key, make_bufs, n_streams: int):
        cache = getattr(self, "_pipe_cache", None)
        if cache is None:
            cache = self._pipe_cache = {}
        if key not in cache:
            torch = self._torch
            cache[key] = {"streams": [torch.cuda.Stream(device=self._device) for _ in range(n_streams)],
                          "bufs": [make_bufs() for _ in range(n_streams)]}
        return cache[key]

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
