"""``SymbolicIK`` -- drop-in for ``reachy2_symbolic_ik.symbolic_ik.SymbolicIK`` whose geometry
runs in the sm_100a CUDA library (``libr2ik.so``), plus batched entry points.

Reference interface mirrored here: constructor ``symbolic_ik.py:26-83``; ``is_reachable``
``:121-282``; ``is_reachable_no_limits`` ``:85-119``; ``get_joints`` ``:697-863``;
``get_elbow_position`` ``:684-695``.  The scalar calls are N = 1 launches of the same kernels
as the batched calls; there is no CPU code path.

Semantics that differ from the reference on purpose: ``get_joints`` of the
reference mutates the solver when the elbow projection fires, so a second call on the same
solved pose returns different joints.  Here every ``get_joints`` / closure call is a fresh
solve of the last pose (the value the reference returns on its *first* call).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Optional, Tuple

import numpy as np

from . import _abi, _native
from .states import STATE_STRINGS

_ZERO7 = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


@dataclass
class BatchResult:
    """Outputs of a batched solve.  Arrays are torch CUDA tensors when the input was a CUDA
    tensor, NumPy arrays otherwise."""

    reachable: Any        # (N,) bool
    theta_interval: Any   # (N, 2) float64, NaN when unreachable; [0] > [1] means wrapped
    state: Any            # (N,) uint8, see states.STATE_STRINGS
    joints: Any           # (N, 7) float64, NaN when unreachable
    elbow: Any            # (N, 3) float64
    n_escalated: Any = None   # precision="fp32" only: poses re-solved by the FP64 solver (int, or a CUDA int32 tensor)

    def state_strings(self):
        s = self.state
        if hasattr(s, "cpu"):
            s = s.cpu().numpy()
        return [STATE_STRINGS[int(c)] for c in s]


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def _precision_dtype(torch, precision: str):
    if precision == "fp64":
        return torch.float64
    if precision == "fp32":
        return torch.float32
    raise ValueError(f"precision must be 'fp64' or 'fp32', got {precision!r}")


def normalise_poses(torch, poses, device, dtype=None):
    """Accept (N,4,4) / (N,16) homogeneous matrices or (N,2,3) / (N,6) reference goal poses, as
    NumPy, CPU tensor or CUDA tensor.  Returns (device tensor (N,k) of ``dtype`` (default float64), kind,
    was_cuda)."""
    dtype = torch.float64 if dtype is None else dtype
    was_cuda = hasattr(poses, "is_cuda") and poses.is_cuda
    if not hasattr(poses, "is_cuda"):
        poses = torch.from_numpy(np.ascontiguousarray(poses))
    if poses.dtype != dtype:
        poses = poses.to(dtype)
    shp = tuple(poses.shape)
    if len(shp) == 3 and shp[1:] == (4, 4) or len(shp) == 2 and shp[1] == 16:
        kind, k = _abi.POSE_MAT4, 16
    elif len(shp) == 3 and shp[1:] == (2, 3) or len(shp) == 2 and shp[1] == 6:
        kind, k = _abi.POSE_EULER6, 6
    else:
        raise ValueError(f"poses must be (N,4,4), (N,16), (N,2,3) or (N,6); got {shp}")
    poses = poses.reshape(shp[0], k).contiguous()
    if not was_cuda:
        poses = poses.to(device, non_blocking=True)
    elif poses.data_ptr() % 16:
        poses = poses.clone()   # the C ABI reads poses with 128-bit loads
    return poses, kind, was_cuda


class SymbolicIK:
    def __init__(
        self,
        arm: str = "r_arm",
        ik_parameters: dict[str, Any] = {},
        elbow_limit: int = 127,
        wrist_limit: np.float64 = np.float64(42.5),
        projection_margin: float = 1e-8,
        backward_limit: float = 0.02,
        normal_vector_margin: float = 1e-7,
        singularity_offset: float = 0.03,
        singularity_limit_coeff: float = 1.0,
        device: Optional[int] = None,
    ) -> None:
        if ik_parameters == {}:
            print("Using default parameters")
            ik_parameters = _abi.DEFAULT_IK_PARAMETERS
        if arm not in ["r_arm", "l_arm"]:
            raise ValueError("arm should be either 'r_arm' or 'l_arm'")
        self.arm = arm
        n = arm[0]
        self.shoulder_position = ik_parameters[f"{n}_shoulder_position"]
        self.shoulder_orientation_offset = ik_parameters[f"{n}_shoulder_orientation"]
        self.upper_arm_size = ik_parameters[f"{n}_upper_arm_size"]
        self.forearm_size = ik_parameters[f"{n}_forearm_size"]
        self.tip_position = ik_parameters[f"{n}_tip_position"]
        self.torso_pose = np.array([0.0, 0.0, 0.0])
        self.projection_margin = projection_margin
        self.normal_vector_margin = normal_vector_margin
        self.backward_limit = backward_limit
        self.elbow_limit = elbow_limit
        self.wrist_limit = wrist_limit
        self.singularity_offset = singularity_offset
        self.singularity_limit_coeff = singularity_limit_coeff

        cfg = _abi.make_arm_config(arm, ik_parameters, elbow_limit, wrist_limit, projection_margin, backward_limit,
                                   normal_vector_margin, singularity_offset, singularity_limit_coeff)
        self._torch = _native.require_cuda()
        self._handle = _native.Handle(cfg, device)
        self._device = self._torch.device("cuda", self._handle.device)
        k = self._handle.constants
        self.gripper_size = np.float64(k.gripper_size)
        self.max_arm_length = np.float64(k.max_arm_length)
        self.shoulder_wrist_min_distance = np.float64(k.shoulder_wrist_min_distance)
        self.elbow_singularity_position = np.array(k.elbow_singularity_position[:])
        self.wrist_singularity_position = np.array(k.wrist_singularity_position[:])

        # state of the scalar API: the last pose handed to is_reachable / is_reachable_no_limits
        self.goal_pose: Optional[np.ndarray] = None
        self.elbow_position: Optional[np.ndarray] = None
        self._no_limits = False

    # ------------------------------------------------------------------ batched API
    def is_reachable_batch(self, poses, theta=None, previous_joints=None, want_joints: bool = True,
                           precision: str = "fp64") -> BatchResult:
        """``is_reachable`` + ``theta_to_joints_func(theta)`` for N poses in one launch.

        poses: (N,4,4) homogeneous matrices (converted like the reference's ControlIK front
        end, ``R.from_matrix(M[:3,:3]).as_euler("xyz")``) or (N,2,3)/(N,6) reference goal poses
        ``[[x,y,z],[roll,pitch,yaw]]``.  theta: None -> ``theta_interval[0]``; else (N,).
        precision: "fp64" (the correctness reference, 1e-9 rad) or "fp32" (the fast path: float32 poses in,
        float32 results out, within 1e-4 rad on well-conditioned poses; states are the FP64 path's because poses
        FP32 cannot decide are re-solved in FP64 -- their count is ``n_escalated``).
        """
        torch = self._torch
        dt = _precision_dtype(torch, precision)
        with torch.cuda.device(self._device):
            P, kind, was_cuda = normalise_poses(torch, poses, self._device, dt)
            n = P.shape[0]
            th = None
            if theta is not None:
                th = torch.as_tensor(theta).to(self._device, dt).reshape(n).contiguous()
            pj = None
            if previous_joints is not None:
                pj = torch.as_tensor(previous_joints).to(self._device, dt).reshape(7).contiguous()
            reach = torch.empty(n, dtype=torch.uint8, device=self._device)
            state = torch.empty(n, dtype=torch.uint8, device=self._device)
            interval = torch.empty((n, 2), dtype=dt, device=self._device)
            joints = torch.empty((n, 7), dtype=dt, device=self._device) if want_joints else None
            elbow = torch.empty((n, 3), dtype=dt, device=self._device) if want_joints else None
            n_esc = None
            if precision == "fp32":
                n_esc = torch.empty(1, dtype=torch.int32, device=self._device)
                self.solve_into_f32(P, kind, th, pj, reach, state, interval, joints, elbow, n_esc)
            else:
                self.solve_into(P, kind, th, pj, reach, state, interval, joints, elbow)
            res = BatchResult(reach.bool(), interval, state, joints, elbow, n_esc)
            if was_cuda:
                return res
            out = [None if x is None else x.cpu().numpy() for x in
                   (res.reachable, res.theta_interval, res.state, res.joints, res.elbow)]
            return BatchResult(*out, None if n_esc is None else int(n_esc.item()))

    def solve_into(self, poses_dev, kind, theta_dev, prev_dev, reach, state, interval, joints, elbow, stream=None):
        """Raw launch on device tensors (no allocation, asynchronous): the C-ABI call itself."""
        torch = self._torch
        s = torch.cuda.current_stream(self._device).cuda_stream if stream is None else stream
        rc = self._handle.lib.r2ik_symik_solve_f64(
            self._handle.h, kind, _ptr(poses_dev), _ptr(theta_dev), _ptr(prev_dev), C.c_int64(poses_dev.shape[0]),
            _ptr(reach), _ptr(state), _ptr(interval), _ptr(joints), _ptr(elbow), C.c_void_p(s))
        _native.check(rc, "r2ik_symik_solve_f64")

    def solve_into_f32(self, poses_dev, kind, theta_dev, prev_dev, reach, state, interval, joints, elbow, n_escalated=None,
                       stream=None, scratch=None):
        """Raw launch of the FP32 fast path on float32 device tensors (``r2ik_symik_solve_f32``).  ``n_escalated``: int32
        device tensor (1,) the call sets to the number of poses re-solved in FP64; ``scratch``: int32 device tensor of
        at least N entries (the list of those poses); both are allocated / cached here when omitted."""
        torch = self._torch
        n = poses_dev.shape[0]
        s = torch.cuda.current_stream(self._device).cuda_stream if stream is None else stream
        if scratch is None:   # one cached list per stream: calls on one stream are ordered, calls on two may overlap
            cache = self.__dict__.setdefault("_esc_scratch", {})
            scratch = cache.get(s)
            if scratch is None or scratch.numel() < n:
                scratch = cache[s] = torch.empty(max(n, 1), dtype=torch.int32, device=self._device)
        if n_escalated is None:
            n_escalated = torch.empty(1, dtype=torch.int32, device=self._device)
        rc = self._handle.lib.r2ik_symik_solve_f32(
            self._handle.h, kind, _ptr(poses_dev), _ptr(theta_dev), _ptr(prev_dev), C.c_int64(n),
            _ptr(reach), _ptr(state), _ptr(interval), _ptr(joints), _ptr(elbow), _ptr(scratch), _ptr(n_escalated),
            C.c_void_p(s))
        _native.check(rc, "r2ik_symik_solve_f32")

    # ------------------------------------------------------------------ host-buffer pipeline
    def alloc_host_outputs(self, n: int, precision: str = "fp64") -> BatchResult:
        """Pinned host output buffers for ``is_reachable_batch_host`` (reusable across calls)."""
        torch = self._torch
        dt = _precision_dtype(torch, precision)
        return BatchResult(
            reachable=torch.empty(n, dtype=torch.uint8).pin_memory(),
            theta_interval=torch.empty((n, 2), dtype=dt).pin_memory(),
            state=torch.empty(n, dtype=torch.uint8).pin_memory(),
            joints=torch.empty((n, 7), dtype=dt).pin_memory(),
            elbow=torch.empty((n, 3), dtype=dt).pin_memory())

    def is_reachable_batch_host(self, poses_host, out: Optional[BatchResult] = None, chunk: Optional[int] = None,
                                n_streams: int = 3, precision: str = "fp64") -> BatchResult:
        """Host-to-host batched solve: ``poses_host`` is a CPU tensor (N,16)/(N,4,4)/(N,6)/(N,2,3),
        ideally pinned; results land in ``out`` (pinned CPU tensors; ``reachable`` is uint8 0/1).
        The batch is cut into chunks that flow H2D -> K1 -> D2H on ``n_streams`` CUDA streams so the
        PCIe copies in both directions overlap the kernel.  Synchronous on return.  precision="fp32": float32
        poses and outputs (half the PCIe bytes), the FP32 fast path of K1."""
        torch = self._torch
        dt = _precision_dtype(torch, precision)
        if chunk is None:   # measured optimum on B200 / PCIe 5 (scripts/exp_e2e.py): 128 k poses (FP64), 256 k (FP32)
            chunk = 1 << 17 if precision == "fp64" else 1 << 18
        if not hasattr(poses_host, "is_cuda"):
            poses_host = torch.from_numpy(np.ascontiguousarray(poses_host))
        if poses_host.dtype != dt:
            poses_host = poses_host.to(dt)
        shp = tuple(poses_host.shape)
        k = 16 if (shp[1:] == (4, 4) or shp[1:] == (16,)) else 6
        if k == 6 and shp[1:] not in ((2, 3), (6,)):
            raise ValueError(f"poses must be (N,4,4), (N,16), (N,2,3) or (N,6); got {shp}")
        kind = _abi.POSE_MAT4 if k == 16 else _abi.POSE_EULER6
        P = poses_host.reshape(shp[0], k)
        n = shp[0]
        if out is None:
            out = self.alloc_host_outputs(n, precision)
        fp32 = precision == "fp32"
        with torch.cuda.device(self._device):
            pipe = self._pipeline(chunk, n_streams, k, dt)
            cur = torch.cuda.current_stream(self._device)
            for s in pipe["streams"]:
                s.wait_stream(cur)
            for ci, lo in enumerate(range(0, n, chunk)):
                hi = min(n, lo + chunk)
                m = hi - lo
                slot = ci % n_streams
                s = pipe["streams"][slot]
                b = pipe["bufs"][slot]
                with torch.cuda.stream(s):
                    b["poses"][:m].copy_(P[lo:hi], non_blocking=True)
                    if fp32:
                        self.solve_into_f32(b["poses"][:m], kind, None, None, b["reach"], b["state"], b["interval"], b["joints"],
                                            b["elbow"], b["n_esc"], stream=s.cuda_stream, scratch=b["esc"])
                    else:
                        self.solve_into(b["poses"][:m], kind, None, None, b["reach"], b["state"], b["interval"], b["joints"],
                                        b["elbow"], stream=s.cuda_stream)
                    out.reachable[lo:hi].copy_(b["reach"][:m], non_blocking=True)
                    out.state[lo:hi].copy_(b["state"][:m], non_blocking=True)
                    out.theta_interval[lo:hi].copy_(b["interval"][:m], non_blocking=True)
                    out.joints[lo:hi].copy_(b["joints"][:m], non_blocking=True)
                    out.elbow[lo:hi].copy_(b["elbow"][:m], non_blocking=True)
            for s in pipe["streams"]:
                s.synchronize()
        return out

    def _pipeline(self, chunk: int, n_streams: int, k: int, dt=None):
        dt = self._torch.float64 if dt is None else dt
        key = (chunk, n_streams, k, dt)
        cache = getattr(self, "_pipe_cache", None)
        if cache is None:
            cache = self._pipe_cache = {}
        if key not in cache:
            torch = self._torch
            d = self._device
            cache[key] = {
                "streams": [torch.cuda.Stream(device=d) for _ in range(n_streams)],
                "bufs": [dict(poses=torch.empty((chunk, k), dtype=dt, device=d),
                              reach=torch.empty(chunk, dtype=torch.uint8, device=d),
                              state=torch.empty(chunk, dtype=torch.uint8, device=d),
                              interval=torch.empty((chunk, 2), dtype=dt, device=d),
                              joints=torch.empty((chunk, 7), dtype=dt, device=d),
                              elbow=torch.empty((chunk, 3), dtype=dt, device=d),
                              esc=torch.empty(chunk, dtype=torch.int32, device=d),
                              n_esc=torch.empty(1, dtype=torch.int32, device=d)) for _ in range(n_streams)],
            }
        return cache[key]

    def is_reachable_no_limits_batch(self, poses, theta):
        """``is_reachable_no_limits`` + ``get_joints(theta)``: returns (joints (N,7), elbow (N,3))."""
        torch = self._torch
        with torch.cuda.device(self._device):
            P, kind, was_cuda = normalise_poses(torch, poses, self._device)
            n = P.shape[0]
            th = torch.as_tensor(theta, dtype=torch.float64).to(self._device).reshape(n).contiguous()
            joints = torch.empty((n, 7), dtype=torch.float64, device=self._device)
            elbow = torch.empty((n, 3), dtype=torch.float64, device=self._device)
            s = torch.cuda.current_stream(self._device).cuda_stream
            rc = self._handle.lib.r2ik_symik_no_limits_f64(self._handle.h, kind, _ptr(P), _ptr(th), C.c_int64(n),
                                                           _ptr(joints), _ptr(elbow), C.c_void_p(s))
            _native.check(rc, "r2ik_symik_no_limits_f64")
            if was_cuda:
                return joints, elbow
            return joints.cpu().numpy(), elbow.cpu().numpy()

    def get_elbow_position_batch(self, poses, thetas):
        """``get_elbow_position`` for K thetas per pose after ``is_reachable``: (N,K,3), NaN if unreachable."""
        torch = self._torch
        with torch.cuda.device(self._device):
            P, kind, was_cuda = normalise_poses(torch, poses, self._device)
            n = P.shape[0]
            th = torch.as_tensor(thetas, dtype=torch.float64).to(self._device).reshape(n, -1).contiguous()
            K = th.shape[1]
            out = torch.empty((n, K, 3), dtype=torch.float64, device=self._device)
            s = torch.cuda.current_stream(self._device).cuda_stream
            rc = self._handle.lib.r2ik_elbow_positions_f64(self._handle.h, kind, _ptr(P), _ptr(th), C.c_int32(K),
                                                           C.c_int64(n), _ptr(out), C.c_void_p(s))
            _native.check(rc, "r2ik_elbow_positions_f64")
            return out if was_cuda else out.cpu().numpy()

    def reach_map(self, n: int = 256, orientations_euler=None, n_orientations: int = 512, **kw):
        """Workspace reachability map: int32 CUDA tensor (n,n,n) of per-voxel reachable-orientation
        counts (see ``workspace.reach_map``; all-reduced over ranks when ``dist=torch.distributed``)."""
        from . import workspace

        return workspace.reach_map(self, n=n, orientations_euler=orientations_euler, n_orientations=n_orientations, **kw)

    def task_space_test(self, arm_length: float = 0.5, precision: str = "fp64", **steps):
        """Batched form of the reference's ``task_space_test`` sweep (``src/benchmark/ik_comparison.py:137-181``):
        returns (goal_poses, BatchResult) for the position x Euler-angle grid (see ``workspace.task_space_grid``)."""
        from . import workspace

        return workspace.task_space_test(self, arm_length=arm_length, precision=precision, **steps)

    # ------------------------------------------------------------------ scalar API (reference signatures)
    @staticmethod
    def _pose6(goal_pose) -> np.ndarray:
        gp = np.array([np.asarray(goal_pose[0], dtype=np.float64), np.asarray(goal_pose[1], dtype=np.float64)])
        return gp.reshape(1, 6)

    def is_reachable(self, goal_pose) -> Tuple[bool, np.ndarray, Optional[Any], str]:
        """Check if the goal pose ``[[x,y,z],[roll,pitch,yaw]]`` is reachable taking the wrist and
        elbow limits into account; returns (is_reachable, theta_interval, theta_to_joints_func, state)."""
        P = self._pose6(goal_pose)
        res = self.is_reachable_batch(P, want_joints=False)
        self.goal_pose = P.reshape(2, 3).copy()
        self._no_limits = False
        state = STATE_STRINGS[int(res.state[0])]
        if bool(res.reachable[0]):
            return True, np.array(res.theta_interval[0]), self.get_joints, state
        return False, np.array([]), None, state

    def is_reachable_no_limits(self, goal_pose) -> Tuple[bool, np.ndarray, Optional[Any]]:
        """Reachability without the wrist / elbow limits (unreachable goals are projected);
        always (True, [-pi, pi], get_joints)."""
        P = self._pose6(goal_pose)
        self.goal_pose = P.reshape(2, 3).copy()
        self._no_limits = True
        return True, np.array([-np.pi, np.pi]), self.get_joints

    def get_joints(self, theta: float, previous_joints: list = _ZERO7) -> Tuple[np.ndarray, np.ndarray]:
        """Joints for the elbow angle theta on the last solved pose (fresh solve per call)."""
        if self.goal_pose is None:
            raise AttributeError("get_joints called before is_reachable / is_reachable_no_limits")
        P = self.goal_pose.reshape(1, 6)
        if self._no_limits:
            joints, elbow = self.is_reachable_no_limits_batch(P, np.array([float(theta)]))
            j, e = joints[0], elbow[0]
        else:
            res = self.is_reachable_batch(P, theta=np.array([float(theta)]), previous_joints=previous_joints)
            j, e = res.joints[0], res.elbow[0]
        self.elbow_position = np.array(e)
        return np.array(j), self.elbow_position

    def get_elbow_position(self, theta: float) -> np.ndarray:
        """Elbow position [x, y, z, 1] on the elbow circle of the last solved pose."""
        if self.goal_pose is None:
            raise AttributeError("get_elbow_position called before is_reachable")
        if self._no_limits:
            raise NotImplementedError("get_elbow_position after is_reachable_no_limits is not exposed")
        e = self.get_elbow_position_batch(self.goal_pose.reshape(1, 6), np.array([[float(theta)]]))[0, 0]
        return np.array([e[0], e[1], e[2], 1.0])
