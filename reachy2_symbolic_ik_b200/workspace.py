"""Workspace reachability map (BASELINE.json configs[4]) and the multi-GPU partitioning helpers.

The reference has no such component; its closest code is the task-space grid sweep of
``src/benchmark/ik_comparison.py:137-181`` (``is_reachable`` over a position grid).  Here a
voxel grid x orientation set is swept by the K4 kernel (``r2ik_reach_map_u32``): counts[v] =
number of orientations for which ``SymbolicIK.is_reachable(voxel centre, orientation)`` is True.

Multi-GPU: poses / trajectories are independent, so batches are cut into contiguous slices
(``shard_range``) with no exchange.  The reach map shards the ORIENTATION set; every rank fills a
full count volume, and one all-reduce (NCCL over NVLink / NVSwitch) sums them in place -- the only
collective of the whole path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np

from . import _native
from .fk import fibonacci_orientations


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced slice [begin, end) of ``n_items`` owned by ``rank`` (first ``n % world``
    ranks get one extra item)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_items), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reach_grid(shoulder_position, max_arm_length: float, n: int):
    """Cell-centred n^3 grid over the bounding cube of the reach sphere (shoulder +- max_arm_length)."""
    step = 2.0 * float(max_arm_length) / n
    origin = np.asarray(shoulder_position, dtype=np.float64) - float(max_arm_length) + step / 2
    return origin, np.array([step, step, step]), np.array([n, n, n], dtype=np.int32)


def sharded_sum(launch: Callable[[int, int], "object"], n_orientations: int, dist=None, group=None, mark=None):
    """Run ``launch(ori_begin, ori_end)`` on this rank's orientation slice and sum the returned
    count tensor over the ranks in place.  ``dist`` is ``torch.distributed`` (initialised) or None.
    ``mark``: optional callable invoked between the kernel and the collective (benchmarks record an event there)."""
    rank, world = 0, 1
    if dist is not None and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_range(n_orientations, rank, world)
    counts = launch(b, e)
    if mark is not None:
        mark()
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def reach_map(solver, n: int = 256, orientations_euler=None, n_orientations: int = 512, origin=None, step=None,
              dims=None, dist=None, group=None, out=None, all_fp64: bool = False, mark=None):
    """Reachability count volume of ``solver`` (a ``SymbolicIK``): int32 CUDA tensor (d0, d1, d2).

    orientations_euler: (n_ori, 3) xyz Euler angles (default: ``fibonacci_orientations(n_orientations)``).
    With an initialised ``torch.distributed`` passed as ``dist`` the orientation set is sharded over the
    ranks and the volume is all-reduced; every rank returns the full map.
    all_fp64: decide every (voxel, orientation) pair with the FP64 flag solve instead of the mixed-precision test with
    FP64 escalation (identical counts, slower: the cross-check)."""
    torch = solver._torch
    if orientations_euler is None:
        orientations_euler = fibonacci_orientations(n_orientations)
    if origin is None or step is None or dims is None:
        origin, step, dims = reach_grid(solver.shoulder_position, solver.max_arm_length, n)
    origin = np.ascontiguousarray(origin, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    dims = np.ascontiguousarray(dims, dtype=np.int32)
    dev = solver._device
    with torch.cuda.device(dev):
        if hasattr(orientations_euler, "is_cuda"):
            ori = orientations_euler.to(dev, torch.float64).contiguous()
        else:
            ori = torch.from_numpy(np.ascontiguousarray(orientations_euler, dtype=np.float64)).to(dev)
        n_ori = ori.shape[0]
        if out is None:
            out = torch.empty(tuple(int(d) for d in dims), dtype=torch.int32, device=dev)

        def launch(b: int, e: int):
            s = torch.cuda.current_stream(dev).cuda_stream
            entry = solver._handle.lib.r2ik_reach_map_f64_u32 if all_fp64 else solver._handle.lib.r2ik_reach_map_u32
            rc = entry(
                solver._handle.h, origin.ctypes.data_as(C.POINTER(C.c_double)), step.ctypes.data_as(C.POINTER(C.c_double)),
                dims.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(ori.data_ptr()), C.c_int32(b), C.c_int32(e),
                C.c_void_p(out.data_ptr()), C.c_void_p(s))
            _native.check(rc, "r2ik_reach_map_u32")
            return out

        return sharded_sum(launch, n_ori, dist, group, mark)


def task_space_grid(shoulder_position, arm_length: float = 0.5, x_step: float = 0.15, y_step: float = 0.15,
                    z_step: float = 0.15, roll_step: int = 45, pitch_step: int = 45, yaw_step: int = 45) -> np.ndarray:
    """The goal poses of the reference's ``task_space_test`` (``src/benchmark/ik_comparison.py:137-170``): a position
    grid over the cube shoulder +- arm_length, kept where it lies inside the sphere of that radius and in front of the
    robot (x >= 0), crossed with an Euler-angle grid in degrees.  Returns (N, 2, 3) ``[[x, y, z], [roll, pitch, yaw]]``
    in the reference's loop order."""
    s = np.asarray(shoulder_position, dtype=np.float64)
    xs = np.arange(s[0] - arm_length, s[0] + arm_length + x_step, x_step)
    ys = np.arange(s[1] - arm_length, s[1] + arm_length + y_step, y_step)
    zs = np.arange(s[2] - arm_length, s[2] + arm_length + z_step, z_step)
    P = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1).reshape(-1, 3)
    keep = ~((np.linalg.norm(P - s, axis=1) > arm_length) | (P[:, 0] < 0))
    P = P[keep]
    ang = [np.radians(np.arange(0, 360, st)) for st in (roll_step, pitch_step, yaw_step)]
    E = np.stack(np.meshgrid(*ang, indexing="ij"), axis=-1).reshape(-1, 3)
    out = np.empty((len(P), len(E), 2, 3))
    out[:, :, 0, :] = P[:, None, :]
    out[:, :, 1, :] = E[None, :, :]
    return out.reshape(-1, 2, 3)


def task_space_test(solver, arm_length: float = 0.5, precision: str = "fp64", **steps):
    """Batched ``task_space_test`` (``ik_comparison.py:137-181``): ``is_reachable`` on every pose of ``task_space_grid``
    in one launch.  Returns (goal_poses (N,2,3), BatchResult); ``int(result.reachable.sum())`` is the reference's
    "reachable poses" count."""
    poses = task_space_grid(solver.shoulder_position, arm_length, **steps)
    return poses, solver.is_reachable_batch(poses, want_joints=False, precision=precision)


def save_reach_map(path: str, counts, origin, step, orientations_euler, arm: str = "") -> None:
    """Write a reachability count volume with its grid (cell centres = origin + index * step) and orientation set to a
    compressed ``.npz``."""
    c = counts.cpu().numpy() if hasattr(counts, "cpu") else np.asarray(counts)
    np.savez_compressed(path, counts=c.astype(np.uint32), origin=np.asarray(origin, dtype=np.float64),
                        step=np.asarray(step, dtype=np.float64), orientations_euler=np.asarray(orientations_euler, dtype=np.float64),
                        arm=np.array(arm))


def load_reach_map(path: str) -> dict:
    """Read a volume written by ``save_reach_map``; adds ``fraction`` = counts / number of orientations."""
    with np.load(path, allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    d["arm"] = str(d["arm"])
    d["fraction"] = d["counts"].astype(np.float64) / max(len(d["orientations_euler"]), 1)
    return d
