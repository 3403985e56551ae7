"""GPU tests of K4 (workspace reachability map) and of the FK kernel, through the facade / C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ARMS = ("r_arm", "l_arm")


@pytest.mark.parametrize("arm", ARMS)
def test_reach_map_matches_oracle(oracle, arm):
    """Every voxel count equals the oracle's (sum over orientations of is_reachable flags); voxels whose
    flag flips under a 3e-13 perturbation of the pose are boundary cases and only counted."""
    from reachy2_symbolic_ik_b200 import SymbolicIK, fk, workspace

    ik = SymbolicIK(arm=arm)
    n = 20
    ori = fk.fibonacci_orientations(24)
    origin, step, dims = workspace.reach_grid(ik.shoulder_position, ik.max_arm_length, n)
    got = ik.reach_map(n=n, orientations_euler=ori).cpu().numpy()
    want = oracle.reach_map(oracle.arm_config(arm), origin, step, dims, ori)
    assert got.shape == (n, n, n) and got.dtype == np.int32
    diff = got.astype(np.int64) - want.astype(np.int64)
    # allow a handful of boundary voxels (|diff| <= 1) but nothing systematic
    assert np.abs(diff).max() <= 1, np.abs(diff).max()
    assert (diff != 0).sum() <= 3, (diff != 0).sum()
    assert want.sum() > 0 and (want == 0).sum() > 0
    # orientation slices add up (the multi-GPU decomposition of workspace.sharded_sum)
    parts = np.zeros_like(got)
    for r in range(3):
        b, e = workspace.shard_range(len(ori), r, 3)
        parts += ik.reach_map(n=n, orientations_euler=ori[b:e]).cpu().numpy()
    assert np.array_equal(parts, got)


def test_reach_map_flags_are_is_reachable(oracle):
    """counts with ONE orientation == the K1 reachability flag of the same poses."""
    from reachy2_symbolic_ik_b200 import SymbolicIK, workspace

    ik = SymbolicIK(arm="r_arm")
    n = 16
    e = np.array([[0.3, -1.2, 0.4]])
    origin, step, dims = workspace.reach_grid(ik.shoulder_position, ik.max_arm_length, n)
    got = ik.reach_map(n=n, orientations_euler=e).cpu().numpy().reshape(-1)
    idx = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    P = np.concatenate([origin + idx * step, np.broadcast_to(e, (n ** 3, 3))], axis=1)
    res = ik.is_reachable_batch(P, want_joints=False)
    assert np.array_equal(got.astype(bool), res.reachable)


@pytest.mark.parametrize("arm", ARMS)
def test_fk_kernel_matches_host_fk(arm):
    import torch

    from reachy2_symbolic_ik_b200 import fk

    rng = np.random.default_rng(11)
    q = fk.sample_fk_joints(5000, rng)
    want = fk.forward_kinematics(q, arm)
    got = fk.forward_kinematics_device(torch.from_numpy(q).cuda(), arm).cpu().numpy()
    assert got.shape == (5000, 4, 4)
    np.testing.assert_allclose(got, want, atol=1e-14, rtol=0)
    assert np.array_equal(got[:, 3], np.broadcast_to([0, 0, 0, 1.0], (5000, 4)))


@pytest.mark.parametrize("arm", ARMS)
def test_fk_round_trip_full_size(arm):
    """Size-independent property at BASELINE size (1M poses): FK(get_joints(theta)) reproduces the goal pose
    whenever no goal-shifting branch fired (ControlIK's singularity_offset = -1.01 disables the elbow
    projection).  The URDF's truncated rpy literals bound the agreement at ~1e-5."""
    import torch

    from reachy2_symbolic_ik_b200 import SymbolicIK, fk

    n = 1_000_000
    rng = np.random.default_rng(21)
    q = torch.from_numpy(fk.sample_fk_joints(n, rng)).cuda()
    M = fk.forward_kinematics_device(q, arm)
    ik = SymbolicIK(arm=arm, singularity_offset=-1.01)
    res = ik.is_reachable_batch(M)
    ok = res.reachable
    assert 0.3 < ok.double().mean().item() < 0.7   # about half of raw FK samples are backward poses
    M2 = fk.forward_kinematics_device(res.joints[ok], arm)
    # poses whose wrist was pushed forward / outward by is_reachable are legitimately shifted: compare
    # orientation always, position only where the solver did not move the goal
    rot_err = (M2[:, :3, :3] - M[ok][:, :3, :3]).abs().amax(dim=(1, 2))
    assert rot_err.quantile(0.999).item() < 1e-4
    pos_err = (M2[:, :3, 3] - M[ok][:, :3, 3]).norm(dim=1)
    assert pos_err.median().item() < 1e-5
    assert (pos_err < 1e-4).double().mean().item() > 0.9


@pytest.mark.parametrize("arm", ARMS)
def test_reach_map_mixed_equals_all_fp64(arm):
    """K4 decides a (voxel, orientation) pair with an FP64 front end + FP32 linking test and hands the pairs it cannot
    call to the FP64 flag solve: the volume must equal the all-FP64 kernel's, voxel by voxel."""
    from reachy2_symbolic_ik_b200 import SymbolicIK, fk

    ik = SymbolicIK(arm=arm)
    ori = fk.fibonacci_orientations(96)
    a = ik.reach_map(n=112, orientations_euler=ori).cpu().numpy()
    b = ik.reach_map(n=112, orientations_euler=ori, all_fp64=True).cpu().numpy()
    assert a.sum() > 1_000_000
    np.testing.assert_array_equal(a, b)
    # an orientation slice, as a rank of the sharded map computes it
    from reachy2_symbolic_ik_b200 import workspace
    origin, step, dims = workspace.reach_grid(ik.shoulder_position, ik.max_arm_length, 64)
    c = ik.reach_map(orientations_euler=ori, origin=origin + 1e-3, step=step, dims=dims).cpu().numpy()
    d = ik.reach_map(orientations_euler=ori, origin=origin + 1e-3, step=step, dims=dims, all_fp64=True).cpu().numpy()
    np.testing.assert_array_equal(c, d)


def test_reach_map_full_size_mixed_equals_all_fp64():
    """BASELINE configs[4] at full size (256^3 voxels x 512 orientations, 2.1e9 live pairs): the mixed-precision volume
    equals the all-FP64 one voxel by voxel.  (Two pairs of this very map sit 3e-9 from the discriminant's zero with the
    planes 3 degrees from parallel; they fixed the width of the FP32 error band of the test.)"""
    from reachy2_symbolic_ik_b200 import SymbolicIK, fk

    ik = SymbolicIK(arm="r_arm")
    ori = fk.fibonacci_orientations(512)
    a = ik.reach_map(n=256, orientations_euler=ori).clone()
    b = ik.reach_map(n=256, orientations_euler=ori, all_fp64=True)
    assert int((a != b).sum().item()) == 0
    assert int(a.sum().item()) == 856520774


def test_task_space_sweep_matches_reference():
    """SymbolicIK.task_space_test = the reference's task_space_test sweep (ik_comparison.py:137-181) in one launch:
    flags and states of all 37 376 poses equal the reference's (tests/golden/task_space.npz), FP64 and FP32 paths."""
    from parity import load
    from reachy2_symbolic_ik_b200 import SymbolicIK

    g = load("task_space.npz")
    ik = SymbolicIK()
    want = np.unpackbits(g["reachable_packed"])[: int(g["n_poses"])].astype(bool)
    for precision in ("fp64", "fp32"):
        poses, res = ik.task_space_test(precision=precision)
        assert len(poses) == 37376
        # the grid angles are multiples of 45 degrees: many poses sit exactly on the reference's special cases
        # (the FP32 path sees the float32-rounded grid; its states may differ only where that rounding crosses a boundary)
        mism = int((res.state != g["state"]).sum())
        assert mism == 0 if precision == "fp64" else mism <= 40, (precision, mism)
        if precision == "fp64":
            assert np.array_equal(res.reachable, want) and int(res.reachable.sum()) == 13492


def _sharded_worker(rank, world, port, tmp):
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from reachy2_symbolic_ik_b200 import SymbolicIK, fk

        ik = SymbolicIK(arm="l_arm", device=rank)
        for n, n_ori in ((40, 37), (33, 64)):           # even and odd plane sizes (the odd one takes the unsplit path)
            ori = fk.fibonacci_orientations(n_ori)
            single = ik.reach_map(n=n, orientations_euler=ori)
            for shard in ("voxels", "orientations"):
                timing = {}
                sharded = ik.reach_map(n=n, orientations_euler=ori, dist=dist, timing=timing, shard=shard)
                assert torch.equal(sharded, single), f"{shard}-sharded map (16-bit, live range, slabs) differs from the single-GPU map ({n}, {n_ori})"
                assert timing["exchanged_bytes"] <= 2 * n ** 3
            plain = ik.reach_map(n=n, orientations_euler=ori, dist=dist, plain_allreduce=True)
            assert torch.equal(plain, single), "plain all-reduce differs from the single-GPU map"
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_reach_map_sharded_two_gpus(tmp_path):
    """cfg 5 as it runs on several GPUs: orientation shards, 16-bit counts two per all-reduce lane, only the x-range that
    can hold a reachable voxel exchanged, slab-pipelined -- equal to the single-GPU volume and to the plain all-reduce."""
    import socket

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))
