"""Size-independent properties of K2 and K3 at BASELINE.json's full sizes (configs[2]: 1M poses x 360 samples,
configs[3]: 65 536 trajectories x 1 000 waypoints), where the oracle would take hours: the two forms of each kernel
agree, results that must be constant are constant, joints reproduce the goal pose through forward kinematics, and a
trajectory that has not latched an emergency stop moves by less than the continuity limits per waypoint.
(K1 and K4 have theirs in tests/test_gpu_workspace.py.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_discrete_full_size_properties():
    import torch

    from reachy2_symbolic_ik_b200 import ControlIK, fk

    n, K = 1_000_000, 360
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    ctl.nb_search_points = K
    M = torch.from_numpy(fk.sample_fk_poses(n, "r_arm", seed=3)).cuda()
    a = ctl.symbolic_inverse_kinematics_batch("r_arm", M, "discrete", compact=True)
    b = ctl.symbolic_inverse_kinematics_batch("r_arm", M, "discrete", compact=False)
    for x, y in zip(a, b):      # three dense passes == one kernel, byte for byte
        assert torch.equal(x.view(torch.uint8) if x.dtype == torch.bool else x, y.view(torch.uint8) if y.dtype == torch.bool else y)
    joints, reach, state, emg = a
    assert 0.55 < reach.double().mean().item() < 0.72
    assert bool(((state == 0) == reach).all()) and not bool(emg.any())
    # every pose without a valid theta returns the same row: the current joints through the safety chain
    lost = joints[~reach]
    assert bool((lost == lost[0]).all())
    # joints of the found poses reproduce the goal through FK (URDF rpy literals are truncated: ~1e-5), unless the
    # Orbita3D cone limit changed the wrist
    idx = torch.nonzero(reach).flatten()
    M2 = fk.forward_kinematics_device(joints[idx], "r_arm")
    rot_err = (M2[:, :3, :3] - M.reshape(n, 4, 4)[idx][:, :3, :3]).abs().amax(dim=(1, 2))
    pos_err = (M2[:, :3, 3] - M.reshape(n, 4, 4)[idx][:, :3, 3]).norm(dim=1)
    assert (rot_err < 1e-4).double().mean().item() > 0.9      # the rest: wrist clamped to the 42.5 degree cone
    assert pos_err[rot_err < 1e-4].median().item() < 1e-5


def test_continuous_full_size_properties():
    import torch

    from reachy2_symbolic_ik_b200 import ControlIK, fk

    T, W = 65_536, 1_000
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    M = fk.sinusoidal_trajectories_device(T, W, "r_arm", seed=4)
    a = ctl.symbolic_inverse_kinematics_batch("r_arm", M, "continuous", phased=True)
    torch.cuda.synchronize()
    joints, reach, state, st = a
    # per-trajectory scans on winding codes == the one-thread-per-trajectory recursion (a quarter of the batch: the serial
    # kernel is the slow cross-check)
    Tq = T // 4
    b = ctl.symbolic_inverse_kinematics_batch("r_arm", M[:Tq], "continuous", phased=False)
    assert torch.equal(reach[:Tq], b[1]) and torch.equal(state[:Tq], b[2])
    assert (joints[:Tq] - b[0]).abs().max().item() < 1e-12
    for f in ("emergency_stop", "emergency_bits", "init", "has_previous_sol"):
        assert np.array_equal(st[f][:Tq], b[3][f])
    del b
    # continuity: wherever the controller has not latched an emergency stop, consecutive waypoints differ by less than
    # the limits of continuity_check (utils.py:571-589)
    max_step = torch.tensor([0.5, 0.5, 0.5, 0.5, 1.0, 1.0, 1.0], dtype=torch.float64, device=joints.device)
    moving = torch.from_numpy(st["emergency_stop"] == 0).to(joints.device)
    worst = torch.zeros(7, dtype=torch.float64, device=joints.device)
    for lo in range(0, T, 8192):
        d = (joints[lo:lo + 8192, 1:] - joints[lo:lo + 8192, :-1]).abs()
        d = d[moving[lo:lo + 8192]]
        if d.numel():
            worst = torch.maximum(worst, d.amax(dim=(0, 1)))
    assert bool((worst <= max_step + 1e-12).all()), worst
    assert moving.double().mean().item() > 0.5
    # a latched trajectory repeats its last solution and reports the emergency state
    halted = torch.nonzero(~moving).flatten()[:64]
    for t in halted.tolist():
        assert int(state[t, -1]) == 8 and torch.equal(joints[t, -1], joints[t, -2])
    # reached waypoints carry the empty state string, and nothing is NaN (no invalid rotation in the workload)
    assert bool((state[reach] == 7).all()) and not bool(torch.isnan(joints).any())
