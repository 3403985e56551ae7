"""Every kernel of libr2ik.so at small sizes, for compute-sanitizer (memcheck / racecheck / synccheck on the
kernels that use shared memory, warp votes and atomics).  Ragged sizes on purpose (partial warps, partial tiles).

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitizer_driver.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import ControlIK, SymbolicIK, fk  # noqa: E402

torch.cuda.init()
ran = []
for arm in ("r_arm", "l_arm"):
    ik = SymbolicIK(arm=arm)
    M = np.concatenate([fk.sample_fk_poses(1500, arm, seed=3), fk.sample_task_space_poses(1237, arm, seed=4)])
    r = ik.is_reachable_batch(M); ran.append("k_symik_solve<MAT4>")
    ik.is_reachable_batch(np.ascontiguousarray(M[:, :3, :])); ran.append("k_symik_solve<MAT34>")
    from scipy.spatial.transform import Rotation as R

    gp = np.concatenate([M[:, :3, 3], R.from_matrix(M[:, :3, :3]).as_euler("xyz")], axis=1)
    ik.is_reachable_batch(gp, theta=np.linspace(-3, 3, len(gp))); ran.append("k_symik_solve<EULER6>")
    ik.is_reachable_batch(M.astype(np.float32), precision="fp32"); ran.append("k_symik_solve_f32 + k_symik_escalated_f32")
    ik.is_reachable_batch(gp.astype(np.float32), precision="fp32")
    ik.is_reachable_no_limits_batch(M, np.zeros(len(M)), np.ones((len(M), 7)), with_projected=True); ran.append("k_symik_no_limits")
    ik.get_elbow_position_batch(M, np.zeros((len(M), 3)), with_projected=True); ik.get_elbow_position_batch(M, np.zeros((len(M), 3)), no_limits=True)
    ran.append("k_elbow_positions")
    ok, itv, f, st = ik.is_reachable([[0.3, -0.1 if arm == "r_arm" else 0.1, 0.1], [0.3, -0.8, 0.3]])
    if ok:
        f(itv[0])
    ik.is_reachable_no_limits([[0.9, 0.0, 0.0], [0, 0, 0]]); ik.get_elbow_position(0.3); ran.append("k_symik_scalar")
    ik.reach_map(n=24, n_orientations=40); ik.reach_map(n=24, n_orientations=40, all_fp64=True); ran.append("k_reach_map<u32>, k_reach_map_f64")
    Mh = torch.from_numpy(M).pin_memory()
    ik.is_reachable_batch_host(Mh, chunk=700); ik.is_reachable_batch_host(torch.from_numpy(gp).pin_memory(), chunk=512, want=ik.LEAN)
    ik.is_reachable_batch_host(Mh.float(), chunk=700, precision="fp32"); ran.append("r2ik_pipeline_symik_f64 / f32")
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf"); ran.append("k_ctl_ctor_theta")
    ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete", compact=False); ran.append("k_ctl_discrete")
    ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete", compact=True); ran.append("k_disc_consts/classify/search/finish")
    ctl.nb_search_points = 360
    ctl.symbolic_inverse_kinematics_batch(arm, M[:1100], "discrete", exhaustive=True); ran.append("k_ctl_discrete_scan")
    ctl.symbolic_inverse_kinematics_batch_host(arm, Mh, "discrete", chunk=600); ran.append("r2ik_pipeline_ctl_discrete_f64")
    ctl.symbolic_inverse_kinematics(arm, M[0], "discrete"); ctl.symbolic_inverse_kinematics(arm, M[1], "continuous")
    MT = fk.sinusoidal_trajectories(77, 53, arm, seed=5)[0].copy()
    MT[3, 13:, :3, :3] = MT[3, 13:, :3, :3] @ np.diag([-1.0, -1.0, 1.0])     # emergency latch
    MT[5, 7, :3, :3] = np.diag([-1.0, 1.0, 1.0])                             # invalid rotation
    for phased in (False, True):
        ctl.symbolic_inverse_kinematics_batch(arm, MT, "continuous", phased=phased)
        ctl.symbolic_inverse_kinematics_batch(arm, MT, "continuous", phased=phased, _test_force_serial_mod=7) if phased else None
    ran.append("k_ctl_continuous, k_cont_targets / thetas / raw_joints_codes / finish_codes / apply_windings")
    ctl.symbolic_inverse_kinematics_batch_host(arm, torch.from_numpy(MT), "continuous", chunk=16)
    fk.forward_kinematics_device(torch.zeros((333, 7), dtype=torch.float64, device="cuda"), arm); ran.append("k_fk")
torch.cuda.synchronize()
print("sanitizer driver ran:", "; ".join(dict.fromkeys(ran)))
