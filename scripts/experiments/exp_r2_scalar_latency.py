"""Scalar-call latency of the drop-in classes on the GPU box (what a control loop pays per tick), beside the unmodified
Python reference from baseline/_ref when it is installed.  Prints one JSON line.

    python scripts/experiments/exp_r2_scalar_latency.py [n_calls]
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def timeit(fn, n, warm=20):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    return {"median_us": float(np.median(ts)), "p90_us": float(np.quantile(ts, 0.9)), "min_us": float(ts.min())}


def measure(n=400):
    from reachy2_symbolic_ik_b200 import ControlIK, SymbolicIK, fk

    with contextlib.redirect_stdout(io.StringIO()):
        ik = SymbolicIK(arm="r_arm")
        ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    goal = [[0.3, -0.1, 0.1], [np.radians(20), np.radians(-50), np.radians(20)]]     # src/benchmark/ik_benchmarks.py:13-14
    M = fk.sample_fk_poses(64, "r_arm", seed=5)
    out = {}

    def symik():
        ok, itv, f, _ = ik.is_reachable(goal)
        if ok:
            f(itv[0])

    out["SymbolicIK.is_reachable+get_joints"] = timeit(symik, n)
    out["SymbolicIK.is_reachable"] = timeit(lambda: ik.is_reachable(goal), n)
    k = [0]

    def discrete():
        k[0] += 1
        ctl.symbolic_inverse_kinematics("r_arm", M[k[0] % 64], "discrete")

    out["ControlIK.symbolic_inverse_kinematics(discrete)"] = timeit(discrete, n)
    traj = fk.sinusoidal_trajectories(1, n + 40, "r_arm", seed=6)[0][0]
    k[0] = 0

    def continuous():
        ctl.symbolic_inverse_kinematics("r_arm", traj[k[0] % len(traj)], "continuous")
        k[0] += 1

    out["ControlIK.symbolic_inverse_kinematics(continuous)"] = timeit(continuous, n)

    ref = os.path.join(REPO, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "reachy2_symbolic_ik")):
        sys.path.insert(0, ref)
        with contextlib.redirect_stdout(io.StringIO()):
            from reachy2_symbolic_ik.control_ik import ControlIK as RC
            from reachy2_symbolic_ik.symbolic_ik import SymbolicIK as RS

            rik = RS(arm="r_arm")
            rctl = RC(urdf=open(fk.bundled_urdf_path()).read())

            def rsym():
                ok, itv, f, _ = rik.is_reachable(np.array(goal))
                if ok:
                    f(itv[0])

            out["reference SymbolicIK.is_reachable+get_joints"] = timeit(rsym, 100, 5)
            k[0] = 0

            def rdis():
                k[0] += 1
                rctl.symbolic_inverse_kinematics("r_arm", M[k[0] % 64], "discrete")

            out["reference ControlIK(discrete)"] = timeit(rdis, 100, 5)
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    real = os.dup(1)
    os.dup2(2, 1)
    res = measure(n)
    os.write(real, (json.dumps({"scalar_latency": res}) + "\n").encode())
