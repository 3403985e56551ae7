"""Host soak of the FP32 fast path (the kernel source of csrc/r2ik_device_f32.cuh compiled for the host, tests/hostsim) against
the FP64 oracle on the same float32 inputs: the stated bound of include/r2ik.h -- states identical, joints / intervals within
1e-4 rad for >= 99.99 % of the poses and within 3e-4 rad for all -- on millions of poses.

    python scripts/soak_f32_bound.py [poses_per_set] [seed]
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]

import test_hostsim_parity as T  # noqa: E402
from oracle import oracle as O  # noqa: E402
from reachy2_symbolic_ik_b200 import fk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
O.build()
O.use_all_host_threads()
subprocess.run(["make", "-C", T.HS_DIR], check=True, capture_output=True)
hs = C.CDLL(os.path.join(T.HS_DIR, "_build", "libr2ik_hostsim.so"))
t0 = time.time()
tot = dict(n=0, reach=0, over1=0, over3=0, mism=0, esc=0)
worst = 0.0
for arm in ("r_arm", "l_arm"):
    for kind, P in (("fk x>0.05", fk.sample_fk_poses(n, arm, seed=seed)), ("fk all", fk.sample_fk_poses(n, arm, seed=seed + 1, min_x=None)),
                    ("task space", fk.sample_task_space_poses(n, arm, seed=seed + 2)), ("euler layout", None)):
        if P is None:
            from scipy.spatial.transform import Rotation as R

            M = fk.sample_fk_poses(n, arm, seed=seed + 3)
            P32 = np.concatenate([M[:, :3, 3], R.from_matrix(M[:, :3, :3]).as_euler("xyz")], axis=1).astype(np.float32).reshape(n, 2, 3)
        else:
            P32 = P.astype(np.float32)
        w = O.symik_batch(O.arm_config(arm), P32.astype(np.float64))
        r, itv, st, j, e, esc = T.hs_symik_f32(hs, T.cfg_for(arm), P32)
        err = np.maximum(np.nan_to_num(np.abs(j - w[3])).max(axis=1), np.nan_to_num(np.abs(itv - w[1])).max(axis=1))
        ok = w[2] == 0
        mism = int((st != w[2]).sum())
        q = np.quantile(err[ok], 0.9999) if ok.any() else 0.0
        print(f"{arm} {kind:12s}: n {n} reachable {ok.mean():.3f} state mismatches {mism}; err p50 {np.median(err[ok]):.1e} p99.99 {q:.2e} "
              f"p99.999 {np.quantile(err[ok], 0.99999):.2e} max {err.max():.2e}; over 1e-4: {int((err > 1e-4).sum())}, over 3e-4: {int((err > 3e-4).sum())}; "
              f"escalated {esc.mean():.4f} [{time.time() - t0:.0f}s]", flush=True)
        tot["n"] += n; tot["reach"] += int(ok.sum()); tot["over1"] += int((err > 1e-4).sum()); tot["over3"] += int((err > 3e-4).sum())
        tot["mism"] += mism; tot["esc"] += int(esc.sum()); worst = max(worst, float(err.max()))
print(f"TOTAL: {tot['n']} poses ({tot['reach']} reachable), state mismatches {tot['mism']}, over 1e-4: {tot['over1']} "
      f"({tot['over1'] / max(tot['reach'], 1):.2e} of the reachable), over 3e-4: {tot['over3']}, max {worst:.2e}, escalated {tot['esc'] / tot['n']:.4f}")
sys.exit(1 if (tot["mism"] or tot["over3"] or tot["over1"] > 1e-4 * tot["reach"]) else 0)
