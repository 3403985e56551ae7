"""K1 variants: scalar-store vs staged (transposed 128-bit) stores, MAT4 (128 B) vs MAT34 (96 B) vs EULER6 (48 B) input, lean outputs.
    python scripts/experiments/exp_r2_k1.py [lib.so]      (run under ncu for sectors / request and pipe utilisation)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import _native  # noqa: E402

args = sys.argv[1:]
tag = "in-tree"
if args and args[0].endswith(".so"):
    tag = args[0].split("libr2ik_")[-1][:-3]
    _native.use_library(args.pop(0))
from reachy2_symbolic_ik_b200 import SymbolicIK, _abi, fk  # noqa: E402
from oracle import oracle as O  # noqa: E402
from scipy.spatial.transform import Rotation as R  # noqa: E402

n = 1_000_000
M = fk.sample_fk_poses(n, "r_arm", seed=1)
ik = SymbolicIK(arm="r_arm")
dev = torch.device("cuda")
lay = {"mat4": (torch.from_numpy(M.reshape(n, 16)).to(dev), _abi.POSE_MAT4),
       "mat34": (torch.from_numpy(np.ascontiguousarray(M[:, :3, :].reshape(n, 12))).to(dev), _abi.POSE_MAT34),
       "euler6": (torch.from_numpy(np.ascontiguousarray(np.concatenate([M[:, :3, 3], R.from_matrix(M[:, :3, :3]).as_euler("xyz")], axis=1))).to(dev), _abi.POSE_EULER6)}
reach = torch.empty(n, dtype=torch.uint8, device=dev); state = torch.empty_like(reach)
itv = torch.empty((n, 2), dtype=torch.float64, device=dev); j = torch.empty((n, 7), dtype=torch.float64, device=dev)
e = torch.empty((n, 3), dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
want = O.symik_batch(O.arm_config("r_arm"), M[:50_000])
for name, (P, kind) in lay.items():
    for outs, oa in (("all", (reach, state, itv, j, e)), ("lean", (None, state, None, j, None))):
        for _ in range(3):
            ik.solve_into(P, kind, None, None, *oa)
        ts = []
        for _ in range(12):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); ik.solve_into(P, kind, None, None, *oa); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        dj = np.nanmax(np.abs(j[:50_000].cpu().numpy() - want[3]))
        ok = np.array_equal(state[:50_000].cpu().numpy(), want[2])
        print(f"{tag:10s} {name:7s} outputs {outs:4s}: {sorted(ts)[len(ts) // 2]:6.1f} us / 1M (min {min(ts):.1f})  max|dj| {dj:.1e} states ok {ok}", flush=True)
