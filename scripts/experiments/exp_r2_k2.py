"""K2 (ControlIK discrete, 1M poses x 360 samples): the one-kernel form against the compacted three-pass form.
    python scripts/experiments/exp_r2_k2.py [lib.so] [--once]     (--once: one call of each form, for ncu)"""
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
once = "--once" in args
from reachy2_symbolic_ik_b200 import ControlIK, fk  # noqa: E402

n = 1_000_000
ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
ctl.nb_search_points = 360
M = torch.from_numpy(fk.sample_fk_poses(n, "r_arm", seed=3)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for compact in (False, True):
    out = None
    for _ in range(1 if once else 3):
        out = ctl.symbolic_inverse_kinematics_batch("r_arm", M, "discrete", out=out, compact=compact)
    ts = []
    for _ in range(0 if once else 15):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ctl.symbolic_inverse_kinematics_batch("r_arm", M, "discrete", out=out, compact=compact); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    # back to back without flushing (the bench's step)
    if not once:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ctl.symbolic_inverse_kinematics_batch("r_arm", M, "discrete", out=out, compact=compact)
        e1.record(); torch.cuda.synchronize()
        print(f"{tag:12s} compact={compact!s:5s}: {sorted(ts)[len(ts) // 2]:7.1f} us / 1M x 360 after an L2 flush (min {min(ts):.1f}); "
              f"{e0.elapsed_time(e1) * 1e3 / 50:7.1f} us back to back", flush=True)
    res[compact] = [x.cpu().numpy() for x in out]
same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res[False], res[True]))
print(f"{tag:12s} identical outputs: {same}; found {res[True][1].mean():.3f}")
if not once:
    for nn in (1 << 12, 1 << 14, 1 << 15, 1 << 16, 1 << 18):
        line = []
        for compact in (False, True):
            Ms = M[:nn]
            o = ctl.symbolic_inverse_kinematics_batch("r_arm", Ms, "discrete", compact=compact)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                ctl.symbolic_inverse_kinematics_batch("r_arm", Ms, "discrete", out=o, compact=compact)
            e1.record(); torch.cuda.synchronize()
            line.append(e0.elapsed_time(e1) * 10)
        print(f"{tag:12s} n = {nn:7d}: single {line[0]:6.1f} us, compact {line[1]:6.1f} us")
