"""Find the voxels where the mixed-precision reach map differs from the all-FP64 one (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reachy2_symbolic_ik_b200 import SymbolicIK, fk
for arm in ("r_arm", "l_arm"):
    ik = SymbolicIK(arm=arm)
    ori = torch.from_numpy(fk.fibonacci_orientations(512)).cuda()
    a = ik.reach_map(n=256, orientations_euler=ori).clone()
    b = ik.reach_map(n=256, orientations_euler=ori, all_fp64=True)
    d = (a - b)
    idx = torch.nonzero(d).cpu().numpy()
    print(arm, "sum", int(a.sum()), int(b.sum()), "differing voxels", len(idx))
    for i in idx[:20]:
        print("  voxel", i.tolist(), "mixed", int(a[tuple(i)]), "f64", int(b[tuple(i)]))
