"""K4 timing (development aid): reach map 256^3 x 512 of $R2IK_LIB or the in-tree library, checked against a reference sum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reachy2_symbolic_ik_b200 import SymbolicIK, fk

ik = SymbolicIK(arm="r_arm")
ori = torch.from_numpy(fk.fibonacci_orientations(512)).cuda()
out = torch.empty((256,) * 3, dtype=torch.int32, device="cuda")
for _ in range(2):
    ik.reach_map(n=256, orientations_euler=ori, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ik.reach_map(n=256, orientations_euler=ori, out=out)
e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('R2IK_LIB', 'in-tree'):70s} K4 {e0.elapsed_time(e1) / 5:7.3f} ms  total={int(out.sum().item())}", flush=True)
