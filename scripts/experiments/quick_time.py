"""Quick device-side timing of the K1 kernel (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import SymbolicIK, fk, _native
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = fk.sample_fk_poses(n, "r_arm", seed=1)
ik = SymbolicIK(arm="r_arm")
Md = torch.from_numpy(M).cuda().reshape(n, 16)
reach = torch.empty(n, dtype=torch.uint8, device="cuda"); state = torch.empty_like(reach)
itv = torch.empty((n, 2), dtype=torch.float64, device="cuda"); j = torch.empty((n, 7), dtype=torch.float64, device="cuda")
e = torch.empty((n, 3), dtype=torch.float64, device="cuda")
for kind, P in ((1, Md),):
    for _ in range(3):
        ik.solve_into(P, kind, None, None, reach, state, itv, j, e)
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 10
    for _ in range(reps):
        ik.solve_into(P, kind, None, None, reach, state, itv, j, e)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"K1 kind={kind} n={n}: {ms:.3f} ms/launch -> {n / ms * 1e3:.3e} poses/s; reachable {reach.float().mean().item():.3f}")
ms = C.c_double(); fl = C.c_double()
_native.check(_native.load().r2ik_dfma_probe(0, 200000, C.byref(ms), C.byref(fl), None), "probe")
print(f"DFMA probe: {fl.value / ms.value / 1e9:.2f} TFLOP/s FP64 ({ms.value:.2f} ms)")
