"""K1 at the three L2 fetch granularities (cudaLimitMaxL2FetchGranularity 32 / 64 / 128 B): the pose rows are 128 B of which
96 are read -- does the unused fourth sector still come from DRAM?  Run under ncu --metrics dram__bytes_read.sum for the bytes.

    python scripts/experiments/exp_r2_l2fetch.py <granularity>
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import SymbolicIK, fk  # noqa: E402

gran = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
if gran:
    rc = rt.cudaDeviceSetLimit(C.c_int(5), C.c_size_t(gran))      # cudaLimitMaxL2FetchGranularity = 0x05
    val = C.c_size_t()
    rt.cudaDeviceGetLimit(C.byref(val), C.c_int(5))
    print("cudaDeviceSetLimit rc", rc, "-> granularity", val.value)
n = 1_000_000
M = fk.sample_fk_poses(n, "r_arm", seed=1)
ik = SymbolicIK(arm="r_arm")
Md = torch.from_numpy(M).cuda().reshape(n, 16)
reach = torch.empty(n, dtype=torch.uint8, device="cuda"); state = torch.empty_like(reach)
itv = torch.empty((n, 2), dtype=torch.float64, device="cuda"); j = torch.empty((n, 7), dtype=torch.float64, device="cuda")
e = torch.empty((n, 3), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ik.solve_into(Md, 1, None, None, reach, state, itv, j, e)
ts = []
for _ in range(10):
    flush.zero_()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    ik.solve_into(Md, 1, None, None, reach, state, itv, j, e)
    ev1.record(); torch.cuda.synchronize()
    ts.append(ev0.elapsed_time(ev1) * 1e3)
print(f"granularity {gran or 'default'}: K1 {sorted(ts)[len(ts) // 2]:.1f} us / 1M poses (L2 flushed), min {min(ts):.1f}")
