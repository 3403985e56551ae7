"""K1 reading its poses straight from pinned host memory (UVA): only the 96 bytes of a 4x4 matrix the kernel loads cross
PCIe, against 128 B for a DMA copy of the whole rows.  Times the kernel alone, and with a D2H copy of the previous
chunk's outputs running beside it (the pipeline's steady state).
    python scripts/experiments/exp_r2_zero_copy_in.py"""
import sys

import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import SymbolicIK, _abi, fk  # noqa: E402

n = 1 << 20
M = fk.sample_fk_poses(n, "r_arm", seed=1)
ik = SymbolicIK(arm="r_arm")
dev = torch.device("cuda")
host = torch.from_numpy(M.reshape(n, 16)).pin_memory()
dM = host.to(dev)
reach = torch.empty(n, dtype=torch.uint8, device=dev); state = torch.empty_like(reach)
itv = torch.empty((n, 2), dtype=torch.float64, device=dev); j = torch.empty((n, 7), dtype=torch.float64, device=dev)
e = torch.empty((n, 3), dtype=torch.float64, device=dev)
hj = torch.empty((n, 7), dtype=torch.float64).pin_memory(); hi = torch.empty((n, 2), dtype=torch.float64).pin_memory()
he = torch.empty((n, 3), dtype=torch.float64).pin_memory()
side = torch.cuda.Stream()


def timed(fn, reps=8):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def d2h():
    with torch.cuda.stream(side):
        hj.copy_(j, non_blocking=True); hi.copy_(itv, non_blocking=True); he.copy_(e, non_blocking=True)


for name, P in (("device-resident poses", dM), ("host-resident poses (zero copy)", host)):
    ik.solve_into(P, _abi.POSE_MAT4, None, None, reach, state, itv, j, e)
    t = timed(lambda: ik.solve_into(P, _abi.POSE_MAT4, None, None, reach, state, itv, j, e))
    print(f"{name}: {t * 1e3:8.1f} us per {n} poses = {n / (t * 1e-3):.3e} poses/s; 96 B/pose -> {96 * n / t / 1e6:.1f} GB/s, 128 B/pose -> {128 * n / t / 1e6:.1f} GB/s")

    def both():
        d2h()
        ik.solve_into(P, _abi.POSE_MAT4, None, None, reach, state, itv, j, e)
        torch.cuda.current_stream().wait_stream(side)
    t = timed(both)
    print(f"   with the D2H of 96 B/pose of outputs beside it: {t * 1e3:8.1f} us = {n / (t * 1e-3):.3e} poses/s")
t = timed(lambda: dM.copy_(host, non_blocking=True))
print(f"DMA H2D of the 128 B rows alone: {t * 1e3:8.1f} us = {n / (t * 1e-3):.3e} poses/s ({128 * n / t / 1e6:.1f} GB/s)")


def dma_both():
    d2h()
    dM.copy_(host, non_blocking=True)
    torch.cuda.current_stream().wait_stream(side)


t = timed(dma_both)
print(f"DMA H2D with the D2H beside it: {t * 1e3:8.1f} us = {n / (t * 1e-3):.3e} poses/s")
