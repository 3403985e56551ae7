"""Timeline of the lean host pipeline (48 B in, 57 B out per pose) rebuilt with torch streams + timing events, with and without
the kernel, to see where the link idles."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import SymbolicIK, fk, _abi  # noqa: E402
from scipy.spatial.transform import Rotation as R  # noqa: E402

n = 1_000_000
M = fk.sample_fk_poses(n, "r_arm", seed=1)
ik = SymbolicIK(arm="r_arm")
gp = torch.from_numpy(np.ascontiguousarray(np.concatenate([M[:, :3, 3], R.from_matrix(M[:, :3, :3]).as_euler("xyz")], axis=1))).pin_memory()
h_j = torch.empty((n, 7), dtype=torch.float64).pin_memory()
h_s = torch.empty(n, dtype=torch.uint8).pin_memory()
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
C = 1 << 18
slots = [dict(p=torch.empty((C, 6), dtype=torch.float64, device="cuda"), j=torch.empty((C, 7), dtype=torch.float64, device="cuda"),
              s=torch.empty(C, dtype=torch.uint8, device="cuda")) for _ in range(3)]


def run(kernel, state_copy, timeline=False):
    evs = []
    k_done = [None] * 3; d_done = [None] * 3
    t_origin = torch.cuda.Event(enable_timing=True); t_origin.record(s_in)
    for ci, lo in enumerate(range(0, n, C)):
        hi = min(n, lo + C); m = hi - lo
        b = slots[ci % 3]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        if k_done[ci % 3] is not None:
            s_in.wait_event(k_done[ci % 3])
        with torch.cuda.stream(s_in):
            e[0].record(s_in); b["p"][:m].copy_(gp[lo:hi], non_blocking=True); e[1].record(s_in)
        s_k.wait_event(e[1])
        if d_done[ci % 3] is not None:
            s_k.wait_event(d_done[ci % 3])
        e[2].record(s_k)
        if kernel:
            ik.solve_into(b["p"][:m], _abi.POSE_EULER6, None, None, None, b["s"], None, b["j"], None, stream=s_k.cuda_stream)
        e[3].record(s_k)
        k_done[ci % 3] = e[3]
        s_out.wait_event(e[3])
        with torch.cuda.stream(s_out):
            e[4].record(s_out)
            h_j[lo:hi].copy_(b["j"][:m], non_blocking=True)
            if state_copy:
                h_s[lo:hi].copy_(b["s"][:m], non_blocking=True)
            e[5].record(s_out)
        d_done[ci % 3] = e[5]
        evs.append(e)
    torch.cuda.synchronize()
    if timeline:
        for ci, e in enumerate(evs):
            print("   chunk", ci, " ".join(f"{t_origin.elapsed_time(x):7.3f}" for x in e), "(H2D start/end, K start/end, D2H start/end, ms)")


for kernel, state_copy in ((False, False), (True, False), (True, True)):
    run(kernel, state_copy); run(kernel, state_copy)
    t0 = time.perf_counter()
    for _ in range(8):
        run(kernel, state_copy)
    dt = (time.perf_counter() - t0) / 8
    print(f"kernel={kernel} state_copy={state_copy}: {dt * 1e3:.2f} ms / 1M -> {n / dt:.3e} poses/s", flush=True)
    run(kernel, state_copy, timeline=True)
