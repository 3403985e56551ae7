"""End-to-end pipeline sweep (development aid): chunk size x stream count of SymbolicIK.is_reachable_batch_host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reachy2_symbolic_ik_b200 import SymbolicIK, fk

n = 1_000_000
ik = SymbolicIK(arm="r_arm")
M = fk.sample_fk_poses(n, "r_arm", seed=1)
for prec, dt in (("fp64", torch.float64), ("fp32", torch.float32)):
    hin = torch.from_numpy(M).reshape(n, 16).to(dt).pin_memory()
    hout = ik.alloc_host_outputs(n, prec)
    for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18):
        for ns in (2, 3, 4, 6):
            t = []
            for _ in range(5):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                ik.is_reachable_batch_host(hin, hout, chunk=chunk, n_streams=ns, precision=prec)
                t.append(time.perf_counter() - t0)
            print(f"{prec} chunk={chunk:7d} streams={ns}: {min(t) * 1e3:7.3f} ms/1M -> {n / min(t):.3e} poses/s", flush=True)
