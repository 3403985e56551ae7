"""Lean / fat e2e of SymbolicIK.is_reachable_batch_host at N ranks (torchrun) for several chunk sizes, beside plain chunked
copies of the same byte counts: where does the N-rank e2e go?"""
import json
import os
import sys
import time

import numpy as np
import torch

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
import torch.distributed as dist

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import SymbolicIK, fk  # noqa: E402
from scipy.spatial.transform import Rotation as R  # noqa: E402

n = 1_000_000
M = fk.sample_fk_poses(n, "r_arm", seed=1 + rank)
ik = SymbolicIK(arm="r_arm", device=local)
mat = torch.from_numpy(M.reshape(n, 16)).pin_memory()
gp = torch.from_numpy(np.ascontiguousarray(np.concatenate([M[:, :3, 3], R.from_matrix(M[:, :3, :3]).as_euler("xyz")], axis=1))).pin_memory()


def agg(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return float(t)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
for name, inp, want in (("lean", gp, SymbolicIK.LEAN), ("fat", mat, None)):
    out = ik.alloc_host_outputs(n, want=want)
    for chunk in (1 << 16, 1 << 18, 1 << 20):
        for _ in range(2):
            ik.is_reachable_batch_host(inp, out, chunk=chunk, want=want)
        barrier()
        t0 = time.perf_counter()
        for _ in range(6):
            ik.is_reachable_batch_host(inp, out, chunk=chunk, want=want)
        dt = (time.perf_counter() - t0) / 6
        barrier()
        res[f"{name}_chunk{chunk}"] = agg(n / dt)
# plain chunked copies of the lean byte counts (48 B in, 56 B out per pose), no kernel
d_in = torch.empty(n * 48, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n * 56, dtype=torch.uint8, device="cuda")
h_in = torch.empty(n * 48, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n * 56, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for chunk in (1 << 16, 1 << 18, 1 << 20):
    def once():
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            with torch.cuda.stream(s1):
                d_in[lo * 48:hi * 48].copy_(h_in[lo * 48:hi * 48], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[lo * 56:hi * 56].copy_(d_out[lo * 56:hi * 56], non_blocking=True)
        torch.cuda.synchronize()
    once(); barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        once()
    dt = (time.perf_counter() - t0) / 6
    barrier()
    res[f"plain_lean_bytes_chunk{chunk}"] = agg(n / dt)
if rank == 0:
    print(json.dumps({"n_ranks": world, "cpus": os.cpu_count(), "aggregate_poses_per_s": res}), flush=True)
if world > 1:
    dist.destroy_process_group()
