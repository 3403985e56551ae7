"""Plain pinned-memory copy ceiling at N ranks of one node: what H2D, D2H and both at once deliver per rank and in
aggregate when 1 / 2 / 4 / 8 processes drive their own GPU at the same time.  This is the bound of every e2e number of
bench.py (the kernels are 10-40x faster than the link).  Launch like bench.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/experiments/exp_pcie_nrank.py            (or plain `python ...` for N = 1)

Rank 0 prints one JSON line: per-direction GB/s (min over ranks and aggregate) and the poses/s each record format could
reach on that link (fat: (N,4,4) in + all outputs = 128 + 98 B; lean: (N,6) in + state + joints = 48 + 57 B)."""
import json
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from reachy2_symbolic_ik_b200 import hostmem  # noqa: E402

binding = hostmem.bind_to_gpu_numa(local) if "--bind" in sys.argv else {"bound": False}
MB = 256
n = MB << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def run(mode, reps=8):
    def once():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = n * reps / dt / 1e9           # per direction
    if dist is not None:
        t = torch.tensor([gbs, -gbs, gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t[:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(t[1:2], op=dist.ReduceOp.MAX)
        return {"aggregate_GBps_per_direction": float(t[0]), "slowest_rank_GBps": float(-t[1])}
    return {"aggregate_GBps_per_direction": gbs, "slowest_rank_GBps": gbs}


res = {m: run(m) for m in ("h2d", "d2h", "both")}
if rank == 0:
    both = res["both"]["aggregate_GBps_per_direction"] * 1e9
    line = {"n_ranks": world, "copy_MB": MB, "binding": binding, **res,
            "poses_per_s_bound": {"fat_128in_98out": both / 128.0, "goal_pose_48in_98out": both / 98.0, "lean_48in_57out": both / 57.0},
            "note": "both = H2D and D2H streams running at the same time; the poses/s bound divides the per-direction rate by the "
                    "larger of the two byte counts of a record"}
    print(json.dumps(line), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
