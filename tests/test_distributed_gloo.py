"""World-size-2 CPU tests (gloo) of the multi-GPU host logic: contiguous pose slices need no
exchange, and the reach map's orientation shards sum to the full map through one all-reduce.
The per-rank compute is played by the CPU oracle here (the product has no CPU path); on the GPU
box the same ``workspace.sharded_sum`` wraps the K4 launch."""
import os
import socket

import numpy as np
import pytest

from parity import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    import sys

    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    from oracle import oracle as O
    from reachy2_symbolic_ik_b200 import fk, workspace

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = O.arm_config("r_arm")
        ori = fk.fibonacci_orientations(13)
        origin, step, dims = workspace.reach_grid([0.0, -0.2, 0.0], 0.66, 12)

        def launch(b, e):
            return torch.from_numpy(O.reach_map(cfg, origin, step, dims, ori, b, e).astype(np.int32))

        counts = workspace.sharded_sum(launch, len(ori), dist)
        full = O.reach_map(cfg, origin, step, dims, ori).astype(np.int32)
        assert np.array_equal(counts.numpy(), full), "all-reduced shards differ from the full map"

        # the sharded map's wire format: 16-bit counts, two per int32 lane, live x-range only, slab by slab (async)
        from types import SimpleNamespace

        solver = SimpleNamespace(shoulder_position=[0.0, -0.2, 0.0], max_arm_length=0.66, backward_limit=0.02)
        lo, hi = workspace.live_x_range(solver, origin, step, dims)
        assert not full[:lo].any() and not full[hi:].any() and 0 < hi - lo < int(dims[0])
        b, e = workspace.shard_range(len(ori), rank, world)
        mine16 = torch.from_numpy(O.reach_map(cfg, origin, step, dims, ori, b, e).astype(np.int16).reshape(-1))
        plane = int(dims[1] * dims[2])
        edges = [lo + (hi - lo) * k // 3 for k in range(4)]
        works = [workspace.allreduce_u16_pairs(mine16, x0 * plane, x1 * plane, dist) for x0, x1 in zip(edges[:-1], edges[1:])]
        for w in works:
            w.wait()
        assert np.array_equal(mine16.numpy().astype(np.int32).reshape(full.shape), full), "packed 16-bit all-reduce differs"
        big = torch.full((6,), 30000 if rank == 0 else 30001, dtype=torch.int16)      # 60 001 > int16 max: the top bit is a count bit
        workspace.allreduce_u16_pairs(big, 0, 5, dist).wait()                          # odd tail: the padding element rides along
        assert np.array_equal(big.numpy().view(np.uint16), np.full(6, 60001, np.uint16))

        # pose batch: contiguous slices, no exchange; gathering the slices reproduces the full batch
        M = fk.sample_fk_poses(1001, "r_arm", seed=9)
        b, e = workspace.shard_range(len(M), rank, world)
        mine = O.symik_batch(cfg, M[b:e])[3]
        gathered = [None] * world
        dist.all_gather_object(gathered, (b, e, mine))
        whole = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        want = O.symik_batch(cfg, M)[3]
        assert np.array_equal(np.nan_to_num(whole), np.nan_to_num(want))
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from reachy2_symbolic_ik_b200.workspace import shard_range

    for n in (0, 1, 7, 512, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_world_size_2_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
