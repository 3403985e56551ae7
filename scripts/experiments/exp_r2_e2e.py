"""e2e throughput of SymbolicIK.is_reachable_batch_host (native pipeline) for record formats x chunk sizes x slots.
    python scripts/experiments/exp_r2_e2e.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import SymbolicIK, fk  # noqa: E402
from scipy.spatial.transform import Rotation as R  # noqa: E402

n = 1_000_000
M = fk.sample_fk_poses(n, "r_arm", seed=1)
ik = SymbolicIK(arm="r_arm")
mat = torch.from_numpy(M.reshape(n, 16)).pin_memory()
gp = torch.from_numpy(np.ascontiguousarray(np.concatenate([M[:, :3, 3], R.from_matrix(M[:, :3, :3]).as_euler("xyz")], axis=1))).pin_memory()
ref = ik.is_reachable_batch(M)
for name, inp, want, bts in (("fat  mat4 + all outputs (128 in / 98 out)", mat, None, (128, 98)),
                             ("goal pose + all outputs (48 in / 98 out)", gp, None, (48, 98)),
                             ("lean goal pose + state + joints (48 in / 57 out)", gp, SymbolicIK.LEAN, (48, 57))):
    out = ik.alloc_host_outputs(n, want=want)
    for chunk in (1 << 16, 1 << 17, 1 << 18, 1 << 19):
        for slots in (3,):
            for _ in range(2):
                ik.is_reachable_batch_host(inp, out, chunk=chunk, n_streams=slots, want=want)
            t0 = time.perf_counter()
            reps = 8
            for _ in range(reps):
                ik.is_reachable_batch_host(inp, out, chunk=chunk, n_streams=slots, want=want)
            dt = (time.perf_counter() - t0) / reps
            ok = np.array_equal(out.state.numpy(), ref.state)
            print(f"{name:52s} chunk {chunk:7d} slots {slots}: {n / dt:.3e} poses/s  ({dt * 1e3:.2f} ms / 1M; in {bts[0] * n / dt / 1e9:.1f} GB/s out {bts[1] * n / dt / 1e9:.1f} GB/s) states ok {ok}", flush=True)

# both arms enqueued before either is waited for (what bench.py's e2e step does)
ik2 = SymbolicIK(arm="l_arm")
M2 = fk.sample_fk_poses(n, "l_arm", seed=2)
gp2 = torch.from_numpy(np.ascontiguousarray(np.concatenate([M2[:, :3, 3], R.from_matrix(M2[:, :3, :3]).as_euler("xyz")], axis=1))).pin_memory()
mat2 = torch.from_numpy(M2.reshape(n, 16)).pin_memory()
for name, a, b, want in (("fat two arms async", mat, mat2, None), ("lean two arms async", gp, gp2, SymbolicIK.LEAN)):
    o1, o2 = ik.alloc_host_outputs(n, want=want), ik2.alloc_host_outputs(n, want=want)
    for chunk in (1 << 17, 1 << 18):
        def step():
            ik.is_reachable_batch_host(a, o1, chunk=chunk, want=want, wait=False)
            ik2.is_reachable_batch_host(b, o2, chunk=chunk, want=want, wait=False)
            ik.wait_host(); ik2.wait_host()
        step(); step()
        t0 = time.perf_counter()
        for _ in range(8):
            step()
        dt = (time.perf_counter() - t0) / 8
        print(f"{name:52s} chunk {chunk:7d}: {2 * n / dt:.3e} poses/s ({dt * 1e3:.2f} ms / 2M)", flush=True)
