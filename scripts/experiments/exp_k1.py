"""K1 tuning experiments on the GPU box (development aid; bench.py is the contract).

    python scripts/exp_k1.py time            # this process: time K1 of $R2IK_LIB (or the in-tree lib) + parity vs oracle
    python scripts/exp_k1.py sweep           # every lib in lib/variants + the in-tree lib, one subprocess each
    python scripts/exp_k1.py zerocopy        # K1 reading / writing pinned host memory directly over PCIe
"""
import glob
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def setup(n):
    import torch
    from reachy2_symbolic_ik_b200 import SymbolicIK, fk

    M = fk.sample_fk_poses(n, "r_arm", seed=1)
    ik = SymbolicIK(arm="r_arm")
    Md = torch.from_numpy(M).cuda().reshape(n, 16)
    outs = dict(reach=torch.empty(n, dtype=torch.uint8, device="cuda"), state=torch.empty(n, dtype=torch.uint8, device="cuda"),
                itv=torch.empty((n, 2), dtype=torch.float64, device="cuda"), j=torch.empty((n, 7), dtype=torch.float64, device="cuda"),
                e=torch.empty((n, 3), dtype=torch.float64, device="cuda"))
    return torch, ik, M, Md, outs


def time_launch(torch, fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cmd_time():
    n = 1_000_000
    torch, ik, M, Md, o = setup(n)
    ms = time_launch(torch, lambda: ik.solve_into(Md, 1, None, None, o["reach"], o["state"], o["itv"], o["j"], o["e"]))
    from oracle import oracle as O
    m = 50_000
    want = O.symik_batch(O.arm_config("r_arm"), M[:m])
    ej = np.abs(o["j"][:m].cpu().numpy() - want[3]); ei = np.abs(o["itv"][:m].cpu().numpy() - want[1])
    mism = int((o["state"][:m].cpu().numpy() != want[2]).sum())
    print(f"{os.environ.get('R2IK_LIB', 'in-tree'):60s} K1 {ms * 1e3:8.1f} us/1M  {n / ms * 1e3:.3e} poses/s  "
          f"state_mismatch={mism} max_dj={np.nanmax(ej):.2e} max_di={np.nanmax(ei):.2e}", flush=True)
    # K2 (ControlIK discrete, K = 360) and K1-f32 of the same library
    from reachy2_symbolic_ik_b200 import ControlIK, fk
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    ctl.nb_search_points = 360
    M3 = torch.from_numpy(fk.sample_fk_poses(n, "r_arm", seed=3)).cuda().reshape(n, 16)
    out = [None]
    def k2():
        out[0] = ctl.symbolic_inverse_kinematics_batch("r_arm", M3, "discrete", out=out[0])
    ms2 = time_launch(torch, k2, reps=10, warm=3)
    P32 = Md.float()
    o32 = dict(itv=torch.empty((n, 2), device="cuda"), j=torch.empty((n, 7), device="cuda"), e=torch.empty((n, 3), device="cuda"),
               ne=torch.empty(1, dtype=torch.int32, device="cuda"), sc=torch.empty(n, dtype=torch.int32, device="cuda"))
    ms3 = time_launch(torch, lambda: ik.solve_into_f32(P32, 1, None, None, o["reach"], o["state"], o32["itv"], o32["j"], o32["e"],
                                                       o32["ne"], scratch=o32["sc"]))
    print(f"{'':60s} K2 {ms2 * 1e3:8.1f} us/1M (K=360)   K1-f32 {ms3 * 1e3:8.1f} us/1M", flush=True)


def cmd_sweep():
    libs = [None] + sorted(glob.glob(os.path.join(REPO, "reachy2_symbolic_ik_b200", "lib", "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["R2IK_LIB"] = lib
        for generic in (("0", "1") if lib is None else ("1",)):
            env["R2IK_K1_GENERIC"] = generic
            print("generic" if generic == "1" else "stream ", end=" ", flush=True)
            subprocess.run([sys.executable, __file__, "time"], env=env)


def cmd_zerocopy():
    n = 1_000_000
    torch, ik, M, Md, o = setup(n)
    hin = torch.from_numpy(M).reshape(n, 16).pin_memory()
    ho = dict(reach=torch.empty(n, dtype=torch.uint8).pin_memory(), state=torch.empty(n, dtype=torch.uint8).pin_memory(),
              itv=torch.empty((n, 2), dtype=torch.float64).pin_memory(), j=torch.empty((n, 7), dtype=torch.float64).pin_memory(),
              e=torch.empty((n, 3), dtype=torch.float64).pin_memory())
    cases = {"dev->dev": (Md, o), "host->dev": (hin, o), "dev->host": (Md, ho), "host->host": (hin, ho)}
    for name, (pi, po) in cases.items():
        ms = time_launch(torch, lambda: ik.solve_into(pi, 1, None, None, po["reach"], po["state"], po["itv"], po["j"], po["e"]), reps=5, warm=2)
        print(f"zero-copy {name:12s}: {ms:8.3f} ms/1M poses -> {n / ms * 1e3:.3e} poses/s ({n * 128 / ms / 1e6:.1f} GB/s in, {n * 98 / ms / 1e6:.1f} GB/s out)", flush=True)
    assert torch.equal(ho["j"].nan_to_num(), o["j"].cpu().nan_to_num())
    # staged pipeline for comparison
    hout = ik.alloc_host_outputs(n)
    for chunk in (1 << 16, 1 << 17, 1 << 18):
        t = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ik.is_reachable_batch_host(hin, hout, chunk=chunk)
            t.append(time.perf_counter() - t0)
        print(f"staged pipeline chunk={chunk}: {min(t) * 1e3:.3f} ms/1M -> {n / min(t):.3e} poses/s", flush=True)


if __name__ == "__main__":
    {"time": cmd_time, "sweep": cmd_sweep, "zerocopy": cmd_zerocopy}[sys.argv[1]]()
