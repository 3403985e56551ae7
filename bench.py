#!/usr/bin/env python
"""bench.py -- headline benchmark of the Reachy2 symbolic IK hot path on B200.

Workload (BASELINE.json configs[1]): 1 M FK-sampled poses per arm (r_arm and l_arm), FP64,
reachability flag + theta interval + 7 joints at theta_interval[0] + elbow position.  One *step*
= one pass of K1 (`r2ik_symik_solve_f64`) over both arms' batches = 2 M solves per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* own arm: `value` = poses/s with inputs resident in HBM (CUDA events on the launching stream,
  max over ranks); `e2e` = the same metric through the public facade with pinned HOST buffers,
  H2D and D2H copies inside the timed region; `roofline` / `roofline_fp64` for the K1 kernel;
  `cpu_baseline` = the C oracle (a port of the reference algorithm) on the box's host cores.
* `--impl reference`: the reference algorithm's CPU implementation (oracle port, OpenMP over all
  host threads) on the same workload; the pure-Python reference itself cannot travel to the GPU
  box (its measured speed in the build container is recorded in BASELINE.md / DESIGN.md).
* N > 1: launched by torchrun, one rank per GPU, contiguous slices of the pose batch per rank,
  no data-path collective (weak scaling: 2 M solves per GPU per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

POSES_PER_ARM = 1_000_000
ARMS = ("r_arm", "l_arm")
SEEDS = {"r_arm": 1, "l_arm": 2}
BYTES_IN = 128                              # row-major 4x4 float64 per pose
BYTES_OUT = 1 + 1 + 16 + 56 + 24            # reachable, state, interval, joints, elbow
FLOP_EQ_PER_SOLVE = 2100.0                  # weighted FP64 flop-equivalents
METRIC = "ik_poses_per_sec"
UNIT = "poses/s"


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for k, nme in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(rank: int):
    """Synthetic FK-sampled poses, distinct per rank (contiguous slices of the global batch)."""
    from reachy2_symbolic_ik_b200 import fk

    return {arm: fk.sample_fk_poses(POSES_PER_ARM, arm, seed=SEEDS[arm] + 1000 * rank) for arm in ARMS}


def cpu_baseline(poses, repeats: int = 3):
    """The oracle port on the host cores (all OpenMP threads) over the full per-GPU workload."""
    from oracle import oracle as O

    cfgs = {arm: O.arm_config(arm) for arm in ARMS}
    O.symik_batch(cfgs["r_arm"], poses["r_arm"][:20000])  # warm-up (thread pool, page-in)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for arm in ARMS:
            O.symik_batch(cfgs[arm], poses[arm])
        best = min(best, time.perf_counter() - t0)
    n = sum(len(poses[a]) for a in ARMS)
    return {"value": n / best, "unit": UNIT, "cores": O.max_threads(), "kind": "port",
            "sample": f"full per-GPU workload ({n} poses: {POSES_PER_ARM} per arm), best of {repeats}, C oracle + OpenMP"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O

    poses = make_workload(0)
    cfgs = {arm: O.arm_config(arm) for arm in ARMS}
    n = sum(len(poses[a]) for a in ARMS)
    for _ in range(args.warmup):
        for arm in ARMS:
            O.symik_batch(cfgs[arm], poses[arm])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for arm in ARMS:
            O.symik_batch(cfgs[arm], poses[arm])
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.max_threads(), "kind": "port",
                         "sample": f"{n} poses per step ({POSES_PER_ARM} per arm), C oracle port of the reference algorithm, OpenMP"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python (NumPy/SciPy) and is not installable on the GPU box; this arm times the "
                "C port of its algorithm (oracle/) on all host threads. The Python reference itself measured "
                "~590 poses/s/core in the build container (BASELINE.md).",
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus: int):
    return {"workload": "configs[1]: 1M FK-sampled poses per arm (r_arm + l_arm), FP64, flag + theta interval + joints at "
                        "theta_interval[0] + elbow, (N,4,4) pose input", "poses_per_step_per_gpu": 2 * POSES_PER_ARM,
            "global_poses_per_step": 2 * POSES_PER_ARM * n_gpus, "pose_layout": "mat4_rowmajor_f64",
            "l2_policy": "inputs+outputs per step (452 MB) exceed the 126 MB L2; the two arms' batches alternate",
            "parallelism": f"pose-slices x{n_gpus}, no collective"}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        # one reference step is ~0.25 s of all host cores: bound the run to a few minutes
        args.steps = min(args.steps, 20)
        args.warmup = min(args.warmup, 2)
        return run_reference(args)

    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from reachy2_symbolic_ik_b200 import SymbolicIK, _abi, _native

    dev = torch.device("cuda", local_rank)
    poses = make_workload(rank)
    solvers = {arm: SymbolicIK(arm=arm, device=local_rank) for arm in ARMS}
    n = POSES_PER_ARM
    dpose = {arm: torch.from_numpy(poses[arm]).reshape(n, 16).to(dev) for arm in ARMS}
    outs = {arm: dict(reach=torch.empty(n, dtype=torch.uint8, device=dev), state=torch.empty(n, dtype=torch.uint8, device=dev),
                      interval=torch.empty((n, 2), dtype=torch.float64, device=dev),
                      joints=torch.empty((n, 7), dtype=torch.float64, device=dev),
                      elbow=torch.empty((n, 3), dtype=torch.float64, device=dev)) for arm in ARMS}

    def step():
        for arm in ARMS:
            o = outs[arm]
            solvers[arm].solve_into(dpose[arm], _abi.POSE_MAT4, None, None, o["reach"], o["state"], o["interval"], o["joints"], o["elbow"])
        return 2  # kernel launches

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            launches += step()
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    poses_per_step = 2 * n * world
    value = poses_per_step * args.steps / (ms_total * 1e-3)
    kernel_ms = ms_total / launches  # the step is nothing but K1 launches: average launch duration

    # ---- end to end through the public facade with pinned host buffers (copies inside the timed region)
    host_in = {arm: torch.from_numpy(poses[arm]).reshape(n, 16).pin_memory() for arm in ARMS}
    host_out = {arm: solvers[arm].alloc_host_outputs(n) for arm in ARMS}
    for _ in range(2):
        for arm in ARMS:
            solvers[arm].is_reachable_batch_host(host_in[arm], host_out[arm])
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for arm in ARMS:
            solvers[arm].is_reachable_batch_host(host_in[arm], host_out[arm])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = poses_per_step * e2e_steps / e2e_s
    # the e2e results must be the device-resident results
    chk = solvers["r_arm"].is_reachable_batch_host(host_in["r_arm"], host_out["r_arm"])
    assert torch.equal(chk.joints[:1000].nan_to_num(), outs["r_arm"]["joints"][:1000].cpu().nan_to_num())

    # ---- parity spot check against the oracle on this rank's data (not timed)
    parity = None
    cpu = None
    if rank == 0:
        from oracle import oracle as O

        m = 100_000
        want = O.symik_batch(O.arm_config("r_arm"), poses["r_arm"][:m])
        got_j = outs["r_arm"]["joints"][:m].cpu().numpy()
        got_i = outs["r_arm"]["interval"][:m].cpu().numpy()
        got_s = outs["r_arm"]["state"][:m].cpu().numpy()
        ej = np.abs(got_j - want[3]); ei = np.abs(got_i - want[1])
        parity = {"checked_poses": m, "state_mismatches": int((got_s != want[2]).sum()),
                  "max_abs_err_joints_rad": float(np.nanmax(ej)), "p99.9_abs_err_joints_rad": float(np.nanquantile(ej, 0.999)),
                  "max_abs_err_interval_rad": float(np.nanmax(ei)), "over_1e-9": int((np.nan_to_num(ej).max(axis=1) > 1e-9).sum()),
                  "vs": "CPU oracle (pinned to the reference by tests/golden)"}
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(poses)

    # ---- rooflines for the K1 kernel
    hbm_peak, peak_src = measured_peaks()
    solves_per_launch = n
    alg_bytes = (BYTES_IN + BYTES_OUT) * solves_per_launch
    achieved_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
    import ctypes as C

    pms, pfl = C.c_double(), C.c_double()
    _native.check(_native.load().r2ik_dfma_probe(local_rank, 400000, C.byref(pms), C.byref(pfl), None), "r2ik_dfma_probe")
    fp64_peak = pfl.value / (pms.value * 1e-3) / 1e12
    fp64_ach = FLOP_EQ_PER_SOLVE * solves_per_launch / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(REPO, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * BYTES_IN, "d2h_bytes_per_step": 2 * n * BYTES_OUT,
                    "steps": e2e_steps, "path": "SymbolicIK.is_reachable_batch_host: pinned host -> chunked H2D / K1 / D2H on 3 streams"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": traffic, "kernel": "k_symik_solve<MAT4>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_solve": BYTES_IN + BYTES_OUT, "solves_per_launch": solves_per_launch, "peak_source": peak_src,
                         "note": "K1 is FP64-pipe bound, not HBM bound: see roofline_fp64"},
            "roofline_fp64": {"bound": "fp64", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak,
                              "flop_eq_per_solve": FLOP_EQ_PER_SOLVE, "peak_source": "r2ik_dfma_probe measured in this run (DFMA chains, full grid)"},
            "clocks": clocks.summary(), "parity": parity,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
