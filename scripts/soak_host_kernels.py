"""Large-sample soak of the kernel SOURCE compiled for the host (tests/hostsim) against the CPU oracle -- no reference and no
GPU needed.  K1 FP64 and FP32 on 1 M FK + 1 M task-space poses per arm, K2 on 500 k poses per arm (K = 360 and low_elbow),
K3 (serial, and phased with its finish pass on winding codes) on 3 000 trajectories x 200 waypoints per arm with excursions
and orientation flips.  The committed tests run thousands of poses; this runs millions.

    python scripts/soak_host_kernels.py [seed]
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]

import test_hostsim_parity as T  # noqa: E402
from oracle import oracle as O  # noqa: E402
from parity import ill_conditioned_mask  # noqa: E402
from reachy2_symbolic_ik_b200 import fk  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 901
O.build()
O.use_all_host_threads()
subprocess.run(["make", "-C", T.HS_DIR], check=True, capture_output=True)
hs = C.CDLL(os.path.join(T.HS_DIR, "_build", "libr2ik_hostsim.so"))
params = T.urdf_params()
t0 = time.time()
fails = 0


def genuine(over, run, P):
    """Of the poses whose error exceeds the tolerance, those that are NOT ill-conditioned in tests/parity.py's sense (the
    oracle's own output moves by > 1e-10 under a 3e-13 perturbation of the pose): (genuine, ill-conditioned)."""
    idx = np.nonzero(over)[0]
    if not len(idx):
        return 0, 0
    sub = np.ascontiguousarray(P[idx])
    ill = ill_conditioned_mask(lambda p: run(p.reshape(sub.shape)), sub.reshape(len(idx), -1), n_trials=8)
    return int((~ill).sum()), int(ill.sum())


def report(tag, n_bad, extra):
    global fails
    fails += bool(n_bad)
    print(f"{tag}: {'OK' if not n_bad else str(n_bad) + ' MISMATCHES'}; {extra} [{time.time() - t0:.0f}s]", flush=True)


for arm in ("r_arm", "l_arm"):
    # ---- K1, FP64 and FP32 (states identical; FP64 joints / intervals within 1e-9, FP32 within 1e-4 at p99.9)
    for kind, P in (("fk", fk.sample_fk_poses(1_000_000, arm, seed=seed, min_x=None)),
                    ("task", fk.sample_task_space_poses(1_000_000, arm, seed=seed + 1))):
        w = O.symik_batch(O.arm_config(arm), P)
        r, itv, st, j, e = T.hs_symik(hs, T.cfg_for(arm), P)
        ej = np.nan_to_num(np.abs(j - w[3])).max(axis=1)
        ei = np.nan_to_num(np.abs(itv - w[1])).max(axis=1)
        ocfg1 = O.arm_config(arm)
        g, ill = genuine((ej > 1e-9) | (ei > 1e-9), lambda p: O.symik_batch(ocfg1, p)[:4], P)
        report(f"K1 f64 {arm} {kind}", int((st != w[2]).sum()) + g,
               f"max |d joints| {ej.max():.1e}, max |d interval| {ei.max():.1e}, over 1e-9 and ill-conditioned: {ill}")
        P32 = P.astype(np.float32)
        w32 = O.symik_batch(O.arm_config(arm), P32.astype(np.float64))
        r, itv, st, j, e, esc = T.hs_symik_f32(hs, T.cfg_for(arm), P32)
        ej = np.nan_to_num(np.abs(j - w32[3])).max(axis=1)
        report(f"K1 f32 {arm} {kind}", int((st != w32[2]).sum()) + int(np.quantile(ej, 0.999) > 1e-4),
               f"p99.9 |d joints| {np.quantile(ej, 0.999):.1e}, over 1e-4: {int((ej > 1e-4).sum())}, escalated {esc.mean():.2e}")
    # ---- K2
    n = 500_000
    M = np.ascontiguousarray(np.concatenate([fk.sample_fk_poses(n * 4 // 5, arm, seed=seed + 2, min_x=0.0),
                                             fk.sample_task_space_poses(n // 5, arm, seed=seed + 3)]))
    ocfg = O.arm_config(arm, ik_parameters=params, singularity_offset=-1.01)
    cfg = T.cfg_for(arm, params, -1.01)
    prev = np.array(O.DEFAULT_PREV_JOINTS[arm])
    for kw in (dict(nb_search_points=360), dict(nb_search_points=20, constrained_mode="low_elbow")):
        wj, wr, ws, we = O.ctl_discrete_batch(ocfg, O.ControlParams(arm=arm, **kw), M)
        par = T.ctl_params(O, arm, **kw)
        j = np.empty((n, 7)); r = np.zeros(n, np.uint8); st = np.zeros(n, np.uint8); emg = np.zeros(n, np.uint8)
        hs.hs_ctl_discrete_batch(C.byref(cfg), C.byref(par), T.dp(M), C.c_int64(n), T.dp(prev), T.dp(prev), T.dp(j), T.u8(r),
                                 T.u8(st), T.u8(emg))
        ej = np.abs(j - wj).max(axis=1)
        opar = O.ControlParams(arm=arm, **kw)
        g, ill = genuine(ej > 1e-9, lambda p: O.ctl_discrete_batch(ocfg, opar, p)[:3], M)
        report(f"K2 {arm} {kw}", int((st != ws).sum()) + g + int((emg != we).sum()),
               f"max |d joints| {ej.max():.1e}, over 1e-9 and ill-conditioned: {ill}")
    # ---- K3
    Tn, W = 3000, 200
    M = fk.sinusoidal_trajectories(Tn, W, arm, seed=seed + 4)[0].copy()
    rng = np.random.default_rng(seed + 5)
    for t in rng.choice(Tn, Tn // 10, replace=False):
        if t % 2:
            M[t, :, 0, 3] += 0.5 * np.sin(np.pi * np.linspace(0, 1, W)) ** 2          # out of the workspace and back
        else:
            M[t, int(rng.integers(5, W - 5)):, :3, :3] = M[t, -1, :3, :3] @ np.diag([-1.0, -1.0, 1.0])   # jump: emergency latch
    wj, wr, ws, wst = O.ctl_continuous_batch(ocfg, O.ControlParams(arm=arm), M)
    par = T.ctl_params(O, arm)
    for name, kw, sl in (("serial", {}, slice(None)), ("phased", dict(phased=True), slice(None))):
        j, r, s, st = T.hs_continuous(hs, O, cfg, par, arm, M[sl], **kw)
        ej = np.abs(j - wj[sl]).max(axis=2)
        straight = np.abs(wj[sl][..., 3]) < 1e-3          # kinematic singularity: counted, not compared
        n_bad = int((s != ws[sl]).sum()) + int((r != wr[sl]).sum()) + int(((ej > 1e-9) & ~straight).sum()) + \
            int((st["emergency_stop"] != wst["emergency_stop"][sl]).sum())
        report(f"K3 {name} {arm}", n_bad, f"{int(st['emergency_stop'].sum())} latched, {int(straight.sum())} straight-arm waypoints")
print(f"soak_host_kernels seed {seed}: {'OK' if not fails else str(fails) + ' FAILED SECTIONS'}")
sys.exit(1 if fails else 0)
