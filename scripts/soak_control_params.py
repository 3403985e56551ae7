"""Randomised soak of ControlIK's per-call parameters against the unmodified reference (build container only: imports
/root/reference/src): discrete mode with random K / preferred_theta / constrained_mode / DVT, continuous mode with random
d_theta_max / preferred_theta / constrained_mode / DVT incl. excursions out of the workspace.  Checked: the CPU oracle,
the kernel source compiled for the host (the serial recursion and the phased form with its finish pass on winding codes).

    PYTHONDONTWRITEBYTECODE=1 python scripts/soak_control_params.py [seed] [trials]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "tests", "golden")]
sys.dont_write_bytecode = True

import gen_golden as G  # noqa: E402
import test_hostsim_parity as T  # noqa: E402
from oracle import oracle as O  # noqa: E402
from parity import Report, ill_conditioned_mask  # noqa: E402
from reachy2_symbolic_ik_b200 import fk  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 99
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 14
O.build()
subprocess.run(["make", "-C", T.HS_DIR], check=True, capture_output=True)
hs = C.CDLL(os.path.join(T.HS_DIR, "_build", "libr2ik_hostsim.so"))
params = {k: np.asarray(v) for k, v in G.urdf_params().items()}
rng = np.random.default_rng(seed)
bad = 0

# ---------------------------------------------------------------- discrete mode
for trial in range(trials):
    arm = ("r_arm", "l_arm")[trial % 2]
    K = int(rng.choice([2, 3, 5, 20, 51, 100]))
    pref = float(rng.uniform(-np.pi, np.pi)) if trial % 3 else float(rng.uniform(-7, 7))
    mode = ("unconstrained", "low_elbow")[int(rng.integers(2))]
    dvt = trial % 5 == 4
    n = 300
    M = np.ascontiguousarray(np.concatenate([fk.sample_fk_poses(n - 60, arm, seed=seed + 500 + trial, min_x=0.05),
                                             fk.sample_task_space_poses(60, arm, seed=seed + 600 + trial)]))
    ctl = G.new_control(is_dvt=dvt)
    ctl.nb_search_points = K
    J = np.zeros((n, 7)); F = np.zeros(n, bool); S = np.zeros(n, np.uint8)
    with G._Quiet():
        for i in range(n):
            j, ok, st = ctl.symbolic_inverse_kinematics(arm, M[i], "discrete", constrained_mode=mode, preferred_theta=pref)
            J[i], F[i], S[i] = j, ok, G.state_code(st)
    off = 0.03 if dvt else -1.01
    kw = dict(nb_search_points=K, preferred_theta=pref, constrained_mode=mode)
    ocfg = O.arm_config(arm, ik_parameters=params, singularity_offset=off)
    opar = O.ControlParams(arm=arm, **kw)
    ill = ill_conditioned_mask(lambda p: O.ctl_discrete_batch(ocfg, opar, p.reshape(M.shape))[:3], M.reshape(n, -1))
    j, r, st, _ = O.ctl_discrete_batch(ocfg, opar, M)
    reps = [Report(f"oracle discrete trial {trial}", n, ill)]
    reps[0].exact("reach", r, F); reps[0].exact("state", st, S); reps[0].close("joints", j, J)
    cfg, par = T.cfg_for(arm, params, off), T.ctl_params(O, arm, **kw)
    prev = np.array(O.DEFAULT_PREV_JOINTS[arm])
    hj = np.empty((n, 7)); hr = np.zeros(n, np.uint8); hst = np.zeros(n, np.uint8); he = np.zeros(n, np.uint8)
    hs.hs_ctl_discrete_batch(C.byref(cfg), C.byref(par), T.dp(M), C.c_int64(n), T.dp(prev), T.dp(prev), T.dp(hj), T.u8(hr),
                             T.u8(hst), T.u8(he))
    reps.append(Report(f"hostsim discrete trial {trial}", n, ill))
    reps[1].exact("reach", hr.astype(bool), F); reps[1].exact("state", hst, S); reps[1].close("joints", hj, J)
    ok = not any(rp.bad.any() for rp in reps)
    for rp in reps:
        if rp.bad.any():
            print(rp.summary())
    bad += not ok
    print(f"discrete trial {trial} {arm} K={K} pref={pref:.3f} {mode} dvt={dvt}: found {F.mean():.2f}, limited by shoulder "
          f"{(S == 6).mean():.2f}, ill {int(ill.sum())} -> {'ok' if ok else 'BAD'}")

# ---------------------------------------------------------------- continuous mode
G.ref_control.time = G.FakeTime()
for trial in range(trials):
    arm = ("r_arm", "l_arm")[trial % 2]
    dth = float(10 ** rng.uniform(-3.3, -0.5))
    pref = float(rng.uniform(-np.pi, np.pi))
    mode = ("unconstrained", "low_elbow")[int(rng.integers(2))]
    dvt = trial % 4 == 3
    Tn, W = 3, 150
    Ms = fk.sinusoidal_trajectories(Tn, W, arm, seed=seed + 700 + trial)[0].copy()
    if trial % 3 == 0:   # an excursion out of the workspace in the middle of trajectory 1
        Ms[1, :, 0, 3] += 0.5 * np.sin(np.pi * np.linspace(0, 1, W)) ** 2
    J = np.zeros((Tn, W, 7)); F = np.zeros((Tn, W), bool); S = np.zeros((Tn, W), np.uint8); E = np.zeros(Tn, bool); TH = np.zeros(Tn)
    kw = dict(preferred_theta=pref, constrained_mode=mode, d_theta_max=dth)
    for t in range(Tn):
        ctl = G.new_control(is_dvt=dvt)
        with G._Quiet():
            for w in range(W):
                j, ok, st = ctl.symbolic_inverse_kinematics(arm, Ms[t, w], "continuous", **kw)
                J[t, w], F[t, w], S[t, w] = j, ok, G.state_code(st)
        E[t], TH[t] = ctl.emergency_stop, ctl.previous_theta[arm]
    off = 0.03 if dvt else -1.01
    ocfg = O.arm_config(arm, ik_parameters=params, singularity_offset=off)
    cfg, par = T.cfg_for(arm, params, off), T.ctl_params(O, arm, **kw)
    runs = {"oracle": O.ctl_continuous_batch(ocfg, O.ControlParams(arm=arm, **kw), Ms),
            "hostsim serial": T.hs_continuous(hs, O, cfg, par, arm, Ms),
            "hostsim phased": T.hs_continuous(hs, O, cfg, par, arm, Ms, phased=True)}
    ok = True
    for who, (jj, rr, ss, stt) in runs.items():
        for t in range(Tn):
            rep = Report(f"{who} continuous trial {trial} traj {t}", W, np.abs(J[t, :, 3]) < 1e-3)   # straight-arm singularity
            rep.exact("reach", rr[t], F[t]); rep.exact("state", ss[t], S[t]); rep.close("joints", jj[t], J[t])
            if rep.bad.any():
                ok = False
                print(rep.summary())
        if not (np.array_equal(stt["emergency_stop"].astype(bool), E) and np.allclose(stt["previous_theta"], TH, atol=1e-9)):
            ok = False
            print(who, "final controller state differs")
    bad += not ok
    print(f"continuous trial {trial} {arm} d_theta_max={dth:.4f} pref={pref:.3f} {mode} dvt={dvt}: reachable {F.mean():.2f}, "
          f"emergency {E.tolist()} -> {'ok' if ok else 'BAD'}")
print(f"soak_control_params seed {seed}: {'OK' if not bad else str(bad) + ' BAD TRIALS'}")
sys.exit(1 if bad else 0)
