"""Soak of the CPU oracle AND the kernel source (tests/hostsim) against the unmodified reference on fresh seeds
(build container only: imports /root/reference/src).  Not part of the test suite -- the committed fixtures are a
fixed sample; this draws new ones.

    PYTHONDONTWRITEBYTECODE=1 python scripts/soak_reference.py [seed] [n_poses] [n_traj]
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "tests", "golden")]
sys.dont_write_bytecode = True

import gen_golden as G  # noqa: E402  (the generator's helpers drive the reference)
from oracle import oracle as O  # noqa: E402
from parity import Report, ill_conditioned_mask  # noqa: E402
from reachy2_symbolic_ik_b200 import _abi, fk  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 12345
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 6
O.build()
HS_DIR = os.path.join(REPO, "tests", "hostsim")
subprocess.run(["make", "-C", HS_DIR], check=True, capture_output=True)
hs = C.CDLL(os.path.join(HS_DIR, "_build", "libr2ik_hostsim.so"))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))  # noqa: E731
fails = 0


def close(rep, max_ill=0.03):
    global fails
    print(rep.summary())
    if rep.bad.any() or rep.ill.mean() > max_ill:
        fails += 1
        print("   ^^^ FAIL", np.nonzero(rep.bad)[0][:10])


t0 = time.time()
params = {k: np.asarray(v) for k, v in G.urdf_params().items()}
for ai, arm in enumerate(("r_arm", "l_arm")):
    # ---- SymbolicIK, default constructor: FK-sampled + task-space poses
    M = np.concatenate([fk.sample_fk_poses(n // 2, arm, seed=seed + ai, min_x=None),
                        fk.sample_task_space_poses(n - n // 2, arm, seed=seed + 10 + ai)])
    gp = np.array([G.euler_pose_from_matrix(m) for m in M])
    with G._Quiet():
        ik = G.SymbolicIK(arm=arm)
    flag, state, interval, joints, elbow = G.run_symik(ik, gp)
    ocfg = O.arm_config(arm)
    ill = ill_conditioned_mask(lambda p: O.symik_batch(ocfg, p.reshape(M.shape))[:4], M.reshape(len(M), -1))
    cfg = _abi.make_arm_config(arm, _abi.DEFAULT_IK_PARAMETERS, 127, 42.5, 1e-8, 0.02, 1e-7, 0.03, 1.0)
    for who in ("oracle", "hostsim"):
        if who == "oracle":
            r, itv, st, j, e = O.symik_batch(ocfg, M)
        else:
            Mc = np.ascontiguousarray(M)
            r = np.zeros(len(M), np.uint8); st = np.zeros(len(M), np.uint8)
            itv = np.empty((len(M), 2)); j = np.empty((len(M), 7)); e = np.empty((len(M), 3))
            hs.hs_symik_batch(C.byref(cfg), _abi.POSE_MAT4, dp(Mc), None, C.c_int64(len(M)), u8(r), u8(st), dp(itv), dp(j), dp(e))
            r = r.astype(bool)
        rep = Report(f"soak {who} symik {arm} seed {seed}", len(M), ill)
        rep.exact("reachable", r, flag); rep.exact("state", st, state)
        rep.close("interval", itv, interval); rep.close("joints", j, joints); rep.close("elbow", e, elbow)
        close(rep)
    # ---- ControlIK discrete (K = 20) on the first n/4 poses
    nd = n // 4
    ctl = G.new_control()
    dj, df, ds = G.run_discrete(ctl, arm, M[:nd])
    ocfg2 = O.arm_config(arm, ik_parameters=params, singularity_offset=-1.01)
    opar = O.ControlParams(arm=arm)
    Md = np.ascontiguousarray(M[:nd])
    ill = ill_conditioned_mask(lambda p: O.ctl_discrete_batch(ocfg2, opar, p.reshape(Md.shape))[:3], Md.reshape(nd, -1))
    j, r, st, emg = O.ctl_discrete_batch(ocfg2, opar, Md)
    rep = Report(f"soak oracle discrete {arm} seed {seed}", nd, ill)
    rep.exact("reachable", r, df); rep.exact("state", st, ds); rep.close("joints", j, dj)
    close(rep)
    cfg2 = _abi.make_arm_config(arm, params, 127, 42.5, 1e-8, 0.02, 1e-7, -1.01, 1.0)
    par = _abi.CtlParams()
    C.memmove(C.byref(par), C.byref(opar._c), C.sizeof(par))
    prev = np.array(O.DEFAULT_PREV_JOINTS[arm])
    j = np.empty((nd, 7)); r = np.zeros(nd, np.uint8); st = np.zeros(nd, np.uint8); emg = np.zeros(nd, np.uint8)
    hs.hs_ctl_discrete_batch(C.byref(cfg2), C.byref(par), dp(Md), C.c_int64(nd), dp(prev), dp(prev), dp(j), u8(r), u8(st), u8(emg))
    rep = Report(f"soak hostsim discrete {arm} seed {seed}", nd, ill)
    rep.exact("reachable", r.astype(bool), df); rep.exact("state", st, ds); rep.close("joints", j, dj)
    close(rep)
    # ---- ControlIK continuous on fresh sinusoidal trajectories
    G.ref_control.time = G.FakeTime()
    W = 200
    Ms, q = fk.sinusoidal_trajectories(T, W, arm, seed=seed + 20 + ai)
    J = np.zeros((T, W, 7)); F = np.zeros((T, W), bool); S = np.zeros((T, W), np.uint8); E = np.zeros(T, bool); TH = np.zeros(T)
    for t in range(T):
        J[t], F[t], S[t], E[t], TH[t] = G.run_continuous(arm, Ms[t])
    G.ref_control.time = time
    oj, orr, os_, ost = O.ctl_continuous_batch(ocfg2, opar, Ms)
    cj = np.empty((T, 7)); cp = np.empty((T, 4, 4))
    cj[:] = O.DEFAULT_PREV_JOINTS[arm]; cp[:] = O.DEFAULT_CURRENT_POSE[arm]
    hst = np.zeros(T, dtype=_abi.TRAJ_STATE_DTYPE); hst["init"] = 1
    hj = np.empty((T, W, 7)); hr = np.zeros((T, W), np.uint8); hss = np.zeros((T, W), np.uint8)
    Mc = np.ascontiguousarray(Ms)
    hs.hs_ctl_continuous_batch(C.byref(cfg2), C.byref(par), dp(Mc), C.c_int64(T), C.c_int32(W), dp(cj), dp(cp),
                               hst.ctypes.data_as(C.c_void_p), dp(hj), u8(hr), u8(hss))
    for who, (jj, rr, ss, stt) in (("oracle", (oj, orr, os_, ost)), ("hostsim", (hj, hr.astype(bool), hss, hst))):
        for t in range(T):
            # a nearly straight arm (|elbow pitch| < 1e-3 rad) is a kinematic singularity: the reference's own elbow / wrist
            # yaw there moves by ~1e-6 rad under a 3e-13 perturbation of the pose; such waypoints are counted, not compared
            rep = Report(f"soak {who} continuous {arm} seed {seed} traj {t}", W, np.abs(J[t, :, 3]) < 1e-3)
            rep.exact("reachable", rr[t], F[t]); rep.exact("state", ss[t], S[t]); rep.close("joints", jj[t], J[t])
            if rep.bad.any() or rep.ill.any():
                close(rep, max_ill=0.1)
        ok = np.array_equal(stt["emergency_stop"].astype(bool), E) and np.allclose(stt["previous_theta"], TH, atol=1e-9)
        print(f"soak {who} continuous {arm}: {T} trajectories x {W}, final states {'ok' if ok else 'DIFFER'}")
        fails += not ok
print(f"soak seed {seed}: {'OK' if not fails else str(fails) + ' FAILURES'} in {time.time() - t0:.0f}s")
sys.exit(1 if fails else 0)
