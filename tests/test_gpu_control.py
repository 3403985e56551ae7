"""GPU parity of K2 (ControlIK discrete) and K3 (ControlIK continuous) through the facade /
C ABI, against the reference's golden outputs and the CPU oracle."""
import numpy as np
import pytest

from parity import Report, ill_conditioned_mask, load
from parity import COND_TOL, PERTURB, TOL  # noqa: E402

pytestmark = pytest.mark.gpu
ARMS = ("r_arm", "l_arm")


def urdf_params():
    u = load("symik_urdf.npz")
    return {k[len("param_"):]: u[k] for k in u.files if k.startswith("param_")}


@pytest.fixture(scope="module")
def controls():
    from reachy2_symbolic_ik_b200 import ControlIK

    return {False: ControlIK(urdf_path="../config_files/reachy2.urdf"),
            True: ControlIK(urdf_path="../config_files/reachy2.urdf", is_dvt=True)}


def test_constructor_contract():
    from reachy2_symbolic_ik_b200 import ControlIK

    with pytest.raises(ValueError, match="No URDF provided"):
        ControlIK()
    with pytest.raises(ValueError, match="Unknown Reachy model"):
        ControlIK(urdf_path="../config_files/reachy2.urdf", reachy_model="nope")
    with pytest.raises(ValueError, match="Error while parsing URDF"):
        ControlIK(urdf="<robot><joint></robot>")
    c = ControlIK(urdf_path="../config_files/reachy2.urdf", reachy_model="starter_kit_right")
    assert list(c.symbolic_ik_solver) == ["r_arm"]
    with pytest.raises(ValueError, match="Unknown type"):
        c.symbolic_inverse_kinematics("r_arm", np.eye(4), "bogus")
    np.testing.assert_allclose(c.preferred_theta["r_arm"], -2 * np.pi / 3)


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("variant", ["k20", "k360", "low", "dvt"])
def test_discrete_golden(controls, oracle, arm, variant):
    g = load(f"ctl_discrete_{arm}.npz")
    params = urdf_params()
    dvt = variant == "dvt"
    want_j, want_f, want_s = g[f"joints_{variant}"], g[f"reachable_{variant}"], g[f"state_{variant}"]
    M = np.ascontiguousarray(g["M"][: len(want_j)])
    kw = dict(nb_search_points=360 if variant == "k360" else 20,
              constrained_mode="low_elbow" if variant == "low" else "unconstrained")
    ocfg = oracle.arm_config(arm, ik_parameters=params, singularity_offset=0.03 if dvt else -1.01)
    opar = oracle.ControlParams(arm=arm, **kw)
    ill = ill_conditioned_mask(lambda p: oracle.ctl_discrete_batch(ocfg, opar, p.reshape(M.shape))[:3], M.reshape(len(M), -1))
    ctl = controls[dvt]
    ctl.nb_search_points = kw["nb_search_points"]
    try:
        joints, reach, state, emg = ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete", constrained_mode=kw["constrained_mode"])
    finally:
        ctl.nb_search_points = 20
    rep = Report(f"gpu ctl discrete {arm} {variant}", len(M), ill)
    rep.exact("reachable", reach, want_f)
    rep.exact("state", state, want_s)
    rep.close("joints", joints, want_j)
    assert not emg.any()
    rep.check(max_ill_fraction=0.02)


@pytest.mark.parametrize("arm", ARMS)
def test_discrete_vs_oracle_k360(controls, oracle, arm):
    """BASELINE config 3 shape (K = 360 samples) at a size the oracle finishes in seconds."""
    from reachy2_symbolic_ik_b200 import fk

    M = np.concatenate([fk.sample_fk_poses(15000, arm, seed=31, min_x=0.0), fk.sample_task_space_poses(5000, arm, seed=32)])
    params = urdf_params()
    ocfg = oracle.arm_config(arm, ik_parameters=params, singularity_offset=-1.01)
    opar = oracle.ControlParams(arm=arm, nb_search_points=360)
    ill = ill_conditioned_mask(lambda p: oracle.ctl_discrete_batch(ocfg, opar, p.reshape(M.shape))[:3], M.reshape(len(M), -1), n_trials=2)
    wj, wr, ws, we = oracle.ctl_discrete_batch(ocfg, opar, M)
    ctl = controls[False]
    ctl.nb_search_points = 360
    try:
        joints, reach, state, emg = ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete")
    finally:
        ctl.nb_search_points = 20
    rep = Report(f"gpu ctl discrete vs oracle K=360 {arm}", len(M), ill)
    rep.exact("reachable", reach, wr)
    rep.exact("state", state, ws)
    rep.close("joints", joints, wj)
    rep.check(max_ill_fraction=0.02)


def test_discrete_scalar_api_matches_reference_examples(controls):
    """Survey-session reference outputs for src/example/test_go_to.py pose #1."""
    g = load("symik_named.npz")
    ctl = controls[False]
    pose = g["r_arm_poses"][9]  # [[0.38, -0.2, -0.28], [0, -pi/2, 0]]
    from oracle import oracle as O

    M = np.eye(4)
    M[:3, :3] = O.matrix_from_euler_xyz(pose[1])
    M[:3, 3] = pose[0]
    joints, ok, state = ctl.symbolic_inverse_kinematics("r_arm", M, "discrete")
    assert ok and state == "reachable"
    np.testing.assert_allclose(joints, [-0.00460168557495466, -0.1062724336311957, 0.19960165027845525, -1.5707963267948966,
                                        -0.3622181452116582, 0.06703749820492533, -0.3622181452116573], atol=1e-9)
    # unreachable discrete calls return the previous solution (control_ik.py:457-458)
    M[:3, 3] = [0.0, -0.85, 0.0]
    joints, ok, state = ctl.symbolic_inverse_kinematics("r_arm", M, "discrete")
    assert not ok and state in ("Backward pose", "Pose out of reach")
    np.testing.assert_allclose(joints, [0, 0.2617993877991494, -0.17453292519943295, 0, 0, 0, 0], atol=1e-12)


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("variant", ["default", "cj", "dvt"])
def test_continuous_golden(controls, oracle, arm, variant):
    g = load(f"ctl_continuous_{arm}.npz")
    pre = {"default": "", "cj": "cj_", "dvt": "dvt_"}[variant]
    want_j, want_f, want_s = g[pre + "joints"], g[pre + "reachable"], g[pre + "state"]
    T, W = want_j.shape[:2]
    M = np.ascontiguousarray(g["M"][:T])
    from reachy2_symbolic_ik_b200 import ControlIK

    # fresh controller: previous_pose / previous_sol are per-instance state, like the reference's
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf", is_dvt=(variant == "dvt"))
    kw = {}
    if variant == "cj":
        kw = dict(current_joints=g["cj_current_joints"], current_pose=g["cj_current_pose"])
    joints, reach, state, st = ctl.symbolic_inverse_kinematics_batch(arm, M, "continuous", **kw)
    for t in range(T):
        rep = Report(f"gpu ctl continuous {arm} {variant} traj {t}", W)
        rep.exact("reachable", reach[t], want_f[t])
        rep.exact("state", state[t], want_s[t])
        rep.close("joints", joints[t], want_j[t])
        rep.check()
    np.testing.assert_array_equal(st["emergency_stop"].astype(bool), g[pre + "emergency"])
    np.testing.assert_allclose(st["previous_theta"], g[pre + "final_theta"], atol=1e-9)


def test_continuous_resume_is_identical(controls):
    """Chunked trajectories (state struct passed back in) reproduce the one-shot result: bit for bit with the serial kernel
    (a pure recursion, like the reference), to rounding with the phased pipeline (an ordinary waypoint's joints are its raw
    joints plus whole turns there, and the first waypoint of a call is the reference's prev + angle_diff(j, prev))."""
    g = load("ctl_continuous_r_arm.npz")
    from reachy2_symbolic_ik_b200 import ControlIK

    M = np.ascontiguousarray(g["M"])
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    W = M.shape[1]
    for phased in (False, True):
        j_all, r_all, s_all, st_all = ctl.symbolic_inverse_kinematics_batch("r_arm", M, "continuous", phased=phased)
        j1, r1, s1, st1 = ctl.symbolic_inverse_kinematics_batch("r_arm", M[:, : W // 3], "continuous", phased=phased)
        j2, r2, s2, st2 = ctl.symbolic_inverse_kinematics_batch("r_arm", M[:, W // 3:], "continuous", states=st1, phased=phased)
        np.testing.assert_array_equal(np.concatenate([s1, s2], axis=1), s_all)
        np.testing.assert_array_equal(np.concatenate([r1, r2], axis=1), r_all)
        if not phased:
            np.testing.assert_array_equal(np.concatenate([j1, j2], axis=1), j_all)
            assert st2.tobytes() == st_all.tobytes()
        else:
            np.testing.assert_allclose(np.concatenate([j1, j2], axis=1), j_all, rtol=0, atol=1e-12)
            np.testing.assert_allclose(st2["previous_sol"], st_all["previous_sol"], rtol=0, atol=1e-12)
            for f in ("previous_theta", "has_previous_sol", "init", "emergency_stop", "emergency_bits"):
                np.testing.assert_array_equal(st2[f], st_all[f])


@pytest.mark.parametrize("arm", ARMS)
def test_continuous_vs_oracle_many(controls, oracle, arm):
    """256 sinusoidal trajectories x 200 waypoints against the oracle; continuity invariant."""
    from reachy2_symbolic_ik_b200 import fk

    M, q = fk.sinusoidal_trajectories(256, 200, arm, seed=41)
    params = urdf_params()
    ocfg = oracle.arm_config(arm, ik_parameters=params, singularity_offset=-1.01)
    wj, wr, ws, wst = oracle.ctl_continuous_batch(ocfg, oracle.ControlParams(arm=arm), M)
    from reachy2_symbolic_ik_b200 import ControlIK

    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    joints, reach, state, st = ctl.symbolic_inverse_kinematics_batch(arm, M, "continuous")
    assert np.array_equal(reach, wr) and np.array_equal(state, ws)
    # every waypoint over 1e-9 is classified, as in the K1 / K2 tests (tests/parity.py): it must be ill-conditioned -- the
    # oracle's own answer for it moves by > 1e-10 when the trajectory is perturbed by 3e-13 (a trajectory is a recursion:
    # one such waypoint contaminates its tail) -- else it is a genuine failure
    err = np.abs(joints - wj).max(axis=2)
    over = err > TOL
    ill = np.zeros_like(over)
    rng = np.random.default_rng(4100)
    for _ in range(3):
        pj, pr, ps, _ = oracle.ctl_continuous_batch(ocfg, oracle.ControlParams(arm=arm), M + rng.uniform(-PERTURB, PERTURB, size=M.shape))
        ill |= (np.abs(pj - wj).max(axis=2) > COND_TOL) | (pr != wr) | (ps != ws)
    genuine = over & ~ill
    print(f"[gpu ctl continuous vs oracle {arm}] {over.size} waypoints: over {TOL:g}: {int(over.sum())} "
          f"({int((over & ill).sum())} ill-conditioned, {int(genuine.sum())} genuine); ill-conditioned waypoints {int(ill.sum())}; "
          f"max|err| well-conditioned {np.where(ill, 0.0, err).max():.2e} (all {err.max():.2e})")
    assert not genuine.any(), np.argwhere(genuine)[:10]
    assert ill.mean() <= 0.10, ill.mean()      # sinusoids sweep through the straight arm: ~6 % of the waypoints
    np.testing.assert_array_equal(st["emergency_stop"], wst["emergency_stop"])
    ok = ~st["emergency_stop"].astype(bool)
    step = np.abs(np.diff(joints[ok], axis=1))[:, 1:]  # waypoint 0 -> 1 follows the unconstrained init
    assert (step.max(axis=(0, 1)) <= np.array([0.5, 0.5, 0.5, 0.5, 1.0, 1.0, 1.0])).all()


def test_continuous_scalar_api(controls):
    """Scalar continuous calls (wall-clock timeout logic of control_ik.py:296-304 on the host)
    walk the same recursion as one batched trajectory."""
    from reachy2_symbolic_ik_b200 import ControlIK

    g = load("ctl_continuous_r_arm.npz")
    M = g["M"][0][:40]
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    out = [ctl.symbolic_inverse_kinematics("r_arm", M[w], "continuous") for w in range(len(M))]
    joints = np.array([o[0] for o in out])
    np.testing.assert_allclose(joints, g["joints"][0][:40], atol=1e-9)
    assert all(o[1] for o in out) and all(o[2] == "" for o in out)


def test_host_pipelines_match_device_calls(controls):
    """symbolic_inverse_kinematics_batch_host (pinned host in / out, chunked over 3 streams) returns exactly
    what the device-resident batched call returns, for ragged chunk counts, in both modes."""
    import torch

    from reachy2_symbolic_ik_b200 import fk

    ctl = controls[False]
    M = fk.sample_fk_poses(10_001, "r_arm", seed=31)
    want = ctl.symbolic_inverse_kinematics_batch("r_arm", M, "discrete")
    hin = torch.from_numpy(M).reshape(-1, 16).pin_memory()
    got = ctl.symbolic_inverse_kinematics_batch_host("r_arm", hin, "discrete", chunk=3000)
    np.testing.assert_array_equal(got[0].numpy(), want[0])
    np.testing.assert_array_equal(got[1].numpy().astype(bool), want[1])
    np.testing.assert_array_equal(got[2].numpy(), want[2])
    np.testing.assert_array_equal(got[3].numpy(), want[3])

    Mt = fk.sinusoidal_trajectories(37, 50, "l_arm", seed=32)[0]
    want = ctl.symbolic_inverse_kinematics_batch("l_arm", Mt, "continuous")
    got = ctl.symbolic_inverse_kinematics_batch_host("l_arm", torch.from_numpy(Mt).pin_memory(), "continuous", chunk=8)
    # the pipeline cuts the waypoint axis: flags / states identical, joints to rounding (see test_continuous_resume_is_identical)
    np.testing.assert_allclose(got[0].numpy(), want[0], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(got[1].numpy().astype(bool), want[1])
    np.testing.assert_array_equal(got[2].numpy(), want[2])
    gst = got[3].numpy().reshape(-1).view(want[3].dtype)
    np.testing.assert_allclose(gst["previous_sol"], want[3]["previous_sol"], rtol=0, atol=1e-12)
    for f in ("previous_theta", "has_previous_sol", "init", "emergency_stop", "emergency_bits"):
        np.testing.assert_array_equal(gst[f], want[3][f])


@pytest.mark.parametrize("W", [120, 30, 33, 300])   # multiple of the scan tile; even with a 2-waypoint tail; odd (unaligned code rows); several 128-waypoint blocks
@pytest.mark.parametrize("arm", ARMS)
def test_continuous_phased_equals_serial_kernel(controls, arm, W):
    """The phased K3 (per-waypoint kernels + per-trajectory scans) and the one-thread-per-trajectory K3 are the
    same arithmetic regrouped: flags, state codes and the integer controller state must be identical, and the joints
    / thetas equal to rounding (the two forms are separate compilations, so FMA contraction may differ in the last
    bit: 1e-12 rad is allowed), also when a trajectory latches an emergency stop, hits invalid rotations or is
    resumed from a previous state."""
    from reachy2_symbolic_ik_b200 import fk

    ctl = controls[False]
    M = fk.sinusoidal_trajectories(300, W, arm, seed=41)[0].copy()
    flip = np.diag([-1.0, -1.0, 1.0])          # half a turn about the tool axis from waypoint 13 on: the wrist
    M[3, 13:, :3, :3] = M[3, 13:, :3, :3] @ flip  # jumps by pi -> continuity violation -> emergency latch on trajectory 3
    M[5, 7, :3, :3] = np.diag([-1.0, 1.0, 1.0])    # det < 0 in the middle of trajectory 5
    M[6, 0, :3, :3] = np.diag([1.0, -1.0, 1.0])    # ... and at the first waypoint of trajectory 6
    a = ctl.symbolic_inverse_kinematics_batch(arm, M, "continuous", phased=True)
    b = ctl.symbolic_inverse_kinematics_batch(arm, M, "continuous", phased=False)
    def same(p, q):
        np.testing.assert_allclose(p[0], q[0], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(p[1], q[1])
        np.testing.assert_array_equal(p[2], q[2])
        for f in ("has_previous_sol", "init", "emergency_stop", "emergency_bits"):
            np.testing.assert_array_equal(p[3][f], q[3][f])
        np.testing.assert_allclose(p[3]["previous_theta"], q[3]["previous_theta"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(p[3]["previous_sol"], q[3]["previous_sol"], rtol=0, atol=1e-12)

    same(a, b)
    assert a[3]["emergency_stop"][3] == 1 and (a[2][3, -10:] == 8).all()
    # resume from the returned states
    a2 = ctl.symbolic_inverse_kinematics_batch(arm, M[:, ::-1].copy(), "continuous", states=a[3], phased=True)
    b2 = ctl.symbolic_inverse_kinematics_batch(arm, M[:, ::-1].copy(), "continuous", states=b[3], phased=False)
    same(a2, b2)


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("K", [20, 360, 1000])
def test_discrete_analytic_search_equals_exhaustive_scan(controls, arm, K):
    """K2's default elbow search locates the arg-min over the K samples from the crossings of the two elbow tests;
    the scan kernel visits every sample (warp-cooperative, shuffle arg-min).  Same theta, hence identical outputs."""
    from reachy2_symbolic_ik_b200 import fk

    ctl = controls[False]
    M = np.concatenate([fk.sample_fk_poses(150_000, arm, seed=51 + K), fk.sample_task_space_poses(50_000, arm, seed=52 + K)])
    old = ctl.nb_search_points
    try:
        ctl.nb_search_points = K
        a = ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete")
        b = ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete", exhaustive=True)
        for mode in ("low_elbow",):
            a2 = ctl.symbolic_inverse_kinematics_batch(arm, M[:50_000], "discrete", constrained_mode=mode)
            b2 = ctl.symbolic_inverse_kinematics_batch(arm, M[:50_000], "discrete", constrained_mode=mode, exhaustive=True)
    finally:
        ctl.nb_search_points = old
    searched = int((a[2] == 6).sum() + (a[1] & (a[2] == 0)).sum())
    assert searched > 10_000
    for (x, y) in list(zip(a, b)) + list(zip(a2, b2)):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("arm", ARMS)
def test_continuous_serial_route_and_fixup(controls, arm):
    """A waypoint whose get_joints needs the previous solution (exact singularities) stops the lane-parallel finish
    scan of its trajectory; a fixup kernel resumes it serially.  No physical pose triggers that, so the library's test
    hook sends every m-th waypoint down that route: outputs and final states must not change."""
    from reachy2_symbolic_ik_b200 import fk

    ctl = controls[False]
    M = fk.sinusoidal_trajectories(257, 90, arm, seed=43)[0].copy()
    flip = np.diag([-1.0, -1.0, 1.0])
    M[3, 13:, :3, :3] = M[3, 13:, :3, :3] @ flip       # emergency latch on trajectory 3
    M[5, 7, :3, :3] = np.diag([-1.0, 1.0, 1.0])        # invalid rotation in trajectory 5
    want = ctl.symbolic_inverse_kinematics_batch(arm, M, "continuous")
    for m in (1, 7, 97):
        got = ctl.symbolic_inverse_kinematics_batch(arm, M, "continuous", _test_force_serial_mod=m)
        np.testing.assert_allclose(got[0], want[0], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(got[1], want[1])
        np.testing.assert_array_equal(got[2], want[2])
        for f in ("has_previous_sol", "init", "emergency_stop", "emergency_bits"):
            np.testing.assert_array_equal(got[3][f], want[3][f])
        np.testing.assert_allclose(got[3]["previous_sol"], want[3]["previous_sol"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("arm", ARMS)
def test_reference_example_matrices(arm):
    """The goal matrices printed in the reference's examples (src/example/test_continuous_ik.py:345-370,
    test_go_to.py:250-257; tests/golden/ctl_examples.npz) through the library: SymbolicIK, ControlIK discrete (scalar
    and batched), ControlIK continuous."""
    from reachy2_symbolic_ik_b200 import ControlIK, SymbolicIK

    g = load("ctl_examples.npz")
    M = np.ascontiguousarray(g[f"{arm}_M"])
    res = SymbolicIK(arm=arm).is_reachable_batch(M)
    assert np.array_equal(res.state, g[f"{arm}_sym_state"]) and np.array_equal(res.reachable, g[f"{arm}_sym_reachable"])
    np.testing.assert_allclose(res.theta_interval, g[f"{arm}_sym_interval"], atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(res.joints, g[f"{arm}_sym_joints"], atol=1e-9, equal_nan=True)
    ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
    j, r, s, e = ctl.symbolic_inverse_kinematics_batch(arm, M, "discrete")
    assert np.array_equal(s, g[f"{arm}_dis_state"]) and np.array_equal(r, g[f"{arm}_dis_reachable"])
    np.testing.assert_allclose(j, g[f"{arm}_dis_joints"], atol=1e-9)
    for i in range(len(M)):   # the scalar API, a fresh controller per call like the fixture
        js, ok, st = ControlIK(urdf_path="../config_files/reachy2.urdf").symbolic_inverse_kinematics(arm, M[i], "discrete")
        np.testing.assert_allclose(js, g[f"{arm}_dis_joints"][i], atol=1e-9)
        assert ok == bool(g[f"{arm}_dis_reachable"][i])
    W = g[f"{arm}_con_joints"].shape[1]
    MT = np.ascontiguousarray(np.repeat(M[:, None], W, axis=1))
    for phased in (True, False):
        cj, cr, cs, st = ControlIK(urdf_path="../config_files/reachy2.urdf").symbolic_inverse_kinematics_batch(
            arm, MT, "continuous", phased=phased)
        assert np.array_equal(cs, g[f"{arm}_con_state"]) and np.array_equal(cr, g[f"{arm}_con_reachable"])
        np.testing.assert_allclose(cj, g[f"{arm}_con_joints"], atol=1e-9)
        assert np.array_equal(st["emergency_stop"].astype(bool), g[f"{arm}_con_emergency"])
