"""GPU parity of K1 (SymbolicIK.is_reachable + theta_to_joints) through the facade / C ABI,
against the reference's golden outputs and against the CPU oracle on larger seeded batches."""
import numpy as np
import pytest

from parity import Report, check_f32, ill_conditioned_mask, load

pytestmark = pytest.mark.gpu
ARMS = ("r_arm", "l_arm")


def urdf_params():
    u = load("symik_urdf.npz")
    return {k[len("param_"):]: u[k] for k in u.files if k.startswith("param_")}


@pytest.fixture(scope="module")
def solvers():
    from reachy2_symbolic_ik_b200 import SymbolicIK

    return {arm: SymbolicIK(arm=arm) for arm in ARMS}


def test_reference_ci_test(solvers):
    """tests/test_ik.py:12-79 of the reference, verbatim assertions, scalar API."""
    symbolic_ik = solvers["r_arm"]
    goal_pose = [[0.4, 0.2, 0.1], [np.radians(-60), np.radians(-90), np.radians(20)]]
    result = symbolic_ik.is_reachable(goal_pose)
    assert not (result[0])
    assert len(result[1]) == 0
    assert result[2] is None

    goal_pose = [[0.3, -0.2, -0.3], [np.radians(0), np.radians(-90), np.radians(0)]]
    result = symbolic_ik.is_reachable(goal_pose)
    assert result[0]
    assert result[1][0] >= -np.pi
    assert result[1][1] <= np.pi
    assert result[2] is not None
    joints, elbow_position = result[2](result[1][0])
    assert len(joints) == 7

    goal_pose = [[0.02, -0.2, -0.65], [0.0, 0.0, 0.0]]
    result = symbolic_ik.is_reachable(goal_pose)
    assert result[0]
    assert np.all(result[1] == [-np.pi, np.pi])
    assert result[2] is not None
    joints, elbow_position = result[2](0)
    assert len(joints) == 7

    result = symbolic_ik.is_reachable([[0.0, -0.2, -0.65], [0.0, 0.0, 0.0]])
    assert not (result[0])
    result = symbolic_ik.is_reachable([[0.87, -0.2, -0.0], [0.0, -np.pi / 2, 0.0]])
    assert not (result[0])
    result = symbolic_ik.is_reachable([[0.35, -0.2, -0.28], [0.0, -np.pi / 2, 0.0]])
    assert result[0]


def test_readme_example(solvers):
    """README.md:65-80 of the reference."""
    ik = solvers["r_arm"]
    goal_pose = [[0.55, -0.3, -0.15], [0, -np.pi / 2, 0]]
    is_reachable, theta_interval, theta_to_joints_func, state = ik.is_reachable(goal_pose)
    assert is_reachable and state == "reachable"
    np.testing.assert_allclose(theta_interval, [2.18952378, -0.22393633], atol=1e-8)
    joints, elbow_position = theta_to_joints_func(theta_interval[0])
    np.testing.assert_allclose(
        joints, [-1.52495747, -0.68439452, -3.93117303, -1.04866976, -0.4404594, 0.61794447, -2.28174079], atol=1e-8)
    assert np.asarray(ik.get_elbow_position(theta_interval[0])).shape == (4,)


def test_constructor_errors():
    from reachy2_symbolic_ik_b200 import SymbolicIK

    with pytest.raises(ValueError, match="arm should be either 'r_arm' or 'l_arm'"):
        SymbolicIK(arm="x_arm")


@pytest.mark.parametrize("arm", ARMS)
def test_named_poses(solvers, oracle, arm):
    g = load("symik_named.npz")
    P = g[f"{arm}_poses"]
    ocfg = oracle.arm_config(arm)
    ill = ill_conditioned_mask(lambda p: oracle.symik_batch(ocfg, p.reshape(P.shape))[:4], P.reshape(len(P), -1))
    res = solvers[arm].is_reachable_batch(P)
    rep = Report(f"gpu named {arm}", len(P), ill)
    rep.exact("reachable", res.reachable, g[f"{arm}_reachable"])
    rep.exact("state", res.state, g[f"{arm}_state"])
    rep.close("interval", res.theta_interval, g[f"{arm}_interval"])
    rep.close("joints", res.joints, g[f"{arm}_joints"])
    rep.close("elbow", res.elbow, g[f"{arm}_elbow"])
    res0 = solvers[arm].is_reachable_batch(P, theta=np.zeros(len(P)))
    rep.close("joints(theta=0)", res0.joints, g[f"{arm}_joints_theta0"])
    rep.check(max_ill_fraction=0.1)


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("layout", ["euler", "mat4"])
def test_random_golden(solvers, oracle, arm, layout):
    g = load(f"symik_random_{arm}.npz")
    P = g["goal_pose"] if layout == "euler" else g["M"]
    ocfg = oracle.arm_config(arm)
    ill = ill_conditioned_mask(lambda p: oracle.symik_batch(ocfg, p.reshape(P.shape))[:4], P.reshape(len(P), -1))
    res = solvers[arm].is_reachable_batch(P)
    rep = Report(f"gpu random golden {arm} {layout}", len(P), ill)
    rep.exact("reachable", res.reachable, g["reachable"])
    rep.exact("state", res.state, g["state"])
    rep.close("interval", res.theta_interval, g["interval"])
    rep.close("joints@interval[0]", res.joints, g["joints"])
    rep.close("elbow", res.elbow, g["elbow"])
    res2 = solvers[arm].is_reachable_batch(P, theta=g["theta2"])
    rep.close("joints@theta2", res2.joints, g["joints_theta2"])
    rep.close("elbow@theta2", res2.elbow, g["elbow_theta2"])
    rep.check()


@pytest.mark.parametrize("arm", ARMS)
def test_urdf_params_and_no_limits(oracle, arm):
    from reachy2_symbolic_ik_b200 import SymbolicIK

    g = load("symik_urdf.npz")
    params = urdf_params()
    M = g[f"{arm}_M"]
    th = g[f"{arm}_nl_theta"]
    ocfg = oracle.arm_config(arm, ik_parameters=params, singularity_offset=-1.01)
    run = lambda p: oracle.symik_batch(ocfg, p.reshape(M.shape))[:4] + oracle.symik_no_limits_batch(  # noqa: E731
        ocfg, p.reshape(M.shape), th)
    ill = ill_conditioned_mask(run, M.reshape(len(M), -1))
    ik = SymbolicIK(arm=arm, ik_parameters=params, singularity_offset=-1.01)
    res = ik.is_reachable_batch(M)
    rep = Report(f"gpu urdf {arm}", len(M), ill)
    rep.exact("reachable", res.reachable, g[f"{arm}_reachable"])
    rep.exact("state", res.state, g[f"{arm}_state"])
    rep.close("interval", res.theta_interval, g[f"{arm}_interval"])
    rep.close("joints", res.joints, g[f"{arm}_joints"])
    nj, ne = ik.is_reachable_no_limits_batch(M, th)
    rep.close("no_limits joints", nj, g[f"{arm}_nl_joints"])
    rep.close("no_limits elbow", ne, g[f"{arm}_nl_elbow"])
    rep.check()
    # attributes the reference exposes
    np.testing.assert_allclose(ik.max_arm_length, 0.66, atol=1e-15)
    np.testing.assert_allclose(ik.shoulder_wrist_min_distance, 0.2498707753414929, atol=1e-15)


@pytest.mark.parametrize("arm", ARMS)
def test_vs_oracle_100k(solvers, oracle, arm):
    """Seeded FK + task-space batch at a size the oracle finishes in seconds."""
    from reachy2_symbolic_ik_b200 import fk

    seed = 7 if arm == "r_arm" else 8
    M = np.concatenate([fk.sample_fk_poses(60000, arm, seed=seed, min_x=None),
                        fk.sample_task_space_poses(40000, arm, seed=seed + 100)])
    ocfg = oracle.arm_config(arm)
    ill = ill_conditioned_mask(lambda p: oracle.symik_batch(ocfg, p.reshape(M.shape))[:4], M.reshape(len(M), -1), n_trials=2)
    want = oracle.symik_batch(ocfg, M)
    res = solvers[arm].is_reachable_batch(M)
    rep = Report(f"gpu vs oracle 100k {arm}", len(M), ill)
    rep.exact("reachable", res.reachable, want[0])
    rep.exact("state", res.state, want[2])
    rep.close("interval", res.theta_interval, want[1])
    rep.close("joints", res.joints, want[3])
    rep.close("elbow", res.elbow, want[4])
    rep.check()


def test_edge_cases(solvers, oracle):
    ik = solvers["r_arm"]
    # empty batch
    res = ik.is_reachable_batch(np.zeros((0, 4, 4)))
    assert res.reachable.shape == (0,) and res.joints.shape == (0, 7)
    # ragged size (not a multiple of the block) and a left-handed rotation (scipy raises ValueError)
    from reachy2_symbolic_ik_b200 import fk

    M = fk.sample_fk_poses(131, "r_arm", seed=3)
    M[5, :3, :3] = np.diag([-1.0, 1.0, 1.0])
    res = ik.is_reachable_batch(M)
    assert res.state[5] == 9 and not res.reachable[5] and np.isnan(res.joints[5]).all()
    want = oracle.symik_batch(oracle.arm_config("r_arm"), M)
    assert np.array_equal(res.state, want[2])
    np.testing.assert_allclose(res.joints, want[3], atol=1e-9, equal_nan=True)
    # truncated (non-orthonormal) matrix like src/example/test_go_to.py:250-257 (scipy: SVD projection)
    Mt = np.array([[[0.36861, 0.089736, -0.92524, 0.37213], [-0.068392, 0.99525, 0.069279, -0.028012],
                    [0.92706, 0.037742, 0.373, -0.38572], [0, 0, 0, 1]]])
    res = ik.is_reachable_batch(Mt)
    want = oracle.symik_batch(oracle.arm_config("r_arm"), Mt)
    assert np.array_equal(res.state, want[2])
    np.testing.assert_allclose(res.theta_interval, want[1], atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(res.joints, want[3], atol=1e-9, equal_nan=True)


def test_cuda_tensor_in_cuda_tensor_out(solvers):
    import torch

    from reachy2_symbolic_ik_b200 import fk

    M = torch.from_numpy(fk.sample_fk_poses(1000, "r_arm", seed=11)).cuda()
    res = solvers["r_arm"].is_reachable_batch(M)
    assert res.joints.is_cuda and res.reachable.dtype == torch.bool
    res_np = solvers["r_arm"].is_reachable_batch(M.cpu().numpy())
    np.testing.assert_array_equal(res.joints.cpu().numpy(), res_np.joints)


def test_mirror_symmetry(solvers):
    """Invariant used by the reference's examples (test_continuous_ik.py:148-155): with
    M_l = S M_r S, S = diag(1,-1,1), the l_arm solution at theta_l = pi - theta_r is the r_arm
    solution with joints 1, 2, 4, 6 negated."""
    from reachy2_symbolic_ik_b200 import fk

    Mr = fk.sample_fk_poses(20000, "r_arm", seed=21)
    S = np.diag([1.0, -1.0, 1.0, 1.0])
    Ml = S @ Mr @ S
    r = solvers["r_arm"].is_reachable_batch(Mr)
    th_r = np.where(r.reachable, r.theta_interval[:, 0], 0.0)
    r = solvers["r_arm"].is_reachable_batch(Mr, theta=th_r)
    l = solvers["l_arm"].is_reachable_batch(Ml, theta=np.pi - th_r)
    both = r.reachable & l.reachable
    assert (r.reachable == l.reachable).mean() > 0.9999
    sign = np.array([1, -1, -1, 1, -1, 1, -1.0])
    d = np.abs(np.angle(np.exp(1j * (l.joints[both] - r.joints[both] * sign))))
    assert np.quantile(d.max(axis=1), 0.999) < 1e-9


@pytest.mark.parametrize("n", [1, 31, 33, 127, 4099, 200_001])
def test_ragged_sizes_and_misaligned_outputs(solvers, n):
    """Ragged batch sizes; output buffers that start off a 16-byte boundary (only the pose buffer has an
    alignment contract) must give the same results as aligned ones."""
    import torch

    from reachy2_symbolic_ik_b200 import _abi, fk

    ik = solvers["r_arm"]
    M = np.concatenate([fk.sample_fk_poses(n - n // 3, "r_arm", seed=n), fk.sample_task_space_poses(n // 3, "r_arm", seed=n + 1)])
    dev = torch.device("cuda", 0)

    def run(offset):
        P = torch.from_numpy(M).reshape(-1).to(dev)
        reach = torch.zeros(n + 16, dtype=torch.uint8, device=dev)[offset:offset + n]
        state = torch.zeros(n + 16, dtype=torch.uint8, device=dev)[offset:offset + n]
        itv = torch.zeros(n * 2 + 2, dtype=torch.float64, device=dev)[offset:offset + 2 * n]
        j = torch.zeros(n * 7 + 2, dtype=torch.float64, device=dev)[offset:offset + 7 * n]
        e = torch.zeros(n * 3 + 2, dtype=torch.float64, device=dev)[offset:offset + 3 * n]

        class V:   # solve_into only needs data_ptr() / shape[0]
            def __init__(s, t, rows): s.t, s.shape = t, (rows,)
            def data_ptr(s): return s.t.data_ptr()
        ik.solve_into(V(P, n), _abi.POSE_MAT4, None, None, reach, state, itv, j, e)
        torch.cuda.synchronize()
        return [x.cpu().numpy().copy() for x in (reach, state, itv, j, e)]

    aligned, misaligned = run(0), run(1)
    for a, b in zip(aligned, misaligned):
        np.testing.assert_array_equal(a, b)
    assert aligned[0].sum() > 0 or n < 4


def test_misaligned_pose_pointer_is_an_argument_error(solvers):
    import torch

    from reachy2_symbolic_ik_b200 import _abi, _native

    ik = solvers["r_arm"]
    buf = torch.zeros(16 * 4 + 1, dtype=torch.float64, device="cuda")
    P = buf[1:]

    class V:
        shape = (4,)
        def data_ptr(self): return P.data_ptr()
    o = ik.alloc_host_outputs(4)
    with pytest.raises(_native.R2ikError, match="16-byte aligned"):
        ik.solve_into(V(), _abi.POSE_MAT4, None, None, o.reachable, o.state, o.theta_interval, o.joints, o.elbow)
    # the facade re-aligns such a view instead of failing
    res = ik.is_reachable_batch(P.reshape(4, 16))
    assert res.state.shape == (4,)


# ---------------------------------------------------------------------------------------------------
# FP32 fast path (r2ik_symik_solve_f32): float32 poses in, float32 results out; checked against the FP64
# oracle on the same inputs widened to double.  Tolerance 1e-4 rad (tests/parity.py TOL_F32), states exact.
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("arm", ARMS)
def test_fp32_vs_oracle_100k(solvers, oracle, arm):
    from reachy2_symbolic_ik_b200 import fk

    seed = 31 if arm == "r_arm" else 32
    P32 = np.concatenate([fk.sample_fk_poses(70000, arm, seed=seed), fk.sample_task_space_poses(30000, arm, seed=seed + 100)]
                         ).astype(np.float32)
    res = solvers[arm].is_reachable_batch(P32, precision="fp32")
    assert res.joints.dtype == np.float32 and res.theta_interval.dtype == np.float32
    n_esc = res.n_escalated
    esc = np.zeros(len(P32), bool)
    esc[:n_esc] = True   # only the count is known on the GPU path; check_f32 uses its mean
    check_f32(f"gpu f32 vs oracle 100k {arm}", oracle, arm, P32, (res.reachable, res.theta_interval, res.state, res.joints,
                                                                    res.elbow, esc))


@pytest.mark.parametrize("arm", ARMS)
def test_fp32_golden_layouts_and_theta(solvers, oracle, arm):
    g = load(f"symik_random_{arm}.npz")
    th = g["theta2"].astype(np.float32)
    for P in (g["goal_pose"], g["M"]):
        P32 = P.astype(np.float32)
        for theta in (None, th):
            res = solvers[arm].is_reachable_batch(P32, theta=theta, precision="fp32")
            esc = np.zeros(len(P32), bool); esc[:res.n_escalated] = True
            check_f32(f"gpu f32 golden {arm} {P.shape[1:]} theta={'given' if theta is not None else 'i0'}", oracle, arm, P32,
                      (res.reachable, res.theta_interval, res.state, res.joints, res.elbow, esc), theta=theta)


def test_fp32_named_degenerate_and_plumbing(solvers, oracle):
    import torch

    from reachy2_symbolic_ik_b200 import fk

    g = load("symik_named.npz")
    for arm in ARMS:
        P32 = g[f"{arm}_poses"].astype(np.float32)
        res = solvers[arm].is_reachable_batch(P32, precision="fp32")
        want = oracle.symik_batch(oracle.arm_config(arm), P32.astype(np.float64))
        assert np.array_equal(res.state, want[2]) and np.array_equal(res.reachable, want[0])
    ik = solvers["r_arm"]
    # scaled / left-handed / null rotations are handed to the FP64 solver
    M = np.tile(np.eye(4, dtype=np.float32), (4, 1, 1))
    M[:, :3, 3] = [0.3, -0.2, -0.3]
    M[1, :3, :3] *= 1.5
    M[2, 0, 0] = -1.0
    M[3, :3, :3] = 0.0
    res = ik.is_reachable_batch(M, precision="fp32")
    want = oracle.symik_batch(oracle.arm_config("r_arm"), M.astype(np.float64))
    assert np.array_equal(res.state, want[2]) and res.n_escalated >= 3
    assert np.nanmax(np.abs(res.joints - want[3])) < 1e-5
    # empty and ragged batches; CUDA tensors in -> CUDA tensors out; float64 input is narrowed
    assert ik.is_reachable_batch(np.zeros((0, 4, 4), np.float32), precision="fp32").joints.shape == (0, 7)
    M = fk.sample_fk_poses(1001, "r_arm", seed=12)
    a = ik.is_reachable_batch(torch.from_numpy(M).cuda(), precision="fp32")
    assert a.joints.is_cuda and a.joints.dtype == torch.float32
    b = ik.is_reachable_batch(M.astype(np.float32), precision="fp32")
    np.testing.assert_array_equal(a.joints.cpu().numpy(), b.joints)
    # host pipeline (pinned float32 buffers) gives the device call's results
    P = fk.sample_fk_poses(300_001, "r_arm", seed=13).astype(np.float32)
    out = ik.is_reachable_batch_host(torch.from_numpy(P).pin_memory(), precision="fp32", chunk=1 << 16)
    d = ik.is_reachable_batch(P, precision="fp32")
    np.testing.assert_array_equal(out.joints.numpy(), d.joints)
    np.testing.assert_array_equal(out.state.numpy(), d.state)
    with pytest.raises(ValueError):
        ik.is_reachable_batch(P[:4], precision="fp16")
