"""Host-side forward kinematics of the Reachy2 arms and synthetic pose generators.

The reference library has no FK (its examples call the Reachy SDK's,
src/example/test_continuous_ik.py:146).  Synthetic workloads for tests and ``bench.py``
 are built here with NumPy from the arm chains of the bundled
``config_files/reachy2_arms.urdf`` (torso -> ``{r,l}_arm_tip``).  This is input generation,
not the IK hot path.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from functools import lru_cache

import numpy as np

_URDF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_files", "reachy2_arms.urdf")

JOINT_NAMES = ("shoulder_pitch", "shoulder_roll", "elbow_yaw", "elbow_pitch", "wrist_roll", "wrist_pitch", "wrist_yaw")


def bundled_urdf_path() -> str:
    return _URDF


def _rpy_matrix(rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    # URDF fixed-axis roll/pitch/yaw: R = Rz(yaw) Ry(pitch) Rx(roll)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


def _origin(xyz, rpy) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = _rpy_matrix(rpy)
    T[:3, 3] = xyz
    return T


@lru_cache(maxsize=4)
def arm_chain(arm: str, urdf_path: str = _URDF):
    """Segments of the chain torso -> tip: list of (fixed 4x4 prefix, revolute axis or None)."""
    prefix = arm[0]
    root = ET.parse(urdf_path).getroot()
    by_child = {j.find("child").attrib["link"]: j for j in root.findall("joint")}
    chain, link = [], f"{prefix}_arm_tip"
    while link != "torso":
        j = by_child[link]
        chain.append(j)
        link = j.find("parent").attrib["link"]
    chain.reverse()
    segs, acc = [], np.eye(4)
    for j in chain:
        o = j.find("origin").attrib
        acc = acc @ _origin([float(v) for v in o["xyz"].split()], [float(v) for v in o["rpy"].split()])
        if j.attrib["type"] == "revolute":
            axis = np.array([float(v) for v in j.find("axis").attrib["xyz"].split()])
            segs.append((acc, axis))
            acc = np.eye(4)
    segs.append((acc, None))
    n_rev = sum(1 for _, a in segs if a is not None)
    if n_rev != 7:
        raise ValueError(f"expected 7 revolute joints on the {arm} chain, found {n_rev}")
    return tuple(segs)


def _axis_rot(axis: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Rodrigues rotation about a unit axis for a vector of angles -> (N,4,4)."""
    n = q.shape[0]
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3)[None] + np.sin(q)[:, None, None] * K[None] + (1 - np.cos(q))[:, None, None] * (K @ K)[None]
    T = np.zeros((n, 4, 4))
    T[:, :3, :3] = R
    T[:, 3, 3] = 1.0
    return T


def forward_kinematics(joints: np.ndarray, arm: str = "r_arm") -> np.ndarray:
    """FK of the arm tip in the torso frame.  joints: (..., 7) -> poses (..., 4, 4)."""
    q = np.asarray(joints, dtype=np.float64)
    lead = q.shape[:-1]
    q = q.reshape(-1, 7)
    T = np.broadcast_to(np.eye(4), (q.shape[0], 4, 4)).copy()
    k = 0
    for fixed, axis in arm_chain(arm):
        T = T @ fixed
        if axis is not None:
            T = T @ _axis_rot(axis, q[:, k])
            k += 1
    return T.reshape(*lead, 4, 4)


def sample_fk_joints(n: int, rng: np.random.Generator) -> np.ndarray:
    """Joint distribution of SURVEY.md 8(d): U(-pi,pi)^7, elbow pitch U(-2.2,0), wrist roll/pitch U(-0.7,0.7)."""
    q = rng.uniform(-np.pi, np.pi, size=(n, 7))
    q[:, 3] = rng.uniform(-2.2, 0.0, size=n)
    q[:, 4:6] = rng.uniform(-0.7, 0.7, size=(n, 2))
    return q


def sample_fk_poses(n: int, arm: str = "r_arm", seed: int = 0, min_x: float | None = 0.05) -> np.ndarray:
    """n FK-sampled tip poses (n,4,4).  min_x rejects poses behind the torso plane
    (about half of the raw samples are 'Backward pose'); None keeps everything."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 4, 4))
    filled = 0
    while filled < n:
        m = max(1024, int((n - filled) * 2.2))
        T = forward_kinematics(sample_fk_joints(m, rng), arm)
        if min_x is not None:
            T = T[T[:, 0, 3] >= min_x]
        take = min(n - filled, T.shape[0])
        out[filled:filled + take] = T[:take]
        filled += take
    return out


def random_rotations(n: int, rng: np.random.Generator) -> np.ndarray:
    """Haar-distributed rotation matrices (n,3,3) from normalised Gaussian quaternions."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def sample_task_space_poses(n: int, arm: str = "r_arm", seed: int = 0) -> np.ndarray:
    """Uniform task-space box x Haar orientation (stresses every early-out state)."""
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(-0.1, 0.8, n), rng.uniform(-0.9, 0.5, n), rng.uniform(-0.7, 0.7, n)], axis=1)
    if arm.startswith("l"):
        p[:, 1] = -p[:, 1]
    T = np.zeros((n, 4, 4))
    T[:, :3, :3] = random_rotations(n, rng)
    T[:, :3, 3] = p
    T[:, 3, 3] = 1.0
    return T


def sinusoidal_trajectories(T: int, W: int, arm: str = "r_arm", seed: int = 4, dt: float = 1.0 / 120.0):
    """Joint-space sinusoids through FK -> (T, W, 4, 4) goal poses, after the reference's manual
    continuity test (src/example/test_continuous_ik.py:96-113); random phase per trajectory."""
    rng = np.random.default_rng(seed)
    q0 = np.deg2rad([-25.0, -40.0, 0.0, -45.0, 0.0, 0.0, 0.0])
    A = np.deg2rad([20.0, 20.0, 30.0, 45.0, 25.0, 25.0, 90.0])
    f = np.array([0.6, 0.34, 0.78, 0.18, 0.31, 0.47, 0.25])
    if arm.startswith("l"):
        mirror = np.array([1, -1, -1, 1, -1, 1, -1.0])
        q0, A = q0 * mirror, A * mirror
    phase = rng.uniform(0, 2 * np.pi, size=(T, 1, 7))
    t = (np.arange(W) * dt)[None, :, None]
    q = q0 + A * np.sin(2 * np.pi * f * t + phase)
    return forward_kinematics(q, arm), q


def fibonacci_orientations(n: int, seed: int = 5) -> np.ndarray:
    """Deterministic, well-spread set of n orientations as xyz Euler angles (n,3):
    Fibonacci sphere for the approach axis, golden-ratio spin about it."""
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    golden = (1 + 5 ** 0.5) / 2
    lam = 2 * np.pi * i / golden
    spin = 2 * np.pi * ((i * (golden - 1) + 0.1 * seed) % 1.0)
    # yaw = lam, pitch = phi - pi/2, roll = spin  (extrinsic xyz)
    e = np.stack([spin - np.pi, phi - np.pi / 2, (lam % (2 * np.pi)) - np.pi], axis=1)
    return e


# --------------------------------------------------------------------------------------------
# Device FK (libr2ik.so: r2ik_fk_f64) -- the same chain evaluated by a CUDA kernel, for
# workloads too large to build on the host (cfg 4: 65 536 x 1 000 waypoints) and for FK
# round-trip checks at full batch size.  No CPU fallback: needs the built library and a GPU.
# --------------------------------------------------------------------------------------------
def fk_chain_struct(arm: str):
    """R2ikFkChain for an arm from the bundled URDF."""
    from . import _abi

    ch = _abi.FkChain()
    k = 0
    for fixed, axis in arm_chain(arm):
        ch.fixed[k][:] = [float(v) for v in fixed[:3, :].reshape(-1)]
        if axis is not None:
            ch.axis[k][:] = [float(v) for v in axis / np.linalg.norm(axis)]
        k += 1
    return ch


def forward_kinematics_device(joints, arm: str = "r_arm", out=None):
    """joints: CUDA float64 tensor (..., 7) -> tip poses (..., 4, 4) CUDA float64 (asynchronous)."""
    import ctypes as C

    from . import _native

    torch = _native.require_cuda()
    if not (hasattr(joints, "is_cuda") and joints.is_cuda):
        raise _native.R2ikError("forward_kinematics_device needs a CUDA tensor (use forward_kinematics on the host)")
    q = joints.to(torch.float64).contiguous()
    lead = tuple(q.shape[:-1])
    n = q.numel() // 7
    if out is None:
        out = torch.empty((*lead, 4, 4), dtype=torch.float64, device=q.device)
    ch = fk_chain_struct(arm)
    with torch.cuda.device(q.device):
        s = torch.cuda.current_stream(q.device).cuda_stream
        rc = _native.load().r2ik_fk_f64(C.byref(ch), q.device.index, C.c_void_p(q.data_ptr()), C.c_int64(n),
                                        C.c_void_p(out.data_ptr()), C.c_void_p(s))
        _native.check(rc, "r2ik_fk_f64")
    return out


def sinusoidal_trajectories_device(T: int, W: int, arm: str = "r_arm", seed: int = 4, dt: float = 1.0 / 120.0,
                                   device=None, chunk: int = 4096):
    """Device version of ``sinusoidal_trajectories``: same joint-space sinusoids (phases from the
    same NumPy generator), FK on the GPU.  Returns (T, W, 4, 4) CUDA float64."""
    from . import _native

    torch = _native.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rng = np.random.default_rng(seed)
    q0 = np.deg2rad([-25.0, -40.0, 0.0, -45.0, 0.0, 0.0, 0.0])
    A = np.deg2rad([20.0, 20.0, 30.0, 45.0, 25.0, 25.0, 90.0])
    f = np.array([0.6, 0.34, 0.78, 0.18, 0.31, 0.47, 0.25])
    if arm.startswith("l"):
        mirror = np.array([1, -1, -1, 1, -1, 1, -1.0])
        q0, A = q0 * mirror, A * mirror
    phase = torch.from_numpy(rng.uniform(0, 2 * np.pi, size=(T, 1, 7))).to(dev)
    t = (torch.arange(W, dtype=torch.float64, device=dev) * dt)[None, :, None]
    q0d, Ad, fd = (torch.from_numpy(x).to(dev) for x in (q0, A, f))
    out = torch.empty((T, W, 4, 4), dtype=torch.float64, device=dev)
    for lo in range(0, T, chunk):
        hi = min(T, lo + chunk)
        q = q0d + Ad * torch.sin(2 * np.pi * fd * t + phase[lo:hi])
        forward_kinematics_device(q, arm, out=out[lo:hi])
    return out
