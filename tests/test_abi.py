"""CPU checks of the drop-in boundary: libr2ik.so builds for sm_100a, loads, and exports every
symbol include/r2ik.h declares; the ctypes structs match the C layouts; the product path refuses
to run without a GPU (no CPU fallback).  No compute entry is called here."""
import ctypes as C
import os
import re

import pytest

from parity import REPO
from reachy2_symbolic_ik_b200 import _abi, _native, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _native.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(REPO, "include", "r2ik.h")).read()
    declared = set(re.findall(r"\b(r2ik_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_sizes(lib):
    assert lib.r2ik_abi_version() == _abi.ABI_VERSION
    assert C.sizeof(_abi.ArmConfig) == 8 * 18 + 8
    assert C.sizeof(_abi.ArmConstants) == 8 * 9
    assert C.sizeof(_abi.CtlParams) == 8 * 6 + 8
    assert C.sizeof(_abi.TrajState) == 80


def test_interval_limit_host_entry(lib, oracle):
    import numpy as np
    for side in (1, -1):
        for low in (False, True):
            got = _native.interval_limit(side, low)
            want = np.zeros(2)
            oracle.lib().orc_interval_limit(side, int(low), want.ctypes.data_as(C.POINTER(C.c_double)))
            assert got == want.tolist()
    # unconstrained l_arm -> [-2pi/3, pi/4]
    np.testing.assert_allclose(_native.interval_limit(-1, False), [-2 * np.pi / 3, np.pi / 4], atol=1e-15)


def test_argument_errors_do_not_need_a_device(lib):
    h = C.c_void_p()
    assert lib.r2ik_create(None, 0, C.byref(h)) == 1  # R2IK_ERR_NULL
    assert b"null" in lib.r2ik_last_error()
    # argument checks come before any CUDA call
    assert lib.r2ik_ctl_discrete_compact_f64(None, None, None, C.c_int64(4), None, None, None, None, None, None, None, C.c_int64(0), None) == 1
    assert b"r2ik_ctl_discrete_compact_f64" in lib.r2ik_last_error()


def test_discrete_workspace_size_is_a_host_function(lib):
    """r2ik_ctl_discrete_workspace_bytes: header (two 128-byte lines) + per pose a 64-byte search plan, a theta and two
    list entries."""
    lib.r2ik_ctl_discrete_workspace_bytes.restype = C.c_int64
    for n in (0, 1, 1000, (1 << 31) - 1):
        assert lib.r2ik_ctl_discrete_workspace_bytes(C.c_int64(n)) == 256 + 80 * n
    assert lib.r2ik_ctl_discrete_workspace_bytes(C.c_int64(-5)) == -1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from reachy2_symbolic_ik_b200 import SymbolicIK

    with pytest.raises(_native.R2ikError):
        SymbolicIK(arm="r_arm")


def test_sass_is_sm100a_fp64():
    """The shipped cubin targets sm_100a and the hot kernel is FP64 (DFMA) code."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build.build()
    out = subprocess.run([cuobjdump, "-sass", _native.lib_path()], capture_output=True, text=True).stdout
    assert "arch = sm_100a" in out
    k1 = out.split("k_symik_solveILi1E")[1].split("Function :")[0]
    assert k1.count("DFMA") > 500


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's contract: stdout carries exactly one JSON line.  The reference arm (C oracle on the host cores)
    needs no GPU, so it is checked here."""
    import json
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ik_poses_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


BOUNDARY_METHODS = {   # what a user of the reference calls
    "SymbolicIK": ["__init__", "is_reachable", "is_reachable_no_limits", "get_joints", "get_elbow_position"],
    "ControlIK": ["__init__", "symbolic_inverse_kinematics"],
}


def test_facade_signatures_match_the_reference():
    """Parameter names, order and defaults of the drop-in classes against those of the reference's classes
    (tests/golden/api_surface.json, written by gen_golden.py from the unmodified reference).  The facade may append
    parameters of its own (device=...) after the reference's."""
    import inspect
    import json

    from reachy2_symbolic_ik_b200 import ControlIK, SymbolicIK

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(repo, "tests", "golden", "api_surface.json")))
    for cname, cls in (("SymbolicIK", SymbolicIK), ("ControlIK", ControlIK)):
        for m in BOUNDARY_METHODS[cname]:
            want = ref[cname][m]
            sig = inspect.signature(getattr(cls, m))
            mine = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)] for p in sig.parameters.values()]
            assert mine[: len(want)] == want, f"{cname}.{m}: {mine} != {want}"
            assert all(d is not None for _, d in mine[len(want):]), f"{cname}.{m}: extra parameters must have defaults"
