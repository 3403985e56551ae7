"""K3 timing at cfg 4 (65 536 x 1 000): tiled vs the four-kernel phased form vs the serial kernel; checks they agree.
    python scripts/experiments/exp_r2_k3.py [lib.so] [modes...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import _native  # noqa: E402

args = sys.argv[1:]
if args and args[0].endswith(".so"):
    _native.use_library(args.pop(0))
from reachy2_symbolic_ik_b200 import ControlIK, _abi, fk  # noqa: E402

modes = args or ["codes", "serial"]
T, W = 65536, 1000
ctl = ControlIK(urdf_path="../config_files/reachy2.urdf")
dM = fk.sinusoidal_trajectories_device(T, W, "r_arm", seed=4, device=torch.device("cuda"))
st0 = np.zeros(T, dtype=_abi.TRAJ_STATE_DTYPE); st0["init"] = 1
st0 = torch.from_numpy(st0.view(np.uint8).reshape(T, -1)).cuda()
ref = None
for mode in modes:
    phased = {"codes": True, "serial": False}[mode]
    st = st0.clone()
    out = None
    ts = []
    for it in range(5):
        st.copy_(st0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ctl.symbolic_inverse_kinematics_batch("r_arm", dM, "continuous", states=st, out=out, phased=phased)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    j, r, s = out[0], out[1], out[2]
    msg = ""
    if ref is None:
        ref = (j.clone(), r.clone(), s.clone(), st.clone())
    else:
        dj = (j - ref[0]).abs().nan_to_num().max().item()
        dps = (st.view(torch.float64)[:, :8] - ref[3].view(torch.float64)[:, :8]).abs().max().item()
        msg = f"max|d previous_sol/theta| {dps:.1e}, max|dj| vs {modes[0]} {dj:.2e}, flags equal {bool((r == ref[1]).all())}, states equal {bool((s == ref[2]).all())}, traj states equal {bool((st[:, 64:] == ref[3][:, 64:]).all())}"
    print(f"{mode:8s} {min(ts[1:]):8.3f} ms (runs {['%.2f' % t for t in ts]})  {T * W / min(ts[1:]) * 1e3:.3e} waypoints/s  reach {r.float().mean().item():.3f} {msg}", flush=True)
