"""Experiment: is a strided host->device copy of the 96 useful bytes of every 128-byte pose (cudaMemcpy2DAsync,
width 96, source pitch 128) faster than the contiguous 128-byte copy?  (PCIe is the end-to-end bound of K1.)"""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reachy2_symbolic_ik_b200 import _native

L = _native.load()
n = 1 << 20
host = torch.randn(n, 16, dtype=torch.float64).pin_memory()
dev16 = torch.empty(n, 16, dtype=torch.float64, device="cuda")
dev12 = torch.empty(n, 12, dtype=torch.float64, device="cuda")
out_d = torch.randn(n, 98, device="cuda").to(torch.uint8)
out_h = torch.empty(n, 98, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def contiguous():
    with torch.cuda.stream(s1):
        dev16.copy_(host, non_blocking=True)


def strided(rows_per_call=n):
    for lo in range(0, n, rows_per_call):
        m = min(rows_per_call, n - lo)
        rc = L.r2ik_copy2d_async(C.c_void_p(dev12.data_ptr() + lo * 96), 96, C.c_void_p(host.data_ptr() + lo * 128), 128, 96, m,
                                 C.c_void_p(s1.cuda_stream))
        assert rc == 0


def d2h():
    with torch.cuda.stream(s2):
        out_h.copy_(out_d, non_blocking=True)


for name, fn, nbytes in (("H2D contiguous 128 B/pose", contiguous, 128), ("H2D 2-D 96 of 128 B/pose", strided, 96),
                         ("H2D 2-D, 128k-row calls", lambda: strided(1 << 17), 96), ("D2H contiguous 98 B/pose", d2h, 98)):
    t = timeit(fn)
    print(f"{name}: {t*1e3:.3f} ms per 1M poses = {n/t:.3e} poses/s, {n*nbytes/t/1e9:.1f} GB/s of useful bytes")
for name, fn in (("contiguous + D2H", lambda: (contiguous(), d2h())), ("2-D + D2H", lambda: (strided(1 << 17), d2h()))):
    t = timeit(fn)
    print(f"both directions, {name}: {t*1e3:.3f} ms per 1M poses = {n/t:.3e} poses/s")
assert torch.equal(dev12.cpu(), host[:, :12])
