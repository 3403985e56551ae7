"""Why does the lean pipeline reach 39 GB/s out + 33 GB/s in when plain 256 MB copies reach 50 + 50?  Copy patterns without kernels."""
import time
import torch

torch.cuda.init()
n = 1_000_000
h_in = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
h_j = torch.empty(n * 56, dtype=torch.uint8).pin_memory()
h_s = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
d_j = torch.empty(n * 56, dtype=torch.uint8, device="cuda")
d_s = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:70s} {dt * 1e3:6.2f} ms / 1M poses -> {n / dt:.3e} poses/s (out {57 * n / dt / 1e9:.1f} GB/s, in {48 * n / dt / 1e9:.1f} GB/s)", flush=True)


def whole(h2d=True, d2h=True):
    if h2d:
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            h_j.copy_(d_j, non_blocking=True)
            h_s.copy_(d_s, non_blocking=True)


def chunked(c, deps, two_out=True):
    ev = None
    for lo in range(0, n, c):
        hi = min(n, lo + c)
        with torch.cuda.stream(s1):
            d_in[lo * 48:hi * 48].copy_(h_in[lo * 48:hi * 48], non_blocking=True)
            if deps:
                ev = torch.cuda.Event(); ev.record(s1)
        with torch.cuda.stream(s2):
            if deps:
                s2.wait_event(ev)
            h_j[lo * 56:hi * 56].copy_(d_j[lo * 56:hi * 56], non_blocking=True)
            if two_out:
                h_s[lo:hi].copy_(d_s[lo:hi], non_blocking=True)


run("D2H only, whole", lambda: whole(False, True))
run("H2D only, whole", lambda: whole(True, False))
run("both, whole buffers, no dependency", whole)
for c in (1 << 16, 1 << 18):
    run(f"both, chunks of {c}, no dependency", lambda: chunked(c, False))
    run(f"both, chunks of {c}, D2H(i) after H2D(i)", lambda: chunked(c, True))
    run(f"both, chunks of {c}, D2H(i) after H2D(i), joints only", lambda: chunked(c, True, False))
