"""Build tuning variants of libr2ik.so into reachy2_symbolic_ik_b200/lib/variants/ (development aid).

    python scripts/build_variants.py name=-DFOO=1,-DBAR=2 name2=...
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reachy2_symbolic_ik_b200 import build as b

out_dir = os.path.join(b.PKG, "lib", "variants")
os.makedirs(out_dir, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    extra = [f for f in flags.split(",") if f]
    out = os.path.join(out_dir, f"libr2ik_{name}.so")
    cmd = [b.nvcc_path(), *b.NVCC_FLAGS, *extra, "-Xptxas=-v", "-o", out, *[os.path.join(b.CSRC, s) for s in b.SOURCES]]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    log = p.communicate()[0]
    lines = log.splitlines()
    for i, l in enumerate(lines):
        if os.environ.get("R2IK_VARIANT_KERNEL", "k_symik_solveILi1") in l and "Compiling" in l:
            print(name, "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
    if p.returncode != 0:
        print(name, "FAILED\n", log[-2000:])
