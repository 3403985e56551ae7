"""bench.py on a variant library (development aid):  python scripts/experiments/bench_with_lib.py <lib.so> [bench args]"""
import json
import os
import sys

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import _native  # noqa: E402

lib = sys.argv[1]
_native.use_library(lib)
sys.argv = ["bench.py"] + sys.argv[2:]
import bench  # noqa: E402

import io  # noqa: E402
import contextlib  # noqa: E402

buf = io.StringIO()
real = sys.stdout
rc = bench.main()
