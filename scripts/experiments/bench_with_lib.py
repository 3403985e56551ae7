"""bench.py on a variant library (development aid):  python scripts/experiments/bench_with_lib.py <lib.so> [bench args]"""
import sys

sys.path.insert(0, ".")
from reachy2_symbolic_ik_b200 import _native  # noqa: E402

_native.use_library(sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
import bench  # noqa: E402

sys.exit(bench.main())
