"""bench.py on an A/B build of the library (mcm_b200.build.build_variant): python tools/bench_with_lib.py <lib.so> [bench args]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcm_b200 import _lib  # noqa: E402

_lib.use_library(os.path.abspath(sys.argv[1]))
sys.argv = ["bench.py"] + sys.argv[2:]
import bench  # noqa: E402

bench.main()
