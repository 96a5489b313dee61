#!/usr/bin/env python
"""In-stream durations of the steps of an iteration (NMFGPU_PROFILE_ITERATION: CUDA events between the steps, no graphs).

    python tools/profile_iteration.py [m n k [iterations]]        (default: BASELINE configs[1], 40 iterations)
"""
import os
import sys

os.environ["NMFGPU_PROFILE_ITERATION"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                                   # noqa: E402
from nmfgpu_b200.workloads import uniform_block               # noqa: E402

m, n, k = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (100_000, 10_000, 64)
iters = int(sys.argv[4]) if len(sys.argv) >= 5 else 40
L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
dev = L.lib.nmfgpu_b200_device_alloc(m * n * 4)
assert L.lib.nmfgpu_b200_device_uniform_f32(dev, m, n, m, 42, m, 0, 0) == 0
s = api.Session(L, "mu", m, n, k, device_ptr=dev, ld_v=m)
s.set_factors(uniform_block(43, m, k), uniform_block(44, k, n, total_rows=k))
s.iterate(iters)
s.synchronize()
s.close()          # the report is printed when the engine is destroyed
L.lib.nmfgpu_b200_device_free(dev)
L.finalize()
