#!/usr/bin/env python
"""e2e_timing.py -- where the wall clock of one nmfgpu_compute_single call goes (bench.py's e2e leg), with
NMFGPU_TIMING=1 phase marks of the library on stderr.  GPU box only."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                     # noqa: E402
from nmfgpu_b200.workloads import uniform_block  # noqa: E402

M, N, K = 100000, 10000, 64
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
nbytes = M * N * 4
hp = L.lib.nmfgpu_b200_host_alloc(nbytes)
Vh = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_float)), shape=(N, M)).T
tmp = L.lib.nmfgpu_b200_device_alloc(nbytes)
assert L.lib.nmfgpu_b200_device_uniform_f32(tmp, M, N, M, 42, M, 0, 0) == 0
assert L.lib.nmfgpu_b200_device_download(hp, tmp, nbytes) == 0
L.lib.nmfgpu_b200_device_free(tmp)
W0 = uniform_block(43, M, K)
H0 = uniform_block(44, K, N)
L.compute(Vh, K, W0=W0, H0=H0, iterations=2)
for timing in (False, True):
    if timing:
        os.environ["NMFGPU_TIMING"] = "1"
    t0 = time.perf_counter()
    r = L.compute(Vh, K, W0=W0, H0=H0, iterations=iters)
    wall = time.perf_counter() - t0
    print("%d iterations: wall %.1f ms (%.1f it/s), library loop %.1f ms, timing marks %s"
          % (iters, wall * 1e3, iters / wall, r["elapsed"] * 1e3, "on" if timing else "off"), flush=True)
L.finalize()
