#!/usr/bin/env python
"""e2e_pageable.py -- wall clock of nmfgpu_compute_single (20 iterations, cfg 2) from PAGEABLE host memory for several
settings of the staged upload (NMFGPU_UPLOAD_THREADS), with NMFGPU_TIMING phase marks for the default.  GPU box only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                     # noqa: E402
from nmfgpu_b200.workloads import uniform_block  # noqa: E402

M, N, K = 100000, 10000, 64
L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
nbytes = M * N * 4
tmp = L.lib.nmfgpu_b200_device_alloc(nbytes)
assert L.lib.nmfgpu_b200_device_uniform_f32(tmp, M, N, M, 42, M, 0, 0) == 0
Vt = np.empty((N, M), dtype=np.float32)
assert L.lib.nmfgpu_b200_device_download(Vt.ctypes.data, tmp, nbytes) == 0
L.lib.nmfgpu_b200_device_free(tmp)
V = Vt.T
W0 = uniform_block(43, M, K)
H0 = uniform_block(44, K, N)
L.compute(V, K, W0=W0, H0=H0, iterations=2)
print("hardware threads:", os.cpu_count())
for threads in ("", "4", "8", "12", "16", "24", "32"):
    if threads:
        os.environ["NMFGPU_UPLOAD_THREADS"] = threads
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        r = L.compute(V, K, W0=W0, H0=H0, iterations=20)
        best = min(best, time.perf_counter() - t0)
    print("upload threads %-8s: %.1f ms = %.1f it/s" % (threads or "default", best * 1e3, 20 / best), flush=True)
os.environ.pop("NMFGPU_UPLOAD_THREADS")
os.environ["NMFGPU_TIMING"] = "1"
L.compute(V, K, W0=W0, H0=H0, iterations=20)
L.finalize()
