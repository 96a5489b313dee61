#!/usr/bin/env python
"""baseline_numbers.py -- the numbers BASELINE.md sections 2 and 4 ask for, measured on the GPU box (one JSON object on stdout):
TF32 and fp32-FFMA dense matmul peaks (torch.matmul 8192^3, cuBLAS: library calls, used only as roofline denominators), the
reference build (oracle/_ref) on cfg 1 and the fp64 OpenMP oracle on cfg 1.  cfg 2 of both comes from bench.py.

    python tools/baseline_numbers.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                           # noqa: E402
from nmfgpu_b200.workloads import cfg1_inputs         # noqa: E402
from oracle import binding as orc                     # noqa: E402

out = {}
n = 8192
a = torch.rand(n, n, device="cuda", dtype=torch.float32)
b = torch.rand(n, n, device="cuda", dtype=torch.float32)
for name, tf32 in (("tf32_tflops", True), ("fp32_ffma_tflops", False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for _ in range(3):
        torch.matmul(a, b)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_burst"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if tf32 else 30
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    out[name + "_sustained"] = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
del a, b

V, W0, H0 = cfg1_inputs()
flops1 = 4.0 * 1000 * 500 * 10 + 4.0 * 100 * 1500
ref_so = os.path.join(ROOT, "oracle", "_ref", "libnmfgpu64_ref.so")
for label, path in (("reference_cfg1", ref_so), ("ours_cfg1", None)):
    if path is not None and not os.path.exists(path):
        continue
    L = api.Library(path)
    L.set_verbosity(api.Verbosity.NoOutput)
    assert L.initialize() == 0
    L.compute(V, 10, W0=W0, H0=H0, iterations=100)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        r = L.compute(V, 10, W0=W0, H0=H0, iterations=100)
    wall = (time.perf_counter() - t0) / reps
    out[label] = {"iterations_per_s_whole_call": 100.0 / wall, "frobenius": r["frobenius"], "effective_gflops": flops1 * 100.0 / wall / 1e9}
    L.finalize()
t0 = time.perf_counter()
o = orc.run_nmf("mu", V, W0, H0, 100)
wall = time.perf_counter() - t0
out["oracle_cfg1"] = {"iterations_per_s": 100.0 / wall, "cores": orc.num_threads(), "frobenius": float(o["frob"][-1]), "effective_gflops": flops1 * 100.0 / wall / 1e9}
print(json.dumps(out))
