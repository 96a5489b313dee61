#!/usr/bin/env python
"""run_products.py -- the two V-sized products (and optionally whole MU iterations) on the bench workload with
V generated on the device; the command profiled with ncu (see profiles/).  GPU box only.

    python tools/run_products.py [--m M --n N --k K] [--reps R] [--iters I] [--mode auto|fp32|3xtf32|tf32]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                     # noqa: E402
from nmfgpu_b200.workloads import uniform_block  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=100000)
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--k", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--iters", type=int, default=0)
ap.add_argument("--mode", default="auto")
ap.add_argument("--algo", default="mu")
ap.add_argument("--lib", default=None, help="alternative libnmfgpu64.so (A/B comparisons)")
a = ap.parse_args()

L = api.Library(a.lib)
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
L.set_precision(a.mode)
dev = L.lib.nmfgpu_b200_device_alloc(a.m * a.n * 4)
assert dev
assert L.lib.nmfgpu_b200_device_uniform_f32(dev, a.m, a.n, a.m, 42, a.m, 0, 0) == 0
PARAMS = {"mu": {}, "gdcls": {"lambda": 0.01}, "als": {}, "acls": {"lambdaW": 0.01, "lambdaH": 0.01},
          "ahcls": {"lambdaW": 0.01, "lambdaH": 0.01, "alphaW": 0.01, "alphaH": 0.01}, "nsnmf": {"theta": 0.5}}
s = api.Session(L, a.algo, a.m, a.n, a.k, device_ptr=dev, ld_v=a.m, params=PARAMS[a.algo])
s.set_factors(uniform_block(43, a.m, a.k), uniform_block(44, a.k, a.n))
for _ in range(a.reps):
    _, _, t1, t2 = s.products(want_wtv=False, want_vht=False)
    print("W^T V %.4f ms   V H^T %.4f ms" % (t1, t2), flush=True)
if a.iters:
    s.iterate(3)
    s.synchronize()
    ms = s.time_iterations(a.iters)
    print("%d iterations: %.4f ms each, %.1f it/s" % (a.iters, ms / a.iters, 1000.0 * a.iters / ms))
    print("residual after: %r" % (s.iterate_with_error(),))
i = s.info()
print("tensor cores %d, slots %d/%d, launches %d" % (i.uses_tensor_cores, i.splits_wtv, i.splits_vht, i.kernel_launches))
s.close()
L.lib.nmfgpu_b200_device_free(dev)
L.finalize()
