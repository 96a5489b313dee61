#!/usr/bin/env python
"""ls_stability.py -- how far the least-squares algorithms (ALS, ACLS, GDCLS, AHCLS) land from the fp64 oracle, for this
library (explicit k x k inverse, or NMFGPU_LS_SOLVE=qr: Q^T + back substitution per right-hand side) and for the reference
build (cuSOLVER geqrf/ormqr + trsm), on a set of problems of different conditioning.  GPU box only.

    python tools/ls_stability.py            (run twice: with and without NMFGPU_LS_SOLVE=qr)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                                   # noqa: E402
from nmfgpu_b200.workloads import planted_inputs, uniform_block  # noqa: E402
from oracle import binding as orc                             # noqa: E402
from tests.test_oracle import PARAMS                          # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
REF = None
ref_so = os.path.join(ROOT, "oracle", "_ref", "libnmfgpu64_ref.so")
if os.path.exists(ref_so):
    REF = api.Library(ref_so)
    REF.set_verbosity(api.Verbosity.NoOutput)
    assert REF.initialize() == 0

problems = []
for seed in (31, 32, 33):
    problems.append(("planted rank 12, k 12, seed %d" % seed, planted_inputs(700, 450, 12, seed=seed)))
V, W0, H0 = planted_inputs(700, 450, 12, seed=31, noise=0.2)
problems.append(("planted rank 12 + noise 0.2, k 12", (V, W0, H0)))
V, _, _ = planted_inputs(700, 450, 20, seed=41, noise=0.05)
problems.append(("planted rank 20, k 8", (V, uniform_block(42, 700, 8), uniform_block(43, 8, 450))))
V, _, _ = planted_inputs(2000, 1500, 40, seed=51, noise=0.05)
problems.append(("planted rank 40, k 32, 2000x1500", (V, uniform_block(52, 2000, 32), uniform_block(53, 32, 1500))))

print("solve = %s" % os.environ.get("NMFGPU_LS_SOLVE", "inverse"))
for name, (V, W0, H0) in problems:
    k = W0.shape[1]
    for algo in ("als", "acls", "gdcls", "ahcls"):
        o = orc.run_nmf(algo, V, W0, H0, 30, params=PARAMS[algo])
        G = o["W"].T @ o["W"]
        new = L.compute(V, k, algorithm=algo, W0=W0, H0=H0, iterations=30, params=PARAMS[algo])
        line = "%-36s %-6s cond(W^T W) %.1e  ours: res %.1e W %.1e H %.1e" % (
            name, algo, np.linalg.cond(G), abs(new["frobenius"] - o["frob"][-1]) / o["frob"][-1], rel(new["W"], o["W"]), rel(new["H"], o["H"]))
        if REF is not None:
            ref = REF.compute(V, k, algorithm=algo, W0=W0, H0=H0, iterations=30, params=PARAMS[algo])
            line += "   reference: res %.1e W %.1e H %.1e" % (
                abs(ref["frobenius"] - o["frob"][-1]) / o["frob"][-1], rel(ref["W"], o["W"]), rel(ref["H"], o["H"]))
        print(line, flush=True)
L.finalize()
if REF is not None:
    REF.finalize()
