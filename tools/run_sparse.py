#!/usr/bin/env python
"""run_sparse.py -- BASELINE.json configs[4]: MU on a sparse CSR term-document matrix through nmfgpu_compute_single
(compressed execution, csrc/spmm.cu).  Default: 1 000 000 x 100 000 at 0.1 % density, k = 100 (GPU box only; the host
needs ~6 GB to build the matrix).  Prints iterations/s from the library's own ExecutionRecord (excludes setup, like the
reference's timing, SingleGpuDispatcher.cpp:166-218) and the wall clock of the whole call.

    python tools/run_sparse.py [--m M --n N --density D --k K --iters I]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                     # noqa: E402
from nmfgpu_b200.workloads import uniform_block  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=1000000)
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--density", type=float, default=1e-3)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--blocks", default="", help="comma separated NMFGPU_SPARSE_BLOCKS settings to compare (default: the library's choice)")
ap.add_argument("--check", action="store_true", help="explicit ||V - W H||_F over the stored entries and the W H mass (host, slow)")
a = ap.parse_args()

rng = np.random.default_rng(42)
t0 = time.time()
# per row: Poisson(density * n) distinct-ish uniform columns (SURVEY.md 8d cfg 5); duplicates are merged by the sort below
counts = rng.poisson(a.density * a.n, size=a.m).astype(np.int64)
nnz = int(counts.sum())
rows = np.repeat(np.arange(a.m, dtype=np.int64), counts)
cols = rng.integers(0, a.n, size=nnz, dtype=np.int64)
key = np.unique(rows * a.n + cols)
rows, cols = (key // a.n).astype(np.int32), (key % a.n).astype(np.int32)
nnz = len(key)
vals = (1.0 - rng.random(nnz, dtype=np.float32)).astype(np.float32)       # (0, 1]
indptr = np.zeros(a.m + 1, dtype=np.int32)
np.cumsum(np.bincount(rows, minlength=a.m), out=indptr[1:])
print("matrix %d x %d, nnz %d (density %.4g), built in %.1f s" % (a.m, a.n, nnz, nnz / (a.m * a.n), time.time() - t0), flush=True)

L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
desc = api.sparse_description(api.StorageFormat.CSR, a.m, a.n, vals, indptr, cols, 0)
W0 = uniform_block(43, a.m, a.k)
H0 = uniform_block(44, a.k, a.n)
runs = [(None, 10)] + [(b or None, a.iters) for b in (a.blocks.split(",") if a.blocks else [""])]   # the first call also pays context and kernel loading
for blocks, iters in runs:
    if blocks is not None:
        os.environ["NMFGPU_SPARSE_BLOCKS"] = blocks
    t0 = time.time()
    r = L.compute(None, a.k, W0=W0, H0=H0, iterations=iters, sparse=(desc, np.dtype(np.float32)))
    wall = time.time() - t0
    assert r["rc"] == 0, r
    print("%d iterations%s: library time %.3f s = %.2f iterations/s (%.3f ms each), wall %.2f s, frobenius %.6g"
          % (iters, " (row blocks %s)" % blocks if blocks else "", r["elapsed"], iters / max(r["elapsed"], 1e-3), 1000.0 * r["elapsed"] / iters, wall,
             r["frobenius"]), flush=True)
flops = 4.0 * nnz * a.k + 4.0 * a.k * a.k * (a.m + a.n)
print("algorithmic work per iteration: %.3g FLOP, %.3g bytes" % (flops, 2 * (8.0 * nnz + 4.0 * (a.m + 1)) + 16.0 * a.k * (a.m + a.n)))
if a.check:
    W, H = r["W"].astype(np.float64), r["H"].astype(np.float64)
    # ||V - WH||^2 = sum over entries (v - wh)^2 - (wh)^2  +  ||WH||^2, with ||WH||^2 = tr((W^T W)(H H^T))
    wh = np.einsum("ij,ji->i", W[rows], H[:, cols]) if nnz < 5e6 else None
    if wh is not None:
        total = ((vals - wh) ** 2 - wh ** 2).sum() + np.trace((W.T @ W) @ (H @ H.T))
        print("explicit residual %.6g" % np.sqrt(total))
L.finalize()
