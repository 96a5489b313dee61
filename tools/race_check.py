#!/usr/bin/env python
"""race_check.py -- the products are deterministic by construction (static stream-K slots, no atomics): repeat them
and compare every result bitwise with the first one.  Any difference is a synchronisation bug.  GPU box only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                      # noqa: E402
from nmfgpu_b200.workloads import dense_inputs   # noqa: E402

L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (m, n, k) in [(100000, 2000, 64), (20000, 5000, 64), (20000, 6000, 128)]:
    V, W0, H0 = dense_inputs(m, n, k, seed=11)
    bad = 0
    worst = 0.0
    first = None
    for r in range(reps):
        s = api.Session(L, "mu", m, n, k, V=V)
        s.set_factors(W0, H0)
        wtv, vht, t1, t2 = s.products()
        times = (t1, t2)
        s.close()
        if np.isnan(wtv).any() or np.isnan(vht).any():
            rr, cc = np.nonzero(np.isnan(wtv))
            print("  rep %d: NaN in wtv: %d elements, cols %s; in vht: %d" % (r, len(rr), sorted(set(cc))[:12], int(np.isnan(vht).sum())), flush=True)
        if first is None:
            first = (wtv.copy(), vht.copy())
            ref = (W0.astype(np.float64).T @ V.astype(np.float64), V.astype(np.float64) @ H0.astype(np.float64).T)
            print("%s first run rel err %.2e %.2e" % ((m, n, k), np.linalg.norm(wtv - ref[0]) / np.linalg.norm(ref[0]),
                                                     np.linalg.norm(vht - ref[1]) / np.linalg.norm(ref[1])), flush=True)
        else:
            for a, b, name in ((wtv, first[0], "wtv"), (vht, first[1], "vht")):
                if not np.array_equal(a, b):
                    bad += 1
                    d = np.abs(a.astype(np.float64) - b)
                    idx = np.unravel_index(np.argmax(d), d.shape)
                    worst = max(worst, float(d.max() / np.abs(b).max()))
                    rr, cc = np.nonzero(a != b)
                    print("  rep %d %s differs: %d elements in %d rows x %d cols (cols %s), max at %s (%.6g vs %.6g)"
                          % (r, name, len(rr), len(set(rr)), len(set(cc)), sorted(set(cc))[:8], idx, a[idx], b[idx]), flush=True)
    print("%s: %d mismatching results in %d repetitions, worst relative difference %.2e  (last: %.3f / %.3f ms)"
          % ((m, n, k), bad, reps, worst, times[0], times[1]), flush=True)
L.finalize()
