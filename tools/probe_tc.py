#!/usr/bin/env python
"""probe_tc.py -- accuracy and timing of the two tensor-core products against fp64 numpy, swept over the
accumulator flush interval (NMFGPU_TC_FLUSH_STAGES) and the precision mode.  GPU box only (gpurun).

    python tools/probe_tc.py [small|big|all]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                     # noqa: E402
from nmfgpu_b200.workloads import uniform_block  # noqa: E402


def stats(x, ref):
    x = np.asarray(x, dtype=np.float64)
    d = x - ref
    return {"rel_fro": float(np.linalg.norm(d) / np.linalg.norm(ref)), "max_rel": float(np.max(np.abs(d) / np.abs(ref).clip(1e-30))),
            "bias": float(np.mean(d / ref.clip(1e-30)))}


def reference(V, W, H, block=500):
    m, n = V.shape
    wtv = np.empty((W.shape[1], n))
    vht = np.zeros((m, H.shape[0]))
    W64 = W.astype(np.float64)
    for c0 in range(0, n, block):
        Vb = V[:, c0:c0 + block].astype(np.float64)
        wtv[:, c0:c0 + block] = W64.T @ Vb
        vht += Vb @ H[:, c0:c0 + block].astype(np.float64).T
    return wtv, vht


def run(L, m, n, k, modes, flushes, reps=3):
    t0 = time.time()
    V = np.empty((m, n), dtype=np.float32, order="F")
    for c0 in range(0, n, 500):
        c1 = min(n, c0 + 500)
        V[:, c0:c1] = uniform_block(42, m, c1 - c0, total_rows=m, col0=c0)
    W = uniform_block(43, m, k)
    H = uniform_block(44, k, n)
    wtv_ref, vht_ref = reference(V, W, H)
    print("# %dx%d k=%d: inputs + fp64 reference in %.1f s" % (m, n, k, time.time() - t0), flush=True)
    for mode in modes:
        for F in (flushes if mode != "fp32" else [0]):
            os.environ["NMFGPU_TC_FLUSH_STAGES"] = str(F)
            L.set_precision(mode)
            try:
                s = api.Session(L, "mu", m, n, k, V=V)
                s.set_factors(W, H)
                wtv, vht, a, b = s.products()
                ta, tb = [], []
                for _ in range(reps):
                    _, _, x, y = s.products(want_wtv=False, want_vht=False)
                    ta.append(x)
                    tb.append(y)
                info = s.info()
                rec = {"shape": [m, n, k], "mode": mode, "flush": F, "tc": int(info.uses_tensor_cores), "slots": [info.splits_wtv, info.splits_vht],
                       "wtv": stats(wtv, wtv_ref), "vht": stats(vht, vht_ref), "ms_wtv": min(ta), "ms_vht": min(tb)}
                s.close()
            except Exception as e:  # noqa: BLE001
                rec = {"shape": [m, n, k], "mode": mode, "flush": F, "error": str(e)}
            print(json.dumps(rec), flush=True)
    L.set_precision("auto")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    L = api.Library()
    L.set_verbosity(api.Verbosity.NoOutput)
    assert L.initialize() == 0
    if which in ("small", "all"):
        run(L, 256, 256, 16, ["3xtf32"], [8])
        run(L, 1000, 500, 10, ["3xtf32", "fp32"], [8])
        run(L, 4096, 2048, 64, ["3xtf32", "tf32", "fp32"], [0, 8])
        run(L, 3000, 1700, 100, ["3xtf32"], [8])
    if which in ("big", "all"):
        run(L, 100000, 10000, 64, ["3xtf32", "fp32"], [0, 32, 8, 4, 2, 1])
    L.finalize()


if __name__ == "__main__":
    main()
