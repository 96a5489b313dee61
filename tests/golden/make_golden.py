"""Generates tests/golden/nmf_traces.json from the numpy restatement (tests/np_restatement.py).

The reference has no golden vectors and cannot run without a GPU, so these fixtures pin the C++ oracle
to an independently written fp64 restatement.  Run from the repo root: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import np_restatement as npr  # noqa: E402
from tests.test_oracle import PARAMS  # noqa: E402
from tests.workloads import planted_inputs  # noqa: E402

out = {}
for algo in PARAMS:
    m, n, k, seed, iters = 257, 131, 6, 21, 40
    V, W0, H0 = planted_inputs(m, n, k, seed=seed)
    W, H, checks = npr.run(algo, V, W0, H0, iters, float(np.finfo(np.float32).eps), PARAMS[algo])
    out[algo] = dict(m=m, n=n, k=k, seed=seed, iterations=iters,
                     frob=[c[1] for c in checks], frob_explicit=[c[2] for c in checks],
                     w_colsum=np.abs(W).sum(axis=0).tolist(), h_rowsum=H.sum(axis=1).tolist())
with open(os.path.join(os.path.dirname(__file__), "nmf_traces.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out), "traces")
