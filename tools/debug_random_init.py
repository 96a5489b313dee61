import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api
from nmfgpu_b200.api import NmfInitializationMethod
from nmfgpu_b200.workloads import planted_inputs
from oracle import binding as orc
from tests.test_oracle import PARAMS

L = api.Library(); L.set_verbosity(0); assert L.initialize() == 0
REF = api.Library(os.path.join(ROOT, "oracle", "_ref", "libnmfgpu64_ref.so")); REF.set_verbosity(0); assert REF.initialize() == 0
V, _, _ = planted_inputs(640, 320, 8, seed=13)
for it in (1, 2, 5, 20):
    ref = REF.compute(V, 8, init=NmfInitializationMethod.AllRandomValues, iterations=it, seed=77)
    out = [ref["frobenius"]]
    for mode in ("auto", "fp32"):
        L.set_precision(mode)
        new = L.compute(V, 8, init=NmfInitializationMethod.AllRandomValues, iterations=it, seed=77)
        out.append(new["frobenius"])
        out.append(float(np.linalg.norm(new["W"] - ref["W"]) / np.linalg.norm(ref["W"])))
        out.append(float(np.linalg.norm(new["H"] - ref["H"]) / np.linalg.norm(ref["H"])))
    print("random init it=%d ref %.6f | auto %.6f dW %.2e dH %.2e | fp32 %.6f dW %.2e dH %.2e" % (it, *out))
# ALS / ACLS
V, W0, H0 = planted_inputs(700, 450, 12, seed=31)
for algo in ("als", "acls", "gdcls"):
    o = orc.run_nmf(algo, V, W0, H0, 30, params=PARAMS[algo])
    for it, idx in ((10, 0), (30, 2)):
        ref = REF.compute(V, 12, algorithm=algo, W0=W0, H0=H0, iterations=it, params=PARAMS[algo])
        row = [o["frob"][idx], ref["frobenius"]]
        for mode in ("auto", "fp32"):
            L.set_precision(mode)
            new = L.compute(V, 12, algorithm=algo, W0=W0, H0=H0, iterations=it, params=PARAMS[algo])
            row.append(new["frobenius"])
        print("%s it=%d oracle %.6f ref %.6f auto %.6f fp32 %.6f" % (algo, it, *row))
