"""GPU parity at the sizes BASELINE.json names (or their shape class), asserted -- not just printed by the bench:

  * configs[1]  100 000 x 10 000, k = 64, MU: 20 iterations through nmfgpu_compute_single from pageable host buffers, side
                by side with the reference build (oracle/_ref) from the same W0 / H0: residual and factors.
  * configs[3]  the shape CLASS of 50 000 x 20 000, k = 128 (4 064 x 3 000: k = 128 and an odd number of 32-row blocks, which
                is what the reference's k-means launch geometry is sensitive to), k-means initialisation from
                the same seed, GDCLS / AHCLS / nsNMF / MU: this library and the reference against the fp64 oracle.
  * configs[4]  a CSR input with more than 10^6 non-zeros in compressed execution against the fp64 oracle on the densified
                matrix, and ExecutionRecord.sparsityW / sparsityH against numpy.

Tolerances: MU / nsNMF residual 2e-5, factors 2e-4 (tests/test_parity_gpu.py); the least-squares family is compared with
the reference run, whose own distance from fp64 sets the scale (tools/ls_stability.py).
"""
import ctypes
import os

import numpy as np
import pytest

from nmfgpu_b200 import api
from nmfgpu_b200.api import NmfInitializationMethod, ResultType
from oracle import binding as orc
from tests.test_oracle import PARAMS
from tests.workloads import planted_inputs, uniform_block

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnmfgpu64_ref.so")


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def L():
    lib = api.Library()
    lib.set_verbosity(api.Verbosity.NoOutput)
    assert lib.initialize() == ResultType.Success
    yield lib
    lib.finalize()


@pytest.fixture(scope="module")
def REF():
    if not os.path.exists(REF_SO):
        pytest.skip("reference build oracle/_ref/libnmfgpu64_ref.so not present")
    try:
        lib = api.Library(REF_SO)
    except OSError as e:
        pytest.skip("reference build does not load: %s" % e)
    lib.set_verbosity(api.Verbosity.NoOutput)
    assert lib.initialize() == ResultType.Success
    yield lib
    lib.finalize()


def test_cfg2_residual_and_factors_match_reference(L, REF):
    """BASELINE configs[1] at full size, both libraries through nmfgpu_compute_single from the same host buffers."""
    m, n, k, iters = 100_000, 10_000, 64, 20
    dev = L.lib.nmfgpu_b200_device_alloc(m * n * 4)
    assert dev
    try:
        assert L.lib.nmfgpu_b200_device_uniform_f32(dev, m, n, m, 42, m, 0, 0) == 0
        Vt = np.empty((n, m), dtype=np.float32)                       # (n, m) C-order = (m, n) column-major
        assert L.lib.nmfgpu_b200_device_download(Vt.ctypes.data, dev, m * n * 4) == 0
    finally:
        L.lib.nmfgpu_b200_device_free(dev)
    V = Vt.T
    np.testing.assert_array_equal(V[70_000:70_004, 5000:5003], uniform_block(42, 4, 3, total_rows=m, row0=70_000, col0=5000))
    W0 = uniform_block(43, m, k)
    H0 = uniform_block(44, k, n)
    ref = REF.compute(V, k, W0=W0, H0=H0, iterations=iters)
    new = L.compute(V, k, W0=W0, H0=H0, iterations=iters)
    assert ref["rc"] == ResultType.Success and new["rc"] == ResultType.Success
    e = abs(new["frobenius"] - ref["frobenius"]) / ref["frobenius"]
    print("cfg 2 after %d iterations: ours %.6f, reference %.6f (%.1e apart)" % (iters, new["frobenius"], ref["frobenius"], e))
    assert e <= 2e-5
    # the reference's plain fp32 cuBLAS sums over 100 000 terms are the less accurate side here
    assert rel(new["W"], ref["W"].astype(np.float64)) <= 2e-4
    assert rel(new["H"], ref["H"].astype(np.float64)) <= 2e-4
    np.testing.assert_allclose((new["W"].astype(np.float64) ** 2).sum(axis=0), 1.0, rtol=1e-5)


@pytest.mark.parametrize("algo", ["mu", "gdcls", "ahcls", "nsnmf"])
def test_cfg4_shape_kmeans_init_three_way(L, REF, algo):
    """k = 128, 127 row blocks of 32 (odd: SURVEY.md B-9), KMeansAndNonNegativeWTV from the same seed, 20 iterations: this
    library and the reference build against the fp64 oracle started from the oracle's own (bit-identical) k-means
    centroids.  The data are 128 prototypes + noise so that the k x k systems of the least-squares family are well posed
    (on rank-deficient data both libraries are 20-60 % apart from each other: the solves amplify fp32 rounding by
    cond(W^T W + lambda I), tools/ls_stability.py)."""
    m, n, k = 4064, 3000, 128
    rng = np.random.default_rng(61)
    V = np.asfortranarray(rng.random((m, k)).astype(np.float32)[:, rng.integers(0, k, n)] + 0.3 * rng.random((m, n)).astype(np.float32))
    init = NmfInitializationMethod.KMeansAndNonNegativeWTV
    ref = REF.compute(V, k, algorithm=algo, init=init, iterations=20, seed=9, params=PARAMS[algo])
    new = L.compute(V, k, algorithm=algo, init=init, iterations=20, seed=9, params=PARAMS[algo])
    assert ref["rc"] == ResultType.Success and new["rc"] == ResultType.Success
    assert ref["seed"] == new["seed"]                    # the seed chain hands k-means the same seed (written back, MU.h:137)
    km = orc.run_kmeans(V, k, seed=new["seed"], maxiter=100, threshold=0.005)     # KMeansStrategy.cpp:54-58
    W0 = np.asfortranarray(km["centroids"].astype(np.float32))
    H0 = np.asfortranarray(np.maximum(0.0, W0.astype(np.float64).T @ V.astype(np.float64)).astype(np.float32))
    o = orc.run_nmf(algo, V, W0, H0, 20, params=PARAMS[algo])
    e_ref = abs(ref["frobenius"] - o["frob"][-1]) / o["frob"][-1]
    e_new = abs(new["frobenius"] - o["frob"][-1]) / o["frob"][-1]
    print("cfg 4 shape, %-5s: oracle %.6f, ours %.6f (%.1e), reference %.6f (%.1e); W %.1e vs %.1e"
          % (algo, o["frob"][-1], new["frobenius"], e_new, ref["frobenius"], e_ref, rel(new["W"], o["W"]), rel(ref["W"], o["W"])))
    tol = 2e-5 if algo in ("mu", "nsnmf") else 1e-4
    assert e_new <= max(2 * e_ref, tol), (algo, e_new, e_ref)
    assert rel(new["W"], o["W"]) <= max(2 * rel(ref["W"], o["W"]), 10 * tol), algo


def test_sparse_million_nonzeros_against_oracle(L, monkeypatch):
    """CSR input with 1.02e6 non-zeros, compressed execution (csrc/spmm.cu), against the fp64 oracle on the dense matrix"""
    import scipy.sparse as sp
    monkeypatch.setenv("NMFGPU_SPARSE", "1")
    m, n, k = 16000, 8000, 32
    S = sp.random(m, n, density=0.008, format="csr", dtype=np.float32, random_state=np.random.default_rng(21),
                  data_rvs=lambda size: (0.1 + np.random.default_rng(22).random(size)).astype(np.float32))
    assert S.nnz > 1_000_000
    W0 = uniform_block(23, m, k)
    H0 = uniform_block(24, k, n)
    desc = api.sparse_description(api.StorageFormat.CSR, m, n, S.data, S.indptr.astype(np.int32), S.indices.astype(np.int32), 0)
    r = L.compute(None, k, W0=W0, H0=H0, iterations=20, sparse=(desc, np.dtype(np.float32)))
    assert r["rc"] == ResultType.Success
    o = orc.run_nmf("mu", np.asfortranarray(S.toarray()), W0, H0, 20)
    e = abs(r["frobenius"] - o["frob"][-1]) / o["frob"][-1]
    print("sparse %d non-zeros: residual %.6f, oracle %.6f (%.1e apart)" % (S.nnz, r["frobenius"], o["frob"][-1], e))
    assert e <= 2e-5
    assert rel(r["W"], o["W"]) <= 2e-4 and rel(r["H"], o["H"]) <= 2e-4


def test_sparsity_fields_of_the_execution_record(L):
    """ExecutionRecord.sparsityW / sparsityH (declared by the reference header, never written by the reference): Hoyer
    sparseness of the factors the call returns"""
    V, W0, H0 = planted_inputs(700, 450, 12, seed=31)

    def hoyer(X):
        x = np.abs(X.astype(np.float64)).ravel()
        return (np.sqrt(x.size) - x.sum() / np.sqrt((x * x).sum())) / (np.sqrt(x.size) - 1.0)

    for algo in ("mu", "ahcls", "nsnmf"):
        r = L.compute(V, 12, algorithm=algo, W0=W0, H0=H0, iterations=30, params=PARAMS[algo])
        assert r["rc"] == ResultType.Success
        assert abs(r["sparsity_w"] - hoyer(r["W"])) <= 1e-6, algo
        assert abs(r["sparsity_h"] - hoyer(r["H"])) <= 1e-6, algo
        assert 0.0 < r["sparsity_w"] < 1.0 and 0.0 < r["sparsity_h"] < 1.0
