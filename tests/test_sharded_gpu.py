"""The sharded dataflows of csrc/engine.cu on the GPU, against the single-GPU engine.

The ranks are THREADS of this process (nmfgpu_b200_dist_local_unique_id, csrc/dist.h), all on cuda:0, so the test runs on a
single-GPU box.  What executes is the product code: the same engine, the same kernels and the same exchange protocol as
with one process per GPU -- peer stores into the other ranks' exchange buffers, epoch flags, last-block signals, in-kernel
waits.  Two things differ: the setup-time collectives go through memcpy instead of NCCL, and because the ranks share one GPU
they meet on the host (stream synchronise + barrier) before every kernel that waits for another rank, and no CUDA graphs are
used: a kernel spinning on the shared GPU can keep the rank it waits for from running at all (csrc/dist.h
ranksMayShareDevice, csrc/engine.cu m_hostLockstep).  The waits still execute and find their flags set -- a flag that is
never raised would hit the 10 s watchdog and fail the test.  Waits that really wait are what tools/dist_check.py and
bench.py --gpus N exercise with one process per GPU.
"""
import ctypes
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nmfgpu_b200 import api                                      # noqa: E402
from nmfgpu_b200.workloads import dense_inputs, shard_columns    # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

# factors and residual of a sharded run against the single-GPU run of the same problem: the arithmetic is the same, only
# the summation order of the partial products differs
TOL_FACTOR, TOL_RESIDUAL = 1e-5, 5e-6   # measured: 8e-7 / 3e-8


def _single(L, algorithm, V, W0, H0, iters, params=None):
    s = api.Session(L, algorithm, V.shape[0], V.shape[1], W0.shape[1], V=V, params=params)
    s.set_factors(W0, H0)
    s.iterate(iters)
    f, _ = s.iterate_with_error()
    W, H = s.get_factors()
    info = s.info()
    s.close()
    return W, H, f, info


def _run_ranks(world, body):
    """Runs body(rank, L, sync) on `world` threads, each with its own library context and a local communicator."""
    L0 = api.Library()
    uid = (ctypes.c_ubyte * 128)()
    assert L0.lib.nmfgpu_b200_dist_local_unique_id(uid) == 0
    sync = threading.Barrier(world, timeout=300)
    results, errors = [None] * world, []

    def worker(rank):
        L = api.Library()
        try:
            L.set_verbosity(api.Verbosity.NoOutput)
            assert L.initialize() == 0
            assert L.choose_gpu(0) == 0
            assert L.lib.nmfgpu_b200_dist_init(rank, world, uid) == 0
            results[rank] = body(rank, L, sync)
        except BaseException as e:  # noqa: BLE001
            errors.append((rank, repr(e)))
            sync.abort()
        finally:
            L.lib.nmfgpu_b200_dist_finalize()
            L.finalize()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return results


@pytest.fixture(scope="module")
def lib():
    L = api.Library()
    L.set_verbosity(api.Verbosity.NoOutput)
    assert L.initialize() == 0
    yield L
    L.finalize()


@pytest.mark.parametrize("world,m,n,k,iters", [(2, 6000, 4096, 32, 20), (3, 2500, 1000, 10, 30), (2, 20000, 4096, 64, 12), (4, 5000, 1500, 100, 10)])
@pytest.mark.parametrize("mode", ["rowblocks", "allreduce"])
def test_mu_shards_match_single_gpu(lib, world, m, n, k, iters, mode, monkeypatch):
    if mode == "allreduce":
        monkeypatch.setenv("NMFGPU_DIST_MODE", "allreduce")
    else:
        monkeypatch.delenv("NMFGPU_DIST_MODE", raising=False)
    V, W0, H0 = dense_inputs(m, n, k, seed=5)
    W1, H1, f1, info1 = _single(lib, "mu", V, W0, H0, iters)
    assert info1.uses_tensor_cores

    def body(rank, L, sync):
        c0, c1 = shard_columns(n, world, rank)
        assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
        s = api.Session(L, "mu", m, c1 - c0, k, V=np.asfortranarray(V[:, c0:c1]))
        s.set_factors(W0, np.asfortranarray(H0[:, c0:c1]))
        info = s.info()
        sync.wait()
        s.iterate(iters)
        f, _ = s.iterate_with_error()
        sync.wait()
        W, H = s.get_factors()
        sync.wait()
        s.close()
        return W, H, f, bool(info.row_owners), (c0, c1)

    for W2, H2, f2, row_blocks, (c0, c1) in _run_ranks(world, body):
        assert row_blocks == (mode == "rowblocks")
        eW = np.linalg.norm(W2 - W1) / np.linalg.norm(W1)
        eH = np.linalg.norm(H2 - H1[:, c0:c1]) / np.linalg.norm(H1[:, c0:c1])
        ef = abs(f2 - f1) / f1
        assert eW <= TOL_FACTOR and eH <= TOL_FACTOR and ef <= TOL_RESIDUAL, (eW, eH, ef)


def test_unequal_shards_and_constant_w(lib):
    """Column shards of different widths (the regrouping into row blocks takes any partition) and a constant W."""
    m, n, k, iters = 3000, 900, 16, 15
    V, W0, H0 = dense_inputs(m, n, k, seed=9)
    cuts = [0, 100, 640, 900]

    def single(constant):
        s = api.Session(lib, "mu", m, n, k, V=V, constant_w=constant)
        s.set_factors(W0, H0)
        s.iterate(iters)
        f, _ = s.iterate_with_error()
        W, H = s.get_factors()
        s.close()
        return W, H, f

    for constant in (False, True):
        W1, H1, f1 = single(constant)
        if constant:
            np.testing.assert_allclose(W1, W0, rtol=0, atol=0)   # a constant W is neither updated nor normalised (MU.h:218-228)

        def body(rank, L, sync):
            c0, c1 = cuts[rank], cuts[rank + 1]
            assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
            s = api.Session(L, "mu", m, c1 - c0, k, V=np.asfortranarray(V[:, c0:c1]), constant_w=constant)
            s.set_factors(W0, np.asfortranarray(H0[:, c0:c1]))
            sync.wait()
            s.iterate(iters)
            f, _ = s.iterate_with_error()
            sync.wait()
            W, H = s.get_factors()
            sync.wait()
            s.close()
            return W, H, f

        for rank, (W2, H2, f2) in enumerate(_run_ranks(3, body)):
            c0, c1 = cuts[rank], cuts[rank + 1]
            assert np.linalg.norm(W2 - W1) / np.linalg.norm(W1) <= TOL_FACTOR
            assert np.linalg.norm(H2 - H1[:, c0:c1]) / np.linalg.norm(H1[:, c0:c1]) <= TOL_FACTOR
            assert abs(f2 - f1) / f1 <= TOL_RESIDUAL


def test_reference_api_over_shards(lib):
    """nmfgpu_compute_single on column shards (the call bench.py's e2e leg makes on N GPUs): same factors as one GPU."""
    m, n, k, iters = 4000, 2048, 24, 30
    V, W0, H0 = dense_inputs(m, n, k, seed=3)
    one = lib.compute(V, k, W0=W0, H0=H0, iterations=iters)
    assert one["rc"] == 0

    def body(rank, L, sync):
        c0, c1 = shard_columns(n, 2, rank)
        assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
        r = L.compute(np.asfortranarray(V[:, c0:c1]), k, W0=W0, H0=np.asfortranarray(H0[:, c0:c1]), iterations=iters)
        return r, (c0, c1)

    for r, (c0, c1) in _run_ranks(2, body):
        assert r["rc"] == 0
        assert np.linalg.norm(r["W"] - one["W"]) / np.linalg.norm(one["W"]) <= TOL_FACTOR
        assert np.linalg.norm(r["H"] - one["H"][:, c0:c1]) / np.linalg.norm(one["H"][:, c0:c1]) <= TOL_FACTOR
        assert abs(r["frobenius"] - one["frobenius"]) / one["frobenius"] <= TOL_RESIDUAL
        assert r["iterations"] == one["iterations"] == iters


@pytest.mark.parametrize("algorithm,params", [("gdcls", {"lambda": 0.01}), ("acls", {"lambdaW": 0.01, "lambdaH": 0.01}), ("nsnmf", {"theta": 0.3})])
def test_other_algorithms_over_shards(lib, algorithm, params):
    """The all-reduce dataflow (m x k partial V H^T, k x k partial H H^T) carries every algorithm besides MU."""
    m, n, k, iters = 3000, 1024, 12, 10
    V, W0, H0 = dense_inputs(m, n, k, seed=11)
    W1, H1, f1, _ = _single(lib, algorithm, V, W0, H0, iters, params)

    def body(rank, L, sync):
        c0, c1 = shard_columns(n, 2, rank)
        assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
        s = api.Session(L, algorithm, m, c1 - c0, k, V=np.asfortranarray(V[:, c0:c1]), params=params)
        s.set_factors(W0, np.asfortranarray(H0[:, c0:c1]))
        sync.wait()
        s.iterate(iters)
        f, _ = s.iterate_with_error()
        sync.wait()
        W, H = s.get_factors()
        sync.wait()
        s.close()
        return W, H, f, (c0, c1)

    for W2, H2, f2, (c0, c1) in _run_ranks(2, body):
        assert np.linalg.norm(W2 - W1) / np.linalg.norm(W1) <= 2e-4
        assert np.linalg.norm(H2 - H1[:, c0:c1]) / np.linalg.norm(H1[:, c0:c1]) <= 2e-4
        assert abs(f2 - f1) / f1 <= 2e-5


@pytest.mark.parametrize("world,cuts", [(2, None), (3, [0, 130, 131, 500])])
def test_kmeans_over_shards_is_bit_exact(lib, world, cuts):
    """nmfgpu_compute_kmeans_single on column shards (csrc/kmeans.cu: global Forgy seeds, local assignments, centroid sums
    chained through the ranks in sample order): memberships and centroids are the bits of the single-GPU run"""
    rng = np.random.default_rng(17)
    m, n, k = 1000, 500, 6              # 32 row blocks; (m = 1030: 33 blocks, the reference's odd-block quirk B-9) below
    for rows in (m, 1030):
        X = (rng.random((rows, k)).astype(np.float32)[:, rng.integers(0, k, n)] + 0.3 * rng.random((rows, n)).astype(np.float32))
        one = lib.compute_kmeans(X, k, iterations=40, seed=4, threshold=0.0)
        assert one["rc"] == 0

        def body(rank, L, sync):
            c0, c1 = (cuts[rank], cuts[rank + 1]) if cuts else shard_columns(n, world, rank)
            assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
            r = L.compute_kmeans(np.asfortranarray(X[:, c0:c1]), k, iterations=40, seed=4, threshold=0.0)
            return r, (c0, c1)

        for r, (c0, c1) in _run_ranks(world, body):
            assert r["rc"] == 0
            np.testing.assert_array_equal(r["memberships"], one["memberships"][c0:c1])
            np.testing.assert_array_equal(r["centroids"], one["centroids"])


@pytest.mark.parametrize("algorithm,params", [("mu", {}), ("gdcls", {"lambda": 0.01}), ("ahcls", {"lambdaW": 0.01, "lambdaH": 0.01, "alphaW": 0.01, "alphaH": 0.01})])
def test_kmeans_initialisation_over_shards(lib, algorithm, params):
    """BASELINE configs[3] in small: k-means initialisation (KMeansAndRandomValues) of a column-sharded problem gives the
    factorisation the single GPU gives from the same seed"""
    m, n, k, iters = 2048, 1200, 16, 20
    V, _, _ = dense_inputs(m, n, k, seed=21)
    init = api.NmfInitializationMethod.KMeansAndRandomValues
    one = lib.compute(V, k, algorithm=algorithm, init=init, iterations=iters, seed=6, params=params)
    assert one["rc"] == 0

    def body(rank, L, sync):
        c0, c1 = shard_columns(n, 2, rank)
        assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
        r = L.compute(np.asfortranarray(V[:, c0:c1]), k, algorithm=algorithm, init=init, iterations=iters, seed=6, params=params)
        return r, (c0, c1)

    for r, (c0, c1) in _run_ranks(2, body):
        assert r["rc"] == 0
        assert np.linalg.norm(r["W"] - one["W"]) / np.linalg.norm(one["W"]) <= 2e-4
        assert np.linalg.norm(r["H"] - one["H"][:, c0:c1]) / np.linalg.norm(one["H"][:, c0:c1]) <= 2e-4
        assert abs(r["frobenius"] - one["frobenius"]) / one["frobenius"] <= 2e-5
