"""CPU tests of the host logic behind the C ABI: return codes, validation order and struct write-backs
(reference source/common/Interface.cpp:53-77,214-347,359-421).  Without a GPU a well-formed compute call
must fail loudly (ErrorDeviceSelection / ErrorExternalLibrary), never fall back to a CPU path."""
import numpy as np
import pytest

from nmfgpu_b200 import api
from nmfgpu_b200 import build as nbuild
from nmfgpu_b200.api import NmfAlgorithm, NmfInitializationMethod, ResultType
from tests.workloads import planted_inputs, shard_columns, uniform_block


def _gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def L():
    nbuild.build(verbose=False)
    lib = api.Library()
    lib.set_verbosity(api.Verbosity.NoOutput)
    return lib


def test_version_and_lifecycle(L):
    assert L.version() == (0 << 24) | (2 << 16) | 3 == 131075
    V, W0, H0 = planted_inputs(40, 30, 3)
    assert L.compute(V, 3, W0=W0, H0=H0, iterations=5)["rc"] == ResultType.ErrorNotInitialized
    assert L.compute_kmeans(V, 3)["rc"] == ResultType.ErrorNotInitialized
    assert L.finalize() == ResultType.ErrorNotInitialized
    assert L.initialize() == ResultType.Success
    assert L.initialize() == ResultType.ErrorAlreadyInitialized
    assert L.finalize() == ResultType.Success
    assert L.finalize() == ResultType.ErrorNotInitialized


def test_argument_validation_precedes_device_work(L):
    assert L.initialize() == ResultType.Success
    try:
        V, W0, H0 = planted_inputs(40, 30, 3)
        # features > columns (Interface.cpp:228-232)
        assert L.compute(V, 31, iterations=5, init=NmfInitializationMethod.AllRandomValues)["rc"] == ResultType.ErrorInvalidArgument
        # required named parameters (Interface.cpp:246-328)
        for algo, params in [(NmfAlgorithm.GDCLS, {}), (NmfAlgorithm.ACLS, {"lambdaW": 0.1}),
                             (NmfAlgorithm.AHCLS, {"lambdaW": 0.1, "lambdaH": 0.1, "alphaW": 0.1}), (NmfAlgorithm.nsNMF, {"lambda": 1.0})]:
            r = L.compute(V, 3, algorithm=algo, W0=W0, H0=H0, iterations=5, params=params)
            assert r["rc"] == ResultType.ErrorInvalidArgument, algo
        assert L.compute(V, 3, algorithm=17, W0=W0, H0=H0, iterations=5)["rc"] == ResultType.ErrorInvalidArgument
        # CopyExisting with several runs is reduced to one run, written back to the caller's struct (Interface.cpp:221-225)
        r = L.compute(V, 3, W0=W0, H0=H0, iterations=5, runs=4)
        assert r["runs"] == 1
        # k-means validation (Interface.cpp:365-389)
        assert L.compute_kmeans(V, 0)["rc"] == ResultType.ErrorInvalidArgument
        assert L.compute_kmeans(V, 30)["rc"] == ResultType.ErrorInvalidArgument
        assert L.lib.nmfgpu_create_summary(None) == ResultType.ErrorInvalidArgument
        assert L.lib.nmfgpu_get_information_for_gpu_index(0, None) == ResultType.ErrorInvalidArgument
    finally:
        L.finalize()


@pytest.mark.skipif(_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu(L):
    assert L.initialize() == ResultType.Success
    try:
        V, W0, H0 = planted_inputs(40, 30, 3)
        r = L.compute(V, 3, W0=W0, H0=H0, iterations=5)
        assert r["rc"] in (ResultType.ErrorDeviceSelection, ResultType.ErrorExternalLibrary)
        np.testing.assert_array_equal(r["W"], W0)  # untouched: nothing was computed anywhere
        assert L.compute_kmeans(V, 3)["rc"] in (ResultType.ErrorDeviceSelection, ResultType.ErrorExternalLibrary)
        assert L.number_of_gpu() == 0
        assert L.choose_gpu(0) == ResultType.ErrorDeviceSelection
    finally:
        L.finalize()


def test_summary_vtable_roundtrip(L):
    s = L.create_summary()
    assert s.record_count() == 0 and s.best_run() == 0
    s.destroy()


def test_workload_generator_blocks_are_consistent():
    full = uniform_block(42, 64, 48)
    assert full.dtype == np.float32 and full.min() > 0 and full.max() <= 1
    c0, c1 = shard_columns(48, 4, 2)
    np.testing.assert_array_equal(uniform_block(42, 64, c1 - c0, total_rows=64, col0=c0), full[:, c0:c1])
    np.testing.assert_array_equal(uniform_block(42, 16, 48, total_rows=64, row0=32), full[32:48, :])
    spans = [shard_columns(10, 4, r) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
