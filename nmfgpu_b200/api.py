"""ctypes mirror of include/nmfgpu.h and include/nmfgpu_b200.h.

Same names, field order and packing (`#pragma pack(4)`) as the reference header, so the parity tests read
like a C caller of the reference: fill an NmfDescription, call nmfgpu_compute_single, read the summary.
`Library` can wrap either this repo's libnmfgpu64.so or a build of the reference library -- both export
the same twelve C symbols -- which is how the tests run the two side by side.
"""
import ctypes
from ctypes import POINTER, c_bool, c_char_p, c_double, c_float, c_int, c_size_t, c_uint, c_void_p

import numpy as np

from . import _lib

# ---- enums (include/nmfgpu.h) -------------------------------------------------------------------------------


class ResultType:
    Success, ErrorAlreadyInitialized, ErrorNotInitialized, ErrorInvalidArgument, ErrorNotEnoughHostMemory, \
        ErrorNotEnoughDeviceMemory, ErrorExternalLibrary, ErrorUserInterrupt, ErrorDeviceSelection = range(9)


class NmfInitializationMethod:
    CopyExisting, AllRandomValues, MeanColumns, KMeansAndRandomValues, KMeansAndAbsoluteWTV, \
        KMeansAndNonNegativeWTV, EInNMF = range(7)


class NmfThresholdType:
    Frobenius, RMSD = range(2)


class NmfAlgorithm:
    Multiplicative, GDCLS, ALS, ACLS, AHCLS, nsNMF = range(6)


class Verbosity:
    NoOutput, Summary, Informative, Debugging = range(4)


class IndexBase:
    Zero, One = range(2)


class StorageFormat:
    Dense, CSR, CSC, COO = range(4)


ALGORITHM_BY_NAME = {"mu": 0, "gdcls": 1, "als": 2, "acls": 3, "ahcls": 4, "nsnmf": 5}

# ---- structs ------------------------------------------------------------------------------------------------------


class _Dense(ctypes.Structure):
    _pack_ = 4
    _fields_ = [("values", c_void_p), ("leadingDimension", c_uint)]


class _Sparse(ctypes.Structure):
    _pack_ = 4
    _fields_ = [("values", c_void_p), ("ptrA", c_void_p), ("ptrB", c_void_p), ("nnz", c_uint), ("base", c_int)]


class _MatrixUnion(ctypes.Union):
    _pack_ = 4
    _fields_ = [("dense", _Dense), ("csr", _Sparse), ("csc", _Sparse), ("coo", _Sparse)]


class MatrixDescription(ctypes.Structure):
    _pack_ = 4
    _anonymous_ = ("u",)
    _fields_ = [("rows", c_uint), ("columns", c_uint), ("format", c_int), ("u", _MatrixUnion)]


class Parameter(ctypes.Structure):
    _pack_ = 4
    _fields_ = [("name", c_char_p), ("value", c_double)]


UserInterruptCallback = ctypes.CFUNCTYPE(c_bool)


class NmfDescription(ctypes.Structure):
    _pack_ = 4
    _fields_ = [
        ("algorithm", c_int),
        ("useConstantBasisVectors", c_bool),
        ("inputMatrix", MatrixDescription),
        ("inputLabels", c_void_p),
        ("outputMatrixW", MatrixDescription),
        ("outputMatrixH", MatrixDescription),
        ("features", c_uint),
        ("initMethod", c_int),
        ("numIterations", c_uint),
        ("numRuns", c_uint),
        ("seed", c_uint),
        ("thresholdType", c_int),
        ("thresholdValue", c_double),
        ("callbackUserInterrupt", c_void_p),
        ("parameters", POINTER(Parameter)),
        ("numParameters", c_uint),
    ]


class KMeansDescription(ctypes.Structure):
    _pack_ = 4
    _fields_ = [
        ("inputMatrix", MatrixDescription),
        ("outputMatrixClusters", MatrixDescription),
        ("outputMemberships", c_void_p),
        ("numClusters", c_uint),
        ("numIterations", c_uint),
        ("seed", c_uint),
        ("thresholdValue", c_double),
    ]


class ExecutionRecord(ctypes.Structure):
    _pack_ = 4
    _fields_ = [("frobenius", c_double), ("rmsd", c_double), ("elapsedTime", c_double), ("sparsityW", c_double),
                ("sparsityH", c_double), ("numIterations", c_uint)]


class GpuInformation(ctypes.Structure):
    _pack_ = 4
    _fields_ = [("name", ctypes.c_char * 256), ("totalMemory", c_size_t), ("freeMemory", c_size_t)]


class NamedValue(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("value", c_double)]


class SessionInfo(ctypes.Structure):
    _fields_ = [("uses_tensor_cores", c_int), ("splits_wtv", c_uint), ("splits_vht", c_uint),
                ("kernel_launches", ctypes.c_ulonglong), ("collective_calls", ctypes.c_ulonglong),
                ("ld_v", c_size_t), ("ld_w", c_size_t), ("ld_h", c_size_t), ("row_owners", c_int)]


C_SYMBOLS = ["nmfgpu_initialize", "nmfgpu_finalize", "nmfgpu_version", "nmfgpu_set_verbosity", "nmfgpu_create_summary",
             "nmfgpu_compute_single", "nmfgpu_compute_double", "nmfgpu_compute_kmeans_single",
             "nmfgpu_compute_kmeans_double", "nmfgpu_choose_gpu", "nmfgpu_get_number_of_gpu",
             "nmfgpu_get_information_for_gpu_index"]

EXT_SYMBOLS = ["nmfgpu_b200_set_precision", "nmfgpu_b200_dist_unique_id", "nmfgpu_b200_dist_local_unique_id", "nmfgpu_b200_dist_init",
               "nmfgpu_b200_session_time_run",
               "nmfgpu_b200_dist_set_shard", "nmfgpu_b200_dist_finalize", "nmfgpu_b200_session_create_f32",
               "nmfgpu_b200_session_set_factors_f32", "nmfgpu_b200_session_get_factors_f32", "nmfgpu_b200_session_initialize",
               "nmfgpu_b200_session_iterate", "nmfgpu_b200_session_iterate_with_error",
               "nmfgpu_b200_session_time_iterations", "nmfgpu_b200_session_products_f32",
               "nmfgpu_b200_session_synchronize", "nmfgpu_b200_session_get_info", "nmfgpu_b200_session_destroy",
               "nmfgpu_b200_device_alloc", "nmfgpu_b200_device_free", "nmfgpu_b200_device_uniform_f32",
               "nmfgpu_b200_flush_l2", "nmfgpu_b200_device_download", "nmfgpu_b200_device_upload", "nmfgpu_b200_host_alloc",
               "nmfgpu_b200_host_free", "nmfgpu_b200_plan_segments"]


class SummaryHandle:
    """An ISummary* seen from C: the four virtuals are called through the Itanium vtable
    (order fixed by include/nmfgpu.h: destroy, bestRun, record, recordCount)."""

    def __init__(self, ptr):
        self.ptr = ptr
        vtable = ctypes.cast(ptr, POINTER(c_void_p))[0]
        slots = ctypes.cast(vtable, POINTER(c_void_p))
        self._destroy = ctypes.CFUNCTYPE(None, c_void_p)(slots[0])
        self._best = ctypes.CFUNCTYPE(c_uint, c_void_p)(slots[1])
        self._record = ctypes.CFUNCTYPE(None, c_void_p, c_uint, POINTER(ExecutionRecord))(slots[2])
        self._count = ctypes.CFUNCTYPE(c_uint, c_void_p)(slots[3])

    def best_run(self):
        return self._best(self.ptr)

    def record_count(self):
        return self._count(self.ptr)

    def record(self, index):
        r = ExecutionRecord()
        self._record(self.ptr, index, ctypes.byref(r))
        return r

    def destroy(self):
        if self.ptr:
            self._destroy(self.ptr)
            self.ptr = None


def _dense(arr):
    """MatrixDescription of a Fortran-ordered 2-D numpy array (no copy)."""
    assert arr.flags.f_contiguous and arr.ndim == 2
    d = MatrixDescription()
    d.rows, d.columns = arr.shape
    d.format = StorageFormat.Dense
    d.dense.values = arr.ctypes.data
    d.dense.leadingDimension = max(1, arr.strides[1] // arr.itemsize) if arr.shape[1] > 1 else arr.shape[0]
    return d


def sparse_description(fmt, rows, cols, values, ptr_a, ptr_b, base=IndexBase.Zero):
    d = MatrixDescription()
    d.rows, d.columns, d.format = rows, cols, fmt
    d.csr.values = values.ctypes.data
    d.csr.ptrA = ptr_a.ctypes.data
    d.csr.ptrB = ptr_b.ctypes.data
    d.csr.nnz = len(values)
    d.csr.base = base
    d._keep = (values, ptr_a, ptr_b)   # the description only holds raw pointers: keep the arrays alive with it
    return d


class Library:
    """One loaded libnmfgpu64-compatible shared object."""

    def __init__(self, path=None):
        self.lib = _lib.load(path)
        L = self.lib
        for name in ("nmfgpu_initialize", "nmfgpu_finalize", "nmfgpu_version", "nmfgpu_get_number_of_gpu"):
            getattr(L, name).restype = c_int
        L.nmfgpu_get_number_of_gpu.restype = c_uint
        L.nmfgpu_set_verbosity.argtypes = [c_int]
        L.nmfgpu_set_verbosity.restype = None
        L.nmfgpu_create_summary.argtypes = [POINTER(c_void_p)]
        L.nmfgpu_compute_single.argtypes = [POINTER(NmfDescription), c_void_p]
        L.nmfgpu_compute_double.argtypes = [POINTER(NmfDescription), c_void_p]
        L.nmfgpu_compute_kmeans_single.argtypes = [POINTER(KMeansDescription)]
        L.nmfgpu_compute_kmeans_double.argtypes = [POINTER(KMeansDescription)]
        L.nmfgpu_choose_gpu.argtypes = [c_uint]
        L.nmfgpu_get_information_for_gpu_index.argtypes = [c_uint, POINTER(GpuInformation)]
        self.has_extensions = hasattr(L, "nmfgpu_b200_session_create_f32")
        if self.has_extensions:
            L.nmfgpu_b200_device_alloc.restype = c_void_p
            L.nmfgpu_b200_device_alloc.argtypes = [c_size_t]
            L.nmfgpu_b200_device_free.argtypes = [c_void_p]
            L.nmfgpu_b200_device_free.restype = None
            L.nmfgpu_b200_device_uniform_f32.argtypes = [c_void_p, c_uint, c_uint, c_size_t, ctypes.c_ulonglong,
                                                         ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong]
            L.nmfgpu_b200_session_create_f32.argtypes = [c_int, c_uint, c_uint, c_uint, c_void_p, c_uint, c_int, c_int,
                                                         POINTER(NamedValue), c_uint, POINTER(c_void_p)]
            L.nmfgpu_b200_session_set_factors_f32.argtypes = [c_void_p, c_void_p, c_uint, c_void_p, c_uint]
            L.nmfgpu_b200_session_get_factors_f32.argtypes = [c_void_p, c_void_p, c_uint, c_void_p, c_uint]
            L.nmfgpu_b200_session_initialize.argtypes = [c_void_p, c_int, c_uint, POINTER(c_float)]
            L.nmfgpu_b200_session_iterate.argtypes = [c_void_p, c_uint]
            L.nmfgpu_b200_session_iterate_with_error.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double)]
            L.nmfgpu_b200_session_time_iterations.argtypes = [c_void_p, c_uint, POINTER(c_float)]
            L.nmfgpu_b200_session_time_run.argtypes = [c_void_p, c_uint, POINTER(c_float), POINTER(c_double)]
            L.nmfgpu_b200_dist_local_unique_id.argtypes = [c_void_p]
            L.nmfgpu_b200_session_products_f32.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_float), POINTER(c_float)]
            L.nmfgpu_b200_session_synchronize.argtypes = [c_void_p]
            L.nmfgpu_b200_session_get_info.argtypes = [c_void_p, POINTER(SessionInfo)]
            L.nmfgpu_b200_session_destroy.argtypes = [c_void_p]
            L.nmfgpu_b200_session_destroy.restype = None
            L.nmfgpu_b200_dist_unique_id.argtypes = [c_void_p]
            L.nmfgpu_b200_dist_init.argtypes = [c_int, c_int, c_void_p]
            L.nmfgpu_b200_dist_set_shard.argtypes = [c_uint, c_uint]
            L.nmfgpu_b200_set_precision.argtypes = [c_int]
            L.nmfgpu_b200_device_download.argtypes = [c_void_p, c_void_p, c_size_t]
            L.nmfgpu_b200_device_upload.argtypes = [c_void_p, c_void_p, c_size_t]
            L.nmfgpu_b200_host_alloc.restype = c_void_p
            L.nmfgpu_b200_host_alloc.argtypes = [c_size_t]
            L.nmfgpu_b200_host_free.argtypes = [c_void_p]
            L.nmfgpu_b200_host_free.restype = None

    # -- reference API ------------------------------------------------------------------------------------
    def initialize(self):
        return self.lib.nmfgpu_initialize()

    def finalize(self):
        return self.lib.nmfgpu_finalize()

    def version(self):
        return self.lib.nmfgpu_version()

    def set_verbosity(self, v):
        self.lib.nmfgpu_set_verbosity(v)

    def choose_gpu(self, index):
        return self.lib.nmfgpu_choose_gpu(index)

    def number_of_gpu(self):
        return self.lib.nmfgpu_get_number_of_gpu()

    def gpu_information(self, index):
        info = GpuInformation()
        rc = self.lib.nmfgpu_get_information_for_gpu_index(index, ctypes.byref(info))
        return rc, info

    def create_summary(self):
        p = c_void_p()
        rc = self.lib.nmfgpu_create_summary(ctypes.byref(p))
        if rc != ResultType.Success:
            raise RuntimeError("nmfgpu_create_summary -> %d" % rc)
        return SummaryHandle(p.value)

    def compute(self, V, features, algorithm=NmfAlgorithm.Multiplicative, W0=None, H0=None,
                init=NmfInitializationMethod.CopyExisting, iterations=100, runs=1, seed=0,
                threshold_type=NmfThresholdType.Frobenius, threshold_value=0.0, params=None, constant_w=False,
                callback=None, sparse=None):
        """Fill an NmfDescription and call nmfgpu_compute_single / _double (by V's dtype).

        V: (m, n) float32/float64 array (copied to column-major), or None with `sparse` = a ready
        MatrixDescription plus dtype via W0.  Returns dict(rc, W, H, frobenius, rmsd, iterations, elapsed,
        seed, runs, record_count).
        """
        if isinstance(algorithm, str):
            algorithm = ALGORITHM_BY_NAME[algorithm]
        if sparse is not None:
            in_desc, dtype = sparse
            m, n = in_desc.rows, in_desc.columns
        else:
            V = np.asfortranarray(V)
            dtype = V.dtype
            m, n = V.shape
            in_desc = _dense(V)
        assert dtype in (np.float32, np.float64)
        k = features
        W = np.zeros((m, k), dtype=dtype, order="F")
        H = np.zeros((k, n), dtype=dtype, order="F")
        if W0 is not None:
            W[:, :] = W0
        if H0 is not None:
            H[:, :] = H0
        d = NmfDescription()
        d.algorithm = algorithm
        d.useConstantBasisVectors = bool(constant_w)
        d.inputMatrix = in_desc
        d.inputLabels = None
        d.outputMatrixW = _dense(W)
        d.outputMatrixH = _dense(H)
        d.features = k
        d.initMethod = init
        d.numIterations = iterations
        d.numRuns = runs
        d.seed = seed
        d.thresholdType = threshold_type
        d.thresholdValue = threshold_value
        cb = UserInterruptCallback(callback) if callback is not None else None
        d.callbackUserInterrupt = ctypes.cast(cb, c_void_p).value if cb is not None else None
        params = params or {}
        arr = (Parameter * max(1, len(params)))()
        keep = []
        for i, (name, value) in enumerate(params.items()):
            b = name.encode()
            keep.append(b)
            arr[i].name = b
            arr[i].value = float(value)
        d.parameters = arr
        d.numParameters = len(params)
        summary = self.create_summary()
        fn = self.lib.nmfgpu_compute_single if dtype == np.float32 else self.lib.nmfgpu_compute_double
        rc = fn(ctypes.byref(d), summary.ptr)
        out = dict(rc=rc, W=W, H=H, seed=d.seed, runs=d.numRuns, record_count=summary.record_count())
        if rc == ResultType.Success and summary.record_count() > 0:
            r = summary.record(summary.best_run())
            out.update(frobenius=r.frobenius, rmsd=r.rmsd, iterations=r.numIterations, elapsed=r.elapsedTime,
                       sparsity_w=r.sparsityW, sparsity_h=r.sparsityH, best_run=summary.best_run())
        summary.destroy()
        return out

    def compute_kmeans(self, X, clusters, iterations=100, seed=0, threshold=0.0):
        X = np.asfortranarray(X)
        m, n = X.shape
        C = np.zeros((m, clusters), dtype=X.dtype, order="F")
        memb = np.zeros(n, dtype=np.uint32)
        d = KMeansDescription()
        d.inputMatrix = _dense(X)
        d.outputMatrixClusters = _dense(C)
        d.outputMemberships = memb.ctypes.data
        d.numClusters = clusters
        d.numIterations = iterations
        d.seed = seed
        d.thresholdValue = threshold
        fn = self.lib.nmfgpu_compute_kmeans_single if X.dtype == np.float32 else self.lib.nmfgpu_compute_kmeans_double
        rc = fn(ctypes.byref(d))
        return dict(rc=rc, centroids=C, memberships=memb)

    # -- extensions (include/nmfgpu_b200.h) --------------------------------------------------------------------
    def set_precision(self, mode):
        return self.lib.nmfgpu_b200_set_precision({"auto": 0, "fp32": 1, "3xtf32": 2, "tf32": 3}.get(mode, mode))


class Session:
    """A factorisation whose V stays resident in HBM (include/nmfgpu_b200.h sessions)."""

    def __init__(self, library, algorithm, m, n, k, V=None, device_ptr=None, ld_v=None, constant_w=False, params=None):
        self.L = library
        self.m, self.n, self.k = m, n, k
        if isinstance(algorithm, str):
            algorithm = ALGORITHM_BY_NAME[algorithm]
        params = params or {}
        arr = (NamedValue * max(1, len(params)))()
        self._keep = []
        for i, (name, value) in enumerate(params.items()):
            b = name.encode()
            self._keep.append(b)
            arr[i].name = b
            arr[i].value = float(value)
        h = c_void_p()
        if device_ptr is not None:
            rc = library.lib.nmfgpu_b200_session_create_f32(algorithm, m, n, k, device_ptr, ld_v or m, 1, int(constant_w), arr,
                                                            len(params), ctypes.byref(h))
        else:
            V = np.asfortranarray(V, dtype=np.float32)
            self._V = V
            rc = library.lib.nmfgpu_b200_session_create_f32(algorithm, m, n, k, V.ctypes.data, V.strides[1] // 4 if n > 1 else m, 0,
                                                            int(constant_w), arr, len(params), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("nmfgpu_b200_session_create_f32 -> %d" % rc)
        self.h = h

    def set_factors(self, W, H):
        W = np.asfortranarray(W, dtype=np.float32)
        H = np.asfortranarray(H, dtype=np.float32)
        rc = self.L.lib.nmfgpu_b200_session_set_factors_f32(self.h, W.ctypes.data, self.m, H.ctypes.data, self.k)
        if rc != 0:
            raise RuntimeError("set_factors -> %d" % rc)

    def initialize(self, init_method, seed):
        """initial factors by one of the reference's strategies on the resident V; returns the milliseconds it took"""
        ms = c_float()
        rc = self.L.lib.nmfgpu_b200_session_initialize(self.h, init_method, seed, ctypes.byref(ms))
        if rc != 0:
            raise RuntimeError("session_initialize -> %d" % rc)
        return ms.value

    def get_factors(self):
        W = np.zeros((self.m, self.k), dtype=np.float32, order="F")
        H = np.zeros((self.k, self.n), dtype=np.float32, order="F")
        rc = self.L.lib.nmfgpu_b200_session_get_factors_f32(self.h, W.ctypes.data, self.m, H.ctypes.data, self.k)
        if rc != 0:
            raise RuntimeError("get_factors -> %d" % rc)
        return W, H

    def iterate(self, iterations):
        rc = self.L.lib.nmfgpu_b200_session_iterate(self.h, iterations)
        if rc != 0:
            raise RuntimeError("iterate -> %d" % rc)

    def iterate_with_error(self):
        f, r = c_double(), c_double()
        rc = self.L.lib.nmfgpu_b200_session_iterate_with_error(self.h, ctypes.byref(f), ctypes.byref(r))
        if rc != 0:
            raise RuntimeError("iterate_with_error -> %d" % rc)
        return f.value, r.value

    def time_iterations(self, iterations):
        ms = c_float()
        rc = self.L.lib.nmfgpu_b200_session_time_iterations(self.h, iterations, ctypes.byref(ms))
        if rc != 0:
            raise RuntimeError("time_iterations -> %d" % rc)
        return ms.value

    def time_run(self, iterations):
        """Milliseconds of `iterations` iterations with the reference's residual cadence, and the last residual."""
        ms, f = c_float(), c_double()
        rc = self.L.lib.nmfgpu_b200_session_time_run(self.h, iterations, ctypes.byref(ms), ctypes.byref(f))
        if rc != 0:
            raise RuntimeError("time_run -> %d" % rc)
        return ms.value, f.value

    def products(self, want_wtv=True, want_vht=True):
        wtv = np.zeros((self.k, self.n), dtype=np.float32, order="F") if want_wtv else None
        vht = np.zeros((self.m, self.k), dtype=np.float32, order="F") if want_vht else None
        a, b = c_float(), c_float()
        rc = self.L.lib.nmfgpu_b200_session_products_f32(self.h, wtv.ctypes.data if want_wtv else None,
                                                         vht.ctypes.data if want_vht else None, ctypes.byref(a), ctypes.byref(b))
        if rc != 0:
            raise RuntimeError("products -> %d" % rc)
        return wtv, vht, a.value, b.value

    def synchronize(self):
        self.L.lib.nmfgpu_b200_session_synchronize(self.h)

    def info(self):
        i = SessionInfo()
        self.L.lib.nmfgpu_b200_session_get_info(self.h, ctypes.byref(i))
        return i

    def close(self):
        if self.h:
            self.L.lib.nmfgpu_b200_session_destroy(self.h)
            self.h = None
