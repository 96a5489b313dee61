"""ctypes binding of oracle/nmf_oracle.cpp (the CPU checker).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(nmfgpu_b200/) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRCS = [os.path.join(_HERE, "nmf_oracle.cpp"), os.path.join(_HERE, "kmeans_oracle.cpp")]

ALGORITHMS = {"mu": 0, "gdcls": 1, "als": 2, "acls": 3, "ahcls": 4, "nsnmf": 5}


def build(force=False):
    """Compile the oracle with g++ -fopenmp (a few seconds).  Building the checker is not using it."""
    srcs = [s for s in _SRCS if os.path.exists(s)]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", _SO] + srcs
    subprocess.run(cmd, check=True)
    return _SO


class OracleConfig(ctypes.Structure):
    _fields_ = [
        ("algorithm", ctypes.c_int),
        ("m", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int),
        ("num_iterations", ctypes.c_int),
        ("use_constant_w", ctypes.c_int),
        ("threshold_type", ctypes.c_int),
        ("threshold_value", ctypes.c_double),
        ("eps", ctypes.c_double),
        ("lambda_", ctypes.c_double),
        ("lambdaW", ctypes.c_double), ("lambdaH", ctypes.c_double),
        ("alphaW", ctypes.c_double), ("alphaH", ctypes.c_double),
        ("theta", ctypes.c_double),
        ("v_is_float", ctypes.c_int),
        ("explicit_residual", ctypes.c_int),
        ("num_threads", ctypes.c_int),
    ]


class OracleTrace(ctypes.Structure):
    _fields_ = [
        ("capacity", ctypes.c_int),
        ("num_checks", ctypes.c_int),
        ("iterations_done", ctypes.c_int),
        ("iteration", ctypes.POINTER(ctypes.c_int)),
        ("frob_reported", ctypes.POINTER(ctypes.c_double)),
        ("frob_explicit", ctypes.POINTER(ctypes.c_double)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_nmf_run.restype = ctypes.c_int
        _lib.oracle_nmf_run.argtypes = [ctypes.POINTER(OracleConfig), ctypes.c_void_p, ctypes.c_long,
                                        ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                                        ctypes.POINTER(OracleTrace)]
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_mu_iterations.restype = ctypes.c_int
        _lib.oracle_mu_iterations.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_long,
                                              ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                                              ctypes.c_int, ctypes.c_double]
        if hasattr(_lib, "oracle_kmeans_f32"):
            _lib.oracle_kmeans_f32.restype = ctypes.c_int
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def run_nmf(algorithm, V, W0, H0, iterations, eps=None, use_constant_w=False, threshold_type=0,
            threshold_value=0.0, params=None, explicit_residual=True, threads=0):
    """Run the oracle from W0/H0 (CopyExisting semantics).

    V: (m, n) array, float32 or float64, any memory order (copied to column-major).
    W0: (m, k), H0: (k, n).  Returns dict(W, H, iteration, frob, frob_explicit, iterations_done).
    eps defaults to the machine epsilon of V's dtype, as in the reference.
    """
    params = params or {}
    V = np.asfortranarray(V)
    assert V.dtype in (np.float32, np.float64)
    m, n = V.shape
    k = W0.shape[1]
    assert W0.shape == (m, k) and H0.shape == (k, n)
    W = np.asfortranarray(W0, dtype=np.float64).copy(order="F")
    H = np.asfortranarray(H0, dtype=np.float64).copy(order="F")
    if eps is None:
        eps = float(np.finfo(V.dtype).eps)
    cfg = OracleConfig()
    cfg.algorithm = ALGORITHMS[algorithm] if isinstance(algorithm, str) else int(algorithm)
    cfg.m, cfg.n, cfg.k = m, n, k
    cfg.num_iterations = iterations
    cfg.use_constant_w = int(use_constant_w)
    cfg.threshold_type = threshold_type
    cfg.threshold_value = threshold_value
    cfg.eps = eps
    cfg.lambda_ = params.get("lambda", 0.0)
    cfg.lambdaW = params.get("lambdaW", 0.0)
    cfg.lambdaH = params.get("lambdaH", 0.0)
    cfg.alphaW = params.get("alphaW", 0.0)
    cfg.alphaH = params.get("alphaH", 0.0)
    cfg.theta = params.get("theta", 0.0)
    cfg.v_is_float = int(V.dtype == np.float32)
    cfg.explicit_residual = int(explicit_residual)
    cfg.num_threads = threads
    cap = iterations // 10 + 2
    it = np.zeros(cap, dtype=np.int32)
    fr = np.zeros(cap, dtype=np.float64)
    fe = np.zeros(cap, dtype=np.float64)
    tr = OracleTrace()
    tr.capacity = cap
    tr.iteration = it.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    tr.frob_reported = fr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    tr.frob_explicit = fe.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    rc = lib().oracle_nmf_run(ctypes.byref(cfg), V.ctypes.data, V.strides[1] // V.itemsize,
                              W.ctypes.data, m, H.ctypes.data, k, ctypes.byref(tr))
    if rc != 0:
        raise RuntimeError("oracle_nmf_run failed: %d" % rc)
    c = tr.num_checks
    return dict(W=W, H=H, iteration=it[:c].copy(), frob=fr[:c].copy(), frob_explicit=fe[:c].copy(),
                iterations_done=tr.iterations_done)


def time_mu_iterations(V32, W0, H0, iters):
    """cpu_baseline leg: run `iters` MU iterations on fp32 V; returns seconds."""
    import time
    V32 = np.asfortranarray(V32, dtype=np.float32)
    m, n = V32.shape
    k = W0.shape[1]
    W = np.asfortranarray(W0, dtype=np.float64).copy(order="F")
    H = np.asfortranarray(H0, dtype=np.float64).copy(order="F")
    t0 = time.perf_counter()
    rc = lib().oracle_mu_iterations(m, n, k, V32.ctypes.data, V32.strides[1] // 4, W.ctypes.data, m,
                                    H.ctypes.data, k, iters, float(np.finfo(np.float32).eps))
    t1 = time.perf_counter()
    if rc != 0:
        raise RuntimeError("oracle_mu_iterations failed")
    return t1 - t0


def run_kmeans(X, k, seed=0, maxiter=100, threshold=0.0, reference_row_coverage=True):
    """CPU k-means with the reference's arithmetic order.  X: (m, n) float32/float64."""
    X = np.asfortranarray(X)
    m, n = X.shape
    C = np.zeros((m, k), dtype=X.dtype, order="F")
    memb = np.zeros(n, dtype=np.uint32)
    rounds = ctypes.c_uint(0)
    fn = lib().oracle_kmeans_f32 if X.dtype == np.float32 else lib().oracle_kmeans_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                   ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_uint)]
    rc = fn(m, n, k, X.ctypes.data, X.strides[1] // X.itemsize if n > 1 else m, C.ctypes.data, m, memb.ctypes.data, seed, maxiter,
            threshold, int(reference_row_coverage), ctypes.byref(rounds))
    if rc != 0:
        raise RuntimeError("oracle_kmeans failed: %d" % rc)
    return dict(centroids=C, memberships=memb, rounds=rounds.value)
