"""Loads libnmfgpu64.so (built in-tree by nmfgpu_b200/build.py).  There is no Python or CPU fallback:
if the CUDA library is missing this raises, and every compute entry point returns an error code when no
GPU is present."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY = os.path.join(HERE, "lib", "libnmfgpu64.so")

_lib = None


def load(path=None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("NMFGPU_LIB") or LIBRARY   # NMFGPU_LIB: load another build (A/B and bisecting tools)
    if not os.path.exists(p):
        raise RuntimeError("%s not found: build it with `python -m nmfgpu_b200.build` (nvcc, sm_100a). "
                           "nmfgpu_b200 has no CPU fallback." % p)
    lib = ctypes.CDLL(p, mode=ctypes.RTLD_LOCAL)
    if path is None:
        _lib = lib
    return lib
