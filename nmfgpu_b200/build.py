"""Builds libnmfgpu64.so (the drop-in library) in-tree with nvcc for sm_100a.

    python -m nmfgpu_b200.build [--force]

The library name is the reference's (`nmfgpu64`, source/CMakeLists.txt:78-92) so callers that locate it
through NMFGPU_ROOT keep working.  Objects are cached per source under nmfgpu_b200/lib/obj/.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBRARY = os.path.join(LIBDIR, "libnmfgpu64.so")
# diagnostic variants: NMFGPU_BUILD_EXTRA="-DNMFGPU_TC_CHECK_BUILD ..." NMFGPU_BUILD_OUT=tools/bin/lib_x.so build next to
# the product library (own object directory), and tools load them through NMFGPU_LIB
if os.environ.get("NMFGPU_BUILD_OUT"):
    LIBRARY = os.path.abspath(os.environ["NMFGPU_BUILD_OUT"])
    OBJDIR = os.path.join(LIBDIR, "obj_" + os.path.basename(LIBRARY).replace(".", "_"))

SOURCES = ["api.cpp", "host.cpp", "dist.cpp", "engine.cu", "kernels.cu", "fused.cu", "tc_gemm.cu", "kmeans.cu", "sparse.cu", "spmm.cu",
           "init_kernels.cu", "session.cu"]

# NMFGPU_TC_TRACE_BUILD=1 compiles the timeline / ablation hooks of tc_gemm.cu in (diagnostic builds only)
EXTRA = ["-DNMFGPU_TC_TRACE_BUILD"] if os.environ.get("NMFGPU_TC_TRACE_BUILD") else []
EXTRA += os.environ.get("NMFGPU_BUILD_EXTRA", "").split()
NVCC_FLAGS = EXTRA + ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas", "-DNMFGPU_EXPORTING",
              "-I", os.path.join(os.path.dirname(HERE), "include")]


def _newest_header():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src, force, header_time):
    obj = os.path.join(OBJDIR, src.replace(".", "_") + ".o")
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), header_time):
        return obj, False
    cmd = ["nvcc"] + NVCC_FLAGS + ["-x", "cu", "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJDIR, exist_ok=True)
    header_time = _newest_header()
    with ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(lambda s: _compile(s, force, header_time), SOURCES))
    objs = [o for o, _ in results]
    if force or any(c for _, c in results) or not os.path.exists(LIBRARY):
        cmd = ["nvcc", "-shared", "-o", LIBRARY] + objs + ["-lcurand", "-lcuda", "-ldl",
                                                           "-Xlinker", "-Bsymbolic", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("built", LIBRARY)
    return LIBRARY


if __name__ == "__main__":
    build(force="--force" in sys.argv)
