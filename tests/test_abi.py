"""CPU tests of the drop-in boundary: struct layouts, enum values, exported symbols,
and that the C-ABI library loads and exports every symbol include/*.h declares.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

from nmfgpu_b200 import api
from nmfgpu_b200 import build as nbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library_path():
    return nbuild.build(verbose=False)


def test_struct_layouts_match_reference_pack4():
    assert ctypes.sizeof(api.MatrixDescription) == 44
    assert ctypes.sizeof(api.NmfDescription) == 200
    assert ctypes.sizeof(api.KMeansDescription) == 116
    assert ctypes.sizeof(api.ExecutionRecord) == 44
    assert ctypes.sizeof(api.Parameter) == 16
    assert ctypes.sizeof(api.GpuInformation) == 272
    off = {n: getattr(api.NmfDescription, n).offset for n, _ in api.NmfDescription._fields_}
    assert off == dict(algorithm=0, useConstantBasisVectors=4, inputMatrix=8, inputLabels=52, outputMatrixW=60,
                       outputMatrixH=104, features=148, initMethod=152, numIterations=156, numRuns=160, seed=164,
                       thresholdType=168, thresholdValue=172, callbackUserInterrupt=180, parameters=188, numParameters=196)
    off = {n: getattr(api.KMeansDescription, n).offset for n, _ in api.KMeansDescription._fields_}
    assert off == dict(inputMatrix=0, outputMatrixClusters=44, outputMemberships=88, numClusters=96, numIterations=100,
                       seed=104, thresholdValue=108)
    assert api.MatrixDescription.u.offset == 12
    assert api._Sparse.nnz.offset == 24 and api._Sparse.base.offset == 28


def test_cxx_layouts_match_ctypes(tmp_path):
    """Compile a probe against include/nmfgpu.h with g++ and compare sizeof/offsetof with the ctypes mirror."""
    src = tmp_path / "probe.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstddef>
#include "nmfgpu.h"
using namespace nmfgpu;
int main() {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(MatrixDescription<float>), sizeof(NmfDescription<float>), sizeof(NmfDescription<double>),
         sizeof(KMeansDescription<double>), sizeof(ExecutionRecord), sizeof(Parameter), sizeof(GpuInformation));
  printf("%zu %zu %zu %zu\n", offsetof(NmfDescription<float>, thresholdValue), offsetof(NmfDescription<float>, parameters),
         offsetof(KMeansDescription<float>, thresholdValue), offsetof(ExecutionRecord, numIterations));
  printf("%d %d %d %d %d\n", (int)ResultType::ErrorDeviceSelection, (int)NmfInitializationMethod::EInNMF, (int)NmfAlgorithm::nsNMF,
         (int)StorageFormat::COO, (int)Verbosity::Debugging);
  return 0;
}''')
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0].split() == ["44", "200", "200", "116", "44", "16", "272"]
    assert out[1].split() == ["172", "188", "108", "40"]
    assert out[2].split() == ["8", "6", "5", "3", "3"]


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(nmfgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(library_path):
    lib = ctypes.CDLL(library_path)
    declared = [s for s in _declared("nmfgpu.h")] + [s for s in _declared("nmfgpu_b200.h")]
    assert len(declared) >= 12 + 15
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(s for s in _declared("nmfgpu.h")) == sorted(api.C_SYMBOLS)
    # the C++ entry points of the reference header are exported under their Itanium names
    names = subprocess.run(["nm", "-D", "--defined-only", library_path], capture_output=True, text=True, check=True).stdout
    for mangled in ["_ZN6nmfgpu10initializeEv", "_ZN6nmfgpu8finalizeEv", "_ZN6nmfgpu7versionEv", "_ZN6nmfgpu9chooseGpuEj",
                    "_ZN6nmfgpu14getNumberOfGpuEv", "_ZN6nmfgpu25getInformationForGpuIndexEjRNS_14GpuInformationE",
                    "_ZN6nmfgpu12setVerbosityENS_9VerbosityE", "_ZN6nmfgpu7computeERNS_14NmfDescriptionIfEEPNS_8ISummaryE",
                    "_ZN6nmfgpu7computeERNS_14NmfDescriptionIdEEPNS_8ISummaryE",
                    "_ZN6nmfgpu13computeKMeansERNS_17KMeansDescriptionIfEEPNS_13KMeansSummaryE",
                    "_ZN6nmfgpu13computeKMeansERNS_17KMeansDescriptionIdEEPNS_13KMeansSummaryE", "_ZN6nmfgpu8ISummary6createEv"]:
        assert mangled in names, mangled


REFERENCE_EXAMPLE = "/root/reference/example/main.cpp"


@pytest.mark.skipif(not os.path.exists(REFERENCE_EXAMPLE), reason="the reference tree is only mounted in the build container")
def test_reference_example_builds_against_this_header_and_library(tmp_path, library_path):
    """Source-level drop-in: the reference's own example caller (example/main.cpp, C++ API: nmfgpu::initialize / chooseGpu /
    compute, ISummary through its vtable), compiled UNMODIFIED where it lies against include/nmfgpu.h and linked with this
    libnmfgpu64.so.  Without a GPU the program must come back with the library's error, not crash."""
    exe = str(tmp_path / "reference_example")
    libdir = os.path.dirname(library_path)
    subprocess.run(["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), REFERENCE_EXAMPLE, "-L", libdir, "-lnmfgpu64",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], input="\n", capture_output=True, text=True, timeout=300)
    assert r.returncode >= 0, "killed by signal %d" % -r.returncode      # ran to its own exit
    if not _has_gpu():
        assert "no CPU fallback" in (r.stdout + r.stderr)


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).returncode == 0
    except FileNotFoundError:
        return False


def test_product_sources_never_touch_the_oracle():
    """The product path may not include, link or call anything under oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "nmfgpu_b200")
    for base, _, files in os.walk(pkg):
        if "lib" in base.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"import\s+oracle|from\s+oracle|#include\s+[\"<][^\n]*oracle|liboracle|oracle\.binding|oracle/", text), os.path.join(base, f)


def test_library_contains_the_blackwell_instructions(library_path):
    """The built libnmfgpu64.so carries sm_100a code whose V-sized products are tcgen05 / TMA kernels: UTCHMMA
    (tcgen05.mma), UTMALDG (TMA tensor load), LDTM / STTM (tcgen05.ld / st) in the SASS of tc_stream_gemm, and no cuBLAS
    among its dependencies.  (What profiles/r02_sass_counts.txt records, checked on every build.)"""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", library_path], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    blocks = sass.split("Function : ")
    gemm = [b for b in blocks if "tc_stream_gemm" in b.split("\n", 1)[0]]
    assert len(gemm) == 4, [b.split("\n", 1)[0] for b in gemm]            # <64|128> x <W^T V | V H^T>
    for b in gemm:
        for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR"):
            assert mnemonic in b, (b.split("\n", 1)[0], mnemonic)
    needed = subprocess.run(["readelf", "-d", library_path], capture_output=True, text=True).stdout
    assert "cublas" not in needed.lower() and "cusolver" not in needed.lower() and "cusparse" not in needed.lower()
