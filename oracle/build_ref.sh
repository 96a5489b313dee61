#!/usr/bin/env bash
# build_ref.sh -- compile the UNMODIFIED reference (razorx89/nmfgpu) from the sources where they lie
# under /root/reference into oracle/_ref/libnmfgpu64_ref.so (git-ignored, travels to the GPU box).
# Test infrastructure: the result is the parity checker and the `--impl reference` bench arm.
# The reference's own CMake (FindCUDA, sm_13..sm_35) is not used; the file list is
# source/CMakeLists.txt:40-62.  No reference file is copied or edited: the three incompatibilities
# with CUDA 12.9 / gcc 13 are bridged by the force-included oracle/ref_compat.h.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NMFGPU_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/source" ]; then
	echo "build_ref.sh: $REF not present (GPU box) -- using prebuilt $OUT if any" >&2
	exit 0
fi
mkdir -p "$OUT/obj"
SRCS="common/Event.cpp common/Interface.cpp common/Logging.cpp common/Matrix.cpp common/Stream.cpp common/Wrapper.cpp
init/CopyStrategy.cpp init/EInNMF.cu init/InitializationStrategy.cpp init/KernelMeanColumn.cu init/KMeansStrategy.cpp
init/MeanColumnStrategy.cpp init/RandomValueStrategy.cpp kmeans/kMeans.cu nmf/Algorithm.cpp nmf/FrobeniusResolver.cpp
nmf/KernelFillMatrix.cu nmf/KernelMakeNonNegative.cu nmf/KernelMultiplyDivide.cu nmf/KernelNormalizeColumns.cu
nmf/KernelTraceMultiplication.cu nmf/SingleGpuDispatcher.cpp nmf/Summary.cpp"
FLAGS="-std=c++14 -O2 -w -Xcompiler -fPIC -gencode arch=compute_100,code=sm_100 -include $HERE/ref_compat.h
-I $REF/include -I $REF/source -DHAVE_CUBLAS -DNMFGPU_EXPORTING"
OBJS=""
pids=""
for s in $SRCS; do
	o="$OUT/obj/$(echo "$s" | tr '/.' '__').o"
	OBJS="$OBJS $o"
	if [ ! -f "$o" ] || [ "$REF/source/$s" -nt "$o" ]; then
		nvcc $FLAGS -x cu -c "$REF/source/$s" -o "$o" &
		pids="$pids $!"
	fi
done
for p in $pids; do wait "$p"; done
nvcc -shared -o "$OUT/libnmfgpu64_ref.so" $OBJS -lcublas -lcurand -lcusparse -lcusolver -lgomp
echo "built $OUT/libnmfgpu64_ref.so"
