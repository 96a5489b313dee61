// sparse.h -- ingestion of the sparse input formats of include/nmfgpu.h (CSR / CSC / COO, zero- or
// one-based) into the engine's dense device matrix.  Contract: reference source/common/Matrix.h:145-232
// (which used the legacy cuSPARSE csr2dense/csc2dense/coo2csr calls, removed in CUDA 12).
#pragma once
#include "common.h"

namespace nmfgpu {
namespace b200 {
namespace sparse {

// zero rows [rows, ld) of every column of a column-major block
template <typename T>
void zeroPadRows(T* A, unsigned rows, unsigned cols, size_t ld, cudaStream_t stream);

// host sparse description -> dense column-major device matrix (ld >= rows); synchronises the stream
template <typename T>
void densify(const MatrixDescription<T>& src, T* dst, size_t ld, cudaStream_t stream);

}  // namespace sparse
}  // namespace b200
}  // namespace nmfgpu
