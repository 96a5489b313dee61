// spmm.h -- the NMF iteration on a SPARSE input matrix without densifying it.
//
// The reference converts every sparse input to a dense device matrix before it starts (source/common/Matrix.h:145-232:
// cusparse csr2dense / csc2dense, COO through coo2csr) and then runs its dense GEMMs, so a 1 000 000 x 100 000 term-document
// matrix at 0.1 % density (BASELINE.json configs[4]; 400 GB dense) cannot run at all.  Here the matrix stays compressed:
//   W^T V  (k x n) : one warp per COLUMN of V walks the column's non-zeros (CSC) and gathers rows of W,
//   V H^T  (m x k) : one warp per ROW of V walks the row's non-zeros (CSR) and gathers columns of H,
// both through the same kernel.  The value and index arrays are read exactly once per product, 32 entries per warp
// load (coalesced); every gathered operand row is k contiguous floats (W is kept in a row-major copy for this), so a
// gather is ceil(k/32) full 128-byte lines.  fp32 (or fp64) FMA in the order of the stored entries: deterministic.
//
// Ingestion keeps the index-base and storage-format contract of the reference (zero- or one-based CSR, CSC, COO):
// entries are expanded to coordinates on the device, stably sorted by row (CSR copy) and then by column (CSC copy,
// rows ascending inside a column) with cub radix sorts.  Out-of-range coordinates are dropped (the dense scatter of
// sparse.cu ignores them too); repeated coordinates add up.
#pragma once
#include "common.h"

namespace nmfgpu {
namespace b200 {
namespace sparse {

template <typename T>
struct DeviceSparse {
	unsigned rows = 0, cols = 0, nnz = 0;
	DeviceBuffer<int> rowPtr, colIdx;   // CSR, zero based
	DeviceBuffer<T> csrVal;
	DeviceBuffer<int> colPtr, rowIdx;   // CSC, zero based
	DeviceBuffer<T> cscVal;
};

// host CSR / CSC / COO description -> both compressed copies on the device; synchronises the stream
template <typename T>
void ingest(const MatrixDescription<T>& src, DeviceSparse<T>& dst, cudaStream_t stream);

// out[r * ldo + c] = sum over the entries e of compressed row/column r of val[e] * D[idx[e] * ldd + c],  c < k <= 128
template <typename T>
void spmmGather(unsigned numMajor, unsigned k, const int* ptr, const int* idx, const T* val, const T* D, size_t ldd, T* out, size_t ldo,
                cudaStream_t stream);

// the same over the entry ranges [ptrBegin[r], ptrEnd[r]): one block of a blocked sweep (buildBlockPointers)
template <typename T>
void spmmGather(unsigned numMajor, unsigned k, const int* ptrBegin, const int* ptrEnd, const int* idx, const T* val, const T* D, size_t ldd, T* out,
                size_t ldo, cudaStream_t stream);

// Blocked sweep (study knob NMFGPU_SPARSE_BLOCKS; measured: no gain, see engine.cu): W^T V block of rows by block of rows --
// one launch per block over all columns, restricted to the entries whose row lies in the block (they are contiguous: rows
// ascend inside a CSC column) -- so that the rows being gathered stay in L2; every block yields one partial product.  blockPtr is (blocks + 1) x numMajor: the first entry of column j at
// or after row b * minorPerBlock, the last row holding the column ends.
void buildBlockPointers(unsigned numMajor, unsigned blocks, unsigned minorPerBlock, const int* ptr, const int* idx, int* blockPtr, cudaStream_t stream);

// B[r * ldb + c] = A[c * lda + r] for r < rows, c < cols  (column-major -> row-major, or the other way round with the roles swapped)
template <typename T>
void transpose(unsigned rows, unsigned cols, const T* A, size_t lda, T* B, size_t ldb, cudaStream_t stream);

// out[r] = sum of squares of the entries of compressed row/column r  (tr(V^T V) per column, MU.h:117-125)
template <typename T>
void majorSquares(unsigned numMajor, const int* ptr, const T* val, T* out, cudaStream_t stream);

}  // namespace sparse
}  // namespace b200
}  // namespace nmfgpu
