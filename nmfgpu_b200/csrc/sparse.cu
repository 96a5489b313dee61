// sparse.cu -- see sparse.h.  Scatter kernels: the value / index arrays are read once, coalesced; the
// dense target is zero-filled first.  Duplicated coordinates accumulate (atomicAdd), as a sum of
// entries is the conventional meaning of a repeated COO coordinate.
#include "sparse.h"

namespace nmfgpu {
namespace b200 {
namespace sparse {

namespace {
template <typename T>
__global__ void zero_pad_kernel(T* A, unsigned rows, unsigned cols, size_t ld) {
	const unsigned pad = (unsigned)(ld - rows);
	const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (size_t)pad * cols) return;
	const unsigned c = (unsigned)(idx / pad), r = rows + (unsigned)(idx % pad);
	A[(size_t)c * ld + r] = T(0);
}

// one warp per compressed row/column: `major` indexes ptr, `minor` comes from idx
template <typename T, bool RowCompressed>
__global__ void scatter_compressed(unsigned numMajor, const int* __restrict__ ptr, const int* __restrict__ idx, const T* __restrict__ val,
                                   int base, unsigned rows, unsigned cols, T* __restrict__ dst, size_t ld) {
	const unsigned major = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (major >= numMajor) return;
	const int begin = ptr[major] - base, end = ptr[major + 1] - base;
	for (int e = begin + (int)lane; e < end; e += 32) {
		const unsigned minor = (unsigned)(idx[e] - base);
		const unsigned r = RowCompressed ? major : minor;
		const unsigned c = RowCompressed ? minor : major;
		if (r < rows && c < cols) atomicAdd(&dst[(size_t)c * ld + r], val[e]);
	}
}

template <typename T>
__global__ void scatter_coo(unsigned nnz, const int* __restrict__ rowIdx, const int* __restrict__ colIdx, const T* __restrict__ val, int base,
                            unsigned rows, unsigned cols, T* __restrict__ dst, size_t ld) {
	const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nnz) return;
	const unsigned r = (unsigned)(rowIdx[e] - base), c = (unsigned)(colIdx[e] - base);
	if (r < rows && c < cols) atomicAdd(&dst[(size_t)c * ld + r], val[e]);
}
}  // namespace

template <typename T>
void zeroPadRows(T* A, unsigned rows, unsigned cols, size_t ld, cudaStream_t stream) {
	if (ld <= rows) return;
	const size_t total = (ld - rows) * cols;
	zero_pad_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(A, rows, cols, ld);
	CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void densify(const MatrixDescription<T>& src, T* dst, size_t ld, cudaStream_t stream) {
	const unsigned rows = src.rows, cols = src.columns;
	CUDA_CHECK(cudaMemsetAsync(dst, 0, ld * cols * sizeof(T), stream));
	// the three sparse members of the union share one shape (values, ptrA, ptrB, nnz, base)
	const unsigned nnz = src.csr.nnz;
	const int base = src.csr.base == IndexBase::One ? 1 : 0;
	if (nnz == 0) {
		CUDA_CHECK(cudaStreamSynchronize(stream));
		return;
	}
	if (src.csr.values == nullptr || src.csr.rowPtr == nullptr || src.csr.columnIndices == nullptr)
		throw EngineError(ResultType::ErrorInvalidArgument, "sparse matrix with null arrays");
	DeviceBuffer<T> val;
	DeviceBuffer<int> a, b;
	val.allocate(nnz);
	CUDA_CHECK(cudaMemcpyAsync(val.get(), src.csr.values, nnz * sizeof(T), cudaMemcpyHostToDevice, stream));
	if (src.format == StorageFormat::CSR || src.format == StorageFormat::CSC) {
		const bool csr = src.format == StorageFormat::CSR;
		const unsigned numMajor = csr ? rows : cols;
		a.allocate(numMajor + 1);
		b.allocate(nnz);
		CUDA_CHECK(cudaMemcpyAsync(a.get(), csr ? src.csr.rowPtr : src.csc.columnPtr, (numMajor + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
		CUDA_CHECK(cudaMemcpyAsync(b.get(), csr ? src.csr.columnIndices : src.csc.rowIndices, nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
		if (csr) scatter_compressed<T, true><<<ceilDiv(numMajor, 8), 256, 0, stream>>>(numMajor, a.get(), b.get(), val.get(), base, rows, cols, dst, ld);
		else scatter_compressed<T, false><<<ceilDiv(numMajor, 8), 256, 0, stream>>>(numMajor, a.get(), b.get(), val.get(), base, rows, cols, dst, ld);
	} else if (src.format == StorageFormat::COO) {
		a.allocate(nnz);
		b.allocate(nnz);
		CUDA_CHECK(cudaMemcpyAsync(a.get(), src.coo.rowIndices, nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
		CUDA_CHECK(cudaMemcpyAsync(b.get(), src.coo.columnIndices, nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
		scatter_coo<T><<<ceilDiv(nnz, 256), 256, 0, stream>>>(nnz, a.get(), b.get(), val.get(), base, rows, cols, dst, ld);
	} else {
		throw EngineError(ResultType::ErrorInvalidArgument, "unknown storage format");
	}
	CUDA_CHECK(cudaGetLastError());
	CUDA_CHECK(cudaStreamSynchronize(stream));
}

template void zeroPadRows<float>(float*, unsigned, unsigned, size_t, cudaStream_t);
template void zeroPadRows<double>(double*, unsigned, unsigned, size_t, cudaStream_t);
template void zeroPadRows<unsigned>(unsigned*, unsigned, unsigned, size_t, cudaStream_t);
template void densify<float>(const MatrixDescription<float>&, float*, size_t, cudaStream_t);
template void densify<double>(const MatrixDescription<double>&, double*, size_t, cudaStream_t);

}  // namespace sparse
}  // namespace b200
}  // namespace nmfgpu
