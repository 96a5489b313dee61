// spmm.cu -- see spmm.h.
#include "spmm.h"

#include <algorithm>

#include <cub/cub.cuh>

#include <cstdint>
#include <cstdlib>

namespace nmfgpu {
namespace b200 {
namespace sparse {

namespace {

// ---- ingestion ---------------------------------------------------------------------------------------------
// compressed pointer array -> one major index per entry (one warp per compressed row/column)
__global__ void expand_major_kernel(unsigned numMajor, const int* __restrict__ ptr, int base, unsigned nnz, int* __restrict__ major) {
	const unsigned r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	if (r >= numMajor) return;
	const long long begin = (long long)ptr[r] - base, end = (long long)ptr[r + 1] - base;
	for (long long e = begin + (threadIdx.x % 32); e < end; e += 32)
		if (e >= 0 && e < (long long)nnz) major[e] = (int)r;
}

// zero-based coordinates; an entry outside the matrix is turned into an explicit zero at (0, 0)
template <typename T>
__global__ void rebase_kernel(unsigned nnz, int* __restrict__ row, int* __restrict__ col, T* __restrict__ val, int rowBase, int colBase,
                              unsigned rows, unsigned cols) {
	const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nnz) return;
	const long long r = (long long)row[e] - rowBase, c = (long long)col[e] - colBase;
	if (r < 0 || c < 0 || r >= (long long)rows || c >= (long long)cols) {
		row[e] = 0;
		col[e] = 0;
		val[e] = T(0);
	} else {
		row[e] = (int)r;
		col[e] = (int)c;
	}
}

__global__ void iota_kernel(unsigned count, unsigned* __restrict__ out) {
	const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < count) out[e] = e;
}

template <typename T>
__global__ void gather_kernel(unsigned count, const unsigned* __restrict__ perm, const int* __restrict__ idxIn, const T* __restrict__ valIn,
                              int* __restrict__ idxOut, T* __restrict__ valOut) {
	const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= count) return;
	const unsigned p = perm[e];
	idxOut[e] = idxIn[p];
	valOut[e] = valIn[p];
}

// ptr[r] = first position whose (sorted) key is >= r, for r in [0, numMajor]
__global__ void lower_bound_kernel(unsigned numMajor, unsigned count, const unsigned* __restrict__ sortedKeys, int* __restrict__ ptr) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > numMajor) return;
	unsigned lo = 0, hi = count;
	while (lo < hi) {
		const unsigned mid = lo + (hi - lo) / 2;
		if (sortedKeys[mid] < r) lo = mid + 1;
		else hi = mid;
	}
	ptr[r] = (int)lo;
}

unsigned bitsFor(unsigned count) {
	unsigned b = 1;
	while (b < 32 && (1ull << b) < count) ++b;
	return b;
}

// stable sort of the entry positions by `keys` (values < count): perm <- sorted positions, sortedKeys <- sorted keys
void sortPositions(unsigned nnz, const unsigned* keys, unsigned keyCount, unsigned* sortedKeys, unsigned* perm, cudaStream_t stream) {
	DeviceBuffer<unsigned> iota;
	iota.allocate(nnz);
	iota_kernel<<<ceilDiv(nnz, 256), 256, 0, stream>>>(nnz, iota.get());
	size_t tempBytes = 0;
	CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys, sortedKeys, iota.get(), perm, (int)nnz, 0, (int)bitsFor(keyCount), stream));
	DeviceBuffer<unsigned char> temp;
	temp.allocate(tempBytes);
	CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp.get(), tempBytes, keys, sortedKeys, iota.get(), perm, (int)nnz, 0, (int)bitsFor(keyCount), stream));
	CUDA_CHECK(cudaStreamSynchronize(stream));
}

// ---- products ------------------------------------------------------------------------------------------------
// One warp per compressed row/column.  The 32 lanes load 32 (index, value) pairs at once and hand them round by
// shuffle; every gather of an operand row is ONE 16-byte load per lane (lane l holds the output entries 4 l .. 4 l + 3 in
// fp32, 2 l, 2 l + 1 in fp64; k = 100: 25 lanes, 400 bytes, four 128-byte lines).  Four entries are in flight per step to
// cover the L2 / HBM latency of the gathers; per output entry the FMAs run in stored order (deterministic, and the same
// bits as the scalar variant below).  ncu on a quarter of configs[4] (profiles/r02_spmm_ncu_summary.json) showed the scalar
// variant -- four 4-byte loads per lane and gather -- at 62-72 % issue-slot utilisation with the L2 at 42-48 %: instruction
// bound before it is gather bound, hence one wide load instead of four.
template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };

template <typename T, int VQ>
__global__ void __launch_bounds__(256) spmm_gather_vec_kernel(unsigned numMajor, unsigned k, const int* __restrict__ ptrBegin,
                                                              const int* __restrict__ ptrEnd, const int* __restrict__ idx, const T* __restrict__ val,
                                                              const T* __restrict__ D, size_t ldd, T* __restrict__ out, size_t ldo) {
	using V = typename Vec16<T>::type;
	constexpr int E = Vec16<T>::N;
	const unsigned r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (r >= numMajor) return;
	const unsigned vecs = (k + E - 1) / E;   // 16-byte pieces of an operand row (the rows are padded to 32 entries: always readable)
	const int begin = ptrBegin[r], end = ptrEnd[r];
	T acc[VQ][E];
#pragma unroll
	for (int q = 0; q < VQ; ++q)
#pragma unroll
		for (int e = 0; e < E; ++e) acc[q][e] = T(0);
	auto load = [&](int jj, T (&d)[VQ][E]) {
		const V* row = reinterpret_cast<const V*>(D + (size_t)jj * ldd);
#pragma unroll
		for (int q = 0; q < VQ; ++q) {
			if (lane + 32 * q < vecs) {
				const V x = row[lane + 32 * q];
				const T* xe = reinterpret_cast<const T*>(&x);
#pragma unroll
				for (int e = 0; e < E; ++e) d[q][e] = xe[e];
			} else {
#pragma unroll
				for (int e = 0; e < E; ++e) d[q][e] = T(0);
			}
		}
	};
	for (int base = begin; base < end; base += 32) {
		const int e = base + (int)lane;
		const int j = e < end ? idx[e] : 0;
		const T v = e < end ? val[e] : T(0);
		const int cnt = min(32, end - base);
		int t = 0;
		for (; t + 4 <= cnt; t += 4) {
			T d[4][VQ][E], vv[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int jj = __shfl_sync(0xffffffffu, j, t + u);
				vv[u] = __shfl_sync(0xffffffffu, v, t + u);
				load(jj, d[u]);
			}
#pragma unroll
			for (int u = 0; u < 4; ++u)
#pragma unroll
				for (int q = 0; q < VQ; ++q)
#pragma unroll
					for (int x = 0; x < E; ++x) acc[q][x] = fma(vv[u], d[u][q][x], acc[q][x]);
		}
		for (; t < cnt; ++t) {
			const int jj = __shfl_sync(0xffffffffu, j, t);
			const T vv = __shfl_sync(0xffffffffu, v, t);
			T d[VQ][E];
			load(jj, d);
#pragma unroll
			for (int q = 0; q < VQ; ++q)
#pragma unroll
				for (int x = 0; x < E; ++x) acc[q][x] = fma(vv, d[q][x], acc[q][x]);
		}
	}
	T* orow = out + (size_t)r * ldo;
#pragma unroll
	for (int q = 0; q < VQ; ++q) {
		const unsigned e0 = (lane + 32 * q) * E;
		if (e0 + E <= k) {
			V x;
			T* xe = reinterpret_cast<T*>(&x);
#pragma unroll
			for (int e = 0; e < E; ++e) xe[e] = acc[q][e];
			*reinterpret_cast<V*>(orow + e0) = x;
		} else {
#pragma unroll
			for (int e = 0; e < E; ++e)
				if (e0 + e < k) orow[e0 + e] = acc[q][e];
		}
	}
}

// the scalar variant (NMFGPU_SPMM_SCALAR=1: A/B comparisons): lane l accumulates the output entries l, l + 32, ... (KQ of them)
template <typename T, int KQ>
__global__ void __launch_bounds__(256) spmm_gather_kernel(unsigned numMajor, unsigned k, const int* __restrict__ ptrBegin,
                                                          const int* __restrict__ ptrEnd, const int* __restrict__ idx, const T* __restrict__ val,
                                                          const T* __restrict__ D, size_t ldd, T* __restrict__ out, size_t ldo) {
	const unsigned r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (r >= numMajor) return;
	const int begin = ptrBegin[r], end = ptrEnd[r];
	T acc[KQ];
#pragma unroll
	for (int q = 0; q < KQ; ++q) acc[q] = T(0);
	for (int base = begin; base < end; base += 32) {
		const int e = base + (int)lane;
		const int j = e < end ? idx[e] : 0;
		const T v = e < end ? val[e] : T(0);
		const int cnt = min(32, end - base);
		int t = 0;
		for (; t + 4 <= cnt; t += 4) {
			T d[4][KQ], vv[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int jj = __shfl_sync(0xffffffffu, j, t + u);
				vv[u] = __shfl_sync(0xffffffffu, v, t + u);
				const T* row = D + (size_t)jj * ldd;
#pragma unroll
				for (int q = 0; q < KQ; ++q) d[u][q] = (lane + 32 * q < k) ? row[lane + 32 * q] : T(0);
			}
#pragma unroll
			for (int u = 0; u < 4; ++u)
#pragma unroll
				for (int q = 0; q < KQ; ++q) acc[q] = fma(vv[u], d[u][q], acc[q]);
		}
		for (; t < cnt; ++t) {
			const int jj = __shfl_sync(0xffffffffu, j, t);
			const T vv = __shfl_sync(0xffffffffu, v, t);
			const T* row = D + (size_t)jj * ldd;
#pragma unroll
			for (int q = 0; q < KQ; ++q)
				if (lane + 32 * q < k) acc[q] = fma(vv, row[lane + 32 * q], acc[q]);
		}
	}
#pragma unroll
	for (int q = 0; q < KQ; ++q)
		if (lane + 32 * q < k) out[(size_t)r * ldo + lane + 32 * q] = acc[q];
}

// blockPtr[b * numMajor + j] = first entry of compressed column j whose (ascending) minor index is >= b * minorPerBlock;
// row `blocks` holds the end of the column
__global__ void block_pointers_kernel(unsigned numMajor, unsigned blocks, unsigned minorPerBlock, const int* __restrict__ ptr,
                                      const int* __restrict__ idx, int* __restrict__ blockPtr) {
	const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned b = blockIdx.y;
	if (j >= numMajor) return;
	int lo = ptr[j], hi = ptr[j + 1];
	if (b >= blocks) {
		blockPtr[(size_t)b * numMajor + j] = hi;
		return;
	}
	const long long bound = (long long)b * minorPerBlock;
	while (lo < hi) {
		const int mid = lo + (hi - lo) / 2;
		if ((long long)idx[mid] < bound) lo = mid + 1;
		else hi = mid;
	}
	blockPtr[(size_t)b * numMajor + j] = lo;
}

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(unsigned rows, unsigned cols, const T* __restrict__ A, size_t lda, T* __restrict__ B, size_t ldb) {
	__shared__ T tile[32][33];
	const unsigned r0 = blockIdx.x * 32;
	const unsigned tx = threadIdx.x % 32, ty = threadIdx.x / 32;
	// the column tiles stride over gridDim.y (at most 65535 blocks; a transposed k x m matrix has m / 32 of them)
	for (unsigned c0 = blockIdx.y * 32; c0 < cols; c0 += gridDim.y * 32) {
		for (unsigned cc = ty; cc < 32; cc += 8) {
			const unsigned r = r0 + tx, c = c0 + cc;
			tile[cc][tx] = (r < rows && c < cols) ? A[(size_t)c * lda + r] : T(0);
		}
		__syncthreads();
		for (unsigned rr = ty; rr < 32; rr += 8) {
			const unsigned r = r0 + rr, c = c0 + tx;
			if (r < rows && c < cols) B[(size_t)r * ldb + c] = tile[tx][rr];
		}
		__syncthreads();
	}
}

template <typename T>
__global__ void major_squares_kernel(unsigned numMajor, const int* __restrict__ ptr, const T* __restrict__ val, T* __restrict__ out) {
	const unsigned r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (r >= numMajor) return;
	T acc = T(0);
	for (int e = ptr[r] + (int)lane; e < ptr[r + 1]; e += 32) acc = fma(val[e], val[e], acc);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) out[r] = acc;
}

}  // namespace

template <typename T>
void ingest(const MatrixDescription<T>& src, DeviceSparse<T>& dst, cudaStream_t stream) {
	const unsigned rows = src.rows, cols = src.columns;
	// the three sparse members of the union share one shape (values, ptrA, ptrB, nnz, base)
	const unsigned nnz = src.csr.nnz;
	const int base = src.csr.base == IndexBase::One ? 1 : 0;
	dst.rows = rows;
	dst.cols = cols;
	dst.nnz = nnz;
	dst.rowPtr.allocate(rows + 1);
	dst.colPtr.allocate(cols + 1);
	dst.rowPtr.zero(stream);
	dst.colPtr.zero(stream);
	if (nnz == 0) {
		CUDA_CHECK(cudaStreamSynchronize(stream));
		return;
	}
	if (nnz > 0x7fffffffu) throw EngineError(ResultType::ErrorInvalidArgument, "sparse matrix with more than 2^31 - 1 entries");
	if (src.csr.values == nullptr || src.csr.rowPtr == nullptr || src.csr.columnIndices == nullptr)
		throw EngineError(ResultType::ErrorInvalidArgument, "sparse matrix with null arrays");

	// ---- coordinates of every entry
	DeviceBuffer<int> row, col;
	DeviceBuffer<T> val;
	row.allocate(nnz);
	col.allocate(nnz);
	val.allocate(nnz);
	CUDA_CHECK(cudaMemcpyAsync(val.get(), src.csr.values, (size_t)nnz * sizeof(T), cudaMemcpyHostToDevice, stream));
	int rowBase = base, colBase = base;
	if (src.format == StorageFormat::CSR || src.format == StorageFormat::CSC) {
		const bool csr = src.format == StorageFormat::CSR;
		const unsigned numMajor = csr ? rows : cols;
		DeviceBuffer<int> ptr;
		ptr.allocate(numMajor + 1);
		CUDA_CHECK(cudaMemcpyAsync(ptr.get(), csr ? src.csr.rowPtr : src.csc.columnPtr, (size_t)(numMajor + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
		int* major = csr ? row.get() : col.get();
		int* minor = csr ? col.get() : row.get();
		CUDA_CHECK(cudaMemsetAsync(major, 0xFF, (size_t)nnz * sizeof(int), stream));   // entries no pointer range covers: out of range -> dropped
		CUDA_CHECK(cudaMemcpyAsync(minor, csr ? src.csr.columnIndices : src.csc.rowIndices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
		expand_major_kernel<<<ceilDiv(numMajor, 8), 256, 0, stream>>>(numMajor, ptr.get(), base, nnz, major);
		(csr ? rowBase : colBase) = 0;   // the expanded index is zero based already
		CUDA_CHECK(cudaStreamSynchronize(stream));   // ptr goes out of scope
	} else if (src.format == StorageFormat::COO) {
		CUDA_CHECK(cudaMemcpyAsync(row.get(), src.coo.rowIndices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
		CUDA_CHECK(cudaMemcpyAsync(col.get(), src.coo.columnIndices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
	} else {
		throw EngineError(ResultType::ErrorInvalidArgument, "unknown storage format");
	}
	rebase_kernel<T><<<ceilDiv(nnz, 256), 256, 0, stream>>>(nnz, row.get(), col.get(), val.get(), rowBase, colBase, rows, cols);
	CUDA_CHECK(cudaGetLastError());

	// ---- CSR copy: stable sort by row
	DeviceBuffer<unsigned> sortedKeys, perm;
	sortedKeys.allocate(nnz);
	perm.allocate(nnz);
	dst.colIdx.allocate(nnz);
	dst.csrVal.allocate(nnz);
	sortPositions(nnz, reinterpret_cast<const unsigned*>(row.get()), rows, sortedKeys.get(), perm.get(), stream);
	gather_kernel<T><<<ceilDiv(nnz, 256), 256, 0, stream>>>(nnz, perm.get(), col.get(), val.get(), dst.colIdx.get(), dst.csrVal.get());
	lower_bound_kernel<<<ceilDiv(rows + 1, 256), 256, 0, stream>>>(rows, nnz, sortedKeys.get(), dst.rowPtr.get());
	// row index of every entry in CSR order (the payload of the second sort)
	CUDA_CHECK(cudaMemcpyAsync(row.get(), sortedKeys.get(), (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, stream));

	// ---- CSC copy: stable sort of the CSR order by column, so rows ascend inside a column
	dst.rowIdx.allocate(nnz);
	dst.cscVal.allocate(nnz);
	sortPositions(nnz, reinterpret_cast<const unsigned*>(dst.colIdx.get()), cols, sortedKeys.get(), perm.get(), stream);
	gather_kernel<T><<<ceilDiv(nnz, 256), 256, 0, stream>>>(nnz, perm.get(), row.get(), dst.csrVal.get(), dst.rowIdx.get(), dst.cscVal.get());
	lower_bound_kernel<<<ceilDiv(cols + 1, 256), 256, 0, stream>>>(cols, nnz, sortedKeys.get(), dst.colPtr.get());
	CUDA_CHECK(cudaGetLastError());
	CUDA_CHECK(cudaStreamSynchronize(stream));
}

template <typename T>
void spmmGather(unsigned numMajor, unsigned k, const int* ptrBegin, const int* ptrEnd, const int* idx, const T* val, const T* D, size_t ldd, T* out,
                size_t ldo, cudaStream_t stream) {
	if (numMajor == 0) return;
	if (k > 128) throw EngineError(ResultType::ErrorInvalidArgument, "sparse products support at most 128 features");
	const unsigned grid = ceilDiv(numMajor, 8);
	static const bool scalar = [] {
		const char* e = getenv("NMFGPU_SPMM_SCALAR");
		return e != nullptr && atoi(e) != 0;
	}();
	const bool aligned = ldd % (16 / sizeof(T)) == 0 && ldo % (16 / sizeof(T)) == 0 && reinterpret_cast<uintptr_t>(D) % 16 == 0 &&
	                     reinterpret_cast<uintptr_t>(out) % 16 == 0;
	if (!scalar && aligned) {
		const unsigned vecs = ceilDiv(k, (unsigned)(16 / sizeof(T)));
		if (vecs <= 32) spmm_gather_vec_kernel<T, 1><<<grid, 256, 0, stream>>>(numMajor, k, ptrBegin, ptrEnd, idx, val, D, ldd, out, ldo);
		else spmm_gather_vec_kernel<T, 2><<<grid, 256, 0, stream>>>(numMajor, k, ptrBegin, ptrEnd, idx, val, D, ldd, out, ldo);
		CUDA_CHECK(cudaGetLastError());
		return;
	}
	switch (ceilDiv(k, 32)) {
	case 1: spmm_gather_kernel<T, 1><<<grid, 256, 0, stream>>>(numMajor, k, ptrBegin, ptrEnd, idx, val, D, ldd, out, ldo); break;
	case 2: spmm_gather_kernel<T, 2><<<grid, 256, 0, stream>>>(numMajor, k, ptrBegin, ptrEnd, idx, val, D, ldd, out, ldo); break;
	case 3: spmm_gather_kernel<T, 3><<<grid, 256, 0, stream>>>(numMajor, k, ptrBegin, ptrEnd, idx, val, D, ldd, out, ldo); break;
	default: spmm_gather_kernel<T, 4><<<grid, 256, 0, stream>>>(numMajor, k, ptrBegin, ptrEnd, idx, val, D, ldd, out, ldo); break;
	}
	CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void spmmGather(unsigned numMajor, unsigned k, const int* ptr, const int* idx, const T* val, const T* D, size_t ldd, T* out, size_t ldo,
                cudaStream_t stream) {
	spmmGather<T>(numMajor, k, ptr, ptr + 1, idx, val, D, ldd, out, ldo, stream);
}

void buildBlockPointers(unsigned numMajor, unsigned blocks, unsigned minorPerBlock, const int* ptr, const int* idx, int* blockPtr, cudaStream_t stream) {
	if (numMajor == 0) return;
	const dim3 grid(ceilDiv(numMajor, 256), blocks + 1);
	block_pointers_kernel<<<grid, 256, 0, stream>>>(numMajor, blocks, minorPerBlock, ptr, idx, blockPtr);
	CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void transpose(unsigned rows, unsigned cols, const T* A, size_t lda, T* B, size_t ldb, cudaStream_t stream) {
	if (rows == 0 || cols == 0) return;
	const dim3 grid(ceilDiv(rows, 32), std::min(ceilDiv(cols, 32), 65535u));
	transpose_kernel<T><<<grid, 256, 0, stream>>>(rows, cols, A, lda, B, ldb);
	CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void majorSquares(unsigned numMajor, const int* ptr, const T* val, T* out, cudaStream_t stream) {
	if (numMajor == 0) return;
	major_squares_kernel<T><<<ceilDiv(numMajor, 8), 256, 0, stream>>>(numMajor, ptr, val, out);
	CUDA_CHECK(cudaGetLastError());
}

#define NMF_INSTANTIATE(T)                                                                                                                    \
	template void ingest<T>(const MatrixDescription<T>&, DeviceSparse<T>&, cudaStream_t);                                                     \
	template void spmmGather<T>(unsigned, unsigned, const int*, const int*, const T*, const T*, size_t, T*, size_t, cudaStream_t);           \
	template void spmmGather<T>(unsigned, unsigned, const int*, const int*, const int*, const T*, const T*, size_t, T*, size_t, cudaStream_t); \
	template void transpose<T>(unsigned, unsigned, const T*, size_t, T*, size_t, cudaStream_t);                                               \
	template void majorSquares<T>(unsigned, const int*, const T*, T*, cudaStream_t);
NMF_INSTANTIATE(float)
NMF_INSTANTIATE(double)
#undef NMF_INSTANTIATE

}  // namespace sparse
}  // namespace b200
}  // namespace nmfgpu
