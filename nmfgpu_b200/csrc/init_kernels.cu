// init_kernels.cu -- see init_kernels.h.
#include "init_kernels.h"

#include <functional>
#include <random>
#include <vector>

namespace nmfgpu {
namespace b200 {
namespace init {

namespace {
constexpr unsigned kMeanCount = 5;  // MeanColumnStrategy.h default

template <typename T>
__global__ void mean_columns_kernel(unsigned m, unsigned k, const T* __restrict__ V, size_t ldv, T* __restrict__ W, size_t ldw,
                                    const unsigned* __restrict__ picks) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (i >= m || c >= k) return;
	T s = T(0);
	for (unsigned q = 0; q < kMeanCount; ++q) s += V[(size_t)picks[c * kMeanCount + q] * ldv + i];
	W[(size_t)c * ldw + i] = s / T(kMeanCount);
}
}  // namespace

template <typename T>
void meanColumns(unsigned m, unsigned n, unsigned k, const T* V, size_t ldv, T* W, size_t ldw, unsigned seed, cudaStream_t stream) {
	// Same generator as the reference (uniform_int over mt19937(seed)).  The reference leaves the very last
	// index unset (std::generate stops one short); here every index is drawn.
	std::vector<unsigned> picks((size_t)kMeanCount * k);
	auto draw = std::bind(std::uniform_int_distribution<unsigned>(0, n - 1), std::mt19937(seed));
	for (auto& p : picks) p = draw();
	DeviceBuffer<unsigned> dev;
	dev.allocate(picks.size());
	CUDA_CHECK(cudaMemcpyAsync(dev.get(), picks.data(), picks.size() * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
	dim3 grid(ceilDiv(m, 256), k);
	mean_columns_kernel<T><<<grid, 256, 0, stream>>>(m, k, V, ldv, W, ldw, dev.get());
	CUDA_CHECK(cudaGetLastError());
	CUDA_CHECK(cudaStreamSynchronize(stream));
}

template void meanColumns<float>(unsigned, unsigned, unsigned, const float*, size_t, float*, size_t, unsigned, cudaStream_t);
template void meanColumns<double>(unsigned, unsigned, unsigned, const double*, size_t, double*, size_t, unsigned, cudaStream_t);

}  // namespace init
}  // namespace b200
}  // namespace nmfgpu
