// kmeans.cu -- see kmeans.h for the contract.
#include "kmeans.h"

#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#include "dist.h"

namespace nmfgpu {
namespace b200 {
namespace kmeans {

namespace {

template <typename T>
__device__ __forceinline__ T butterfly(T v) {
	v += __shfl_xor_sync(0xffffffffu, v, 16);
	v += __shfl_xor_sync(0xffffffffu, v, 8);
	v += __shfl_xor_sync(0xffffffffu, v, 4);
	v += __shfl_xor_sync(0xffffffffu, v, 2);
	v += __shfl_xor_sync(0xffffffffu, v, 1);
	return v;
}

// One warp per sample.  The sample's rows are kept in registers in chunks so that every centroid pass
// re-reads only the centroid (L1/L2 resident), not the sample.  Arithmetic order per (sample, cluster,
// lane) is exactly the reference's: rows lane, lane+32, ... in increasing order, fused multiply-add.
template <typename T>
__global__ void __launch_bounds__(256) assign_kernel(unsigned n, unsigned m, unsigned k, const T* __restrict__ data, size_t ldData,
                                                    const T* __restrict__ centroids, size_t ldC, unsigned* __restrict__ membership,
                                                    unsigned* __restrict__ changeCount) {
	const unsigned sample = blockIdx.x * 8 + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (sample >= n) return;
	const T* x = data + (size_t)sample * ldData;
	unsigned best = 0;
	T bestDist = T(0);
	for (unsigned c = 0; c < k; ++c) {
		const T* cc = centroids + (size_t)c * ldC;
		T sum = T(0);
		for (unsigned i = lane; i < m; i += 32) {
			const T diff = x[i] - cc[i];
			sum = fma(diff, diff, sum);
		}
		sum = butterfly(sum);
		if (c == 0 || sum < bestDist) {
			bestDist = sum;
			best = c;
		}
	}
	if (lane == 0 && membership[sample] != best) {
		membership[sample] = best;
		atomicAdd(changeCount, 1u);
	}
}

// stable bucketing, one warp per cluster: count, then (after the prefix) scatter in ascending sample index
__global__ void __launch_bounds__(256) bucket_count_kernel(unsigned n, unsigned k, const unsigned* __restrict__ membership,
                                                          unsigned* __restrict__ count) {
	const unsigned c = blockIdx.x * 8 + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (c >= k) return;
	unsigned total = 0;
	for (unsigned j0 = 0; j0 < n; j0 += 32) {
		const unsigned j = j0 + lane;
		const bool hit = j < n && membership[j] == c;
		total += __popc(__ballot_sync(0xffffffffu, hit));
	}
	if (lane == 0) count[c] = total;
}

__global__ void bucket_prefix_kernel(unsigned k, const unsigned* __restrict__ count, unsigned* __restrict__ entry) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		unsigned run = 0;
		for (unsigned c = 0; c < k; ++c) {
			entry[c] = run;
			run += count[c];
		}
	}
}

__global__ void __launch_bounds__(256) bucket_scatter_kernel(unsigned n, unsigned k, const unsigned* __restrict__ membership,
                                                            const unsigned* __restrict__ entry, unsigned* __restrict__ sorted) {
	const unsigned c = blockIdx.x * 8 + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (c >= k) return;
	unsigned pos = entry[c];
	for (unsigned j0 = 0; j0 < n; j0 += 32) {
		const unsigned j = j0 + lane;
		const bool hit = j < n && membership[j] == c;
		const unsigned mask = __ballot_sync(0xffffffffu, hit);
		if (hit) sorted[pos + __popc(mask & ((1u << lane) - 1u))] = j;
		pos += __popc(mask);
	}
}

// thread per (row, cluster): sequential mean over the members in ascending sample index
template <typename T>
__global__ void __launch_bounds__(256) centroid_kernel(unsigned rowLimit, unsigned k, const T* __restrict__ data, size_t ldData,
                                                      T* __restrict__ centroids, size_t ldC, const unsigned* __restrict__ sorted,
                                                      const unsigned* __restrict__ entry, const unsigned* __restrict__ count) {
	const unsigned row = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (row >= rowLimit || c >= k) return;
	const unsigned cnt = count[c];
	if (cnt == 0) return;  // empty clusters keep their previous centroid (kMeans.cu:92-94)
	const unsigned* ids = sorted + entry[c];
	T sum = T(0);
	for (unsigned q = 0; q < cnt; ++q) sum += data[(size_t)ids[q] * ldData + row];
	sum /= T(cnt);
	centroids[(size_t)c * ldC + row] = sum;
}

// Column shards: the same sequential sums, continued from rank to rank.  `running` holds the sums over the members owned
// by the ranks before this one (zeros on rank 0); this rank adds its own members in ascending sample index -- the exact
// chain of additions a single GPU performs, so the centroids (and with them every later assignment) stay bit-identical
// to the single-GPU run, which an all-reduce of per-rank sums would not give.
template <typename T>
__global__ void __launch_bounds__(256) centroid_accumulate_kernel(unsigned rowLimit, unsigned k, const T* __restrict__ data, size_t ldData,
                                                                 T* __restrict__ running, size_t ldC, const unsigned* __restrict__ sorted,
                                                                 const unsigned* __restrict__ entry, const unsigned* __restrict__ count) {
	const unsigned row = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (row >= rowLimit || c >= k) return;
	const unsigned cnt = count[c];
	const unsigned* ids = sorted + entry[c];
	T sum = running[(size_t)c * ldC + row];
	for (unsigned q = 0; q < cnt; ++q) sum += data[(size_t)ids[q] * ldData + row];
	running[(size_t)c * ldC + row] = sum;
}

// the last rank of the chain: mean over the members of all ranks; empty clusters keep their previous centroid
template <typename T>
__global__ void __launch_bounds__(256) centroid_finish_kernel(unsigned rowLimit, unsigned k, const T* __restrict__ running, T* __restrict__ centroids,
                                                             size_t ldC, const unsigned* __restrict__ globalCount) {
	const unsigned row = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (row >= rowLimit || c >= k) return;
	const unsigned cnt = globalCount[c];
	if (cnt == 0) return;
	T sum = running[(size_t)c * ldC + row];
	sum /= T(cnt);
	centroids[(size_t)c * ldC + row] = sum;
}

// point-to-point transfers of the Communicator are counted in floats
template <typename T>
Communicator::Transfer transferOf(T* buffer, size_t count, int peer) {
	return Communicator::Transfer{reinterpret_cast<float*>(buffer), count * (sizeof(T) / sizeof(float)), peer};
}

}  // namespace

template <typename T>
unsigned run(unsigned m, unsigned n, unsigned k, const T* data, size_t ldData, T* centroids, size_t ldCentroids, unsigned* membership,
             unsigned seed, unsigned maxIterations, double threshold, cudaStream_t stream, Communicator* comm, bool referenceRowCoverage) {
	// Column shards (dist.h): rank g holds the samples [c0, c0 + n) of N; the centroids are replicated.  Seeding picks global
	// sample indices, assignments are local, the centroid sums are chained through the ranks in rank order (see
	// centroid_accumulate_kernel), the change count is summed: memberships and centroids are those of the single-GPU run.
	const bool sharded = comm != nullptr && comm->worldSize() > 1;
	const unsigned G = sharded ? (unsigned)comm->worldSize() : 1u, rank = sharded ? (unsigned)comm->rank() : 0u;
	const unsigned N = sharded ? comm->globalColumns() : n, c0 = sharded ? comm->columnOffset() : 0u;
	if (k == 0 || k > N) throw EngineError(ResultType::ErrorInvalidArgument, "cluster count must be in [1, columns]");
	if (sharded) {
		struct Shard {
			unsigned offset, columns;
		};
		const Shard mine = {c0, n};
		std::vector<Shard> shards(G);
		comm->allGatherHost(&mine, sizeof(mine), shards.data());
		unsigned next = 0;
		for (const Shard& sh : shards) {   // "ascending sample index" must mean the same on one GPU and on G
			if (sh.offset != next) throw EngineError(ResultType::ErrorInvalidArgument, "k-means over column shards needs contiguous shards in rank order");
			next += sh.columns;
		}
		if (next != N) throw EngineError(ResultType::ErrorInvalidArgument, "the column shards of the ranks do not add up to the global column count");
	}

	// Phase 1: Forgy seeding (kMeans.cu:136-146)
	std::vector<unsigned> order(N);
	std::iota(order.begin(), order.end(), 0u);
	std::mt19937 generator(seed);
	std::shuffle(order.begin(), order.end(), generator);
	if (sharded) CUDA_CHECK(cudaMemsetAsync(centroids, 0, ldCentroids * (size_t)k * sizeof(T), stream));
	for (unsigned c = 0; c < k; ++c)
		if (order[c] >= c0 && order[c] - c0 < n)
			CUDA_CHECK(cudaMemcpyAsync(centroids + (size_t)c * ldCentroids, data + (size_t)(order[c] - c0) * ldData, (size_t)m * sizeof(T),
			                           cudaMemcpyDeviceToDevice, stream));
	if (sharded) comm->allReduceSum(centroids, ldCentroids * (size_t)k, stream);   // every seed has one owner: x + 0 + ... + 0 is exact

	DeviceBuffer<unsigned> changeCount, count, entry, sorted, globalCount;
	DeviceBuffer<T> running;
	changeCount.allocate(1);
	count.allocate(k);
	entry.allocate(k);
	sorted.allocate(std::max(1u, n));
	if (sharded) {
		globalCount.allocate(k);
		running.allocate(ldCentroids * (size_t)k);
	}
	// the reference starts from an uninitialised membership buffer; "no cluster" makes round 0 count every sample
	CUDA_CHECK(cudaMemsetAsync(membership, 0xFF, (size_t)n * sizeof(unsigned), stream));

	// rows the centroid update covers
	unsigned rowLimit = m;
	if (referenceRowCoverage) {
		const unsigned blocks = std::max(1u, ceilDiv(m, 32) / 2u);
		rowLimit = std::min<size_t>(m, (size_t)blocks * 64);
	}
	const dim3 centroidGrid(ceilDiv(rowLimit, 256), k);

	auto assign = [&]() {
		changeCount.zero(stream);
		if (n > 0) assign_kernel<T><<<ceilDiv(n, 8), 256, 0, stream>>>(n, m, k, data, ldData, centroids, ldCentroids, membership, changeCount.get());
		CUDA_CHECK(cudaGetLastError());
	};

	unsigned iteration = 0;
	double fraction = 0.0;
	unsigned changed = 0;
	do {
		assign();
		CUDA_CHECK(cudaMemcpyAsync(&changed, changeCount.get(), sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		if (sharded) changed = (unsigned)comm->allReduceSumHost((double)changed);
		fraction = changed / double(N);
		if (changed > 0) {
			bucket_count_kernel<<<ceilDiv(k, 8), 256, 0, stream>>>(n, k, membership, count.get());
			bucket_prefix_kernel<<<1, 32, 0, stream>>>(k, count.get(), entry.get());
			bucket_scatter_kernel<<<ceilDiv(k, 8), 256, 0, stream>>>(n, k, membership, entry.get(), sorted.get());
			if (!sharded) {
				centroid_kernel<T><<<centroidGrid, 256, 0, stream>>>(rowLimit, k, data, ldData, centroids, ldCentroids, sorted.get(), entry.get(), count.get());
				CUDA_CHECK(cudaGetLastError());
			} else {
				// members per cluster over all ranks
				std::vector<unsigned> mine(k), all((size_t)k * G), total(k, 0u);
				CUDA_CHECK(cudaMemcpyAsync(mine.data(), count.get(), (size_t)k * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
				CUDA_CHECK(cudaStreamSynchronize(stream));
				comm->allGatherHost(mine.data(), (size_t)k * sizeof(unsigned), all.data());
				for (unsigned g = 0; g < G; ++g)
					for (unsigned c = 0; c < k; ++c) total[c] += all[(size_t)g * k + c];
				CUDA_CHECK(cudaMemcpyAsync(globalCount.get(), total.data(), (size_t)k * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
				// the chain: round r moves the running sums from rank r to rank r + 1 (every rank takes part in every round:
				// the thread transport of dist.h is collective)
				if (rank == 0) running.zero(stream);
				const size_t words = ldCentroids * (size_t)k;
				for (unsigned r = 0; r + 1 < G; ++r) {
					std::vector<Communicator::Transfer> sends, recvs;
					if (rank == r) {
						centroid_accumulate_kernel<T><<<centroidGrid, 256, 0, stream>>>(rowLimit, k, data, ldData, running.get(), ldCentroids, sorted.get(),
						                                                                entry.get(), count.get());
						CUDA_CHECK(cudaGetLastError());
						sends.push_back(transferOf(running.get(), words, (int)(r + 1)));
					}
					if (rank == r + 1) recvs.push_back(transferOf(running.get(), words, (int)r));
					comm->exchange(sends, recvs, stream);
				}
				if (rank == G - 1) {
					centroid_accumulate_kernel<T><<<centroidGrid, 256, 0, stream>>>(rowLimit, k, data, ldData, running.get(), ldCentroids, sorted.get(),
					                                                                entry.get(), count.get());
					centroid_finish_kernel<T><<<centroidGrid, 256, 0, stream>>>(rowLimit, k, running.get(), centroids, ldCentroids, globalCount.get());
					CUDA_CHECK(cudaGetLastError());
				} else {
					CUDA_CHECK(cudaMemsetAsync(centroids, 0, words * sizeof(T), stream));
				}
				comm->allReduceSum(centroids, words, stream);   // broadcast of the last rank's centroids (the others add zeros)
			}
		}
	} while (++iteration < maxIterations && fraction > threshold);

	if (fraction > 0.0) assign();  // final assignment against the last centroids (kMeans.cu:269-277)
	CUDA_CHECK(cudaStreamSynchronize(stream));
	return iteration;
}

template unsigned run<float>(unsigned, unsigned, unsigned, const float*, size_t, float*, size_t, unsigned*, unsigned, unsigned, double,
                             cudaStream_t, Communicator*, bool);
template unsigned run<double>(unsigned, unsigned, unsigned, const double*, size_t, double*, size_t, unsigned*, unsigned, unsigned, double,
                              cudaStream_t, Communicator*, bool);

}  // namespace kmeans
}  // namespace b200
}  // namespace nmfgpu
