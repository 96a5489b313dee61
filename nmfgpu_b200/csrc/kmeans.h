// kmeans.h -- Lloyd k-means on the columns of a dense matrix (reference source/kmeans/kMeans.cu:125-278),
// used by nmfgpu_compute_kmeans_* and by the k-means based NMF initialisations.
//
// Bit-exactness contract (BASELINE.json north star: "k-means cluster assignments must be bit-exact"):
//   * Forgy seeding with libstdc++ std::mt19937(seed) + std::shuffle, first k indices (kMeans.cu:136-146);
//   * squared distances accumulate per lane over rows l, l+32, ... with fma(diff, diff, sum), followed by
//     the xor butterfly 16,8,4,2,1 (kMeans.cu:40-50, KernelHelper.cuh:31-43), argmin with strict `<`;
//   * centroids are sums over members in ascending sample index, divided by the count (kMeans.cu:105-121);
//     empty clusters keep their centroid;
//   * the reference's launch geometry never updates the last 32-row block when ceil(m/32) is odd
//     (kMeans.cu:218-222).  `referenceRowCoverage` reproduces that (default) or fixes it.
// What is NOT taken from the reference: the per-round D2H -> host std::sort -> H2D round trip is replaced
// by a stable device-side bucketing.
#pragma once
#include "common.h"

namespace nmfgpu {
namespace b200 {
class Communicator;
namespace kmeans {

// returns the number of Lloyd rounds executed; membership (n entries, device) holds the final assignment
template <typename T>
unsigned run(unsigned m, unsigned n, unsigned k, const T* data, size_t ldData, T* centroids, size_t ldCentroids, unsigned* membership,
             unsigned seed, unsigned maxIterations, double threshold, cudaStream_t stream, Communicator* comm,
             bool referenceRowCoverage = true);

}  // namespace kmeans
}  // namespace b200
}  // namespace nmfgpu
