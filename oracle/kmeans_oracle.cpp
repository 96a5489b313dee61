// kmeans_oracle.cpp -- CPU restatement of the reference's Lloyd k-means (source/kmeans/kMeans.cu:125-278)
// with the reference's exact fp32 arithmetic order, so cluster assignments can be compared bit for bit.
//
// TEST INFRASTRUCTURE ONLY (see nmf_oracle.cpp for the rule).  Parity unpinned by the reference itself
// (it has no tests); pinned on the GPU box against the compiled reference (oracle/_ref) by
// tests/test_parity_gpu.py::test_kmeans_matches_reference.
//
// Arithmetic restated:
//   distance (kMeans.cu:40-50, KernelHelper.cuh:31-43): 32 lanes, lane l accumulates rows l, l+32, ...
//     with sum = fma(diff, diff, sum) (nvcc contracts `sum += diff*diff`), then the xor butterfly
//     16, 8, 4, 2, 1 in which every lane adds its partner's value -- all lanes end with the same bits;
//   argmin (kMeans.cu:62-71): strict `<`, lowest index wins ties;
//   centroid (kMeans.cu:105-121): sequential sum over members in ascending sample index, then a division;
//     empty clusters keep their centroid; rows beyond the launch coverage keep their value (SURVEY B-9).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>

namespace {
template <typename T>
T butterfly_sum(T lanes[32]) {
	for (int off = 16; off > 0; off >>= 1) {
		T next[32];
		for (int l = 0; l < 32; ++l) next[l] = lanes[l] + lanes[l ^ off];
		for (int l = 0; l < 32; ++l) lanes[l] = next[l];
	}
	return lanes[0];
}

template <typename T>
T distance_sq(const T* x, const T* c, unsigned m) {
	T lanes[32];
	for (unsigned l = 0; l < 32; ++l) {
		T sum = T(0);
		for (unsigned i = l; i < m; i += 32) {
			const T diff = x[i] - c[i];
			sum = std::fma(diff, diff, sum);
		}
		lanes[l] = sum;
	}
	return butterfly_sum(lanes);
}

template <typename T>
int kmeans_impl(unsigned m, unsigned n, unsigned k, const T* data, long ld, T* centroids, long ldc, unsigned* membership,
                unsigned seed, unsigned maxiter, double threshold, int reference_row_coverage, unsigned* rounds_out) {
	if (k == 0 || k > n) return -1;
	std::vector<unsigned> order(n);
	std::iota(order.begin(), order.end(), 0u);
	std::mt19937 gen(seed);
	std::shuffle(order.begin(), order.end(), gen);
	for (unsigned c = 0; c < k; ++c)
		for (unsigned i = 0; i < m; ++i) centroids[(size_t)c * ldc + i] = data[(size_t)order[c] * ld + i];
	for (unsigned j = 0; j < n; ++j) membership[j] = 0xFFFFFFFFu;

	unsigned rowLimit = m;
	if (reference_row_coverage) {
		const unsigned blocks = std::max(1u, ((m + 31) / 32) / 2u);
		rowLimit = (unsigned)std::min<size_t>(m, (size_t)blocks * 64);
	}
	auto assign = [&]() -> unsigned {
		unsigned changed = 0;
#pragma omp parallel for reduction(+ : changed) schedule(dynamic, 16)
		for (unsigned j = 0; j < n; ++j) {
			unsigned best = 0;
			T bestD = distance_sq(data + (size_t)j * ld, centroids, m);
			for (unsigned c = 1; c < k; ++c) {
				const T d = distance_sq(data + (size_t)j * ld, centroids + (size_t)c * ldc, m);
				if (d < bestD) { bestD = d; best = c; }
			}
			if (membership[j] != best) { membership[j] = best; ++changed; }
		}
		return changed;
	};
	unsigned iteration = 0;
	double fraction = 0.0;
	do {
		const unsigned changed = assign();
		fraction = changed / double(n);
		if (changed > 0) {
			std::vector<std::vector<unsigned>> members(k);
			for (unsigned j = 0; j < n; ++j) members[membership[j]].push_back(j);
#pragma omp parallel for schedule(dynamic, 1)
			for (unsigned c = 0; c < k; ++c) {
				if (members[c].empty()) continue;
				for (unsigned i = 0; i < rowLimit; ++i) {
					T sum = T(0);
					for (unsigned id : members[c]) sum += data[(size_t)id * ld + i];
					sum /= T(members[c].size());
					centroids[(size_t)c * ldc + i] = sum;
				}
			}
		}
	} while (++iteration < maxiter && fraction > threshold);
	if (fraction > 0.0) assign();
	if (rounds_out) *rounds_out = iteration;
	return 0;
}
}  // namespace

extern "C" {
int oracle_kmeans_f32(unsigned m, unsigned n, unsigned k, const float* data, long ld, float* centroids, long ldc, unsigned* membership,
                      unsigned seed, unsigned maxiter, double threshold, int reference_row_coverage, unsigned* rounds) {
	return kmeans_impl<float>(m, n, k, data, ld, centroids, ldc, membership, seed, maxiter, threshold, reference_row_coverage, rounds);
}
int oracle_kmeans_f64(unsigned m, unsigned n, unsigned k, const double* data, long ld, double* centroids, long ldc, unsigned* membership,
                      unsigned seed, unsigned maxiter, double threshold, int reference_row_coverage, unsigned* rounds) {
	return kmeans_impl<double>(m, n, k, data, ld, centroids, ldc, membership, seed, maxiter, threshold, reference_row_coverage, rounds);
}
}
