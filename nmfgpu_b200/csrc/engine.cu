// engine.cu -- buffer plan and per-algorithm iteration sequences (see engine.h).
//
// Iteration semantics follow SURVEY.md appendix A, which restates the reference:
//   MU     source/nmf/AlgorithmMultiplicativeFrobenius.h:149-248
//   GDCLS  source/nmf/AlgorithmGradientDescentConstrainedLeastSquares.h:159-265
//   ALS    source/nmf/AlgorithmAlternatingLeastSquares.h:145-234
//   ACLS / AHCLS  source/nmf/AlgorithmAlternatingHoyerConstrainedLeastSquares.h:170-295
//   nsNMF  source/nmf/AlgorithmNonSmoothNMF.h:173-225
// What differs from the reference is HOW: one stream, split-K partial products that the update
// kernels consume directly (no separate numerator/denominator matrices, no multiplyDivide pass,
// no 8-CTA normalisation kernel), and tensor-core contractions for fp32.
#include "engine.h"

#include <curand.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>

#include "dist.h"
#include "init_kernels.h"
#include "kernels.h"
#include "kmeans.h"
#include "sparse.h"
#include "tc_gemm.h"

namespace nmfgpu {
namespace b200 {

namespace {
unsigned pickSplits(unsigned tiles, unsigned reduceLen) {
	// enough CTAs for two waves on 148 SMs, but never slices shorter than 64 reduction steps
	unsigned s = ceilDiv(148, std::max(1u, tiles));
	s = std::min(s, std::max(1u, reduceLen / 256));
	return std::max(1u, s);
}
// The k x k Gram products are latency bound (one 64 x 64 tile, a long reduction): four CTAs per SM hide the load
// latency that one CTA of 8 warps cannot (measured: 28 % FMA utilisation at one CTA per SM), slices down to 64 steps.
unsigned pickGramSplits(unsigned tiles, unsigned reduceLen) {
	unsigned s = ceilDiv(4 * 148, std::max(1u, tiles));
	s = std::min(s, std::max(1u, reduceLen / 64));
	return std::max(1u, s);
}

// H2D of a dense column-major host matrix.  Callers of the reference API (nmfgpu4R) hand over ordinary pageable memory;
// cudaMemcpy from pageable memory goes through the driver's single-threaded staging at a fraction of the PCIe rate, and at
// 20 iterations that copy IS the call (4 GB against 30 ms of iterations).  Here the matrix moves in column chunks through
// two page-locked staging buffers (pooled between calls): a few host threads fill one buffer while the DMA engine drains
// the other.  Page-locked / registered sources take the direct path.
template <typename T>
void uploadDense(T* dev, size_t ldDev, const T* host, size_t ldHost, unsigned rows, unsigned cols, unsigned ranksOnThisHost, cudaStream_t stream) {
	const size_t columnBytes = (size_t)rows * sizeof(T);
	cudaPointerAttributes attr;
	bool pageable = true;
	if (cudaPointerGetAttributes(&attr, host) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
	else cudaGetLastError();
	if (const char* e = getenv("NMFGPU_STAGED_UPLOAD")) pageable = pageable && atoi(e) != 0;
	if (!pageable || columnBytes * cols < (32u << 20)) {
		CUDA_CHECK(cudaMemcpy2DAsync(dev, ldDev * sizeof(T), host, ldHost * sizeof(T), columnBytes, cols, cudaMemcpyHostToDevice, stream));
		return;
	}
	const size_t chunkBytes = 64u << 20;
	const unsigned chunkCols = (unsigned)std::max<size_t>(1, chunkBytes / columnBytes);
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	// measured on the 16-thread host of a B200 box, 4 GB: 4 threads 188 ms per call, 8: 158, 12: 144, 16: 137, 24: 140 (profiles/r02_notes.md)
	unsigned threads = std::max(2u, std::min(16u, hw / std::max(1u, ranksOnThisHost)));
	if (const char* e = getenv("NMFGPU_UPLOAD_THREADS")) threads = std::max(1u, std::min(64u, (unsigned)atoi(e)));   // study knob
	PinnedBuffer<T> staging[2];
	cudaEvent_t drained[2] = {nullptr, nullptr};
	for (int b = 0; b < 2; ++b) {
		staging[b].allocate((size_t)chunkCols * rows);
		CUDA_CHECK(cudaEventCreateWithFlags(&drained[b], cudaEventDisableTiming));
	}
	unsigned chunk = 0;
	for (unsigned c0 = 0; c0 < cols; c0 += chunkCols, ++chunk) {
		const unsigned nc = std::min(chunkCols, cols - c0);
		const int b = (int)(chunk & 1);
		if (chunk >= 2) CUDA_CHECK(cudaEventSynchronize(drained[b]));   // the DMA of two chunks ago has left this buffer
		T* stage = staging[b].get();
		auto fill = [&](unsigned t) {
			// thread t copies an equal share of the chunk's bytes, column by column (the host matrix may have ld > rows)
			const size_t total = (size_t)nc * columnBytes, begin = total * t / threads, end = total * (t + 1) / threads;
			size_t at = begin;
			while (at < end) {
				const size_t col = at / columnBytes, off = at % columnBytes;
				const size_t len = std::min(end - at, columnBytes - off);
				memcpy(reinterpret_cast<char*>(stage) + col * columnBytes + off,
				       reinterpret_cast<const char*>(host + (size_t)(c0 + col) * ldHost) + off, len);
				at += len;
			}
		};
		std::vector<std::thread> pool;
		for (unsigned t = 1; t < threads; ++t) pool.emplace_back(fill, t);
		fill(0);
		for (std::thread& th : pool) th.join();
		CUDA_CHECK(cudaMemcpy2DAsync(dev + (size_t)c0 * ldDev, ldDev * sizeof(T), stage, columnBytes, columnBytes, nc, cudaMemcpyHostToDevice, stream));
		CUDA_CHECK(cudaEventRecord(drained[b], stream));
	}
	CUDA_CHECK(cudaStreamSynchronize(stream));   // the staging buffers go back to the pool
	for (int b = 0; b < 2; ++b) cudaEventDestroy(drained[b]);
}
}  // namespace

template <typename T>
struct Engine<T>::TcPlan {
	tc::Plan plan;
};

template <typename T>
Engine<T>::Engine(const EngineConfig& cfg) : m_cfg(cfg), m_eps(std::numeric_limits<T>::epsilon()) {
	CUDA_CHECK(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking));
	m_profile = getenv("NMFGPU_PROFILE_ITERATION") != nullptr;
}

template <typename T>
void Engine<T>::stamp(const char* what) {
	if (!m_profile || m_stamps.size() > 4000) return;
	cudaEvent_t e;
	CUDA_CHECK(cudaEventCreate(&e));
	CUDA_CHECK(cudaEventRecord(e, m_stream));
	m_stamps.push_back({what, e});
}

// average in-stream duration of every step over the stamped iterations (the first two iterations are warm-up)
template <typename T>
void Engine<T>::reportStamps() {
	if (m_stamps.empty()) return;
	cudaStreamSynchronize(m_stream);
	const bool quiet = m_cfg.comm != nullptr && m_cfg.comm->rank() != 0;   // one report per job
	std::vector<std::pair<const char*, double>> sums;
	unsigned iterations = 0;
	for (size_t i = 1; i < m_stamps.size(); ++i) {
		if (strcmp(m_stamps[i - 1].first, "begin") == 0) ++iterations;
		if (strcmp(m_stamps[i].first, "begin") == 0 || iterations <= 2) continue;
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, m_stamps[i - 1].second, m_stamps[i].second) != cudaSuccess) {
			cudaGetLastError();
			continue;
		}
		bool found = false;
		for (auto& s : sums)
			if (s.first == m_stamps[i].first) {
				s.second += ms;
				found = true;
			}
		if (!found) sums.push_back({m_stamps[i].first, ms});
	}
	const double count = iterations > 2 ? iterations - 2 : 1;
	double total = 0.0;
	for (auto& s : sums) total += s.second / count;
	if (!quiet) {
		for (auto& s : sums) errorf("[iteration] %-28s %8.1f us\n", s.first, 1000.0 * s.second / count);
		errorf("[iteration] %-28s %8.1f us over %u iterations\n", "total", 1000.0 * total, (unsigned)count);
	}
	for (auto& s : m_stamps) cudaEventDestroy(s.second);
	m_stamps.clear();
}

template <typename T>
Engine<T>::~Engine() {
	if (m_stream) {
		reportStamps();
		cudaStreamSynchronize(m_stream);
		releaseFused();
		for (auto& row : m_graphExec)
			for (cudaGraphExec_t& g : row)
				if (g) cudaGraphExecDestroy(g);
		cudaStreamDestroy(m_stream);
	}
}

template <typename T>
void Engine<T>::synchronize() {
	CUDA_CHECK(cudaStreamSynchronize(m_stream));
}

template <typename T>
void Engine<T>::setup(const MatrixDescription<T>& V, bool vOnDevice) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	if (k == 0 || m == 0 || n == 0) throw EngineError(ResultType::ErrorInvalidArgument, "empty problem");
	PhaseTimer timer;
	// leading dimensions padded to 32 elements, as the reference's DeviceMatrix (Matrix.h:450-452):
	// keeps cuRAND-initialised factors bit-identical and every column 128-byte aligned for TMA.
	m_ldW = roundUp(m, 32);
	m_ldH = roundUp(k, 32);

	decideSparse(V, vOnDevice);
	if (m_sparse) {
		sparse::ingest(V, m_S, m_stream);
		m_ldV = 0;
		m_Wt.allocate(m_ldH * m);
		m_Pt.allocate(m_ldH * m);
		m_Wt.zero(m_stream);
		m_Pt.zero(m_stream);
		// Blocked sweep of W^T V (spmm.h), a study knob: measured at configs[4] (W copy 512 MB, 13 blocks of 40 MB) it
		// changes nothing, 13.1 vs 13.0 ms per iteration -- the gathers run at the L2-to-SM request rate whether the rows
		// come from L2 or from HBM (profiles/r01_notes.md 15) -- so the default is the single sweep.
		unsigned blocks = 1;
		if (const char* e = getenv("NMFGPU_SPARSE_BLOCKS")) blocks = std::max(1, std::min(64, atoi(e)));
		m_sparseBlocks = blocks;
		if (blocks > 1) {
			m_sparseBlockRows = (unsigned)roundUp(ceilDiv(m, blocks), 32);
			m_sparseBlocks = ceilDiv(m, m_sparseBlockRows);
			m_blockPtr.allocate((size_t)(m_sparseBlocks + 1) * n);
			sparse::buildBlockPointers(n, m_sparseBlocks, m_sparseBlockRows, m_S.colPtr.get(), m_S.rowIdx.get(), m_blockPtr.get(), m_stream);
		}
	} else if (vOnDevice) {
		if (V.format != StorageFormat::Dense) throw EngineError(ResultType::ErrorInvalidArgument, "device-resident V must be dense");
		m_ldV = V.dense.leadingDimension;
		m_V.adopt(V.dense.values, m_ldV * n);
	} else {
		m_ldV = roundUp(m, 32);
		m_V.allocate(m_ldV * n);
		if (V.format == StorageFormat::Dense) {
			if (V.dense.leadingDimension < m) throw EngineError(ResultType::ErrorInvalidArgument, "leading dimension of V too small");
			uploadDense<T>(m_V.get(), m_ldV, V.dense.values, V.dense.leadingDimension, m, n, m_cfg.comm ? (unsigned)m_cfg.comm->worldSize() : 1u, m_stream);
			if (m_ldV != m) sparse::zeroPadRows(m_V.get(), m, n, m_ldV, m_stream);
		} else {
			sparse::densify(V, m_V.get(), m_ldV, m_stream);
		}
	}

	timer.mark("  V on the device");
	for (int b = 0; b < 2; ++b) {
		m_W[b].allocate(m_ldW * k);
		m_W[b].zero(m_stream);
		m_H[b].allocate(m_ldH * n);
		m_H[b].zero(m_stream);
	}
	m_G.allocate((size_t)k * k);
	m_Gsaved.allocate((size_t)k * k);
	m_B.allocate((size_t)k * k);
	{
		// the k x k systems of the least-squares family are factorised in one block's shared memory (kernels.h qrFactor)
		const NmfAlgorithm a = m_cfg.algorithm;
		const bool ls = a == NmfAlgorithm::GDCLS || a == NmfAlgorithm::ALS || a == NmfAlgorithm::ACLS || a == NmfAlgorithm::AHCLS;
		if (ls && ((size_t)k * k + k) * sizeof(T) + 64 > (size_t)227 * 1024)
			throw EngineError(ResultType::ErrorInvalidArgument,
			                  std::is_same<T, float>::value ? "GDCLS / ALS / ACLS / AHCLS support up to 240 features in single precision"
			                                                : "GDCLS / ALS / ACLS / AHCLS support up to 169 features in double precision");
	}
	m_qr.allocate((size_t)k * k + k);
	m_inverse.allocate((size_t)k * k);
	if (std::is_same<T, float>::value) m_qrWork.allocate(3 * (size_t)k * k + k);

	// tensor-core eligibility: fp32, rank that fits one UMMA N, TMA-compatible strides
	m_useTC = false;
	if (std::is_same<T, float>::value && m_cfg.precision != Precision::Exact && !m_sparse) {
		const bool ok = tc::shapeSupported(m, n, k, m_ldV, m_ldW) && (reinterpret_cast<uintptr_t>(m_V.get()) % 16 == 0);
		if (!ok && (m_cfg.precision == Precision::Tf32x3 || m_cfg.precision == Precision::Tf32x1))
			throw EngineError(ResultType::ErrorInvalidArgument, "tensor-core precision requested for an unsupported shape");
		m_useTC = ok;
	}

	timer.mark("  factor buffers");
	m_fused = decideFused();
	// split-K plans of the two V-sized products and of the Gram products
	if (m_fused) {
		// MU on the tensor cores: buffers, row blocks and plans in setupFused() below
	} else if (m_useTC) {
		m_tc.reset(new TcPlan());
		m_ldHt = roundUp(n, 32);
		m_Whi.allocate(m_ldW * k);
		m_Wlo.allocate(m_ldW * k);
		m_HtHi.allocate(m_ldHt * k);
		m_HtLo.allocate(m_ldHt * k);
		m_Whi.zero(m_stream);
		m_Wlo.zero(m_stream);
		m_HtHi.zero(m_stream);
		m_HtLo.zero(m_stream);
		// centre of the products: the mean of V (any constant is exact algebra; the mean keeps the accumulators smallest)
		float center = tc::meanOf(reinterpret_cast<const float*>(m_V.get()), m, n, m_ldV, m_stream);
		if (const char* e = getenv("NMFGPU_TC_CENTER")) center = (float)atof(e) * center;   // study knob: 0 switches centring off
		tc::makePlan(m_tc->plan, m, n, k, reinterpret_cast<const float*>(m_V.get()), m_ldV, m_Whi.get(), m_Wlo.get(), m_ldW, m_HtHi.get(),
		             m_HtLo.get(), m_ldHt, m_cfg.precision == Precision::Tf32x1, center);
		m_corrN = reinterpret_cast<const T*>(m_tc->plan.corrN);
		m_corrP = reinterpret_cast<const T*>(m_tc->plan.corrP);
		m_splitsN = m_tc->plan.wtv.maxSlots;
		m_splitsP = m_tc->plan.vht.maxSlots;
		m_slotsN = m_tc->plan.wtv.slotCount;
		m_slotsP = m_tc->plan.vht.slotCount;
	} else if (m_sparse) {
		m_splitsN = m_sparseBlocks;  // one partial product per row block of the sweep (1: complete when written)
		m_splitsP = 1;
	} else {
		m_splitsN = kern::effectiveSplits(m, pickSplits(ceilDiv(k, 64) * ceilDiv(n, 64), m));
		m_splitsP = kern::effectiveSplits(n, pickSplits(ceilDiv(m, 64) * ceilDiv(k, 64), n));
	}
	timer.mark("  mean of V, TMA plans");
	m_splitsGW = kern::effectiveSplits(m, pickGramSplits(ceilDiv(k, 64) * ceilDiv(k, 64), m));
	m_splitsGH = kern::effectiveSplits(n, pickGramSplits(ceilDiv(k, 64) * ceilDiv(k, 64), n));
	m_strideN = m_ldH * n;
	m_strideP = m_ldW * k;
	if (!m_fused) {
		m_Npart.allocate(m_strideN * (m_splitsN + 1));   // +1: landing zone of the pre-reduced product (many partials per tile)
		m_Ppart.allocate(m_strideP * (m_splitsP + 1));  // +1: slot for the summed / all-reduced product
	}
	m_kkScratch.allocate((size_t)k * k * std::max(m_splitsGW, m_splitsGH));
	m_colSqPartials.allocate((size_t)ceilDiv(m, 128) * k);
	m_colSumPartials.allocate((size_t)ceilDiv(m, 128) * k);
	m_rowSumPartials.allocate((size_t)ceilDiv(n, 64) * k);
	m_colSq.allocate(k);
	// residual terms: per column of the caller's shard at setup (tr V^T V), per updated column in the iterations
	const unsigned globalColumns = m_cfg.comm != nullptr ? std::max(n, m_cfg.comm->globalColumns()) : n;
	m_partN.allocate(std::max(globalColumns, k));
	m_partK.allocate(k);
	m_hostSecond.allocate(std::max(globalColumns, k));
	m_hostThird.allocate(k);
	if (m_cfg.algorithm == NmfAlgorithm::nsNMF) {
		m_smoothW.allocate(m_ldW * k);
		m_smoothH.allocate(m_ldH * n);
		m_smoothW.zero(m_stream);
		m_smoothH.zero(m_stream);
	}

	timer.mark("  scratch buffers");
	// tr(V^T V) per column, sorted ascending on the host (MU.h:117-125)
	if (m_sparse) sparse::majorSquares<T>(n, m_S.colPtr.get(), m_S.cscVal.get(), m_partN.get(), m_stream);
	else kern::columnDots<T>(m, n, m_V.get(), m_ldV, m_V.get(), m_ldV, m_partN.get(), m_stream);
	m_vtvSorted.resize(n);
	CUDA_CHECK(cudaMemcpyAsync(m_hostSecond.get(), m_partN.get(), n * sizeof(T), cudaMemcpyDeviceToHost, m_stream));
	synchronize();
	std::copy(m_hostSecond.get(), m_hostSecond.get() + n, m_vtvSorted.begin());
	std::sort(m_vtvSorted.begin(), m_vtvSorted.end());
	m_vtvSum = 0.0;
	for (T v : m_vtvSorted) m_vtvSum += (double)v;
	timer.mark("  tr(V^T V)");
	if (m_fused) setupFused();
	timer.mark("  row blocks, exchange buffer, plans");
}

// Sparse inputs run compressed when densifying is wasteful or impossible (spmm.h); the reference always densifies.
template <typename T>
void Engine<T>::decideSparse(const MatrixDescription<T>& V, bool vOnDevice) {
	m_sparse = false;
	if (vOnDevice || V.format == StorageFormat::Dense || m_cfg.needsDenseV || m_cfg.k > 128) return;
	if (const char* e = getenv("NMFGPU_SPARSE")) {
		m_sparse = atoi(e) != 0;
		return;
	}
	const double cells = (double)m_cfg.m * (double)m_cfg.n;
	size_t freeBytes = 0, totalBytes = 0;
	CUDA_CHECK(cudaMemGetInfo(&freeBytes, &totalBytes));
	m_sparse = (double)V.csr.nnz <= 0.02 * cells || cells * sizeof(T) > 0.5 * (double)freeBytes;
}

// ---- fused MU: one GPU, or row blocks over several (fused.h, dist.h) ------------------------------------------------
// The caller hands in column shards (the north star's contract).  m >> n in the workloads this library serves, so the
// shards are regrouped ONCE into row blocks V[I_g, :]; from then on
//   W^T V    : rank g forms the partial W[I_g]^T V[I_g, :] (k x N) and its tcgen05 kernel stores every tile straight into
//              the memory of the rank that owns those columns of H (NVLink peer stores),
//   H update : the owners add the partials up, update their columns and store them (and their TF32 split) to every rank,
//   V H^T    : V[I_g, :] H^T, the rank's own rows, no reduction,
//   W update : rows I_g only; the unit-column scaling is applied when W is read, its k*k + k statistics travel as peer
//              stores.  W is never gathered during the iterations.
// Exchanged per iteration and rank: (G-1)/G of k x N partials out, k x N/G columns of H to G-1 ranks, 2 (k*k + k) statistics.
template <typename T>
bool Engine<T>::decideFused() {
	if (!std::is_same<T, float>::value || !m_useTC || m_cfg.algorithm != NmfAlgorithm::Multiplicative) return false;
	if (const char* e = getenv("NMFGPU_FUSED"))
		if (atoi(e) == 0) return false;
	Communicator* comm = m_cfg.comm;
	if (comm == nullptr || comm->worldSize() <= 1) return true;
	const unsigned G = (unsigned)comm->worldSize();
	const char* mode = getenv("NMFGPU_DIST_MODE");
	const unsigned mrPad = (unsigned)roundUp(ceilDiv(m_cfg.m, G), 256);
	bool ok = G <= fused::kMaxRanks && !(mode != nullptr && strcmp(mode, "allreduce") == 0) && (unsigned long long)(G - 1) * mrPad < m_cfg.m;
	// every rank must take the same path: the two dataflows exchange different things
	return comm->allReduceSumHost(ok ? 1.0 : 0.0) == (double)G;
}

template <typename T>
void Engine<T>::setupFused() {
	Communicator* comm = (m_cfg.comm != nullptr && m_cfg.comm->worldSize() > 1) ? m_cfg.comm : nullptr;
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	const unsigned G = comm ? (unsigned)comm->worldSize() : 1u, rank = comm ? (unsigned)comm->rank() : 0u;
	const unsigned N = comm ? comm->globalColumns() : n;
	m_globalN = N;
	m_peers.world = G;
	m_peers.rank = rank;
	fused::configure();

	// ---- the row block of this rank
	if (comm == nullptr) {
		m_mrPad = m;
		m_r0 = 0;
		m_mr = m;
		m_ldVr = m_ldV;
		m_Vblock = reinterpret_cast<const float*>(m_V.get());
	} else {
		m_mrPad = (unsigned)roundUp(ceilDiv(m, G), 256);
		m_r0 = rank * m_mrPad;
		m_mr = std::min(m_mrPad, m - m_r0);
		m_ldVr = roundUp(m_mr, 32);
		auto rowsOf = [&](unsigned g) { return std::min(m_mrPad, m - g * m_mrPad); };
		// who holds which columns (shards may be unequal)
		struct Shard {
			unsigned offset, columns;
		};
		const Shard mine = {comm->columnOffset(), n};
		std::vector<Shard> shards(G);
		comm->allGatherHost(&mine, sizeof(mine), shards.data());
		unsigned long long covered = 0;
		for (const Shard& sh : shards) covered += sh.columns;
		if (covered != N) throw EngineError(ResultType::ErrorInvalidArgument, "the column shards of the ranks do not add up to the global column count");
		// V[I_g, :] from the column shards: rank g sends V[I_h, J_g] to every h, packed with the receiver's leading dimension
		m_Vr.allocate(m_ldVr * N);
		m_Vr.zero(m_stream);
		size_t total = 0;
		for (unsigned h = 0; h < G; ++h)
			if (h != rank) total += roundUp(rowsOf(h), 32) * n;
		DeviceBuffer<float> pack;
		pack.allocate(total);
		pack.zero(m_stream);
		std::vector<Communicator::Transfer> sends, recvs;
		size_t at = 0;
		const float* V = reinterpret_cast<const float*>(m_V.get());
		for (unsigned h = 0; h < G; ++h) {
			const size_t ldh = roundUp(rowsOf(h), 32);
			if (h == rank) {
				CUDA_CHECK(cudaMemcpy2DAsync(m_Vr.get() + m_ldVr * (size_t)mine.offset, m_ldVr * sizeof(float), V + m_r0, m_ldV * sizeof(float),
				                             (size_t)m_mr * sizeof(float), n, cudaMemcpyDeviceToDevice, m_stream));
				continue;
			}
			CUDA_CHECK(cudaMemcpy2DAsync(pack.get() + at, ldh * sizeof(float), V + (size_t)h * m_mrPad, m_ldV * sizeof(float),
			                             (size_t)rowsOf(h) * sizeof(float), n, cudaMemcpyDeviceToDevice, m_stream));
			sends.push_back({pack.get() + at, ldh * n, (int)h});
			if (shards[h].columns > 0) recvs.push_back({m_Vr.get() + m_ldVr * (size_t)shards[h].offset, m_ldVr * shards[h].columns, (int)h});
			at += ldh * n;
		}
		if (n == 0) sends.clear();
		comm->exchange(sends, recvs, m_stream);
		synchronize();
		m_Vblock = m_Vr.get();
	}

	// ---- the columns of H this rank updates (128-column granularity: a panel never straddles a stream-K tile)
	m_colsPerRank = (unsigned)roundUp(ceilDiv(N, G), 128);
	m_c0 = (unsigned)std::min<unsigned long long>(N, (unsigned long long)rank * m_colsPerRank);
	m_nOwn = (unsigned)std::min<unsigned long long>(N, (unsigned long long)m_c0 + m_colsPerRank) - m_c0;

	// ---- operands of the products and their plan.  The centre must be the same number on every rank (the rank-one terms
	// it leaves out are added once, after the partials of all ranks have been summed).
	m_Whi.allocate(m_ldW * k);
	m_Wlo.allocate(m_ldW * k);
	m_Whi.zero(m_stream);
	m_Wlo.zero(m_stream);
	m_ldHtFull = roundUp(N, 32);
	float center = tc::meanOf(m_Vblock, m_mr, N, m_ldVr, m_stream);
	if (comm != nullptr) center = (float)(comm->allReduceSumHost((double)center * (double)m_mr * (double)N) / ((double)m * (double)N));
	if (const char* e = getenv("NMFGPU_TC_CENTER")) center = (float)atof(e) * center;   // study knob: 0 switches centring off

	// slots per rank follow from the shape-only work split: plan once with placeholder operands to learn it
	unsigned info[5];
	m_slotsPerRank = 1;
	{
		std::vector<unsigned char> perTile(ceilDiv(N, 128));
		int sms = 0, dev = 0;
		CUDA_CHECK(cudaGetDevice(&dev));
		CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
		tc::enumerateSegments(N, comm ? m_mrPad : m, (unsigned)roundUp(k, 16), (unsigned)sms, nullptr, 0, info, perTile.data(), (unsigned)perTile.size());
		for (unsigned char c : perTile) m_slotsPerRank = std::max<unsigned>(m_slotsPerRank, c);
	}

	// ---- exchange buffer: the same layout on every rank
	auto align = [](size_t x) { return roundUp(x, 256); };
	m_lay.statLen = (unsigned)roundUp((size_t)k * k + k + 1, 32);
	size_t at = 0;
	m_lay.flagsN = at; at = align(at + fused::kMaxRanks * sizeof(unsigned));
	m_lay.flagsH = at; at = align(at + fused::kMaxRanks * sizeof(unsigned));
	m_lay.statW = at; at = align(at + (size_t)G * m_lay.statLen * sizeof(float));
	m_lay.statH = at; at = align(at + (size_t)G * m_lay.statLen * sizeof(float));
	m_lay.trace = at; at = align(at + fused::kMaxRanks * sizeof(double));
	m_lay.H = at; at = align(at + m_ldH * (size_t)G * m_colsPerRank * sizeof(float));   // room for an in-place all-gather of the owners' blocks
	m_lay.HtHi = at; at = align(at + m_ldHtFull * (size_t)k * sizeof(float));
	m_lay.HtLo = at; at = align(at + m_ldHtFull * (size_t)k * sizeof(float));
	// partial products of W^T V: one GPU -- the stream-K slots of the tensor-core kernel; several -- one partial per rank
	// for the own columns (the slots of the own row block stay in m_Nlocal, fused::pushN sends their sum to the owners)
	m_lay.slots = at; at = align(at + (size_t)(comm ? G : m_slotsPerRank) * m_ldH * m_colsPerRank * sizeof(float));
	m_lay.bytes = at;
	m_sym = static_cast<char*>(pooledDeviceAlloc(m_lay.bytes));
	CUDA_CHECK(cudaMemsetAsync(m_sym, 0, m_lay.bytes, m_stream));
	synchronize();   // zeroed before any other rank can store into it (openPeers is a collective)
	if (comm != nullptr) {
		m_peerPtrs = comm->openPeers(m_sym);
		for (unsigned g = 0; g < G; ++g) m_peers.base[g] = static_cast<char*>(m_peerPtrs[g]);
	} else {
		m_peers.base[0] = m_sym;
	}
	if (comm != nullptr) m_Nlocal.allocate((size_t)m_slotsPerRank * m_ldH * N);

	m_tc.reset(new TcPlan());
	float* HtHi = reinterpret_cast<float*>(m_sym + m_lay.HtHi);
	float* HtLo = reinterpret_cast<float*>(m_sym + m_lay.HtLo);
	tc::makePlan(m_tc->plan, m_mr, N, k, m_Vblock, m_ldVr, m_Whi.get() + m_r0, m_Wlo.get() + m_r0, m_ldW, HtHi, HtLo, m_ldHtFull,
	             m_cfg.precision == Precision::Tf32x1, center, comm ? m_mrPad : 0);
	if (m_tc->plan.wtv.maxSlots > m_slotsPerRank) throw EngineError(ResultType::ErrorExternalLibrary, "work split of W^T V changed between two plans of the same shape");
	m_splitsN = m_slotsPerRank;
	m_slotsN = m_tc->plan.wtv.slotCount;
	m_slotsP = m_tc->plan.vht.slotCount;
	m_corrN = reinterpret_cast<const T*>(m_tc->plan.corrN);
	m_corrP = reinterpret_cast<const T*>(m_tc->plan.corrP);
	m_strideN = m_ldH * (size_t)(comm ? N : m_colsPerRank);   // between the stream-K slots of W^T V
	m_splitsPr = m_splitsP = m_tc->plan.vht.maxSlots;
	m_ldPr = roundUp(m_mr, 32);
	m_stridePr = m_ldPr * k;
	m_PpartR.allocate(m_stridePr * m_splitsPr);

	m_statSum.allocate(m_lay.statLen);
	m_inv.allocate(roundUp(k, 32));
	m_statPartH.allocate((size_t)std::max(1u, ceilDiv(m_nOwn, fused::panelColumnsH(m_nOwn))) * ((size_t)k * k + k));
	m_statPartW.allocate((size_t)std::max(1u, ceilDiv(m_mr, 128)) * ((size_t)k * k + k));
	m_ctlWords.allocate(32);
	m_ctlWords.zero(m_stream);
	m_ctl.epoch = m_ctlWords.get();
	m_ctl.error = m_ctlWords.get() + 1;
	m_ctl.tickets = m_ctlWords.get() + 4;
	m_wStatFlag = reinterpret_cast<float*>(m_ctlWords.get() + 16);
	// Ranks that share a GPU (the thread transport of the tests): a kernel that waits for another rank's signal can keep
	// that rank's kernels from being scheduled -- or from being submitted at all -- for as long as it spins.  There the
	// ranks proceed in lockstep on the HOST at the two points of an iteration where they wait for each other (stream
	// synchronise + barrier, no CUDA graphs); the in-kernel waits still run and find their flags already set.
	m_hostLockstep = comm != nullptr && comm->ranksMayShareDevice();
	m_hostFlags.allocate(2);
	m_hostFlags.get()[0] = m_hostFlags.get()[1] = 0;
	m_hostTrace.allocate(fused::kMaxRanks);
	m_vtvTotal = comm ? comm->allReduceSumHost(m_vtvSum) : m_vtvSum;
	synchronize();
}

template <typename T>
void Engine<T>::releaseFused() {
	if (m_sym == nullptr) return;
	cudaStreamSynchronize(m_stream);
	Communicator* comm = (m_cfg.comm != nullptr && m_cfg.comm->worldSize() > 1) ? m_cfg.comm : nullptr;
	if (comm != nullptr && !m_peerPtrs.empty()) {
		try {
			comm->barrier();   // nobody stores into a buffer that is about to be recycled
			comm->closePeers(m_peerPtrs);
			comm->barrier();
		} catch (...) {
		}
	}
	pooledDeviceFree(m_sym, m_lay.bytes);
	m_sym = nullptr;
}

// Initial factors are in m_W (all rows on every rank) and m_H (the caller's columns): derive what the iteration keeps
// up to date itself -- the TF32 split and the statistics of the own rows of W (flag 0: the reference uses the initial W
// as it is and normalises only after an update, MU.h:247), the full H with its transposed split.
template <typename T>
void Engine<T>::finishInitialisationFused() {
	Communicator* comm = (m_cfg.comm != nullptr && m_cfg.comm->worldSize() > 1) ? m_cfg.comm : nullptr;
	const unsigned k = m_cfg.k, n = m_cfg.n, N = m_globalN;
	if (comm != nullptr) {   // no rank is still iterating on the buffers that are rewritten below
		synchronize();
		comm->barrier();
	}
	float* W = reinterpret_cast<float*>(m_W[m_wCur].get()) + m_r0;
	kern::splitTf32(m_mr, k, W, m_ldW, m_Whi.get() + m_r0, m_Wlo.get() + m_r0, m_ldW, m_stream);
	m_wUpdated = false;
	m_blocksW = fused::updateW(m_peers, m_lay, 0.f, m_mr, k, nullptr, nullptr, nullptr, W, m_ldW, m_Whi.get() + m_r0, m_Wlo.get() + m_r0, nullptr, 0, 0,
	                           nullptr, 0.f, m_statPartW.get(), m_wStatFlag, false, m_stream);
	fused::reducePush(m_peers, m_lay.statW, m_lay.statLen, m_statPartW.get(), m_blocksW, k * k + k, 0.f, fused::kNoSignal, m_ctl, 2, m_stream);
	float* Hfull = reinterpret_cast<float*>(m_sym + m_lay.H);
	const float* Hmine = reinterpret_cast<const float*>(m_H[m_hCur].get());
	if (comm == nullptr) {
		CUDA_CHECK(cudaMemcpyAsync(Hfull, Hmine, m_ldH * (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, m_stream));
	} else {
		CUDA_CHECK(cudaMemsetAsync(Hfull, 0, m_ldH * (size_t)N * sizeof(float), m_stream));
		CUDA_CHECK(cudaMemcpyAsync(Hfull + m_ldH * (size_t)comm->columnOffset(), Hmine, m_ldH * (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, m_stream));
		comm->allReduceSum(Hfull, m_ldH * (size_t)N, m_stream);   // every column has exactly one non-zero contribution
	}
	tc::splitTransposeH(k, N, Hfull, m_ldH, reinterpret_cast<float*>(m_sym + m_lay.HtHi), reinterpret_cast<float*>(m_sym + m_lay.HtLo), m_ldHtFull, m_stream);
	tc::refreshCorrectionH(m_tc->plan, Hfull, m_ldH, m_stream);   // centring term of V H^T for the initial H (debugProducts; the iteration recomputes it)
	m_launches += 6;
	if (comm != nullptr) {
		synchronize();
		comm->barrier();
	}
}

// the stream-K slots of W^T V of the own row block: into the exchange buffer on one GPU (the H update reads them there),
// into m_Nlocal over several ranks (pushN sums them on their way to the owners)
template <typename T>
void Engine<T>::productWtVFused() {
	float* slots = m_peers.world > 1 ? m_Nlocal.get() : reinterpret_cast<float*>(m_sym + m_lay.slots);
	tc::gemmWtV(m_tc->plan, slots, m_ldH, m_strideN, m_stream);
}

template <typename T>
void Engine<T>::iterateMUFused(bool err) {
	const unsigned k = m_cfg.k;
	tc::Plan& plan = m_tc->plan;
	float* G = reinterpret_cast<float*>(m_G.get());
	float* B = reinterpret_cast<float*>(m_B.get());
	stamp("begin");
	// ---- H <- H o (W^T V) / ((W^T W) H + eps)   (MU.h:164-198)
	const bool several = m_peers.world > 1;
	productWtVFused();
	stamp("product W^T V");
	if (several) {
		// (the same launch sums and sends the statistics of the last W update: no launch of their own over several ranks)
		fused::pushN(m_peers, m_lay, m_ctl, plan.kp, m_globalN, m_colsPerRank, m_ldH, m_Nlocal.get(), m_strideN, plan.wtv.slotCount, m_statPartW.get(), m_blocksW,
		             k * k + k, m_wStatFlag, m_stream);
		stamp("partials to the owners");
		m_launches += 1;
	}
	if (several && m_hostLockstep) {
		synchronize();
		m_cfg.comm->barrier();
	}
	const unsigned blocksH = fused::updateH(m_peers, m_lay, m_ctl, plan.center, k, m_c0, m_nOwn, m_colsPerRank, m_ldH, m_ldHtFull, several ? 1u : m_slotsPerRank,
	                                        several ? nullptr : plan.wtv.slotCount, G, m_inv.get(), plan.corrN, (float)m_eps,
	                                        err ? reinterpret_cast<float*>(m_partN.get()) : nullptr, m_statPartH.get(), m_stream);
	stamp("update H");
	if (err && several) {
		fused::traceSumPush(m_peers, m_lay, reinterpret_cast<const float*>(m_partN.get()), m_nOwn, m_stream);
		m_launches += 1;
	}
	fused::reducePush(m_peers, m_lay.statH, m_lay.statLen, m_statPartH.get(), blocksH, k * k + k, -1.f, m_lay.flagsH, m_ctl, 1, m_stream);
	stamp("H statistics");
	m_launches += 3;
	if (several && m_hostLockstep) {
		synchronize();
		m_cfg.comm->barrier();
	}
	if (err || m_cfg.constantW) {
		// H H^T before the W update (the trace term), or without one: as a kernel of its own (it waits for the other ranks)
		fused::finishH(m_peers, m_lay, m_ctl, k, plan.center, B, plan.corrP, m_stream);
		m_launches += 1;
	}
	if (err) {
		kern::traceKK<T>(k, m_B.get(), m_G.get(), m_partK.get(), m_stream);                     // tr(HH^T W^T W) MU.h:203-216
		m_launches += 1;
	}
	// ---- W <- W o (V H^T) / (W (H H^T) + eps), unit columns (MU.h:200-248)
	if (!m_cfg.constantW) {
		tc::Gate gate;
		gate.flags = reinterpret_cast<const unsigned*>(m_sym + m_lay.flagsH);
		gate.epoch = m_ctl.epoch;
		gate.error = m_ctl.error;
		gate.count = m_peers.world;
		tc::gemmVHt(plan, m_PpartR.get(), m_ldPr, m_stridePr, m_stream, several ? &gate : nullptr);
		stamp("product V H^T");
		float* W = reinterpret_cast<float*>(m_W[m_wCur].get()) + m_r0;
		m_wUpdated = true;
		m_blocksW = fused::updateW(m_peers, m_lay, plan.center, m_mr, k, B, plan.corrP, m_inv.get(), W, m_ldW, m_Whi.get() + m_r0, m_Wlo.get() + m_r0,
		                           m_PpartR.get(), m_ldPr, m_stridePr, plan.vht.slotCount, (float)m_eps, m_statPartW.get(), m_wStatFlag, true, m_stream);
		stamp("update W");
		m_launches += 2;
		if (!several) {   // several ranks: the next iteration's pushN does this
			fused::reducePush(m_peers, m_lay.statW, m_lay.statLen, m_statPartW.get(), m_blocksW, k * k + k, 1.f, fused::kNoSignal, m_ctl, 2, m_stream);
			stamp("W statistics");
			m_launches += 1;
		}
	}
	if (err) resolveError(m_nOwn);
}

// The best run's factors for the caller (MU.h:251-254): W with unit columns -- materialised here, the iterations never
// do it -- gathered over the row blocks; H: the caller's columns.
template <typename T>
void Engine<T>::storeFused(const MatrixDescription<T>& hostW, const MatrixDescription<T>& hostH) {
	Communicator* comm = (m_cfg.comm != nullptr && m_cfg.comm->worldSize() > 1) ? m_cfg.comm : nullptr;
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	if (comm != nullptr) {
		// over several ranks the statistics of the last W update travel with the NEXT iteration's pushN: send them now
		fused::reducePush(m_peers, m_lay.statW, m_lay.statLen, m_statPartW.get(), m_blocksW, k * k + k, m_wUpdated ? 1.f : 0.f, fused::kNoSignal, m_ctl, 2,
		                  m_stream);
		// and H: every rank keeps only its own columns current (fused.cu update_h_fused); in-place all-gather of the owners' blocks
		float* Hall = reinterpret_cast<float*>(m_sym + m_lay.H);
		const size_t block = m_ldH * (size_t)m_colsPerRank;
		comm->allGather(Hall + block * m_peers.rank, Hall, block, m_stream);
	}
	synchronize();
	if (comm != nullptr) comm->barrier();   // the statistics of every rank's last W update have landed
	fused::prepH(m_peers, m_lay, k, m_tc->plan.center, m_statSum.get(), reinterpret_cast<float*>(m_Gsaved.get()), m_inv.get(), m_tc->plan.corrN, m_stream);
	float* spare = reinterpret_cast<float*>(m_W[1 - m_wCur].get());
	const float* W = reinterpret_cast<const float*>(m_W[m_wCur].get()) + m_r0;
	if (comm == nullptr) {
		fused::scaleRows(m, (unsigned)m_ldW, k, W, m_ldW, m_inv.get(), spare, m_stream);
	} else {
		DeviceBuffer<float> block, gathered;
		block.allocate((size_t)m_mrPad * k);
		gathered.allocate((size_t)m_mrPad * k * comm->worldSize());
		fused::scaleRows(m_mr, m_mrPad, k, W, m_ldW, m_inv.get(), block.get(), m_stream);
		comm->allGather(block.get(), gathered.get(), (size_t)m_mrPad * k, m_stream);
		fused::unpackRows(m, k, m_mrPad, gathered.get(), spare, m_ldW, m_stream);
		synchronize();
	}
	if (hostW.dense.values != nullptr)
		CUDA_CHECK(cudaMemcpy2DAsync(hostW.dense.values, (size_t)hostW.dense.leadingDimension * sizeof(T), spare, m_ldW * sizeof(T), (size_t)m * sizeof(T), k,
		                             cudaMemcpyDeviceToHost, m_stream));
	const float* Hfull = reinterpret_cast<const float*>(m_sym + m_lay.H) + m_ldH * (size_t)(comm ? comm->columnOffset() : 0);
	if (hostH.dense.values != nullptr)
		CUDA_CHECK(cudaMemcpy2DAsync(hostH.dense.values, (size_t)hostH.dense.leadingDimension * sizeof(T), Hfull, m_ldH * sizeof(T), (size_t)k * sizeof(T), n,
		                             cudaMemcpyDeviceToHost, m_stream));
	measureSparsity(reinterpret_cast<const T*>(spare), m_ldW, reinterpret_cast<const T*>(Hfull), m_ldH);
	synchronize();
	finishSparsity();
	checkDeviceFlags();
}

// A wait inside a kernel that gave up (a tensor-core barrier, tc_gemm.cu, or another rank's signal, fused.cu) leaves
// garbage behind; the counters are read at the host's synchronisation points and turned into an error code.
template <typename T>
void Engine<T>::checkDeviceFlags() {
	if (!m_useTC) {
		synchronize();
		return;
	}
	unsigned* host = m_hostFlags.get();
	if (host == nullptr) {
		m_hostFlags.allocate(2);
		host = m_hostFlags.get();
	}
	host[0] = host[1] = 0;
	CUDA_CHECK(cudaMemcpyAsync(host, tc::timeoutCounter(), sizeof(unsigned), cudaMemcpyDeviceToHost, m_stream));
	if (m_ctl.error != nullptr) CUDA_CHECK(cudaMemcpyAsync(host + 1, m_ctl.error, sizeof(unsigned), cudaMemcpyDeviceToHost, m_stream));
	synchronize();
	if (host[0] != 0) throw EngineError(ResultType::ErrorExternalLibrary, "a barrier wait inside the tensor-core product timed out: the factors are not valid (NMFGPU_TC_DEBUG=1 reports where)");
	if (host[1] != 0) throw EngineError(ResultType::ErrorExternalLibrary, "another rank did not answer within 10 s: the factors are not valid");
}

// ---- initial factors --------------------------------------------------------------------------------

template <typename T>
void Engine<T>::loadW(const MatrixDescription<T>& hostW) {
	if (hostW.format != StorageFormat::Dense || hostW.dense.values == nullptr || hostW.dense.leadingDimension < m_cfg.m)
		throw EngineError(ResultType::ErrorInvalidArgument, "matrix W must be a dense m x k host matrix");
	CUDA_CHECK(cudaMemcpy2DAsync(m_W[m_wCur].get(), m_ldW * sizeof(T), hostW.dense.values, (size_t)hostW.dense.leadingDimension * sizeof(T),
	                             (size_t)m_cfg.m * sizeof(T), m_cfg.k, cudaMemcpyHostToDevice, m_stream));
}

template <typename T>
void Engine<T>::loadH(const MatrixDescription<T>& hostH) {
	if (hostH.format != StorageFormat::Dense || hostH.dense.values == nullptr || hostH.dense.leadingDimension < m_cfg.k)
		throw EngineError(ResultType::ErrorInvalidArgument, "matrix H must be a dense k x n host matrix");
	CUDA_CHECK(cudaMemcpy2DAsync(m_H[m_hCur].get(), m_ldH * sizeof(T), hostH.dense.values, (size_t)hostH.dense.leadingDimension * sizeof(T),
	                             (size_t)m_cfg.k * sizeof(T), m_cfg.n, cudaMemcpyHostToDevice, m_stream));
}

namespace {
// cuRAND XORWOW over ld x cols elements: the stream the reference draws (RandomValueStrategy.cpp:29-38),
// so AllRandomValues runs start from the same factors as the reference for the same seed.
void curandFill(float* p, size_t count, unsigned seed, cudaStream_t stream) {
	curandGenerator_t gen;
	if (curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS)
		throw EngineError(ResultType::ErrorExternalLibrary, "curandCreateGenerator failed");
	curandSetStream(gen, stream);
	// the reference passes the seed through an `int` parameter (RandomValueStrategy.cpp:29,41): sign-extended to 64 bits
	curandSetPseudoRandomGeneratorSeed(gen, (unsigned long long)(long long)(int)seed);
	const curandStatus_t st = curandGenerateUniform(gen, p, count);
	curandDestroyGenerator(gen);
	if (st != CURAND_STATUS_SUCCESS) throw EngineError(ResultType::ErrorExternalLibrary, "curandGenerateUniform failed");
}
void curandFill(double* p, size_t count, unsigned seed, cudaStream_t stream) {
	curandGenerator_t gen;
	if (curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS)
		throw EngineError(ResultType::ErrorExternalLibrary, "curandCreateGenerator failed");
	curandSetStream(gen, stream);
	curandSetPseudoRandomGeneratorSeed(gen, (unsigned long long)(long long)(int)seed);
	const curandStatus_t st = curandGenerateUniformDouble(gen, p, count);
	curandDestroyGenerator(gen);
	if (st != CURAND_STATUS_SUCCESS) throw EngineError(ResultType::ErrorExternalLibrary, "curandGenerateUniformDouble failed");
}
}  // namespace

template <typename T>
void Engine<T>::randomW(unsigned seed) {
	curandFill(m_W[m_wCur].get(), m_ldW * m_cfg.k, seed, m_stream);
	if (m_ldW != m_cfg.m) sparse::zeroPadRows(m_W[m_wCur].get(), m_cfg.m, m_cfg.k, m_ldW, m_stream);
}

template <typename T>
void Engine<T>::randomH(unsigned seed) {
	// With column shards every rank draws the full-width stream and keeps its own columns, so the
	// factors do not depend on the number of GPUs.
	Communicator* comm = m_cfg.comm;
	if (comm == nullptr || comm->worldSize() == 1) {
		curandFill(m_H[m_hCur].get(), m_ldH * m_cfg.n, seed, m_stream);
	} else {
		DeviceBuffer<T> full;
		full.allocate(m_ldH * comm->globalColumns());
		curandFill(full.get(), m_ldH * comm->globalColumns(), seed, m_stream);
		CUDA_CHECK(cudaMemcpyAsync(m_H[m_hCur].get(), full.get() + m_ldH * comm->columnOffset(), m_ldH * m_cfg.n * sizeof(T),
		                           cudaMemcpyDeviceToDevice, m_stream));
		synchronize();
	}
	if (m_ldH != m_cfg.k) sparse::zeroPadRows(m_H[m_hCur].get(), m_cfg.k, m_cfg.n, m_ldH, m_stream);
}

template <typename T>
void Engine<T>::meanColumnsW(unsigned seed) {
	if (m_sparse) throw EngineError(ResultType::ErrorInvalidArgument, "MeanColumns initialisation needs the dense matrix (NMFGPU_SPARSE=0)");
	// W is replicated: every rank would average columns of its own shard and start from a different W
	if (m_cfg.comm != nullptr && m_cfg.comm->worldSize() > 1)
		throw EngineError(ResultType::ErrorInvalidArgument, "MeanColumns initialisation is not available with column shards");
	init::meanColumns<T>(m_cfg.m, m_cfg.n, m_cfg.k, m_V.get(), m_ldV, m_W[m_wCur].get(), m_ldW, seed, m_stream);
}

template <typename T>
void Engine<T>::kmeansW(unsigned seed) {
	if (m_sparse) throw EngineError(ResultType::ErrorInvalidArgument, "k-means initialisation needs the dense matrix (NMFGPU_SPARSE=0)");
	// k-means on the data columns; the centroids become W (KMeansStrategy.cpp:54-58: 100 rounds, 0.5 %)
	DeviceBuffer<unsigned> membership;
	membership.allocate(m_cfg.n);
	kmeans::run<T>(m_cfg.m, m_cfg.n, m_cfg.k, m_V.get(), m_ldV, m_W[m_wCur].get(), m_ldW, membership.get(), seed, 100, 0.005, m_stream,
	               m_cfg.comm);
	synchronize();
}

template <typename T>
void Engine<T>::hFromWtV(bool absolute) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	if (m_sparse) {
		productWtV(m_W[m_wCur].get());
		kern::sumSplits<T>(k, n, m_Npart.get(), m_ldH, m_splitsN, m_strideN, m_H[m_hCur].get(), m_ldH, m_stream);
		if (absolute) kern::absInPlace<T>(k, n, m_H[m_hCur].get(), m_ldH, m_stream);
		else kern::clampNonNegative<T>(k, n, m_H[m_hCur].get(), m_ldH, m_stream);
		return;
	}
	// exact fp32 SIMT product into scratch slices of its own (the iteration's slot buffers belong to the tensor-core plan)
	const unsigned eff = kern::effectiveSplits(m, std::min(16u, pickSplits(ceilDiv(k, 64) * ceilDiv(n, 64), m)));
	const size_t stride = m_ldH * (size_t)n;
	DeviceBuffer<T> slices;
	slices.allocate(stride * eff);
	kern::gemmTN<T>(m, k, n, m_W[m_wCur].get(), m_ldW, m_V.get(), m_ldV, slices.get(), m_ldH, eff, stride, m_stream);
	kern::sumSplits<T>(k, n, slices.get(), m_ldH, eff, stride, m_H[m_hCur].get(), m_ldH, m_stream);
	if (absolute) kern::absInPlace<T>(k, n, m_H[m_hCur].get(), m_ldH, m_stream);
	else kern::clampNonNegative<T>(k, n, m_H[m_hCur].get(), m_ldH, m_stream);
	synchronize();
}

template <typename T>
void Engine<T>::finishInitialisation() {
	if (!m_useTC) return;
	if (m_fused) {
		finishInitialisationFused();
		return;
	}
	float* W = reinterpret_cast<float*>(m_W[m_wCur].get());
	float* H = reinterpret_cast<float*>(m_H[m_hCur].get());
	kern::splitTf32(m_cfg.m, m_cfg.k, W, m_ldW, m_Whi.get(), m_Wlo.get(), m_ldW, m_stream);
	tc::splitTransposeH(m_cfg.k, m_cfg.n, H, m_ldH, m_HtHi.get(), m_HtLo.get(), m_ldHt, m_stream);
	operandChangedW(m_W[m_wCur].get());
	operandChangedH(m_H[m_hCur].get());
}

// the tensor-core products read hi/lo copies of W resp. H; their rank-one centring terms follow the same values
template <typename T>
void Engine<T>::operandChangedW(const T* W) {
	if (!m_useTC) return;
	tc::refreshCorrectionW(m_tc->plan, reinterpret_cast<const float*>(W), m_ldW, m_stream);
	m_launches += 2;
}

template <typename T>
void Engine<T>::operandChangedH(const T* H) {
	if (!m_useTC) return;
	tc::refreshCorrectionH(m_tc->plan, reinterpret_cast<const float*>(H), m_ldH, m_stream);
	m_launches += 2;
}

// ---- shared building blocks ---------------------------------------------------------------------------

template <typename T>
void Engine<T>::gramW(const T* W, T* G) {
	const unsigned k = m_cfg.k;
	kern::gemmTN<T>(m_cfg.m, k, k, W, m_ldW, W, m_ldW, m_kkScratch.get(), k, m_splitsGW, (size_t)k * k, m_stream);
	kern::sumSplits<T>(k, k, m_kkScratch.get(), k, m_splitsGW, (size_t)k * k, G, k, m_stream);
	m_launches += 2;
}

template <typename T>
void Engine<T>::gramH(const T* H, size_t ldh, T* B) {
	const unsigned k = m_cfg.k;
	kern::gemmNT<T>(k, m_cfg.n, k, H, ldh, H, ldh, m_kkScratch.get(), k, m_splitsGH, (size_t)k * k, m_stream);
	kern::sumSplits<T>(k, k, m_kkScratch.get(), k, m_splitsGH, (size_t)k * k, B, k, m_stream);
	m_launches += 2;
	if (m_cfg.comm && m_cfg.comm->worldSize() > 1) m_cfg.comm->allReduceSum(B, (size_t)k * k, m_stream);
}

template <typename T>
void Engine<T>::productWtV(const T* W) {
	if (m_sparse) {   // N[:, j] = sum over the entries of column j of v * W[i, :]: gathers rows of the row-major copy of W
		sparse::transpose<T>(m_cfg.m, m_cfg.k, W, m_ldW, m_Wt.get(), m_ldH, m_stream);
		if (m_sparseBlocks <= 1) {
			sparse::spmmGather<T>(m_cfg.n, m_cfg.k, m_S.colPtr.get(), m_S.rowIdx.get(), m_S.cscVal.get(), m_Wt.get(), m_ldH, m_Npart.get(), m_ldH, m_stream);
		} else {
			for (unsigned b = 0; b < m_sparseBlocks; ++b)
				sparse::spmmGather<T>(m_cfg.n, m_cfg.k, m_blockPtr.get() + (size_t)b * m_cfg.n, m_blockPtr.get() + (size_t)(b + 1) * m_cfg.n, m_S.rowIdx.get(),
				                      m_S.cscVal.get(), m_Wt.get(), m_ldH, m_Npart.get() + m_strideN * b, m_ldH, m_stream);
		}
		m_launches += m_sparseBlocks;
	} else if (m_useTC) {
		static const bool poison = getenv("NMFGPU_TC_POISON") != nullptr;   // debugging: an unwritten slot shows up as NaN
		if (poison) CUDA_CHECK(cudaMemsetAsync(m_Npart.get(), 0xFF, m_Npart.bytes(), m_stream));
		tc::gemmWtV(m_tc->plan, reinterpret_cast<float*>(m_Npart.get()), m_ldH, m_strideN, m_stream);
	} else {
		kern::gemmTN<T>(m_cfg.m, m_cfg.k, m_cfg.n, W, m_ldW, m_V.get(), m_ldV, m_Npart.get(), m_ldH, m_splitsN, m_strideN, m_stream);
	}
	m_launches += 1;
}

template <typename T>
void Engine<T>::productVHt(const T* H, size_t ldh) {
	if (m_sparse) {   // P[i, :] = sum over the entries of row i of v * H[:, j]: gathers columns of H, then back to column-major
		sparse::spmmGather<T>(m_cfg.m, m_cfg.k, m_S.rowPtr.get(), m_S.colIdx.get(), m_S.csrVal.get(), H, ldh, m_Pt.get(), m_ldH, m_stream);
		sparse::transpose<T>(m_cfg.k, m_cfg.m, m_Pt.get(), m_ldH, m_Ppart.get(), m_ldW, m_stream);
		m_launches += 1;
	} else if (m_useTC) {
		static const bool poison = getenv("NMFGPU_TC_POISON") != nullptr;
		if (poison) CUDA_CHECK(cudaMemsetAsync(m_Ppart.get(), 0xFF, m_Ppart.bytes(), m_stream));
		tc::gemmVHt(m_tc->plan, reinterpret_cast<float*>(m_Ppart.get()), m_ldW, m_strideP, m_stream);
	} else {
		kern::gemmNT<T>(m_cfg.m, m_cfg.n, m_cfg.k, m_V.get(), m_ldV, H, ldh, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, m_stream);
	}
	m_launches += 1;
}

template <typename T>
void Engine<T>::normaliseW(unsigned blocks, bool haveColumnSums) {
	// with the per-block column sums of the update kernel the centring term of W^T V falls out of the norm kernel
	const bool fused = haveColumnSums && m_useTC;
	kern::finishColumnNorms<T>(m_cfg.k, blocks, m_colSqPartials.get(), m_colSq.get(), m_stream, fused ? m_colSumPartials.get() : nullptr,
	                           fused ? m_tc->plan.center : 0.f, fused ? m_tc->plan.corrN : nullptr);
	kern::scaleColumns<T>(m_cfg.m, m_cfg.k, m_W[m_wCur].get(), m_ldW, m_colSq.get(), m_useTC ? m_Whi.get() : nullptr,
	                      m_useTC ? m_Wlo.get() : nullptr, m_stream);
	m_launches += 2;
	if (!fused) operandChangedW(m_W[m_wCur].get());
}

// W <- W o P / (W B + eps) then unit columns; P is read as split partials, or -- with column shards --
// summed and all-reduced into the spare slot first.
template <typename T>
void Engine<T>::multiplicativeW(const T* B) {
	const T* P = m_Ppart.get();
	unsigned splits = m_splitsP;
	const unsigned char* slots = m_slotsP;
	const T* corr = m_corrP;
	if (m_cfg.comm && m_cfg.comm->worldSize() > 1) {
		T* sum = m_Ppart.get() + m_strideP * m_splitsP;
		kern::sumSplits<T>(m_cfg.m, m_cfg.k, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, sum, m_ldW, m_stream, m_slotsP, true, m_corrP);
		m_cfg.comm->allReduceSum(sum, m_strideP, m_stream);
		m_launches += 1;
		P = sum;
		splits = 1;
		slots = nullptr;
		corr = nullptr;
	}
	const unsigned blocks = kern::updateW<T>(m_cfg.m, m_cfg.k, B, m_W[m_wCur].get(), m_W[1 - m_wCur].get(), m_ldW, P, m_ldW, splits, m_strideP,
	                                         m_eps, m_colSqPartials.get(), m_stream, slots, corr, m_useTC ? m_colSumPartials.get() : nullptr);
	m_launches += 1;
	m_wCur = 1 - m_wCur;
	normaliseW(blocks, true);
}

// A tile of W^T V may receive many partial products (4 reduction chunks x 5 CTAs on one GPU, 30 CTAs per tile on an
// 8-GPU shard).  The update kernel has only n/64 blocks and would walk them one memory latency after the other (68 us
// measured at 8 GPUs, 50 us at one); one thread per element adds them up first.
template <typename T>
void Engine<T>::preReduceN(const T*& N, unsigned& splits, const unsigned char*& slots, const T*& corr) {
	if (m_splitsN <= 24) return;   // up to ~20 partials the update kernel walks them faster than a separate pass (50 vs 58 us at one GPU)
	T* sum = m_Npart.get() + m_strideN * m_splitsN;
	kern::sumSplits<T>(m_cfg.k, m_cfg.n, m_Npart.get(), m_ldH, m_splitsN, m_strideN, sum, m_ldH, m_stream, m_slotsN, false, m_corrN);
	m_launches += 1;
	N = sum;
	splits = 1;
	slots = nullptr;
	corr = nullptr;
}

// ---- MU -------------------------------------------------------------------------------------------------
template <typename T>
void Engine<T>::iterateMU(bool err) {
	const unsigned n = m_cfg.n, k = m_cfg.k;
	float* htHi = m_useTC ? m_HtHi.get() : nullptr;
	float* htLo = m_useTC ? m_HtLo.get() : nullptr;

	stamp("begin");
	gramW(m_W[m_wCur].get(), m_G.get());                                                    // A = W^T W      MU.h:168/176
	stamp("gram W^T W");
	productWtV(m_W[m_wCur].get());                                                          // N = W^T V      MU.h:187
	stamp("product W^T V");
	const T* N = m_Npart.get();
	unsigned splits = m_splitsN;
	const unsigned char* slots = m_slotsN;
	const T* corr = m_corrN;
	preReduceN(N, splits, slots, corr);
	kern::updateH<T>(k, n, m_G.get(), m_H[m_hCur].get(), m_H[1 - m_hCur].get(), m_ldH, N, m_ldH, splits, m_strideN, m_eps,
	                 err ? m_partN.get() : nullptr, htHi, htLo, m_ldHt, m_stream, slots, corr,
	                 m_useTC ? m_rowSumPartials.get() : nullptr);                           // H update  MU.h:181-197
	m_launches += 1;
	m_hCur = 1 - m_hCur;
	if (m_useTC) {   // centring term of V H^T from the row sums the update kernel left per 64-column block
		kern::finishPartialSums(k, ceilDiv(n, 64), reinterpret_cast<const float*>(m_rowSumPartials.get()), m_tc->plan.center, m_tc->plan.corrP, m_stream);
		m_launches += 1;
	}
	stamp("update H (+ row sums)");

	gramH(m_H[m_hCur].get(), m_ldH, m_B.get());                                             // B = H H^T      MU.h:208/231
	stamp("gram H H^T");
	if (err) {
		kern::traceKK<T>(k, m_B.get(), m_G.get(), m_partK.get(), m_stream);                 // tr(HH^T W^T W) MU.h:203-216
		m_launches += 1;
	}
	if (!m_cfg.constantW) {
		productVHt(m_H[m_hCur].get(), m_ldH);                                               // N2 = V H^T     MU.h:240
		stamp("product V H^T");
		multiplicativeW(m_B.get());                                                         // MU.h:235-247
		stamp("update W, norms, scale");
	}
	if (err) resolveError(n);
}

// ---- nsNMF ------------------------------------------------------------------------------------------------
template <typename T>
void Engine<T>::iterateNsNMF(bool err) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	const T theta = (T)m_cfg.params.theta;
	float* htHi = m_useTC ? m_HtHi.get() : nullptr;
	float* htLo = m_useTC ? m_HtLo.get() : nullptr;

	kern::smoothRight<T>(m, k, m_W[m_wCur].get(), m_ldW, m_smoothW.get(), m_ldW, theta, m_stream);   // W~ = W S   nsNMF.h:174
	m_launches += 1;
	if (m_useTC) {  // the tensor-core product reads the hi/lo split of its left factor
		kern::splitTf32(m, k, reinterpret_cast<float*>(m_smoothW.get()), m_ldW, m_Whi.get(), m_Wlo.get(), m_ldW, m_stream);
		m_launches += 1;
		operandChangedW(m_smoothW.get());
	}
	gramW(m_smoothW.get(), m_G.get());                                                      // W~^T W~
	productWtV(m_smoothW.get());                                                            // W~^T V
	// the H^T split written here is overwritten below by the split of S H (what V H~^T consumes)
	kern::updateH<T>(k, n, m_G.get(), m_H[m_hCur].get(), m_H[1 - m_hCur].get(), m_ldH, m_Npart.get(), m_ldH, m_splitsN, m_strideN, m_eps,
	                 err ? m_partN.get() : nullptr, nullptr, nullptr, m_ldHt, m_stream, m_slotsN, m_corrN);
	m_launches += 1;
	m_hCur = 1 - m_hCur;
	if (!err && m_cfg.constantW) return;                                                     // nsNMF.h:193-195

	kern::smoothLeft<T>(k, n, m_H[m_hCur].get(), m_ldH, m_smoothH.get(), m_ldH, theta, m_stream);   // H~ = S H   nsNMF.h:197
	m_launches += 1;
	gramH(m_smoothH.get(), m_ldH, m_B.get());                                               // H~ H~^T
	if (err) {
		gramW(m_W[m_wCur].get(), m_Gsaved.get());                                           // W^T W (unsmoothed) nsNMF.h:202-203
		kern::traceKK<T>(k, m_B.get(), m_Gsaved.get(), m_partK.get(), m_stream);
		m_launches += 1;
	}
	if (!m_cfg.constantW) {
		if (m_useTC) {
			tc::splitTransposeH(k, n, reinterpret_cast<float*>(m_smoothH.get()), m_ldH, htHi, htLo, m_ldHt, m_stream);
			m_launches += 1;
			operandChangedH(m_smoothH.get());
		}
		productVHt(m_smoothH.get(), m_ldH);                                                 // V H~^T     nsNMF.h:211
		multiplicativeW(m_B.get());                                                         // nsNMF.h:212-217
	}
	if (err) resolveError(n);
}

// ---- GDCLS / ALS / ACLS / AHCLS -------------------------------------------------------------------------------
template <typename T>
void Engine<T>::iterateLS(bool err) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	const NmfAlgorithm algo = m_cfg.algorithm;
	const AlgorithmParams& p = m_cfg.params;
	const bool multi = m_cfg.comm && m_cfg.comm->worldSize() > 1;
	T betaW = 0, betaH = 0;
	if (algo == NmfAlgorithm::AHCLS) {  // AHCLS.h:81-84
		betaW = (T)((1 - p.alphaW) * std::sqrt((double)k) + p.alphaW); betaW *= betaW;
		betaH = (T)((1 - p.alphaH) * std::sqrt((double)k) + p.alphaH); betaH *= betaH;
	}

	// ---- H <- max(0, (W^T W + C_H)^-1 W^T V)
	gramW(m_W[m_wCur].get(), m_G.get());
	if (err) CUDA_CHECK(cudaMemcpyAsync(m_Gsaved.get(), m_G.get(), (size_t)k * k * sizeof(T), cudaMemcpyDeviceToDevice, m_stream));
	if (algo == NmfAlgorithm::GDCLS) kern::addConstraint<T>(k, m_G.get(), T(0), (T)p.lambda, m_stream);
	else if (algo == NmfAlgorithm::ACLS) kern::addConstraint<T>(k, m_G.get(), T(0), (T)p.lambdaH, m_stream);
	else if (algo == NmfAlgorithm::AHCLS) kern::addConstraint<T>(k, m_G.get(), -(T)p.lambdaH, (T)p.lambdaH * betaH - (T)p.lambdaH, m_stream);
	kern::qrFactor<T>(k, m_G.get(), m_qr.get(), m_stream, n >= 4 * k ? m_inverse.get() : nullptr, m_qrWork.get());
	productWtV(m_W[m_wCur].get());
	T* H = m_H[m_hCur].get();
	kern::sumSplits<T>(k, n, m_Npart.get(), m_ldH, m_splitsN, m_strideN, H, m_ldH, m_stream, m_slotsN, false, m_corrN);
	kern::qrSolveClamp<T>(k, m_qr.get(), H, m_ldH, n, false, m_stream, n >= 4 * k ? m_inverse.get() : nullptr);
	m_launches += 4;
	if (m_useTC) {
		tc::splitTransposeH(k, n, reinterpret_cast<float*>(H), m_ldH, m_HtHi.get(), m_HtLo.get(), m_ldHt, m_stream);
		m_launches += 1;
		operandChangedH(H);
	}

	gramH(H, m_ldH, m_B.get());
	if (err) {
		kern::traceKK<T>(k, m_B.get(), m_Gsaved.get(), m_partK.get(), m_stream);            // GDCLS.h:216-227, AHCLS.h:217-224
		m_launches += 1;
	}

	T* Psum = m_Ppart.get() + m_strideP * m_splitsP;
	if (algo == NmfAlgorithm::GDCLS) {
		// ---- W by the multiplicative rule; residual term from P and the NEW W (GDCLS.h:236-264)
		if (!m_cfg.constantW) {
			productVHt(H, m_ldH);
			if (err && !multi) {  // multiplicativeW consumes the partials; keep a summed copy for the trace
				kern::sumSplits<T>(m, k, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, Psum, m_ldW, m_stream, m_slotsP, true, m_corrP);
				m_launches += 1;
			}
			multiplicativeW(m_B.get());
		} else if (err) {
			// constant W: the reference reads a V H^T that only the skipped W update would have written (GDCLS.h:260-264, stale
			// deviceMR: undefined); the term is formed from the actual product instead
			productVHt(H, m_ldH);
			kern::sumSplits<T>(m, k, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, Psum, m_ldW, m_stream, m_slotsP, true, m_corrP);
			if (multi) m_cfg.comm->allReduceSum(Psum, m_strideP, m_stream);
			m_launches += 1;
		}
		if (err) {
			kern::columnDots<T>(m, k, Psum, m_ldW, m_W[m_wCur].get(), m_ldW, m_partN.get(), m_stream);
			m_launches += 1;
		}
	} else {
		// ---- W <- max(0, V H^T (H H^T + C_W)^-1), residual term from W BEFORE the update (AHCLS.h:226-284)
		if (!m_cfg.constantW) {
			if (algo == NmfAlgorithm::ACLS) kern::addConstraint<T>(k, m_B.get(), T(0), (T)p.lambdaW, m_stream);
			else if (algo == NmfAlgorithm::AHCLS) kern::addConstraint<T>(k, m_B.get(), -(T)p.lambdaW, (T)p.lambdaW * betaW - (T)p.lambdaW, m_stream);
			kern::qrFactor<T>(k, m_B.get(), m_qr.get(), m_stream, m >= 4 * k ? m_inverse.get() : nullptr, m_qrWork.get());
			productVHt(H, m_ldH);
			T* Wnext = m_W[1 - m_wCur].get();
			kern::sumSplits<T>(m, k, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, Wnext, m_ldW, m_stream, m_slotsP, true, m_corrP);
			m_launches += 2;
			if (multi) m_cfg.comm->allReduceSum(Wnext, m_strideP, m_stream);
			if (err) {
				kern::columnDots<T>(m, k, m_W[m_wCur].get(), m_ldW, Wnext, m_ldW, m_partN.get(), m_stream);
				m_launches += 1;
			}
			kern::qrSolveClamp<T>(k, m_qr.get(), Wnext, m_ldW, m, true, m_stream, m >= 4 * k ? m_inverse.get() : nullptr);
			m_wCur = 1 - m_wCur;
			const unsigned blocks = kern::columnSquares<T>(m, k, m_W[m_wCur].get(), m_ldW, m_colSqPartials.get(), m_stream);
			m_launches += 2;
			normaliseW(blocks);
		} else if (err) {
			// reference quirk (AHCLS.h:259-269 with a constant W): W itself stands in for V H^T
			kern::columnDots<T>(m, k, m_W[m_wCur].get(), m_ldW, m_W[m_wCur].get(), m_ldW, m_partN.get(), m_stream);
			m_launches += 1;
		}
	}
	if (err) resolveError(k);
}

// D2H of the partial sums, then the reference's host-side combine (FrobeniusResolver.cpp:30-51):
// ascending sort of each array, interleaved accumulation in double, sqrt.
template <typename T>
void Engine<T>::resolveError(unsigned secondLen) {
	const unsigned k = m_cfg.k;
	const bool multi = m_cfg.comm && m_cfg.comm->worldSize() > 1;
	const double mn = m_cfg.comm ? (double)m_cfg.m * (double)m_cfg.comm->globalColumns() : (double)m_cfg.m * (double)m_cfg.n;
	if (m_fused && m_peers.world > 1) {
		// Row blocks: every rank holds the fp64 sums of all ranks' per-column terms (fused::traceSumPush; the wait in finishH
		// made sure they have landed), so the residual needs no collective: sum of the column norms of V over all ranks
		// - 2 (sums in rank order) + the k trace terms in the reference's order.  The same bits on every rank.
		CUDA_CHECK(cudaMemcpyAsync(m_hostTrace.get(), m_sym + m_lay.trace, m_peers.world * sizeof(double), cudaMemcpyDeviceToHost, m_stream));
		CUDA_CHECK(cudaMemcpyAsync(m_hostThird.get(), m_partK.get(), k * sizeof(T), cudaMemcpyDeviceToHost, m_stream));
		checkDeviceFlags();   // synchronises the stream
		T* third = m_hostThird.get();
		std::sort(third, third + k);
		double acc = m_vtvTotal;
		for (unsigned g = 0; g < m_peers.world; ++g) acc -= 2.0 * m_hostTrace.get()[g];
		for (unsigned j = 0; j < k; ++j) acc += third[j];
		m_frobenius = std::sqrt(acc);
		m_rmsd = m_frobenius / std::sqrt(mn);
		return;
	}
	// the per-column terms arrive sorted: a radix sort on the device instead of a std::sort of n floats on the host
	const bool deviceSort = secondLen >= 2048;
	const T* secondSource = m_partN.get();
	if (deviceSort) {
		if (m_sortedSecond.count() < secondLen) m_sortedSecond.allocate(secondLen);
		const size_t needed = kern::sortAscending<T>(m_partN.get(), m_sortedSecond.get(), secondLen, m_sortTemp.get(), m_sortTemp.count(), m_stream);
		if (needed > m_sortTemp.count()) {
			m_sortTemp.allocate(needed);
			kern::sortAscending<T>(m_partN.get(), m_sortedSecond.get(), secondLen, m_sortTemp.get(), m_sortTemp.count(), m_stream);
		}
		secondSource = m_sortedSecond.get();
		m_launches += 2;
	}
	CUDA_CHECK(cudaMemcpyAsync(m_hostSecond.get(), secondSource, secondLen * sizeof(T), cudaMemcpyDeviceToHost, m_stream));
	CUDA_CHECK(cudaMemcpyAsync(m_hostThird.get(), m_partK.get(), k * sizeof(T), cudaMemcpyDeviceToHost, m_stream));
	checkDeviceFlags();   // synchronises the stream
	T* second = m_hostSecond.get();
	T* third = m_hostThird.get();
	if (!deviceSort) std::sort(second, second + secondLen);
	std::sort(third, third + k);
	double acc = 0.0;
	if (!multi) {
		const size_t len = std::max<size_t>(m_vtvSorted.size(), std::max<size_t>(secondLen, k));
		for (size_t j = 0; j < len; ++j) {
			if (j < m_vtvSorted.size()) acc += m_vtvSorted[j];
			if (j < secondLen) acc -= 2.f * second[j];
			if (j < k) acc += third[j];
		}
	} else {
		// column shards: the per-column terms are local, the k x k trace terms are replicated
		const bool secondIsLocal = m_cfg.algorithm == NmfAlgorithm::Multiplicative || m_cfg.algorithm == NmfAlgorithm::nsNMF;
		double local = 0.0;
		const size_t len = std::max<size_t>(m_vtvSorted.size(), secondIsLocal ? secondLen : 0);
		for (size_t j = 0; j < len; ++j) {
			if (j < m_vtvSorted.size()) local += m_vtvSorted[j];
			if (secondIsLocal && j < secondLen) local -= 2.f * second[j];
		}
		acc = m_cfg.comm->allReduceSumHost(local);
		for (size_t j = 0; j < k; ++j) {
			if (!secondIsLocal && j < secondLen) acc -= 2.f * second[j];
			acc += third[j];
		}
	}
	m_frobenius = std::sqrt(acc);
	m_rmsd = m_frobenius / std::sqrt(mn);
}

template <typename T>
void Engine<T>::iterate(bool computeError) {
	switch (m_cfg.algorithm) {
	case NmfAlgorithm::Multiplicative: m_fused ? iterateMUFused(computeError) : iterateMU(computeError); break;
	case NmfAlgorithm::nsNMF: iterateNsNMF(computeError); break;
	case NmfAlgorithm::GDCLS:
	case NmfAlgorithm::ALS:
	case NmfAlgorithm::ACLS:
	case NmfAlgorithm::AHCLS: iterateLS(computeError); break;
	default: throw EngineError(ResultType::ErrorInvalidArgument, "unknown algorithm");
	}
}

// Batches of iterations without residual go through a CUDA graph of TWO iterations (after two, the W/H ping-pong
// buffers are back where they started, so the recorded pointers stay valid); one graph per buffer parity.
template <typename T>
void Engine<T>::iterateNoError(unsigned count) {
	static const bool graphs = [] {
		const char* e = getenv("NMFGPU_GRAPHS");
		return e == nullptr || atoi(e) != 0;
	}();
	// (the in-stream profiler records events, which a capture cannot hold)
	const bool collectivesOk = m_fused || m_cfg.comm == nullptr || m_cfg.comm->capturable();   // the fused iteration calls no collective
	if (graphs && collectivesOk && !m_hostLockstep && count >= 4 && !m_profile && getenv("NMFGPU_TC_DEBUG") == nullptr) {
		cudaGraphExec_t& exec = m_graphExec[m_wCur][m_hCur];
		if (exec == nullptr) {
			iterate(false);   // an eager pair first: lazily set kernel attributes must not happen inside the capture
			iterate(false);
			count -= 2;
			const unsigned long long before = m_launches;
			cudaGraph_t graph = nullptr;
			CUDA_CHECK(cudaStreamBeginCapture(m_stream, cudaStreamCaptureModeThreadLocal));
			try {
				iterate(false);
				iterate(false);
			} catch (...) {
				cudaStreamEndCapture(m_stream, &graph);
				if (graph) cudaGraphDestroy(graph);
				throw;
			}
			CUDA_CHECK(cudaStreamEndCapture(m_stream, &graph));
			m_graphLaunches = m_launches - before;
			m_launches = before;   // recorded, not executed
			const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
			cudaGraphDestroy(graph);
			CUDA_CHECK(e);
		}
		while (count >= 2) {
			CUDA_CHECK(cudaGraphLaunch(exec, m_stream));
			m_launches += m_graphLaunches;
			count -= 2;
		}
	}
	for (unsigned i = 0; i < count; ++i) iterate(false);
}

template <typename T>
void Engine<T>::store(const MatrixDescription<T>& hostW, const MatrixDescription<T>& hostH) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	if (hostW.format != StorageFormat::Dense || hostH.format != StorageFormat::Dense)
		throw EngineError(ResultType::ErrorInvalidArgument, "output matrices must be dense");
	if (m_fused) {
		storeFused(hostW, hostH);
		return;
	}
	const T* W = m_W[m_wCur].get();
	if (m_cfg.algorithm == NmfAlgorithm::nsNMF) {  // the returned basis is W S (nsNMF.h:221-225)
		kern::smoothRight<T>(m, k, W, m_ldW, m_smoothW.get(), m_ldW, (T)m_cfg.params.theta, m_stream);
		W = m_smoothW.get();
	}
	if (hostW.dense.values != nullptr)
		CUDA_CHECK(cudaMemcpy2DAsync(hostW.dense.values, (size_t)hostW.dense.leadingDimension * sizeof(T), W, m_ldW * sizeof(T), (size_t)m * sizeof(T), k,
		                             cudaMemcpyDeviceToHost, m_stream));
	if (hostH.dense.values != nullptr)
		CUDA_CHECK(cudaMemcpy2DAsync(hostH.dense.values, (size_t)hostH.dense.leadingDimension * sizeof(T), m_H[m_hCur].get(), m_ldH * sizeof(T),
		                             (size_t)k * sizeof(T), n, cudaMemcpyDeviceToHost, m_stream));
	measureSparsity(W, m_ldW, m_H[m_hCur].get(), m_ldH);
	synchronize();
	finishSparsity();
}

// Sparseness of the factors that store() hands out (engine.h).  W is whole on every rank; H is this rank's columns, its two
// norms are summed over the ranks.
constexpr unsigned kSparsityBlocks = 128;
template <typename T>
void Engine<T>::measureSparsity(const T* W, size_t ldw, const T* H, size_t ldh) {
	if (m_sparsityPartials.get() == nullptr) {
		m_sparsityPartials.allocate(4 * kSparsityBlocks);
		m_hostSparsity.allocate(4 * kSparsityBlocks);
	}
	kern::absSquareSums<T>(m_cfg.m, m_cfg.k, W, ldw, m_sparsityPartials.get(), kSparsityBlocks, m_stream);
	kern::absSquareSums<T>(m_cfg.k, m_cfg.n, H, ldh, m_sparsityPartials.get() + 2 * kSparsityBlocks, kSparsityBlocks, m_stream);
	CUDA_CHECK(cudaMemcpyAsync(m_hostSparsity.get(), m_sparsityPartials.get(), 4 * kSparsityBlocks * sizeof(double), cudaMemcpyDeviceToHost, m_stream));
	m_launches += 2;
}

template <typename T>
void Engine<T>::finishSparsity() {
	const double* p = m_hostSparsity.get();
	double sums[4] = {0.0, 0.0, 0.0, 0.0};   // |W|_1, |W|_2^2, |H|_1, |H|_2^2
	for (unsigned f = 0; f < 2; ++f)
		for (unsigned b = 0; b < kSparsityBlocks; ++b) {
			sums[2 * f] += p[2 * (f * kSparsityBlocks + b)];
			sums[2 * f + 1] += p[2 * (f * kSparsityBlocks + b) + 1];
		}
	double columns = (double)m_cfg.n;
	if (m_cfg.comm != nullptr && m_cfg.comm->worldSize() > 1) {
		sums[2] = m_cfg.comm->allReduceSumHost(sums[2]);
		sums[3] = m_cfg.comm->allReduceSumHost(sums[3]);
		columns = (double)m_cfg.comm->globalColumns();
	}
	auto hoyer = [](double l1, double l2sq, double count) {
		if (count <= 1.0) return 0.0;
		if (l2sq <= 0.0) return 1.0;   // nothing but zeros
		const double root = std::sqrt(count);
		return (root - l1 / std::sqrt(l2sq)) / (root - 1.0);
	};
	m_sparsityW = hoyer(sums[0], sums[1], (double)m_cfg.m * (double)m_cfg.k);
	m_sparsityH = hoyer(sums[2], sums[3], (double)m_cfg.k * columns);
}

template <typename T>
void Engine<T>::debugProducts(T* wtv, T* vht, float* msWtV, float* msVHt, cudaEvent_t e0, cudaEvent_t e1) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	if (m_fused) {
		tc::Plan& plan = m_tc->plan;
		if (m_peers.world > 1 && (wtv != nullptr || vht != nullptr))
			throw EngineError(ResultType::ErrorInvalidArgument, "with row blocks over several ranks the products can only be timed");
		if (wtv != nullptr || msWtV != nullptr) {
			CUDA_CHECK(cudaEventRecord(e0, m_stream));
			productWtVFused();
			CUDA_CHECK(cudaEventRecord(e1, m_stream));
			CUDA_CHECK(cudaEventSynchronize(e1));
			if (msWtV) CUDA_CHECK(cudaEventElapsedTime(msWtV, e0, e1));
			if (wtv) {
				float* sum = reinterpret_cast<float*>(m_H[1 - m_hCur].get());   // spare H buffer as the landing zone
				fused::prepH(m_peers, m_lay, k, plan.center, m_statSum.get(), reinterpret_cast<float*>(m_Gsaved.get()), m_inv.get(), plan.corrN, m_stream);
				fused::collectN(m_peers, m_lay, k, m_c0, m_nOwn, m_colsPerRank, m_ldH, m_slotsPerRank, plan.wtv.slotCount, m_inv.get(), plan.corrN, sum, m_ldH, m_stream);
				CUDA_CHECK(cudaMemcpy2DAsync(wtv, (size_t)k * sizeof(T), sum, m_ldH * sizeof(T), (size_t)k * sizeof(T), n, cudaMemcpyDeviceToHost, m_stream));
				synchronize();
			}
		}
		if (vht != nullptr || msVHt != nullptr) {
			CUDA_CHECK(cudaEventRecord(e0, m_stream));
			tc::gemmVHt(plan, m_PpartR.get(), m_ldPr, m_stridePr, m_stream);
			CUDA_CHECK(cudaEventRecord(e1, m_stream));
			CUDA_CHECK(cudaEventSynchronize(e1));
			if (msVHt) CUDA_CHECK(cudaEventElapsedTime(msVHt, e0, e1));
			if (vht) {
				T* sum = m_W[1 - m_wCur].get();
				kern::sumSplits<T>(m, k, reinterpret_cast<const T*>(m_PpartR.get()), m_ldPr, m_splitsPr, m_stridePr, sum, m_ldW, m_stream, plan.vht.slotCount, true,
				                   reinterpret_cast<const T*>(plan.corrP));
				CUDA_CHECK(cudaMemcpy2DAsync(vht, (size_t)m * sizeof(T), sum, m_ldW * sizeof(T), (size_t)m * sizeof(T), k, cudaMemcpyDeviceToHost, m_stream));
				synchronize();
			}
		}
		return;
	}
	if (wtv != nullptr || msWtV != nullptr) {
		CUDA_CHECK(cudaEventRecord(e0, m_stream));
		productWtV(m_W[m_wCur].get());
		CUDA_CHECK(cudaEventRecord(e1, m_stream));
		CUDA_CHECK(cudaEventSynchronize(e1));
		if (msWtV) CUDA_CHECK(cudaEventElapsedTime(msWtV, e0, e1));
		if (wtv) {
			T* sum = m_H[1 - m_hCur].get();  // spare H buffer as the landing zone
			kern::sumSplits<T>(k, n, m_Npart.get(), m_ldH, m_splitsN, m_strideN, sum, m_ldH, m_stream, m_slotsN, false, m_corrN);
			CUDA_CHECK(cudaMemcpy2DAsync(wtv, (size_t)k * sizeof(T), sum, m_ldH * sizeof(T), (size_t)k * sizeof(T), n, cudaMemcpyDeviceToHost, m_stream));
			synchronize();
		}
	}
	if (vht != nullptr || msVHt != nullptr) {
		CUDA_CHECK(cudaEventRecord(e0, m_stream));
		productVHt(m_H[m_hCur].get(), m_ldH);
		CUDA_CHECK(cudaEventRecord(e1, m_stream));
		CUDA_CHECK(cudaEventSynchronize(e1));
		if (msVHt) CUDA_CHECK(cudaEventElapsedTime(msVHt, e0, e1));
		if (vht) {
			T* sum = m_Ppart.get() + m_strideP * m_splitsP;
			kern::sumSplits<T>(m, k, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, sum, m_ldW, m_stream, m_slotsP, true, m_corrP);
			CUDA_CHECK(cudaMemcpy2DAsync(vht, (size_t)m * sizeof(T), sum, m_ldW * sizeof(T), (size_t)m * sizeof(T), k, cudaMemcpyDeviceToHost, m_stream));
			synchronize();
		}
	}
}

template class Engine<float>;
template class Engine<double>;

}  // namespace b200
}  // namespace nmfgpu
