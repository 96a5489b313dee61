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
		cudaEventElapsedTime(&ms, m_stamps[i - 1].second, m_stamps[i].second);
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
			CUDA_CHECK(cudaMemcpy2DAsync(m_V.get(), m_ldV * sizeof(T), V.dense.values, (size_t)V.dense.leadingDimension * sizeof(T),
			                             (size_t)m * sizeof(T), n, cudaMemcpyHostToDevice, m_stream));
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
	m_qr.allocate((size_t)k * k + k);
	m_inverse.allocate((size_t)k * k);

	// tensor-core eligibility: fp32, rank that fits one UMMA N, TMA-compatible strides
	m_useTC = false;
	if (std::is_same<T, float>::value && m_cfg.precision != Precision::Exact && !m_sparse) {
		const bool ok = tc::shapeSupported(m, n, k, m_ldV, m_ldW) && (reinterpret_cast<uintptr_t>(m_V.get()) % 16 == 0);
		if (!ok && (m_cfg.precision == Precision::Tf32x3 || m_cfg.precision == Precision::Tf32x1))
			throw EngineError(ResultType::ErrorInvalidArgument, "tensor-core precision requested for an unsupported shape");
		m_useTC = ok;
	}

	timer.mark("  factor buffers");
	// split-K plans of the two V-sized products and of the Gram products
	if (m_useTC) {
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
	m_Npart.allocate(m_strideN * (m_splitsN + 1));   // +1: landing zone of the pre-reduced product (many partials per tile)
	m_Ppart.allocate(m_strideP * (m_splitsP + 1));  // +1: slot for the summed / all-reduced product
	m_kkScratch.allocate((size_t)k * k * std::max(m_splitsGW, m_splitsGH));
	m_colSqPartials.allocate((size_t)ceilDiv(m, 128) * k);
	m_colSumPartials.allocate((size_t)ceilDiv(m, 128) * k);
	m_rowSumPartials.allocate((size_t)ceilDiv(n, 64) * k);
	m_colSq.allocate(k);
	m_partN.allocate(std::max(n, k));
	m_partK.allocate(k);
	m_hostSecond.allocate(std::max(n, k));
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
	timer.mark("  tr(V^T V)");
	setupRowOwners();
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

// ---- row-owner dataflow for column shards (dist.h) ----------------------------------------------------------
// The all-reduce dataflow repeats the whole W update on every rank and moves 2 x 4mk bytes per rank and iteration.
// Here rank g also keeps the row block V[I_g, :] (one grouped send/recv of the column shards at setup; V is then
// resident twice, 2 x 4mn/G bytes per GPU), so that
//   H update : columns J_g from W^T V[:, J_g] as before, then all-gather of H (4kn bytes),
//   W update : rows I_g from V[I_g, :] H^T -- no reduction --, all-reduce of k*k + k statistics of the un-normalised
//              block (Gram matrix, whose diagonal holds the column norms, and column sums), all-gather of the unit-column
//              blocks (4mk bytes).
// Every W-side kernel works on m/G rows; W^T W falls out of the statistics.  Needs equal column shards.
template <typename T>
void Engine<T>::setupRowOwners() {
	Communicator* comm = m_cfg.comm;
	if (comm == nullptr || comm->worldSize() <= 1) return;
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
	const unsigned G = (unsigned)comm->worldSize(), rank = (unsigned)comm->rank();
	const unsigned N = comm->globalColumns();
	const unsigned mrPad = (unsigned)roundUp(ceilDiv(m, G), 256);
	const char* mode = getenv("NMFGPU_DIST_MODE");
	bool ok = std::is_same<T, float>::value && m_useTC && m_cfg.algorithm == NmfAlgorithm::Multiplicative && !m_cfg.constantW &&
	          !(mode != nullptr && strcmp(mode, "allreduce") == 0) && (unsigned long long)n * G == N && comm->columnOffset() == rank * n &&
	          (unsigned long long)(G - 1) * mrPad < m;
	if (ok) {
		const unsigned r0 = rank * mrPad, mr = std::min(mrPad, m - r0);
		ok = tc::shapeSupported(mr, N, k, roundUp(mr, 32), m_ldW);
	}
	// every rank must take the same path: the collectives of the two dataflows do not match
	if (comm->allReduceSumHost(ok ? 1.0 : 0.0) != (double)G) return;

	m_rowOwners = true;
	m_globalN = N;
	m_mrPad = mrPad;
	m_r0 = rank * mrPad;
	m_mr = std::min(mrPad, m - m_r0);
	m_ldVr = roundUp(m_mr, 32);
	auto rowsOf = [&](unsigned g) { return std::min(mrPad, m - g * mrPad); };

	// ---- V[I_g, :] from the column shards: rank g sends V[I_h, J_g] to every h, packed with the receiver's leading dimension
	m_Vr.allocate(m_ldVr * N);
	m_Vr.zero(m_stream);
	{
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
				CUDA_CHECK(cudaMemcpy2DAsync(m_Vr.get() + m_ldVr * ((size_t)rank * n), m_ldVr * sizeof(float), V + m_r0, m_ldV * sizeof(float),
				                             (size_t)m_mr * sizeof(float), n, cudaMemcpyDeviceToDevice, m_stream));
				continue;
			}
			CUDA_CHECK(cudaMemcpy2DAsync(pack.get() + at, ldh * sizeof(float), V + (size_t)h * mrPad, m_ldV * sizeof(float),
			                             (size_t)rowsOf(h) * sizeof(float), n, cudaMemcpyDeviceToDevice, m_stream));
			sends.push_back({pack.get() + at, ldh * n, (int)h});
			recvs.push_back({m_Vr.get() + m_ldVr * ((size_t)h * n), m_ldVr * n, (int)h});
			at += ldh * n;
		}
		comm->exchange(sends, recvs, m_stream);
		synchronize();
	}

	// ---- operands and plan of V[I_g, :] H^T
	m_ldHtFull = roundUp(N, 32);
	m_Hfull.allocate(m_ldH * N);
	m_HtHiFull.allocate(m_ldHtFull * k);
	m_HtLoFull.allocate(m_ldHtFull * k);
	m_Hfull.zero(m_stream);
	m_HtHiFull.zero(m_stream);
	m_HtLoFull.zero(m_stream);
	m_tcR.reset(new TcPlan());
	const float center = tc::meanOf(m_Vr.get(), m_mr, N, m_ldVr, m_stream);
	tc::makePlan(m_tcR->plan, m_mr, N, k, m_Vr.get(), m_ldVr, m_Whi.get() + m_r0, m_Wlo.get() + m_r0, m_ldW, m_HtHiFull.get(), m_HtLoFull.get(),
	             m_ldHtFull, m_cfg.precision == Precision::Tf32x1, center);
	m_splitsPr = m_tcR->plan.vht.maxSlots;
	m_ldPr = roundUp(m_mr, 32);
	m_stridePr = m_ldPr * k;
	m_PpartR.allocate(m_stridePr * m_splitsPr);
	m_statLen = roundUp((size_t)k * k + k, 32);
	m_stat.allocate(m_statLen);
	m_statPart.allocate(m_statLen);
	m_statGath.allocate(m_statLen * G);
	m_statPart.zero(m_stream);
	m_Wblk.allocate((size_t)mrPad * k);
	m_Wgath.allocate((size_t)mrPad * k * G);
	m_splitsGHfull = kern::effectiveSplits(N, pickGramSplits(ceilDiv(k, 64) * ceilDiv(k, 64), N));
	m_splitsGWrows = kern::effectiveSplits(m_mr, pickGramSplits(ceilDiv(k, 64) * ceilDiv(k, 64), m_mr));
	m_kkScratch.allocate((size_t)k * k * std::max(std::max(m_splitsGW, m_splitsGH), std::max(m_splitsGHfull, m_splitsGWrows)));
}

// Everything the W update derives from H: the full matrix and its transposed TF32 split, H H^T and the centring term.
// The k*k + k statistics of the local columns (H_g H_g^T and the row sums) travel with the all-gather of H in one NCCL
// group and are added up in rank order on every rank: no separate all-reduce and no replicated work on all n columns.
template <typename T>
void Engine<T>::gatherH(bool haveRowSums) {
	const unsigned k = m_cfg.k, n = m_cfg.n, N = m_globalN;
	const unsigned G = (unsigned)m_cfg.comm->worldSize();
	const T* Hloc = m_H[m_hCur].get();
	float* part = m_statPart.get();
	kern::gemmNT<T>(k, n, k, Hloc, m_ldH, Hloc, m_ldH, m_kkScratch.get(), k, m_splitsGH, (size_t)k * k, m_stream);
	kern::sumSplits<T>(k, k, m_kkScratch.get(), k, m_splitsGH, (size_t)k * k, reinterpret_cast<T*>(part), k, m_stream);
	if (haveRowSums) kern::finishPartialSums(k, ceilDiv(n, 64), reinterpret_cast<const float*>(m_rowSumPartials.get()), 1.f, part + (size_t)k * k, m_stream);
	else tc::rowSums(m_tcR->plan, reinterpret_cast<const float*>(Hloc), n, m_ldH, part + (size_t)k * k, m_stream);
	m_cfg.comm->allGatherPair(reinterpret_cast<const float*>(Hloc), m_Hfull.get(), m_ldH * n, part, m_statGath.get(), m_statLen, m_stream);
	tc::splitTransposeH(k, N, m_Hfull.get(), m_ldH, m_HtHiFull.get(), m_HtLoFull.get(), m_ldHtFull, m_stream);
	kern::sumGathered((unsigned)((size_t)k * k + k), G, m_statLen, m_statGath.get(), m_stat.get(), m_stream);
	kern::finishStatsH(k, m_stat.get(), m_tcR->plan.center, reinterpret_cast<float*>(m_B.get()), m_tcR->plan.corrP, m_stream);
	m_launches += 7;
}

template <typename T>
void Engine<T>::iterateMURowOwners(bool err) {
	const unsigned n = m_cfg.n, k = m_cfg.k;
	Communicator* comm = m_cfg.comm;
	// ---- H <- H o (W^T V) / ((W^T W) H + eps) on the columns of this rank (MU.h:164-198); m_G is up to date
	stamp("begin");
	productWtV(m_W[m_wCur].get());
	stamp("product W^T V (own columns)");
	const T* N = m_Npart.get();
	unsigned splits = m_splitsN;
	const unsigned char* slots = m_slotsN;
	const T* corr = m_corrN;
	preReduceN(N, splits, slots, corr);
	kern::updateH<T>(k, n, m_G.get(), m_H[m_hCur].get(), m_H[1 - m_hCur].get(), m_ldH, N, m_ldH, splits, m_strideN, m_eps,
	                 err ? m_partN.get() : nullptr, nullptr, nullptr, m_ldHt, m_stream, slots, corr, m_rowSumPartials.get());
	m_launches += 1;
	m_hCur = 1 - m_hCur;
	stamp("update H");
	gatherH(true);
	stamp("H statistics, all-gather, split");
	if (err) {
		kern::traceKK<T>(k, m_B.get(), m_G.get(), m_partK.get(), m_stream);                 // tr(HH^T W^T W) MU.h:203-216
		m_launches += 1;
	}

	// ---- W <- W o (V H^T) / (W (H H^T) + eps) on the rows of this rank (MU.h:200-248)
	float* stat = m_stat.get();
	tc::gemmVHt(m_tcR->plan, m_PpartR.get(), m_ldPr, m_stridePr, m_stream);
	stamp("product V H^T (own rows)");
	T* Wnext = m_W[1 - m_wCur].get();
	const unsigned wBlocks = kern::updateW<T>(m_mr, k, m_B.get(), m_W[m_wCur].get() + m_r0, Wnext + m_r0, m_ldW, reinterpret_cast<const T*>(m_PpartR.get()),
	                                          m_ldPr, m_splitsPr, m_stridePr, m_eps, m_colSqPartials.get(), m_stream, m_tcR->plan.vht.slotCount,
	                                          reinterpret_cast<const T*>(m_tcR->plan.corrP), m_colSumPartials.get());
	// statistics of the un-normalised block: Gram matrix (its diagonal = the column sums of squares) and column sums.
	// They travel with the all-gather of the (still un-normalised) blocks in one NCCL group; every rank adds them up in
	// rank order and the unpack kernel divides by the norms.
	float* part = m_statPart.get();
	kern::gemmTN<T>(m_mr, k, k, Wnext + m_r0, m_ldW, Wnext + m_r0, m_ldW, m_kkScratch.get(), k, m_splitsGWrows, (size_t)k * k, m_stream);
	kern::sumSplits<T>(k, k, m_kkScratch.get(), k, m_splitsGWrows, (size_t)k * k, reinterpret_cast<T*>(part), k, m_stream);
	kern::finishPartialSums(k, wBlocks, reinterpret_cast<const float*>(m_colSumPartials.get()), 1.f, part + (size_t)k * k, m_stream);
	kern::scalePackRows(m_mr, m_mrPad, k, reinterpret_cast<const float*>(Wnext) + m_r0, m_ldW, nullptr, m_Wblk.get(), m_stream);
	stamp("update W rows, statistics, pack");
	comm->allGatherPair(m_Wblk.get(), m_Wgath.get(), (size_t)m_mrPad * k, part, m_statGath.get(), m_statLen, m_stream);
	stamp("all-gather W + statistics");
	kern::sumGathered((unsigned)((size_t)k * k + k), (unsigned)comm->worldSize(), m_statLen, m_statGath.get(), stat, m_stream);
	kern::finishStats(k, stat, m_tc->plan.center, reinterpret_cast<float*>(m_G.get()), m_tc->plan.corrN, m_stream);
	kern::unpackSplit(m_cfg.m, k, m_mrPad, m_Wgath.get(), reinterpret_cast<float*>(Wnext), m_ldW, m_Whi.get(), m_Wlo.get(), m_stream, stat);
	stamp("norms, unpack W, hi/lo");
	m_launches += 10;
	m_wCur = 1 - m_wCur;
	if (err) resolveError(n);
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
	const unsigned splits = kern::effectiveSplits(m, std::min(m_splitsN, 16u));
	// m_Npart has room for m_splitsN slices; reuse them
	const unsigned use = std::min(splits, m_splitsN);
	kern::gemmTN<T>(m, k, n, m_W[m_wCur].get(), m_ldW, m_V.get(), m_ldV, m_Npart.get(), m_ldH, use, m_strideN, m_stream);
	const unsigned eff = kern::effectiveSplits(m, use);
	kern::sumSplits<T>(k, n, m_Npart.get(), m_ldH, eff, m_strideN, m_H[m_hCur].get(), m_ldH, m_stream);
	if (absolute) kern::absInPlace<T>(k, n, m_H[m_hCur].get(), m_ldH, m_stream);
	else kern::clampNonNegative<T>(k, n, m_H[m_hCur].get(), m_ldH, m_stream);
}

template <typename T>
void Engine<T>::finishInitialisation() {
	if (!m_useTC) return;
	float* W = reinterpret_cast<float*>(m_W[m_wCur].get());
	float* H = reinterpret_cast<float*>(m_H[m_hCur].get());
	kern::splitTf32(m_cfg.m, m_cfg.k, W, m_ldW, m_Whi.get(), m_Wlo.get(), m_ldW, m_stream);
	tc::splitTransposeH(m_cfg.k, m_cfg.n, H, m_ldH, m_HtHi.get(), m_HtLo.get(), m_ldHt, m_stream);
	operandChangedW(m_W[m_wCur].get());
	operandChangedH(m_H[m_hCur].get());
	if (m_rowOwners) {   // what the iteration keeps up to date itself: W^T W and everything derived from the full H
		gramW(m_W[m_wCur].get(), m_G.get());
		gatherH(false);
	}
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
	kern::qrFactor<T>(k, m_G.get(), m_qr.get(), m_stream);
	productWtV(m_W[m_wCur].get());
	T* H = m_H[m_hCur].get();
	kern::sumSplits<T>(k, n, m_Npart.get(), m_ldH, m_splitsN, m_strideN, H, m_ldH, m_stream, m_slotsN, false, m_corrN);
	kern::qrSolveClamp<T>(k, m_qr.get(), H, m_ldH, n, false, m_stream, m_inverse.get());
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
			kern::qrFactor<T>(k, m_B.get(), m_qr.get(), m_stream);
			productVHt(H, m_ldH);
			T* Wnext = m_W[1 - m_wCur].get();
			kern::sumSplits<T>(m, k, m_Ppart.get(), m_ldW, m_splitsP, m_strideP, Wnext, m_ldW, m_stream, m_slotsP, true, m_corrP);
			m_launches += 2;
			if (multi) m_cfg.comm->allReduceSum(Wnext, m_strideP, m_stream);
			if (err) {
				kern::columnDots<T>(m, k, m_W[m_wCur].get(), m_ldW, Wnext, m_ldW, m_partN.get(), m_stream);
				m_launches += 1;
			}
			kern::qrSolveClamp<T>(k, m_qr.get(), Wnext, m_ldW, m, true, m_stream, m_inverse.get());
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
	CUDA_CHECK(cudaMemcpyAsync(m_hostSecond.get(), m_partN.get(), secondLen * sizeof(T), cudaMemcpyDeviceToHost, m_stream));
	CUDA_CHECK(cudaMemcpyAsync(m_hostThird.get(), m_partK.get(), k * sizeof(T), cudaMemcpyDeviceToHost, m_stream));
	synchronize();
	T* second = m_hostSecond.get();
	T* third = m_hostThird.get();
	std::sort(second, second + secondLen);
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
	const double mn = m_cfg.comm ? (double)m_cfg.m * (double)m_cfg.comm->globalColumns() : (double)m_cfg.m * (double)m_cfg.n;
	m_rmsd = m_frobenius / std::sqrt(mn);
}

template <typename T>
void Engine<T>::iterate(bool computeError) {
	switch (m_cfg.algorithm) {
	case NmfAlgorithm::Multiplicative: m_rowOwners ? iterateMURowOwners(computeError) : iterateMU(computeError); break;
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
	if (graphs && count >= 6 && getenv("NMFGPU_TC_DEBUG") == nullptr) {
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
	synchronize();
}

template <typename T>
void Engine<T>::debugProducts(T* wtv, T* vht, float* msWtV, float* msVHt, cudaEvent_t e0, cudaEvent_t e1) {
	const unsigned m = m_cfg.m, n = m_cfg.n, k = m_cfg.k;
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
