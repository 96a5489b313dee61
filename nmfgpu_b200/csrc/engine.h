// engine.h -- device-resident state and iteration sequencing of one factorisation problem.
//
// An Engine<T> owns V, W, H and all scratch on one GPU and knows how to run one iteration of each
// algorithm of include/nmfgpu.h.  It plays the role of the reference's IAlgorithm subclasses
// (source/nmf/Algorithm*.h) but is one class with one buffer plan, a single stream, fused kernels
// and -- for fp32 with k <= 128 -- the tcgen05 tensor-core contractions of tc_gemm.cu.
// The run loop that drives it (error cadence, stop rule, best-run bookkeeping) is driver.cpp.
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <random>
#include <vector>

#include "common.h"
#include "fused.h"
#include "spmm.h"
#include "tc_gemm.h"

namespace nmfgpu {
namespace b200 {

class Communicator;  // dist.h: the ranks of a column-sharded run (nullptr = single GPU)

enum class Precision {
	Auto,     // fp32: tensor cores (3xTF32) when the shape allows, else SIMT; fp64: SIMT
	Exact,    // SIMT FFMA/DFMA only ("fp32-exact mode")
	Tf32x3,   // force the tcgen05 3xTF32 path (error if the shape does not allow it)
	Tf32x1,   // single-pass TF32 (diagnostic: shows why the split is needed)
};

struct AlgorithmParams {
	double lambda = 0.0;                    // GDCLS
	double lambdaW = 0.0, lambdaH = 0.0;    // ACLS / AHCLS
	double alphaW = 0.0, alphaH = 0.0;      // AHCLS
	double theta = 0.0;                     // nsNMF
};

struct EngineConfig {
	NmfAlgorithm algorithm = NmfAlgorithm::Multiplicative;
	unsigned m = 0, n = 0, k = 0;           // n = columns held by THIS rank
	bool constantW = false;
	AlgorithmParams params;
	Precision precision = Precision::Auto;
	Communicator* comm = nullptr;
	bool needsDenseV = false;               // the initialisation reads V as a dense matrix (k-means, mean columns): no sparse execution
};

template <typename T>
class Engine {
public:
	explicit Engine(const EngineConfig& cfg);
	~Engine();

	// Allocates every buffer and ingests V: host dense / CSR / CSC / COO (reference
	// source/common/Matrix.h:145-232) or, when vOnDevice, a dense device pointer that is adopted
	// without a copy.  Also computes the per-column squared norms of V (MU.h:117-125).
	void setup(const MatrixDescription<T>& V, bool vOnDevice);

	// initial factors ------------------------------------------------------------
	void loadW(const MatrixDescription<T>& hostW);            // CopyStrategy.h:41-46
	void loadH(const MatrixDescription<T>& hostH);
	void randomW(unsigned seed);                              // RandomValueStrategy.cpp:29-70
	void randomH(unsigned seed);
	void meanColumnsW(unsigned seed);                         // MeanColumnStrategy.cpp:41-56
	void kmeansW(unsigned seed);                              // KMeansStrategy.cpp:54-58
	void hFromWtV(bool absolute);                             // KMeansStrategy.cpp:35-37 (+ documented |W^T V|)
	void finishInitialisation();                              // derived buffers (TF32 splits) after W/H changed

	// one iteration; if computeError the residual is resolved on the host afterwards
	void iterate(bool computeError);
	double frobenius() const { return m_frobenius; }
	double rmsd() const { return m_rmsd; }
	// Hoyer sparseness (sqrt(N) - |x|_1 / |x|_2) / (sqrt(N) - 1) of the factors last handed out by store(), N = entries of
	// the whole factor (all shards): 0 = all entries equal, 1 = a single non-zero.  The reference declares the two fields
	// of ExecutionRecord and never writes them (SingleGpuDispatcher.cpp:217-222).
	double sparsityW() const { return m_sparsityW; }
	double sparsityH() const { return m_sparsityH; }

	// enqueue `count` iterations without any error computation (bench / session API)
	void iterateNoError(unsigned count);
	void synchronize();

	void store(const MatrixDescription<T>& hostW, const MatrixDescription<T>& hostH);  // MU.h:251-254
	cudaStream_t stream() const { return m_stream; }
	const EngineConfig& config() const { return m_cfg; }
	bool usesTensorCores() const { return m_useTC; }
	bool rowBlocks() const { return m_fused && m_peers.world > 1; }   // column shards regrouped into row blocks (dist.h)
	bool fusedMU() const { return m_fused; }
	bool sparseExecution() const { return m_sparse; }
	unsigned long long kernelLaunches() const { return m_launches; }
	unsigned splitsWtV() const { return m_splitsN; }
	unsigned splitsVHt() const { return m_splitsP; }
	// test / bench hook: both V-sized products for the current factors, summed over slices, to the host
	void debugProducts(T* wtv, T* vht, float* msWtV, float* msVHt, cudaEvent_t e0, cudaEvent_t e1);

	T* deviceV() const { return m_V.get(); }
	size_t ldV() const { return m_ldV; }
	T* deviceW() const { return m_W[m_wCur].get(); }
	size_t ldW() const { return m_ldW; }
	T* deviceH() const { return m_H[m_hCur].get(); }
	size_t ldH() const { return m_ldH; }

private:
	void iterateMU(bool err);
	// MU on the tensor-core path, one GPU or row blocks over several (fused.h, dist.h): six launches per iteration
	bool decideFused();
	void setupFused();
	void finishInitialisationFused();
	void iterateMUFused(bool err);
	void productWtVFused();
	void storeFused(const MatrixDescription<T>& hostW, const MatrixDescription<T>& hostH);
	void releaseFused();
	void checkDeviceFlags();                                   // barrier / peer waits that timed out surface as ErrorExternalLibrary
	void iterateNsNMF(bool err);
	void iterateLS(bool err);
	void resolveError(unsigned secondLen);
	void measureSparsity(const T* W, size_t ldw, const T* H, size_t ldh);   // enqueues the reductions; finishSparsity() after the next synchronize()
	void finishSparsity();

	void gramW(const T* W, T* G);                              // G = W^T W
	void gramH(const T* H, size_t ldh, T* B);                  // B = H H^T (all-reduced over shards)
	void productWtV(const T* W);                               // m_Npart / m_splitsN <- W^T V
	void productVHt(const T* H, size_t ldh);                   // m_Ppart / m_splitsP <- V H^T (all-reduced)
	void normaliseW(unsigned blocks, bool haveColumnSums = false);
	void preReduceN(const T*& N, unsigned& splits, const unsigned char*& slots, const T*& corr);   // many partials of W^T V -> one
	void operandChangedW(const T* W);                          // refresh what the tensor-core products derive from W resp. H
	void operandChangedH(const T* H);
	void multiplicativeW(const T* B);                          // W <- W o P / (W B + eps), normalise

	// NMFGPU_PROFILE_ITERATION=1: CUDA events between the steps of MU iterations run through iterate() (not the graph
	// batches), in-stream durations on stderr when the engine is destroyed.  ncu flushes the caches between kernels, which
	// triples the time of the small kernels; this does not.
	void stamp(const char* what);
	void reportStamps();
	std::vector<std::pair<const char*, cudaEvent_t>> m_stamps;
	bool m_profile = false;

	EngineConfig m_cfg;
	T m_eps;
	cudaStream_t m_stream = nullptr;
	bool m_useTC = false;
	unsigned long long m_launches = 0, m_graphLaunches = 0;
	cudaGraphExec_t m_graphExec[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // two recorded iterations, per (W, H) buffer parity

	size_t m_ldV = 0, m_ldW = 0, m_ldH = 0;
	DeviceBuffer<T> m_V, m_W[2], m_H[2];
	int m_wCur = 0, m_hCur = 0;
	DeviceBuffer<T> m_G, m_Gsaved, m_B, m_kkScratch, m_qr, m_inverse;
	DeviceBuffer<double> m_qrWork;   // fp32: the k x k inverse is formed in fp64 (kernels.h qrFactor)
	DeviceBuffer<T> m_Npart, m_Ppart, m_smoothW, m_smoothH;
	unsigned m_splitsN = 1, m_splitsP = 1, m_splitsGW = 1, m_splitsGH = 1;
	size_t m_strideN = 0, m_strideP = 0;
	const unsigned char* m_slotsN = nullptr;   // per-tile partial counts of the stream-K tensor-core products (tc_gemm.h)
	const unsigned char* m_slotsP = nullptr;
	const T* m_corrN = nullptr;   // rank-one terms of the mean-centred tensor-core products (tc_gemm.h), device, [k]
	const T* m_corrP = nullptr;
	DeviceBuffer<T> m_colSqPartials, m_colSumPartials, m_rowSumPartials, m_colSq;
	DeviceBuffer<T> m_partN, m_partK;
	PinnedBuffer<T> m_hostSecond, m_hostThird;
	std::vector<T> m_vtvSorted;
	double m_vtvSum = 0.0, m_vtvTotal = 0.0;   // sum of the sorted column norms of V: own shard, all ranks
	DeviceBuffer<T> m_sortedSecond;            // the per-column residual terms, sorted on the device before their D2H
	DeviceBuffer<unsigned char> m_sortTemp;
	PinnedBuffer<double> m_hostTrace;
	double m_frobenius = 0.0, m_rmsd = 0.0;
	double m_sparsityW = 0.0, m_sparsityH = 0.0;
	DeviceBuffer<double> m_sparsityPartials;
	PinnedBuffer<double> m_hostSparsity;

	// tensor-core operands (fp32 only): TF32 hi/lo splits of W (m x k) and of H^T (n x k, n contiguous)
	DeviceBuffer<float> m_Whi, m_Wlo, m_HtHi, m_HtLo;
	size_t m_ldHt = 0;
	struct TcPlan;
	std::unique_ptr<TcPlan> m_tc;

	// sparse execution (spmm.h): V stays compressed (CSR + CSC), the two V-sized products are gathers.  Chosen for sparse
	// inputs below 2 % density or too large to densify; NMFGPU_SPARSE=1 / 0 forces / forbids it.
	bool m_sparse = false;
	sparse::DeviceSparse<T> m_S;
	DeviceBuffer<T> m_Wt, m_Pt;   // row-major W (the gather operand of W^T V) and row-major V H^T, leading dimension m_ldH
	DeviceBuffer<int> m_blockPtr; // blocked sweep of W^T V (spmm.h): (m_sparseBlocks + 1) x n entry positions
	unsigned m_sparseBlocks = 1, m_sparseBlockRows = 0;
	void decideSparse(const MatrixDescription<T>& V, bool vOnDevice);

	// fused MU (fused.h).  This rank holds the row block V[I, :], I = [m_r0, m_r0 + m_mr) (all of V on one GPU), updates
	// those rows of W, and the columns [m_c0, m_c0 + m_nOwn) of H (all of them on one GPU).  H, its transposed TF32 split,
	// the partial products of W^T V and the k*k + k statistics live in the exchange buffer m_sym, which every other rank
	// stores into directly.  W stays UN-NORMALISED between iterations: m_inv holds 1 / column norm, applied where W is read.
	bool m_fused = false;
	fused::Peers m_peers;
	fused::Layout m_lay;
	fused::Control m_ctl;
	char* m_sym = nullptr;
	std::vector<void*> m_peerPtrs;
	unsigned m_r0 = 0, m_mr = 0, m_mrPad = 0, m_globalN = 0, m_colsPerRank = 0, m_c0 = 0, m_nOwn = 0, m_slotsPerRank = 1, m_splitsPr = 1;
	size_t m_ldVr = 0, m_ldHtFull = 0, m_ldPr = 0, m_stridePr = 0;
	DeviceBuffer<float> m_Vr, m_Nlocal, m_PpartR, m_statSum, m_inv, m_statPartH, m_statPartW;
	DeviceBuffer<unsigned> m_ctlWords;
	float* m_wStatFlag = nullptr;      // device: 1 when m_statPartW describes an updated W (columns get normalised), 0 for the initial factors
	unsigned m_blocksW = 0;            // block partials in m_statPartW
	bool m_wUpdated = false;           // the host's copy of that flag
	bool m_hostLockstep = false;       // ranks share a GPU (test transport): host barriers where the ranks wait for each other, no graphs
	PinnedBuffer<unsigned> m_hostFlags;
	const float* m_Vblock = nullptr;   // V[I, :]: m_Vr, or V itself on one GPU
};

}  // namespace b200
}  // namespace nmfgpu
