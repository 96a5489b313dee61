// tc_gemm.h -- the two V-sized contractions of the NMF iteration on the B200 tensor cores.
//
//   gemmWtV :  N  (k x n) = W^T V      replaces cublasSgemm(T,N)  reference MU.h:187-188  (G3 in SURVEY.md 2.2)
//   gemmVHt :  N2 (m x k) = V H^T      replaces cublasSgemm(N,T)  reference MU.h:240-241  (G6)
//
// Both stream the 4 GB matrix V exactly once per call and are HBM-bound by design (k/2 FLOP per byte
// of V).  To stay within fp32 tolerance on TF32 tensor cores they run the 3xTF32 scheme
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   x_hi = rna_tf32(x), x_lo = x - x_hi
// with the split of the big operand (V) done on the fly in registers and handed to the tensor core
// through TENSOR MEMORY (tcgen05.mma with the A operand in TMEM), so V costs one TMA write and one
// shared-memory read per element and nothing else.  The small operands (W, H^T) are pre-split in
// global memory by the kernels that produce them.  See tc_gemm.cu for the kernel anatomy.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace nmfgpu {
namespace b200 {
namespace tc {

// One of the two products, decomposed stream-K style: the (tile, reduction-stage) space is cut into
// equal contiguous ranges, one per CTA of a persistent grid.  A tile (128 columns of V for W^T V,
// 128 rows of V for V H^T) that is covered by several CTAs receives one partial product per CTA, in
// consecutive "slots" of the output; slotCount[tile] says how many.  Everything is static (a function
// of the shape only), so the summation order is deterministic.
struct Product {
	unsigned tiles = 0;            // 128-row tiles of the A operand
	unsigned stagesPerTile = 0;    // reduction stages (32 elements) per tile
	// Reduction chunks: the reduction range is cut into `chunks` pieces of chunkStages stages and the units are ordered
	// chunk-major, so at any time all CTAs read the same piece of the small operand (W for W^T V) and it stays in L2.
	// Without chunks every CTA walks its own part of the whole range, W hi/lo (51 MB at 100 000 x 64) keep getting evicted
	// by the stream of V and W^T V fetched 0.55 GB of W from DRAM per launch (ncu: 26 % misses on the evict_last lines).
	unsigned chunks = 1, chunkStages = 0;
	unsigned grid = 0;             // persistent CTAs
	unsigned maxSlots = 1;         // partial products a consumer may have to add per tile
	unsigned char* slotCount = nullptr;   // device, [tiles]
	size_t slotCountBytes = 0;
	alignas(64) unsigned char mapV[128];  // TMA descriptors (CUtensorMap)
	alignas(64) unsigned char mapBhi[128];
	alignas(64) unsigned char mapBlo[128];
};

struct Plan {
	unsigned m = 0, n = 0, k = 0;
	unsigned kp = 0;               // rank padded to the UMMA N granularity (multiple of 16, <= 128)
	unsigned passes = 3;           // 3 = 3xTF32, 1 = single-pass TF32 (diagnostic)
	unsigned flushStages = 16;
	unsigned prefetchStages = 0;   // L2 prefetch distance of the V tiles, in stages (measured: no gain, >4 hurts)     // reduction stages accumulated inside the tensor core before the fp32 flush
	// Mean centring: the kernels multiply (V - center) instead of V, so the tensor-core accumulators hover around
	// zero instead of growing monotonically.  The tensor core truncates its fp32 accumulator after every MMA; on
	// the all-positive data of an NMF that is a systematic -2.4e-7 per accumulated stage, on centred data it is an
	// unbiased error of a much smaller accumulator.  The omitted rank-one terms
	//     W^T V = W^T (V - c 1 1^T) + c (W^T 1) 1^T ,     V H^T = (V - c 1 1^T) H^T + c 1 (H 1)^T
	// are k numbers each (corrN, corrP), recomputed in fp64 whenever W resp. H changes and added by the consumers.
	float center = 0.f;
	float* corrN = nullptr;        // device, [k]: center * column sums of W
	float* corrP = nullptr;        // device, [k]: center * row sums of H
	double* sumScratch = nullptr;
	Product wtv, vht;
	unsigned long long* trace = nullptr;   // device buffer of the optional kernel timeline (environment NMFGPU_TC_TRACE=<file>)
	~Plan();
};

// The work split of one product for `sms` CTAs, on the host and without a device (what tests/test_stream_k_plan.py checks):
// every segment {cta, 256-row tile, first stage, stages, slot} in execution order (5 words each, at most `capacity` are
// written; the return value is their number), info = {tiles, stages per tile, chunks, stages per chunk, grid} and the
// partial products per 128-wide tile that the consumers will add up.
unsigned enumerateSegments(unsigned rowsA, unsigned kdim, unsigned kp, unsigned sms, unsigned* segments, unsigned capacity, unsigned info[5],
                           unsigned char* slotsPerTile, unsigned tileCapacity);

// fp32 problem shapes the tensor-core path covers (others run the SIMT kernels)
bool shapeSupported(unsigned m, unsigned n, unsigned k, size_t ldV, size_t ldW);

// reduceLenWtV: plan the work split of W^T V for a reduction of that many rows (>= m; the surplus reads TMA zeros)
void makePlan(Plan& plan, unsigned m, unsigned n, unsigned k, const float* V, size_t ldV, const float* Whi, const float* Wlo, size_t ldW,
              const float* HtHi, const float* HtLo, size_t ldHt, bool singlePass, float center, unsigned reduceLenWtV = 0);

// mean of all elements of V (fp64 accumulation; synchronises the stream)
float meanOf(const float* V, unsigned m, unsigned n, size_t ldV, cudaStream_t stream);

// corrN <- center * column sums of W (m x k) ; corrP <- center * row sums of H (k x n).  Call after every change of the
// operand the products will read (the hi/lo copies must stem from the same values).
void refreshCorrectionW(Plan& plan, const float* W, size_t ldW, cudaStream_t stream);
void refreshCorrectionH(Plan& plan, const float* H, size_t ldH, cudaStream_t stream);
// out[c] = sum of the first `rows` entries of column c of W (fp64 accumulation), c < plan.k
void columnSums(Plan& plan, const float* W, unsigned rows, size_t ldW, float* out, cudaStream_t stream);
// out[r] = sum of the first `cols` entries of row r of H (k x cols), r < plan.k
void rowSums(Plan& plan, const float* H, unsigned cols, size_t ldH, float* out, cudaStream_t stream);

// Npart + slot*slotStride (k x n, leading dimension ldn) receives the partial products of W^T V
void gemmWtV(const Plan& plan, float* Npart, size_t ldn, size_t slotStride, cudaStream_t stream);

// device address of the counter of tensor-core barrier waits that timed out (0 = healthy); the engine reads it with the
// residual terms and turns a non-zero value into ErrorExternalLibrary instead of returning garbage with Success
const unsigned* timeoutCounter();
// Row blocks over several ranks (dist.h): H^T is stored into this GPU's memory by the ranks that own its columns, and a
// flag word per rank says up to which iteration.  With a gate the product waits INSIDE the kernel -- only the producer of
// the H^T tiles, before its first load, while the V pipeline fills -- instead of behind a kernel that does nothing but wait.
struct Gate {
	const unsigned* flags = nullptr;   // device, [count]: last epoch every rank has signalled
	const unsigned* epoch = nullptr;   // device: the epoch to wait for
	unsigned* error = nullptr;         // device: set to 1 when a flag did not arrive within 10 s
	unsigned count = 0;
};
// Ppart + slot*slotStride (m x k, leading dimension ldp)
void gemmVHt(const Plan& plan, float* Ppart, size_t ldp, size_t slotStride, cudaStream_t stream, const Gate* gate = nullptr);

// hi/lo TF32 split of H (k x n, column-major) written transposed: Ht[c * ldht + j] = H[c + j * ldh]
void splitTransposeH(unsigned k, unsigned n, const float* H, size_t ldh, float* hi, float* lo, size_t ldht, cudaStream_t stream);

}  // namespace tc
}  // namespace b200
}  // namespace nmfgpu
