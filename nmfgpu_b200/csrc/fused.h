// fused.h -- the non-GEMM kernels of the MU iteration on the tensor-core path, one GPU or row blocks over several
// (dist.h).  Together with the two tcgen05 products (tc_gemm.h) one iteration is eight launches:
//
//   1  tc::gemmWtV        partial W_g^T V_g per stream-K slot
//   1b pushN              (several ranks only) the slots summed, each column stored into the memory of the rank that owns it
//                         (NVLink peer stores in whole 256-byte rows: the reduce-scatter)
//   2  prepH              signal "my partials and statistics are out", wait for every rank's signal; then the statistics
//                         of the UN-NORMALISED W (Gram matrix + column sums, summed over ranks) become the column scales
//                         1/||w_c||, W^T W of the unit-column matrix and the centring term of W^T V
//   3  updateH            owner columns: adds up the partials of all ranks and slots, (W^T W) H, H <- H o N / (D + eps),
//                         residual term; the new columns, their transposed TF32 split go to EVERY rank (peer stores:
//                         the all-gather is the epilogue); per-block Gram and row sums of the new columns
//   4  reducePush         block partials -> this rank's H statistics, stored to every rank
//   5  finishH            signal / wait, H H^T and the centring term of V H^T from all ranks' statistics
//   6  tc::gemmVHt        V_g H^T: the rank's own rows, no reduction across ranks
//   7  updateW            W_g <- (W_g/||.||) o P / ((W_g/||.||) (H H^T) + eps) with the column scale applied when W is READ,
//                         so there is no normalisation pass (MU.h:247, KernelNormalizeColumns.cu); writes the new
//                         un-normalised rows and their TF32 split; per-block Gram and column sums of the new rows
//   8  reducePush         block partials -> this rank's W statistics, stored to every rank
//
// Replaces, fused: cublasSsyrk/Ssymm G1, G2, G4, G5, multiplyDivide, normalizeColumns, traceMultiplication of the
// reference's MU iteration (MU.h:164-248).  The Gram products are SIMT fp32 with fixed-order sums (exact, deterministic).
// Nothing here calls NCCL; with one rank the signals and waits vanish and "every rank" is this GPU.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace nmfgpu {
namespace b200 {
namespace fused {

constexpr unsigned kMaxRanks = 8;

// every rank's exchange buffer as mapped into this process (dist.h openPeers); base[rank] is the local one
struct Peers {
	unsigned world = 1, rank = 0;
	char* base[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// byte offsets inside an exchange buffer (identical on every rank)
struct Layout {
	size_t flagsN = 0, flagsH = 0;   // kMaxRanks unsigned each: epoch of the last signal of every rank
	size_t statW = 0, statH = 0;     // [world][statLen] floats: k*k Gram + k sums (+ 1 flag for W: 1 = columns get normalised)
	size_t H = 0, HtHi = 0, HtLo = 0;
	size_t slots = 0;                // partial products of W^T V.  One rank: [slots][ldh * n], written by the tensor-core kernel;
	                                 // several: [world][ldh * colsPerRank], one partial per rank for the own columns (pushN)
	size_t bytes = 0;
	unsigned statLen = 0;
};

// local (not exchanged) device state of the protocol
struct Control {
	unsigned* epoch = nullptr;   // incremented by prepH once per iteration
	unsigned* error = nullptr;   // set when a wait for another rank timed out
};

// opt-in to > 48 KB of dynamic shared memory for the kernels below on the current device (call at setup, not in a capture)
void configure();

// step 2.  G (k x k), inv (k), corrN (k), statSum (k*k + k + 1 scratch) are local device buffers.
// signal = false: only the statistics part (no epoch, no flags), used when W is materialised outside an iteration.
void prepH(const Peers& peers, const Layout& lay, const Control& ctl, unsigned k, float center, float* statSum, float* G, float* inv, float* corrN,
           bool signal, cudaStream_t stream);

// step 1b.  localSlots: [slots][ldh * N] partial products of this rank's row block (slotCount per 128-column tile)
void pushN(const Peers& peers, const Layout& lay, unsigned kp, unsigned N, unsigned colsPerRank, size_t ldh, const float* localSlots, size_t localStride,
           const unsigned char* slotCount, cudaStream_t stream);

// columns per block of updateH for a rank that owns nOwn columns (statPart needs ceil(nOwn / that) * (k*k + k) floats)
unsigned panelColumnsH(unsigned nOwn);

// step 3.  Columns [c0, c0 + nOwn) of the k x N matrix H are this rank's.  slotCount[t]: partial products per rank of
// the 128-column tile t (global column index), nullptr = one per rank.  tracePartials (nOwn, or nullptr), statPart ([blocks][k*k + k]).
// Returns the number of blocks (= partials in statPart).
unsigned updateH(const Peers& peers, const Layout& lay, unsigned k, unsigned c0, unsigned nOwn, unsigned colsPerRank, size_t ldh, size_t ldht,
                 unsigned slotsPerRank, const unsigned char* slotCount, const float* G, const float* inv, const float* corrN, float eps,
                 float* tracePartials, float* statPart, cudaStream_t stream);

// steps 4 and 8.  out[x] = sum over the blocks of partials[b * count + x] (fixed order) -> (float*)(base[g] + dstOffset)[rank * statLen + x]
// on every rank g; flag >= 0 is stored behind the sums (index count).
void reducePush(const Peers& peers, size_t dstOffset, unsigned statLen, const float* partials, unsigned blocks, unsigned count, float flag,
                cudaStream_t stream);

// step 5.  B (k x k) and corrP (k) are local device buffers.
void finishH(const Peers& peers, const Layout& lay, const Control& ctl, unsigned k, float center, float* B, float* corrP, cudaStream_t stream);

// step 7 on `rows` rows (W, Whi, Wlo point at the first of them; in place).  P: partial products of V H^T (ldp, slot stride,
// slotCount per 128-row tile).  update = false: no update, only the statistics of W as it is (initial factors).
// Returns the number of blocks (= partials in statPart, [blocks][k*k + k]).
unsigned updateW(unsigned rows, unsigned k, const float* B, const float* inv, float* W, size_t ldw, float* Whi, float* Wlo, const float* Ppart,
                 size_t ldp, size_t slotStride, const unsigned char* slotCount, const float* corrP, float eps, float* statPart, bool update,
                 cudaStream_t stream);

// out[r + j * ldo] = inv[r] * (sum of all partials of W^T V) + corrN[r] for the own columns j (diagnostics: tests, bench)
void collectN(const Peers& peers, const Layout& lay, unsigned k, unsigned c0, unsigned nOwn, unsigned colsPerRank, size_t ldh, unsigned slotsPerRank,
              const unsigned char* slotCount, const float* inv, const float* corrN, float* out, size_t ldo, cudaStream_t stream);

// block[c * rowsPadded + r] = W[r, c] * inv[c] (zero beyond `rows`): the unit-column rows of this rank, ready for an all-gather
void scaleRows(unsigned rows, unsigned rowsPadded, unsigned k, const float* W, size_t ldw, const float* inv, float* block, cudaStream_t stream);
// W (m x k) <- gathered[(g * k + c) * rowsPadded + r], global row = g * rowsPadded + r
void unpackRows(unsigned m, unsigned k, unsigned rowsPadded, const float* gathered, float* W, size_t ldw, cudaStream_t stream);

}  // namespace fused
}  // namespace b200
}  // namespace nmfgpu
