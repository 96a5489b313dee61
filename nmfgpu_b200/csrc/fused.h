// fused.h -- the non-GEMM kernels of the MU iteration on the tensor-core path, one GPU or row blocks over several
// (dist.h).  Together with the two tcgen05 products (tc_gemm.h) one iteration is SIX launches, on one GPU and over several:
//
//   1  tc::gemmWtV        partial W_g^T V_g per stream-K slot
//   1b pushN              (several ranks only) the slots summed, each column stored into the memory of the rank that owns it
//                         (NVLink peer stores in whole 256-byte rows: the reduce-scatter); extra CTAs of the same launch sum
//                         the block partials of the last W update and store this rank's W statistics to every rank (step 6
//                         folded in); the last block signals "my partials and the statistics of my rows of W are out"
//   2  updateH            every block: waits for every rank's signal, turns the statistics of the UN-NORMALISED W (Gram
//                         matrix + column sums, summed over ranks) into the column scales 1/||w_c||, W^T W of the
//                         unit-column matrix and the centring term of W^T V; then its panel of the own columns: adds up
//                         the partials of all ranks, (W^T W) H, H <- H o N / (D + eps), residual term; the new columns
//                         stay here, their transposed TF32 split goes to EVERY rank (peer stores: the all-gather is the epilogue);
//                         per-block Gram and row sums of the new columns
//   3  reducePush         block partials -> this rank's H statistics, stored to every rank; last block signals "H is out"
//   4  tc::gemmVHt        V_g H^T: the rank's own rows, no reduction across ranks.  Its B-operand producer waits for every
//                         rank's H signal before the first load of H^T (tc::Gate); the V pipeline fills meanwhile
//   5  updateW            every block: H H^T and the centring term of V H^T from all ranks' H statistics; then
//                         W_g <- (W_g/||.||) o P / ((W_g/||.||) (H H^T) + eps) with the column scale applied when W is READ,
//                         so there is no normalisation pass (MU.h:247, KernelNormalizeColumns.cu); writes the new
//                         un-normalised rows and their TF32 split; per-block Gram and column sums of the new rows
//   6  reducePush         (one rank only) block partials -> the W statistics
//
// Replaces, fused: cublasSsyrk/Ssymm G1, G2, G4, G5, multiplyDivide, normalizeColumns, traceMultiplication of the
// reference's MU iteration (MU.h:164-248).  The Gram products are SIMT fp32 with fixed-order sums (exact, deterministic).
// Nothing here calls NCCL; no kernel exists only to signal or to wait; with one rank the signals and waits vanish and
// "every rank" is this GPU.
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
	size_t trace = 0;                // [world] doubles: every rank's sum of the per-column residual terms (residual iterations)
	size_t H = 0, HtHi = 0, HtLo = 0;
	size_t slots = 0;                // partial products of W^T V.  One rank: [slots][ldh * n], written by the tensor-core kernel;
	                                 // several: [world][ldh * colsPerRank], one partial per rank for the own columns (pushN)
	size_t bytes = 0;
	unsigned statLen = 0;
};

// local (not exchanged) device state of the protocol
struct Control {
	unsigned* epoch = nullptr;     // iteration counter of the protocol: advanced by the last block of pushN
	unsigned* error = nullptr;     // set when a wait for another rank timed out
	unsigned* tickets = nullptr;   // [4] grid-completion counters of the signalling kernels (pushN: 0, reducePush: 1, 2)
};
constexpr size_t kNoSignal = ~(size_t)0;

// opt-in to > 48 KB of dynamic shared memory for the kernels below on the current device (call at setup, not in a capture)
void configure();

// Outside the iteration (store of the factors, diagnostics): the statistics part of updateH as a kernel of its own.
// G (k x k), inv (k), corrN (k), statSum (k*k + k + 1 scratch) are local device buffers.
void prepH(const Peers& peers, const Layout& lay, unsigned k, float center, float* statSum, float* G, float* inv, float* corrN, cudaStream_t stream);

// step 1b.  localSlots: [slots][ldh * N] partial products of this rank's row block (slotCount per 128-column tile).
// Extra CTAs of the same launch do what reducePush does for the W statistics (statPartials: [statBlocks][statCount] block
// partials of updateW, *statFlag the flag updateW left on the device) -- over several ranks the iteration has no separate
// launch for them.
void pushN(const Peers& peers, const Layout& lay, const Control& ctl, unsigned kp, unsigned N, unsigned colsPerRank, size_t ldh, const float* localSlots,
           size_t localStride, const unsigned char* slotCount, const float* statPartials, unsigned statBlocks, unsigned statCount, const float* statFlag,
           cudaStream_t stream);

// residual iterations over several ranks, between updateH and the reducePush that signals: the fp64 sum of tracePartials
// (count = own columns) stored to lay.trace[rank] on every rank
void traceSumPush(const Peers& peers, const Layout& lay, const float* tracePartials, unsigned count, cudaStream_t stream);

// columns per block of updateH for a rank that owns nOwn columns (statPart needs max(1, ceil(nOwn / that)) * (k*k + k) floats)
unsigned panelColumnsH(unsigned nOwn);

// step 2.  Columns [c0, c0 + nOwn) of the k x N matrix H are this rank's.  slotCount[t]: partial products per rank of
// the 128-column tile t (global column index), nullptr = one per rank.  G (k x k), inv (k), corrN (k): written by block 0
// for the trace term, the W update and the store.  tracePartials (nOwn, or nullptr), statPart ([blocks][k*k + k]).
// Returns the number of blocks (= partials in statPart; at least one).
unsigned updateH(const Peers& peers, const Layout& lay, const Control& ctl, float center, unsigned k, unsigned c0, unsigned nOwn, unsigned colsPerRank,
                 size_t ldh, size_t ldht, unsigned slotsPerRank, const unsigned char* slotCount, float* G, float* inv, float* corrN, float eps,
                 float* tracePartials, float* statPart, cudaStream_t stream);

// steps 3 and 6.  out[x] = sum over the blocks of partials[b * count + x] (fixed order) -> (float*)(base[g] + dstOffset)[rank * statLen + x]
// on every rank g; flag >= 0 is stored behind the sums (index count).  signalFlags != kNoSignal: the last block signals
// the current epoch on that flag array (ticket: which of ctl.tickets counts the blocks).
void reducePush(const Peers& peers, size_t dstOffset, unsigned statLen, const float* partials, unsigned blocks, unsigned count, float flag,
                size_t signalFlags, const Control& ctl, unsigned ticket, cudaStream_t stream);

// H H^T (B, k x k) and the centring term of V H^T (corrP, k) from all ranks' H statistics, after waiting for their H
// signals -- as a kernel of its own: residual iterations (the trace term needs B before the W update) and a constant W
void finishH(const Peers& peers, const Layout& lay, const Control& ctl, unsigned k, float center, float* B, float* corrP, cudaStream_t stream);

// step 5 on `rows` rows (W, Whi, Wlo point at the first of them; in place).  P: partial products of V H^T (ldp, slot stride,
// slotCount per 128-row tile).  B (k x k) and corrP (k) are written by block 0.  update = false: no update, only the
// statistics of W as it is (initial factors).  Returns the number of blocks (= partials in statPart, [blocks][k*k + k]).
// statFlag (device, may be nullptr): set to 1 (update) / 0 (initial factors) for the launch that pushes these statistics.
unsigned updateW(const Peers& peers, const Layout& lay, float center, unsigned rows, unsigned k, float* B, float* corrP, const float* inv, float* W, size_t ldw,
                 float* Whi, float* Wlo, const float* Ppart, size_t ldp, size_t slotStride, const unsigned char* slotCount, float eps, float* statPart,
                 float* statFlag, bool update, cudaStream_t stream);

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
