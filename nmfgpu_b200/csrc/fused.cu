// fused.cu -- see fused.h.  sm_100a; SIMT fp32 with fixed-order reductions (bit-reproducible from run to run).
#include "fused.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.h"

namespace nmfgpu {
namespace b200 {
namespace fused {

namespace {

// ---- system-scope flag protocol between ranks ---------------------------------------------------------------------
// A rank signals by storing its epoch into a flag word that lives in the RECEIVER's exchange buffer (so waiting spins on
// local memory).  The data the flag announces was written by earlier kernels of the same stream (peer stores); the
// kernel boundary plus the system fence order it before the flag.
// Flags are written and polled with RELAXED system-scope accesses between explicit fences: a st.release.sys per peer is
// a fence per peer, and each of them waits for the previous peer's store to be acknowledged over NVLink -- eight
// sequential round trips, 16 of the 28 us the H-side signal took at 8 GPUs.  One fence, then eight independent stores.
__device__ __forceinline__ void storeRelaxedSystem(unsigned* p, unsigned v) {
	asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned loadRelaxedSystem(const unsigned* p) {
	unsigned v;
	asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long globalTimer() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// one thread: tell every rank "epoch e of this rank is out"
__device__ void signalPeers(const Peers& peers, size_t flagOffset, unsigned epoch) {
	__threadfence_system();   // everything this rank stored (observed by this thread) is performed before any flag
	for (unsigned g = 0; g < peers.world; ++g)
		storeRelaxedSystem(reinterpret_cast<unsigned*>(peers.base[g] + flagOffset) + peers.rank, epoch);
}

// one thread: wait until every rank has signalled `epoch`.  A rank that never answers (it failed) must not hang the GPU:
// after 10 s the wait is abandoned and *error set; the host turns that into ErrorExternalLibrary at its next
// synchronisation point.
__device__ void waitPeers(const Peers& peers, size_t flagOffset, unsigned epoch, unsigned* error) {
	const unsigned* mine = reinterpret_cast<const unsigned*>(peers.base[peers.rank] + flagOffset);
	const unsigned long long t0 = globalTimer();
	for (unsigned g = 0; g < peers.world; ++g) {
		unsigned spins = 0;
		// epochs only grow; "reached" must survive a wrap of the 32-bit counter
		while ((int)(loadRelaxedSystem(mine + g) - epoch) < 0) {
			if ((++spins & 0x3FF) == 0 && globalTimer() - t0 > 10000000000ull) {
				atomicExch(error, 1u);
				return;
			}
		}
	}
	__threadfence_system();   // acquire: nothing after this is satisfied from before the flags were seen
}

__device__ __forceinline__ unsigned currentEpoch(const Control& ctl) { return *reinterpret_cast<volatile const unsigned*>(ctl.epoch); }

// Called by ALL threads of EVERY block of a 1-D grid after their (peer) stores: the block that finishes last tells every
// rank -- the signal costs no launch of its own and leaves as soon as the data is out.  newEpoch: the iteration counter
// advances with this signal (once per iteration, pushN).
__device__ void lastBlockSignals(const Peers& peers, size_t flagOffset, const Control& ctl, unsigned ticket, bool newEpoch) {
	__syncthreads();   // every store of the block happens-before thread 0's fence: one system fence per block, not per thread
	                   // (a membar.sys from each of 160 000 threads made pushN 34 us at 8 GPUs)
	if (threadIdx.x == 0) {
		__threadfence_system();   // cumulative: the block's stores are performed, at every rank, before the ticket is taken
		unsigned* t = ctl.tickets + ticket;
		if (atomicAdd(t, 1u) == gridDim.x - 1) {
			*t = 0;   // ready for the next launch (or graph replay)
			unsigned e = currentEpoch(ctl);
			if (newEpoch) {
				e += 1u;
				*reinterpret_cast<volatile unsigned*>(ctl.epoch) = e;
			}
			signalPeers(peers, flagOffset, e);
		}
	}
}

// out[i] (+)= sum over the ranks, in rank order, of four consecutive statistics: one 16-byte load per rank, all of them in
// flight at once (a scalar loop over 8 ranks x 4 160 entries per block was a chain of 33 L2 round trips, 25 us at 8 GPUs)
template <bool CACHED>
__device__ __forceinline__ float4 sumRanks4(const Peers& peers, const float* stat, unsigned statLen, unsigned index) {
	float4 v[kMaxRanks];
#pragma unroll
	for (unsigned g = 0; g < kMaxRanks; ++g)
		if (g < peers.world) {
			const float4* p = reinterpret_cast<const float4*>(stat + (size_t)g * statLen + index);
			v[g] = CACHED ? __ldg(p) : __ldcg(p);
		}
	float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
	for (unsigned g = 0; g < kMaxRanks; ++g)
		if (g < peers.world) {
			s.x += v[g].x; s.y += v[g].y; s.z += v[g].z; s.w += v[g].w;
		}
	return s;
}

__device__ __forceinline__ float tf32Hi(float x) {
	uint32_t u;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
	return __uint_as_float(u);
}

// ---- W statistics -> column scales (outside the iteration: store, diagnostics; inside, updateH does this per block) -------
// statSum = sum over the ranks of [W_un^T W_un, column sums of W_un, flag].  Unit columns (KernelNormalizeColumns.cu:52-58:
// divide by the norm where the sum of squares is positive) are applied lazily: inv[c] = 1 / ||w_c||, so
//   W^T W of the unit-column matrix  = statSum[i, j] * inv[i] * inv[j]
//   centring term of W^T V           = center * column sum * inv
// flag == 0 (initial factors, constant W): W is used as it is (the reference normalises only after a W update, MU.h:247).
__global__ void __launch_bounds__(1024) prep_h_kernel(Peers peers, size_t statW, unsigned statLen, unsigned k, float center,
                                                      float* __restrict__ statSum, float* __restrict__ G, float* __restrict__ inv,
                                                      float* __restrict__ corrN) {
	__shared__ float invS[128];
	const unsigned count = k * k + k + 1;
	const float* local = reinterpret_cast<const float*>(peers.base[peers.rank] + statW);
	for (unsigned idx = threadIdx.x; idx < count; idx += blockDim.x) {
		float s = 0.f;
		for (unsigned g = 0; g < peers.world; ++g) s += __ldcg(local + (size_t)g * statLen + idx);
		statSum[idx] = s;
	}
	__syncthreads();
	const bool normalise = statSum[k * k + k] > 0.5f;
	for (unsigned c = threadIdx.x; c < k; c += blockDim.x) {
		const float d = statSum[(size_t)c * k + c];
		const float v = (normalise && d > 0.f) ? 1.0f / sqrtf(d) : 1.0f;
		invS[c] = v;
		inv[c] = v;
		corrN[c] = center * (statSum[(size_t)k * k + c] * v);
	}
	__syncthreads();
	for (unsigned idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
		const unsigned r = idx % k, t = idx / k;
		G[idx] = statSum[idx] * invS[r] * invS[t];
	}
}

// ---- H statistics -> H H^T and the centring term, as a kernel of its own: residual iterations (the trace term needs
// H H^T before the W update) and a constant W (no W update that would do it per block) ---------------------------------------
__global__ void __launch_bounds__(1024) finish_h_kernel(Peers peers, size_t flagsH, size_t statH, unsigned statLen, Control ctl, unsigned k, float center,
                                                        float* __restrict__ B, float* __restrict__ corrP) {
	if (peers.world > 1 && threadIdx.x == 0) waitPeers(peers, flagsH, currentEpoch(ctl), ctl.error);
	__syncthreads();
	const float* local = reinterpret_cast<const float*>(peers.base[peers.rank] + statH);
	for (unsigned idx = threadIdx.x; idx < k * k + k; idx += blockDim.x) {
		float s = 0.f;
		for (unsigned g = 0; g < peers.world; ++g) s += __ldcg(local + (size_t)g * statLen + idx);   // rank order: identical on every rank
		if (idx < k * k) B[idx] = s;
		else corrP[idx - k * k] = center * s;
	}
}

// residual iterations over several ranks: the sum of the per-column terms of the own columns (fp64, fixed order) goes to
// every rank, so that each of them can form the residual without a host-side collective
__global__ void __launch_bounds__(1024) trace_sum_push_kernel(Peers peers, size_t oTrace, const float* __restrict__ tracePartials, unsigned count) {
	__shared__ double red[32];
	double a = 0.0;
	for (unsigned i = threadIdx.x; i < count; i += 1024) a += (double)tracePartials[i];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
	if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = a;
	__syncthreads();
	if (threadIdx.x == 0) {
		double total = 0.0;
		for (int w = 0; w < 32; ++w) total += red[w];
		for (unsigned g = 0; g < peers.world; ++g) reinterpret_cast<double*>(peers.base[g] + oTrace)[peers.rank] = total;
	}
}

__device__ void reduceAndPush(const Peers& peers, size_t dstOffset, unsigned statLen, const float* __restrict__ partials, unsigned blocks, unsigned count,
                              float flag, unsigned bx);

// ---- step 1b (several ranks) ---------------------------------------------------------------------------------------------
// The partial product of the own row block, summed over its stream-K slots, goes to the rank that owns the columns: one
// float4 per thread, a column (kp contiguous values) per group of threads, so the peer stores leave as whole 256-byte
// rows.  (Stored straight from the tensor-core kernel -- one 16-byte piece per thread at a 256-byte stride -- the same
// bytes cost 50 us over NVLink at 8 GPUs: W^T V 144 us against 93 us for the same launch without peers.)
__global__ void __launch_bounds__(256) push_n_kernel(Peers peers, size_t oSlots, size_t flagsN, Control ctl, unsigned kp, unsigned N, unsigned colsPerRank,
                                                     size_t ldh, const float* __restrict__ local, size_t localStride,
                                                     const unsigned char* __restrict__ slotCount, unsigned pushBlocks, size_t statW, unsigned statLen,
                                                     const float* __restrict__ statPartials, unsigned statBlocks, unsigned statCount,
                                                     const float* __restrict__ statFlag) {
	if (blockIdx.x >= pushBlocks) {
		// the CTAs behind the pushing ones: the statistics of this rank's rows of W (block partials of the last W update,
		// or of the initial factors) summed and stored to every rank -- they travel with the same signal
		reduceAndPush(peers, statW, statLen, statPartials, statBlocks, statCount, *statFlag, blockIdx.x - pushBlocks);
		lastBlockSignals(peers, flagsN, ctl, 0, true);
		return;
	}
	const unsigned perCol = kp / 4;
	const unsigned long long idx = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
	const unsigned j = (unsigned)(idx / perCol), q = (unsigned)(idx % perCol);
	if (j < N) {
		const unsigned splits = slotCount[j >> 7];
		const float* src = local + (size_t)j * ldh + 4 * q;
		float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
		unsigned sl = 0;
		for (; sl + 4 <= splits; sl += 4) {   // four loads in flight: the sum of 5..15 slots is a chain of L2 latencies otherwise
			const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sl * localStride));
			const float4 y = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sl + 1) * localStride));
			const float4 z = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sl + 2) * localStride));
			const float4 w = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sl + 3) * localStride));
			a.x += x.x + z.x; a.y += x.y + z.y; a.z += x.z + z.z; a.w += x.w + z.w;
			b.x += y.x + w.x; b.y += y.y + w.y; b.z += y.z + w.z; b.w += y.w + w.w;
		}
		for (; sl < splits; ++sl) {
			const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sl * localStride));
			a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
		}
		const unsigned owner = j / colsPerRank;
		float* dst = reinterpret_cast<float*>(peers.base[owner] + oSlots) + ((size_t)peers.rank * colsPerRank + (j - owner * colsPerRank)) * ldh + 4 * q;
		*reinterpret_cast<float4*>(dst) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
	}
	// "my partials -- and, earlier in this stream, the statistics of my rows of W -- are out": a new iteration's first signal
	lastBlockSignals(peers, flagsN, ctl, 0, true);
}

// ---- step 2 ------------------------------------------------------------------------------------------------------------
// A COLS-column panel of the own columns per block of 4 COLS threads (64 columns; 32 or 16 when the rank owns so few
// columns that 64-wide panels would leave most SMs idle).  D = G H is a register-tiled product out of shared memory
// (thread = KP/16 rows x 4 columns); the numerators are the partial products of all ranks and slots, fetched several
// partials at a time so that their latencies overlap; the multiplicative update (KernelMultiplyDivide.cu:42: multiply,
// then divide), the residual term and the stores to every rank are the epilogue; then the Gram matrix of the new panel.
// slotCount == nullptr: one partial per rank (several ranks: fused::pushN has summed the slots).
template <int KP, int COLS>
__global__ void __launch_bounds__(COLS * 4) update_h_fused(Peers peers, size_t oH, size_t oHtHi, size_t oHtLo, size_t oSlots, size_t flagsN, size_t statW,
                                                          unsigned statLen, Control ctl, float center, unsigned k, unsigned c0, unsigned nOwn, size_t ldh,
                                                          size_t ldht, unsigned slotsPerRank, size_t slotStride, const unsigned char* __restrict__ slotCount,
                                                          float* __restrict__ Gout, float* __restrict__ invOut, float* __restrict__ corrNout, float eps,
                                                          float* __restrict__ tracePartials, float* __restrict__ statPart) {
	constexpr int NT = COLS * 4, RPT = KP / 16, LDJ = COLS + 4;
	const unsigned jl0 = blockIdx.x * COLS;   // first column of the panel: local index, global index
	const unsigned j0 = c0 + jl0;
	const unsigned splits = (slotCount != nullptr && jl0 < nOwn) ? slotCount[j0 >> 7] : 1u;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ float invS[128], corrS[128];
	__shared__ float flagS;
	float* Gs = reinterpret_cast<float*>(smem_raw);  // [KP t][KP r]: Gs[t*KP + r] = G[r + t*k]
	float* Hs = Gs + KP * KP;                         // [KP t][LDJ]:  Hs[t*LDJ + j] = H[t, j0 + j]
	const unsigned tid = threadIdx.x;
	const float* Hloc = reinterpret_cast<const float*>(peers.base[peers.rank] + oH);
	const float* Nloc = reinterpret_cast<const float*>(peers.base[peers.rank] + oSlots);
	// every rank's partials of W^T V and the statistics of its rows of W have landed here
	if (peers.world > 1) {
		if (tid == 0) waitPeers(peers, flagsN, currentEpoch(ctl), ctl.error);
		__syncthreads();
	}
	// Statistics of the UN-NORMALISED W, summed over the ranks in rank order (identical on every rank and in every block):
	// Gram matrix, column sums, flag.  Unit columns (KernelNormalizeColumns.cu:52-58: divide by the norm where the sum of
	// squares is positive) are applied lazily: inv[c] = 1 / ||w_c||, W^T W of the unit-column matrix = G[i, j] inv[i] inv[j],
	// centring term of W^T V = center * column sum * inv.  flag == 0 (initial factors, constant W): W is used as it is
	// (the reference normalises only after a W update, MU.h:247).
	{
		const float* stat = reinterpret_cast<const float*>(peers.base[peers.rank] + statW);
		if (k % 4 == 0) {
			for (unsigned idx = tid; idx < KP * KP / 4; idx += NT) {
				const unsigned r = (idx % (KP / 4)) * 4, t = idx / (KP / 4);
				float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
				if (r < k && t < k) s = sumRanks4<false>(peers, stat, statLen, t * k + r);
				*reinterpret_cast<float4*>(Gs + t * KP + r) = s;
			}
		} else {
			for (unsigned idx = tid; idx < KP * KP; idx += NT) {
				const unsigned r = idx % KP, t = idx / KP;
				float s = 0.f;
				if (r < k && t < k)
					for (unsigned g = 0; g < peers.world; ++g) s += __ldcg(stat + (size_t)g * statLen + (size_t)t * k + r);
				Gs[idx] = s;
			}
		}
		for (unsigned c = tid; c < k; c += NT) {
			float s = 0.f;
			for (unsigned g = 0; g < peers.world; ++g) s += __ldcg(stat + (size_t)g * statLen + (size_t)k * k + c);
			corrS[c] = s;
		}
		if (tid == 0) {
			float f = 0.f;
			for (unsigned g = 0; g < peers.world; ++g) f += __ldcg(stat + (size_t)g * statLen + (size_t)k * k + k);
			flagS = f;
		}
	}
	for (unsigned idx = tid; idx < COLS * KP; idx += NT) {
		const unsigned t = idx % KP, j = idx / KP;
		Hs[t * LDJ + j] = (jl0 + j < nOwn && t < k) ? Hloc[(size_t)(j0 + j) * ldh + t] : 0.f;
	}
	__syncthreads();
	for (unsigned c = tid; c < KP; c += NT) {
		float v = 1.f, cs = 0.f;
		if (c < k) {
			const float d = Gs[c * KP + c];
			v = (flagS > 0.5f && d > 0.f) ? 1.0f / sqrtf(d) : 1.0f;
			cs = center * (corrS[c] * v);
		}
		invS[c] = v;
		corrS[c] = cs;
	}
	__syncthreads();
	for (unsigned idx = tid; idx < KP * KP; idx += NT) {
		const unsigned r = idx % KP, t = idx / KP;
		Gs[idx] = Gs[idx] * invS[r] * invS[t];
	}
	if (blockIdx.x == 0) {   // for the trace term, the W update and the store of the factors
		for (unsigned c = tid; c < k; c += NT) {
			invOut[c] = invS[c];
			corrNout[c] = corrS[c];
		}
	}
	__syncthreads();
	if (blockIdx.x == 0)
		for (unsigned idx = tid; idx < k * k; idx += NT) Gout[idx] = Gs[(idx / k) * KP + idx % k];
	const unsigned rx = tid % 16, jx = tid / 16;      // rows rx*RPT.., columns jx*4..
	float numv[RPT][4];
#pragma unroll
	for (int q = 0; q < 4; ++q)
#pragma unroll
		for (int i = 0; i < RPT; ++i) numv[i][q] = 0.f;
	{
		const unsigned total = peers.world * splits;
		const bool vec = RPT % 4 == 0 && rx * RPT + RPT <= k;
		auto slotOf = [&](unsigned p) {
			const unsigned g = p / splits, sl = p - g * splits;
			return Nloc + ((size_t)g * slotsPerRank + sl) * slotStride;
		};
		auto fetch = [&](const float* slot, float (&x)[RPT][4]) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const unsigned jl = jl0 + jx * 4 + q;
				if (jl < nOwn) {
					const float* src = slot + (size_t)jl * ldh + rx * RPT;
					if (vec) {
#pragma unroll
						for (int i = 0; i < RPT; i += 4) {
							const float4 v = __ldcg(reinterpret_cast<const float4*>(src + i));
							x[i][q] = v.x; x[i + 1][q] = v.y; x[i + 2][q] = v.z; x[i + 3][q] = v.w;
						}
					} else {
#pragma unroll
						for (int i = 0; i < RPT; ++i) x[i][q] = (rx * RPT + i < k) ? __ldcg(src + i) : 0.f;
					}
				} else {
#pragma unroll
					for (int i = 0; i < RPT; ++i) x[i][q] = 0.f;
				}
			}
		};
		constexpr int BATCH = RPT <= 4 ? 4 : 2;   // partials in flight per thread
		unsigned p = 0;
		for (; p + BATCH <= total; p += BATCH) {
			float x[BATCH][RPT][4];
#pragma unroll
			for (int b = 0; b < BATCH; ++b) fetch(slotOf(p + b), x[b]);
#pragma unroll
			for (int b = 0; b < BATCH; ++b)
#pragma unroll
				for (int q = 0; q < 4; ++q)
#pragma unroll
					for (int i = 0; i < RPT; ++i) numv[i][q] += x[b][i][q];
		}
		for (; p < total; ++p) {
			float x0[RPT][4];
			fetch(slotOf(p), x0);
#pragma unroll
			for (int q = 0; q < 4; ++q)
#pragma unroll
				for (int i = 0; i < RPT; ++i) numv[i][q] += x0[i][q];
		}
	}
	// N = diag(inv) (sum of the partials) + centring term: what W^T V of the unit-column matrix would have been
#pragma unroll
	for (int i = 0; i < RPT; ++i) {
		const unsigned r = rx * RPT + i;
		const float s = r < k ? invS[r] : 0.f, c = r < k ? corrS[r] : 0.f;
#pragma unroll
		for (int q = 0; q < 4; ++q) numv[i][q] = (jl0 + jx * 4 + q < nOwn) ? fmaf(s, numv[i][q], c) : 0.f;
	}
	float acc[RPT][4];
#pragma unroll
	for (int i = 0; i < RPT; ++i)
#pragma unroll
		for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
#pragma unroll 8
	for (int t = 0; t < KP; ++t) {
		float a[RPT];
		if (RPT % 4 == 0) {
#pragma unroll
			for (int i = 0; i < RPT; i += 4) {
				const float4 g = *reinterpret_cast<const float4*>(Gs + t * KP + rx * RPT + i);
				a[i] = g.x; a[i + 1] = g.y; a[i + 2] = g.z; a[i + 3] = g.w;
			}
		} else {
#pragma unroll
			for (int i = 0; i < RPT; ++i) a[i] = Gs[t * KP + rx * RPT + i];
		}
		const float4 b = *reinterpret_cast<const float4*>(Hs + t * LDJ + jx * 4);
#pragma unroll
		for (int i = 0; i < RPT; ++i) {
			acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
			acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
			acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
			acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
		}
	}
	// epilogue: the new values replace the old ones in the H tile (every thread overwrites exactly the entries it read)
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const unsigned jl = jl0 + jx * 4 + q;
		float tr = 0.f;
#pragma unroll
		for (int i = 0; i < RPT; ++i) {
			const unsigned r = rx * RPT + i;
			const float num = numv[i][q];
			const float v = Hs[r * LDJ + jx * 4 + q] * num / (acc[i][q] + eps);   // padding entries: 0 * 0 / eps = 0
			Hs[r * LDJ + jx * 4 + q] = v;
			tr = fmaf(v, num, tr);
		}
		tr += __shfl_xor_sync(0xffffffffu, tr, 1);
		tr += __shfl_xor_sync(0xffffffffu, tr, 2);
		tr += __shfl_xor_sync(0xffffffffu, tr, 4);
		tr += __shfl_xor_sync(0xffffffffu, tr, 8);
		if (tracePartials != nullptr && rx == 0 && jl < nOwn) tracePartials[jl] = tr;   // MU.h:194-197
	}
	__syncthreads();
	// the new columns and their transposed TF32 split go to every rank: the all-gather of H is this kernel's epilogue
	{   // H itself is needed by its owner only (the next update of these columns); the other ranks read H^T hi/lo.  The
		// store of the factors gathers H (engine.cu storeFused)
		float* Hout = reinterpret_cast<float*>(peers.base[peers.rank] + oH);
		for (unsigned idx = tid; idx < COLS * KP; idx += NT) {
			const unsigned t = idx % KP, j = idx / KP;
			if (jl0 + j < nOwn && t < k) Hout[(size_t)(j0 + j) * ldh + t] = Hs[t * LDJ + j];
		}
	}
	for (unsigned g = 0; g < peers.world; ++g) {
		float* HtHi = reinterpret_cast<float*>(peers.base[g] + oHtHi);
		float* HtLo = reinterpret_cast<float*>(peers.base[g] + oHtLo);
		for (unsigned idx = tid; idx < COLS * KP; idx += NT) {
			const unsigned j = idx % COLS, r = idx / COLS;
			if (r < k && jl0 + j < nOwn) {
				const float v = Hs[r * LDJ + j];
				const float hi = tf32Hi(v);
				HtHi[(size_t)r * ldht + j0 + j] = hi;
				HtLo[(size_t)r * ldht + j0 + j] = v - hi;
			}
		}
	}
	// statistics of the new panel: Gram matrix (thread = rows a + 16 i x rows b + BG q: conflict-free float4 reads) and row sums
	float* stat = statPart + (size_t)blockIdx.x * ((size_t)k * k + k);
	{
		constexpr int BG = NT / 16, QN = KP / BG, QB = QN > 8 ? 8 : QN;
		const unsigned a = tid % 16, b = tid / 16;
		for (int q0 = 0; q0 < QN; q0 += QB) {
			float gr[RPT][QB];
#pragma unroll
			for (int i = 0; i < RPT; ++i)
#pragma unroll
				for (int q = 0; q < QB; ++q) gr[i][q] = 0.f;
			for (int j = 0; j < COLS; j += 4) {
				float4 av[RPT];
#pragma unroll
				for (int i = 0; i < RPT; ++i) av[i] = *reinterpret_cast<const float4*>(Hs + (a + 16 * i) * LDJ + j);
#pragma unroll
				for (int q = 0; q < QB; ++q) {
					const float4 bv = *reinterpret_cast<const float4*>(Hs + (b + BG * (q0 + q)) * LDJ + j);
#pragma unroll
					for (int i = 0; i < RPT; ++i) gr[i][q] = fmaf(av[i].w, bv.w, fmaf(av[i].z, bv.z, fmaf(av[i].y, bv.y, fmaf(av[i].x, bv.x, gr[i][q]))));
				}
			}
#pragma unroll
			for (int q = 0; q < QB; ++q)
#pragma unroll
				for (int i = 0; i < RPT; ++i) {
					const unsigned r1 = a + 16 * i, r2 = b + BG * (q0 + q);
					if (r1 < k && r2 < k) stat[(size_t)r2 * k + r1] = gr[i][q];
				}
		}
	}
	{
		const unsigned lane = tid % 32;
		for (unsigned r = tid / 32; r < k; r += NT / 32) {
			float sum = 0.f;
			for (unsigned j = lane; j < COLS; j += 32) sum += Hs[r * LDJ + j];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
			if (lane == 0) stat[(size_t)k * k + r] = sum;
		}
	}
}

// ---- steps 3 and 6 -------------------------------------------------------------------------------------------------------
// 32 entries x 8 block groups per CTA (bx = index of the CTA among those that reduce); every thread keeps four loads in
// flight; fixed summation order.  All threads of the CTA must call it.
__device__ void reduceAndPush(const Peers& peers, size_t dstOffset, unsigned statLen, const float* __restrict__ partials, unsigned blocks, unsigned count,
                              float flag, unsigned bx) {
	__shared__ float red[8][33];
	const unsigned lane = threadIdx.x % 32, grp = threadIdx.x / 32;
	const unsigned x = bx * 32 + lane;
	float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
	if (x < count) {
		unsigned b = grp;
		for (; b + 24 < blocks; b += 32) {
			a0 += partials[(size_t)b * count + x];
			a1 += partials[(size_t)(b + 8) * count + x];
			a2 += partials[(size_t)(b + 16) * count + x];
			a3 += partials[(size_t)(b + 24) * count + x];
		}
		for (; b < blocks; b += 8) a0 += partials[(size_t)b * count + x];
	}
	red[grp][lane] = (a0 + a1) + (a2 + a3);
	__syncthreads();
	if (grp == 0 && x < count) {
		const float v = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) + ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
		for (unsigned g = 0; g < peers.world; ++g) reinterpret_cast<float*>(peers.base[g] + dstOffset)[(size_t)peers.rank * statLen + x] = v;
	}
	if (flag >= 0.f && bx == 0 && threadIdx.x == 0)
		for (unsigned g = 0; g < peers.world; ++g) reinterpret_cast<float*>(peers.base[g] + dstOffset)[(size_t)peers.rank * statLen + count] = flag;
}

__global__ void __launch_bounds__(256) reduce_push_kernel(Peers peers, size_t dstOffset, unsigned statLen, const float* __restrict__ partials,
                                                          unsigned blocks, unsigned count, float flag, size_t signalFlags, Control ctl, unsigned ticket) {
	reduceAndPush(peers, dstOffset, statLen, partials, blocks, count, flag, blockIdx.x);
	// H side: "my columns of H (the update kernel before this one) and their statistics are out"
	if (signalFlags != kNoSignal && peers.world > 1) lastBlockSignals(peers, signalFlags, ctl, ticket, false);
}

// ---- step 5 ------------------------------------------------------------------------------------------------------------
// Every block walks 128-row panels blockIdx.x, blockIdx.x + gridDim.x, ... (the grid is sized to one wave, so 782 panels of
// the 100 000-row problem are 391 blocks of two).  H H^T comes into shared memory once per block.  Per panel: the tile is
// read with the column scale applied (unit columns of the previous update, KernelNormalizeColumns.cu:52-58, without a pass
// of their own), D = W (H H^T) is a register-tiled product out of shared memory (thread = 4 rows x KP/8 columns), the
// multiplicative update writes the new un-normalised rows and their TF32 split (what the next W^T V reads) and leaves them
// in the tile; then the Gram matrix and the column sums of the new rows are ACCUMULATED in registers over the block's
// panels -- the statistics the next updateH turns into norms, W^T W and the centring term; one partial per block.
template <int KP, bool UPDATE>
__global__ void __launch_bounds__(256, KP <= 64 ? 3 : 1) update_w_fused(Peers peers, size_t statH, unsigned statLen, float center, unsigned m, unsigned k,
                                                     float* __restrict__ Bout, float* __restrict__ corrPout, const float* __restrict__ inv,
                                                     float* __restrict__ W, size_t ldw, float* __restrict__ Whi, float* __restrict__ Wlo,
                                                     const float* __restrict__ Ppart, size_t ldp, size_t slotStride,
                                                     const unsigned char* __restrict__ slotCount, float eps, float* __restrict__ statPart,
                                                     float* __restrict__ statFlag) {
	constexpr int ROWS = 128, CPT = KP / 8, LDW = ROWS + 4, RPT = KP / 16, CSUM = (KP + 7) / 8;
	// 1: these statistics describe an updated W whose columns get normalised; 0: the initial factors, used as they are
	if (blockIdx.x == 0 && threadIdx.x == 0 && statFlag != nullptr) *statFlag = UPDATE ? 1.f : 0.f;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ float corrS[128], invS[128];
	float* Ws = reinterpret_cast<float*>(smem_raw);  // [KP t][LDW]: Ws[t*LDW + r] = W[i0 + r, t] (scaled)
	float* Bs = Ws + KP * LDW;                        // [KP t][KP c]: Bs[t*KP + c] = B[t + c*k]
	const unsigned tid = threadIdx.x, lane = tid % 32;
	const unsigned panels = (m + ROWS - 1) / ROWS;
	for (unsigned c = tid; c < KP; c += 256) invS[c] = (UPDATE && c < k) ? inv[c] : 1.f;
	if (UPDATE) {
		// H H^T and the centring term of V H^T from the statistics of every rank's columns of H, summed in rank order.  The
		// gate in the tensor-core kernel BEFORE this one has waited for them, so they were complete when this kernel
		// started: ordinary cached loads
		const float* stat = reinterpret_cast<const float*>(peers.base[peers.rank] + statH);
		if (k % 4 == 0) {   // (H H^T is symmetric bit for bit: both triangles are the same fma chains with the factors swapped)
			for (unsigned idx = tid; idx < KP * KP / 4; idx += 256) {
				const unsigned c = (idx % (KP / 4)) * 4, t = idx / (KP / 4);
				float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
				if (c < k && t < k) sum = sumRanks4<true>(peers, stat, statLen, t * k + c);
				*reinterpret_cast<float4*>(Bs + t * KP + c) = sum;
			}
		} else {
			for (unsigned idx = tid; idx < KP * KP; idx += 256) {
				const unsigned c = idx % KP, t = idx / KP;
				float sum = 0.f;
				if (c < k && t < k)
					for (unsigned g = 0; g < peers.world; ++g) sum += __ldg(stat + (size_t)g * statLen + (size_t)c * k + t);
				Bs[idx] = sum;
			}
		}
		for (unsigned c = tid; c < KP; c += 256) {
			float sum = 0.f;
			if (c < k)
				for (unsigned g = 0; g < peers.world; ++g) sum += __ldg(stat + (size_t)g * statLen + (size_t)k * k + c);
			corrS[c] = center * sum;
		}
	}
	__syncthreads();
	if (UPDATE && blockIdx.x == 0) {   // for the diagnostics (debugProducts) and the residual iterations
		for (unsigned idx = tid; idx < k * k; idx += 256) Bout[idx] = Bs[(idx % k) * KP + idx / k];
		for (unsigned c = tid; c < k; c += 256) corrPout[c] = corrS[c];
	}
	// statistics accumulated over the panels of this block: Gram entries (columns a + 16 i) x (columns b + 16 q), and per
	// lane partial column sums of the columns warp + 8 j
	const unsigned ga = tid % 16, gb = tid / 16;
	float gr[RPT][RPT], csum[CSUM];
#pragma unroll
	for (int i = 0; i < RPT; ++i)
#pragma unroll
		for (int q = 0; q < RPT; ++q) gr[i][q] = 0.f;
#pragma unroll
	for (int j = 0; j < CSUM; ++j) csum[j] = 0.f;

	for (unsigned panel = blockIdx.x; panel < panels; panel += gridDim.x) {
		const unsigned i0 = panel * ROWS;
		__syncthreads();   // the previous panel's statistics are done with the tile
		for (unsigned idx = tid; idx < KP * ROWS; idx += 256) {
			const unsigned r = idx % ROWS, t = idx / ROWS;
			float v = 0.f;
			if (i0 + r < m && t < k) v = W[(size_t)t * ldw + i0 + r] * invS[t];
			Ws[t * LDW + r] = v;
		}
		__syncthreads();
		if (UPDATE) {
			const unsigned splits = slotCount != nullptr ? slotCount[panel] : 1u;
			const unsigned tx = tid % 32, ty = tid / 32;      // rows tx*4.., columns ty*CPT..
			float acc[4][CPT];
#pragma unroll
			for (int i = 0; i < 4; ++i)
#pragma unroll
				for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
#pragma unroll 4
			for (int t = 0; t < KP; ++t) {
				const float4 a = *reinterpret_cast<const float4*>(Ws + t * LDW + tx * 4);
				float b[CPT];
				if (CPT % 4 == 0) {
#pragma unroll
					for (int j = 0; j < CPT; j += 4) {
						const float4 x = *reinterpret_cast<const float4*>(Bs + t * KP + ty * CPT + j);
						b[j] = x.x; b[j + 1] = x.y; b[j + 2] = x.z; b[j + 3] = x.w;
					}
				} else {
#pragma unroll
					for (int j = 0; j < CPT; ++j) b[j] = Bs[t * KP + ty * CPT + j];
				}
#pragma unroll
				for (int j = 0; j < CPT; ++j) {
					acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
					acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
					acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
					acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
				}
			}
			__syncthreads();   // every thread is done reading the old tile: it can now be overwritten with the new values
			const unsigned r0 = i0 + tx * 4;
#pragma unroll
			for (int j = 0; j < CPT; ++j) {
				const unsigned c = ty * CPT + j;
				if (c < k) {   // warp-uniform
					const float base = corrS[c];
					float p[4] = {base, base, base, base};
					if (r0 + 3 < m) {
						for (unsigned sl = 0; sl < splits; ++sl) {
							const float4 x = __ldcg(reinterpret_cast<const float4*>(Ppart + sl * slotStride + (size_t)c * ldp + r0));
							p[0] += x.x; p[1] += x.y; p[2] += x.z; p[3] += x.w;
						}
					} else {
						for (unsigned sl = 0; sl < splits; ++sl)
							for (int i = 0; i < 4; ++i)
								if (r0 + i < m) p[i] += __ldcg(Ppart + sl * slotStride + (size_t)c * ldp + r0 + i);
					}
					const float4 w = *reinterpret_cast<const float4*>(Ws + c * LDW + tx * 4);
					float wn[4];
					wn[0] = w.x * p[0] / (acc[0][j] + eps);   // KernelMultiplyDivide.cu:42: multiply first, then divide
					wn[1] = w.y * p[1] / (acc[1][j] + eps);
					wn[2] = w.z * p[2] / (acc[2][j] + eps);
					wn[3] = w.w * p[3] / (acc[3][j] + eps);
					float hi[4];
					if (r0 + 3 < m) {
#pragma unroll
						for (int i = 0; i < 4; ++i) hi[i] = tf32Hi(wn[i]);
						*reinterpret_cast<float4*>(W + (size_t)c * ldw + r0) = make_float4(wn[0], wn[1], wn[2], wn[3]);
						*reinterpret_cast<float4*>(Whi + (size_t)c * ldw + r0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
						*reinterpret_cast<float4*>(Wlo + (size_t)c * ldw + r0) = make_float4(wn[0] - hi[0], wn[1] - hi[1], wn[2] - hi[2], wn[3] - hi[3]);
					} else {
						for (int i = 0; i < 4; ++i) {
							if (r0 + i < m) {
								const float h = tf32Hi(wn[i]);
								W[(size_t)c * ldw + r0 + i] = wn[i];
								Whi[(size_t)c * ldw + r0 + i] = h;
								Wlo[(size_t)c * ldw + r0 + i] = wn[i] - h;
							} else {
								wn[i] = 0.f;
							}
						}
					}
					*reinterpret_cast<float4*>(Ws + c * LDW + tx * 4) = make_float4(wn[0], wn[1], wn[2], wn[3]);
				}
			}
			__syncthreads();
		}
		// statistics of the rows now in the tile
#pragma unroll 2
		for (int r = 0; r < ROWS; r += 4) {
			float4 av[RPT], bv[RPT];
#pragma unroll
			for (int i = 0; i < RPT; ++i) av[i] = *reinterpret_cast<const float4*>(Ws + (ga + 16 * i) * LDW + r);
#pragma unroll
			for (int q = 0; q < RPT; ++q) bv[q] = *reinterpret_cast<const float4*>(Ws + (gb + 16 * q) * LDW + r);
#pragma unroll
			for (int i = 0; i < RPT; ++i)
#pragma unroll
				for (int q = 0; q < RPT; ++q)
					gr[i][q] = fmaf(av[i].w, bv[q].w, fmaf(av[i].z, bv[q].z, fmaf(av[i].y, bv[q].y, fmaf(av[i].x, bv[q].x, gr[i][q]))));
		}
#pragma unroll
		for (int j = 0; j < CSUM; ++j) {
			const unsigned c = tid / 32 + 8 * j;
			if (c < KP) {
				const float* col = Ws + c * LDW;
				csum[j] += (col[lane] + col[lane + 32]) + (col[lane + 64] + col[lane + 96]);
			}
		}
	}
	float* stat = statPart + (size_t)blockIdx.x * ((size_t)k * k + k);
#pragma unroll
	for (int q = 0; q < RPT; ++q)
#pragma unroll
		for (int i = 0; i < RPT; ++i) {
			const unsigned c1 = ga + 16 * i, c2 = gb + 16 * q;
			if (c1 < k && c2 < k) stat[(size_t)c2 * k + c1] = gr[i][q];
		}
#pragma unroll
	for (int j = 0; j < CSUM; ++j) {
		const unsigned c = tid / 32 + 8 * j;
		float sum = csum[j];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
		if (lane == 0 && c < k) stat[(size_t)k * k + c] = sum;
	}
}

__global__ void scale_rows_kernel(unsigned rows, unsigned rowsPadded, unsigned k, const float* __restrict__ W, size_t ldw, const float* __restrict__ inv,
                                  float* __restrict__ block) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (r >= rowsPadded) return;
	block[(size_t)c * rowsPadded + r] = r < rows ? W[(size_t)c * ldw + r] * inv[c] : 0.f;
}

__global__ void unpack_rows_kernel(unsigned m, unsigned k, unsigned rowsPadded, const float* __restrict__ gathered, float* __restrict__ W, size_t ldw) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (i >= m) return;
	const unsigned g = i / rowsPadded, r = i - g * rowsPadded;
	W[(size_t)c * ldw + i] = gathered[((size_t)g * k + c) * rowsPadded + r];
}

// out[r + j * ldo] = inv[r] * (sum of the partials of W^T V of all ranks and slots) + corrN[r] for the own columns: the
// numerator of the H update as a matrix (tests, bench)
__global__ void collect_n_kernel(Peers peers, size_t oSlots, unsigned k, unsigned c0, unsigned nOwn, size_t ldh, unsigned slotsPerRank, size_t slotStride,
                                 const unsigned char* __restrict__ slotCount, const float* __restrict__ inv, const float* __restrict__ corrN,
                                 float* __restrict__ out, size_t ldo) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= k) return;
	const float* Nloc = reinterpret_cast<const float*>(peers.base[peers.rank] + oSlots);
	for (unsigned jl = blockIdx.y; jl < nOwn; jl += gridDim.y) {
		const unsigned splits = slotCount != nullptr ? slotCount[(c0 + jl) >> 7] : 1u;
		float s = 0.f;
		for (unsigned g = 0; g < peers.world; ++g)
			for (unsigned sl = 0; sl < splits; ++sl) s += __ldcg(Nloc + ((size_t)g * slotsPerRank + sl) * slotStride + (size_t)jl * ldh + r);
		out[(size_t)jl * ldo + r] = fmaf(inv[r], s, corrN[r]);
	}
}

template <int KP, int COLS>
constexpr size_t smemUpdateH() {
	return sizeof(float) * ((size_t)KP * KP + (size_t)KP * (COLS + 4));
}
template <int KP>
constexpr size_t smemUpdateW() {
	return sizeof(float) * ((size_t)KP * (128 + 4) + (size_t)KP * KP);
}

template <int KP>
void configureRank() {
	CUDA_CHECK(cudaFuncSetAttribute(update_h_fused<KP, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemUpdateH<KP, 64>()));
	CUDA_CHECK(cudaFuncSetAttribute(update_h_fused<KP, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemUpdateH<KP, 32>()));
	CUDA_CHECK(cudaFuncSetAttribute(update_h_fused<KP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemUpdateH<KP, 16>()));
	CUDA_CHECK(cudaFuncSetAttribute(update_w_fused<KP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemUpdateW<KP>()));
	CUDA_CHECK(cudaFuncSetAttribute(update_w_fused<KP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemUpdateW<KP>()));
}

#define launchCheck() CUDA_CHECK(cudaGetLastError())   // a macro: the message names the launch site

}  // namespace

void configure() {
	configureRank<16>();
	configureRank<32>();
	configureRank<64>();
	configureRank<128>();
}

void prepH(const Peers& peers, const Layout& lay, unsigned k, float center, float* statSum, float* G, float* inv, float* corrN, cudaStream_t stream) {
	prep_h_kernel<<<1, 1024, 0, stream>>>(peers, lay.statW, lay.statLen, k, center, statSum, G, inv, corrN);
	launchCheck();
}

template <int KP, int COLS>
static unsigned launchUpdateHCols(const Peers& peers, const Layout& lay, const Control& ctl, float center, unsigned k, unsigned c0, unsigned nOwn,
                                  unsigned colsPerRank, size_t ldh, size_t ldht, unsigned slotsPerRank, const unsigned char* slotCount, float* G, float* inv,
                                  float* corrN, float eps, float* tracePartials, float* statPart, cudaStream_t stream) {
	const unsigned blocks = std::max(1u, ceilDiv(nOwn, COLS));   // a rank without columns still derives G, inv and corrN (block 0)
	update_h_fused<KP, COLS><<<blocks, COLS * 4, smemUpdateH<KP, COLS>(), stream>>>(peers, lay.H, lay.HtHi, lay.HtLo, lay.slots, lay.flagsN, lay.statW,
	                                                                                 lay.statLen, ctl, center, k, c0, nOwn, ldh, ldht, slotsPerRank,
	                                                                                 ldh * (size_t)colsPerRank, slotCount, G, inv, corrN, eps, tracePartials,
	                                                                                 statPart);
	launchCheck();
	return blocks;
}

// panel width: 64 columns unless that leaves more than half of the SMs without a block (a rank of an 8-GPU run owns
// 1 280 of 10 000 columns: 20 panels of 64 took 56 us, latency-bound)
unsigned panelColumnsH(unsigned nOwn) {
	static const unsigned forced = [] {
		const char* e = getenv("NMFGPU_UPDATE_H_COLS");   // study knob
		return e != nullptr ? (unsigned)atoi(e) : 0u;
	}();
	if (forced == 16 || forced == 32 || forced == 64) return forced;
	return ceilDiv(nOwn, 64) >= 74 ? 64 : 32;   // measured on 1 280 own columns: 64 -> 29 us, 32 -> 24.6 us, 16 -> 30 us
}

template <int KP>
static unsigned launchUpdateH(const Peers& peers, const Layout& lay, const Control& ctl, float center, unsigned k, unsigned c0, unsigned nOwn,
                              unsigned colsPerRank, size_t ldh, size_t ldht, unsigned slotsPerRank, const unsigned char* slotCount, float* G, float* inv,
                              float* corrN, float eps, float* tracePartials, float* statPart, cudaStream_t stream) {
#define NMF_ARGS peers, lay, ctl, center, k, c0, nOwn, colsPerRank, ldh, ldht, slotsPerRank, slotCount, G, inv, corrN, eps, tracePartials, statPart, stream
	switch (panelColumnsH(nOwn)) {
	case 64: return launchUpdateHCols<KP, 64>(NMF_ARGS);
	case 32: return launchUpdateHCols<KP, 32>(NMF_ARGS);
	default: return launchUpdateHCols<KP, 16>(NMF_ARGS);
	}
#undef NMF_ARGS
}

unsigned updateH(const Peers& peers, const Layout& lay, const Control& ctl, float center, unsigned k, unsigned c0, unsigned nOwn, unsigned colsPerRank,
                 size_t ldh, size_t ldht, unsigned slotsPerRank, const unsigned char* slotCount, float* G, float* inv, float* corrN, float eps,
                 float* tracePartials, float* statPart, cudaStream_t stream) {
#define NMF_ARGS peers, lay, ctl, center, k, c0, nOwn, colsPerRank, ldh, ldht, slotsPerRank, slotCount, G, inv, corrN, eps, tracePartials, statPart, stream
	if (k <= 16) return launchUpdateH<16>(NMF_ARGS);
	if (k <= 32) return launchUpdateH<32>(NMF_ARGS);
	if (k <= 64) return launchUpdateH<64>(NMF_ARGS);
	if (k <= 128) return launchUpdateH<128>(NMF_ARGS);
#undef NMF_ARGS
	throw EngineError(ResultType::ErrorInvalidArgument, "the fused MU kernels cover ranks up to 128");
}

void pushN(const Peers& peers, const Layout& lay, const Control& ctl, unsigned kp, unsigned N, unsigned colsPerRank, size_t ldh, const float* localSlots,
           size_t localStride, const unsigned char* slotCount, const float* statPartials, unsigned statBlocks, unsigned statCount, const float* statFlag,
           cudaStream_t stream) {
	const unsigned long long threads = (unsigned long long)N * (kp / 4);
	const unsigned pushBlocks = (unsigned)std::max<unsigned long long>(1, (threads + 255) / 256);
	push_n_kernel<<<pushBlocks + ceilDiv(statCount, 32), 256, 0, stream>>>(peers, lay.slots, lay.flagsN, ctl, kp, N, colsPerRank, ldh, localSlots, localStride,
	                                                                        slotCount, pushBlocks, lay.statW, lay.statLen, statPartials, statBlocks, statCount,
	                                                                        statFlag);
	launchCheck();
}

void reducePush(const Peers& peers, size_t dstOffset, unsigned statLen, const float* partials, unsigned blocks, unsigned count, float flag,
                size_t signalFlags, const Control& ctl, unsigned ticket, cudaStream_t stream) {
	reduce_push_kernel<<<ceilDiv(count, 32), 256, 0, stream>>>(peers, dstOffset, statLen, partials, blocks, count, flag, signalFlags, ctl, ticket);
	launchCheck();
}

void traceSumPush(const Peers& peers, const Layout& lay, const float* tracePartials, unsigned count, cudaStream_t stream) {
	trace_sum_push_kernel<<<1, 1024, 0, stream>>>(peers, lay.trace, tracePartials, count);
	launchCheck();
}

void finishH(const Peers& peers, const Layout& lay, const Control& ctl, unsigned k, float center, float* B, float* corrP, cudaStream_t stream) {
	finish_h_kernel<<<1, 1024, 0, stream>>>(peers, lay.flagsH, lay.statH, lay.statLen, ctl, k, center, B, corrP);
	launchCheck();
}

template <int KP>
static unsigned launchUpdateW(const Peers& peers, const Layout& lay, float center, unsigned rows, unsigned k, float* B, float* corrP, const float* inv, float* W,
                              size_t ldw, float* Whi, float* Wlo, const float* Ppart, size_t ldp, size_t slotStride, const unsigned char* slotCount, float eps,
                              float* statPart, float* statFlag, bool update, cudaStream_t stream) {
	const unsigned panels = ceilDiv(rows, 128);
	if (panels == 0) return 0;
	// one wave: as many blocks as are resident at once (shared memory: 50 KB at k = 64, 4 per SM), each walking
	// ceil(panels / blocks) panels and leaving ONE partial of the statistics
	static int sms = 0;
	if (sms == 0) {
		int dev = 0;
		CUDA_CHECK(cudaGetDevice(&dev));
		CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	}
	const unsigned perSm = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / (smemUpdateW<KP>() + 2048)));
	const unsigned resident = perSm * (unsigned)sms;
	const unsigned blocks = ceilDiv(panels, ceilDiv(panels, resident));
	if (update)
		update_w_fused<KP, true><<<blocks, 256, smemUpdateW<KP>(), stream>>>(peers, lay.statH, lay.statLen, center, rows, k, B, corrP, inv, W, ldw, Whi, Wlo, Ppart,
		                                                                     ldp, slotStride, slotCount, eps, statPart, statFlag);
	else
		update_w_fused<KP, false><<<blocks, 256, smemUpdateW<KP>(), stream>>>(peers, lay.statH, lay.statLen, center, rows, k, B, corrP, inv, W, ldw, Whi, Wlo, Ppart,
		                                                                      ldp, slotStride, slotCount, eps, statPart, statFlag);
	launchCheck();
	return blocks;
}

unsigned updateW(const Peers& peers, const Layout& lay, float center, unsigned rows, unsigned k, float* B, float* corrP, const float* inv, float* W, size_t ldw,
                 float* Whi, float* Wlo, const float* Ppart, size_t ldp, size_t slotStride, const unsigned char* slotCount, float eps, float* statPart,
                 float* statFlag, bool update, cudaStream_t stream) {
#define NMF_ARGS peers, lay, center, rows, k, B, corrP, inv, W, ldw, Whi, Wlo, Ppart, ldp, slotStride, slotCount, eps, statPart, statFlag, update, stream
	if (k <= 16) return launchUpdateW<16>(NMF_ARGS);
	if (k <= 32) return launchUpdateW<32>(NMF_ARGS);
	if (k <= 64) return launchUpdateW<64>(NMF_ARGS);
	if (k <= 128) return launchUpdateW<128>(NMF_ARGS);
#undef NMF_ARGS
	throw EngineError(ResultType::ErrorInvalidArgument, "the fused MU kernels cover ranks up to 128");
}

void collectN(const Peers& peers, const Layout& lay, unsigned k, unsigned c0, unsigned nOwn, unsigned colsPerRank, size_t ldh, unsigned slotsPerRank,
              const unsigned char* slotCount, const float* inv, const float* corrN, float* out, size_t ldo, cudaStream_t stream) {
	if (nOwn == 0) return;
	dim3 grid(ceilDiv(k, 64), std::min(nOwn, 65535u));
	collect_n_kernel<<<grid, 64, 0, stream>>>(peers, lay.slots, k, c0, nOwn, ldh, slotsPerRank, ldh * (size_t)colsPerRank, slotCount, inv, corrN, out, ldo);
	launchCheck();
}

void scaleRows(unsigned rows, unsigned rowsPadded, unsigned k, const float* W, size_t ldw, const float* inv, float* block, cudaStream_t stream) {
	dim3 grid(ceilDiv(rowsPadded, 256), k);
	scale_rows_kernel<<<grid, 256, 0, stream>>>(rows, rowsPadded, k, W, ldw, inv, block);
	launchCheck();
}

void unpackRows(unsigned m, unsigned k, unsigned rowsPadded, const float* gathered, float* W, size_t ldw, cudaStream_t stream) {
	dim3 grid(ceilDiv(m, 256), k);
	unpack_rows_kernel<<<grid, 256, 0, stream>>>(m, k, rowsPadded, gathered, W, ldw);
	launchCheck();
}

}  // namespace fused
}  // namespace b200
}  // namespace nmfgpu
