// tc_gemm.cu -- W^T V and V H^T on the Blackwell tensor cores (tcgen05 / TMEM / TMA), 3xTF32.
//
// Both products are the same machine:   D[128 x kp] += A[128 x 32] * B[kp x 32]^T   per reduction stage,
//   W^T V :  A rows = 128 columns of V, reduction over the rows of V (contiguous in memory),  B = W      (hi, lo)
//   V H^T :  A rows = 128 rows of V,    reduction over the columns of V,                     B = H^T    (hi, lo)
// computing N^T resp. N2 so that the big operand V is always the 128-row A operand (UMMA M = 128 runs
// the tensor pipe at full rate, M = 64 at half) and the rank k is the UMMA N dimension.
//
// Kernel anatomy (one persistent CTA per SM, 384 threads, stream-K work split, see tc_gemm.h):
//   warp 0      TMA producer of the V tiles (16 KB per stage, EVICT_FIRST: V is streamed once per product)
//   warp 2      TMEM allocation, then TMA producer of the B tiles (hi and lo, kp x 32 each, EVICT_LAST)
//   warp 1      MMA issuer: per stage 4 k-steps x {A_hi B_hi, A_lo B_hi, A_hi B_lo}, tcgen05.mma kind::tf32 with
//               A in TENSOR MEMORY and B in shared memory (128B swizzle); accumulators in TMEM
//   warps 4-7, 8-11   two worker warpgroups taking alternate stages: read the fp32 V tile from shared memory
//               (each thread owns one A row = one TMEM lane), split every value into TF32 hi/lo in registers
//               and tcgen05.st both halves into an A slot of TMEM.  V is never written back anywhere.
// The tensor core accumulates only `flushStages` stages at a time; the workers then tcgen05.ld the
// accumulator and add it to fp32 running sums in registers with round-to-nearest (two accumulator
// buffers, one per warpgroup, so the flush overlaps the next chunk's MMAs).  That keeps the rounding of
// a 100 000-term reduction at fp32 level regardless of how the tensor core rounds its accumulator.
#include "tc_gemm.h"

#include <cuda.h>

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"

namespace nmfgpu {
namespace b200 {
namespace tc {

namespace {

constexpr int TILE_ROWS = 128;        // A rows per tile (= TMEM lanes)
constexpr int STAGE_K = 32;           // reduction elements per stage (128 bytes of fp32: one swizzle row)
constexpr int V_STAGE_BYTES = TILE_ROWS * STAGE_K * 4;
constexpr int SLOTS = 4;              // operand slots: TMEM A ring (64 columns each: 32 hi + 32 lo) + smem B ring
constexpr int SLOTS_SHIFT = 2;
constexpr int ACC_COLS = 128;         // TMEM columns reserved per accumulator buffer
constexpr int A_BASE_COL = 2 * ACC_COLS;
constexpr int NUM_THREADS = 384;
constexpr int FLUSH_LOOKAHEAD = 2;    // stages a worker keeps splitting past a chunk end before it flushes
constexpr uint64_t POLICY_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t POLICY_EVICT_LAST = 0x14F0000000000000ull;

template <int KPM> struct Rings;
template <> struct Rings<64> { static constexpr int SV = 8; };    // V tile ring depth (16 KB each)
template <> struct Rings<128> { static constexpr int SV = 4; };

struct KParams {
	alignas(64) CUtensorMap mapV;
	alignas(64) CUtensorMap mapBhi;
	alignas(64) CUtensorMap mapBlo;
	float* out;
	unsigned long long ldOut, slotStride, units;
	unsigned rowsA, k, kp, tiles, stagesPerTile, flushStages, passes, grid;
};

// ---- stream-K bookkeeping shared by host and device ---------------------------------------------------
__host__ __device__ inline unsigned long long unitStart(unsigned cta, unsigned grid, unsigned long long units) {
	return units * cta / grid;
}
// the CTA whose range contains unit x: the largest c with floor(c * units / grid) <= x
__host__ __device__ inline unsigned ctaOfUnit(unsigned long long x, unsigned grid, unsigned long long units) {
	return (unsigned)(((x + 1) * grid - 1) / units);
}

struct Segment {
	unsigned tile, stage0, len, slot;
};

struct SegmentWalker {
	unsigned long long u, uEnd, units;
	unsigned stagesPerTile, grid, cta;
	__device__ SegmentWalker(const KParams& p, unsigned ctaIdx)
	    : u(unitStart(ctaIdx, p.grid, p.units)), uEnd(unitStart(ctaIdx + 1, p.grid, p.units)), units(p.units), stagesPerTile(p.stagesPerTile),
	      grid(p.grid), cta(ctaIdx) {}
	__device__ bool next(Segment& s) {
		if (u >= uEnd) return false;
		s.tile = (unsigned)(u / stagesPerTile);
		s.stage0 = (unsigned)(u % stagesPerTile);
		const unsigned long long room = stagesPerTile - s.stage0;
		s.len = (unsigned)((uEnd - u) < room ? (uEnd - u) : room);
		s.slot = cta - ctaOfUnit((unsigned long long)s.tile * stagesPerTile, grid, units);
		u += s.len;
		return true;
	}
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarArriveExpectTx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spins on the phase with parity `parity`; a watchdog turns a protocol bug into a trap instead of a hung GPU.
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	unsigned long long t0 = 0;
	for (uint32_t spin = 1; !done; ++spin) {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done)
		    : "r"(bar), "r"(parity)
		    : "memory");
		if (!done && (spin & 0xFF) == 0) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
			if (t0 == 0) t0 = now;
			else if (now - t0 > 4000000000ull) __trap();   // 4 s without progress
		}
	}
}
__device__ __forceinline__ void tmaLoad2D(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t policy) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
	    "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
	    : "memory");
}
__device__ __forceinline__ bool electOne() {
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcCommit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, issued by one thread for the whole CTA
__device__ __forceinline__ void mmaTf32(uint32_t d, uint32_t a, uint64_t bDesc, uint32_t iDesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
	    "r"(a), "l"(bDesc), "r"(iDesc), "r"(accumulate)
	    : "memory");
}
// 16 consecutive TMEM columns of this thread's lane <- 16 registers
__device__ __forceinline__ void tmemStore16(uint32_t addr, const uint32_t* r) {
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(addr),
	    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
	    "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
	    : "memory");
}
__device__ __forceinline__ void tmemLoad16(uint32_t addr, uint32_t* r) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
	      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
	    : "r"(addr)
	    : "memory");
}
__device__ __forceinline__ void tmemWaitStore() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmemWaitLoad() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major operand tile stored as rows of 128 bytes with the 128B
// swizzle (what TMA writes for a {32 floats, rows} box): 8-row groups are 1024 bytes apart.
__device__ __forceinline__ uint64_t smemDescSw128(uint32_t addr) {
	uint64_t d = 0;
	d |= (uint64_t)((addr & 0x3FFFF) >> 4);       // start address
	d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
	d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups
	d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
	d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
	return d;
}

// TF32 hi/lo split in integer arithmetic: hi = round-to-nearest (ties away) to 10 mantissa bits, the
// same value cvt.rna.tf32.f32 produces; lo = v - hi is exact in fp32.
__device__ __forceinline__ void splitValue(float v, uint32_t& hi, uint32_t& lo) {
	hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
	lo = __float_as_uint(v - __uint_as_float(hi));
}

struct __align__(8) Barriers {
	uint64_t vFull[8], vEmpty[8];          // V tile ring: TMA -> workers
	uint64_t full[SLOTS], empty[SLOTS];    // operand slots: A (tensor memory, written by the workers) + B (shared memory, TMA) -> MMA
	uint64_t accFull[2], accEmpty[2];      // accumulator buffers: MMA -> flushing warpgroup
	uint32_t tmemBase;
};

// ---- the kernel ----------------------------------------------------------------------------------------
// V_COLS_ARE_ROWS = true : W^T V (A rows are columns of V; V tile in smem is [128 cols][32 rows], 128B swizzle)
//                 = false: V H^T (A rows are rows of V;    V tile in smem is [32 cols][128 rows], linear)
template <int KPM, bool V_COLS_ARE_ROWS>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_stream_gemm(const __grid_constant__ KParams p) {
	constexpr int SV = Rings<KPM>::SV;                 // power of two
	constexpr int SV_SHIFT = SV == 8 ? 3 : 2;
	constexpr int B_HALF_BYTES = KPM * STAGE_K * 4;
	extern __shared__ unsigned char smemRaw[];
	unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~(uintptr_t)1023);
	unsigned char* vRing = smem;
	unsigned char* bRing = smem + SV * V_STAGE_BYTES;
	Barriers* bars = reinterpret_cast<Barriers*>(bRing + SLOTS * 2 * B_HALF_BYTES);

	const unsigned warp = threadIdx.x / 32, lane = threadIdx.x % 32;
	const unsigned F = p.flushStages;

	if (threadIdx.x == 0) {
		for (int i = 0; i < SV; ++i) {
			mbarInit(smemAddr(&bars->vFull[i]), 1);
			mbarInit(smemAddr(&bars->vEmpty[i]), 128);
		}
		for (int i = 0; i < SLOTS; ++i) {
			mbarInit(smemAddr(&bars->full[i]), 128 + 1);     // 128 worker threads (A slot written) + the B producer (expect_tx)
			mbarInit(smemAddr(&bars->empty[i]), 1);          // tcgen05.commit of the stage's MMAs
		}
		for (int i = 0; i < 2; ++i) {
			mbarInit(smemAddr(&bars->accFull[i]), 1);
			mbarInit(smemAddr(&bars->accEmpty[i]), 128);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 2) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&bars->tmemBase)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tcFenceBefore();
	__syncthreads();
	tcFenceAfter();
	const uint32_t tmem = bars->tmemBase;
	const uint32_t barBase = smemAddr(bars);
	const uint32_t vFullBar = barBase + offsetof(Barriers, vFull), vEmptyBar = barBase + offsetof(Barriers, vEmpty);
	const uint32_t fullBar = barBase + offsetof(Barriers, full), emptyBar = barBase + offsetof(Barriers, empty);
	const uint32_t accFullBar = barBase + offsetof(Barriers, accFull), accEmptyBar = barBase + offsetof(Barriers, accEmpty);

	if (warp == 0) {
		// ===== V producer =====
		if (lane == 0) {
			SegmentWalker walk(p, blockIdx.x);
			Segment s;
			unsigned g = 0;
			const uint32_t vBase = smemAddr(vRing);
			while (walk.next(s)) {
				const int rIdx = (int)(s.tile * TILE_ROWS);
				int kIdx = (int)(s.stage0 * STAGE_K);
				for (unsigned ls = 0; ls < s.len; ++ls, ++g, kIdx += STAGE_K) {
					const unsigned sv = g & (SV - 1);
					mbarWait(vEmptyBar + sv * 8, ((g >> SV_SHIFT) & 1) ^ 1);
					const uint32_t full = vFullBar + sv * 8;
					mbarArriveExpectTx(full, V_STAGE_BYTES);
					if (V_COLS_ARE_ROWS) tmaLoad2D(vBase + sv * V_STAGE_BYTES, &p.mapV, full, kIdx, rIdx, POLICY_EVICT_FIRST);
					else tmaLoad2D(vBase + sv * V_STAGE_BYTES, &p.mapV, full, rIdx, kIdx, POLICY_EVICT_FIRST);
				}
			}
		}
	} else if (warp == 2) {
		// ===== B producer (hi and lo tiles of W resp. H^T): lands on the same barrier the workers arrive on =====
		if (lane == 0) {
			SegmentWalker walk(p, blockIdx.x);
			Segment s;
			unsigned g = 0;
			const uint32_t bytes = 2u * p.kp * STAGE_K * 4u;
			const uint32_t bBase = smemAddr(bRing);
			while (walk.next(s)) {
				int kIdx = (int)(s.stage0 * STAGE_K);
				for (unsigned ls = 0; ls < s.len; ++ls, ++g, kIdx += STAGE_K) {
					const unsigned sl = g & (SLOTS - 1);
					mbarWait(emptyBar + sl * 8, ((g >> SLOTS_SHIFT) & 1) ^ 1);
					const uint32_t full = fullBar + sl * 8;
					mbarArriveExpectTx(full, bytes);
					const uint32_t dst = bBase + sl * 2 * B_HALF_BYTES;
					tmaLoad2D(dst, &p.mapBhi, full, kIdx, 0, POLICY_EVICT_LAST);
					tmaLoad2D(dst + B_HALF_BYTES, &p.mapBlo, full, kIdx, 0, POLICY_EVICT_LAST);
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer =====
		// The whole warp walks the loop convergently so that every address and descriptor lives in uniform
		// registers; one elected lane issues the 12 MMAs of a stage back to back (a divergent single-thread
		// loop costs ~140 cycles per MMA in R2UR traffic; this form issues at the tensor pipe's 32 cycles).
		const bool leader = electOne();
		// instruction descriptor: D fp32, A/B tf32, both K-major, N = kp, M = 128
		const uint32_t iDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.kp >> 3) << 17) | ((TILE_ROWS >> 4) << 24);
		const uint64_t bDesc0 = smemDescSw128(smemAddr(bRing));
		const bool threePass = p.passes == 3;
		SegmentWalker walk(p, blockIdx.x);
		Segment s;
		unsigned g = 0, gc = 0;
		while (walk.next(s)) {
			unsigned inChunk = 0;
			uint32_t acc = 0;
			for (unsigned ls = 0; ls < s.len; ++ls, ++g) {
				if (inChunk == 0) {
					const unsigned b = gc & 1;
					mbarWait(accEmptyBar + b * 8, ((gc >> 1) & 1) ^ 1);
					acc = tmem + b * ACC_COLS;
				}
				const unsigned sl = g & (SLOTS - 1);
				mbarWait(fullBar + sl * 8, (g >> SLOTS_SHIFT) & 1);
				tcFenceAfter();
				if (leader) {
					const uint32_t aHi = tmem + A_BASE_COL + sl * 64, aLo = aHi + 32;
					const uint64_t dHi = bDesc0 + (uint64_t)((sl * 2 * B_HALF_BYTES) >> 4), dLo = dHi + (B_HALF_BYTES >> 4);
					if (threePass) {
#pragma unroll
						for (int q = 0; q < STAGE_K / 8; ++q) {
							mmaTf32(acc, aHi + q * 8, dHi + q * 2, iDesc, q == 0 ? (uint32_t)(inChunk != 0) : 1u);
							mmaTf32(acc, aLo + q * 8, dHi + q * 2, iDesc, 1);
							mmaTf32(acc, aHi + q * 8, dLo + q * 2, iDesc, 1);
						}
					} else {
#pragma unroll
						for (int q = 0; q < STAGE_K / 8; ++q) mmaTf32(acc, aHi + q * 8, dHi + q * 2, iDesc, q == 0 ? (uint32_t)(inChunk != 0) : 1u);
					}
					tcCommit(emptyBar + sl * 8);
				}
				++inChunk;
				if (inChunk == F || ls + 1 == s.len) {
					if (leader) tcCommit(accFullBar + (gc & 1) * 8);
					++gc;
					inChunk = 0;
				}
				__syncwarp();
			}
		}
	} else if (warp >= 4) {
		// ===== workers: V tile -> TF32 hi/lo -> TMEM A slot; accumulator flush; output =====
		const unsigned wg = (warp - 4) / 4;                 // 0 or 1
		const unsigned row = (warp % 4) * 32 + lane;        // A row = TMEM lane owned by this thread
		const uint32_t laneBase = ((warp % 4) * 32) << 16;
		const unsigned kp = p.kp;
		float sum[KPM];
#pragma unroll
		for (int c = 0; c < KPM; ++c) sum[c] = 0.f;

		auto split = [&](unsigned g) {
			const unsigned sv = g & (SV - 1), sl = g & (SLOTS - 1);
			mbarWait(vFullBar + sv * 8, (g >> SV_SHIFT) & 1);
			const unsigned char* tile = vRing + sv * V_STAGE_BYTES;
			float v[STAGE_K];
			if (V_COLS_ARE_ROWS) {
				// row `row` of the [128][32] tile; 16-byte chunk c lives at chunk c ^ (row & 7)
				const unsigned char* base = tile + row * 128;
#pragma unroll
				for (int c = 0; c < 8; ++c) {
					const float4 x = *reinterpret_cast<const float4*>(base + ((c ^ (row & 7)) << 4));
					v[4 * c + 0] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
				}
			} else {
				const float* base = reinterpret_cast<const float*>(tile) + row;
#pragma unroll
				for (int j = 0; j < STAGE_K; ++j) v[j] = base[j * TILE_ROWS];
			}
			mbarArrive(vEmptyBar + sv * 8);                     // the tile is in registers: release the slot
			uint32_t hi[STAGE_K], lo[STAGE_K];
#pragma unroll
			for (int e = 0; e < STAGE_K; ++e) splitValue(v[e], hi[e], lo[e]);
			mbarWait(emptyBar + sl * 8, ((g >> SLOTS_SHIFT) & 1) ^ 1);
			tcFenceAfter();
			const uint32_t aSlot = tmem + laneBase + A_BASE_COL + sl * 64;
			tmemStore16(aSlot, hi);
			tmemStore16(aSlot + 16, hi + 16);
			tmemStore16(aSlot + 32, lo);
			tmemStore16(aSlot + 48, lo + 16);
			tmemWaitStore();
			tcFenceBefore();
			mbarArrive(fullBar + sl * 8);
		};

		auto flush = [&](unsigned gc) {
			const unsigned b = gc & 1;
			mbarWait(accFullBar + b * 8, (gc >> 1) & 1);
			tcFenceAfter();
			const uint32_t acc = tmem + laneBase + b * ACC_COLS;
#pragma unroll
			for (int q = 0; q < KPM / 32; ++q) {
				if (q * 32 < (int)kp) {
					uint32_t r[32];
					tmemLoad16(acc + q * 32, r);
					if (q * 32 + 16 < (int)kp) tmemLoad16(acc + q * 32 + 16, r + 16);
					tmemWaitLoad();
#pragma unroll
					for (int e = 0; e < 16; ++e) sum[q * 32 + e] += __uint_as_float(r[e]);
					if (q * 32 + 16 < (int)kp) {
#pragma unroll
						for (int e = 16; e < 32; ++e) sum[q * 32 + e] += __uint_as_float(r[e]);
					}
				}
			}
			tcFenceBefore();
			mbarArrive(accEmptyBar + b * 8);
		};

		SegmentWalker walk(p, blockIdx.x);
		Segment s;
		unsigned gBase = 0, gcBase = 0;
		while (walk.next(s)) {
			// chunk c of this segment covers local stages [c F, (c+1) F); global chunk gcBase + c is flushed by
			// warpgroup (gcBase + c) & 1 once that warpgroup has split past the chunk's end
			const unsigned nChunks = (s.len + F - 1) / F;
			unsigned nextChunk = (gcBase & 1) == wg ? 0u : 1u;          // my first chunk of this segment
			unsigned flushAt = (nextChunk + 1) * F + FLUSH_LOOKAHEAD;    // local stage index from which it may be flushed
			for (unsigned ls = (gBase & 1) == wg ? 0u : 1u; ls < s.len; ls += 2) {
				split(gBase + ls);
				if (ls >= flushAt && nextChunk + 1 < nChunks) {
					flush(gcBase + nextChunk);
					nextChunk += 2;
					flushAt += 2 * F;
				}
			}
			for (; nextChunk < nChunks; nextChunk += 2) flush(gcBase + nextChunk);
			gBase += s.len;
			gcBase += nChunks;

			// ---- output of this segment's partial product: warpgroup 0 stores, warpgroup 1 adds its share
			float* out = p.out + (size_t)s.slot * p.slotStride;
			const unsigned r = s.tile * TILE_ROWS + row;
			for (unsigned phase = 0; phase < 2; ++phase) {
				if (phase == wg && r < p.rowsA) {
					if (V_COLS_ARE_ROWS) {
						float4* dst = reinterpret_cast<float4*>(out + (size_t)r * p.ldOut);     // column r of N: kp contiguous values
#pragma unroll
						for (int c = 0; c < KPM / 4; ++c) {
							if (c * 4 < (int)kp) {
								float4 x = make_float4(sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]);
								if (phase == 1) {
									const float4 y = dst[c];
									x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
								}
								dst[c] = x;
							}
						}
					} else {
#pragma unroll
						for (int c = 0; c < KPM; ++c) {
							if (c < (int)p.k) {
								float* dst = out + (size_t)c * p.ldOut + r;                      // row r of N2: coalesced across the warp
								*dst = phase == 1 ? *dst + sum[c] : sum[c];
							}
						}
					}
				}
				if (phase == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
			}
#pragma unroll
			for (int c = 0; c < KPM; ++c) sum[c] = 0.f;
		}
	}

	tcFenceBefore();
	__syncthreads();
	if (warp == 2) {
		tcFenceAfter();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
	}
}

template <int KPM>
size_t smemBytes() {
	return 1024 + (size_t)Rings<KPM>::SV * V_STAGE_BYTES + (size_t)SLOTS * 2 * KPM * STAGE_K * 4 + sizeof(Barriers);
}

// ---- H -> H^T hi/lo ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_transpose_kernel(unsigned k, unsigned n, const float* __restrict__ H, size_t ldh, float* __restrict__ hi,
                                                             float* __restrict__ lo, size_t ldht) {
	__shared__ float tile[32][33];
	const unsigned j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
	const unsigned tx = threadIdx.x % 32, ty = threadIdx.x / 32;
	for (unsigned jj = ty; jj < 32; jj += 8) {
		const unsigned j = j0 + jj, c = c0 + tx;
		tile[jj][tx] = (j < n && c < k) ? H[(size_t)j * ldh + c] : 0.f;
	}
	__syncthreads();
	for (unsigned cc = ty; cc < 32; cc += 8) {
		const unsigned c = c0 + cc, j = j0 + tx;
		if (c < k && j < n) {
			const float v = tile[tx][cc];
			uint32_t h, l;
			splitValue(v, h, l);
			hi[(size_t)c * ldht + j] = __uint_as_float(h);
			lo[(size_t)c * ldht + j] = __uint_as_float(l);
		}
	}
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled() {
	static EncodeTiledFn fn = nullptr;
	if (fn == nullptr) {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
		if (p == nullptr || q != cudaDriverEntryPointSuccess) throw EngineError(ResultType::ErrorExternalLibrary, "cuTensorMapEncodeTiled is not available");
		fn = reinterpret_cast<EncodeTiledFn>(p);
	}
	return fn;
}

// 2-D fp32 tensor map over a column-major matrix: dim0 = rows (contiguous), dim1 = columns (stride ld)
void makeMap(unsigned char* out, const float* base, unsigned rows, unsigned cols, size_t ld, unsigned boxRows, unsigned boxCols, bool swizzle128) {
	CUtensorMap map;
	const cuuint64_t dims[2] = {rows, cols};
	const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
	const cuuint32_t box[2] = {boxRows, boxCols};
	const cuuint32_t elem[2] = {1, 1};
	const CUresult r = encodeTiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem,
	                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
	                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		char buf[160];
		snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d) for a %u x %u matrix, ld %zu, box %u x %u", (int)r, rows, cols, ld, boxRows, boxCols);
		throw EngineError(ResultType::ErrorExternalLibrary, buf);
	}
	static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
	memcpy(out, &map, sizeof(map));
}

int smCount() {
	int dev = 0, sms = 0;
	CUDA_CHECK(cudaGetDevice(&dev));
	CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	return sms;
}

void planProduct(Product& prod, unsigned rowsA, unsigned kdim) {
	prod.tiles = ceilDiv(rowsA, TILE_ROWS);
	prod.stagesPerTile = ceilDiv(kdim, STAGE_K);
	const unsigned long long units = (unsigned long long)prod.tiles * prod.stagesPerTile;
	prod.grid = (unsigned)std::min<unsigned long long>(units, (unsigned long long)smCount());
	std::vector<unsigned char> counts(prod.tiles);
	prod.maxSlots = 1;
	for (unsigned t = 0; t < prod.tiles; ++t) {
		const unsigned first = ctaOfUnit((unsigned long long)t * prod.stagesPerTile, prod.grid, units);
		const unsigned last = ctaOfUnit((unsigned long long)(t + 1) * prod.stagesPerTile - 1, prod.grid, units);
		counts[t] = (unsigned char)(last - first + 1);
		prod.maxSlots = std::max(prod.maxSlots, last - first + 1);
	}
	if (prod.slotCount) cudaFree(prod.slotCount);
	CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&prod.slotCount), prod.tiles));
	CUDA_CHECK(cudaMemcpy(prod.slotCount, counts.data(), prod.tiles, cudaMemcpyHostToDevice));
}

template <int KPM, bool VC>
void launch(const Plan& plan, const Product& prod, unsigned rowsA, float* out, size_t ldOut, size_t slotStride, cudaStream_t stream) {
	static bool configured = false;
	const size_t smem = smemBytes<KPM>();
	if (!configured) {
		CUDA_CHECK(cudaFuncSetAttribute(tc_stream_gemm<KPM, VC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		configured = true;
	}
	KParams p;
	memcpy(&p.mapV, prod.mapV, 128);
	memcpy(&p.mapBhi, prod.mapBhi, 128);
	memcpy(&p.mapBlo, prod.mapBlo, 128);
	p.out = out;
	p.ldOut = ldOut;
	p.slotStride = slotStride;
	p.units = (unsigned long long)prod.tiles * prod.stagesPerTile;
	p.rowsA = rowsA;
	p.k = plan.k;
	p.kp = plan.kp;
	p.tiles = prod.tiles;
	p.stagesPerTile = prod.stagesPerTile;
	p.flushStages = plan.flushStages;
	p.passes = plan.passes;
	p.grid = prod.grid;
	tc_stream_gemm<KPM, VC><<<prod.grid, NUM_THREADS, smem, stream>>>(p);
	CUDA_CHECK(cudaGetLastError());
}

}  // namespace

Plan::~Plan() {
	if (wtv.slotCount) cudaFree(wtv.slotCount);
	if (vht.slotCount) cudaFree(vht.slotCount);
}

bool shapeSupported(unsigned m, unsigned n, unsigned k, size_t ldV, size_t ldW) {
	if (k == 0 || k > 128 || m == 0 || n == 0) return false;
	if (ldV % 4 != 0 || ldW % 4 != 0) return false;   // TMA: 16-byte global strides
	int dev = 0, major = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
	return major == 10;
}

void makePlan(Plan& plan, unsigned m, unsigned n, unsigned k, const float* V, size_t ldV, const float* Whi, const float* Wlo, size_t ldW,
              const float* HtHi, const float* HtLo, size_t ldHt, bool singlePass) {
	plan.m = m;
	plan.n = n;
	plan.k = k;
	plan.kp = (unsigned)roundUp(k, 16);
	plan.passes = singlePass ? 1 : 3;
	plan.flushStages = 8;
	if (const char* e = getenv("NMFGPU_TC_FLUSH_STAGES")) {   // tuning knob: 0 = accumulate whole segments inside the tensor core
		const long v = strtol(e, nullptr, 10);
		plan.flushStages = v <= 0 ? 0x40000000u : (unsigned)v;
	}
	// W^T V: A rows = columns of V, reduction over m
	planProduct(plan.wtv, n, m);
	makeMap(plan.wtv.mapV, V, m, n, ldV, STAGE_K, TILE_ROWS, true);
	makeMap(plan.wtv.mapBhi, Whi, m, k, ldW, STAGE_K, plan.kp, true);
	makeMap(plan.wtv.mapBlo, Wlo, m, k, ldW, STAGE_K, plan.kp, true);
	// V H^T: A rows = rows of V, reduction over n
	planProduct(plan.vht, m, n);
	makeMap(plan.vht.mapV, V, m, n, ldV, TILE_ROWS, STAGE_K, false);
	makeMap(plan.vht.mapBhi, HtHi, n, k, ldHt, STAGE_K, plan.kp, true);
	makeMap(plan.vht.mapBlo, HtLo, n, k, ldHt, STAGE_K, plan.kp, true);
}

void gemmWtV(const Plan& plan, float* Npart, size_t ldn, size_t slotStride, cudaStream_t stream) {
	if (plan.kp <= 64) launch<64, true>(plan, plan.wtv, plan.n, Npart, ldn, slotStride, stream);
	else launch<128, true>(plan, plan.wtv, plan.n, Npart, ldn, slotStride, stream);
}

void gemmVHt(const Plan& plan, float* Ppart, size_t ldp, size_t slotStride, cudaStream_t stream) {
	if (plan.kp <= 64) launch<64, false>(plan, plan.vht, plan.m, Ppart, ldp, slotStride, stream);
	else launch<128, false>(plan, plan.vht, plan.m, Ppart, ldp, slotStride, stream);
}

void splitTransposeH(unsigned k, unsigned n, const float* H, size_t ldh, float* hi, float* lo, size_t ldht, cudaStream_t stream) {
	const dim3 grid(ceilDiv(n, 32), ceilDiv(k, 32));
	split_transpose_kernel<<<grid, 256, 0, stream>>>(k, n, H, ldh, hi, lo, ldht);
	CUDA_CHECK(cudaGetLastError());
}

}  // namespace tc
}  // namespace b200
}  // namespace nmfgpu
