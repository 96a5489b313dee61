// tc_gemm.cu -- W^T V and V H^T on the Blackwell tensor cores (tcgen05 / TMEM / TMA), 3xTF32.
//
// Both products are the same machine:   D[256 x kp] += A[256 x 32] * B[kp x 32]^T   per reduction stage,
//   W^T V :  A rows = 256 columns of V, reduction over the rows of V (contiguous in memory),  B = W      (hi, lo)
//   V H^T :  A rows = 256 rows of V,    reduction over the columns of V,                     B = H^T    (hi, lo)
// computing N^T resp. N2 so that the big operand V is always the A operand in two 128-row UMMA tiles
// (M = 128 runs the tensor pipe at full rate, M = 64 at half) and the rank k is the UMMA N dimension.
// Two A tiles share every B tile: B comes out of L2, and its traffic and its latency -- not HBM -- bound
// the one-tile version of this kernel (profiles/r01_notes.md).
//
// Kernel anatomy (one persistent CTA per SM, 640 threads, stream-K work split in reduction chunks, see tc_gemm.h):
//   warp 0      TMA producer of the V tiles (16 KB each, two per stage, one lane and one ring per A tile; EVICT_FIRST: V
//               is streamed once per product)
//   warp 2      TMEM allocation, then TMA producer of the B tiles (hi and lo, kp x 32 each, EVICT_LAST)
//   warps 1, 3  MMA issuers, one per A tile: 4 k-steps x {A_hi B_hi, A_lo B_hi, A_hi B_lo}, tcgen05.mma kind::tf32 with
//               A in TENSOR MEMORY and B in shared memory (128B swizzle); accumulators in TMEM
//   warps 4-11  two splitter warpgroups, one per A tile: read the fp32 V tile from shared memory (each thread owns one
//               A row = one TMEM lane), subtract the centre, split every value into TF32 hi/lo in registers and
//               tcgen05.st both halves into an A slot of TMEM.  V is never written back anywhere.
//   warps 12-19 two flusher warpgroups, one per A tile: own the fp32 running sums and write the partial products.
// The tensor core accumulates only `flushStages` stages at a time; the flushers then tcgen05.ld the tile's accumulator
// and add it to the running sums in registers with round-to-nearest (double-buffered accumulators for kp <= 64, so the
// flush overlaps the next chunk's MMAs).  The tensor core truncates its fp32 accumulator after every MMA (measured: bias
// of -2.4e-7 per accumulated stage), so a 100 000-term reduction kept in TMEM would be off by 4e-4; chunked and centred
// it stays at fp32 level (tc_gemm.h).
#include "tc_gemm.h"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
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
constexpr int PAIR_ROWS = 2 * TILE_ROWS;   // A rows per stream-K tile: two UMMA tiles that share their B tiles
constexpr int SLOTS = 4;              // TMEM A operand ring, in tiles (64 columns each: 32 hi + 32 lo)
constexpr int SLOTS_SHIFT = 2;
constexpr int SVH = 5;                // shared-memory ring of V tiles (16 KB each): SVH slots per A tile (= per splitter warpgroup)
constexpr int SV = 2 * SVH;
constexpr int A_BASE_COL = 256;       // TMEM columns [0, 256): accumulators, [256, 512): A slots
constexpr uint64_t POLICY_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t POLICY_EVICT_LAST = 0x14F0000000000000ull;

template <int KPM> struct Rings;
// shared-memory rings: SV tiles of V (16 KB each) and SB tile pairs of B (hi + lo, KPM x 32 fp32 each).  Both
// must cover the loaded TMA latency (~2700 cycles measured under full HBM traffic) at one stage per 400-650 cycles.
// SB: shared-memory ring of B tile pairs (hi + lo, KPM x 32 fp32 each); ACC_BUFS: accumulator buffers per A tile.
// Warp roles beyond the four helpers: SPLIT_WGS warpgroups turn V tiles into TMEM A slots (round robin over the
// tile steps), FLUSH_WGS warpgroups own the running sums.  Register budget: setmaxnreg only redistributes what the
// CTA was launched with, 640 threads x 96 = 61440 (an .inc beyond that pool never returns):
//   KPM  64: 128 x 48 + 2 x 128 x 80 + 2 x 128 x 136 = 61440      KPM 128: 128 x 40 + 2 x 128 x 72 + 2 x 128 x 144 = 60416
// SPLIT_WGS must divide the number of A slots: a warpgroup has to be the only writer of its slots, because the parity
// wait on `empty` only tells two consecutive phases apart (three warpgroups on four slots let a fast one get two
// phases ahead of a slow one and overwrite a slot that was never consumed -- seen as a deadlock, profiles/r01_notes.md)
template <> struct Rings<64> { static constexpr int SB = 3, ACC_BUFS = 2, SPLIT_WGS = 2, FLUSH_WGS = 2, HELPER_REGS = 48, SPLIT_REGS = 80, FLUSH_REGS = 136, THREADS = 640; };
template <> struct Rings<128> { static constexpr int SB = 2, ACC_BUFS = 1, SPLIT_WGS = 2, FLUSH_WGS = 2, HELPER_REGS = 40, SPLIT_REGS = 72, FLUSH_REGS = 144, THREADS = 640; };

// position in a ring of arbitrary depth: slot index plus the parity of the number of completed laps
struct RingPos {
	unsigned idx = 0, lap = 0;
	__device__ __forceinline__ void advance(unsigned depth) {
		if (++idx == depth) {
			idx = 0;
			lap ^= 1;
		}
	}
};

struct KParams {
	alignas(64) CUtensorMap mapV;
	alignas(64) CUtensorMap mapBhi;
	alignas(64) CUtensorMap mapBlo;
	float* out;
	unsigned long long* trace;   // optional timeline of CTA 0 (NMFGPU_TC_TRACE), nullptr otherwise
	unsigned long long ldOut, slotStride, units;
	unsigned rowsA, k, kp, tiles, stagesPerTile, flushStages, passes, grid;
	unsigned chunks, chunkStages;   // reduction chunks (tc_gemm.h): `units` = tiles * chunkStages units per chunk
	float center;   // subtracted from every element of V before the split (see tc_gemm.h)
	unsigned prefetchStages;   // how many stages ahead of the TMA loads the V tiles are prefetched into L2
	// Gate of the small operand (tc_gemm.h Gate): the B producer waits until gateFlags[0 .. gateCount) have all reached
	// *gateEpoch before its first load -- the other ranks are still storing H^T into this GPU's memory when the kernel starts
	const unsigned* gateFlags;
	const unsigned* gateEpoch;
	unsigned* gateError;
	unsigned gateCount;
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

// Every CTA takes the same share [units * cta / grid, units * (cta + 1) / grid) of EVERY chunk, one chunk after the other
// (tc_gemm.h), so all CTAs sweep the same part of the reduction range at the same time.  The last chunk of a tile may be
// padded; a segment that lies entirely in the padding has len 0 and still owns a slot (it writes zeros).
struct SegmentWalker {
	unsigned long long u, uBegin, uEnd, units;
	unsigned stagesPerTile, chunkStages, chunks, chunk, grid, cta;
	__host__ __device__ SegmentWalker(unsigned long long unitsPerChunk, unsigned gridSize, unsigned stagesPerTile_, unsigned chunkStages_, unsigned chunks_,
	                                  unsigned ctaIdx)
	    : u(unitStart(ctaIdx, gridSize, unitsPerChunk)), uBegin(u), uEnd(unitStart(ctaIdx + 1, gridSize, unitsPerChunk)), units(unitsPerChunk),
	      stagesPerTile(stagesPerTile_), chunkStages(chunkStages_), chunks(chunks_), chunk(0), grid(gridSize), cta(ctaIdx) {}
	__device__ SegmentWalker(const KParams& p, unsigned ctaIdx) : SegmentWalker(p.units, p.grid, p.stagesPerTile, p.chunkStages, p.chunks, ctaIdx) {}
	__host__ __device__ bool next(Segment& s) {
		if (u >= uEnd) {
			if (++chunk >= chunks || uBegin >= uEnd) return false;
			u = uBegin;
		}
		s.tile = (unsigned)(u / chunkStages);
		const unsigned a = (unsigned)(u % chunkStages);
		const unsigned long long room = chunkStages - a;
		const unsigned len = (unsigned)((uEnd - u) < room ? (uEnd - u) : room);
		s.stage0 = chunk * chunkStages + a;
		s.len = s.stage0 >= stagesPerTile ? 0u : (len < stagesPerTile - s.stage0 ? len : stagesPerTile - s.stage0);
		// the CTAs that share a tile are the same in every chunk: `perTile` partial products per chunk
		const unsigned long long f = (unsigned long long)s.tile * chunkStages;
		const unsigned firstCta = ctaOfUnit(f, grid, units), perTile = ctaOfUnit(f + chunkStages - 1, grid, units) - firstCta + 1;
		s.slot = chunk * perTile + (cta - firstCta);
		u += len;
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
// barrier waits that timed out: count, then up to 63 records {barrier smem address, parity, thread, block}; read by the
// host with NMFGPU_TC_DEBUG=1
__device__ unsigned g_waitTimeout[4 * 64];

// Spins on the phase with parity `parity`; a watchdog turns a protocol bug into a recorded timeout (the wait is
// abandoned, the results are garbage, the next launch with NMFGPU_TC_DEBUG=1 reports it) instead of a hung GPU.
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
			else if (now - t0 > 2000000000ull) {            // 2 s without progress
				if ((threadIdx.x & 31) == 0 || (threadIdx.x < 128)) {
					const unsigned slot = atomicAdd(&g_waitTimeout[0], 1u) + 1u;
					if (slot < 64) {
						g_waitTimeout[4 * slot + 0] = bar;
						g_waitTimeout[4 * slot + 1] = parity;
						g_waitTimeout[4 * slot + 2] = threadIdx.x;
						g_waitTimeout[4 * slot + 3] = blockIdx.x;
					}
				}
				asm volatile("exit;");   // this thread gives up; the others follow within their own 2 s
			}
		}
	}
}
__device__ __forceinline__ void tmaLoad2D(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t policy) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
	    "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
	    : "memory");
}
__device__ __forceinline__ void tmaPrefetchL2(const CUtensorMap* map, int c0, int c1) {
	asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
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

constexpr unsigned TRACE_STAGES = 512, TRACE_EVENTS = 16;
// Timeline hooks are compiled in only with -DNMFGPU_TC_TRACE_BUILD: even a never-taken hook in the worker loop
// costs double-digit percents (measured: five extra hooks took the products from 0.93 to 1.6 ms).
__device__ __forceinline__ void traceEvent(unsigned long long* trace, unsigned g, unsigned ev) {
#ifdef NMFGPU_TC_TRACE_BUILD
	if (trace != nullptr && blockIdx.x == 0 && g < TRACE_STAGES) trace[g * TRACE_EVENTS + ev] = clock64();
#endif
}

struct __align__(8) Barriers {
	uint64_t vFull[SV], vEmpty[SV];        // V tile rings: TMA -> splitters (slot 2 i + w belongs to A tile w)
	uint64_t bFull[4], bEmpty[4];          // B tile ring: TMA -> MMA
	uint64_t full[SLOTS], empty[SLOTS];    // A operand slots in tensor memory: splitters -> MMA
	uint64_t accFull[2], accEmpty[2];      // accumulator buffers: MMA -> flushers
	uint32_t tmemBase;
};

// ---- the kernel ----------------------------------------------------------------------------------------
// V_COLS_ARE_ROWS = true : W^T V (A rows are columns of V; V tile in smem is [128 cols][32 rows], 128B swizzle)
//                 = false: V H^T (A rows are rows of V;    V tile in smem is [32 cols][128 rows], linear)
// Counters: g = stage of this CTA (32 reduction elements of one 256-row pair tile), tile step t = 2 g + w for
// the A tile w of that stage; V ring slot = t mod 8, A slot = t mod 4, B slot = g mod SB.
template <int KPM, bool V_COLS_ARE_ROWS>
__global__ void __launch_bounds__(Rings<KPM>::THREADS, 1) tc_stream_gemm(const __grid_constant__ KParams p) {
	using R = Rings<KPM>;
	constexpr int SB = R::SB, ACC_BUFS = R::ACC_BUFS, SPLIT_WGS = R::SPLIT_WGS, FLUSH_WGS = R::FLUSH_WGS;
	static_assert(4 + 4 * (SPLIT_WGS + FLUSH_WGS) == R::THREADS / 32, "warp roles");
	static_assert(SPLIT_WGS == 2, "splitter warpgroup w owns the A slots {w, w + 2} and the V ring of A tile w");
	constexpr int B_HALF_BYTES = KPM * STAGE_K * 4;
	extern __shared__ unsigned char smemRaw[];
	unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~(uintptr_t)1023);
	unsigned char* vRing = smem;
	unsigned char* bRing = smem + SV * V_STAGE_BYTES;
	Barriers* bars = reinterpret_cast<Barriers*>(bRing + SB * 2 * B_HALF_BYTES);

	const unsigned warp = threadIdx.x / 32, lane = threadIdx.x % 32;
	const unsigned F = p.flushStages;

	if (threadIdx.x == 0) {
		for (int i = 0; i < SV; ++i) {
			mbarInit(smemAddr(&bars->vFull[i]), 1);
			mbarInit(smemAddr(&bars->vEmpty[i]), 4);         // one arrival per warp of the warpgroup that read the tile
		}
		for (int i = 0; i < SB; ++i) {
			mbarInit(smemAddr(&bars->bFull[i]), 1);
			mbarInit(smemAddr(&bars->bEmpty[i]), 2);         // tcgen05.commit of both issuers' MMAs of the stage
		}
		for (int i = 0; i < SLOTS; ++i) {
			mbarInit(smemAddr(&bars->full[i]), 4);           // one arrival per warp once its 32 lanes of the A slot are written
			mbarInit(smemAddr(&bars->empty[i]), 1);          // tcgen05.commit of the tile's MMAs
		}
		for (int i = 0; i < 2; ++i) {
			mbarInit(smemAddr(&bars->accFull[i]), 2);        // tcgen05.commit of both issuers' last MMAs of the chunk
			mbarInit(smemAddr(&bars->accEmpty[i]), 4 * FLUSH_WGS);   // the flusher warps have read the accumulators
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
	const uint32_t bFullBar = barBase + offsetof(Barriers, bFull), bEmptyBar = barBase + offsetof(Barriers, bEmpty);
	const uint32_t fullBar = barBase + offsetof(Barriers, full), emptyBar = barBase + offsetof(Barriers, empty);
	const uint32_t accFullBar = barBase + offsetof(Barriers, accFull), accEmptyBar = barBase + offsetof(Barriers, accEmpty);

	if (warp == 0) {
		// ===== V producer =====
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::HELPER_REGS));
		if (lane < 2) {
			// lane w loads the tiles of A tile w into that tile's own ring: a slow splitter warpgroup only ever holds
			// back its own loads
			const unsigned w = lane;
			SegmentWalker walk(p, blockIdx.x);
			Segment s;
			RingPos v;
			const uint32_t vBase = smemAddr(vRing);
			while (walk.next(s)) {
				const int rIdx = (int)(s.tile * PAIR_ROWS + w * TILE_ROWS);
				int kIdx = (int)(s.stage0 * STAGE_K);
				for (unsigned ls = 0; ls < s.len; ++ls, kIdx += STAGE_K, v.advance(SVH)) {
					const unsigned sv = 2 * v.idx + w;
					mbarWait(vEmptyBar + sv * 8, v.lap ^ 1);
					const uint32_t full = vFullBar + sv * 8;
#ifdef NMFGPU_TC_TRACE_BUILD
					if (p.passes & 0x2000) {   // ablation: no V loads (timing experiments, results are garbage)
						mbarArrive(full);
						continue;
					}
#endif
					mbarArriveExpectTx(full, V_STAGE_BYTES);
					if (V_COLS_ARE_ROWS) tmaLoad2D(vBase + sv * V_STAGE_BYTES, &p.mapV, full, kIdx, rIdx, POLICY_EVICT_FIRST);
					else tmaLoad2D(vBase + sv * V_STAGE_BYTES, &p.mapV, full, rIdx, kIdx, POLICY_EVICT_FIRST);
				}
			}
		}
	} else if (warp == 2) {
		// ===== B producer (hi and lo tiles of W resp. H^T) =====
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::HELPER_REGS));
		if (lane == 0) {
			if (p.gateCount != 0) {
				// poll the flags (stored by other GPUs, like the data behind them), acquire at system scope once they are all
				// there, then order the TMA reads -- async proxy -- behind it.  A rank that never signals must not hang the GPU: give up after 10 s.
				const unsigned epoch = *reinterpret_cast<volatile const unsigned*>(p.gateEpoch);
				unsigned long long t0 = 0;
				for (unsigned g = 0; g < p.gateCount; ++g) {
					unsigned spins = 0, seen;
					for (;;) {
						asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.gateFlags + g) : "memory");
						if ((int)(seen - epoch) >= 0) break;
						if ((++spins & 0x3FF) == 0) {
							unsigned long long now;
							asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
							if (t0 == 0) t0 = now;
							else if (now - t0 > 10000000000ull) {
								atomicExch(p.gateError, 1u);
								break;
							}
						}
					}
				}
				asm volatile("fence.acq_rel.sys;" ::: "memory");   // one acquire fence after all flags (relaxed polls)
				asm volatile("fence.proxy.async;" ::: "memory");
			}
			SegmentWalker walk(p, blockIdx.x);
			Segment s;
			RingPos b;
			const uint32_t bytes = 2u * p.kp * STAGE_K * 4u;
			const uint32_t bBase = smemAddr(bRing);
			while (walk.next(s)) {
				int kIdx = (int)(s.stage0 * STAGE_K);
				for (unsigned ls = 0; ls < s.len; ++ls, kIdx += STAGE_K, b.advance(SB)) {
					mbarWait(bEmptyBar + b.idx * 8, b.lap ^ 1);
					const uint32_t full = bFullBar + b.idx * 8;
#ifdef NMFGPU_TC_TRACE_BUILD
					if (p.passes & 0x1000) {   // ablation: no B loads
						mbarArrive(full);
						continue;
					}
#endif
					mbarArriveExpectTx(full, bytes);
					const uint32_t dst = bBase + b.idx * 2 * B_HALF_BYTES;
					tmaLoad2D(dst, &p.mapBhi, full, kIdx, 0, POLICY_EVICT_LAST);
					tmaLoad2D(dst + B_HALF_BYTES, &p.mapBlo, full, kIdx, 0, POLICY_EVICT_LAST);
				}
			}
		}
	} else if (warp == 1 || warp == 3) {
		// ===== MMA issuers: warp 1 feeds A tile 0 of every stage, warp 3 A tile 1 (separate accumulators) =====
		// Each warp walks its loop convergently so that every address and descriptor lives in uniform registers
		// (a divergent single-thread loop costs ~140 cycles per MMA in R2UR traffic); one elected lane issues the
		// 12 MMAs of a tile back to back.  Issuing blocks for about as long as the MMAs execute and every barrier
		// probe costs ~100 cycles, so one issuer leaves the tensor pipe idle during its waits; two issuers cover
		// each other's waits.
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::HELPER_REGS));
		const unsigned w = warp == 1 ? 0u : 1u;
		const bool leader = electOne();
		// instruction descriptor: D fp32, A/B tf32, both K-major, N = kp, M = 128
		const uint32_t iDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.kp >> 3) << 17) | ((TILE_ROWS >> 4) << 24);
		const uint64_t bDesc0 = smemDescSw128(smemAddr(bRing));
		const bool threePass = (p.passes & 0xFF) == 3;
		SegmentWalker walk(p, blockIdx.x);
		Segment s;
		unsigned g = 0, gc = 0;
		RingPos b;
		while (walk.next(s)) {
			unsigned inChunk = 0, buf = 0;
			for (unsigned ls = 0; ls < s.len; ++ls, ++g, b.advance(SB)) {
				if (inChunk == 0) {
					buf = gc % ACC_BUFS;
					mbarWait(accEmptyBar + buf * 8, (((gc / ACC_BUFS) & 1) ^ 1));
				}
				const unsigned t = 2 * g + w, sl = t & (SLOTS - 1);
				mbarWait(bFullBar + b.idx * 8, b.lap);
				mbarWait(fullBar + sl * 8, (t >> SLOTS_SHIFT) & 1);
				tcFenceAfter();
				if (leader) {
					const uint64_t dHi = bDesc0 + (uint64_t)((b.idx * 2 * B_HALF_BYTES) >> 4), dLo = dHi + (B_HALF_BYTES >> 4);
					const uint32_t acc = tmem + (w * ACC_BUFS + buf) * KPM;
					const uint32_t aHi = tmem + A_BASE_COL + sl * 64, aLo = aHi + 32;
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
					tcCommit(bEmptyBar + b.idx * 8);
				}
				++inChunk;
				if (inChunk == F || ls + 1 == s.len) {
					if (leader) tcCommit(accFullBar + buf * 8);
					++gc;
					inChunk = 0;
				}
				__syncwarp();
			}
		}
	} else if (warp < 4 + 4 * SPLIT_WGS) {
		// ===== splitters: V tile -> TF32 hi/lo -> TMEM A slot.  Tile steps go round robin over the warpgroups,
		// so each has SPLIT_WGS tile times for its chain of barrier probes, shared-memory loads, tcgen05.st and
		// wait::st (~250 cycles of latency alone).  They know nothing of tiles or chunks: only ring positions.
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::SPLIT_REGS));
		const unsigned wgi = (warp - 4) / 4;
		const unsigned row = (warp % 4) * 32 + lane;        // A row = TMEM lane owned by this thread
		const uint32_t laneBase = ((warp % 4) * 32) << 16;
		const float center = p.center;
		unsigned tEnd = 0;   // tile steps of this CTA: two per real stage
		{
			SegmentWalker walk(p, blockIdx.x);
			Segment s;
			while (walk.next(s)) tEnd += 2 * s.len;
		}
		float v[STAGE_K];
		RingPos vpos;   // this warpgroup's position in its V ring (SPLIT_WGS == 2: warpgroup wgi reads the ring of A tile wgi)
		auto loadTile = [&](unsigned) {
			const unsigned sv = 2 * vpos.idx + wgi;
			mbarWait(vFullBar + sv * 8, vpos.lap);
			vpos.advance(SVH);
			const uint32_t tile = smemAddr(vRing) + sv * V_STAGE_BYTES;
			if (V_COLS_ARE_ROWS) {
				// row `row` of the [128][32] tile; 16-byte chunk c lives at chunk c ^ (row & 7)
				const uint32_t base = tile + row * 128;
#pragma unroll
				for (int c = 0; c < 8; ++c)
					asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
					             : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
					             : "r"(base + ((c ^ (row & 7)) << 4))
					             : "memory");
			} else {
				const uint32_t base = tile + row * 4;
#pragma unroll
				for (int j = 0; j < STAGE_K; ++j) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[j]) : "r"(base + j * TILE_ROWS * 4) : "memory");
			}
			// Release the slot only after the loads have been PERFORMED.  Issue order is not enough: an LDS that is still
			// queued behind TMA writes and MMA operand reads when the arrive reaches the barrier lets the producer refill
			// the slot under it, and for W^T V the refill is fast (256-byte L2 promotion: every second stage of a column
			// tile is already in L2).  Seen as a few A rows of ~1e-3 of the tile stages carrying wrong data, differently on
			// every run, in W^T V only (tools/race_check.py, profiles/r01_notes.md finding 10).  The CTA-scope fence makes every thread wait for
			// its loads; then one arrival per warp (128 per-thread arrivals on one mbarrier serialise in the barrier unit
			// and delay the TMA completions that share it).
			asm volatile("fence.acq_rel.cta;" ::: "memory");
			__syncwarp();
			if (lane == 0) mbarArrive(vEmptyBar + sv * 8);
		};
		if (wgi < tEnd) loadTile(wgi);
		for (unsigned t = wgi; t < tEnd; t += SPLIT_WGS) {
			const unsigned sl = t & (SLOTS - 1);
			mbarWait(emptyBar + sl * 8, ((t >> SLOTS_SHIFT) & 1) ^ 1);
			tcFenceAfter();
			const uint32_t aSlot = tmem + laneBase + A_BASE_COL + sl * 64;
			// ptxas guards the source registers of a tcgen05.st with a read scoreboard, so hi/lo may be reused for the
			// second half without waiting for the first stores
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				uint32_t hi[16], lo[16];
#pragma unroll
				for (int e = 0; e < 16; ++e) splitValue(v[16 * h + e] - center, hi[e], lo[e]);
				tmemStore16(aSlot + 16 * h, hi);
				tmemStore16(aSlot + 32 + 16 * h, lo);
			}
			tmemWaitStore();
			tcFenceBefore();
			__syncwarp();
			if (lane == 0) mbarArrive(fullBar + sl * 8);
			if (t + SPLIT_WGS < tEnd) loadTile(t + SPLIT_WGS);   // after the publish: a late V tile must not hold the A slot back
		}
	} else {
		// ===== flushers: every chunk, add the tensor-core accumulators to the fp32 running sums; write the segment's
		// partial product at its end.  Flusher warpgroup f owns the A tiles f, f + FLUSH_WGS, ...
		asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R::FLUSH_REGS));
		constexpr int TILES_PER_FLUSHER = 2 / FLUSH_WGS;
		const unsigned fwg = (warp - 4 - 4 * SPLIT_WGS) / 4;
		const unsigned row = (warp % 4) * 32 + lane;
		const uint32_t laneBase = ((warp % 4) * 32) << 16;
		const unsigned kp = p.kp;
		float sum[TILES_PER_FLUSHER][KPM];
#pragma unroll
		for (int i = 0; i < TILES_PER_FLUSHER; ++i)
#pragma unroll
			for (int c = 0; c < KPM; ++c) sum[i][c] = 0.f;

		SegmentWalker walk(p, blockIdx.x);
		Segment s;
		unsigned gc = 0;
		while (walk.next(s)) {
			const unsigned nChunks = (s.len + F - 1) / F;
			for (unsigned c = 0; c < nChunks; ++c, ++gc) {
				const unsigned buf = gc % ACC_BUFS;
				mbarWait(accFullBar + buf * 8, (gc / ACC_BUFS) & 1);
				tcFenceAfter();
#pragma unroll
				for (int i = 0; i < TILES_PER_FLUSHER; ++i) {
					const unsigned w = fwg + i * FLUSH_WGS;
					const uint32_t acc = tmem + laneBase + (w * ACC_BUFS + buf) * KPM;
#pragma unroll
					for (int q = 0; q < KPM / 16; ++q) {
						if (q * 16 < (int)kp) {
							uint32_t r[16];
							tmemLoad16(acc + q * 16, r);
							tmemWaitLoad();
#pragma unroll
							for (int e = 0; e < 16; ++e) sum[i][q * 16 + e] += __uint_as_float(r[e]);
						}
					}
				}
				tcFenceBefore();
				__syncwarp();
				if (lane == 0) mbarArrive(accEmptyBar + buf * 8);
			}
			// ---- output of this segment's partial product
			float* out = p.out + (size_t)s.slot * p.slotStride;
#pragma unroll
			for (int i = 0; i < TILES_PER_FLUSHER; ++i) {
				const unsigned r = s.tile * PAIR_ROWS + (fwg + i * FLUSH_WGS) * TILE_ROWS + row;
				if (r < p.rowsA) {
					if (V_COLS_ARE_ROWS) {
						// column r of N: kp contiguous values.  (Row blocks over several ranks: the partials stay local and
						// fused::pushN sends their sum to the owners of the columns in whole 256-byte rows -- 16-byte stores
						// at a 256-byte stride straight from here cost 50 us over NVLink, profiles/r02_notes.md.)
						float4* dst = reinterpret_cast<float4*>(out + (size_t)r * p.ldOut);
#pragma unroll
						for (int c = 0; c < KPM / 4; ++c)
							if (c * 4 < (int)kp) dst[c] = make_float4(sum[i][4 * c], sum[i][4 * c + 1], sum[i][4 * c + 2], sum[i][4 * c + 3]);
					} else {
#pragma unroll
						for (int c = 0; c < KPM; ++c)
							if (c < (int)p.k) out[(size_t)c * p.ldOut + r] = sum[i][c];          // row r of N2: coalesced across the warp
					}
				}
#pragma unroll
				for (int c = 0; c < KPM; ++c) sum[i][c] = 0.f;
			}
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
	return 1024 + (size_t)SV * V_STAGE_BYTES + (size_t)Rings<KPM>::SB * 2 * KPM * STAGE_K * 4 + sizeof(Barriers);
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

// ---- rank-one correction of the mean-centred products --------------------------------------------------------
// stage 1: partial sums in double over a slice of the long dimension; stage 2: out[x] = scale * sum of the slices
__global__ void __launch_bounds__(256) column_sums_stage1(unsigned rows, unsigned cols, const float* __restrict__ A, size_t lda, unsigned chunk,
                                                         double* __restrict__ partial) {
	__shared__ double red[8];
	const unsigned c = blockIdx.x, s = blockIdx.y;
	const unsigned begin = s * chunk, end = min(rows, begin + chunk);
	const float* a = A + (size_t)c * lda;
	double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
	unsigned i = begin + threadIdx.x;
	for (; i + 768 < end; i += 1024) {
		a0 += (double)a[i];
		a1 += (double)a[i + 256];
		a2 += (double)a[i + 512];
		a3 += (double)a[i + 768];
	}
	for (; i < end; i += 256) a0 += (double)a[i];
	double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
	__syncthreads();
	if (threadIdx.x == 0) partial[(size_t)s * cols + c] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}
__global__ void __launch_bounds__(256) row_sums_stage1(unsigned rows, unsigned cols, const float* __restrict__ H, size_t ldh, unsigned chunk,
                                                      double* __restrict__ partial) {
	__shared__ double red[256];
	const unsigned rp = rows <= 32 ? 32 : rows <= 64 ? 64 : 128;   // threads per column
	const unsigned r = threadIdx.x % rp, part = threadIdx.x / rp, parts = 256 / rp;
	const unsigned begin = blockIdx.x * chunk, end = min(cols, begin + chunk);
	double acc = 0.0;
	if (r < rows) {
		double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;   // independent chains: the loads of a thread overlap
		unsigned j = begin + part;
		for (; j + 3 * parts < end; j += 4 * parts) {
			a0 += (double)H[(size_t)j * ldh + r];
			a1 += (double)H[(size_t)(j + parts) * ldh + r];
			a2 += (double)H[(size_t)(j + 2 * parts) * ldh + r];
			a3 += (double)H[(size_t)(j + 3 * parts) * ldh + r];
		}
		for (; j < end; j += parts) a0 += (double)H[(size_t)j * ldh + r];
		acc = (a0 + a1) + (a2 + a3);
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	if (part == 0 && r < rows) {
		for (unsigned q = 1; q < parts; ++q) acc += red[q * rp + r];
		partial[(size_t)blockIdx.x * rows + r] = acc;
	}
}
// one warp per output entry, lanes stride the slices (a single thread walking 128 slices is a chain of 128 L2 latencies)
__global__ void __launch_bounds__(256) sums_stage2(unsigned count, unsigned slices, const double* __restrict__ partial, float scale, float* __restrict__ out) {
	const unsigned x = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
	if (x >= count) return;
	double acc = 0.0;
	for (unsigned s = lane; s < slices; s += 32) acc += partial[(size_t)s * count + x];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) out[x] = (float)((double)scale * acc);
}
constexpr unsigned SUM_SLICES = 32, ROW_SUM_SLICES = 128;

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
void makeMap(unsigned char* out, const float* base, unsigned rows, unsigned cols, size_t ld, unsigned boxRows, unsigned boxCols, bool swizzle128,
             bool promote256 = true) {
	CUtensorMap map;
	const cuuint64_t dims[2] = {rows, cols};
	const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
	const cuuint32_t box[2] = {boxRows, boxCols};
	const cuuint32_t elem[2] = {1, 1};
	const CUresult r = encodeTiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem,
	                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
	                                 promote256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

// the shape-only part of a product's plan (no CUDA calls): tiles, chunks, grid and the partial products per 128-wide tile
void planShape(Product& prod, unsigned rowsA, unsigned kdim, unsigned kp, unsigned sms, std::vector<unsigned char>& counts) {
	prod.tiles = ceilDiv(rowsA, PAIR_ROWS);
	prod.stagesPerTile = ceilDiv(kdim, STAGE_K);
	// reduction chunks: the hi/lo copies of the small operand that one chunk reads should sit in L2 (~12 MB) while all CTAs
	// sweep that chunk; at least 64 stages per chunk and at most 8 chunks (every chunk adds partial products per tile)
	const double operandBytes = 2.0 * kdim * kp * 4.0;
	unsigned chunks = (unsigned)std::min(8.0, std::max(1.0, std::ceil(operandBytes / 12.0e6)));
	chunks = std::max(1u, std::min(chunks, prod.stagesPerTile / 64));
	// every chunk multiplies the partial products a tile receives (one per CTA that shares the tile), and the consumers
	// walk them one after the other: no more than ~24 per tile (8 GPUs: 5 tiles on 148 CTAs are 30 per chunk already)
	const unsigned perTile = ceilDiv(sms, prod.tiles) + 1;
	chunks = std::max(1u, std::min(chunks, 24u / perTile));
	if (const char* e = getenv("NMFGPU_TC_CHUNKS")) chunks = std::max(1u, std::min((unsigned)atoi(e), prod.stagesPerTile));   // tuning knob
	prod.chunkStages = ceilDiv(prod.stagesPerTile, chunks);
	prod.chunks = ceilDiv(prod.stagesPerTile, prod.chunkStages);
	const unsigned long long units = (unsigned long long)prod.tiles * prod.chunkStages;   // per chunk
	prod.grid = (unsigned)std::min<unsigned long long>(units, (unsigned long long)sms);
	// consumers index the counts by 128-wide tile (kernels.h), the stream-K tiles are 256 wide
	const unsigned tiles128 = ceilDiv(rowsA, TILE_ROWS);
	counts.assign(tiles128, 0);
	prod.maxSlots = 1;
	for (unsigned t = 0; t < prod.tiles; ++t) {
		const unsigned long long f = (unsigned long long)t * prod.chunkStages;
		const unsigned slots = prod.chunks * (ctaOfUnit(f + prod.chunkStages - 1, prod.grid, units) - ctaOfUnit(f, prod.grid, units) + 1);
		if (slots > 255) throw EngineError(ResultType::ErrorInvalidArgument, "more than 255 partial products per tile");
		for (unsigned h = 0; h < 2; ++h)
			if (2 * t + h < tiles128) counts[2 * t + h] = (unsigned char)slots;
		prod.maxSlots = std::max(prod.maxSlots, slots);
	}
}

void planProduct(Product& prod, unsigned rowsA, unsigned kdim, unsigned kp) {
	std::vector<unsigned char> counts;
	planShape(prod, rowsA, kdim, kp, (unsigned)smCount(), counts);
	const unsigned tiles128 = ceilDiv(rowsA, TILE_ROWS);
	if (prod.slotCount) pooledDeviceFree(prod.slotCount, prod.slotCountBytes);
	prod.slotCountBytes = roundUp(tiles128, 256);
	prod.slotCount = static_cast<unsigned char*>(pooledDeviceAlloc(prod.slotCountBytes));
	CUDA_CHECK(cudaMemcpy(prod.slotCount, counts.data(), tiles128, cudaMemcpyHostToDevice));
}

// the opt-in to > 48 KB of dynamic shared memory is per device: set when a plan is made on a device (never inside a
// stream capture), not cached process-wide
template <int KPM, bool VC>
void configureOne() {
	CUDA_CHECK(cudaFuncSetAttribute(tc_stream_gemm<KPM, VC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes<KPM>()));
}
void configureKernels() {
	configureOne<64, true>();
	configureOne<64, false>();
	configureOne<128, true>();
	configureOne<128, false>();
}

template <int KPM, bool VC>
void launch(const Plan& plan, const Product& prod, unsigned rowsA, float* out, size_t ldOut, size_t slotStride, cudaStream_t stream,
            const Gate* gate = nullptr) {
	const size_t smem = smemBytes<KPM>();
	KParams p;
	p.gateFlags = gate ? gate->flags : nullptr;
	p.gateEpoch = gate ? gate->epoch : nullptr;
	p.gateError = gate ? gate->error : nullptr;
	p.gateCount = gate ? gate->count : 0u;
	memcpy(&p.mapV, prod.mapV, 128);
	memcpy(&p.mapBhi, prod.mapBhi, 128);
	memcpy(&p.mapBlo, prod.mapBlo, 128);
	p.out = out;
	p.trace = plan.trace;
	p.ldOut = ldOut;
	p.slotStride = slotStride;
	p.units = (unsigned long long)prod.tiles * prod.chunkStages;   // per chunk
	p.chunks = prod.chunks;
	p.chunkStages = prod.chunkStages;
	p.rowsA = rowsA;
	p.k = plan.k;
	p.kp = plan.kp;
	p.tiles = prod.tiles;
	p.stagesPerTile = prod.stagesPerTile;
	p.flushStages = plan.flushStages;
	p.passes = plan.passes;
	p.grid = prod.grid;
	p.center = plan.center;
	p.prefetchStages = plan.prefetchStages;
	tc_stream_gemm<KPM, VC><<<prod.grid, Rings<KPM>::THREADS, smem, stream>>>(p);
	CUDA_CHECK(cudaGetLastError());
	if (getenv("NMFGPU_TC_DEBUG") != nullptr) {
		static unsigned rec[4 * 64];
		CUDA_CHECK(cudaStreamSynchronize(stream));
		CUDA_CHECK(cudaMemcpyFromSymbol(rec, g_waitTimeout, sizeof(rec)));
		if (rec[0] != 0) {
			errorf("tc_stream_gemm<%d,%d>: %u barrier waits timed out", KPM, (int)VC, rec[0]);
			for (unsigned i = 1; i <= rec[0] && i < 64; ++i)
				if (rec[4 * i + 3] == rec[7])   // the block of the first record
					errorf("  block %u warp %u: barrier +0x%x parity %u", rec[4 * i + 3], rec[4 * i + 2] / 32, rec[4 * i] & 0x3FF, rec[4 * i + 1]);
			memset(rec, 0, sizeof(rec));
			CUDA_CHECK(cudaMemcpyToSymbol(g_waitTimeout, rec, sizeof(rec)));
		}
	}
}

}  // namespace

unsigned enumerateSegments(unsigned rowsA, unsigned kdim, unsigned kp, unsigned sms, unsigned* segments, unsigned capacity, unsigned info[5],
                           unsigned char* slotsPerTile, unsigned tileCapacity) {
	Product prod;
	std::vector<unsigned char> counts;
	planShape(prod, rowsA, kdim, kp, sms, counts);
	info[0] = prod.tiles;
	info[1] = prod.stagesPerTile;
	info[2] = prod.chunks;
	info[3] = prod.chunkStages;
	info[4] = prod.grid;
	for (size_t t = 0; t < counts.size() && t < tileCapacity; ++t) slotsPerTile[t] = counts[t];
	unsigned count = 0;
	const unsigned long long units = (unsigned long long)prod.tiles * prod.chunkStages;
	for (unsigned cta = 0; cta < prod.grid; ++cta) {
		SegmentWalker walk(units, prod.grid, prod.stagesPerTile, prod.chunkStages, prod.chunks, cta);
		Segment s;
		while (walk.next(s)) {
			if (count < capacity) {
				unsigned* o = segments + 5 * (size_t)count;
				o[0] = cta; o[1] = s.tile; o[2] = s.stage0; o[3] = s.len; o[4] = s.slot;
			}
			++count;
		}
	}
	return count;
}

float meanOf(const float* V, unsigned m, unsigned n, size_t ldV, cudaStream_t stream) {
	double* partial = nullptr;
	partial = static_cast<double*>(pooledDeviceAlloc((size_t)n * sizeof(double)));
	column_sums_stage1<<<dim3(n, 1), 256, 0, stream>>>(m, n, V, ldV, m, partial);
	std::vector<double> host(n);
	const cudaError_t e = cudaMemcpyAsync(host.data(), partial, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream);
	const cudaError_t e2 = cudaStreamSynchronize(stream);
	pooledDeviceFree(partial, (size_t)n * sizeof(double));
	CUDA_CHECK(e);
	CUDA_CHECK(e2);
	double total = 0.0;
	for (double v : host) total += v;
	return (float)(total / ((double)m * (double)n));
}

void refreshCorrectionW(Plan& plan, const float* W, size_t ldW, cudaStream_t stream) {
	const unsigned chunk = ceilDiv(plan.m, SUM_SLICES);
	column_sums_stage1<<<dim3(plan.k, SUM_SLICES), 256, 0, stream>>>(plan.m, plan.k, W, ldW, chunk, plan.sumScratch);
	sums_stage2<<<ceilDiv(plan.k, 8), 256, 0, stream>>>(plan.k, SUM_SLICES, plan.sumScratch, plan.center, plan.corrN);
	CUDA_CHECK(cudaGetLastError());
}

void columnSums(Plan& plan, const float* W, unsigned rows, size_t ldW, float* out, cudaStream_t stream) {
	const unsigned chunk = ceilDiv(rows, SUM_SLICES);
	column_sums_stage1<<<dim3(plan.k, SUM_SLICES), 256, 0, stream>>>(rows, plan.k, W, ldW, chunk, plan.sumScratch);
	sums_stage2<<<ceilDiv(plan.k, 8), 256, 0, stream>>>(plan.k, SUM_SLICES, plan.sumScratch, 1.f, out);
	CUDA_CHECK(cudaGetLastError());
}

void rowSums(Plan& plan, const float* H, unsigned cols, size_t ldH, float* out, cudaStream_t stream) {
	const unsigned chunk = ceilDiv(cols, ROW_SUM_SLICES);
	row_sums_stage1<<<ROW_SUM_SLICES, 256, 0, stream>>>(plan.k, cols, H, ldH, chunk, plan.sumScratch);
	sums_stage2<<<ceilDiv(plan.k, 8), 256, 0, stream>>>(plan.k, ROW_SUM_SLICES, plan.sumScratch, 1.f, out);
	CUDA_CHECK(cudaGetLastError());
}

void refreshCorrectionH(Plan& plan, const float* H, size_t ldH, cudaStream_t stream) {
	const unsigned chunk = ceilDiv(plan.n, ROW_SUM_SLICES);
	row_sums_stage1<<<ROW_SUM_SLICES, 256, 0, stream>>>(plan.k, plan.n, H, ldH, chunk, plan.sumScratch);
	sums_stage2<<<ceilDiv(plan.k, 8), 256, 0, stream>>>(plan.k, ROW_SUM_SLICES, plan.sumScratch, plan.center, plan.corrP);
	CUDA_CHECK(cudaGetLastError());
}

Plan::~Plan() {
	if (corrN) pooledDeviceFree(corrN, 128 * sizeof(float));
	if (corrP) pooledDeviceFree(corrP, 128 * sizeof(float));
	if (sumScratch) pooledDeviceFree(sumScratch, (size_t)ROW_SUM_SLICES * 128 * sizeof(double));
	if (trace != nullptr) {
		// diagnostic timeline of CTA 0 of the LAST product launched: one line per stage with the clock64 stamps
		// worker{tile landed, split done, slot free, slot published}, MMA{operands ready, stage issued}
		std::vector<unsigned long long> host(TRACE_STAGES * TRACE_EVENTS);
		if (cudaMemcpy(host.data(), trace, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
			if (FILE* f = fopen(getenv("NMFGPU_TC_TRACE"), "w")) {
				unsigned long long t0 = ~0ull;
				for (unsigned long long v : host) if (v != 0 && v < t0) t0 = v;
				for (unsigned g = 0; g < TRACE_STAGES; ++g) {
					fprintf(f, "%u", g);
					for (unsigned e = 0; e < 14; ++e) fprintf(f, " %lld", host[g * TRACE_EVENTS + e] ? (long long)(host[g * TRACE_EVENTS + e] - t0) : -1ll);
					fprintf(f, "\n");
				}
				fclose(f);
			}
		}
		cudaFree(trace);
	}
	if (wtv.slotCount) pooledDeviceFree(wtv.slotCount, wtv.slotCountBytes);
	if (vht.slotCount) pooledDeviceFree(vht.slotCount, vht.slotCountBytes);
}

bool shapeSupported(unsigned m, unsigned n, unsigned k, size_t ldV, size_t ldW) {
	if (k == 0 || k > 128 || m == 0 || n == 0) return false;
	if (ldV % 4 != 0 || ldW % 4 != 0) return false;   // TMA: 16-byte global strides
	int dev = 0, major = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
	return major == 10;
}

void makePlan(Plan& plan, unsigned m, unsigned n, unsigned k, const float* V, size_t ldV, const float* Whi, const float* Wlo, size_t ldW,
              const float* HtHi, const float* HtLo, size_t ldHt, bool singlePass, float center, unsigned reduceLenWtV) {
	configureKernels();
	plan.m = m;
	plan.n = n;
	plan.k = k;
	plan.kp = (unsigned)roundUp(k, 16);
	plan.passes = singlePass ? 1 : 3;
	plan.prefetchStages = 0;
	if (const char* e = getenv("NMFGPU_TC_PREFETCH")) plan.prefetchStages = (unsigned)strtol(e, nullptr, 10);   // tuning knob
	if (const char* e = getenv("NMFGPU_TC_ABLATE")) plan.passes |= (unsigned)strtol(e, nullptr, 0) & 0xFF00;   // timing experiments only: results are garbage
	plan.flushStages = 16;   // 512 reduction elements per chunk. Centred data: 5e-8 relative error for any value >= 4; all-positive worst case 2.4e-7 per stage
	if (const char* e = getenv("NMFGPU_TC_FLUSH_STAGES")) {   // tuning knob: 0 = accumulate whole segments inside the tensor core
		const long v = strtol(e, nullptr, 10);
		plan.flushStages = v <= 0 ? 0x40000000u : (unsigned)v;
	}
	plan.center = center;
	if (plan.corrN == nullptr) {
		plan.corrN = static_cast<float*>(pooledDeviceAlloc(128 * sizeof(float)));
		plan.corrP = static_cast<float*>(pooledDeviceAlloc(128 * sizeof(float)));
		plan.sumScratch = static_cast<double*>(pooledDeviceAlloc((size_t)ROW_SUM_SLICES * 128 * sizeof(double)));
		CUDA_CHECK(cudaMemset(plan.corrN, 0, 128 * sizeof(float)));
		CUDA_CHECK(cudaMemset(plan.corrP, 0, 128 * sizeof(float)));
	}
	if (getenv("NMFGPU_TC_TRACE") != nullptr && plan.trace == nullptr) {
		CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&plan.trace), TRACE_STAGES * TRACE_EVENTS * sizeof(unsigned long long)));
		CUDA_CHECK(cudaMemset(plan.trace, 0, TRACE_STAGES * TRACE_EVENTS * sizeof(unsigned long long)));
	}
	// W^T V: A rows = columns of V, reduction over m
	// (reduceLenWtV > m: the work split of a longer reduction -- every rank of a row-block run uses the split of the
	// largest block so that the slot layout is the same everywhere; the rows beyond m are TMA out-of-bounds zeros)
	planProduct(plan.wtv, n, std::max(m, reduceLenWtV), plan.kp);
	makeMap(plan.wtv.mapV, V, m, n, ldV, STAGE_K, TILE_ROWS, true);
	makeMap(plan.wtv.mapBhi, Whi, m, k, ldW, STAGE_K, plan.kp, true);
	makeMap(plan.wtv.mapBlo, Wlo, m, k, ldW, STAGE_K, plan.kp, true);
	// V H^T: A rows = rows of V, reduction over n
	planProduct(plan.vht, m, n, plan.kp);
	makeMap(plan.vht.mapV, V, m, n, ldV, TILE_ROWS, STAGE_K, false, false);   // 512-byte rows: promotion only costs bandwidth (tools/tma_stream_bench)
	makeMap(plan.vht.mapBhi, HtHi, n, k, ldHt, STAGE_K, plan.kp, true);
	makeMap(plan.vht.mapBlo, HtLo, n, k, ldHt, STAGE_K, plan.kp, true);
}

void gemmWtV(const Plan& plan, float* Npart, size_t ldn, size_t slotStride, cudaStream_t stream) {
	if (plan.kp <= 64) launch<64, true>(plan, plan.wtv, plan.n, Npart, ldn, slotStride, stream);
	else launch<128, true>(plan, plan.wtv, plan.n, Npart, ldn, slotStride, stream);
}

const unsigned* timeoutCounter() {
	void* p = nullptr;
	CUDA_CHECK(cudaGetSymbolAddress(&p, g_waitTimeout));
	return static_cast<const unsigned*>(p);
}

void gemmVHt(const Plan& plan, float* Ppart, size_t ldp, size_t slotStride, cudaStream_t stream, const Gate* gate) {
	if (plan.kp <= 64) launch<64, false>(plan, plan.vht, plan.m, Ppart, ldp, slotStride, stream, gate);
	else launch<128, false>(plan, plan.vht, plan.m, Ppart, ldp, slotStride, stream, gate);
}

void splitTransposeH(unsigned k, unsigned n, const float* H, size_t ldh, float* hi, float* lo, size_t ldht, cudaStream_t stream) {
	const dim3 grid(ceilDiv(n, 32), ceilDiv(k, 32));
	split_transpose_kernel<<<grid, 256, 0, stream>>>(k, n, H, ldh, hi, lo, ldht);
	CUDA_CHECK(cudaGetLastError());
}

}  // namespace tc
}  // namespace b200
}  // namespace nmfgpu
