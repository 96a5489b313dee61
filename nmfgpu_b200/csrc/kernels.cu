// kernels.cu -- SIMT kernels of the NMF iteration (exact fp32 / fp64 arithmetic), sm_100a.
//
// See kernels.h for what each launcher replaces in the reference.  Design notes:
//   * every reduction has a fixed order (split partials + ordered sums, no float atomics), so a run
//     is bit-reproducible and parity failures are real failures, not noise;
//   * the elementwise MU rule is `x * num / (den + eps)` -- multiply first, then divide -- exactly as
//     source/nmf/KernelMultiplyDivide.cu:42 of the reference;
//   * the fused H/W update kernels keep the k x k Gram matrix in shared memory and one factor
//     row/column in registers, so each factor is read once and written once per update.
#include "kernels.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include <cub/cub.cuh>

#include "common.h"

namespace nmfgpu {
namespace b200 {
namespace kern {

namespace {

// ---------------------------------------------------------------------------------------------------
// tiled SIMT GEMM:  C[M x N] = sum_K A(a,kk) * B(kk,b)
//   A_KC: A(a,kk) = A[a*lda + kk]  else A[kk*lda + a]
//   B_KC: B(kk,b) = B[b*ldb + kk]  else B[kk*ldb + b]
// C is column-major (a fastest).  grid = (M tiles, N tiles, splits).
// ---------------------------------------------------------------------------------------------------
template <typename T, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) gemm_simt(unsigned M, unsigned N, unsigned K, const T* __restrict__ A, size_t lda,
                                                const T* __restrict__ B, size_t ldb, T* __restrict__ C, size_t ldc, unsigned kChunk,
                                                size_t splitStride) {
	constexpr int BM = 64, BN = 64, BK = 32, TM = 4, TN = 4, PAD = 4;
	__shared__ T As[BK][BM + PAD];
	__shared__ T Bs[BK][BN + PAD];
	const unsigned tid = threadIdx.x;
	const unsigned tx = tid % 16, ty = tid / 16;  // tx -> M micro tile, ty -> N micro tile
	const unsigned a0 = blockIdx.x * BM, b0 = blockIdx.y * BN;
	const unsigned kBegin = blockIdx.z * kChunk;
	const unsigned kEnd = min(K, kBegin + kChunk);

	T acc[TM][TN];
#pragma unroll
	for (int i = 0; i < TM; ++i)
#pragma unroll
		for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

	for (unsigned k0 = kBegin; k0 < kEnd; k0 += BK) {
#pragma unroll
		for (int e = 0; e < (BM * BK) / 256; ++e) {
			const unsigned idx = tid + e * 256;
			unsigned a, kk;
			if (A_KC) { kk = idx % BK; a = idx / BK; } else { a = idx % BM; kk = idx / BM; }
			const unsigned ga = a0 + a, gk = k0 + kk;
			T v = T(0);
			if (ga < M && gk < kEnd) v = A_KC ? A[(size_t)ga * lda + gk] : A[(size_t)gk * lda + ga];
			As[kk][a] = v;
		}
#pragma unroll
		for (int e = 0; e < (BN * BK) / 256; ++e) {
			const unsigned idx = tid + e * 256;
			unsigned b, kk;
			if (B_KC) { kk = idx % BK; b = idx / BK; } else { b = idx % BN; kk = idx / BN; }
			const unsigned gb = b0 + b, gk = k0 + kk;
			T v = T(0);
			if (gb < N && gk < kEnd) v = B_KC ? B[(size_t)gb * ldb + gk] : B[(size_t)gk * ldb + gb];
			Bs[kk][b] = v;
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < BK; ++kk) {
			T av[TM], bv[TN];
#pragma unroll
			for (int i = 0; i < TM; ++i) av[i] = As[kk][tx * TM + i];
#pragma unroll
			for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][ty * TN + j];
#pragma unroll
			for (int i = 0; i < TM; ++i)
#pragma unroll
				for (int j = 0; j < TN; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
		}
		__syncthreads();
	}
	T* Cs = C + (size_t)blockIdx.z * splitStride;
#pragma unroll
	for (int j = 0; j < TN; ++j) {
		const unsigned gb = b0 + ty * TN + j;
		if (gb >= N) continue;
#pragma unroll
		for (int i = 0; i < TM; ++i) {
			const unsigned ga = a0 + tx * TM + i;
			if (ga < M) Cs[(size_t)gb * ldc + ga] = acc[i][j];
		}
	}
}

template <typename T>
__global__ void sum_splits_kernel(unsigned rows, unsigned cols, const T* __restrict__ src, size_t ldsrc, unsigned splits,
                                  size_t splitStride, T* __restrict__ dst, size_t lddst, const unsigned char* __restrict__ tileSlots,
                                  bool tilesAlongRows, const T* __restrict__ corr) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= rows) return;
	// the columns stride over gridDim.y (at most 65535 blocks: k x n matrices have more columns than that)
	for (unsigned c = blockIdx.y; c < cols; c += gridDim.y) {
		const unsigned count = tileSlots != nullptr ? tileSlots[(tilesAlongRows ? r : c) >> 7] : splits;
		// corr: rank-one term of a mean-centred tensor-core product (tc_gemm.h), indexed along the short (rank) dimension
		T s = corr != nullptr ? corr[tilesAlongRows ? c : r] : T(0);
		for (unsigned sp = 0; sp < count; ++sp) s += src[sp * splitStride + (size_t)c * ldsrc + r];
		dst[(size_t)c * lddst + r] = s;
	}
}

// many partials of a small matrix (the split-K Gram products): 32 elements x 8 partial groups per block
template <typename T>
__global__ void __launch_bounds__(256) sum_splits_small_kernel(unsigned rows, unsigned cols, const T* __restrict__ src, size_t ldsrc, unsigned splits,
                                                              size_t splitStride, T* __restrict__ dst, size_t lddst) {
	__shared__ T red[8][33];
	const unsigned e = blockIdx.x * 32 + threadIdx.x % 32, g = threadIdx.x / 32;
	const unsigned r = e % rows, c = e / rows;
	T acc = T(0);
	if (c < cols)
		for (unsigned sp = g; sp < splits; sp += 8) acc += src[sp * splitStride + (size_t)c * ldsrc + r];
	red[g][threadIdx.x % 32] = acc;
	__syncthreads();
	if (g == 0 && c < cols) {
		const unsigned x = threadIdx.x;
		dst[(size_t)c * lddst + r] = ((red[0][x] + red[1][x]) + (red[2][x] + red[3][x])) + ((red[4][x] + red[5][x]) + (red[6][x] + red[7][x]));
	}
}

__device__ __forceinline__ float tf32_hi(float x) {
	uint32_t u;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
	return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------------
// H update, generic rank: one thread per element, 32 columns per block, double buffered Hin -> Hout.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) update_h_generic(unsigned k, unsigned n, const T* __restrict__ G, const T* __restrict__ Hin,
                                                       T* __restrict__ Hout, size_t ldh, const T* __restrict__ Npart, size_t ldn,
                                                       unsigned splits, size_t splitStride, T eps, T* __restrict__ tracePartials,
                                                       float* __restrict__ HtHi, float* __restrict__ HtLo, size_t ldht,
                                                       const unsigned char* __restrict__ tileSlots, const T* __restrict__ corr) {
	extern __shared__ unsigned char smem_raw[];
	T* hcol = reinterpret_cast<T*>(smem_raw);  // [8 warps][k]
	const unsigned lane = threadIdx.x % 32, warp = threadIdx.x / 32;
	T* mine = hcol + (size_t)warp * k;
	for (unsigned cc = 0; cc < 4; ++cc) {
		const unsigned j = blockIdx.x * 32 + warp * 4 + cc;
		if (j >= n) break;  // warp-uniform
		if (tileSlots != nullptr) splits = tileSlots[j >> 7];
		for (unsigned t = lane; t < k; t += 32) mine[t] = Hin[(size_t)j * ldh + t];
		__syncwarp();
		T tr = T(0);
		for (unsigned r = lane; r < k; r += 32) {
			T d = T(0);
			for (unsigned t = 0; t < k; ++t) d = fma(G[(size_t)t * k + r], mine[t], d);
			T num = corr != nullptr ? corr[r] : T(0);
			for (unsigned s = 0; s < splits; ++s) num += Npart[s * splitStride + (size_t)j * ldn + r];
			const T hn = mine[r] * num / (d + eps);
			Hout[(size_t)j * ldh + r] = hn;
			tr = fma(hn, num, tr);
			if (HtHi != nullptr) {
				const float hi = tf32_hi((float)hn);
				HtHi[(size_t)r * ldht + j] = hi;
				HtLo[(size_t)r * ldht + j] = (float)hn - hi;
			}
		}
		if (tracePartials != nullptr) {
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
			if (lane == 0) tracePartials[j] = tr;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------
// H update, rank <= KP (fp32): a 64-column panel per block.  D = G H is a register-tiled product out of
// shared memory (thread = KP/16 rows x 4 columns), the multiplicative update, the residual trace term and
// the TF32 hi/lo copy of the new H^T are its epilogue.
// ---------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) update_h_reg(unsigned k, unsigned n, const float* __restrict__ G, const float* __restrict__ Hin,
                                                   float* __restrict__ Hout, size_t ldh, const float* __restrict__ Npart, size_t ldn,
                                                   unsigned splits, size_t splitStride, float eps, float* __restrict__ tracePartials,
                                                   float* __restrict__ HtHi, float* __restrict__ HtLo, size_t ldht,
                                                   const unsigned char* __restrict__ tileSlots, const float* __restrict__ corr,
                                                   float* __restrict__ rowSumPartials) {
	constexpr int COLS = 64, RPT = KP / 16, LDJ = COLS + 4;
	const unsigned j0 = blockIdx.x * COLS;
	if (tileSlots != nullptr) splits = tileSlots[j0 >> 7];
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* Gs = reinterpret_cast<float*>(smem_raw);  // [KP t][KP r]: Gs[t*KP + r] = G[r + t*k]
	float* Hs = Gs + KP * KP;                         // [KP t][LDJ]:  Hs[t*LDJ + j] = H[t, j0 + j]
	const unsigned tid = threadIdx.x;
	for (unsigned idx = tid; idx < KP * KP; idx += 256) {
		const unsigned r = idx % KP, t = idx / KP;
		Gs[idx] = (r < k && t < k) ? G[(size_t)t * k + r] : 0.f;
	}
	for (unsigned idx = tid; idx < COLS * KP; idx += 256) {
		const unsigned t = idx % KP, j = idx / KP;
		Hs[t * LDJ + j] = (j0 + j < n && t < k) ? Hin[(size_t)(j0 + j) * ldh + t] : 0.f;
	}
	const unsigned rx = tid % 16, jx = tid / 16;      // rows rx*RPT.., columns jx*4..
	// the numerators (sum of the partial products + centring term) are fetched before the product: their latency hides behind it
	// (slot loop outside, entries inside: all loads of a slot are in flight together -- with the loops the other way round
	// every entry pays its own chain of `splits` memory latencies)
	float numv[RPT][4];
#pragma unroll
	for (int q = 0; q < 4; ++q)
#pragma unroll
		for (int i = 0; i < RPT; ++i) numv[i][q] = (corr != nullptr && rx * RPT + i < k) ? corr[rx * RPT + i] : 0.f;
	for (unsigned sl = 0; sl < splits; ++sl) {
		const float* slot = Npart + sl * splitStride;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const unsigned j = j0 + jx * 4 + q;
			if (j < n) {
				if (RPT % 4 == 0 && rx * RPT + RPT <= k) {
#pragma unroll
					for (int i = 0; i < RPT; i += 4) {
						const float4 x = *reinterpret_cast<const float4*>(slot + (size_t)j * ldn + rx * RPT + i);
						numv[i][q] += x.x; numv[i + 1][q] += x.y; numv[i + 2][q] += x.z; numv[i + 3][q] += x.w;
					}
				} else {
#pragma unroll
					for (int i = 0; i < RPT; ++i)
						if (rx * RPT + i < k) numv[i][q] += slot[(size_t)j * ldn + rx * RPT + i];
				}
			}
		}
	}
#pragma unroll
	for (int q = 0; q < 4; ++q)
		if (j0 + jx * 4 + q >= n) {
#pragma unroll
			for (int i = 0; i < RPT; ++i) numv[i][q] = 0.f;
		}
	__syncthreads();
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
	// epilogue: the new values replace the old ones in the H tile (every thread overwrites exactly the entries it read),
	// then H, the transposed TF32 split and the row sums leave shared memory with coalesced stores
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const unsigned j = j0 + jx * 4 + q;
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
		if (tracePartials != nullptr && rx == 0 && j < n) tracePartials[j] = tr;
	}
	__syncthreads();
	for (unsigned idx = tid; idx < COLS * KP; idx += 256) {
		const unsigned t = idx % KP, j = idx / KP;
		if (j0 + j < n && t < k) Hout[(size_t)(j0 + j) * ldh + t] = Hs[t * LDJ + j];
	}
	if (HtHi != nullptr) {
		for (unsigned idx = tid; idx < COLS * KP; idx += 256) {
			const unsigned j = idx % COLS, r = idx / COLS;
			if (r < k && j0 + j < n) {
				const float v = Hs[r * LDJ + j];
				const float hi = tf32_hi(v);
				HtHi[(size_t)r * ldht + j0 + j] = hi;
				HtLo[(size_t)r * ldht + j0 + j] = v - hi;
			}
		}
	}
	if (rowSumPartials != nullptr) {   // row sums of the new H over this block's columns (centring term of V H^T, tc_gemm.h)
		const unsigned lane = tid % 32;
		for (unsigned r = tid / 32; r < k; r += 8) {
			float sum = Hs[r * LDJ + lane] + Hs[r * LDJ + 32 + lane];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
			if (lane == 0) rowSumPartials[(size_t)blockIdx.x * k + r] = sum;
		}
	}
}

// out[r] = scale * sum over the blocks of partials[b * count + r]: one warp per entry, lanes stride the blocks, fixed order
__global__ void __launch_bounds__(256) finish_partial_sums_kernel(unsigned count, unsigned blocks, const float* __restrict__ partials, float scale,
                                                                 float* __restrict__ out) {
	const unsigned r = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
	if (r >= count) return;
	double acc = 0.0;
	for (unsigned b = lane; b < blocks; b += 32) acc += (double)partials[(size_t)b * count + r];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) out[r] = (float)((double)scale * acc);
}

template <typename T>
__global__ void clamp_kernel(unsigned rows, unsigned cols, T* A, size_t lda) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= rows) return;
	for (unsigned c = blockIdx.y; c < cols; c += gridDim.y) {
		const T v = A[(size_t)c * lda + r];
		A[(size_t)c * lda + r] = v > T(0) ? v : T(0);
	}
}

template <typename T>
__global__ void abs_kernel(unsigned rows, unsigned cols, T* A, size_t lda) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= rows) return;
	for (unsigned c = blockIdx.y; c < cols; c += gridDim.y) {
		const T v = A[(size_t)c * lda + r];
		A[(size_t)c * lda + r] = v < T(0) ? -v : v;
	}
}

// ---------------------------------------------------------------------------------------------------
// W update, generic rank: one thread per row, Win re-read through L1.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) update_w_generic(unsigned m, unsigned k, const T* __restrict__ B, const T* __restrict__ Win,
                                                       T* __restrict__ Wout, size_t ldw, const T* __restrict__ Ppart, size_t ldp,
                                                       unsigned splits, size_t splitStride, T eps, T* __restrict__ colSqPartials,
                                                       const unsigned char* __restrict__ tileSlots, const T* __restrict__ corr) {
	__shared__ T warpSq[4];
	if (tileSlots != nullptr) splits = tileSlots[blockIdx.x];
	const unsigned i = blockIdx.x * 128 + threadIdx.x;
	const bool valid = i < m;
	const unsigned lane = threadIdx.x % 32, warp = threadIdx.x / 32;
	for (unsigned c = 0; c < k; ++c) {
		T wn = T(0);
		if (valid) {
			T d = T(0);
			for (unsigned t = 0; t < k; ++t) d = fma(Win[(size_t)t * ldw + i], B[(size_t)c * k + t], d);
			T p = corr != nullptr ? corr[c] : T(0);
			for (unsigned s = 0; s < splits; ++s) p += Ppart[s * splitStride + (size_t)c * ldp + i];
			wn = Win[(size_t)c * ldw + i] * p / (d + eps);
			Wout[(size_t)c * ldw + i] = wn;
		}
		T sq = wn * wn;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
		if (lane == 0) warpSq[warp] = sq;
		__syncthreads();
		if (threadIdx.x == 0) colSqPartials[(size_t)blockIdx.x * k + c] = (warpSq[0] + warpSq[1]) + (warpSq[2] + warpSq[3]);
		__syncthreads();
	}
}

// W update, rank <= KP (fp32): a 128-row panel per block.  D = W (H H^T) is a register-tiled product out of
// shared memory (thread = 4 rows x KP/8 columns); the multiplicative update and the per-block column sums of
// squares (first half of the normalisation, MU.h:247) are its epilogue.
template <int KP>
__global__ void __launch_bounds__(256) update_w_reg(unsigned m, unsigned k, const float* __restrict__ B, const float* __restrict__ Win,
                                                   float* __restrict__ Wout, size_t ldw, const float* __restrict__ Ppart, size_t ldp,
                                                   unsigned splits, size_t splitStride, float eps, float* __restrict__ colSqPartials,
                                                   const unsigned char* __restrict__ tileSlots, const float* __restrict__ corr,
                                                   float* __restrict__ colSumPartials) {
	constexpr int ROWS = 128, CPT = KP / 8;
	if (tileSlots != nullptr) splits = tileSlots[blockIdx.x];
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* Ws = reinterpret_cast<float*>(smem_raw);  // [KP t][ROWS]: Ws[t*ROWS + r] = W[i0 + r, t]
	float* Bs = Ws + KP * ROWS;                       // [KP t][KP c]: Bs[t*KP + c] = B[t + c*k]
	const unsigned tid = threadIdx.x;
	const unsigned i0 = blockIdx.x * ROWS;
	for (unsigned idx = tid; idx < KP * ROWS; idx += 256) {
		const unsigned r = idx % ROWS, t = idx / ROWS;
		Ws[idx] = (i0 + r < m && t < k) ? Win[(size_t)t * ldw + i0 + r] : 0.f;
	}
	for (unsigned idx = tid; idx < KP * KP; idx += 256) {
		const unsigned c = idx % KP, t = idx / KP;
		Bs[idx] = (c < k && t < k) ? B[(size_t)c * k + t] : 0.f;
	}
	__syncthreads();
	const unsigned tx = tid % 32, ty = tid / 32;      // rows tx*4.., columns ty*CPT..
	float acc[4][CPT];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
#pragma unroll 4
	for (int t = 0; t < KP; ++t) {
		const float4 a = *reinterpret_cast<const float4*>(Ws + t * ROWS + tx * 4);
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
	const unsigned r0 = i0 + tx * 4;
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		const unsigned c = ty * CPT + j;
		float s2 = 0.f, s1 = 0.f;
		if (c < k) {   // warp-uniform
			const float base = corr != nullptr ? corr[c] : 0.f;
			float p[4] = {base, base, base, base};
			if (r0 + 3 < m) {
				for (unsigned sl = 0; sl < splits; ++sl) {
					const float4 x = *reinterpret_cast<const float4*>(Ppart + sl * splitStride + (size_t)c * ldp + r0);
					p[0] += x.x; p[1] += x.y; p[2] += x.z; p[3] += x.w;
				}
			} else {
				for (unsigned sl = 0; sl < splits; ++sl)
					for (int i = 0; i < 4; ++i)
						if (r0 + i < m) p[i] += Ppart[sl * splitStride + (size_t)c * ldp + r0 + i];
			}
			const float4 w = *reinterpret_cast<const float4*>(Ws + c * ROWS + tx * 4);
			float wn[4];
			wn[0] = w.x * p[0] / (acc[0][j] + eps);
			wn[1] = w.y * p[1] / (acc[1][j] + eps);
			wn[2] = w.z * p[2] / (acc[2][j] + eps);
			wn[3] = w.w * p[3] / (acc[3][j] + eps);
			if (r0 + 3 < m) {
				*reinterpret_cast<float4*>(Wout + (size_t)c * ldw + r0) = make_float4(wn[0], wn[1], wn[2], wn[3]);
			} else {
				for (int i = 0; i < 4; ++i) {
					if (r0 + i < m) Wout[(size_t)c * ldw + r0 + i] = wn[i];
					else wn[i] = 0.f;
				}
			}
			s2 = (wn[0] * wn[0] + wn[1] * wn[1]) + (wn[2] * wn[2] + wn[3] * wn[3]);
			s1 = (wn[0] + wn[1]) + (wn[2] + wn[3]);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			s2 += __shfl_xor_sync(0xffffffffu, s2, o);
			s1 += __shfl_xor_sync(0xffffffffu, s1, o);
		}
		if (tx == 0 && c < k) {
			colSqPartials[(size_t)blockIdx.x * k + c] = s2;
			if (colSumPartials != nullptr) colSumPartials[(size_t)blockIdx.x * k + c] = s1;
		}
	}
}

template <typename T>
__global__ void __launch_bounds__(128) column_squares_kernel(unsigned m, unsigned k, const T* __restrict__ W, size_t ldw,
                                                            T* __restrict__ colSqPartials) {
	__shared__ T warpSq[4];
	const unsigned i = blockIdx.x * 128 + threadIdx.x;
	const unsigned lane = threadIdx.x % 32, warp = threadIdx.x / 32;
	for (unsigned c = 0; c < k; ++c) {
		const T v = i < m ? W[(size_t)c * ldw + i] : T(0);
		T sq = v * v;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
		if (lane == 0) warpSq[warp] = sq;
		__syncthreads();
		if (threadIdx.x == 0) colSqPartials[(size_t)blockIdx.x * k + c] = (warpSq[0] + warpSq[1]) + (warpSq[2] + warpSq[3]);
		__syncthreads();
	}
}

// one block per column; fixed-order tree over the row-block partials
template <typename T>
__global__ void __launch_bounds__(256) finish_norms_kernel(unsigned k, unsigned blocks, const T* __restrict__ partials, T* __restrict__ colSq,
                                                           const T* __restrict__ sumPartials, float center, float* __restrict__ corrOut) {
	__shared__ T red[256];
	__shared__ double redSum[256];
	const unsigned c = blockIdx.x;
	T s = T(0);
	double s1 = 0.0;
	for (unsigned b = threadIdx.x; b < blocks; b += 256) {
		s += partials[(size_t)b * k + c];
		if (sumPartials != nullptr) s1 += (double)sumPartials[(size_t)b * k + c];
	}
	red[threadIdx.x] = s;
	redSum[threadIdx.x] = s1;
	__syncthreads();
	for (unsigned o = 128; o > 0; o >>= 1) {
		if (threadIdx.x < o) {
			red[threadIdx.x] += red[threadIdx.x + o];
			redSum[threadIdx.x] += redSum[threadIdx.x + o];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		colSq[c] = red[0];
		// centring term of W^T V for the unit-column matrix that scale_columns writes next (tc_gemm.h)
		if (corrOut != nullptr) corrOut[c] = (float)((double)center * (red[0] > T(0) ? redSum[0] / sqrt((double)red[0]) : redSum[0]));
	}
}

template <typename T>
__global__ void scale_columns_kernel(unsigned m, unsigned k, T* __restrict__ W, size_t ldw, const T* __restrict__ colSq,
                                     float* __restrict__ Whi, float* __restrict__ Wlo) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (i >= m || c >= k) return;
	const T s = colSq[c];
	T v = W[(size_t)c * ldw + i];
	if (s > T(0)) {
		v = v / sqrt(s);
		W[(size_t)c * ldw + i] = v;
	}
	if (Whi != nullptr) {
		const float hi = tf32_hi((float)v);
		Whi[(size_t)c * ldw + i] = hi;
		Wlo[(size_t)c * ldw + i] = (float)v - hi;
	}
}

__global__ void split_tf32_kernel(unsigned rows, unsigned cols, const float* __restrict__ X, size_t ldx, float* __restrict__ hi,
                                  float* __restrict__ lo, size_t ldo) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (i >= rows || c >= cols) return;
	const float v = X[(size_t)c * ldx + i];
	const float h = tf32_hi(v);
	hi[(size_t)c * ldo + i] = h;
	lo[(size_t)c * ldo + i] = v - h;
}

// one warp per column: partial[j] = sum_i A[i,j]*B[i,j]
template <typename T>
__global__ void __launch_bounds__(256) column_dots_kernel(unsigned rows, unsigned cols, const T* __restrict__ A, size_t lda,
                                                         const T* __restrict__ B, size_t ldb, T* __restrict__ partial) {
	const unsigned j = blockIdx.x * 8 + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (j >= cols) return;
	const T* a = A + (size_t)j * lda;
	const T* b = B + (size_t)j * ldb;
	T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
	unsigned i = lane;
	for (; i + 96 < rows; i += 128) {
		s0 = fma(a[i], b[i], s0);
		s1 = fma(a[i + 32], b[i + 32], s1);
		s2 = fma(a[i + 64], b[i + 64], s2);
		s3 = fma(a[i + 96], b[i + 96], s3);
	}
	for (; i < rows; i += 32) s0 = fma(a[i], b[i], s0);
	T s = (s0 + s1) + (s2 + s3);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if (lane == 0) partial[j] = s;
}

template <typename T>
__global__ void trace_kk_kernel(unsigned k, const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ partial) {
	const unsigned d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= k) return;
	T s = T(0);
	for (unsigned i = 0; i < k; ++i) s = fma(A[(size_t)i * k + d], B[(size_t)d * k + i], s);
	partial[d] = s;
}

template <typename T>
__global__ void add_constraint_kernel(unsigned k, T* G, T offdiag, T diag) {
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned c = blockIdx.y;
	if (r < k && c < k) G[(size_t)c * k + r] += (r == c) ? diag : offdiag;
}

// X[i,c] = (1-theta) W[i,c] + (theta/k) * sum_t W[i,t]
template <typename T>
__global__ void __launch_bounds__(128) smooth_right_kernel(unsigned m, unsigned k, const T* __restrict__ W, size_t ldw, T* __restrict__ X,
                                                          size_t ldx, T theta) {
	const unsigned i = blockIdx.x * 128 + threadIdx.x;
	if (i >= m) return;
	T rs = T(0);
	for (unsigned t = 0; t < k; ++t) rs += W[(size_t)t * ldw + i];
	const T off = theta / T(k);
	for (unsigned c = 0; c < k; ++c) X[(size_t)c * ldx + i] = (T(1) - theta) * W[(size_t)c * ldw + i] + off * rs;
}

// Y[r,j] = (1-theta) H[r,j] + (theta/k) * sum_t H[t,j]; one warp per column
template <typename T>
__global__ void __launch_bounds__(256) smooth_left_kernel(unsigned k, unsigned n, const T* __restrict__ H, size_t ldh, T* __restrict__ Y,
                                                         size_t ldy, T theta) {
	const unsigned j = blockIdx.x * 8 + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (j >= n) return;
	T cs = T(0);
	for (unsigned t = lane; t < k; t += 32) cs += H[(size_t)j * ldh + t];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
	const T off = theta / T(k);
	for (unsigned r = lane; r < k; r += 32) Y[(size_t)j * ldy + r] = (T(1) - theta) * H[(size_t)j * ldh + r] + off * cs;
}

// ---------------------------------------------------------------------------------------------------
// k x k Householder QR in one block (shared memory), then independent solves.
// factor layout: [k*k] R (upper) + Householder vectors (below the diagonal, unit leading element
// implied) followed by [k] tau.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) qr_factor_kernel(unsigned k, const T* __restrict__ G, T* __restrict__ factor) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* a = reinterpret_cast<T*>(smem_raw);  // [k][k] column-major
	T* tau = a + (size_t)k * k;
	__shared__ T s_norm, s_tau, s_scale, s_beta;
	const unsigned tid = threadIdx.x, threads = blockDim.x;
	for (unsigned idx = tid; idx < k * k; idx += threads) a[idx] = G[idx];
	__syncthreads();
	const unsigned lane = tid % 32, warp = tid / 32, warps = threads / 32;
	for (unsigned j = 0; j < k; ++j) {
		T* col = a + (size_t)j * k;
		if (warp == 0) {   // Householder vector of column j: one warp, shuffle reduction
			T part = T(0);
			for (unsigned i = j + lane; i < k; i += 32) part += col[i] * col[i];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
			if (lane == 0) {
				const T norm = sqrt(part);
				s_norm = norm;
				if (norm == T(0)) {
					s_tau = T(0);
					s_scale = T(0);
					s_beta = T(0);
				} else {
					const T alpha = col[j];
					const T beta = alpha >= T(0) ? -norm : norm;
					s_tau = (beta - alpha) / beta;
					s_scale = T(1) / (alpha - beta);
					s_beta = beta;
				}
				tau[j] = s_tau;
			}
		}
		__syncthreads();
		if (s_norm != T(0)) {
			for (unsigned i = j + 1 + tid; i < k; i += threads) col[i] *= s_scale;
			__syncthreads();
			if (tid == 0) col[j] = s_beta;
			// apply the reflector (v_j = 1, v_i = col[i]) to the trailing columns: one warp per column, lanes along the
			// rows (contiguous in shared memory), so every warp works and no access is bank-conflicted
			const T tj = s_tau;
			for (unsigned c = j + 1 + warp; c < k; c += warps) {
				T* cc = a + (size_t)c * k;
				T part = T(0);
				for (unsigned i = j + lane; i < k; i += 32) part = fma(i == j ? T(1) : col[i], cc[i], part);
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
				const T dot = part * tj;
				for (unsigned i = j + lane; i < k; i += 32) cc[i] -= dot * (i == j ? T(1) : col[i]);
			}
		}
		__syncthreads();
	}
	for (unsigned idx = tid; idx < k * k + k; idx += threads) factor[idx] = a[idx];
}

// M = R^-1 Q^T (k <= 128): one WARP per column of the identity, the vector spread over the lanes (entry i in lane i % 32,
// register i / 32); the factor in shared memory.  (One thread per right-hand side walking the Householder vectors,
// qr_solve_clamp_kernel, needs 0.98 ms for the 128 columns of the identity in fp64; this needs microseconds.)
template <typename T>
__global__ void __launch_bounds__(1024) qr_invert_kernel(unsigned k, const T* __restrict__ factor, T* __restrict__ M) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* f = reinterpret_cast<T*>(smem_raw);   // [k*k] R + Householder vectors, [k] tau
	for (unsigned idx = threadIdx.x; idx < k * k + k; idx += blockDim.x) f[idx] = factor[idx];
	__syncthreads();
	const T* tau = f + (size_t)k * k;
	const unsigned lane = threadIdx.x % 32, warp = threadIdx.x / 32;
	const unsigned c = blockIdx.x * (blockDim.x / 32) + warp;
	if (c >= k) return;
	T x[4];
#pragma unroll
	for (int i = 0; i < 4; ++i) x[i] = (lane + 32 * i == c) ? T(1) : T(0);
	for (unsigned j = 0; j < k; ++j) {   // x <- Q^T x: reflector j has v_j = 1, v_i = f[j*k + i] for i > j
		const T tj = tau[j];
		if (tj == T(0)) continue;
		const T* col = f + (size_t)j * k;
		T v[4], part = T(0);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const unsigned r = lane + 32 * i;
			v[i] = r < k ? (r > j ? col[r] : (r == j ? T(1) : T(0))) : T(0);
			part = fma(v[i], x[i], part);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
		const T dot = part * tj;
#pragma unroll
		for (int i = 0; i < 4; ++i) x[i] -= dot * v[i];
	}
	for (int j = (int)k - 1; j >= 0; --j) {   // back substitution, column oriented: x_j /= R_jj, then x_i -= R_ij x_j for i < j
		const T* col = f + (size_t)j * k;
		T mine = T(0);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			if (i == j / 32) mine = x[i];
		const T xj = __shfl_sync(0xffffffffu, mine, j % 32) / col[j];
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int r = (int)lane + 32 * i;
			if (r == j) x[i] = xj;
			else if (r < j) x[i] -= col[r] * xj;
		}
	}
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const unsigned r = lane + 32 * i;
		if (r < k) M[(size_t)c * k + r] = x[i];
	}
}

// one thread per right-hand side; the vector lives in shared memory (stride k+1, conflict free)
template <typename T>
__global__ void __launch_bounds__(64) qr_solve_clamp_kernel(unsigned k, const T* __restrict__ factor, T* __restrict__ R, size_t ldr,
                                                           unsigned nrhs, bool transposed, bool clamp) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* xs = reinterpret_cast<T*>(smem_raw);  // [64][k+1]
	const unsigned tid = threadIdx.x;
	const unsigned base = blockIdx.x * 64;
	const unsigned ld = k + 1;
	const T* tau = factor + (size_t)k * k;
	// stage: coalesced along the contiguous direction of R
	if (transposed) {  // R is nrhs x k: element (q, c) at R[q + c*ldr]
		for (unsigned c = 0; c < k; ++c) {
			const unsigned q = base + tid;
			xs[tid * ld + c] = q < nrhs ? R[(size_t)c * ldr + q] : T(0);
		}
	} else {  // R is k x nrhs: element (c, q) at R[c + q*ldr]
		for (unsigned idx = tid; idx < 64 * k; idx += 64) {
			const unsigned c = idx % k, qq = idx / k;
			const unsigned q = base + qq;
			xs[qq * ld + c] = q < nrhs ? R[(size_t)q * ldr + c] : T(0);
		}
	}
	__syncthreads();
	T* x = xs + tid * ld;
	for (unsigned j = 0; j < k; ++j) {  // x <- Q^T x
		const T tj = tau[j];
		if (tj == T(0)) continue;
		const T* col = factor + (size_t)j * k;
		T dot = x[j];
		for (unsigned i = j + 1; i < k; ++i) dot = fma(col[i], x[i], dot);
		dot *= tj;
		x[j] -= dot;
		for (unsigned i = j + 1; i < k; ++i) x[i] -= dot * col[i];
	}
	for (int j = (int)k - 1; j >= 0; --j) {  // back substitution
		T s = x[j];
		for (unsigned c = j + 1; c < k; ++c) s -= factor[(size_t)c * k + j] * x[c];
		x[j] = s / factor[(size_t)j * k + j];
	}
	__syncthreads();
	if (transposed) {
		for (unsigned c = 0; c < k; ++c) {
			const unsigned q = base + tid;
			if (q < nrhs) {
				const T v = xs[tid * ld + c];
				R[(size_t)c * ldr + q] = (v > T(0) || !clamp) ? v : T(0);
			}
		}
	} else {
		for (unsigned idx = tid; idx < 64 * k; idx += 64) {
			const unsigned c = idx % k, qq = idx / k;
			const unsigned q = base + qq;
			if (q < nrhs) {
				const T v = xs[qq * ld + c];
				R[(size_t)q * ldr + c] = (v > T(0) || !clamp) ? v : T(0);
			}
		}
	}
}

#define launchCheck() CUDA_CHECK(cudaGetLastError())   // a macro: the message names the launch site

// opt-in to more than 48 KB of dynamic shared memory (never inside a stream capture).  The attribute belongs to the
// (kernel, device) pair, so the cache is keyed by the kernel's ADDRESS and the device ordinal -- keyed by the pointer TYPE,
// instantiations with the same signature (apply_left_clamp<128> / apply_right_clamp<128>) shared one entry and the second
// one was launched without its opt-in ("invalid argument").
template <typename K>
void allowSmem(K kernel, size_t bytes) {
	if (bytes <= 48 * 1024) return;
	static std::mutex lock;
	static std::map<std::pair<const void*, int>, size_t> allowed;
	int dev = 0;
	CUDA_CHECK(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> guard(lock);
	size_t& mine = allowed[{reinterpret_cast<const void*>(kernel), dev}];
	if (bytes <= mine) return;
	CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	mine = bytes;
}

// ---------------------------------------------------------------------------------------------------
// Least-squares updates with the explicit k x k inverse M = R^-1 Q^T of the (regularised) Gram matrix:
//   X (k x n) <- max(0, M X)          and          X (m x k) <- max(0, X M^T)
// as register-tiled products (the same tiling as the multiplicative updates).  One thread per right-hand
// side walking the Householder vectors (qr_solve_clamp_kernel) costs 2-4 ms at cfg 4; this costs ~0.1 ms.
// ---------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) apply_left_clamp(unsigned k, unsigned n, const float* __restrict__ M, float* __restrict__ X, size_t ldx) {
	constexpr int COLS = 64, RPT = KP / 16, LDJ = COLS + 4;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* Ms = reinterpret_cast<float*>(smem_raw);  // [KP t][KP r]: Ms[t*KP + r] = M[r + t*k]
	float* Xs = Ms + KP * KP;                         // [KP t][LDJ]
	const unsigned tid = threadIdx.x, j0 = blockIdx.x * COLS;
	for (unsigned idx = tid; idx < KP * KP; idx += 256) {
		const unsigned r = idx % KP, t = idx / KP;
		Ms[idx] = (r < k && t < k) ? M[(size_t)t * k + r] : 0.f;
	}
	for (unsigned idx = tid; idx < COLS * KP; idx += 256) {
		const unsigned t = idx % KP, j = idx / KP;
		Xs[t * LDJ + j] = (j0 + j < n && t < k) ? X[(size_t)(j0 + j) * ldx + t] : 0.f;
	}
	__syncthreads();
	const unsigned rx = tid % 16, jx = tid / 16;
	float acc[RPT][4];
#pragma unroll
	for (int i = 0; i < RPT; ++i)
#pragma unroll
		for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
#pragma unroll 8
	for (int t = 0; t < KP; ++t) {
		float a[RPT];
#pragma unroll
		for (int i = 0; i < RPT; ++i) a[i] = Ms[t * KP + rx * RPT + i];
		const float4 b = *reinterpret_cast<const float4*>(Xs + t * LDJ + jx * 4);
#pragma unroll
		for (int i = 0; i < RPT; ++i) {
			acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
			acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
			acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
			acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
		}
	}
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const unsigned j = j0 + jx * 4 + q;
#pragma unroll
		for (int i = 0; i < RPT; ++i) {
			const unsigned r = rx * RPT + i;
			if (j < n && r < k) X[(size_t)j * ldx + r] = fmaxf(acc[i][q], 0.f);
		}
	}
}

template <int KP>
__global__ void __launch_bounds__(256) apply_right_clamp(unsigned m, unsigned k, const float* __restrict__ M, float* __restrict__ X, size_t ldx) {
	constexpr int ROWS = 128, CPT = KP / 8;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* Xs = reinterpret_cast<float*>(smem_raw);  // [KP t][ROWS]
	float* Bs = Xs + KP * ROWS;                       // [KP t][KP c]: Bs[t*KP + c] = M[c + t*k]  (= M^T[t, c])
	const unsigned tid = threadIdx.x, i0 = blockIdx.x * ROWS;
	for (unsigned idx = tid; idx < KP * ROWS; idx += 256) {
		const unsigned r = idx % ROWS, t = idx / ROWS;
		Xs[idx] = (i0 + r < m && t < k) ? X[(size_t)t * ldx + i0 + r] : 0.f;
	}
	for (unsigned idx = tid; idx < KP * KP; idx += 256) {
		const unsigned c = idx % KP, t = idx / KP;
		Bs[idx] = (c < k && t < k) ? M[(size_t)t * k + c] : 0.f;
	}
	__syncthreads();
	const unsigned tx = tid % 32, ty = tid / 32;
	float acc[4][CPT];
#pragma unroll
	for (int i = 0; i < 4; ++i)
#pragma unroll
		for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
#pragma unroll 4
	for (int t = 0; t < KP; ++t) {
		const float4 a = *reinterpret_cast<const float4*>(Xs + t * ROWS + tx * 4);
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
	const unsigned r0 = i0 + tx * 4;
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		const unsigned c = ty * CPT + j;
		if (c >= k) continue;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			if (r0 + i < m) X[(size_t)c * ldx + r0 + i] = fmaxf(acc[i][j], 0.f);
	}
}

}  // namespace

// ===================================================================================================
// launchers
// ===================================================================================================

template <typename T>
void gemmTN(unsigned rows, unsigned ka, unsigned nb, const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc, unsigned splits,
            size_t splitStride, cudaStream_t stream) {
	if (splits == 0) splits = 1;
	unsigned chunk = (unsigned)roundUp(ceilDiv(rows, splits), 32);
	splits = ceilDiv(rows, chunk);
	dim3 grid(ceilDiv(ka, 64), ceilDiv(nb, 64), splits);
	gemm_simt<T, true, true><<<grid, 256, 0, stream>>>(ka, nb, rows, A, lda, B, ldb, C, ldc, chunk, splitStride);
	launchCheck();
}

template <typename T>
void gemmNT(unsigned ma, unsigned cols, unsigned kb, const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc, unsigned splits,
            size_t splitStride, cudaStream_t stream) {
	if (splits == 0) splits = 1;
	unsigned chunk = (unsigned)roundUp(ceilDiv(cols, splits), 32);
	splits = ceilDiv(cols, chunk);
	dim3 grid(ceilDiv(ma, 64), ceilDiv(kb, 64), splits);
	gemm_simt<T, false, false><<<grid, 256, 0, stream>>>(ma, kb, cols, A, lda, B, ldb, C, ldc, chunk, splitStride);
	launchCheck();
}

// number of split partials the two launchers above actually produce for a requested split count
unsigned effectiveSplits(unsigned reduceLen, unsigned splits) {
	if (splits == 0) splits = 1;
	unsigned chunk = (unsigned)roundUp(ceilDiv(reduceLen, splits), 32);
	return ceilDiv(reduceLen, chunk);
}

template <typename T>
void sumSplits(unsigned rows, unsigned cols, const T* src, size_t ldsrc, unsigned splits, size_t splitStride, T* dst, size_t lddst,
               cudaStream_t stream, const unsigned char* tileSlots, bool tilesAlongRows, const T* corr) {
	if (tileSlots == nullptr && corr == nullptr && splits > 8 && (size_t)rows * cols <= 65536) {
		sum_splits_small_kernel<T><<<ceilDiv(rows * cols, 32), 256, 0, stream>>>(rows, cols, src, ldsrc, splits, splitStride, dst, lddst);
		launchCheck();
		return;
	}
	dim3 grid(ceilDiv(rows, 128), std::min(cols, 65535u));
	sum_splits_kernel<T><<<grid, 128, 0, stream>>>(rows, cols, src, ldsrc, splits, splitStride, dst, lddst, tileSlots, tilesAlongRows, corr);
	launchCheck();
}

template <typename T>
static void updateHGeneric(unsigned k, unsigned n, const T* G, const T* Hin, T* Hout, size_t ldh, const T* Npart, size_t ldn,
                           unsigned splits, size_t splitStride, T eps, T* tracePartials, float* HtHi, float* HtLo, size_t ldht,
                           cudaStream_t stream, const unsigned char* tileSlots, const T* corr) {
	const size_t smem = (size_t)8 * k * sizeof(T);
	allowSmem(update_h_generic<T>, smem);
	update_h_generic<T><<<ceilDiv(n, 32), 256, smem, stream>>>(k, n, G, Hin, Hout, ldh, Npart, ldn, splits, splitStride, eps,
	                                                            tracePartials, HtHi, HtLo, ldht, tileSlots, corr);
	launchCheck();
}

template <int KP>
static void updateHReg(unsigned k, unsigned n, const float* G, const float* Hin, float* Hout, size_t ldh, const float* Npart, size_t ldn,
                       unsigned splits, size_t splitStride, float eps, float* tracePartials, float* HtHi, float* HtLo, size_t ldht,
                       cudaStream_t stream, const unsigned char* tileSlots, const float* corr, float* rowSumPartials) {
	const size_t smem = sizeof(float) * ((size_t)KP * KP + (size_t)KP * (64 + 4));
	allowSmem(update_h_reg<KP>, smem);
	update_h_reg<KP><<<ceilDiv(n, 64), 256, smem, stream>>>(k, n, G, Hin, Hout, ldh, Npart, ldn, splits, splitStride, eps, tracePartials,
	                                                          HtHi, HtLo, ldht, tileSlots, corr, rowSumPartials);
	launchCheck();
}

template <>
void updateH<float>(unsigned k, unsigned n, const float* G, const float* Hin, float* Hout, size_t ldh, const float* Npart, size_t ldn,
                    unsigned splits, size_t splitStride, float eps, float* tracePartials, float* HtHi, float* HtLo, size_t ldht,
                    cudaStream_t stream, const unsigned char* tileSlots, const float* corr, float* rowSumPartials) {
#define NMF_ARGS k, n, G, Hin, Hout, ldh, Npart, ldn, splits, splitStride, eps, tracePartials, HtHi, HtLo, ldht, stream, tileSlots, corr
	if (k <= 16) updateHReg<16>(NMF_ARGS, rowSumPartials);
	else if (k <= 32) updateHReg<32>(NMF_ARGS, rowSumPartials);
	else if (k <= 64) updateHReg<64>(NMF_ARGS, rowSumPartials);
	else if (k <= 128) updateHReg<128>(NMF_ARGS, rowSumPartials);
	else {
		if (rowSumPartials != nullptr) throw EngineError(ResultType::ErrorInvalidArgument, "row-sum partials need a rank <= 128");
		updateHGeneric<float>(NMF_ARGS);
	}
#undef NMF_ARGS
}

template <>
void updateH<double>(unsigned k, unsigned n, const double* G, const double* Hin, double* Hout, size_t ldh, const double* Npart,
                     size_t ldn, unsigned splits, size_t splitStride, double eps, double* tracePartials, float* HtHi, float* HtLo,
                     size_t ldht, cudaStream_t stream, const unsigned char* tileSlots, const double* corr, double* rowSumPartials) {
	if (rowSumPartials != nullptr) throw EngineError(ResultType::ErrorInvalidArgument, "row-sum partials are an fp32 feature");
	updateHGeneric<double>(k, n, G, Hin, Hout, ldh, Npart, ldn, splits, splitStride, eps, tracePartials, HtHi, HtLo, ldht, stream, tileSlots, corr);
}

template <typename T>
void clampNonNegative(unsigned rows, unsigned cols, T* A, size_t lda, cudaStream_t stream) {
	dim3 grid(ceilDiv(rows, 128), std::min(cols, 65535u));
	clamp_kernel<T><<<grid, 128, 0, stream>>>(rows, cols, A, lda);
	launchCheck();
}

template <typename T>
void absInPlace(unsigned rows, unsigned cols, T* A, size_t lda, cudaStream_t stream) {
	dim3 grid(ceilDiv(rows, 128), std::min(cols, 65535u));
	abs_kernel<T><<<grid, 128, 0, stream>>>(rows, cols, A, lda);
	launchCheck();
}

// partial[2 b] = sum |x|, partial[2 b + 1] = sum x^2 over the share of block b (fp64, fixed order): the two norms behind
// the Hoyer sparseness of a factor (ExecutionRecord::sparsityW / sparsityH)
template <typename T>
__global__ void __launch_bounds__(256) abs_square_sums_kernel(unsigned rows, unsigned cols, const T* __restrict__ X, size_t ld, double* __restrict__ partial) {
	__shared__ double redA[8], redQ[8];
	double a = 0.0, q = 0.0;
	const unsigned long long total = (unsigned long long)rows * cols;
	for (unsigned long long idx = (unsigned long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (unsigned long long)gridDim.x * 256) {
		const unsigned r = (unsigned)(idx % rows);
		const size_t c = (size_t)(idx / rows);
		const double v = (double)X[c * ld + r];
		a += fabs(v);
		q += v * v;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		a += __shfl_xor_sync(0xffffffffu, a, o);
		q += __shfl_xor_sync(0xffffffffu, q, o);
	}
	if (threadIdx.x % 32 == 0) {
		redA[threadIdx.x / 32] = a;
		redQ[threadIdx.x / 32] = q;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		partial[2 * blockIdx.x] = ((redA[0] + redA[1]) + (redA[2] + redA[3])) + ((redA[4] + redA[5]) + (redA[6] + redA[7]));
		partial[2 * blockIdx.x + 1] = ((redQ[0] + redQ[1]) + (redQ[2] + redQ[3])) + ((redQ[4] + redQ[5]) + (redQ[6] + redQ[7]));
	}
}

template <typename T>
void absSquareSums(unsigned rows, unsigned cols, const T* X, size_t ld, double* partial, unsigned blocks, cudaStream_t stream) {
	abs_square_sums_kernel<T><<<blocks, 256, 0, stream>>>(rows, cols, X, ld, partial);
	launchCheck();
}

template <int KP>
static unsigned updateWReg(unsigned m, unsigned k, const float* B, const float* Win, float* Wout, size_t ldw, const float* Ppart,
                           size_t ldp, unsigned splits, size_t splitStride, float eps, float* colSqPartials, cudaStream_t stream,
                           const unsigned char* tileSlots, const float* corr, float* colSumPartials) {
	const size_t smem = sizeof(float) * ((size_t)KP * KP + (size_t)KP * 128);
	allowSmem(update_w_reg<KP>, smem);
	const unsigned blocks = ceilDiv(m, 128);
	update_w_reg<KP><<<blocks, 256, smem, stream>>>(m, k, B, Win, Wout, ldw, Ppart, ldp, splits, splitStride, eps, colSqPartials, tileSlots, corr, colSumPartials);
	launchCheck();
	return blocks;
}

template <>
unsigned updateW<float>(unsigned m, unsigned k, const float* B, const float* Win, float* Wout, size_t ldw, const float* Ppart, size_t ldp,
                        unsigned splits, size_t splitStride, float eps, float* colSqPartials, cudaStream_t stream, const unsigned char* tileSlots, const float* corr,
                        float* colSumPartials) {
#define NMF_ARGS m, k, B, Win, Wout, ldw, Ppart, ldp, splits, splitStride, eps, colSqPartials, stream, tileSlots, corr, colSumPartials
	if (k <= 16) return updateWReg<16>(NMF_ARGS);
	if (k <= 32) return updateWReg<32>(NMF_ARGS);
	if (k <= 64) return updateWReg<64>(NMF_ARGS);
	if (k <= 128) return updateWReg<128>(NMF_ARGS);
#undef NMF_ARGS
	if (colSumPartials != nullptr) throw EngineError(ResultType::ErrorInvalidArgument, "column-sum partials need a rank <= 128");
	const unsigned blocks = ceilDiv(m, 128);
	update_w_generic<float><<<blocks, 128, 0, stream>>>(m, k, B, Win, Wout, ldw, Ppart, ldp, splits, splitStride, eps, colSqPartials, tileSlots, corr);
	launchCheck();
	return blocks;
}

template <>
unsigned updateW<double>(unsigned m, unsigned k, const double* B, const double* Win, double* Wout, size_t ldw, const double* Ppart,
                         size_t ldp, unsigned splits, size_t splitStride, double eps, double* colSqPartials, cudaStream_t stream,
                         const unsigned char* tileSlots, const double* corr, double* colSumPartials) {
	if (colSumPartials != nullptr) throw EngineError(ResultType::ErrorInvalidArgument, "column-sum partials are an fp32 feature");
	const unsigned blocks = ceilDiv(m, 128);
	update_w_generic<double><<<blocks, 128, 0, stream>>>(m, k, B, Win, Wout, ldw, Ppart, ldp, splits, splitStride, eps, colSqPartials, tileSlots, corr);
	launchCheck();
	return blocks;
}

template <typename T>
void finishColumnNorms(unsigned k, unsigned blocks, const T* colSqPartials, T* colSq, cudaStream_t stream, const T* colSumPartials, float center,
                       float* corrOut) {
	finish_norms_kernel<T><<<k, 256, 0, stream>>>(k, blocks, colSqPartials, colSq, colSumPartials, center, corrOut);
	launchCheck();
}

void finishPartialSums(unsigned count, unsigned blocks, const float* partials, float scale, float* out, cudaStream_t stream) {
	finish_partial_sums_kernel<<<ceilDiv(count, 8), 256, 0, stream>>>(count, blocks, partials, scale, out);
	launchCheck();
}

template <typename T>
unsigned columnSquares(unsigned m, unsigned k, const T* W, size_t ldw, T* colSqPartials, cudaStream_t stream) {
	const unsigned blocks = ceilDiv(m, 128);
	column_squares_kernel<T><<<blocks, 128, 0, stream>>>(m, k, W, ldw, colSqPartials);
	launchCheck();
	return blocks;
}

template <typename T>
void scaleColumns(unsigned m, unsigned k, T* W, size_t ldw, const T* colSq, float* Whi, float* Wlo, cudaStream_t stream) {
	dim3 grid(ceilDiv(m, 256), k);
	scale_columns_kernel<T><<<grid, 256, 0, stream>>>(m, k, W, ldw, colSq, Whi, Wlo);
	launchCheck();
}

template <typename T>
void columnDots(unsigned rows, unsigned cols, const T* A, size_t lda, const T* B, size_t ldb, T* partial, cudaStream_t stream) {
	column_dots_kernel<T><<<ceilDiv(cols, 8), 256, 0, stream>>>(rows, cols, A, lda, B, ldb, partial);
	launchCheck();
}

template <typename T>
void traceKK(unsigned k, const T* A, const T* B, T* partial, cudaStream_t stream) {
	trace_kk_kernel<T><<<ceilDiv(k, 128), 128, 0, stream>>>(k, A, B, partial);
	launchCheck();
}

template <typename T>
void addConstraint(unsigned k, T* G, T offdiag, T diag, cudaStream_t stream) {
	dim3 grid(ceilDiv(k, 128), k);
	add_constraint_kernel<T><<<grid, 128, 0, stream>>>(k, G, offdiag, diag);
	launchCheck();
}

template <typename T>
void smoothRight(unsigned m, unsigned k, const T* W, size_t ldw, T* X, size_t ldx, T theta, cudaStream_t stream) {
	smooth_right_kernel<T><<<ceilDiv(m, 128), 128, 0, stream>>>(m, k, W, ldw, X, ldx, theta);
	launchCheck();
}

template <typename T>
void smoothLeft(unsigned k, unsigned n, const T* H, size_t ldh, T* Y, size_t ldy, T theta, cudaStream_t stream) {
	smooth_left_kernel<T><<<ceilDiv(n, 8), 256, 0, stream>>>(k, n, H, ldh, Y, ldy, theta);
	launchCheck();
}

template <typename T>
static void qrFactorPlain(unsigned k, const T* G, T* factor, cudaStream_t stream) {
	const size_t smem = sizeof(T) * ((size_t)k * k + k);
	allowSmem(qr_factor_kernel<T>, smem);
	qr_factor_kernel<T><<<1, k > 32 ? 1024 : 256, smem, stream>>>(k, G, factor);   // a warp per trailing column: 32 warps at k = 128
	launchCheck();
}

template <typename T>
static void qrInvert(unsigned k, const T* factor, T* M, cudaStream_t stream) {
	const size_t smem = sizeof(T) * ((size_t)k * k + k);
	allowSmem(qr_invert_kernel<T>, smem);
	const unsigned warps = k > 64 ? 32 : 16;
	qr_invert_kernel<T><<<ceilDiv(k, warps), warps * 32, smem, stream>>>(k, factor, M);
	launchCheck();
}

// dst[i] = (To)src[i]; identity != 0: dst = the k x k identity instead (count = k * k)
template <typename From, typename To>
__global__ void __launch_bounds__(256) convert_kernel(unsigned count, const From* __restrict__ src, To* __restrict__ dst, unsigned identityK) {
	const unsigned i = blockIdx.x * 256 + threadIdx.x;
	if (i >= count) return;
	if (identityK != 0) dst[i] = (i % identityK == i / identityK) ? To(1) : To(0);
	else dst[i] = (To)src[i];
}

template <typename T>
static void qrSolveGeneric(unsigned k, const T* factor, T* R, size_t ldr, unsigned nrhs, bool transposed, bool clamp, cudaStream_t stream) {
	const size_t smem = sizeof(T) * 64 * ((size_t)k + 1);
	allowSmem(qr_solve_clamp_kernel<T>, smem);
	qr_solve_clamp_kernel<T><<<ceilDiv(nrhs, 64), 64, smem, stream>>>(k, factor, R, ldr, nrhs, transposed, clamp);
	launchCheck();
}

template <int KP>
static void applyInverse(unsigned k, const float* M, float* R, size_t ldr, unsigned nrhs, bool transposed, cudaStream_t stream) {
	if (transposed) {
		const size_t smem = sizeof(float) * ((size_t)KP * KP + (size_t)KP * 128);
		allowSmem(apply_right_clamp<KP>, smem);
		apply_right_clamp<KP><<<ceilDiv(nrhs, 128), 256, smem, stream>>>(nrhs, k, M, R, ldr);
	} else {
		const size_t smem = sizeof(float) * ((size_t)KP * KP + (size_t)KP * (64 + 4));
		allowSmem(apply_left_clamp<KP>, smem);
		apply_left_clamp<KP><<<ceilDiv(nrhs, 64), 256, smem, stream>>>(k, nrhs, M, R, ldr);
	}
	launchCheck();
}

enum class LsSolve { InverseFp64, InverseFp32, Qr };
// NMFGPU_LS_SOLVE (study knob behind tools/ls_stability.py): "qr" applies Q^T and back-substitutes per right-hand side
// (what the reference's ormqr + trsm do, Matrix.h:587-618); "inverse32" forms M = R^-1 Q^T in fp32; default: M formed in fp64
static LsSolve lsSolveMode() {
	static const LsSolve mode = [] {
		const char* e = getenv("NMFGPU_LS_SOLVE");
		if (e != nullptr && strcmp(e, "qr") == 0) return LsSolve::Qr;
		if (e != nullptr && strcmp(e, "inverse32") == 0) return LsSolve::InverseFp32;
		return LsSolve::InverseFp64;
	}();
	return mode;
}

template <>
void qrFactor<double>(unsigned k, const double* G, double* factor, cudaStream_t stream, double*, double*) {
	qrFactorPlain<double>(k, G, factor, stream);
}

// fp32: the factorisation the per-right-hand-side solve walks, and -- when `inverse` is given -- the explicit inverse
// M = R^-1 Q^T for the tiled products of qrSolveClamp.  M is FORMED in fp64 (G converted, factorised and applied to the
// identity in double; k x k work, one block) and rounded once: formed in fp32 its error, cond(G) * eps * |M|, put ACLS 6-9x
// and ALS 2-4x further from the fp64 oracle than the reference's QR solve (profiles/r02_ls_stability.txt).
template <>
void qrFactor<float>(unsigned k, const float* G, float* factor, cudaStream_t stream, float* inverse, double* work) {
	const LsSolve mode = lsSolveMode();
	if (inverse == nullptr || k > 128 || mode == LsSolve::Qr) {   // the factor itself is what qrSolveClamp will walk
		qrFactorPlain<float>(k, G, factor, stream);
		return;
	}
	const unsigned kk = k * k;
	if (work == nullptr || mode == LsSolve::InverseFp32) {
		qrFactorPlain<float>(k, G, factor, stream);
		qrInvert<float>(k, factor, inverse, stream);
		return;
	}
	double* Gd = work;                 // [k*k]
	double* Fd = work + kk;            // [k*k + k]
	double* Md = work + 2 * (size_t)kk + k;   // [k*k]
	convert_kernel<float, double><<<ceilDiv(kk, 256), 256, 0, stream>>>(kk, G, Gd, 0);
	launchCheck();
	qrFactorPlain<double>(k, Gd, Fd, stream);
	qrInvert<double>(k, Fd, Md, stream);
	convert_kernel<double, float><<<ceilDiv(kk, 256), 256, 0, stream>>>(kk, Md, inverse, 0);
	launchCheck();
}

// `inverse`: M as left by qrFactor<float> (nullptr: solve per right-hand side)
template <>
void qrSolveClamp<float>(unsigned k, const float* factor, float* R, size_t ldr, unsigned nrhs, bool transposed, cudaStream_t stream, float* inverse) {
	if (inverse == nullptr || k > 128 || lsSolveMode() == LsSolve::Qr) {
		qrSolveGeneric<float>(k, factor, R, ldr, nrhs, transposed, true, stream);
		return;
	}
	if (k <= 16) applyInverse<16>(k, inverse, R, ldr, nrhs, transposed, stream);
	else if (k <= 32) applyInverse<32>(k, inverse, R, ldr, nrhs, transposed, stream);
	else if (k <= 64) applyInverse<64>(k, inverse, R, ldr, nrhs, transposed, stream);
	else applyInverse<128>(k, inverse, R, ldr, nrhs, transposed, stream);
}

template <>
void qrSolveClamp<double>(unsigned k, const double* factor, double* R, size_t ldr, unsigned nrhs, bool transposed, cudaStream_t stream, double*) {
	qrSolveGeneric<double>(k, factor, R, ldr, nrhs, transposed, true, stream);
}

void splitTf32(unsigned rows, unsigned cols, const float* X, size_t ldx, float* hi, float* lo, size_t ldo, cudaStream_t stream) {
	dim3 grid(ceilDiv(rows, 256), cols);
	split_tf32_kernel<<<grid, 256, 0, stream>>>(rows, cols, X, ldx, hi, lo, ldo);
	launchCheck();
}

// ascending radix sort of `count` values (the per-column residual terms before their host-side combine,
// FrobeniusResolver.cpp:31-50: a std::sort of 10 000 floats costs 0.2 ms per residual evaluation on the host)
template <typename T>
size_t sortAscending(const T* in, T* out, unsigned count, void* temp, size_t tempBytes, cudaStream_t stream) {
	size_t needed = 0;
	CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, needed, in, out, (int)count, 0, (int)sizeof(T) * 8, stream));
	if (temp == nullptr || tempBytes < needed) return needed;
	CUDA_CHECK(cub::DeviceRadixSort::SortKeys(temp, tempBytes, in, out, (int)count, 0, (int)sizeof(T) * 8, stream));
	return needed;
}

#define NMF_INSTANTIATE(T)                                                                                                             \
	template void gemmTN<T>(unsigned, unsigned, unsigned, const T*, size_t, const T*, size_t, T*, size_t, unsigned, size_t, cudaStream_t); \
	template void gemmNT<T>(unsigned, unsigned, unsigned, const T*, size_t, const T*, size_t, T*, size_t, unsigned, size_t, cudaStream_t); \
	template void sumSplits<T>(unsigned, unsigned, const T*, size_t, unsigned, size_t, T*, size_t, cudaStream_t, const unsigned char*, bool, const T*); \
	template void clampNonNegative<T>(unsigned, unsigned, T*, size_t, cudaStream_t);                                                      \
	template void absInPlace<T>(unsigned, unsigned, T*, size_t, cudaStream_t);                                                            \
	template void absSquareSums<T>(unsigned, unsigned, const T*, size_t, double*, unsigned, cudaStream_t);                                \
	template size_t sortAscending<T>(const T*, T*, unsigned, void*, size_t, cudaStream_t);                                                \
	template void finishColumnNorms<T>(unsigned, unsigned, const T*, T*, cudaStream_t, const T*, float, float*);                                                   \
	template unsigned columnSquares<T>(unsigned, unsigned, const T*, size_t, T*, cudaStream_t);                                           \
	template void scaleColumns<T>(unsigned, unsigned, T*, size_t, const T*, float*, float*, cudaStream_t);                                \
	template void columnDots<T>(unsigned, unsigned, const T*, size_t, const T*, size_t, T*, cudaStream_t);                                \
	template void traceKK<T>(unsigned, const T*, const T*, T*, cudaStream_t);                                                             \
	template void addConstraint<T>(unsigned, T*, T, T, cudaStream_t);                                                                     \
	template void smoothRight<T>(unsigned, unsigned, const T*, size_t, T*, size_t, T, cudaStream_t);                                      \
	template void smoothLeft<T>(unsigned, unsigned, const T*, size_t, T*, size_t, T, cudaStream_t);                                       \

NMF_INSTANTIATE(float)
NMF_INSTANTIATE(double)
#undef NMF_INSTANTIATE

}  // namespace kern
}  // namespace b200
}  // namespace nmfgpu
