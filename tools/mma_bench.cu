// mma_bench.cu -- tcgen05.mma issue-rate microbenchmark (B200): cycles per kind::tf32 MMA as a function of
// N, the number of independent accumulators the MMAs rotate over, and the source of the A operand
// (shared memory descriptor or tensor memory).  Operand contents are irrelevant (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_bench tools/mma_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t descSw128(uint32_t addr) {
	uint64_t d = 0;
	d |= (uint64_t)((addr & 0x3FFFF) >> 4);
	d |= (uint64_t)1 << 16;
	d |= (uint64_t)(1024 >> 4) << 32;
	d |= (uint64_t)1 << 46;
	d |= (uint64_t)2 << 61;
	return d;
}
__device__ __forceinline__ bool electOne() {
	uint32_t pred;
	asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
	return pred != 0;
}
__device__ __forceinline__ void mmaTS(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
	             "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mmaSS(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
	             "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// one CTA: thread 0 issues `count` MMAs rotating over nAcc accumulators of N columns each
__global__ void __launch_bounds__(128, 1) bench(int N, int nAcc, int aTmem, int count, int kind16, long long* cyclesOut) {
	extern __shared__ __align__(1024) unsigned char smemRaw[];
	unsigned char* smem = (unsigned char*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar;
	__shared__ uint32_t tmemBase;
	for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&tmemBase)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;
	if (threadIdx.x < 32) {
		uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
		idesc |= kind16 ? ((1u << 7) | (1u << 10)) : ((2u << 7) | (2u << 10));   // bf16 : tf32
		const uint64_t aDesc = descSw128(smemAddr(smem));              // 128 rows x 128 B = 16 KB
		const uint64_t bDesc = descSw128(smemAddr(smem + 16384));      // up to 256 rows x 128 B = 32 KB
		const uint32_t aT = tmem + 448;                                // A operand columns (garbage)
		long long t0 = clock64();
		const uint32_t accMask = (uint32_t)nAcc - 1;   // nAcc is a power of two
		const bool leader = electOne();
#pragma unroll 4
		for (int i = 0; i < count; ++i) {
			const uint32_t d = tmem + (uint32_t)((i & accMask) * N);
			if (!leader) continue;
			if (kind16) {
				if (aTmem) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(aT), "l"(bDesc), "r"(idesc), "r"(1u) : "memory");
				else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(1u) : "memory");
			} else {
				if (aTmem) mmaTS(d, aT, bDesc, idesc, 1);
				else mmaSS(d, aDesc, bDesc, idesc, 1);
			}
		}
		if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(&bar)) : "memory");
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smemAddr(&bar)), "r"(0) : "memory");
		}
		long long t1 = clock64();
		if (blockIdx.x == 0 && threadIdx.x == 0) *cyclesOut = t1 - t0;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// The issue pattern of the real kernel: per "stage" one elected lane issues 12 MMAs (4 k-steps x 3 terms)
// straight-line, A from tensor memory, rotating over nAcc accumulators per term.
template <int NACC>
__global__ void __launch_bounds__(128, 1) benchBlock(int N, int stages, long long* cyclesOut) {
	extern __shared__ __align__(1024) unsigned char smemRaw[];
	unsigned char* smem = (unsigned char*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar;
	__shared__ uint32_t tmemBase;
	for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&tmemBase)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;
	if (threadIdx.x < 32) {
		const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24) | (2u << 7) | (2u << 10);
		const uint64_t bHi = descSw128(smemAddr(smem + 16384)), bLo = descSw128(smemAddr(smem + 16384 + 8192));
		const uint32_t aHi = tmem + 448, aLo = tmem + 480;
		const bool leader = electOne();
		long long t0 = clock64();
		for (int s = 0; s < stages; ++s) {
			if (leader) {
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					mmaTS(tmem + (0 % NACC) * N, aHi + q * 8, bHi + q * 2, idesc, 1);
					mmaTS(tmem + (1 % NACC) * N, aLo + q * 8, bHi + q * 2, idesc, 1);
					mmaTS(tmem + (2 % NACC) * N, aHi + q * 8, bLo + q * 2, idesc, 1);
				}
			}
			__syncwarp();
		}
		if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(&bar)) : "memory");
		uint32_t done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smemAddr(&bar)), "r"(0) : "memory");
		}
		long long t1 = clock64();
		if (blockIdx.x == 0 && threadIdx.x == 0) *cyclesOut = t1 - t0;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Do tcgen05.st / tcgen05.ld from other warps slow the MMA stream down (shared TMEM ports)?  Warp 0 issues the
// block12 pattern; warps 4-7 concurrently write `storeCols` columns (x16 stores) per MMA stage-equivalent.
__global__ void __launch_bounds__(256, 1) benchContend(int N, int stages, int storesPerIter, int withMma, long long* out) {
	extern __shared__ __align__(1024) unsigned char smemRaw[];
	unsigned char* smem = (unsigned char*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar;
	__shared__ uint32_t tmemBase;
	__shared__ volatile int stop;
	for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
	if (threadIdx.x == 0) {
		stop = 0;
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&tmemBase)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;
	const unsigned warp = threadIdx.x / 32;
	if (warp == 0) {
		const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24) | (2u << 7) | (2u << 10);
		const uint64_t bHi = descSw128(smemAddr(smem + 16384)), bLo = descSw128(smemAddr(smem + 16384 + 8192));
		const uint32_t aHi = tmem + 448, aLo = tmem + 480;
		const bool leader = electOne();
		long long t0 = clock64();
		if (withMma) {
			for (int s = 0; s < stages; ++s) {
				if (leader) {
#pragma unroll
					for (int q = 0; q < 4; ++q) {
						mmaTS(tmem, aHi + q * 8, bHi + q * 2, idesc, 1);
						mmaTS(tmem, aLo + q * 8, bHi + q * 2, idesc, 1);
						mmaTS(tmem, aHi + q * 8, bLo + q * 2, idesc, 1);
					}
				}
				__syncwarp();
			}
			if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(&bar)) : "memory");
			uint32_t done = 0;
			while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smemAddr(&bar)), "r"(0) : "memory");
		} else {
			while (clock64() - t0 < 400000) {}
		}
		long long t1 = clock64();
		if (threadIdx.x == 0) stop = 1;
		if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
	} else if (warp >= 4 && storesPerIter > 0) {
		// each warp writes its 32 lanes: columns 256.. of TMEM (away from the accumulator and the A operand columns)
		const uint32_t base = tmem + (((warp % 4) * 32) << 16) + 256;
		uint32_t r[16];
#pragma unroll
		for (int i = 0; i < 16; ++i) r[i] = i + threadIdx.x;
		long long iters = 0;
		long long t0 = clock64();
		while (!stop) {
			for (int q = 0; q < storesPerIter; ++q)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + (q % 4) * 16),
				             "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			++iters;
		}
		long long t1 = clock64();
		if (blockIdx.x == 0 && threadIdx.x == 128) { out[1] = iters; out[2] = t1 - t0; }
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

void runContend(long long* d) {
	cudaFuncSetAttribute(benchContend, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
	printf("contention: warp 0 issues 512 stages of 12 MMAs (N=64, A in TMEM); warps 4-7 loop {k x tcgen05.st.x16; wait::st}\n");
	for (int withMma : {1, 0})
		for (int stores : {0, 1, 4, 8}) {
			if (!withMma && stores == 0) continue;
			cudaMemset(d, 0, 3 * sizeof(long long));
			benchContend<<<148, 256, 66 * 1024>>>(64, 512, stores, withMma, d);
			cudaError_t e = cudaDeviceSynchronize();
			if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
			long long h[3];
			cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
			printf("  mma %d  stores/iter %d : MMA cycles/stage %8.1f | store loop: %8.1f cycles per iteration (%.1f per x16 store, 4 warps = 8 KB each)\n", withMma, stores,
			       withMma ? (double)h[0] / 512 : 0.0, h[1] ? (double)h[2] / h[1] : 0.0, h[1] && stores ? (double)h[2] / h[1] / stores : 0.0);
		}
}

template <int NACC>
void runBlock(int N, long long* d) {
	const int stages = 512;
	cudaFuncSetAttribute(benchBlock<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
	benchBlock<NACC><<<148, 128, 66 * 1024>>>(N, stages, d);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
	long long c;
	cudaMemcpy(&c, d, sizeof(c), cudaMemcpyDeviceToHost);
	printf("block12 tf32 TS  N=%3d nAcc=%d  cycles/MMA %7.1f  cycles/stage %8.1f\n", N, NACC, (double)c / (stages * 12), (double)c / stages);
}

int main() {
	long long* d;
	cudaMalloc(&d, 4 * sizeof(long long));
	runContend(d);
	cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
	for (int N : {16, 32, 64, 128}) {
		runBlock<1>(N, d);
		runBlock<3>(N, d);
	}
	const int count = 4096;
	printf("kind  A     N   nAcc  grid  cycles/MMA   (ideal 128*N/256 = N/2)\n");
	for (int kind16 = 0; kind16 < 2; ++kind16)
		for (int aTmem = 0; aTmem < 2; ++aTmem)
			for (int N : {32, 64, 128, 256})
				for (int nAcc : {1, 2, 4})
					for (int grid : {1, 148}) {
						if (nAcc * N > 448) continue;
						bench<<<grid, 128, 66 * 1024>>>(N, nAcc, aTmem, count, kind16, d);
						cudaError_t e = cudaDeviceSynchronize();
						if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
						long long c;
						cudaMemcpy(&c, d, sizeof(c), cudaMemcpyDeviceToHost);
						printf("%s  %s  %3d  %d     %3d   %8.1f\n", kind16 ? "bf16" : "tf32", aTmem ? "tmem" : "smem", N, nAcc, grid, (double)c / count);
					}
	return 0;
}
