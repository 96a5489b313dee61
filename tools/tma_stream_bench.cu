// tma_stream_bench.cu -- how fast can TMA stream the 100000 x 10000 fp32 matrix V (column-major) into shared
// memory with the box shapes and traversal orders the NMF products use?  No math: a consumer warp just
// releases each slot.  Answers "what is the HBM roofline for THIS access pattern" on a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tma_stream_bench tools/tma_stream_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	while (!done)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

struct Params {
	alignas(64) CUtensorMap map;
	unsigned tiles, stagesPerTile, boxBytes, depth, mode, tileStep, stageStep;   // mode 0: coords (stage, tile); 1: (tile, stage)
	unsigned long long units;
	int evictFirst;
	alignas(64) CUtensorMap mapB;   // optional second stream: L2-resident small operand, two boxes of bBoxBytes per stage
	unsigned bDepth, bBoxBytes;      // bDepth == 0: no B stream
};

__global__ void __launch_bounds__(128, 1) stream(const __grid_constant__ Params p) {
	extern __shared__ unsigned char smemRaw[];
	unsigned char* smem = (unsigned char*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t full[16], empty[16], bfull[16], bempty[16];
	if (threadIdx.x == 0) {
		for (unsigned i = 0; i < p.depth; ++i) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&full[i])), "r"(1));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&empty[i])), "r"(1));
		}
		for (unsigned i = 0; i < p.bDepth; ++i) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&bfull[i])), "r"(1));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&bempty[i])), "r"(1));
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	unsigned char* bsmem = smem + (size_t)p.depth * p.boxBytes;
	const unsigned long long u0 = p.units * blockIdx.x / gridDim.x, u1 = p.units * (blockIdx.x + 1) / gridDim.x;
	if (threadIdx.x == 0) {
		unsigned idx = 0, lap = 0;
		for (unsigned long long u = u0; u < u1; ++u) {
			const unsigned tile = (unsigned)(u / p.stagesPerTile), stage = (unsigned)(u % p.stagesPerTile);
			mbarWait(smemAddr(&empty[idx]), lap ^ 1);
			const uint32_t bar = smemAddr(&full[idx]);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(p.boxBytes) : "memory");
			const int c0 = p.mode == 0 ? (int)(stage * p.stageStep) : (int)(tile * p.tileStep);
			const int c1 = p.mode == 0 ? (int)(tile * p.tileStep) : (int)(stage * p.stageStep);
			const uint64_t policy = p.evictFirst ? 0x12F0000000000000ull : 0x1000000000000000ull;
			asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
			                 smemAddr(smem + (size_t)idx * p.boxBytes)),
			             "l"(reinterpret_cast<uint64_t>(&p.map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
			             : "memory");
			if (++idx == p.depth) { idx = 0; lap ^= 1; }
		}
	} else if (threadIdx.x == 64 && p.bDepth) {
		unsigned idx = 0, lap = 0;
		for (unsigned long long u = u0; u < u1; ++u) {
			const unsigned stage = (unsigned)(u % p.stagesPerTile);
			mbarWait(smemAddr(&bempty[idx]), lap ^ 1);
			const uint32_t bar = smemAddr(&bfull[idx]);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * p.bBoxBytes) : "memory");
			for (int h = 0; h < 2; ++h)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
				                 smemAddr(bsmem + (size_t)(2 * idx + h) * p.bBoxBytes)),
				             "l"(reinterpret_cast<uint64_t>(&p.mapB)), "r"(bar), "r"((int)(stage * 32)), "r"(h * 64), "l"(0x14F0000000000000ull)
				             : "memory");
			if (++idx == p.bDepth) { idx = 0; lap ^= 1; }
		}
	} else if (threadIdx.x == 96 && p.bDepth) {
		unsigned idx = 0, lap = 0;
		for (unsigned long long u = u0; u < u1; ++u) {
			mbarWait(smemAddr(&bfull[idx]), lap);
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(&bempty[idx])) : "memory");
			if (++idx == p.bDepth) { idx = 0; lap ^= 1; }
		}
	} else if (threadIdx.x == 32) {
		unsigned idx = 0, lap = 0;
		for (unsigned long long u = u0; u < u1; ++u) {
			mbarWait(smemAddr(&full[idx]), lap);
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(&empty[idx])) : "memory");
			if (++idx == p.depth) { idx = 0; lap ^= 1; }
		}
	}
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
	const unsigned m = 100000, n = 10000;
	float* V;
	cudaMalloc(&V, (size_t)m * n * 4);
	cudaMemset(V, 0, (size_t)m * n * 4);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	EncodeFn encode = (EncodeFn)fn;
	cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	printf("pattern                         box(rows x cols)  depth  grid  promo  GB/s\n");
	struct Cfg { const char* name; unsigned boxR, boxC; int mode; bool sw; };
	// mode 0 = W^T V order: tile = 128/256 columns, stages walk down the rows; mode 1 = V H^T: tile = rows, stages walk the columns
	const Cfg cfgs[] = {{"WtV  32r x128c sw128", 32, 128, 0, true}, {"WtV  32r x256c sw128", 32, 256, 0, true}, {"WtV  64r x128c none ", 64, 128, 0, false},
	                    {"VHt 128r x 32c none ", 128, 32, 1, false}, {"VHt 256r x 32c none ", 256, 32, 1, false}, {"VHt 128r x 64c none ", 128, 64, 1, false}};
	for (const Cfg& c : cfgs)
		for (unsigned depth : {4u, 6u})
			for (unsigned grid : {148u})
				for (int promo : {0, 1})
				for (unsigned bDepth : {0u, 2u, 4u, 6u})
				for (unsigned bRows : {64u, 32u, 16u}) {
					if (c.boxR * c.boxC != 4096) continue;
					if (bDepth == 0 && bRows != 64) continue;
					Params p;
					{   // B: the first 128 columns of V seen as a 100000 x 128 matrix (51 MB, L2 resident): box 32 rows x bRows columns, twice per stage
						const cuuint64_t bd[2] = {m, 128};
						const cuuint64_t bs[1] = {(cuuint64_t)m * 4};
						const cuuint32_t bb[2] = {32, bRows};
						const cuuint32_t be[2] = {1, 1};
						encode(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, V, bd, bs, bb, be, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
						       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
						p.bDepth = bDepth;
						p.bBoxBytes = 32 * bRows * 4;
					}
					const cuuint64_t dims[2] = {m, n};
					const cuuint64_t strides[1] = {(cuuint64_t)m * 4};
					const cuuint32_t box[2] = {c.boxR, c.boxC};
					const cuuint32_t elem[2] = {1, 1};
					if (encode(&p.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, V, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
					           c.sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
					           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
					p.boxBytes = c.boxR * c.boxC * 4;
					p.depth = depth;
					p.mode = c.mode;
					if (c.mode == 0) { p.tiles = (n + c.boxC - 1) / c.boxC; p.stagesPerTile = (m + c.boxR - 1) / c.boxR; p.tileStep = c.boxC; p.stageStep = c.boxR; }
					else { p.tiles = (m + c.boxR - 1) / c.boxR; p.stagesPerTile = (n + c.boxC - 1) / c.boxC; p.tileStep = c.boxR; p.stageStep = c.boxC; }
					p.units = (unsigned long long)p.tiles * p.stagesPerTile;
					p.evictFirst = 1;
					const size_t smem = 1024 + (size_t)depth * p.boxBytes + (size_t)bDepth * 2 * p.bBoxBytes;
					if (smem > 200 * 1024 / (grid / 148)) continue;
					stream<<<grid, 128, smem>>>(p);
					cudaEventRecord(e0);
					stream<<<grid, 128, smem>>>(p);
					cudaEventRecord(e1);
					if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
					float ms;
					cudaEventElapsedTime(&ms, e0, e1);
					printf("%s  %4u x %4u      %3u   %4u   %d   %8.1f   B: depth %u x 2 x %5u B -> %7.1f GB/s\n", c.name, c.boxR, c.boxC, depth, grid, promo, (double)m * n * 4 / ms / 1e6,
					       bDepth, p.bBoxBytes, bDepth ? (double)p.units * 2 * p.bBoxBytes / ms / 1e6 : 0.0);
				}
	return 0;
}
