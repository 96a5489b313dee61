// tc_gemm.h -- the two V-sized contractions of the NMF iteration on the B200 tensor cores.
//
//   gemmWtV :  N  (k x n) = W^T V      replaces cublasSgemm(T,N)  reference MU.h:187-188  (G3 in SURVEY.md 2.2)
//   gemmVHt :  N2 (m x k) = V H^T      replaces cublasSgemm(N,T)  reference MU.h:240-241  (G6)
//
// Both stream the 4 GB matrix V exactly once per call and are HBM-bound by design (k/2 FLOP per byte
// of V).  To stay within fp32 tolerance on TF32 tensor cores they run the 3xTF32 scheme
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   x_hi = rn_tf32(x), x_lo = x - x_hi
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

struct Plan {
	unsigned m = 0, n = 0, k = 0;
	unsigned kp = 0;               // rank padded to the UMMA N granularity (16, 32, 64 or 128)
	unsigned splitsWtV = 1;        // partial products written by gemmWtV (one per reduction slice)
	unsigned splitsVHt = 1;
	bool singlePass = false;       // 1xTF32 diagnostic mode
	// work decomposition (filled by makePlan)
	unsigned wtvTilesN = 0, wtvChunksPerSplit = 0;
	unsigned vhtTilesM = 0, vhtChunksPerSplit = 0;
	unsigned gridWtV = 0, gridVHt = 0;
	// TMA descriptors (CUtensorMap, 128 bytes each, 64-byte aligned)
	alignas(64) unsigned char mapV_wtv[128];   // V as [32 rows x 128 cols] boxes, 128B swizzle
	alignas(64) unsigned char mapV_vht[128];   // V as [128 rows x 32 cols] boxes, no swizzle
	alignas(64) unsigned char mapWhi[128];     // W hi/lo: [32 rows x kp cols] boxes, 128B swizzle (K-major B operand)
	alignas(64) unsigned char mapWlo[128];
	alignas(64) unsigned char mapHtHi[128];    // H^T hi/lo (n x k, n contiguous): [32 x kp] boxes, 128B swizzle
	alignas(64) unsigned char mapHtLo[128];
};

// fp32 problem shapes the tensor-core path covers (others run the SIMT kernels)
bool shapeSupported(unsigned m, unsigned n, unsigned k, size_t ldV, size_t ldW);

void makePlan(Plan& plan, unsigned m, unsigned n, unsigned k, const float* V, size_t ldV, const float* Whi, const float* Wlo, size_t ldW,
              const float* HtHi, const float* HtLo, size_t ldHt, bool singlePass);

// Npart + s*splitStride (k x n, leading dimension ldn) receives the partial product of reduction slice s
void gemmWtV(const Plan& plan, float* Npart, size_t ldn, size_t splitStride, cudaStream_t stream);
// Ppart + s*splitStride (m x k, leading dimension ldp)
void gemmVHt(const Plan& plan, float* Ppart, size_t ldp, size_t splitStride, cudaStream_t stream);

// hi/lo TF32 split of H (k x n, column-major) written transposed: Ht[c * ldht + j] = H[c + j * ldh]
void splitTransposeH(unsigned k, unsigned n, const float* H, size_t ldh, float* hi, float* lo, size_t ldht, cudaStream_t stream);

}  // namespace tc
}  // namespace b200
}  // namespace nmfgpu
