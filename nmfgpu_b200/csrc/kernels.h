// kernels.h -- launchers of the SIMT (exact fp32 / fp64) kernels of the NMF iteration.
//
// These are the "fp32-exact mode" of BASELINE.json's north star and the only path for fp64 and for
// factor ranks the tensor-core path does not cover.  They replace, fused, the reference's cuBLAS
// calls G1..G6 and its elementwise kernels.  All matrices are
// column-major with explicit leading dimensions.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace nmfgpu {
namespace b200 {
namespace kern {

// C (ka x nb, col-major) = A^T B reduced over the `rows` dimension; A is rows x ka, B is rows x nb.
// The reduction is cut into `splits` contiguous row ranges; split s writes its partial product to
// C + s * splitStride (deterministic: no atomics).  Covers W^T V (MU.h:187) and W^T W (MU.h:168).
template <typename T>
void gemmTN(unsigned rows, unsigned ka, unsigned nb, const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc,
            unsigned splits, size_t splitStride, cudaStream_t stream);

// C (ma x kb, col-major) = A B^T reduced over `cols`; A is ma x cols, B is kb x cols.
// Covers V H^T (MU.h:240) and H H^T (MU.h:208).
template <typename T>
void gemmNT(unsigned ma, unsigned cols, unsigned kb, const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc,
            unsigned splits, size_t splitStride, cudaStream_t stream);

// number of split partials gemmTN / gemmNT actually write for a requested split count
unsigned effectiveSplits(unsigned reduceLen, unsigned splits);

// dst[i] = sum_s src[s * splitStride + i]  for a rows x cols column-major block (fixed order).
// Partial products written by the stream-K tensor-core kernels (tc_gemm.h) have a per-tile slot count:
// when tileSlots != nullptr the number of partials of an element is tileSlots[t], t = its 128-wide tile
// along the rows (tilesAlongRows) or the columns; the same convention holds for updateH (column tiles)
// and updateW (row tiles).  `corr` (rank entries) is the rank-one term the mean-centred tensor-core products
// leave out (tc_gemm.h): it is added once to every element, indexed along the short dimension.
template <typename T>
void sumSplits(unsigned rows, unsigned cols, const T* src, size_t ldsrc, unsigned splits, size_t splitStride, T* dst, size_t lddst,
               cudaStream_t stream, const unsigned char* tileSlots = nullptr, bool tilesAlongRows = false, const T* corr = nullptr);

// ---- H side ---------------------------------------------------------------------------------------
// Hout = Hin o N / (G Hin + eps), N = sum of `splits` partials  (MU.h:181-191, KernelMultiplyDivide.cu:42).
// If tracePartials != nullptr also tracePartials[j] = sum_r Hout[r,j] * N[r,j]  (MU.h:194-197).
// If HtHi/HtLo != nullptr also writes the transposed (n-contiguous) TF32 hi/lo split of Hout.
template <typename T>
void updateH(unsigned k, unsigned n, const T* G, const T* Hin, T* Hout, size_t ldh, const T* Npart, size_t ldn, unsigned splits,
             size_t splitStride, T eps, T* tracePartials, float* HtHi, float* HtLo, size_t ldht, cudaStream_t stream,
             const unsigned char* tileSlots = nullptr, const T* corr = nullptr, T* rowSumPartials = nullptr);
// rowSumPartials (fp32, rank <= 128): [ceil(n / 64)][k] row sums of Hout per 64-column block, for the centring term of V H^T;
// finishPartialSums adds them up: out[r] = scale * sum_b partials[b * count + r] (fp64 accumulation, fixed order)
void finishPartialSums(unsigned count, unsigned blocks, const float* partials, float scale, float* out, cudaStream_t stream);

// Hout = max(0, N) after N was overwritten by the least-squares solve (GDCLS.h:206-209)
template <typename T>
void clampNonNegative(unsigned rows, unsigned cols, T* A, size_t lda, cudaStream_t stream);

// ---- W side ---------------------------------------------------------------------------------------
// Wout = Win o P / (Win B + eps), P = sum of `splits` partials (MU.h:235-244).  Also accumulates the
// per-block column sums of squares of Wout into colSqPartials[block][k] (first half of MU.h:247).
// Returns the number of row blocks written.
template <typename T>
unsigned updateW(unsigned m, unsigned k, const T* B, const T* Win, T* Wout, size_t ldw, const T* Ppart, size_t ldp, unsigned splits,
                 size_t splitStride, T eps, T* colSqPartials, cudaStream_t stream, const unsigned char* tileSlots = nullptr,
                 const T* corr = nullptr, T* colSumPartials = nullptr);   // colSumPartials: per-block column sums of Wout (fp32, rank <= 128)

// colSq[c] = sum_b colSqPartials[b][c] (fixed order) ; used by scaleColumns
// With colSumPartials also corrOut[c] = center * (column sum of the unit-column matrix), the centring term of W^T V (tc_gemm.h)
template <typename T>
void finishColumnNorms(unsigned k, unsigned blocks, const T* colSqPartials, T* colSq, cudaStream_t stream, const T* colSumPartials = nullptr,
                       float center = 0.f, float* corrOut = nullptr);

// per-block column sums of squares of an m x k matrix (for paths that do not go through updateW)
template <typename T>
unsigned columnSquares(unsigned m, unsigned k, const T* W, size_t ldw, T* colSqPartials, cudaStream_t stream);

// W[:,c] /= sqrt(colSq[c]) where colSq[c] > 0 (KernelNormalizeColumns.cu:52-58); optional TF32 hi/lo copies.
template <typename T>
void scaleColumns(unsigned m, unsigned k, T* W, size_t ldw, const T* colSq, float* Whi, float* Wlo, cudaStream_t stream);

// ---- residual terms -------------------------------------------------------------------------------
// partial[j] = sum_i A[i,j] * B[i,j]   (KernelTraceMultiplication.cu transposeA=true): tr(V^T V), tr(W_old^T P)
template <typename T>
void columnDots(unsigned rows, unsigned cols, const T* A, size_t lda, const T* B, size_t ldb, T* partial, cudaStream_t stream);

// partial[d] = sum_i A[d,i] * B[i,d] on k x k operands (transposeA=false): tr(H H^T  W^T W)
template <typename T>
void traceKK(unsigned k, const T* A, const T* B, T* partial, cudaStream_t stream);

// ---- small k x k / elementwise helpers ---------------------------------------------------------------
// G[r,c] += (r == c ? diag : offdiag)   (KernelFillMatrix.cu ReuseValue=true)
template <typename T>
void addConstraint(unsigned k, T* G, T offdiag, T diag, cudaStream_t stream);

// X (m x k) = W (m x k) * S with S = (1-theta) I + theta/k 11^T applied analytically (nsNMF.h:174)
template <typename T>
void smoothRight(unsigned m, unsigned k, const T* W, size_t ldw, T* X, size_t ldx, T theta, cudaStream_t stream);
// Y (k x n) = S * H
template <typename T>
void smoothLeft(unsigned k, unsigned n, const T* H, size_t ldh, T* Y, size_t ldy, T theta, cudaStream_t stream);

// Solve (k x k) G X = R in place for `nrhs` right-hand sides (R is k x nrhs, col-major) when
// transposed == false, or X G^T = R (R is nrhs x k) when transposed == true; then clamp at zero.
// One thread block factorises G by Householder QR in shared memory (the reference's route:
// geqrf/ormqr/trsm, Matrix.h:565-618) into `factor` (k*k + k values), then every column / row is
// solved independently.
// fp32 with `inverse` (k*k values) and `work` (3 k*k + k doubles), k <= 128: also leaves the explicit inverse M = R^-1 Q^T,
// formed in fp64, for the tiled products of qrSolveClamp.
template <typename T>
void qrFactor(unsigned k, const T* G, T* factor, cudaStream_t stream, T* inverse = nullptr, double* work = nullptr);
template <typename T>
void qrSolveClamp(unsigned k, const T* factor, T* R, size_t ldr, unsigned nrhs, bool transposed, cudaStream_t stream,
                  T* inverse = nullptr);   // the M qrFactor left: fp32 then runs max(0, M R) resp. max(0, R M^T) as a tiled product

// TF32 hi/lo split of a dense block: hi = rn_tf32(x), lo = x - hi
void splitTf32(unsigned rows, unsigned cols, const float* X, size_t ldx, float* hi, float* lo, size_t ldo, cudaStream_t stream);

// |x| and max(0,x) variants for the k-means based initialisations (KMeansStrategy.cpp:31-40)
template <typename T>
void absInPlace(unsigned rows, unsigned cols, T* A, size_t lda, cudaStream_t stream);

// out <- in sorted ascending (radix sort).  Returns the scratch bytes needed; sorts only when temp holds that many.
template <typename T>
size_t sortAscending(const T* in, T* out, unsigned count, void* temp, size_t tempBytes, cudaStream_t stream);

// partial[2 b] = sum |x|, partial[2 b + 1] = sum x^2 over block b's share of the rows x cols matrix X (fp64; `blocks` blocks)
template <typename T>
void absSquareSums(unsigned rows, unsigned cols, const T* X, size_t ld, double* partial, unsigned blocks, cudaStream_t stream);

}  // namespace kern
}  // namespace b200
}  // namespace nmfgpu
