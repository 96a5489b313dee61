// placeholder: replaced by the tcgen05 implementation
#include "tc_gemm.h"
#include "common.h"
namespace nmfgpu { namespace b200 { namespace tc {
bool shapeSupported(unsigned, unsigned, unsigned, size_t, size_t) { return false; }
void makePlan(Plan&, unsigned, unsigned, unsigned, const float*, size_t, const float*, const float*, size_t, const float*, const float*, size_t, bool) {}
void gemmWtV(const Plan&, float*, size_t, size_t, cudaStream_t) {}
void gemmVHt(const Plan&, float*, size_t, size_t, cudaStream_t) {}
void splitTransposeH(unsigned, unsigned, const float*, size_t, float*, float*, size_t, cudaStream_t) {}
}}}
