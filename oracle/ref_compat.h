// ref_compat.h -- force-included (-include) when oracle/build_ref.sh compiles the UNMODIFIED
// reference sources from /root/reference for sm_100.  Test infrastructure; never part of the product.
// It only supplies what CUDA 12.9 / gcc 13 removed since the reference was written:
//   * <limits> is no longer pulled in transitively,
//   * the pre-Volta warp shuffles (__shfl_xor/__shfl_up/__shfl) are gone -> full-mask *_sync forms,
//   * the legacy cuSPARSE dense<->sparse conversions are gone -> stubs returning NOT_SUPPORTED
//     (so the shimmed reference must never be fed sparse input).
#pragma once
#include <limits>
#include <stdexcept>
#include <cstdio>
#include <cusparse.h>

#ifdef __CUDACC__
#define __shfl_xor(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#define __shfl_up(v, d) __shfl_up_sync(0xffffffffu, (v), (d))
#define __shfl(v, l) __shfl_sync(0xffffffffu, (v), (l))
#endif

#define REF_COMPAT_STUB(name, T)                                                                     \
	static inline cusparseStatus_t name(cusparseHandle_t, int, int, const cusparseMatDescr_t, const T*, \
	                                    const int*, const int*, T*, int) {                             \
		return CUSPARSE_STATUS_NOT_SUPPORTED;                                                          \
	}
REF_COMPAT_STUB(cusparseScsr2dense, float)
REF_COMPAT_STUB(cusparseDcsr2dense, double)
REF_COMPAT_STUB(cusparseScsc2dense, float)
REF_COMPAT_STUB(cusparseDcsc2dense, double)
#undef REF_COMPAT_STUB

#define REF_COMPAT_STUB2(name, T)                                                                    \
	static inline cusparseStatus_t name(cusparseHandle_t, int, int, const cusparseMatDescr_t, const T*, \
	                                    int, const int*, T*, int*, int*) {                             \
		return CUSPARSE_STATUS_NOT_SUPPORTED;                                                          \
	}
REF_COMPAT_STUB2(cusparseSdense2csr, float)
REF_COMPAT_STUB2(cusparseDdense2csr, double)
REF_COMPAT_STUB2(cusparseSdense2csc, float)
REF_COMPAT_STUB2(cusparseDdense2csc, double)
#undef REF_COMPAT_STUB2
