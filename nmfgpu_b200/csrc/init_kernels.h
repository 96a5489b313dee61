// init_kernels.h -- the MeanColumns initialisation (reference source/init/MeanColumnStrategy.cpp:41-56,
// source/init/KernelMeanColumn.cu:30-50): every column of W is the mean of five randomly chosen data columns.
#pragma once
#include "common.h"

namespace nmfgpu {
namespace b200 {
namespace init {

template <typename T>
void meanColumns(unsigned m, unsigned n, unsigned k, const T* V, size_t ldv, T* W, size_t ldw, unsigned seed, cudaStream_t stream);

}  // namespace init
}  // namespace b200
}  // namespace nmfgpu
