// dist.h -- column-sharded multi-GPU execution (the reference is single-GPU only).
//
// One process per GPU.  Rank g holds V[:, J_g] and H[:, J_g]; W and the k x k Gram matrices are
// replicated.  Per iteration exactly two buffers cross NVLink: the m x k partial V H^T and the k x k
// partial H H^T (plus one scalar on error iterations).  NCCL is resolved with dlopen at the time a
// communicator is created, so the single-GPU library has no NCCL dependency.
#pragma once
#include "common.h"

namespace nmfgpu {
namespace b200 {

class Communicator {
public:
	// uniqueId: the 128 bytes of an ncclUniqueId created on rank 0 (nmfgpu_b200_dist_unique_id)
	static Communicator* create(int rank, int worldSize, const void* uniqueId, unsigned globalColumns, unsigned columnOffset);
	static void makeUniqueId(void* out128);
	~Communicator();

	int rank() const { return m_rank; }
	int worldSize() const { return m_world; }
	unsigned globalColumns() const { return m_globalColumns; }
	unsigned columnOffset() const { return m_columnOffset; }
	void setShard(unsigned globalColumns, unsigned columnOffset) { m_globalColumns = globalColumns; m_columnOffset = columnOffset; }

	void allReduceSum(float* buffer, size_t count, cudaStream_t stream);
	void allReduceSum(double* buffer, size_t count, cudaStream_t stream);
	double allReduceSumHost(double value);  // blocking; used once per error iteration
	unsigned long long calls() const { return m_calls; }

private:
	Communicator() = default;
	int m_rank = 0, m_world = 1;
	unsigned m_globalColumns = 0, m_columnOffset = 0;
	void* m_comm = nullptr;
	double* m_scalar = nullptr;
	cudaStream_t m_stream = nullptr;
	unsigned long long m_calls = 0;
};

}  // namespace b200
}  // namespace nmfgpu
