// dist.h -- column-sharded multi-GPU execution (the reference is single-GPU only).
//
// One process per GPU.  Rank g holds V[:, J_g] and H[:, J_g]; W and the k x k Gram matrices are
// replicated.  Two dataflows (engine.cu):
//   all-reduce  : per iteration the m x k partial V H^T and the k x k partial H H^T are all-reduced and every rank
//                 repeats the W update (all algorithms);
//   row owners  : every rank additionally holds the row block V[I_g, :] (built once from the column shards by a
//                 grouped send/recv), so V[I_g, :] H^T needs no reduction; per iteration H (k x n) and the
//                 updated row blocks of W (m x k) are all-gathered and k*k + k statistics all-reduced (MU).
// NCCL is resolved with dlopen at the time a communicator is created, so the single-GPU library has no NCCL dependency.
#pragma once
#include <vector>

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
	// recv[r * countPerRank ...] <- send of rank r (equal counts on every rank)
	void allGather(const float* send, float* recv, size_t countPerRank, cudaStream_t stream);
	// two all-gathers in one NCCL group (one launch): a matrix block and the small statistics that travel with it
	void allGatherPair(const float* sendA, float* recvA, size_t countA, const float* sendB, float* recvB, size_t countB, cudaStream_t stream);
	// one grouped point-to-point exchange: every send/recv pair of the group proceeds concurrently
	struct Transfer {
		float* buffer;
		size_t count;
		int peer;
	};
	void exchange(const std::vector<Transfer>& sends, const std::vector<Transfer>& recvs, cudaStream_t stream);
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
