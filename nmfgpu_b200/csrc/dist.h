// dist.h -- multi-GPU execution of column-sharded problems (the reference is single-GPU only).
//
// One rank per GPU.  The CALLER's contract is the north star's: rank g hands in V[:, J_g] and H[:, J_g] and receives
// H[:, J_g] and the full W.  Inside, two dataflows (engine.cu):
//   row blocks  : MU on the tensor-core path.  At setup the column shards are regrouped into row blocks V[I_g, :]
//                 (one grouped send/recv).  Both V-sized products then run on the block: W[I_g]^T V[I_g, :] is a k x n
//                 PARTIAL of W^T V, and V[I_g, :] H^T needs no reduction at all.  m >> n in the workloads this library
//                 serves, so the exchanged object is the small factor: per iteration each rank pushes its partial of
//                 W^T V, tile by tile from inside the tensor-core kernel, into the memory of the rank that owns those
//                 columns (NVLink peer stores), the owners update their columns of H and push them to everyone, and
//                 k*k + k statistics travel the same way.  No NCCL call in the iteration, no m x k exchange.
//   all-reduce  : every other algorithm / precision / sparse execution: the m x k partial V H^T and the k x k partial
//                 H H^T are all-reduced and every rank repeats the W update.
//
// Two transports behind one interface:
//   NCCL + CUDA IPC : one PROCESS per GPU (torchrun).  NCCL (resolved with dlopen) carries the setup-time collectives;
//                     cudaIpc maps every rank's exchange buffer into every process for the peer stores.
//   local           : one THREAD per rank inside one process (unique id from makeLocalUniqueId), any device assignment
//                     including all ranks on ONE GPU.  Pointers are shared directly.  This is what lets
//                     `pytest -m gpu` run the sharded dataflows of engine.cu on a single-GPU box.
#pragma once
#include <vector>

#include "common.h"

namespace nmfgpu {
namespace b200 {

class Communicator {
public:
	// uniqueId: 128 bytes from makeUniqueId (NCCL) or makeLocalUniqueId (threads of this process)
	static Communicator* create(int rank, int worldSize, const void* uniqueId, unsigned globalColumns, unsigned columnOffset);
	static void makeUniqueId(void* out128);
	static void makeLocalUniqueId(void* out128);
	virtual ~Communicator() = default;

	int rank() const { return m_rank; }
	int worldSize() const { return m_world; }
	unsigned globalColumns() const { return m_globalColumns; }
	unsigned columnOffset() const { return m_columnOffset; }
	void setShard(unsigned globalColumns, unsigned columnOffset) { m_globalColumns = globalColumns; m_columnOffset = columnOffset; }
	unsigned long long calls() const { return m_calls; }

	// whether the device collectives below may be recorded into a CUDA graph (the local transport synchronises inside them)
	virtual bool capturable() const = 0;
	// whether two ranks may share one GPU (the thread transport of the tests).  A kernel that waits for another rank can
	// then keep that rank's kernels from running, so the engine lets the ranks meet on the host before every such kernel
	// (engine.cu m_hostLockstep).
	virtual bool ranksMayShareDevice() const = 0;
	// ---- device collectives, enqueued on `stream` (NCCL) or completed before returning (local)
	virtual void allReduceSum(float* buffer, size_t count, cudaStream_t stream) = 0;
	virtual void allReduceSum(double* buffer, size_t count, cudaStream_t stream) = 0;
	// recv[r * countPerRank ...] <- send of rank r (equal counts on every rank)
	virtual void allGather(const float* send, float* recv, size_t countPerRank, cudaStream_t stream) = 0;
	// one grouped point-to-point exchange: every send/recv pair of the group proceeds concurrently.  The i-th send of
	// rank a to rank b is matched with the i-th recv of rank b from rank a.
	struct Transfer {
		float* buffer;
		size_t count;
		int peer;
	};
	virtual void exchange(const std::vector<Transfer>& sends, const std::vector<Transfer>& recvs, cudaStream_t stream) = 0;

	// ---- host collectives (blocking)
	// all[r * bytes ...] <- mine of rank r
	virtual void allGatherHost(const void* mine, size_t bytes, void* all) = 0;
	double allReduceSumHost(double value);   // summed in rank order: the same bits on every rank
	void barrier();

	// ---- peer memory: every rank passes its own exchange buffer (same size everywhere); the result holds, per rank, a
	// pointer this rank's kernels can store to and load from (own entry = localBase).  Blocking collective.
	// closePeers must run on every rank before the buffers are freed.
	virtual std::vector<void*> openPeers(void* localBase) = 0;
	virtual void closePeers(std::vector<void*>& peers) = 0;

protected:
	Communicator() = default;
	int m_rank = 0, m_world = 1;
	unsigned m_globalColumns = 0, m_columnOffset = 0;
	unsigned long long m_calls = 0;
};

}  // namespace b200
}  // namespace nmfgpu
