// dist.cpp -- see dist.h.
#include "dist.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <memory>
#include <mutex>
#include <string>

namespace nmfgpu {
namespace b200 {

namespace {
struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*getUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*commInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*allReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*allGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*groupStart)() = nullptr;
	ncclResult_t (*groupEnd)() = nullptr;
	ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
	const char* (*getErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api() {
	static NcclApi a;
	static std::once_flag once;
	std::call_once(once, [] {
		// a process that already loaded NCCL (torch) gets that copy back; otherwise the system library
		const char* names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char* n : names) {
			a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
			if (a.handle) break;
		}
		if (!a.handle) return;
		a.getUniqueId = reinterpret_cast<decltype(a.getUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
		a.commInitRank = reinterpret_cast<decltype(a.commInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
		a.allReduce = reinterpret_cast<decltype(a.allReduce)>(dlsym(a.handle, "ncclAllReduce"));
		a.allGather = reinterpret_cast<decltype(a.allGather)>(dlsym(a.handle, "ncclAllGather"));
		a.send = reinterpret_cast<decltype(a.send)>(dlsym(a.handle, "ncclSend"));
		a.recv = reinterpret_cast<decltype(a.recv)>(dlsym(a.handle, "ncclRecv"));
		a.groupStart = reinterpret_cast<decltype(a.groupStart)>(dlsym(a.handle, "ncclGroupStart"));
		a.groupEnd = reinterpret_cast<decltype(a.groupEnd)>(dlsym(a.handle, "ncclGroupEnd"));
		a.commDestroy = reinterpret_cast<decltype(a.commDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
		a.getErrorString = reinterpret_cast<decltype(a.getErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
	});
	if (!a.handle || !a.getUniqueId || !a.commInitRank || !a.allReduce || !a.commDestroy || !a.allGather || !a.send || !a.recv || !a.groupStart ||
	    !a.groupEnd)
		throw EngineError(ResultType::ErrorExternalLibrary, "NCCL (libnccl.so.2) could not be loaded");
	return a;
}

void ncclCheck(ncclResult_t r, const char* what) {
	if (r == ncclSuccess) return;
	std::string msg = std::string(what) + " failed: " + (api().getErrorString ? api().getErrorString(r) : "?");
	throw EngineError(ResultType::ErrorExternalLibrary, msg);
}
}  // namespace

void Communicator::makeUniqueId(void* out128) {
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	ncclCheck(api().getUniqueId(&id), "ncclGetUniqueId");
	std::memcpy(out128, &id, sizeof(id));
}

Communicator* Communicator::create(int rank, int worldSize, const void* uniqueId, unsigned globalColumns, unsigned columnOffset) {
	std::unique_ptr<Communicator> c(new Communicator());
	c->m_rank = rank;
	c->m_world = worldSize;
	c->m_globalColumns = globalColumns;
	c->m_columnOffset = columnOffset;
	if (worldSize > 1) {
		ncclUniqueId id;
		std::memcpy(&id, uniqueId, sizeof(id));
		ncclComm_t comm;
		ncclCheck(api().commInitRank(&comm, worldSize, id, rank), "ncclCommInitRank");
		c->m_comm = comm;
		CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->m_scalar), sizeof(double)));
		CUDA_CHECK(cudaStreamCreateWithFlags(&c->m_stream, cudaStreamNonBlocking));
	}
	return c.release();
}

Communicator::~Communicator() {
	if (m_comm) api().commDestroy(static_cast<ncclComm_t>(m_comm));
	if (m_scalar) cudaFree(m_scalar);
	if (m_stream) cudaStreamDestroy(m_stream);
}

void Communicator::allReduceSum(float* buffer, size_t count, cudaStream_t stream) {
	if (m_world <= 1) return;
	ncclCheck(api().allReduce(buffer, buffer, count, ncclFloat, ncclSum, static_cast<ncclComm_t>(m_comm), stream), "ncclAllReduce");
	++m_calls;
}

void Communicator::allReduceSum(double* buffer, size_t count, cudaStream_t stream) {
	if (m_world <= 1) return;
	ncclCheck(api().allReduce(buffer, buffer, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(m_comm), stream), "ncclAllReduce");
	++m_calls;
}

void Communicator::allGather(const float* send, float* recv, size_t countPerRank, cudaStream_t stream) {
	if (m_world <= 1) return;
	ncclCheck(api().allGather(send, recv, countPerRank, ncclFloat, static_cast<ncclComm_t>(m_comm), stream), "ncclAllGather");
	++m_calls;
}

void Communicator::allGatherPair(const float* sendA, float* recvA, size_t countA, const float* sendB, float* recvB, size_t countB, cudaStream_t stream) {
	if (m_world <= 1) return;
	ncclCheck(api().groupStart(), "ncclGroupStart");
	ncclCheck(api().allGather(sendA, recvA, countA, ncclFloat, static_cast<ncclComm_t>(m_comm), stream), "ncclAllGather");
	ncclCheck(api().allGather(sendB, recvB, countB, ncclFloat, static_cast<ncclComm_t>(m_comm), stream), "ncclAllGather");
	ncclCheck(api().groupEnd(), "ncclGroupEnd");
	++m_calls;
}

void Communicator::exchange(const std::vector<Transfer>& sends, const std::vector<Transfer>& recvs, cudaStream_t stream) {
	if (m_world <= 1) return;
	ncclCheck(api().groupStart(), "ncclGroupStart");
	for (const Transfer& t : sends) ncclCheck(api().send(t.buffer, t.count, ncclFloat, t.peer, static_cast<ncclComm_t>(m_comm), stream), "ncclSend");
	for (const Transfer& t : recvs) ncclCheck(api().recv(t.buffer, t.count, ncclFloat, t.peer, static_cast<ncclComm_t>(m_comm), stream), "ncclRecv");
	ncclCheck(api().groupEnd(), "ncclGroupEnd");
	++m_calls;
}

double Communicator::allReduceSumHost(double value) {
	if (m_world <= 1) return value;
	CUDA_CHECK(cudaMemcpyAsync(m_scalar, &value, sizeof(double), cudaMemcpyHostToDevice, m_stream));
	allReduceSum(m_scalar, 1, m_stream);
	double out = 0.0;
	CUDA_CHECK(cudaMemcpyAsync(&out, m_scalar, sizeof(double), cudaMemcpyDeviceToHost, m_stream));
	CUDA_CHECK(cudaStreamSynchronize(m_stream));
	return out;
}

}  // namespace b200
}  // namespace nmfgpu
