// dist.cpp -- see dist.h.  (Compiled as CUDA: the local transport needs one tiny summation kernel.)
#include "dist.h"

#include <dlfcn.h>
#include <nccl.h>

#include <array>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

namespace nmfgpu {
namespace b200 {

double Communicator::allReduceSumHost(double value) {
	if (m_world <= 1) return value;
	std::vector<double> all((size_t)m_world);
	allGatherHost(&value, sizeof(double), all.data());
	double sum = 0.0;
	for (double v : all) sum += v;   // rank order on every rank
	return sum;
}

void Communicator::barrier() {
	if (m_world <= 1) return;
	unsigned char token = 0;
	std::vector<unsigned char> all((size_t)m_world);
	allGatherHost(&token, 1, all.data());
}

namespace {

// =====================================================================================================================
// NCCL + CUDA IPC: one process per GPU
// =====================================================================================================================
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

class NcclCommunicator : public Communicator {
	ncclComm_t m_comm = nullptr;
	cudaStream_t m_stream = nullptr;      // host collectives
	unsigned char* m_stage = nullptr;     // device staging of allGatherHost: [own bytes][world x bytes]
	size_t m_stageBytes = 0;

public:
	NcclCommunicator(int rank, int world, const void* uniqueId, unsigned globalColumns, unsigned columnOffset) {
		m_rank = rank;
		m_world = world;
		m_globalColumns = globalColumns;
		m_columnOffset = columnOffset;
		if (world > 1) {
			ncclUniqueId id;
			std::memcpy(&id, uniqueId, sizeof(id));
			ncclCheck(api().commInitRank(&m_comm, world, id, rank), "ncclCommInitRank");
			CUDA_CHECK(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking));
		}
	}
	~NcclCommunicator() override {
		if (m_comm) api().commDestroy(m_comm);
		if (m_stage) cudaFree(m_stage);
		if (m_stream) cudaStreamDestroy(m_stream);
	}

	bool capturable() const override { return true; }
	bool ranksMayShareDevice() const override { return false; }   // one process per GPU
	void allReduceSum(float* buffer, size_t count, cudaStream_t stream) override {
		if (m_world <= 1) return;
		ncclCheck(api().allReduce(buffer, buffer, count, ncclFloat, ncclSum, m_comm, stream), "ncclAllReduce");
		++m_calls;
	}
	void allReduceSum(double* buffer, size_t count, cudaStream_t stream) override {
		if (m_world <= 1) return;
		ncclCheck(api().allReduce(buffer, buffer, count, ncclDouble, ncclSum, m_comm, stream), "ncclAllReduce");
		++m_calls;
	}
	void allGather(const float* send, float* recv, size_t countPerRank, cudaStream_t stream) override {
		if (m_world <= 1) return;
		ncclCheck(api().allGather(send, recv, countPerRank, ncclFloat, m_comm, stream), "ncclAllGather");
		++m_calls;
	}
	void exchange(const std::vector<Transfer>& sends, const std::vector<Transfer>& recvs, cudaStream_t stream) override {
		if (m_world <= 1) return;
		ncclCheck(api().groupStart(), "ncclGroupStart");
		for (const Transfer& t : sends) ncclCheck(api().send(t.buffer, t.count, ncclFloat, t.peer, m_comm, stream), "ncclSend");
		for (const Transfer& t : recvs) ncclCheck(api().recv(t.buffer, t.count, ncclFloat, t.peer, m_comm, stream), "ncclRecv");
		ncclCheck(api().groupEnd(), "ncclGroupEnd");
		++m_calls;
	}
	void allGatherHost(const void* mine, size_t bytes, void* all) override {
		if (m_world <= 1) {
			std::memcpy(all, mine, bytes);
			return;
		}
		const size_t need = bytes * (size_t)(m_world + 1);
		if (need > m_stageBytes) {
			if (m_stage) CUDA_CHECK(cudaFree(m_stage));
			m_stage = nullptr;
			m_stageBytes = std::max<size_t>(need, 4096);
			CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&m_stage), m_stageBytes));
		}
		CUDA_CHECK(cudaMemcpyAsync(m_stage, mine, bytes, cudaMemcpyHostToDevice, m_stream));
		ncclCheck(api().allGather(m_stage, m_stage + bytes, bytes, ncclChar, m_comm, m_stream), "ncclAllGather");
		CUDA_CHECK(cudaMemcpyAsync(all, m_stage + bytes, bytes * (size_t)m_world, cudaMemcpyDeviceToHost, m_stream));
		CUDA_CHECK(cudaStreamSynchronize(m_stream));
		++m_calls;
	}
	std::vector<void*> openPeers(void* localBase) override {
		std::vector<void*> peers((size_t)m_world, nullptr);
		peers[(size_t)m_rank] = localBase;
		if (m_world <= 1) return peers;
		cudaIpcMemHandle_t mine;
		CUDA_CHECK(cudaIpcGetMemHandle(&mine, localBase));
		std::vector<cudaIpcMemHandle_t> all((size_t)m_world);
		allGatherHost(&mine, sizeof(mine), all.data());
		for (int g = 0; g < m_world; ++g) {
			if (g == m_rank) continue;
			void* p = nullptr;
			const cudaError_t e = cudaIpcOpenMemHandle(&p, all[(size_t)g], cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) {
				cudaGetLastError();
				for (int h = 0; h < g; ++h)
					if (h != m_rank && peers[(size_t)h]) cudaIpcCloseMemHandle(peers[(size_t)h]);
				throw EngineError(ResultType::ErrorExternalLibrary, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e));
			}
			peers[(size_t)g] = p;
		}
		return peers;
	}
	void closePeers(std::vector<void*>& peers) override {
		for (int g = 0; g < (int)peers.size(); ++g)
			if (g != m_rank && peers[(size_t)g]) cudaIpcCloseMemHandle(peers[(size_t)g]);
		peers.clear();
	}
};

// =====================================================================================================================
// local: one thread per rank inside this process
// =====================================================================================================================
constexpr char kLocalMagic[8] = {'N', 'M', 'F', 'L', 'O', 'C', 'A', 'L'};
constexpr int kMaxLocalTransfers = 8;

struct LocalGroup {
	int world = 0;
	std::mutex mu;
	std::condition_variable cv;
	int arrived = 0;
	unsigned long long generation = 0;
	std::vector<std::vector<unsigned char>> slots;

	// a rank that never arrives (it threw) must not hang the others for ever
	void barrier() {
		std::unique_lock<std::mutex> lock(mu);
		const unsigned long long gen = generation;
		if (++arrived == world) {
			arrived = 0;
			++generation;
			cv.notify_all();
			return;
		}
		if (!cv.wait_for(lock, std::chrono::seconds(120), [&] { return generation != gen; }))
			throw EngineError(ResultType::ErrorExternalLibrary, "local communicator: a rank did not reach the barrier within 120 s");
	}
};

std::mutex g_registryLock;
std::map<std::array<unsigned char, 128>, std::weak_ptr<LocalGroup>> g_registry;
unsigned long long g_localIds = 0;

template <typename T>
__global__ void sum_ranks_kernel(T* __restrict__ dst, const T* const* __restrict__ src, int ranks, size_t count) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
		T s = T(0);
		for (int g = 0; g < ranks; ++g) s += src[g][i];
		dst[i] = s;
	}
}

class LocalCommunicator : public Communicator {
	std::shared_ptr<LocalGroup> m_group;
	bool m_shared = false;   // two ranks on one device

	template <typename T>
	void allReduceImpl(T* buffer, size_t count, cudaStream_t stream) {
		if (m_world <= 1) return;
		DeviceBuffer<T> staging;
		DeviceBuffer<const T*> table;
		staging.allocate(count);
		table.allocate((size_t)m_world);
		CUDA_CHECK(cudaMemcpyAsync(staging.get(), buffer, count * sizeof(T), cudaMemcpyDeviceToDevice, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		const T* mine = staging.get();
		std::vector<const T*> all((size_t)m_world);
		allGatherHost(&mine, sizeof(mine), all.data());
		CUDA_CHECK(cudaMemcpyAsync(table.get(), all.data(), all.size() * sizeof(const T*), cudaMemcpyHostToDevice, stream));
		const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 1184);
		sum_ranks_kernel<T><<<std::max(1u, blocks), 256, 0, stream>>>(buffer, table.get(), m_world, count);
		CUDA_CHECK(cudaGetLastError());
		CUDA_CHECK(cudaStreamSynchronize(stream));
		m_group->barrier();   // nobody frees its staging copy while another rank still reads it
		++m_calls;
	}

public:
	LocalCommunicator(int rank, int world, const void* uniqueId, unsigned globalColumns, unsigned columnOffset) {
		m_rank = rank;
		m_world = world;
		m_globalColumns = globalColumns;
		m_columnOffset = columnOffset;
		std::array<unsigned char, 128> key;
		std::memcpy(key.data(), uniqueId, 128);
		{
			std::lock_guard<std::mutex> guard(g_registryLock);
			m_group = g_registry[key].lock();
			if (!m_group) {
				m_group = std::make_shared<LocalGroup>();
				m_group->world = world;
				m_group->slots.resize((size_t)world);
				g_registry[key] = m_group;
			}
		}
		if (m_group->world != world) throw EngineError(ResultType::ErrorInvalidArgument, "local communicator: ranks disagree on the world size");
		// ranks on different devices of this process reach each other's memory directly
		int dev = 0;
		CUDA_CHECK(cudaGetDevice(&dev));
		std::vector<int> devices((size_t)world);
		allGatherHost(&dev, sizeof(dev), devices.data());
		for (int d : devices)
			if (d != dev && cudaDeviceEnablePeerAccess(d, 0) != cudaSuccess) cudaGetLastError();   // already enabled / same device
		for (size_t a = 0; a < devices.size(); ++a)
			for (size_t b = a + 1; b < devices.size(); ++b)
				if (devices[a] == devices[b]) m_shared = true;
	}

	bool capturable() const override { return false; }
	bool ranksMayShareDevice() const override { return m_shared; }
	void allReduceSum(float* buffer, size_t count, cudaStream_t stream) override { allReduceImpl(buffer, count, stream); }
	void allReduceSum(double* buffer, size_t count, cudaStream_t stream) override { allReduceImpl(buffer, count, stream); }

	void allGather(const float* send, float* recv, size_t countPerRank, cudaStream_t stream) override {
		if (m_world <= 1) return;
		CUDA_CHECK(cudaStreamSynchronize(stream));
		std::vector<const float*> all((size_t)m_world);
		allGatherHost(&send, sizeof(send), all.data());
		for (int g = 0; g < m_world; ++g)
			if (all[(size_t)g] != recv + (size_t)g * countPerRank)   // in place: the own block is already where it belongs
				CUDA_CHECK(cudaMemcpyAsync(recv + (size_t)g * countPerRank, all[(size_t)g], countPerRank * sizeof(float), cudaMemcpyDefault, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		m_group->barrier();
		++m_calls;
	}

	void exchange(const std::vector<Transfer>& sends, const std::vector<Transfer>& recvs, cudaStream_t stream) override {
		if (m_world <= 1) return;
		struct Posted {
			int count;
			Transfer t[kMaxLocalTransfers];
		};
		if ((int)sends.size() > kMaxLocalTransfers) throw EngineError(ResultType::ErrorInvalidArgument, "local communicator: too many transfers in one group");
		CUDA_CHECK(cudaStreamSynchronize(stream));
		Posted mine;
		std::memset(&mine, 0, sizeof(mine));
		mine.count = (int)sends.size();
		for (int i = 0; i < mine.count; ++i) mine.t[i] = sends[(size_t)i];
		std::vector<Posted> all((size_t)m_world);
		allGatherHost(&mine, sizeof(mine), all.data());
		std::vector<int> taken((size_t)m_world, 0);   // how many sends of each peer to this rank were matched already
		for (const Transfer& r : recvs) {
			const Posted& p = all[(size_t)r.peer];
			int seen = 0;
			const Transfer* match = nullptr;
			for (int i = 0; i < p.count; ++i)
				if (p.t[i].peer == m_rank && seen++ == taken[(size_t)r.peer]) {
					match = &p.t[i];
					break;
				}
			if (match == nullptr || match->count != r.count) throw EngineError(ResultType::ErrorInvalidArgument, "local communicator: unmatched receive");
			++taken[(size_t)r.peer];
			CUDA_CHECK(cudaMemcpyAsync(r.buffer, match->buffer, r.count * sizeof(float), cudaMemcpyDefault, stream));
		}
		CUDA_CHECK(cudaStreamSynchronize(stream));
		m_group->barrier();
		++m_calls;
	}

	void allGatherHost(const void* mine, size_t bytes, void* all) override {
		if (m_world <= 1) {
			std::memcpy(all, mine, bytes);
			return;
		}
		{
			std::lock_guard<std::mutex> guard(m_group->mu);
			auto& slot = m_group->slots[(size_t)m_rank];
			slot.assign(static_cast<const unsigned char*>(mine), static_cast<const unsigned char*>(mine) + bytes);
		}
		m_group->barrier();
		{
			std::lock_guard<std::mutex> guard(m_group->mu);
			for (int g = 0; g < m_world; ++g) {
				const auto& slot = m_group->slots[(size_t)g];
				if (slot.size() != bytes) throw EngineError(ResultType::ErrorInvalidArgument, "local communicator: ranks disagree on a message size");
				std::memcpy(static_cast<unsigned char*>(all) + (size_t)g * bytes, slot.data(), bytes);
			}
		}
		m_group->barrier();   // the slots may be overwritten again
	}

	std::vector<void*> openPeers(void* localBase) override {
		std::vector<void*> peers((size_t)m_world, nullptr);
		allGatherHost(&localBase, sizeof(localBase), peers.data());
		return peers;
	}
	void closePeers(std::vector<void*>& peers) override { peers.clear(); }
};

}  // namespace

void Communicator::makeUniqueId(void* out128) {
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	ncclCheck(api().getUniqueId(&id), "ncclGetUniqueId");
	std::memcpy(out128, &id, sizeof(id));
}

void Communicator::makeLocalUniqueId(void* out128) {
	std::memset(out128, 0, 128);
	std::memcpy(out128, kLocalMagic, sizeof(kLocalMagic));
	std::lock_guard<std::mutex> guard(g_registryLock);
	const unsigned long long serial = ++g_localIds;
	std::memcpy(static_cast<unsigned char*>(out128) + 8, &serial, sizeof(serial));
}

Communicator* Communicator::create(int rank, int worldSize, const void* uniqueId, unsigned globalColumns, unsigned columnOffset) {
	if (worldSize > 1 && std::memcmp(uniqueId, kLocalMagic, sizeof(kLocalMagic)) == 0)
		return new LocalCommunicator(rank, worldSize, uniqueId, globalColumns, columnOffset);
	return new NcclCommunicator(rank, worldSize, uniqueId, globalColumns, columnOffset);
}

}  // namespace b200
}  // namespace nmfgpu
