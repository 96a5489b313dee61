// common.h -- shared host-side plumbing of the B200 NMF engine: error translation, verbosity-gated
// logging, RAII device/pinned buffers.  Replaces the roles of the reference's source/common/Logging.*
// and source/common/Memory.h with the documented ResultType behaviour (the
// reference only logs CUDA failures and carries on; here they surface as return codes).
#pragma once
#include <chrono>
#include <cstdlib>

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/nmfgpu.h"

namespace nmfgpu {
namespace b200 {

// ---- logging ------------------------------------------------------------------
Verbosity currentVerbosity();
void setCurrentVerbosity(Verbosity v);
void logf(Verbosity level, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void errorf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));

// ---- error translation --------------------------------------------------------
struct EngineError : std::runtime_error {
	ResultType code;
	EngineError(ResultType c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void cudaCheck(cudaError_t e, const char* expr, const char* file, int line) {
	if (e == cudaSuccess) return;
	char buf[512];
	snprintf(buf, sizeof(buf), "%s:%d: %s -> %s", file, line, expr, cudaGetErrorString(e));
	cudaGetLastError();  // clear the sticky-less error state
	throw EngineError(e == cudaErrorMemoryAllocation ? ResultType::ErrorNotEnoughDeviceMemory : ResultType::ErrorExternalLibrary, buf);
}
#define CUDA_CHECK(expr) ::nmfgpu::b200::cudaCheck((expr), #expr, __FILE__, __LINE__)

// ---- buffers --------------------------------------------------------------------
// Device and pinned blocks are recycled between calls on the same device (host.cpp): cudaMalloc / cudaFree of the
// 4 GB input and the ~40 scratch buffers of a compute() call cost ~100 ms, a quarter of a 100-iteration run.  A block
// goes back to the pool when its buffer dies and is handed out again for a request of exactly the same size; everything
// is returned to the driver by nmfgpu_finalize(), or earlier when an allocation fails.  NMFGPU_POOL=0 switches the pool
// off (every buffer is then freed before compute() returns, as in the reference, SingleGpuDispatcher.cpp:237-238).
void* pooledDeviceAlloc(size_t bytes);
void pooledDeviceFree(void* p, size_t bytes);
void* pooledPinnedAlloc(size_t bytes);
void pooledPinnedFree(void* p, size_t bytes);
void releasePooledMemory();

template <typename T>
class DeviceBuffer {
	T* m_ptr = nullptr;
	size_t m_count = 0;
	bool m_owned = true;

public:
	DeviceBuffer() = default;
	DeviceBuffer(const DeviceBuffer&) = delete;
	DeviceBuffer& operator=(const DeviceBuffer&) = delete;
	~DeviceBuffer() { release(); }
	void allocate(size_t count) {
		release();
		if (count == 0) return;
		m_ptr = static_cast<T*>(pooledDeviceAlloc(count * sizeof(T)));
		m_count = count;
		m_owned = true;
	}
	// wrap memory owned by the caller (device-resident V handed in through the session API)
	void adopt(T* ptr, size_t count) {
		release();
		m_ptr = ptr;
		m_count = count;
		m_owned = false;
	}
	void release() {
		if (m_ptr && m_owned) pooledDeviceFree(m_ptr, m_count * sizeof(T));
		m_ptr = nullptr;
		m_count = 0;
	}
	T* get() const { return m_ptr; }
	size_t count() const { return m_count; }
	size_t bytes() const { return m_count * sizeof(T); }
	void zero(cudaStream_t s) { if (m_ptr) CUDA_CHECK(cudaMemsetAsync(m_ptr, 0, bytes(), s)); }
};

template <typename T>
class PinnedBuffer {
	T* m_ptr = nullptr;
	size_t m_count = 0;

public:
	PinnedBuffer() = default;
	PinnedBuffer(const PinnedBuffer&) = delete;
	PinnedBuffer& operator=(const PinnedBuffer&) = delete;
	~PinnedBuffer() { release(); }
	void allocate(size_t count) {
		release();
		if (count == 0) return;
		m_ptr = static_cast<T*>(pooledPinnedAlloc(count * sizeof(T)));
		m_count = count;
	}
	void release() {
		if (m_ptr) pooledPinnedFree(m_ptr, m_count * sizeof(T));
		m_ptr = nullptr;
		m_count = 0;
	}
	T* get() const { return m_ptr; }
	size_t count() const { return m_count; }
};

inline size_t roundUp(size_t x, size_t to) { return (x + to - 1) / to * to; }
inline unsigned ceilDiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

// NMFGPU_TIMING=1: wall-clock milliseconds of the host-visible phases of a compute() call on stderr (each mark
// synchronises the device first, so the run itself gets slower; a diagnostic, not a profiler)
class PhaseTimer {
	bool m_on;
	std::chrono::steady_clock::time_point m_last;

public:
	PhaseTimer() : m_on(getenv("NMFGPU_TIMING") != nullptr), m_last(std::chrono::steady_clock::now()) {}
	void mark(const char* what) {
		if (!m_on) return;
		cudaDeviceSynchronize();
		const auto now = std::chrono::steady_clock::now();
		errorf("[timing] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - m_last).count());
		m_last = now;
	}
};

}  // namespace b200
}  // namespace nmfgpu
