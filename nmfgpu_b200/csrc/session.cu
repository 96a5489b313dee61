// session.cu -- implementation of include/nmfgpu_b200.h (C-ABI extensions: precision, column shards,
// HBM-resident sessions, synthetic-workload helpers).  See that header for the contract.
#include "../../include/nmfgpu_b200.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <new>

#include "host.h"
#include "kernels.h"
#include "tc_gemm.h"

using namespace nmfgpu;
using namespace nmfgpu::b200;

struct nmfgpu_b200_session {
	std::unique_ptr<Engine<float>> engine;
	cudaEvent_t start = nullptr, stop = nullptr;
	~nmfgpu_b200_session() {
		if (start) cudaEventDestroy(start);
		if (stop) cudaEventDestroy(stop);
	}
};

namespace {
template <typename F>
int guarded(F&& body) {
	try {
		body();
		return static_cast<int>(ResultType::Success);
	} catch (const EngineError& e) {
		errorf("[ERROR] %s\n", e.what());
		return static_cast<int>(e.code);
	} catch (const std::bad_alloc&) {
		return static_cast<int>(ResultType::ErrorNotEnoughHostMemory);
	} catch (const std::exception& e) {
		errorf("[ERROR] %s\n", e.what());
		return static_cast<int>(ResultType::ErrorExternalLibrary);
	}
}

__global__ void uniform_kernel(float* __restrict__ dst, unsigned rows, unsigned cols, size_t ld, unsigned long long seed,
                               unsigned long long totalRows, unsigned long long row0, unsigned long long col0) {
	const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (size_t)rows * cols) return;
	const unsigned i = (unsigned)(idx % rows), j = (unsigned)(idx / rows);
	unsigned long long z = ((col0 + j) * totalRows + row0 + i + 1ull) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	dst[(size_t)j * ld + i] = (float)((double)((z >> 40) + 1ull) * 5.9604644775390625e-08);  // * 2^-24, exact in fp32
}

__global__ void flush_kernel(float4* p, size_t count) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
		p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
float4* g_flushBuffer = nullptr;
constexpr size_t kFlushBytes = 256ull << 20;  // twice the 126 MB L2
}  // namespace

extern "C" {

NMFGPU_EXPORT int nmfgpu_b200_set_precision(int mode) {
	Context* ctx = currentContext();
	if (ctx == nullptr) return static_cast<int>(ResultType::ErrorNotInitialized);
	if (mode < 0 || mode > 3) return static_cast<int>(ResultType::ErrorInvalidArgument);
	ctx->precision = static_cast<Precision>(mode);
	return 0;
}

NMFGPU_EXPORT int nmfgpu_b200_dist_unique_id(void* out128) {
	if (out128 == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { Communicator::makeUniqueId(out128); });
}

NMFGPU_EXPORT int nmfgpu_b200_dist_local_unique_id(void* out128) {
	if (out128 == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { Communicator::makeLocalUniqueId(out128); });
}

NMFGPU_EXPORT int nmfgpu_b200_dist_init(int rank, int world_size, const void* unique_id128) {
	Context* ctx = currentContext();
	if (ctx == nullptr) return static_cast<int>(ResultType::ErrorNotInitialized);
	if (world_size < 1 || rank < 0 || rank >= world_size || (world_size > 1 && unique_id128 == nullptr))
		return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		CUDA_CHECK(cudaSetDevice(ctx->deviceId));
		ctx->comm.reset(Communicator::create(rank, world_size, unique_id128, 0, 0));
	});
}

NMFGPU_EXPORT int nmfgpu_b200_dist_set_shard(unsigned global_columns, unsigned column_offset) {
	Context* ctx = currentContext();
	if (ctx == nullptr) return static_cast<int>(ResultType::ErrorNotInitialized);
	if (!ctx->comm) return static_cast<int>(ResultType::ErrorInvalidArgument);
	ctx->comm->setShard(global_columns, column_offset);
	return 0;
}

NMFGPU_EXPORT int nmfgpu_b200_dist_finalize(void) {
	Context* ctx = currentContext();
	if (ctx == nullptr) return static_cast<int>(ResultType::ErrorNotInitialized);
	ctx->comm.reset();
	return 0;
}

NMFGPU_EXPORT int nmfgpu_b200_session_create_f32(int algorithm, unsigned rows, unsigned columns, unsigned features, const float* v, unsigned ld_v,
                                                 int v_on_device, int constant_w, const nmfgpu_b200_named_value* params, unsigned num_params,
                                                 nmfgpu_b200_session** out) {
	Context* ctx = currentContext();
	if (ctx == nullptr) return static_cast<int>(ResultType::ErrorNotInitialized);
	if (out == nullptr || v == nullptr || algorithm < 0 || algorithm > static_cast<int>(NmfAlgorithm::nsNMF) || ld_v < rows)
		return static_cast<int>(ResultType::ErrorInvalidArgument);
	*out = nullptr;
	return guarded([&] {
		if (cudaSetDevice(ctx->deviceId) != cudaSuccess) {
			cudaGetLastError();
			throw EngineError(ResultType::ErrorDeviceSelection, "no usable CUDA device: the NMF engine has no CPU fallback");
		}
		EngineConfig cfg;
		cfg.algorithm = static_cast<NmfAlgorithm>(algorithm);
		cfg.m = rows;
		cfg.n = columns;
		cfg.k = features;
		cfg.constantW = constant_w != 0;
		cfg.precision = ctx->precision;
		cfg.comm = (ctx->comm && ctx->comm->worldSize() > 1) ? ctx->comm.get() : nullptr;
		for (unsigned i = 0; i < num_params; ++i) {
			const char* nm = params[i].name;
			if (nm == nullptr) continue;
			if (!std::strcmp(nm, "lambda")) cfg.params.lambda = params[i].value;
			else if (!std::strcmp(nm, "lambdaW")) cfg.params.lambdaW = params[i].value;
			else if (!std::strcmp(nm, "lambdaH")) cfg.params.lambdaH = params[i].value;
			else if (!std::strcmp(nm, "alphaW")) cfg.params.alphaW = params[i].value;
			else if (!std::strcmp(nm, "alphaH")) cfg.params.alphaH = params[i].value;
			else if (!std::strcmp(nm, "theta")) cfg.params.theta = params[i].value;
		}
		std::unique_ptr<nmfgpu_b200_session> s(new nmfgpu_b200_session());
		s->engine.reset(new Engine<float>(cfg));
		MatrixDescription<float> vd;
		std::memset(&vd, 0, sizeof(vd));
		vd.rows = rows;
		vd.columns = columns;
		vd.format = StorageFormat::Dense;
		vd.dense.values = const_cast<float*>(v);
		vd.dense.leadingDimension = ld_v;
		s->engine->setup(vd, v_on_device != 0);
		CUDA_CHECK(cudaEventCreate(&s->start));
		CUDA_CHECK(cudaEventCreate(&s->stop));
		*out = s.release();
	});
}

NMFGPU_EXPORT int nmfgpu_b200_session_set_factors_f32(nmfgpu_b200_session* s, const float* w, unsigned ld_w, const float* h, unsigned ld_h) {
	if (s == nullptr || w == nullptr || h == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		const EngineConfig& c = s->engine->config();
		MatrixDescription<float> wd, hd;
		std::memset(&wd, 0, sizeof(wd));
		std::memset(&hd, 0, sizeof(hd));
		wd.rows = c.m; wd.columns = c.k; wd.format = StorageFormat::Dense; wd.dense.values = const_cast<float*>(w); wd.dense.leadingDimension = ld_w;
		hd.rows = c.k; hd.columns = c.n; hd.format = StorageFormat::Dense; hd.dense.values = const_cast<float*>(h); hd.dense.leadingDimension = ld_h;
		s->engine->loadW(wd);
		s->engine->loadH(hd);
		s->engine->finishInitialisation();
		s->engine->synchronize();
	});
}

// The reference's initialisation strategies (InitializationStrategy::create, source/init/*.cpp) on the resident V, with the
// seeds host.cpp's run loop would hand them; wall-clock milliseconds of the whole initialisation (k-means synchronises).
NMFGPU_EXPORT int nmfgpu_b200_session_initialize(nmfgpu_b200_session* s, int init_method, unsigned seed, float* milliseconds) {
	if (s == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		Engine<float>& e = *s->engine;
		const NmfAlgorithm algo = e.config().algorithm;
		const bool onlyW = algo == NmfAlgorithm::GDCLS || algo == NmfAlgorithm::ACLS || algo == NmfAlgorithm::AHCLS;   // GDCLS.h:147-157, AHCLS.h:158-168
		e.synchronize();
		const auto t0 = std::chrono::steady_clock::now();
		switch (static_cast<NmfInitializationMethod>(init_method)) {
		case NmfInitializationMethod::AllRandomValues:
			e.randomW(seed);
			if (!onlyW) e.randomH(seed);
			break;
		case NmfInitializationMethod::MeanColumns:
			e.meanColumnsW(seed);
			if (!onlyW) e.randomH(seed);
			break;
		case NmfInitializationMethod::KMeansAndRandomValues:
			e.kmeansW(seed);
			if (!onlyW) e.randomH(seed + 1);
			break;
		case NmfInitializationMethod::KMeansAndAbsoluteWTV:
			e.kmeansW(seed);
			if (!onlyW) e.hFromWtV(true);
			break;
		case NmfInitializationMethod::KMeansAndNonNegativeWTV:
			e.kmeansW(seed);
			if (!onlyW) e.hFromWtV(false);
			break;
		default: throw EngineError(ResultType::ErrorInvalidArgument, "session_initialize: AllRandomValues, MeanColumns or a k-means method");
		}
		e.finishInitialisation();
		e.synchronize();
		if (milliseconds) *milliseconds = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
	});
}

NMFGPU_EXPORT int nmfgpu_b200_session_get_factors_f32(nmfgpu_b200_session* s, float* w, unsigned ld_w, float* h, unsigned ld_h) {
	if (s == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		const EngineConfig& c = s->engine->config();
		MatrixDescription<float> wd, hd;
		std::memset(&wd, 0, sizeof(wd));
		std::memset(&hd, 0, sizeof(hd));
		wd.rows = c.m; wd.columns = c.k; wd.format = StorageFormat::Dense; wd.dense.values = w; wd.dense.leadingDimension = ld_w;
		hd.rows = c.k; hd.columns = c.n; hd.format = StorageFormat::Dense; hd.dense.values = h; hd.dense.leadingDimension = ld_h;
		s->engine->store(wd, hd);
	});
}

NMFGPU_EXPORT int nmfgpu_b200_session_iterate(nmfgpu_b200_session* s, unsigned iterations) {
	if (s == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { s->engine->iterateNoError(iterations); });
}

NMFGPU_EXPORT int nmfgpu_b200_session_iterate_with_error(nmfgpu_b200_session* s, double* frobenius, double* rmsd) {
	if (s == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		s->engine->iterate(true);
		if (frobenius) *frobenius = s->engine->frobenius();
		if (rmsd) *rmsd = s->engine->rmsd();
	});
}

NMFGPU_EXPORT int nmfgpu_b200_session_time_iterations(nmfgpu_b200_session* s, unsigned iterations, float* milliseconds) {
	if (s == nullptr || milliseconds == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		cudaStream_t st = s->engine->stream();
		CUDA_CHECK(cudaEventRecord(s->start, st));
		s->engine->iterateNoError(iterations);
		CUDA_CHECK(cudaEventRecord(s->stop, st));
		CUDA_CHECK(cudaEventSynchronize(s->stop));
		CUDA_CHECK(cudaEventElapsedTime(milliseconds, s->start, s->stop));
	});
}

// `iterations` iterations the way the reference's run loop issues them (SingleGpuDispatcher.cpp:171-201): the residual is
// evaluated on every 10th iteration and on the last one -- partial sums to the host, sort, combine -- and that cost is
// inside the events, as it is inside the reference's elapsedTime
NMFGPU_EXPORT int nmfgpu_b200_session_time_run(nmfgpu_b200_session* s, unsigned iterations, float* milliseconds, double* frobenius) {
	if (s == nullptr || milliseconds == nullptr || iterations == 0) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		cudaStream_t st = s->engine->stream();
		CUDA_CHECK(cudaEventRecord(s->start, st));
		unsigned it = 1;
		while (it <= iterations) {
			const unsigned nextError = std::min(iterations, (it + 9) / 10 * 10);
			if (nextError > it) {
				s->engine->iterateNoError(nextError - it);
				it = nextError;
			}
			s->engine->iterate(true);
			++it;
		}
		CUDA_CHECK(cudaEventRecord(s->stop, st));
		CUDA_CHECK(cudaEventSynchronize(s->stop));
		CUDA_CHECK(cudaEventElapsedTime(milliseconds, s->start, s->stop));
		if (frobenius) *frobenius = s->engine->frobenius();
	});
}

NMFGPU_EXPORT int nmfgpu_b200_session_products_f32(nmfgpu_b200_session* s, float* wtv, float* vht, float* ms_wtv, float* ms_vht) {
	if (s == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { s->engine->debugProducts(wtv, vht, ms_wtv, ms_vht, s->start, s->stop); });
}

NMFGPU_EXPORT unsigned nmfgpu_b200_plan_segments(unsigned rows_a, unsigned reduce_len, unsigned kp, unsigned sms, unsigned* segments, unsigned capacity,
                                                 unsigned* info5, unsigned char* slots_per_tile, unsigned tile_capacity) {
	if (rows_a == 0 || reduce_len == 0 || kp == 0 || sms == 0 || info5 == nullptr || (capacity != 0 && segments == nullptr)) return 0;
	try {
		return tc::enumerateSegments(rows_a, reduce_len, kp, sms, segments, capacity, info5, slots_per_tile, slots_per_tile ? tile_capacity : 0);
	} catch (const std::exception& e) {   // e.g. more than 255 partial products per tile under a forced chunk count
		errorf("[ERROR] %s\n", e.what());
		return 0;
	}
}

NMFGPU_EXPORT int nmfgpu_b200_session_synchronize(nmfgpu_b200_session* s) {
	if (s == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { s->engine->synchronize(); });
}

NMFGPU_EXPORT int nmfgpu_b200_session_get_info(nmfgpu_b200_session* s, nmfgpu_b200_session_info* info) {
	if (s == nullptr || info == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	std::memset(info, 0, sizeof(*info));
	info->uses_tensor_cores = s->engine->usesTensorCores() ? 1 : 0;
	info->splits_wtv = s->engine->splitsWtV();
	info->splits_vht = s->engine->splitsVHt();
	info->kernel_launches = s->engine->kernelLaunches();
	info->collective_calls = s->engine->config().comm ? s->engine->config().comm->calls() : 0;
	info->ld_v = s->engine->ldV();
	info->ld_w = s->engine->ldW();
	info->ld_h = s->engine->ldH();
	info->row_owners = s->engine->rowBlocks() ? 1 : 0;
	return 0;
}

NMFGPU_EXPORT void nmfgpu_b200_session_destroy(nmfgpu_b200_session* s) { delete s; }

NMFGPU_EXPORT void* nmfgpu_b200_device_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

NMFGPU_EXPORT void nmfgpu_b200_device_free(void* p) {
	if (p) cudaFree(p);
}

NMFGPU_EXPORT int nmfgpu_b200_device_uniform_f32(float* dev, unsigned rows, unsigned cols, size_t ld, unsigned long long seed,
                                                 unsigned long long total_rows, unsigned long long row0, unsigned long long col0) {
	if (dev == nullptr || ld < rows) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] {
		const size_t total = (size_t)rows * cols;
		uniform_kernel<<<(unsigned)((total + 255) / 256), 256>>>(dev, rows, cols, ld, seed, total_rows, row0, col0);
		CUDA_CHECK(cudaGetLastError());
		CUDA_CHECK(cudaDeviceSynchronize());
	});
}

NMFGPU_EXPORT int nmfgpu_b200_device_download(void* host, const void* dev, size_t bytes) {
	if (host == nullptr || dev == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { CUDA_CHECK(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost)); });
}

NMFGPU_EXPORT int nmfgpu_b200_device_upload(void* dev, const void* host, size_t bytes) {
	if (host == nullptr || dev == nullptr) return static_cast<int>(ResultType::ErrorInvalidArgument);
	return guarded([&] { CUDA_CHECK(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice)); });
}

NMFGPU_EXPORT void* nmfgpu_b200_host_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

NMFGPU_EXPORT void nmfgpu_b200_host_free(void* p) {
	if (p) cudaFreeHost(p);
}

NMFGPU_EXPORT int nmfgpu_b200_flush_l2(void) {
	return guarded([&] {
		if (g_flushBuffer == nullptr) CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&g_flushBuffer), kFlushBytes));
		flush_kernel<<<148 * 4, 256>>>(g_flushBuffer, kFlushBytes / sizeof(float4));
		CUDA_CHECK(cudaGetLastError());
		CUDA_CHECK(cudaDeviceSynchronize());
	});
}

}  // extern "C"
