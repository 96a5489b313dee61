// api.cpp -- the drop-in boundary: nmfgpu:: C++ entry points and the twelve extern "C" symbols of
// include/nmfgpu.h, with the argument checks and return codes of reference source/common/Interface.cpp.
// Nothing here computes; it validates, builds an Engine (engine.h) and hands over to the run loop (host.cpp).
#include <cstdlib>
#include <cstring>
#include <new>

#include "host.h"
#include "kmeans.h"
#include "sparse.h"

namespace nmfgpu {

namespace b200 {
namespace {
thread_local Context* t_context = nullptr;

// named parameter lookup (reference Interface.cpp:41-49): first exact match wins
bool findParameter(const Parameter* params, unsigned count, const char* name, double& value) {
	if (params == nullptr) return false;
	for (unsigned i = 0; i < count; ++i) {
		if (params[i].name != nullptr && std::strcmp(params[i].name, name) == 0) {
			value = params[i].value;
			return true;
		}
	}
	return false;
}

bool requireParameter(const Parameter* params, unsigned count, const char* algorithm, const char* name, double& value) {
	if (findParameter(params, count, name, value)) return true;
	errorf("[ERROR] %s algorithm requires parameter '%s' to be set!\n", algorithm, name);
	return false;
}

Precision precisionFromEnvironment() {
	const char* e = std::getenv("NMFGPU_PRECISION");
	if (e == nullptr) return Precision::Auto;
	if (!std::strcmp(e, "fp32") || !std::strcmp(e, "exact")) return Precision::Exact;
	if (!std::strcmp(e, "3xtf32")) return Precision::Tf32x3;
	if (!std::strcmp(e, "tf32")) return Precision::Tf32x1;
	return Precision::Auto;
}

template <typename T>
ResultType computeImpl(NmfDescription<T>& desc, ISummary* summary) {
	Context* ctx = t_context;
	if (ctx == nullptr) return ResultType::ErrorNotInitialized;

	// Interface.cpp:221-225 -- CopyExisting has no randomisation, more than one run is pointless (mutates the caller's struct)
	if (desc.initMethod == NmfInitializationMethod::CopyExisting && desc.numRuns > 1) {
		logf(Verbosity::Summary, "[WARNING] When using the CopyExisting initialization method, then no more than one run should be performed because of missing randomization!\n");
		desc.numRuns = 1;
	}
	// Interface.cpp:228-232 (columns = all columns of the data set, also when this rank holds a shard)
	const unsigned globalColumns = (ctx->comm && ctx->comm->worldSize() > 1) ? ctx->comm->globalColumns() : desc.inputMatrix.columns;
	if (!desc.useConstantBasisVectors && desc.features > globalColumns) {
		errorf("[ERROR] Feature count has to be less than the matrix dimensions!\n");
		return ResultType::ErrorInvalidArgument;
	}
	if (desc.features == 0 || desc.inputMatrix.rows == 0 || desc.inputMatrix.columns == 0) return ResultType::ErrorInvalidArgument;
	if (desc.outputMatrixW.format != StorageFormat::Dense || desc.outputMatrixH.format != StorageFormat::Dense) {
		errorf("[ERROR] Output matrices must have a dense storage format!\n");
		return ResultType::ErrorInvalidArgument;
	}

	EngineConfig cfg;
	cfg.algorithm = desc.algorithm;
	cfg.m = desc.inputMatrix.rows;
	cfg.n = desc.inputMatrix.columns;
	cfg.k = desc.features;
	cfg.constantW = desc.useConstantBasisVectors;
	cfg.precision = ctx->precision;
	cfg.comm = (ctx->comm && ctx->comm->worldSize() > 1) ? ctx->comm.get() : nullptr;
	cfg.needsDenseV = desc.initMethod == NmfInitializationMethod::MeanColumns || desc.initMethod == NmfInitializationMethod::KMeansAndRandomValues ||
	                  desc.initMethod == NmfInitializationMethod::KMeansAndAbsoluteWTV || desc.initMethod == NmfInitializationMethod::KMeansAndNonNegativeWTV;

	// required named parameters per algorithm (Interface.cpp:237-336)
	const Parameter* p = desc.parameters;
	const unsigned np = desc.numParameters;
	switch (desc.algorithm) {
	case NmfAlgorithm::Multiplicative:
	case NmfAlgorithm::ALS: break;
	case NmfAlgorithm::ACLS:
		if (!requireParameter(p, np, "ACLS", "lambdaW", cfg.params.lambdaW) || !requireParameter(p, np, "ACLS", "lambdaH", cfg.params.lambdaH))
			return ResultType::ErrorInvalidArgument;
		break;
	case NmfAlgorithm::AHCLS:
		if (!requireParameter(p, np, "AHCLS", "lambdaW", cfg.params.lambdaW) || !requireParameter(p, np, "AHCLS", "lambdaH", cfg.params.lambdaH) ||
		    !requireParameter(p, np, "AHCLS", "alphaW", cfg.params.alphaW) || !requireParameter(p, np, "AHCLS", "alphaH", cfg.params.alphaH))
			return ResultType::ErrorInvalidArgument;
		break;
	case NmfAlgorithm::GDCLS:
		if (!requireParameter(p, np, "GDCLS", "lambda", cfg.params.lambda)) return ResultType::ErrorInvalidArgument;
		break;
	case NmfAlgorithm::nsNMF:
		if (!requireParameter(p, np, "nsNMF", "theta", cfg.params.theta)) return ResultType::ErrorInvalidArgument;
		break;
	default:
		errorf("[ERROR] Chosen algorithm is not implemented!\n");
		return ResultType::ErrorInvalidArgument;
	}
	if (static_cast<int>(desc.initMethod) < 0 || static_cast<int>(desc.initMethod) > static_cast<int>(NmfInitializationMethod::EInNMF))
		return ResultType::ErrorInvalidArgument;

	try {
		if (cudaSetDevice(ctx->deviceId) != cudaSuccess) {
			cudaGetLastError();
			errorf("[ERROR] No usable CUDA device (#%d): the NMF engine has no CPU fallback.\n", ctx->deviceId);
			return ResultType::ErrorDeviceSelection;
		}
		PhaseTimer timer;
		bool finished = false;
		{
			Engine<T> engine(cfg);
			engine.setup(desc.inputMatrix, false);
			timer.mark("setup (alloc, H2D, plans)");
			finished = runFactorisation<T>(desc, engine, static_cast<Summary*>(summary));
			timer.mark("runs (init, loop, store)");
		}
		timer.mark("teardown");
		return finished ? ResultType::Success : ResultType::ErrorUserInterrupt;
	} catch (const EngineError& e) {
		errorf("[ERROR] %s\n", e.what());
		return e.code;
	} catch (const std::bad_alloc&) {
		return ResultType::ErrorNotEnoughHostMemory;
	} catch (const std::exception& e) {
		errorf("[ERROR] %s\n", e.what());
		return ResultType::ErrorExternalLibrary;
	}
}

template <typename T>
ResultType computeKMeansImpl(KMeansDescription<T>& desc, KMeansSummary* /*summary: accepted and ignored, as in Interface.cpp:416-418*/) {
	Context* ctx = t_context;
	if (ctx == nullptr) return ResultType::ErrorNotInitialized;
	// Interface.cpp:365-389
	// column shards (include/nmfgpu_b200.h): the samples of this rank are a slice of the global sample set
	Communicator* comm = (ctx->comm && ctx->comm->worldSize() > 1) ? ctx->comm.get() : nullptr;
	const unsigned globalSamples = comm != nullptr ? comm->globalColumns() : desc.inputMatrix.columns;
	if (desc.numClusters == 0 || desc.numClusters >= globalSamples) {
		errorf(" [ERROR] Number of clusters must be smaller than number of samples in dataset!\n");
		return ResultType::ErrorInvalidArgument;
	}
	if (desc.outputMatrixClusters.format != StorageFormat::Dense) {
		errorf(" [ERROR] Cluster matrix must have a dense storage format!\n");
		return ResultType::ErrorInvalidArgument;
	}
	if (desc.inputMatrix.rows != desc.outputMatrixClusters.rows) {
		errorf(" [ERROR] Input and output matrices must have the same amount of rows!\n");
		return ResultType::ErrorInvalidArgument;
	}
	if (desc.inputMatrix.rows == 0 || desc.outputMatrixClusters.dense.values == nullptr) return ResultType::ErrorInvalidArgument;
	try {
		if (cudaSetDevice(ctx->deviceId) != cudaSuccess) {
			cudaGetLastError();
			errorf("[ERROR] No usable CUDA device (#%d): the k-means engine has no CPU fallback.\n", ctx->deviceId);
			return ResultType::ErrorDeviceSelection;
		}
		const unsigned m = desc.inputMatrix.rows, n = desc.inputMatrix.columns, k = desc.numClusters;
		const size_t ld = roundUp(m, 32);
		cudaStream_t stream;
		CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		struct StreamGuard {
			cudaStream_t s;
			~StreamGuard() { cudaStreamDestroy(s); }
		} guard{stream};
		DeviceBuffer<T> data, centroids;
		DeviceBuffer<unsigned> membership;
		data.allocate(ld * n);
		centroids.allocate(ld * k);
		membership.allocate(n);
		if (desc.inputMatrix.format == StorageFormat::Dense) {
			if (desc.inputMatrix.dense.values == nullptr || desc.inputMatrix.dense.leadingDimension < m) return ResultType::ErrorInvalidArgument;
			CUDA_CHECK(cudaMemcpy2DAsync(data.get(), ld * sizeof(T), desc.inputMatrix.dense.values, (size_t)desc.inputMatrix.dense.leadingDimension * sizeof(T),
			                             (size_t)m * sizeof(T), n, cudaMemcpyHostToDevice, stream));
		} else {
			sparse::densify(desc.inputMatrix, data.get(), ld, stream);
		}
		kmeans::run<T>(m, n, k, data.get(), ld, centroids.get(), ld, membership.get(), desc.seed, desc.numIterations, desc.thresholdValue, stream, comm);
		CUDA_CHECK(cudaMemcpy2DAsync(desc.outputMatrixClusters.dense.values, (size_t)desc.outputMatrixClusters.dense.leadingDimension * sizeof(T), centroids.get(),
		                             ld * sizeof(T), (size_t)m * sizeof(T), k, cudaMemcpyDeviceToHost, stream));
		if (desc.outputMemberships != nullptr)
			CUDA_CHECK(cudaMemcpyAsync(desc.outputMemberships, membership.get(), (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		return ResultType::Success;
	} catch (const EngineError& e) {
		errorf("[ERROR] %s\n", e.what());
		return e.code;
	} catch (const std::bad_alloc&) {
		return ResultType::ErrorNotEnoughHostMemory;
	}
}
}  // namespace

Context* currentContext() { return t_context; }
}  // namespace b200

using namespace b200;

// ---- C++ API (reference Interface.cpp:53-77,144-211,350-352,424-430) ------------------------------------------

NMFGPU_EXPORT ResultType initialize() {
	if (t_context != nullptr) return ResultType::ErrorAlreadyInitialized;
	Context* ctx = new (std::nothrow) Context();
	if (ctx == nullptr) return ResultType::ErrorNotEnoughHostMemory;
	ctx->precision = precisionFromEnvironment();
	int dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess) ctx->deviceId = dev;
	else cudaGetLastError();
	t_context = ctx;
	return ResultType::Success;
}

NMFGPU_EXPORT ResultType finalize() {
	if (t_context == nullptr) return ResultType::ErrorNotInitialized;
	delete t_context;
	t_context = nullptr;
	releasePooledMemory();   // device and pinned blocks kept between calls (common.h) go back to the driver
	return ResultType::Success;
}

NMFGPU_EXPORT int version() { return NMFGPU_VERSION; }

NMFGPU_EXPORT ResultType chooseGpu(unsigned index) {
	if (t_context == nullptr) return ResultType::ErrorNotInitialized;  // the reference dereferences null here
	if (cudaSetDevice(int(index)) != cudaSuccess) {
		cudaGetLastError();
		return ResultType::ErrorDeviceSelection;
	}
	t_context->deviceId = int(index);
	return ResultType::Success;
}

NMFGPU_EXPORT unsigned getNumberOfGpu() {
	int num = 0;
	if (cudaGetDeviceCount(&num) != cudaSuccess) {
		cudaGetLastError();
		return 0u;
	}
	return static_cast<unsigned>(num);
}

NMFGPU_EXPORT ResultType getInformationForGpuIndex(unsigned index, GpuInformation& info) {
	int previous = 0;
	if (cudaGetDevice(&previous) != cudaSuccess || cudaSetDevice(int(index)) != cudaSuccess) {
		cudaGetLastError();
		return ResultType::ErrorDeviceSelection;
	}
	cudaDeviceProp props;
	if (cudaGetDeviceProperties(&props, int(index)) == cudaSuccess) {
		std::strncpy(info.name, props.name, sizeof(info.name) - 1);
		info.name[sizeof(info.name) - 1] = '\0';
	} else {
		std::strcpy(info.name, "N/A");
	}
	const cudaError_t st = cudaMemGetInfo(&info.freeMemory, &info.totalMemory);
	cudaSetDevice(previous);
	if (st != cudaSuccess) {
		cudaGetLastError();
		info.freeMemory = 0;
		info.totalMemory = 0;
		return ResultType::ErrorExternalLibrary;
	}
	return ResultType::Success;
}

NMFGPU_EXPORT void setVerbosity(Verbosity verbosity) { setCurrentVerbosity(verbosity); }

ISummary* ISummary::create() { return new Summary(); }

NMFGPU_EXPORT ResultType compute(NmfDescription<float>& description, ISummary* summary) { return computeImpl(description, summary); }
NMFGPU_EXPORT ResultType compute(NmfDescription<double>& description, ISummary* summary) { return computeImpl(description, summary); }
NMFGPU_EXPORT ResultType computeKMeans(KMeansDescription<float>& desc, KMeansSummary* summary) { return computeKMeansImpl(desc, summary); }
NMFGPU_EXPORT ResultType computeKMeans(KMeansDescription<double>& desc, KMeansSummary* summary) { return computeKMeansImpl(desc, summary); }

}  // namespace nmfgpu

// ---- C ABI (reference Interface.cpp:434-488) ----------------------------------------------------------------------
extern "C" {

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_initialize() { return nmfgpu::initialize(); }
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_finalize() { return nmfgpu::finalize(); }
NMFGPU_EXPORT int nmfgpu_version() { return nmfgpu::version(); }
NMFGPU_EXPORT void nmfgpu_set_verbosity(nmfgpu::Verbosity verbosity) { nmfgpu::setVerbosity(verbosity); }

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_create_summary(nmfgpu::ISummary** summary) {
	if (summary == nullptr) return nmfgpu::ResultType::ErrorInvalidArgument;
	*summary = nmfgpu::ISummary::create();
	return nmfgpu::ResultType::Success;
}

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_single(nmfgpu::NmfDescription<float>* description, nmfgpu::ISummary* summary) {
	if (description == nullptr) return nmfgpu::ResultType::ErrorInvalidArgument;
	return nmfgpu::compute(*description, summary);
}

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_double(nmfgpu::NmfDescription<double>* description, nmfgpu::ISummary* summary) {
	if (description == nullptr) return nmfgpu::ResultType::ErrorInvalidArgument;
	return nmfgpu::compute(*description, summary);
}

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_kmeans_single(nmfgpu::KMeansDescription<float>* desc) {
	if (desc == nullptr) return nmfgpu::ResultType::ErrorInvalidArgument;
	return nmfgpu::computeKMeans(*desc, nullptr);
}

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_kmeans_double(nmfgpu::KMeansDescription<double>* desc) {
	if (desc == nullptr) return nmfgpu::ResultType::ErrorInvalidArgument;
	return nmfgpu::computeKMeans(*desc, nullptr);
}

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_choose_gpu(unsigned index) { return nmfgpu::chooseGpu(index); }
NMFGPU_EXPORT unsigned nmfgpu_get_number_of_gpu() { return nmfgpu::getNumberOfGpu(); }

NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_get_information_for_gpu_index(unsigned index, nmfgpu::GpuInformation* info) {
	if (info == nullptr) return nmfgpu::ResultType::ErrorInvalidArgument;
	return nmfgpu::getInformationForGpuIndex(index, *info);
}

}  // extern "C"
