// host.cpp -- logging, summary and the run loop (see host.h).
#include "host.h"

#include <map>
#include <mutex>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <functional>
#include <limits>
#include <random>

namespace nmfgpu {
namespace b200 {

// ---- logging ----------------------------------------------------------------------------------------------
namespace {
Verbosity g_verbosity = Verbosity::Summary;  // reference default: the summary table is printed
}
Verbosity currentVerbosity() { return g_verbosity; }
void setCurrentVerbosity(Verbosity v) { g_verbosity = v; }

void logf(Verbosity level, const char* fmt, ...) {
	if (static_cast<int>(g_verbosity) < static_cast<int>(level)) return;
	va_list ap;
	va_start(ap, fmt);
	vfprintf(stdout, fmt, ap);
	va_end(ap);
	fflush(stdout);
}

void errorf(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
	fflush(stderr);
}

// ---- block pool (common.h) --------------------------------------------------------------------------------------
namespace {
struct BlockPool {
	std::mutex lock;
	std::multimap<std::pair<int, size_t>, void*> device, pinned;   // (device ordinal, bytes) -> free blocks
	bool enabled() {
		static const bool on = [] {
			const char* e = getenv("NMFGPU_POOL");
			return e == nullptr || atoi(e) != 0;
		}();
		return on;
	}
	static int ordinal() {
		int dev = 0;
		cudaGetDevice(&dev);
		return dev;
	}
};
BlockPool& pool() {
	static BlockPool p;
	return p;
}
void* takeBlock(std::multimap<std::pair<int, size_t>, void*>& blocks, size_t bytes) {
	std::lock_guard<std::mutex> guard(pool().lock);
	auto it = blocks.find({BlockPool::ordinal(), bytes});
	if (it == blocks.end()) return nullptr;
	void* p = it->second;
	blocks.erase(it);
	return p;
}
}  // namespace

void* pooledDeviceAlloc(size_t bytes) {
	if (pool().enabled())
		if (void* p = takeBlock(pool().device, bytes)) return p;
	void* p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes);
	if (e == cudaErrorMemoryAllocation) {   // give the pooled blocks back and try once more
		cudaGetLastError();
		releasePooledMemory();
		e = cudaMalloc(&p, bytes);
	}
	CUDA_CHECK(e);
	return p;
}

void pooledDeviceFree(void* p, size_t bytes) {
	if (!pool().enabled()) {
		cudaFree(p);
		return;
	}
	cudaDeviceSynchronize();   // what cudaFree would have done: nothing in flight may still use the block when it is handed out again
	std::lock_guard<std::mutex> guard(pool().lock);
	pool().device.insert({{BlockPool::ordinal(), bytes}, p});
}

void* pooledPinnedAlloc(size_t bytes) {
	if (pool().enabled())
		if (void* p = takeBlock(pool().pinned, bytes)) return p;
	void* p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) {
		cudaGetLastError();
		throw EngineError(ResultType::ErrorNotEnoughHostMemory, "cudaMallocHost failed");
	}
	return p;
}

void pooledPinnedFree(void* p, size_t bytes) {
	if (!pool().enabled()) {
		cudaFreeHost(p);
		return;
	}
	std::lock_guard<std::mutex> guard(pool().lock);
	pool().pinned.insert({{BlockPool::ordinal(), bytes}, p});
}

void releasePooledMemory() {
	std::lock_guard<std::mutex> guard(pool().lock);
	int current = 0;
	cudaGetDevice(&current);
	for (auto& kv : pool().device) {
		cudaSetDevice(kv.first.first);
		cudaFree(kv.second);
	}
	for (auto& kv : pool().pinned) cudaFreeHost(kv.second);
	pool().device.clear();
	pool().pinned.clear();
	cudaSetDevice(current);
	cudaGetLastError();
}

// ---- summary ------------------------------------------------------------------------------------------------
void Summary::destroy() { delete this; }
unsigned Summary::bestRun() const { return m_bestRun; }
void Summary::record(unsigned index, ExecutionRecord& out) const {
	if (index < m_records.size()) out = m_records[index];
}
unsigned Summary::recordCount() const { return static_cast<unsigned>(m_records.size()); }
void Summary::insert(const ExecutionRecord& rec) {
	// the newest record becomes the best one only if it is strictly better than all earlier ones
	const bool better = std::none_of(m_records.begin(), m_records.end(), [&](const ExecutionRecord& r) { return r.frobenius <= rec.frobenius; });
	m_records.push_back(rec);
	if (better) m_bestRun = static_cast<unsigned>(m_records.size() - 1);
}
void Summary::reset() {
	m_bestRun = 0;
	m_records.clear();
}

// ---- run loop -------------------------------------------------------------------------------------------------
namespace {
const char* algorithmName(NmfAlgorithm a) {
	switch (a) {
	case NmfAlgorithm::Multiplicative: return "Multiplicative Frobenius";
	case NmfAlgorithm::GDCLS: return "Gradient Descent Constrained Least Squares";
	case NmfAlgorithm::ALS: return "Alternating Least Squares";
	case NmfAlgorithm::ACLS: return "Alternating Constrained Least Squares";
	case NmfAlgorithm::AHCLS: return "Alternating Hoyer Constrained Least Squares";
	case NmfAlgorithm::nsNMF: return "non-smooth NMF";
	}
	return "?";
}

void formatDuration(char (&buf)[32], long long ms) {
	snprintf(buf, sizeof(buf), "%02d:%02d:%02d.%03d", int(ms / 3600000), int(ms / 60000 % 60), int(ms / 1000 % 60), int(ms % 1000));
}

// per-run initialisation of W and H (reference: IAlgorithm::initialize of each algorithm + InitializationStrategy::create)
template <typename T>
void initialiseRun(NmfDescription<T>& desc, Engine<T>& engine, std::function<unsigned()>& nextSeed) {
	const NmfAlgorithm algo = desc.algorithm;
	const bool onlyW = algo == NmfAlgorithm::GDCLS || algo == NmfAlgorithm::ACLS || algo == NmfAlgorithm::AHCLS;  // GDCLS.h:147-157, AHCLS.h:158-168
	const bool gdclsConstant = algo == NmfAlgorithm::GDCLS && desc.useConstantBasisVectors;
	if (!gdclsConstant) {
		desc.seed = nextSeed();  // written back to the caller's struct (MU.h:137)
		switch (desc.initMethod) {
		case NmfInitializationMethod::CopyExisting:
			engine.loadW(desc.outputMatrixW);
			if (!onlyW) engine.loadH(desc.outputMatrixH);
			break;
		case NmfInitializationMethod::AllRandomValues:
			engine.randomW(desc.seed);
			if (!onlyW) engine.randomH(desc.seed);  // same seed for both factors (RandomValueStrategy.cpp:53-70)
			break;
		case NmfInitializationMethod::MeanColumns:
			engine.meanColumnsW(desc.seed);
			if (!onlyW) engine.randomH(desc.seed);
			break;
		case NmfInitializationMethod::KMeansAndRandomValues:
			engine.kmeansW(desc.seed);
			if (!onlyW) engine.randomH(desc.seed + 1);  // KMeansStrategy.cpp:45
			break;
		case NmfInitializationMethod::KMeansAndAbsoluteWTV:  // the reference leaves H untouched here (B-10); the header documents |W^T V|
			engine.kmeansW(desc.seed);
			if (!onlyW) engine.hFromWtV(true);
			break;
		case NmfInitializationMethod::KMeansAndNonNegativeWTV:
			engine.kmeansW(desc.seed);
			if (!onlyW) engine.hFromWtV(false);
			break;
		case NmfInitializationMethod::EInNMF:
			throw EngineError(ResultType::ErrorInvalidArgument, "EInNMF initialisation is not provided (reference kernel has undefined warp-shuffle behaviour on sm_70+)");
		default:
			throw EngineError(ResultType::ErrorInvalidArgument, "unknown initialisation method");
		}
	}
	if (desc.useConstantBasisVectors) engine.loadW(desc.outputMatrixW);  // MU.h:144-146
	engine.finishInitialisation();
}
}  // namespace

template <typename T>
bool runFactorisation(NmfDescription<T>& desc, Engine<T>& engine, Summary* summary) {
	if (summary != nullptr) summary->reset();
	const unsigned numIterations = desc.numIterations;
	const unsigned numRuns = desc.numRuns;
	const bool multi = numRuns > 1;
	const Context* ctx = currentContext();

	// seed chain: successive outputs of uniform_int<unsigned>(0, UINT_MAX) over mt19937(seed0) (Algorithm.cpp:26-31)
	std::function<unsigned()> nextSeed =
	    std::bind(std::uniform_int_distribution<unsigned>(0, std::numeric_limits<unsigned>::max()), std::mt19937(desc.seed));

	logf(Verbosity::Summary, " Executing %u run(s) of the '%s' algorithm on CUDA device #%d%s: \n", numRuns, algorithmName(desc.algorithm),
	     ctx ? ctx->deviceId : 0, engine.usesTensorCores() ? " (tcgen05 3xTF32)" : " (SIMT)");

	bool interrupted = false;
	double bestError = std::numeric_limits<double>::max();
	const char* rule = multi ? " -------------------------------------------------------------------------------------------------------------\n"
	                         : " ---------------------------------------------------------------------------------------------------\n";
	for (unsigned run = 1; run <= numRuns; ++run) {
		if (run == 1) {
			logf(Verbosity::Summary, "%s", rule);
			if (multi) logf(Verbosity::Summary, " |   Run   | Iteration |     Frobenius     |       RMSD       |       Delta      | Elapsed Time |   Status   |\n");
			else logf(Verbosity::Summary, " | Iteration |     Frobenius     |       RMSD       |       Delta      | Elapsed Time |   Status   |\n");
			logf(Verbosity::Summary, "%s", rule);
		}
		PhaseTimer timer;
		initialiseRun(desc, engine, nextSeed);
		engine.synchronize();
		timer.mark("  initial factors");

		const auto started = std::chrono::high_resolution_clock::now();
		long long elapsedMs = 0;
		double lastError = 0.0, delta = 0.0;
		unsigned iteration = 1;
		for (; iteration <= numIterations; ++iteration) {
			if (desc.callbackUserInterrupt != nullptr && desc.callbackUserInterrupt()) {  // polled before every iteration (Dispatcher.cpp:171)
				interrupted = true;
				break;
			}
			if (desc.callbackUserInterrupt == nullptr) {
				// no callback to poll: the iterations before the next residual evaluation go out as one batch (CUDA graph)
				const unsigned nextError = std::min(numIterations, (iteration + 9) / 10 * 10);
				if (nextError > iteration) {
					engine.iterateNoError(nextError - iteration);
					iteration = nextError;
				}
			}
			const bool computeError = iteration % 10 == 0 || iteration == numIterations;  // SingleGpuDispatcher.cpp:173
			engine.iterate(computeError);
			if (computeError) {
				elapsedMs = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - started).count();
				const double current = desc.thresholdType == NmfThresholdType::Frobenius ? engine.frobenius() : engine.rmsd();
				delta = current - lastError;
				if (currentVerbosity() >= Verbosity::Informative) {
					char t[32];
					formatDuration(t, elapsedMs);
					logf(Verbosity::Informative, " | %9u | %17.4f | %16.4f | %16.4f | %s |            |\n", iteration, engine.frobenius(), engine.rmsd(), delta, t);
				}
				if (lastError != 0.0 && std::fabs(delta) < desc.thresholdValue) break;  // absolute delta, never on the first check
				lastError = current;
			}
		}
		iteration = std::min(iteration, numIterations);  // SingleGpuDispatcher.cpp:205
		timer.mark("  iterations");

		char t[32];
		formatDuration(t, elapsedMs);
		const char* status = "Aborted";
		if (!interrupted) {
			const bool stored = engine.frobenius() < bestError;  // always Frobenius, also when thresholding on RMSD (:214)
			if (stored) {
				ExecutionRecord rec = ExecutionRecord();
				if (summary != nullptr) {
					rec.elapsedTime = elapsedMs / 1000.0;
					rec.frobenius = engine.frobenius();
					rec.rmsd = engine.rmsd();
					rec.numIterations = iteration;
				}
				engine.store(desc.outputMatrixW, desc.outputMatrixH);
				if (summary != nullptr) {
					rec.sparsityW = engine.sparsityW();   // Hoyer sparseness of the stored factors (the reference leaves both at 0)
					rec.sparsityH = engine.sparsityH();
					summary->insert(rec);
				}
				timer.mark("  store W, H");
				bestError = engine.frobenius();
			}
			status = stored ? "Stored" : "Discarded";
		}
		if (multi) logf(Verbosity::Summary, " | %7u | %9u | %17.4f | %16.4f | %16.4f | %s | %10s |\n", run, iteration, engine.frobenius(), engine.rmsd(), delta, t, status);
		else logf(Verbosity::Summary, " | %9u | %17.4f | %16.4f | %16.4f | %s | %10s |\n", iteration, engine.frobenius(), engine.rmsd(), delta, t, status);
		if (interrupted) break;
	}
	logf(Verbosity::Summary, "%s", rule);
	engine.synchronize();
	return !interrupted;
}

template bool runFactorisation<float>(NmfDescription<float>&, Engine<float>&, Summary*);
template bool runFactorisation<double>(NmfDescription<double>&, Engine<double>&, Summary*);

}  // namespace b200
}  // namespace nmfgpu
