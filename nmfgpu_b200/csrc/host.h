// host.h -- host-side pieces behind the C ABI: the per-thread context, the run summary and the run loop.
#pragma once

#include <memory>
#include <vector>

#include "common.h"
#include "dist.h"
#include "engine.h"

namespace nmfgpu {
namespace b200 {

// Per-thread library context.  The reference keeps its cuBLAS/cuSPARSE/cuSOLVER handles in a __thread
// pointer (source/common/Interface.h:33,44): initialize() is per calling thread, and so it is here.
struct Context {
	int deviceId = 0;
	Precision precision = Precision::Auto;
	std::unique_ptr<Communicator> comm;
};
Context* currentContext();

// ISummary implementation; same observable behaviour as reference source/nmf/Summary.cpp:27-60
class Summary : public ISummary {
	std::vector<ExecutionRecord> m_records;
	unsigned m_bestRun = 0;

public:
	void destroy() override;
	unsigned bestRun() const override;
	void record(unsigned index, ExecutionRecord& record) const override;
	unsigned recordCount() const override;
	void insert(const ExecutionRecord& record);
	void reset();
};

// Run loop (reference source/nmf/SingleGpuDispatcher.cpp:132-241): runs x iterations, residual every 10th
// and on the last iteration, absolute-delta stop rule, best run kept.  Returns false on user interrupt.
template <typename T>
bool runFactorisation(NmfDescription<T>& desc, Engine<T>& engine, Summary* summary);

}  // namespace b200
}  // namespace nmfgpu
