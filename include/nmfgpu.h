// nmfgpu.h -- public boundary of the B200-native NMF engine (libnmfgpu64.so).
//
// This header is the drop-in boundary: it re-declares, field for field and
// enumerator for enumerator, the interface that razorx89/nmfgpu publishes in
// its include/nmfgpu.h (reference lines cited per declaration below) so that
// a caller compiled against the reference header (nmfgpu4R, example/main.cpp)
// links against this library unchanged.  Nothing behind the boundary is shared
// with the reference: the implementation lives in nmfgpu_b200/csrc/.
//
// ABI facts that tests/test_abi.py pins:
//   * every struct is laid out under #pragma pack(4)  (reference nmfgpu.h:47)
//   * enums are `enum class` => int sized, values are positional
//   * ISummary is a vtable ABI: destroy, bestRun, record, recordCount, dtor
//   * twelve extern "C" entry points, see the bottom of this file
#pragma once

#include <cstddef>

// reference nmfgpu.h:28-31
#define NMFGPU_MAJOR 0
#define NMFGPU_MINOR 2
#define NMFGPU_PATCH 3
#define NMFGPU_VERSION ((NMFGPU_MAJOR << 24) | (NMFGPU_MINOR << 16) | NMFGPU_PATCH)

// Linux-only build: symbols are exported through default visibility.
#if defined(NMFGPU_EXPORTING)
#define NMFGPU_EXPORT __attribute__((visibility("default")))
#else
#define NMFGPU_EXPORT
#endif

#pragma pack(push, 4)

namespace nmfgpu {

// ---- enumerations (values are part of the ABI) -----------------------------

// reference nmfgpu.h:52-77
enum class ResultType {
	Success = 0,
	ErrorAlreadyInitialized,     // 1: initialize() twice on one thread
	ErrorNotInitialized,         // 2: any compute before initialize()
	ErrorInvalidArgument,        // 3
	ErrorNotEnoughHostMemory,    // 4
	ErrorNotEnoughDeviceMemory,  // 5
	ErrorExternalLibrary,        // 6: CUDA / NCCL failure
	ErrorUserInterrupt,          // 7: callbackUserInterrupt returned true
	ErrorDeviceSelection,        // 8
};

// reference nmfgpu.h:80-100
enum class NmfInitializationMethod {
	CopyExisting,             // W,H taken from outputMatrixW / outputMatrixH
	AllRandomValues,          // uniform (0,1]
	MeanColumns,              // W column = mean of 5 random data columns, H random
	KMeansAndRandomValues,    // W = k-means centroids, H random
	KMeansAndAbsoluteWTV,     // W = centroids, H = |W^T V|
	KMeansAndNonNegativeWTV,  // W = centroids, H = max(0, W^T V)
	EInNMF,                   // W = centroids, H = fuzzy membership degrees
};

// reference nmfgpu.h:102-105
enum class NmfThresholdType { Frobenius, RMSD };

// reference nmfgpu.h:107-114
enum class NmfAlgorithm { Multiplicative, GDCLS, ALS, ACLS, AHCLS, nsNMF };

// reference nmfgpu.h:117-126
enum class Verbosity { None, Summary, Informative, Debugging };

// reference nmfgpu.h:129-134
enum class IndexBase { Zero, One };

// reference nmfgpu.h:177-186
enum class StorageFormat { Dense, CSR, CSC, COO };

// ---- result records ---------------------------------------------------------

// reference nmfgpu.h:136
typedef bool (*UserInterruptCallback)();

// reference nmfgpu.h:138-147 (44 bytes under pack(4))
struct ExecutionStatistic {
	double frobenius;
	double rmsd;
	double elapsedTime;  // seconds, host clock, from end of init to last error check
	double sparsityW;
	double sparsityH;
	unsigned numIterations;
};
typedef ExecutionStatistic ExecutionRecord;

// reference nmfgpu.h:149-175.  Virtual order is ABI.
class ISummary {
public:
	NMFGPU_EXPORT static ISummary* create();
	virtual void destroy() = 0;
	virtual unsigned bestRun() const = 0;
	virtual void record(unsigned index, ExecutionRecord& record) const = 0;
	virtual unsigned recordCount() const = 0;

protected:
	virtual ~ISummary() {}
};

// ---- matrices ----------------------------------------------------------------

// reference nmfgpu.h:188-232 (44 bytes).  Dense is column-major with an
// explicit leading dimension; the three sparse members share one shape.
template <typename NumericType>
struct MatrixDescription {
	unsigned rows;
	unsigned columns;
	StorageFormat format;
	union {
		struct {
			NumericType* values;
			unsigned leadingDimension;
		} dense;
		struct {
			NumericType* values;
			int* rowPtr;
			int* columnIndices;
			unsigned nnz;
			IndexBase base;
		} csr;
		struct {
			NumericType* values;
			int* columnPtr;
			int* rowIndices;
			unsigned nnz;
			IndexBase base;
		} csc;
		struct {
			NumericType* values;
			int* rowIndices;
			int* columnIndices;
			unsigned nnz;
			IndexBase base;
		} coo;
	};
};

// reference nmfgpu.h:234-237
struct Parameter {
	const char* name;
	double value;
};

// reference nmfgpu.h:239-273 (200 bytes).  Attributes are rows of inputMatrix,
// samples are columns.  The library writes back numRuns and seed.
template <typename NumericType>
struct NmfDescription {
	NmfAlgorithm algorithm;
	bool useConstantBasisVectors;
	MatrixDescription<NumericType> inputMatrix;
	int* inputLabels;
	MatrixDescription<NumericType> outputMatrixW;
	MatrixDescription<NumericType> outputMatrixH;
	unsigned features;
	NmfInitializationMethod initMethod;
	unsigned numIterations;
	unsigned numRuns;
	unsigned seed;
	NmfThresholdType thresholdType;
	double thresholdValue;
	UserInterruptCallback callbackUserInterrupt;
	Parameter* parameters;
	unsigned numParameters;
};

// reference nmfgpu.h:287-291
struct GpuInformation {
	char name[256];
	size_t totalMemory;
	size_t freeMemory;
};

// reference nmfgpu.h:301-310
template <typename NumericType>
struct KMeansDescription {
	MatrixDescription<NumericType> inputMatrix;
	MatrixDescription<NumericType> outputMatrixClusters;
	unsigned* outputMemberships;
	unsigned numClusters;
	unsigned numIterations;
	unsigned seed;
	double thresholdValue;
};

// reference nmfgpu.h:312-327
struct KMeansSummary {
	unsigned iterations;
	double betweenSS;
	double* withinSS;
	double totalWithinSS;
	double totalSS;
};

// ---- C++ entry points (Itanium mangled; reference nmfgpu.h:276-299,329-330) --

NMFGPU_EXPORT ResultType initialize();
NMFGPU_EXPORT ResultType finalize();
NMFGPU_EXPORT int version();
NMFGPU_EXPORT ResultType chooseGpu(unsigned index);
NMFGPU_EXPORT unsigned getNumberOfGpu();
NMFGPU_EXPORT ResultType getInformationForGpuIndex(unsigned index, GpuInformation& info);
NMFGPU_EXPORT void setVerbosity(Verbosity verbosity);
NMFGPU_EXPORT ResultType compute(NmfDescription<float>& description, ISummary* summary);
NMFGPU_EXPORT ResultType compute(NmfDescription<double>& description, ISummary* summary);
NMFGPU_EXPORT ResultType computeKMeans(KMeansDescription<float>& desc, KMeansSummary* summary);
NMFGPU_EXPORT ResultType computeKMeans(KMeansDescription<double>& desc, KMeansSummary* summary);

}  // namespace nmfgpu

// ---- C ABI (reference nmfgpu.h:333-349; implemented in csrc/api.cpp) ----------
extern "C" {
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_initialize();
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_finalize();
NMFGPU_EXPORT int nmfgpu_version();
NMFGPU_EXPORT void nmfgpu_set_verbosity(nmfgpu::Verbosity verbosity);
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_create_summary(nmfgpu::ISummary** summary);
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_single(nmfgpu::NmfDescription<float>* description, nmfgpu::ISummary* summary);
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_double(nmfgpu::NmfDescription<double>* description, nmfgpu::ISummary* summary);
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_kmeans_single(nmfgpu::KMeansDescription<float>* desc);
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_compute_kmeans_double(nmfgpu::KMeansDescription<double>* desc);
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_choose_gpu(unsigned index);
NMFGPU_EXPORT unsigned nmfgpu_get_number_of_gpu();
NMFGPU_EXPORT nmfgpu::ResultType nmfgpu_get_information_for_gpu_index(unsigned index, nmfgpu::GpuInformation* info);
}

#pragma pack(pop)
