/* nmfgpu_b200.h -- C-ABI extensions of libnmfgpu64.so that the reference interface has no equivalent for.
 *
 * The drop-in boundary is include/nmfgpu.h (the reference's own twelve entry points).  The functions here
 * add what a B200 deployment needs on top of it and what bench.py / the parity tests use:
 *   - precision selection (3xTF32 tensor cores vs exact SIMT fp32),
 *   - column-sharded multi-GPU execution, one process per GPU (reference: single GPU only,
 *     source/nmf/SingleGpuDispatcher.h:36),
 *   - a "session": a factorisation whose input matrix stays resident in HBM across calls, so iterations
 *     can be timed without the host<->device copies that nmfgpu_compute_single performs per call
 *     (reference: H2D of V inside every compute(), source/nmf/AlgorithmMultiplicativeFrobenius.h:118).
 * Plain C types only; every function returns an nmfgpu::ResultType value as int (0 = success).
 */
#pragma once
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nmfgpu_b200_session nmfgpu_b200_session;

typedef struct nmfgpu_b200_named_value {
	const char* name; /* "lambda", "lambdaW", "lambdaH", "alphaW", "alphaH", "theta" (reference Interface.cpp:246-328) */
	double value;
} nmfgpu_b200_named_value;

typedef struct nmfgpu_b200_session_info {
	int uses_tensor_cores;            /* 1 when the tcgen05 3xTF32 kernels run the V-sized products */
	unsigned splits_wtv, splits_vht;  /* reduction slices of the two products */
	unsigned long long kernel_launches;   /* kernels launched by this session so far */
	unsigned long long collective_calls;  /* NCCL all-reduces issued so far */
	size_t ld_v, ld_w, ld_h;
	int row_owners;                   /* 1 when column shards were regrouped into row blocks (peer-store exchange, csrc/dist.h), 0 for all-reduce */
} nmfgpu_b200_session_info;

/* 0 = auto (tensor cores when the shape allows), 1 = exact SIMT fp32, 2 = force 3xTF32, 3 = 1xTF32 (diagnostic).
 * Applies to the calling thread's context; also settable with the environment variable
 * NMFGPU_PRECISION = auto | fp32 | 3xtf32 | tf32 read by nmfgpu_initialize(). */
int nmfgpu_b200_set_precision(int mode);

/* ---- multi-GPU: rank 0 creates the id, every rank passes it to dist_init after nmfgpu_initialize() and
 * nmfgpu_choose_gpu().  From then on inputMatrix / outputMatrixH of nmfgpu_compute_* describe this rank's
 * column shard (columnOffset .. columnOffset + inputMatrix.columns of globalColumns); W is replicated on return.
 * MU on the tensor-core path regroups the shards once into row blocks; per iteration the ranks then exchange only k x n
 * partial products, columns of H and k*k + k statistics, as NVLink peer stores issued by the kernels themselves (no NCCL
 * call in the iteration; csrc/dist.h, csrc/fused.h).  Everything else all-reduces V H^T and H H^T with NCCL;
 * NMFGPU_DIST_MODE=allreduce forces that dataflow for MU as well.
 * dist_unique_id: one process per GPU (NCCL + CUDA IPC).  dist_local_unique_id: the ranks are THREADS of the calling
 * process (each with its own nmfgpu_initialize), on any devices -- also all on one, which is how the tests run the sharded
 * dataflows on a single-GPU box. */
int nmfgpu_b200_dist_unique_id(void* out128);
int nmfgpu_b200_dist_local_unique_id(void* out128);
int nmfgpu_b200_dist_init(int rank, int world_size, const void* unique_id128);
int nmfgpu_b200_dist_set_shard(unsigned global_columns, unsigned column_offset);
int nmfgpu_b200_dist_finalize(void);

/* ---- sessions (fp32) */
int nmfgpu_b200_session_create_f32(int algorithm, unsigned rows, unsigned columns, unsigned features, const float* v,
                                   unsigned ld_v, int v_on_device, int constant_w, const nmfgpu_b200_named_value* params,
                                   unsigned num_params, nmfgpu_b200_session** out);
int nmfgpu_b200_session_set_factors_f32(nmfgpu_b200_session* s, const float* w, unsigned ld_w, const float* h, unsigned ld_h);
/* initial factors by one of the reference's strategies on the resident input (init_method: NmfInitializationMethod
 * AllRandomValues, MeanColumns or a KMeans* value; seed: what the run loop's seed chain would pass); reports the wall-clock
 * milliseconds of the initialisation (NULL to skip).  Column shards: k-means runs over all ranks' columns. */
int nmfgpu_b200_session_initialize(nmfgpu_b200_session* s, int init_method, unsigned seed, float* milliseconds);
int nmfgpu_b200_session_get_factors_f32(nmfgpu_b200_session* s, float* w, unsigned ld_w, float* h, unsigned ld_h);
/* enqueue `iterations` iterations without residual evaluation; returns immediately */
int nmfgpu_b200_session_iterate(nmfgpu_b200_session* s, unsigned iterations);
/* one iteration with the residual evaluated the way the reference reports it */
int nmfgpu_b200_session_iterate_with_error(nmfgpu_b200_session* s, double* frobenius, double* rmsd);
/* `iterations` iterations bracketed by CUDA events on the session's stream; milliseconds for all of them */
int nmfgpu_b200_session_time_iterations(nmfgpu_b200_session* s, unsigned iterations, float* milliseconds);
/* `iterations` iterations issued the way the reference's run loop does (SingleGpuDispatcher.cpp:171-201): residual on every
 * 10th and on the last iteration, its host work inside the events; *frobenius (may be NULL) = the last residual */
int nmfgpu_b200_session_time_run(nmfgpu_b200_session* s, unsigned iterations, float* milliseconds, double* frobenius);
/* the two V-sized products for the current factors, summed over their slices, copied to the host:
 * wtv is features x columns (ld = features), vht is rows x features (ld = rows).  Either may be NULL.
 * Also reports the device time of each product in milliseconds (NULL to skip). */
int nmfgpu_b200_session_products_f32(nmfgpu_b200_session* s, float* wtv, float* vht, float* ms_wtv, float* ms_vht);
int nmfgpu_b200_session_synchronize(nmfgpu_b200_session* s);
int nmfgpu_b200_session_get_info(nmfgpu_b200_session* s, nmfgpu_b200_session_info* info);
void nmfgpu_b200_session_destroy(nmfgpu_b200_session* s);

/* ---- diagnostics: the stream-K work split of one tensor-core product (csrc/tc_gemm.h) for `sms` CTAs, computed on the host
 * (no device needed).  A rows = rows_a, reduction length = reduce_len, kp = rank padded to 16.  Writes up to `capacity`
 * segments {cta, 256-row tile, first stage, stages, slot} (5 words each) in execution order and returns their number;
 * info5 = {tiles, stages per tile, chunks, stages per chunk, grid}; slots_per_tile[t] = partial products of 128-wide tile t. */
unsigned nmfgpu_b200_plan_segments(unsigned rows_a, unsigned reduce_len, unsigned kp, unsigned sms, unsigned* segments,
                                   unsigned capacity, unsigned* info5, unsigned char* slots_per_tile, unsigned tile_capacity);

/* ---- device helpers for synthetic workloads (bench.py): no H2D of the 4 GB input */
void* nmfgpu_b200_device_alloc(size_t bytes);
void nmfgpu_b200_device_free(void* p);
/* u[i, j] = f(seed, (col0 + j) * total_rows + row0 + i), the generator of nmfgpu_b200/workloads.py */
int nmfgpu_b200_device_uniform_f32(float* dev, unsigned rows, unsigned cols, size_t ld, unsigned long long seed,
                                   unsigned long long total_rows, unsigned long long row0, unsigned long long col0);
/* plain copies between a host buffer and a device buffer obtained from nmfgpu_b200_device_alloc */
int nmfgpu_b200_device_download(void* host, const void* dev, size_t bytes);
int nmfgpu_b200_device_upload(void* dev, const void* host, size_t bytes);
/* page-locked host memory (so nmfgpu_compute_single's H2D of the input runs at full PCIe speed) */
void* nmfgpu_b200_host_alloc(size_t bytes);
void nmfgpu_b200_host_free(void* p);
/* overwrite a 256 MB scratch buffer (twice the L2) to flush the cache between timed iterations */
int nmfgpu_b200_flush_l2(void);

#ifdef __cplusplus
}
#endif
