// nmf_oracle.cpp -- CPU restatement of the nmfgpu iteration hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under nmfgpu_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker.
//
// What it is: a plain multithreaded C++ (OpenMP) fp64 restatement of the
// algorithms the reference runs on the GPU.  The reference itself has no CPU
// path and no tests or golden vectors, and its
// contractions live in cuBLAS/cuSOLVER (un-vendored, unpinned), so this oracle
// is pinned in two ways only:
//   (1) tests/test_oracle.py: against an independent numpy fp64 restatement
//       and against size-independent identities (trace residual == explicit
//       ||V - W H||_F, unit column norms, monotone MU residual);
//   (2) tests/test_parity_gpu.py: against the reference itself, compiled from
//       /root/reference into oracle/_ref/ (see oracle/build_ref.sh) and run on
//       the B200 from the same CopyExisting W0/H0.
//
// Every function cites the reference file:line (relative to /root/reference/)
// whose semantics it follows.  All matrices are column-major.
//
// Storage: V may be float or double (the 4 GB benchmark matrix is kept in
// fp32); W, H and every intermediate are double.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::vector<double> dvec;

enum Algo { kMU = 0, kGDCLS = 1, kALS = 2, kACLS = 3, kAHCLS = 4, kNsNMF = 5 };  // include/nmfgpu.h NmfAlgorithm

// ---------------------------------------------------------------------------
// dense helpers (column-major, explicit leading dimensions)
// ---------------------------------------------------------------------------

// N (k x n, ld k) = W^T V.   W given row-major packed Wr[m][k] so the inner loop
// over the k features is contiguous.  Replaces cublasXgemm(T,N) at
// source/nmf/AlgorithmMultiplicativeFrobenius.h:187-188.
template <typename T>
void gemm_wt_v(int m, int n, int k, const double* Wr, const T* V, long ldV, double* N) {
#pragma omp parallel for schedule(dynamic, 4)
	for (int j = 0; j < n; ++j) {
		double* acc = N + (size_t)j * k;
		for (int c = 0; c < k; ++c) acc[c] = 0.0;
		const T* v = V + (size_t)j * ldV;
		for (int i = 0; i < m; ++i) {
			const double x = (double)v[i];
			const double* w = Wr + (size_t)i * k;
			for (int c = 0; c < k; ++c) acc[c] += x * w[c];
		}
	}
}

// P (m x k, ld m) = V H^T.  Row blocks keep the m x k accumulator slice in
// cache while all n columns stream past.  Replaces cublasXgemm(N,T) at
// AlgorithmMultiplicativeFrobenius.h:240-241.
template <typename T>
void gemm_v_ht(int m, int n, int k, const T* V, long ldV, const double* H, long ldH, double* P, long ldP) {
	const int RB = 256;
	const int nblocks = (m + RB - 1) / RB;
#pragma omp parallel
	{
		dvec acc((size_t)RB * k);
#pragma omp for schedule(dynamic, 1)
		for (int b = 0; b < nblocks; ++b) {
			const int i0 = b * RB, rows = std::min(RB, m - i0);
			std::fill(acc.begin(), acc.end(), 0.0);
			for (int j = 0; j < n; ++j) {
				const T* v = V + (size_t)j * ldV + i0;
				const double* h = H + (size_t)j * ldH;
				for (int i = 0; i < rows; ++i) {
					const double x = (double)v[i];
					double* a = acc.data() + (size_t)i * k;
					for (int c = 0; c < k; ++c) a[c] += x * h[c];
				}
			}
			for (int c = 0; c < k; ++c)
				for (int i = 0; i < rows; ++i) P[(size_t)c * ldP + i0 + i] = acc[(size_t)i * k + c];
		}
	}
}

// row-major packed copy of a column-major m x k matrix
void pack_rows(int m, int k, const double* W, long ldW, double* Wr) {
#pragma omp parallel for
	for (int i = 0; i < m; ++i)
		for (int c = 0; c < k; ++c) Wr[(size_t)i * k + c] = W[(size_t)c * ldW + i];
}

// G (k x k) = W^T W  from the row-major pack (cublasXsyrk/gemm, MU.h:168,176)
void gram_rows(int m, int k, const double* Wr, double* G) {
	std::fill(G, G + (size_t)k * k, 0.0);
#pragma omp parallel
	{
		dvec loc((size_t)k * k, 0.0);
#pragma omp for nowait
		for (int i = 0; i < m; ++i) {
			const double* w = Wr + (size_t)i * k;
			for (int a = 0; a < k; ++a) {
				const double wa = w[a];
				double* row = loc.data() + (size_t)a * k;
				for (int b = 0; b < k; ++b) row[b] += wa * w[b];
			}
		}
#pragma omp critical
		for (size_t t = 0; t < (size_t)k * k; ++t) G[t] += loc[t];
	}
}

// B (k x k) = H H^T with H k x n column-major (MU.h:208,231) -- each column is a packed "row" of H^T
void gram_cols(int k, int n, const double* H, long ldH, double* B) {
	std::fill(B, B + (size_t)k * k, 0.0);
#pragma omp parallel
	{
		dvec loc((size_t)k * k, 0.0);
#pragma omp for nowait
		for (int j = 0; j < n; ++j) {
			const double* h = H + (size_t)j * ldH;
			for (int a = 0; a < k; ++a) {
				const double ha = h[a];
				double* row = loc.data() + (size_t)a * k;
				for (int b = 0; b < k; ++b) row[b] += ha * h[b];
			}
		}
#pragma omp critical
		for (size_t t = 0; t < (size_t)k * k; ++t) B[t] += loc[t];
	}
}

// X (m x k) = W (m x k) * S (k x k), all column-major
void mul_right_small(int m, int k, const double* W, long ldW, const double* S, double* X, long ldX) {
#pragma omp parallel for
	for (int i = 0; i < m; ++i) {
		double row[1024];
		for (int c = 0; c < k; ++c) {
			double s = 0.0;
			for (int t = 0; t < k; ++t) s += W[(size_t)t * ldW + i] * S[(size_t)c * k + t];
			row[c] = s;
		}
		for (int c = 0; c < k; ++c) X[(size_t)c * ldX + i] = row[c];
	}
}

// Y (k x n) = S (k x k) * H (k x n)
void mul_left_small(int k, int n, const double* S, const double* H, long ldH, double* Y, long ldY) {
#pragma omp parallel for
	for (int j = 0; j < n; ++j) {
		double col[1024];
		for (int r = 0; r < k; ++r) {
			double s = 0.0;
			for (int t = 0; t < k; ++t) s += S[(size_t)t * k + r] * H[(size_t)j * ldH + t];
			col[r] = s;
		}
		for (int r = 0; r < k; ++r) Y[(size_t)j * ldY + r] = col[r];
	}
}

// source/nmf/KernelNormalizeColumns.cu:30-59 -- unit L2 columns, columns with zero norm untouched
void normalize_columns(int m, int k, double* W, long ldW) {
#pragma omp parallel for
	for (int c = 0; c < k; ++c) {
		double* w = W + (size_t)c * ldW;
		double s = 0.0;
		for (int i = 0; i < m; ++i) s += w[i] * w[i];
		if (s > 0.0) {
			s = std::sqrt(s);
			for (int i = 0; i < m; ++i) w[i] = w[i] / s;
		}
	}
}

// source/nmf/KernelTraceMultiplication.cu:30-81 with transposeA=false on k x k operands:
// partial[d] = sum_i A[d,i] * B[i,d]
void trace_partials_kk(int k, const double* A, const double* B, double* partial) {
	for (int d = 0; d < k; ++d) {
		double s = 0.0;
		for (int i = 0; i < k; ++i) s += A[(size_t)i * k + d] * B[(size_t)d * k + i];
		partial[d] = s;
	}
}

// same kernel with transposeA=true: partial[d] = sum_i A[i,d] * B[i,d]
void trace_partials_cols(int rows, int cols, const double* A, long ldA, const double* B, long ldB, double* partial) {
#pragma omp parallel for
	for (int d = 0; d < cols; ++d) {
		double s = 0.0;
		const double* a = A + (size_t)d * ldA;
		const double* b = B + (size_t)d * ldB;
		for (int i = 0; i < rows; ++i) s += a[i] * b[i];
		partial[d] = s;
	}
}

// source/nmf/FrobeniusResolver.cpp:30-51 -- ascending sort, interleaved accumulation, sqrt
double resolve_frobenius(const dvec& vtv_sorted, dvec second, dvec third) {
	std::sort(second.begin(), second.end());
	std::sort(third.begin(), third.end());
	double acc = 0.0;
	const size_t len = std::max(vtv_sorted.size(), std::max(second.size(), third.size()));
	for (size_t j = 0; j < len; ++j) {
		if (j < vtv_sorted.size()) acc += vtv_sorted[j];
		if (j < second.size()) acc -= 2.0 * second[j];
		if (j < third.size()) acc += third[j];
	}
	return std::sqrt(acc);
}

// Householder QR solve of the k x k system G X = R for nrhs right-hand sides, in place in R
// (ld k).  Follows the reference's solve route geqrf -> ormqr(Q^T) -> trsm(upper)
// (source/common/Matrix.h:565-618, GDCLS.h:196-206) rather than assuming G is SPD: the AHCLS
// matrix can be indefinite.
struct SmallQR {
	int k;
	dvec a;    // R in the upper triangle, Householder vectors below
	dvec tau;
	explicit SmallQR(int k_, const double* G) : k(k_), a(G, G + (size_t)k_ * k_), tau(k_, 0.0) {
		for (int j = 0; j < k; ++j) {
			double* col = a.data() + (size_t)j * k;
			double norm = 0.0;
			for (int i = j; i < k; ++i) norm += col[i] * col[i];
			norm = std::sqrt(norm);
			if (norm == 0.0) { tau[j] = 0.0; continue; }
			const double alpha = col[j];
			const double beta = alpha >= 0.0 ? -norm : norm;
			tau[j] = (beta - alpha) / beta;
			const double scale = 1.0 / (alpha - beta);
			for (int i = j + 1; i < k; ++i) col[i] *= scale;
			col[j] = beta;
			for (int c = j + 1; c < k; ++c) {  // apply H_j to the trailing columns
				double* cc = a.data() + (size_t)c * k;
				double dot = cc[j];
				for (int i = j + 1; i < k; ++i) dot += col[i] * cc[i];
				dot *= tau[j];
				cc[j] -= dot;
				for (int i = j + 1; i < k; ++i) cc[i] -= dot * col[i];
			}
		}
	}
	// x <- G^{-1} x  for one vector of length k
	void solve(double* x) const {
		for (int j = 0; j < k; ++j) {  // Q^T x
			if (tau[j] == 0.0) continue;
			const double* col = a.data() + (size_t)j * k;
			double dot = x[j];
			for (int i = j + 1; i < k; ++i) dot += col[i] * x[i];
			dot *= tau[j];
			x[j] -= dot;
			for (int i = j + 1; i < k; ++i) x[i] -= dot * col[i];
		}
		for (int j = k - 1; j >= 0; --j) {  // back substitution with R
			double s = x[j];
			for (int c = j + 1; c < k; ++c) s -= a[(size_t)c * k + j] * x[c];
			x[j] = s / a[(size_t)j * k + j];
		}
	}
};

// add (diag on the diagonal, offdiag elsewhere) to a k x k matrix
// (source/nmf/KernelFillMatrix.cu:29-45 with ReuseValue=true; argument order per Matrix.h:531-533)
void add_constraint(int k, double* G, double offdiag, double diag) {
	for (int c = 0; c < k; ++c)
		for (int r = 0; r < k; ++r) G[(size_t)c * k + r] += (r == c) ? diag : offdiag;
}

template <typename T>
double explicit_residual(int m, int n, int k, const T* V, long ldV, const double* W, long ldW, const double* H, long ldH) {
	dvec Wr((size_t)m * k);
	pack_rows(m, k, W, ldW, Wr.data());
	double total = 0.0;
#pragma omp parallel for reduction(+ : total) schedule(dynamic, 4)
	for (int j = 0; j < n; ++j) {
		const double* h = H + (size_t)j * ldH;
		const T* v = V + (size_t)j * ldV;
		double s = 0.0;
		for (int i = 0; i < m; ++i) {
			const double* w = Wr.data() + (size_t)i * k;
			double wh = 0.0;
			for (int c = 0; c < k; ++c) wh += w[c] * h[c];
			const double d = (double)v[i] - wh;
			s += d * d;
		}
		total += s;
	}
	return std::sqrt(total);
}

}  // namespace

extern "C" {

struct OracleConfig {
	int algorithm;          // NmfAlgorithm value
	int m, n, k;
	int num_iterations;
	int use_constant_w;     // NmfDescription::useConstantBasisVectors
	int threshold_type;     // 0 Frobenius, 1 RMSD
	double threshold_value;
	double eps;             // numeric_limits<T>::epsilon() of the API type (MU.h:191,244)
	double lambda;          // GDCLS
	double lambdaW, lambdaH, alphaW, alphaH;  // ACLS / AHCLS
	double theta;           // nsNMF
	int v_is_float;         // storage type of V
	int explicit_residual;  // also compute ||V - W_t H_t||_F at each check (costs one extra GEMM)
	int num_threads;        // 0 = OpenMP default
};

struct OracleTrace {
	int capacity;            // length of the three arrays
	int num_checks;          // filled
	int iterations_done;     // filled: value the dispatcher would store in ExecutionRecord::numIterations
	int* iteration;          // iteration number of each check
	double* frob_reported;   // the value the reference reports (trace identity, W_{t-1}/H_t mix per algorithm)
	double* frob_explicit;   // ||V - W_t H_t||_F (NaN unless explicit_residual)
};

int oracle_num_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

}  // extern "C"

// One run of one algorithm from the given W0/H0 (CopyExisting semantics), with the dispatcher's
// error cadence and stop rule (source/nmf/SingleGpuDispatcher.cpp:155-235).
template <typename T>
static int run_impl(const OracleConfig& cfg, const T* V, long ldV, double* W, long ldW, double* H, long ldH, OracleTrace* tr) {
	const int m = cfg.m, n = cfg.n, k = cfg.k;
	if (k > 1024) return -1;
#ifdef _OPENMP
	if (cfg.num_threads > 0) omp_set_num_threads(cfg.num_threads);
#endif
	dvec Wr((size_t)m * k), G((size_t)k * k), Gsaved((size_t)k * k), B((size_t)k * k);
	dvec N((size_t)k * n), P((size_t)m * k);
	dvec vtv(n), second, third(k);

	// setup: per-column squared norms of V, sorted ascending (MU.h:117-125)
	{
#pragma omp parallel for
		for (int j = 0; j < n; ++j) {
			const T* v = V + (size_t)j * ldV;
			double s = 0.0;
			for (int i = 0; i < m; ++i) s += (double)v[i] * (double)v[i];
			vtv[j] = s;
		}
		std::sort(vtv.begin(), vtv.end());
	}

	// nsNMF smoothing matrix (AlgorithmNonSmoothNMF.h:131-134)
	dvec S, Wt, Ht;
	if (cfg.algorithm == kNsNMF) {
		S.assign((size_t)k * k, cfg.theta / k);
		for (int c = 0; c < k; ++c) S[(size_t)c * k + c] = (1.0 - cfg.theta) + cfg.theta / k;
		Wt.resize((size_t)m * k);
		Ht.resize((size_t)k * n);
	}
	// Hoyer constants (AlgorithmAlternatingHoyerConstrainedLeastSquares.h:81-84)
	double betaW = 0.0, betaH = 0.0;
	if (cfg.algorithm == kAHCLS) {
		betaW = (1.0 - cfg.alphaW) * std::sqrt((double)k) + cfg.alphaW; betaW *= betaW;
		betaH = (1.0 - cfg.alphaH) * std::sqrt((double)k) + cfg.alphaH; betaH *= betaH;
	}

	tr->num_checks = 0;
	double last_error = 0.0, frob = 0.0;
	unsigned iteration = 1;
	const unsigned total = (unsigned)cfg.num_iterations;
	for (; iteration <= total; ++iteration) {
		const bool check = (iteration % 10 == 0) || iteration == total;  // SingleGpuDispatcher.cpp:173

		switch (cfg.algorithm) {
		case kMU: {
			// ---- H update (MU.h:164-198)
			pack_rows(m, k, W, ldW, Wr.data());
			gram_rows(m, k, Wr.data(), G.data());                       // A = W^T W
			gemm_wt_v(m, n, k, Wr.data(), V, ldV, N.data());            // N = W^T V
#pragma omp parallel for
			for (int j = 0; j < n; ++j) {
				double* h = H + (size_t)j * ldH;
				double d[1024];
				for (int r = 0; r < k; ++r) {                          // D = A H
					double s = 0.0;
					for (int t = 0; t < k; ++t) s += G[(size_t)t * k + r] * h[t];
					d[r] = s;
				}
				for (int r = 0; r < k; ++r)                            // KernelMultiplyDivide.cu:42
					h[r] = h[r] * N[(size_t)j * k + r] / (d[r] + cfg.eps);
			}
			if (check) {
				second.assign(n, 0.0);                                  // tr(H^T N) partials (MU.h:194-197)
				trace_partials_cols(k, n, H, ldH, N.data(), k, second.data());
				gram_cols(k, n, H, ldH, B.data());                      // B = H H^T (new H)
				trace_partials_kk(k, B.data(), G.data(), third.data()); // MU.h:203-216 (A from the old W)
			}
			if (!cfg.use_constant_w) {
				// ---- W update (MU.h:218-248)
				if (!check) gram_cols(k, n, H, ldH, B.data());
				gemm_v_ht(m, n, k, V, ldV, H, ldH, P.data(), m);        // N2 = V H^T
#pragma omp parallel for
				for (int i = 0; i < m; ++i) {
					double d[1024];
					for (int c = 0; c < k; ++c) {                      // D2 = W B
						double s = 0.0;
						for (int t = 0; t < k; ++t) s += W[(size_t)t * ldW + i] * B[(size_t)c * k + t];
						d[c] = s;
					}
					for (int c = 0; c < k; ++c) {
						double& w = W[(size_t)c * ldW + i];
						w = w * P[(size_t)c * m + i] / (d[c] + cfg.eps);
					}
				}
				normalize_columns(m, k, W, ldW);                        // MU.h:247
			}
			break;
		}
		case kNsNMF: {
			// AlgorithmNonSmoothNMF.h:173-218
			mul_right_small(m, k, W, ldW, S.data(), Wt.data(), m);      // W~ = W S
			pack_rows(m, k, Wt.data(), m, Wr.data());
			gram_rows(m, k, Wr.data(), G.data());                       // W~^T W~
			gemm_wt_v(m, n, k, Wr.data(), V, ldV, N.data());            // W~^T V
#pragma omp parallel for
			for (int j = 0; j < n; ++j) {
				double* h = H + (size_t)j * ldH;
				double d[1024];
				for (int r = 0; r < k; ++r) {
					double s = 0.0;
					for (int t = 0; t < k; ++t) s += G[(size_t)t * k + r] * h[t];
					d[r] = s;
				}
				for (int r = 0; r < k; ++r) h[r] = h[r] * N[(size_t)j * k + r] / (d[r] + cfg.eps);
			}
			if (check) {
				second.assign(n, 0.0);
				trace_partials_cols(k, n, H, ldH, N.data(), k, second.data());
			}
			if (!check && cfg.use_constant_w) break;
			mul_left_small(k, n, S.data(), H, ldH, Ht.data(), k);       // H~ = S H
			gram_cols(k, n, Ht.data(), k, B.data());                    // H~ H~^T
			if (check) {
				pack_rows(m, k, W, ldW, Wr.data());
				gram_rows(m, k, Wr.data(), Gsaved.data());              // W^T W (unsmoothed W)
				trace_partials_kk(k, B.data(), Gsaved.data(), third.data());
				if (cfg.use_constant_w) break;
			}
			gemm_v_ht(m, n, k, V, ldV, Ht.data(), k, P.data(), m);      // V H~^T
#pragma omp parallel for
			for (int i = 0; i < m; ++i) {
				double d[1024];
				for (int c = 0; c < k; ++c) {
					double s = 0.0;
					for (int t = 0; t < k; ++t) s += W[(size_t)t * ldW + i] * B[(size_t)c * k + t];
					d[c] = s;
				}
				for (int c = 0; c < k; ++c) {
					double& w = W[(size_t)c * ldW + i];
					w = w * P[(size_t)c * m + i] / (d[c] + cfg.eps);
				}
			}
			normalize_columns(m, k, W, ldW);
			break;
		}
		case kGDCLS:
		case kALS:
		case kACLS:
		case kAHCLS: {
			// ---- H by regularised least squares (GDCLS.h:174-209, AHCLS.h:170-214, ALS.h:145-172)
			pack_rows(m, k, W, ldW, Wr.data());
			gram_rows(m, k, Wr.data(), G.data());
			Gsaved = G;                                                 // W_old^T W_old for the trace
			if (cfg.algorithm == kGDCLS) add_constraint(k, G.data(), 0.0, cfg.lambda);
			else if (cfg.algorithm == kACLS) add_constraint(k, G.data(), 0.0, cfg.lambdaH);
			else if (cfg.algorithm == kAHCLS) add_constraint(k, G.data(), -cfg.lambdaH, cfg.lambdaH * betaH - cfg.lambdaH);
			gemm_wt_v(m, n, k, Wr.data(), V, ldV, N.data());
			{
				SmallQR qr(k, G.data());
#pragma omp parallel for
				for (int j = 0; j < n; ++j) {
					double x[1024];
					for (int r = 0; r < k; ++r) x[r] = N[(size_t)j * k + r];
					qr.solve(x);
					for (int r = 0; r < k; ++r) H[(size_t)j * ldH + r] = x[r] > 0.0 ? x[r] : 0.0;  // KernelMakeNonNegative.cu:30-46
				}
			}
			gram_cols(k, n, H, ldH, B.data());
			if (check) trace_partials_kk(k, B.data(), Gsaved.data(), third.data());  // GDCLS.h:216-227, AHCLS.h:217-224

			if (cfg.algorithm == kGDCLS) {
				// ---- W by the multiplicative rule (GDCLS.h:236-257); trace uses the NEW W (GDCLS.h:260-264)
				if (!cfg.use_constant_w) {
					gemm_v_ht(m, n, k, V, ldV, H, ldH, P.data(), m);
#pragma omp parallel for
					for (int i = 0; i < m; ++i) {
						double d[1024];
						for (int c = 0; c < k; ++c) {
							double s = 0.0;
							for (int t = 0; t < k; ++t) s += W[(size_t)t * ldW + i] * B[(size_t)c * k + t];
							d[c] = s;
						}
						for (int c = 0; c < k; ++c) {
							double& w = W[(size_t)c * ldW + i];
							w = w * P[(size_t)c * m + i] / (d[c] + cfg.eps);
						}
					}
					normalize_columns(m, k, W, ldW);
				} else if (check) {
					// The reference reads a stale V H^T here when W is constant (GDCLS.h:260-264 uses deviceMR, which only the
					// skipped W update writes): undefined.  The defined value of the same term is used instead.
					gemm_v_ht(m, n, k, V, ldV, H, ldH, P.data(), m);
				}
				if (check) {
					second.assign(k, 0.0);
					trace_partials_cols(m, k, P.data(), m, W, ldW, second.data());
				}
			} else {
				// ---- W by least squares (AHCLS.h:226-284, ALS.h:174-222); trace uses W saved BEFORE the update
				dvec Wold;
				if (check) {
					Wold.resize((size_t)m * k);
					for (int c = 0; c < k; ++c) std::memcpy(&Wold[(size_t)c * m], W + (size_t)c * ldW, sizeof(double) * m);
				}
				if (!cfg.use_constant_w) {
					if (cfg.algorithm == kACLS) add_constraint(k, B.data(), 0.0, cfg.lambdaW);
					else if (cfg.algorithm == kAHCLS) add_constraint(k, B.data(), -cfg.lambdaW, cfg.lambdaW * betaW - cfg.lambdaW);
					gemm_v_ht(m, n, k, V, ldV, H, ldH, P.data(), m);
				}
				if (check) {
					second.assign(k, 0.0);
					if (!cfg.use_constant_w) trace_partials_cols(m, k, Wold.data(), m, P.data(), m, second.data());
					else trace_partials_cols(m, k, Wold.data(), m, W, ldW, second.data());
				}
				if (!cfg.use_constant_w) {
					// W = P Q R^{-T} == P G^{-T}: solve G^T x = p_row.  G is symmetric, so G^T = G.
					SmallQR qr(k, B.data());
#pragma omp parallel for
					for (int i = 0; i < m; ++i) {
						double x[1024];
						for (int c = 0; c < k; ++c) x[c] = P[(size_t)c * m + i];
						qr.solve(x);
						for (int c = 0; c < k; ++c) W[(size_t)c * ldW + i] = x[c] > 0.0 ? x[c] : 0.0;
					}
					normalize_columns(m, k, W, ldW);
				}
			}
			break;
		}
		default:
			return -2;
		}

		if (check) {
			frob = resolve_frobenius(vtv, second, third);
			const double rmsd = frob / std::sqrt((double)m * (double)n);
			if (tr->num_checks < tr->capacity) {
				const int c = tr->num_checks;
				tr->iteration[c] = (int)iteration;
				tr->frob_reported[c] = frob;
				double fe = std::nan("");
				if (cfg.explicit_residual) {
					if (cfg.algorithm == kNsNMF) {
						mul_right_small(m, k, W, ldW, S.data(), Wt.data(), m);
						fe = explicit_residual(m, n, k, V, ldV, Wt.data(), m, H, ldH);
					} else {
						fe = explicit_residual(m, n, k, V, ldV, W, ldW, H, ldH);
					}
				}
				tr->frob_explicit[c] = fe;
			}
			tr->num_checks++;
			// stop rule (SingleGpuDispatcher.cpp:184-200): absolute delta, never on the first check
			const double cur = cfg.threshold_type == 0 ? frob : rmsd;
			const double delta = cur - last_error;
			if (last_error != 0.0 && std::fabs(delta) < cfg.threshold_value) break;
			last_error = cur;
		}
	}
	tr->iterations_done = (int)std::min(iteration, total);  // SingleGpuDispatcher.cpp:205

	if (cfg.algorithm == kNsNMF) {  // returned basis is W S (AlgorithmNonSmoothNMF.h:221-225)
		mul_right_small(m, k, W, ldW, S.data(), Wt.data(), m);
		for (int c = 0; c < k; ++c) std::memcpy(W + (size_t)c * ldW, &Wt[(size_t)c * m], sizeof(double) * m);
	}
	return 0;
}

extern "C" {

int oracle_nmf_run(const OracleConfig* cfg, const void* V, long ldV, double* W, long ldW, double* H, long ldH, OracleTrace* trace) {
	if (!cfg || !V || !W || !H || !trace) return -1;
	if (cfg->v_is_float) return run_impl<float>(*cfg, (const float*)V, ldV, W, ldW, H, ldH, trace);
	return run_impl<double>(*cfg, (const double*)V, ldV, W, ldW, H, ldH, trace);
}

// Timing helper for bench.py's cpu_baseline: `iters` full MU iterations (no error checks) on an
// m x n fp32 V; returns 0.  W,H are updated in place.
int oracle_mu_iterations(int m, int n, int k, const float* V, long ldV, double* W, long ldW, double* H, long ldH, int iters, double eps) {
	OracleConfig cfg;
	std::memset(&cfg, 0, sizeof(cfg));
	cfg.algorithm = kMU; cfg.m = m; cfg.n = n; cfg.k = k;
	cfg.num_iterations = iters; cfg.eps = eps; cfg.v_is_float = 1;
	int it[8]; double a[8], b[8];
	OracleTrace tr; tr.capacity = 8; tr.iteration = it; tr.frob_reported = a; tr.frob_explicit = b;
	return run_impl<float>(cfg, V, ldV, W, ldW, H, ldH, &tr);
}

}  // extern "C"
