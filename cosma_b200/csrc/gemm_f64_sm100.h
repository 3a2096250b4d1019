// Internal C++ entry points of the FP64 GEMM kernels (the public door is include/cosma_b200.h).
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/cosma_b200.h"

namespace cosma_b200 {

// path_used (optional): 0 = no GEMM kernel launched (degenerate), 1 = TMA/DMMA persistent kernel,
// 2 = generic (unaligned) kernel.
int dgemm_sm100(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                int* path_used);

int zgemm_sm100(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc,
                int* path_used);

// Host-pointer variant (NN): operands in (ideally pinned) host memory, staged through HBM with panel pipelining.
int gemm_f64_host(cudaStream_t stream, int elem_doubles, int64_t m, int64_t n, int64_t k, const double* alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc,
                  int* launches);
void release_host_gemm_workspace();

}  // namespace cosma_b200
