// Internal C++ entry points of the FP32 GEMM kernels (3xTF32 on tcgen05/TMEM). The public door is include/cosma_b200.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cosma_b200.h"

namespace cosma_b200 {

// path_used (optional): 0 = no GEMM kernel launched (degenerate), 1 = TMA + tcgen05 kernel, 2 = generic kernel
// (unaligned operands; CGEMM with a transposed operand).
int sgemm_sm100(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int* path_used);
// complex: interleaved (re, im) floats; alpha, beta point at 2 floats; leading dimensions in complex elements
int cgemm_sm100(cudaStream_t stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha, const float* A,
                int64_t lda, const float* B, int64_t ldb, const float* beta, float* C, int64_t ldc, int* path_used);

}  // namespace cosma_b200
