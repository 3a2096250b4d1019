// One GEMM whose operands may still be in host memory: PCIe transfers pipelined under the kernel (host_gemm.cu).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>
#include <cuda_runtime.h>

namespace cosma_b200 {

struct StreamGemmArgs {
    char dtype = 'd';  // 's' | 'd' | 'c' | 'z'
    int64_t m = 0, n = 0, k = 0;
    const double* alpha = nullptr;  // 1 double, or (re, im) for complex types (converted to float for 's' / 'c')
    const double* beta = nullptr;
    // device operands (always valid) and their leading dimensions in elements
    void* dA = nullptr; int64_t dlda = 0;
    void* dB = nullptr; int64_t dldb = 0;
    void* dC = nullptr; int64_t dldc = 0;
    // host sources / sink, column-major; nullptr = the device buffer already holds the data / the result stays on the device
    const void* hA = nullptr; int64_t lda = 0;
    const void* hB = nullptr; int64_t ldb = 0;
    const void* hC_in = nullptr;  // read only when beta != 0
    void* hC_out = nullptr;
    int64_t ldc = 0;
};

// C = alpha*A*B + beta*C ('N','N'); asynchronous; `stream` completes when the result is where it was asked to go.
int stream_gemm(cudaStream_t stream, const StreamGemmArgs& args, int* launches);

int launch_gemm_nn(char dtype, cudaStream_t stream, int64_t m, int64_t n, int64_t k, const double* alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, const double* beta, void* C, int64_t ldc, int* path);

std::vector<std::pair<int64_t, int64_t>> stream_panels(char dtype, int64_t n, bool a_streamed, bool c_to_host);

}  // namespace cosma_b200
