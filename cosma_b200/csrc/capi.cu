// extern "C" surface of libcosma_b200.so (declared in include/cosma_b200.h).
#include "../../include/cosma_b200.h"
#include "gemm_f64_sm100.h"
#include "gemm_tf32x3_sm100.h"
#include "exec_internal.h"
#include "cta_budget.h"

#include <string>

namespace {
thread_local std::string g_last_error;
thread_local int g_last_gemm_path = 0;
thread_local int g_last_launches = 0;
}  // namespace

namespace cosma_b200 {
void set_last_error(const std::string& msg) { g_last_error = msg; }
int& reserved_sms() {
    thread_local int r = 0;
    return r;
}
}  // namespace cosma_b200

extern "C" {

const char* cosma_b200_version(void) { return "cosma_b200 0.1.0 sm_100a"; }
const char* cosma_b200_last_error(void) { return g_last_error.c_str(); }
int cosma_b200_last_gemm_path(void) { return g_last_gemm_path; }

int cosma_b200_dgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C,
                     int64_t ldc) {
    if (!alpha || !beta) return COSMA_B200_INVALID_ARG;
    return cosma_b200::dgemm_sm100(static_cast<cudaStream_t>(stream), transa, transb, m, n, k, *alpha, A, lda, B, ldb,
                                   *beta, C, ldc, &g_last_gemm_path);
}

int cosma_b200_zgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C,
                     int64_t ldc) {
    if (!alpha || !beta) return COSMA_B200_INVALID_ARG;
    return cosma_b200::zgemm_sm100(static_cast<cudaStream_t>(stream), transa, transb, m, n, k, alpha, A, lda, B, ldb, beta,
                                   C, ldc, &g_last_gemm_path);
}

int cosma_b200_sgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha, const float* A,
                     int64_t lda, const float* B, int64_t ldb, const float* beta, float* C, int64_t ldc) {
    if (!alpha || !beta) return COSMA_B200_INVALID_ARG;
    return cosma_b200::sgemm_sm100(static_cast<cudaStream_t>(stream), transa, transb, m, n, k, *alpha, A, lda, B, ldb, *beta, C, ldc,
                                   &g_last_gemm_path);
}

int cosma_b200_cgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha, const float* A,
                     int64_t lda, const float* B, int64_t ldb, const float* beta, float* C, int64_t ldc) {
    if (!alpha || !beta) return COSMA_B200_INVALID_ARG;
    return cosma_b200::cgemm_sm100(static_cast<cudaStream_t>(stream), transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc,
                                   &g_last_gemm_path);
}

int cosma_b200_dgemm_host(void* stream, int64_t m, int64_t n, int64_t k, const double* alpha, const double* A, int64_t lda,
                          const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc) {
    return cosma_b200::guarded("cosma_b200_dgemm_host", [&]() -> int {
        if (!alpha || !beta) return COSMA_B200_INVALID_ARG;
        return cosma_b200::gemm_f64_host(static_cast<cudaStream_t>(stream), 1, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc,
                                         &g_last_launches);
    });
}
int cosma_b200_zgemm_host(void* stream, int64_t m, int64_t n, int64_t k, const double* alpha, const double* A, int64_t lda,
                          const double* B, int64_t ldb, const double* beta, double* C, int64_t ldc) {
    return cosma_b200::guarded("cosma_b200_zgemm_host", [&]() -> int {
        if (!alpha || !beta) return COSMA_B200_INVALID_ARG;
        return cosma_b200::gemm_f64_host(static_cast<cudaStream_t>(stream), 2, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc,
                                         &g_last_launches);
    });
}
int cosma_b200_last_launch_count(void) { return g_last_launches; }

static int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return COSMA_B200_OK;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? COSMA_B200_OUT_OF_MEMORY : COSMA_B200_CUDA_ERROR;
}
int cosma_b200_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return COSMA_B200_INVALID_ARG;
    *ptr = nullptr;
    if (bytes == 0) return COSMA_B200_OK;
    return cuda_status(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable), "cudaHostAlloc");
}
int cosma_b200_host_free(void* ptr) { return ptr ? cuda_status(cudaFreeHost(ptr), "cudaFreeHost") : COSMA_B200_OK; }
int cosma_b200_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return COSMA_B200_OK;
    return cuda_status(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable), "cudaHostRegister");
}
int cosma_b200_host_unregister(void* ptr) { return ptr ? cuda_status(cudaHostUnregister(ptr), "cudaHostUnregister") : COSMA_B200_OK; }
int cosma_b200_device_count(int* count) {
    if (!count) return COSMA_B200_INVALID_ARG;
    *count = 0;
    return cuda_status(cudaGetDeviceCount(count), "cudaGetDeviceCount");
}
int cosma_b200_set_device(int device) { return cuda_status(cudaSetDevice(device), "cudaSetDevice"); }
int cosma_b200_device_pci_bus_id(int device, char* out, int out_len) {
    if (!out || out_len < 13) return COSMA_B200_INVALID_ARG;
    const int st = cuda_status(cudaDeviceGetPCIBusId(out, out_len, device), "cudaDeviceGetPCIBusId");
    if (st != COSMA_B200_OK) return st;
    for (char* c = out; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = static_cast<char>(*c - 'A' + 'a');
    return COSMA_B200_OK;
}
int cosma_b200_stream_synchronize(void* stream) {
    return cuda_status(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "cudaStreamSynchronize");
}
void cosma_b200_release_workspace(void) { cosma_b200::release_host_gemm_workspace(); }

}  // extern "C"
