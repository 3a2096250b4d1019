/* cosma_b200 -- C ABI of the B200-native COSMA hot path.
 *
 * This is the only door between host code (C++, Fortran, Python/ctypes, Julia ccall ...) and the
 * sm_100a kernels. Plain pointers and sizes; every entry point returns an int status and never
 * throws. All device work is asynchronous on the caller's stream (a cudaStream_t passed as void*).
 *
 * Each entry point names the reference interface it replaces (paths relative to eth-cscs/COSMA
 * v2.8.4).  INTEGRATION.md shows the binding a reference maintainer would add.
 */
#ifndef COSMA_B200_H
#define COSMA_B200_H

#include <stdint.h>

#if defined(COSMA_B200_BUILD)
#define COSMA_B200_API __attribute__((visibility("default")))
#else
#define COSMA_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum cosma_b200_status {
    COSMA_B200_OK = 0,
    COSMA_B200_INVALID_ARG = 1,
    COSMA_B200_CUDA_ERROR = 2,
    COSMA_B200_NCCL_ERROR = 3,
    COSMA_B200_OUT_OF_MEMORY = 4,
    COSMA_B200_NOT_SUPPORTED = 5,
    COSMA_B200_INTERNAL_ERROR = 6
};

/* Library / build identification: "cosma_b200 <version> sm_100a". */
COSMA_B200_API const char* cosma_b200_version(void);
/* Human-readable text for the last error on the calling thread ("" if none). */
COSMA_B200_API const char* cosma_b200_last_error(void);

/* ---- local GEMM (K1/K2) -------------------------------------------------------------------
 * C = alpha*op(A)*op(B) + beta*C, column-major, DEVICE pointers, alpha/beta HOST pointers.
 * Replaces gpu::gemm(mm_handle&, transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, ...)
 *   (libs/Tiled-MM/src/Tiled-MM/tiled_mm.hpp:69-79; called from src/cosma/local_multiply.cpp:219-269)
 * and cblas_?gemm (src/cosma/blas.cpp:24-130). trans = 'N' | 'T' | 'C'. beta == 0 never reads C.
 * Complex scalars/matrices are interleaved (re, im) doubles, leading dimensions in complex elements.
 */
COSMA_B200_API int cosma_b200_dgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C,
                     int64_t ldc);
COSMA_B200_API int cosma_b200_zgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, const double* beta, double* C,
                     int64_t ldc);
/* Which kernel the last ?gemm call on this thread used: 0 none, 1 TMA+DMMA persistent, 2 generic. */
COSMA_B200_API int cosma_b200_last_gemm_path(void);

/* ---- planning layer (host only, no CUDA) --------------------------------------------------------
 * cosma::Strategy (reference src/cosma/strategy.hpp:16-183) and cosma::Mapper (src/cosma/mapper.hpp:21-126) in
 * C form. Steps are written as the reference miniapp's -s string: "pm2,sn4,pk2" (p|s)(m|n|k)(divisor).
 */
/* Strategy(m,n,k,P[,prefix][,mem_limit in elements, <=0: unlimited]) -> step string in out; *P_out = ranks used
 * (ranks >= *P_out idle, reference strategy.cpp:139-169); *mem_used = elements per rank. */
COSMA_B200_API int cosma_b200_strategy(int m, int n, int k, int P, long long mem_limit, const char* prefix, char* out,
                                       int out_len, int* P_out, long long* mem_used);
/* Mapper(label in 'A'|'B'|'C').complete_layout(): counts[r] = blocks of rank r; out = [r][block]{row_first,row_last,
 * col_first,col_last} (inclusive), flattened in local-buffer order. */
COSMA_B200_API int cosma_b200_mapper_layout(char label, int m, int n, int k, int P, const char* steps, int* counts, int* out,
                                            int out_cap, int* total_blocks);
COSMA_B200_API int cosma_b200_mapper_local_coordinates(char label, int m, int n, int k, int P, const char* steps, int gi,
                                                       int gj, int64_t* local_idx, int* rank);
COSMA_B200_API int cosma_b200_mapper_global_coordinates(char label, int m, int n, int k, int P, const char* steps,
                                                        int64_t local_idx, int rank, int* gi, int* gj);

/* ---- local GEMM with HOST operands ('N','N') -------------------------------------------------
 * Same contract as the reference's GPU base case, which receives host pointers and streams tiles
 * through the device (gpu::gemm, libs/Tiled-MM/src/Tiled-MM/tiled_mm.cpp:492-624; copy_c_back = true).
 * A is uploaded once; column panels of B and C are pipelined (H2D / DMMA kernel / D2H overlap on three
 * streams). Pinned host memory is required for the copies to be asynchronous. Work is ordered after
 * everything already queued on `stream`, and `stream` completes when C is back in host memory.
 */
COSMA_B200_API int cosma_b200_dgemm_host(void* stream, int64_t m, int64_t n, int64_t k, const double* alpha, const double* A,
                                         int64_t lda, const double* B, int64_t ldb, const double* beta, double* C,
                                         int64_t ldc);
COSMA_B200_API int cosma_b200_zgemm_host(void* stream, int64_t m, int64_t n, int64_t k, const double* alpha, const double* A,
                                         int64_t lda, const double* B, int64_t ldb, const double* beta, double* C,
                                         int64_t ldc);
/* Number of GEMM kernels launched by the last *_host call on this thread. */
COSMA_B200_API int cosma_b200_last_launch_count(void);
/* Frees the device staging workspace cached by the *_host entry points (thread-local). */
COSMA_B200_API void cosma_b200_release_workspace(void);

#ifdef __cplusplus
}
#endif
#endif /* COSMA_B200_H */
