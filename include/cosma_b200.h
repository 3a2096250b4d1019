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

/* ---- distributed multiply ------------------------------------------------------------------------
 * cosma::multiply(A, B, C, strategy, comm, alpha, beta) (reference src/cosma/multiply.hpp:47-54, multiply.cpp:222-314)
 * split into its plan-time and run-time halves. One process per GPU; the caller selects the device first.
 *
 * Communicator: the reference turns an MPI_Comm into NCCL communicators by broadcasting an ncclUniqueId over MPI
 * (src/cosma/gpu/nccl_utils.cpp:21-42). Here the host does that broadcast with whatever it has (MPI_Bcast,
 * torch.distributed, a file) and hands the 128 bytes in.
 */
COSMA_B200_API int cosma_b200_nccl_unique_id(uint8_t* out128);                 /* call on one rank, broadcast */
COSMA_B200_API int cosma_b200_comm_create(int rank, int nranks, const uint8_t* id128, void** comm_out);
COSMA_B200_API int cosma_b200_comm_destroy(void* comm);
/* Plan = Strategy + Mapper x3 + compiled schedule + one ring communicator per parallel step (created collectively:
 * every rank of `comm` must call this with the same m, n, k, steps). comm == NULL builds the plan only (any rank of
 * any nranks; no execution) -- used by tests and tools. steps: "" = automatic (Strategy(m,n,k,P)), else e.g.
 * "pm2,pn2,pk2". dtype: 'd' (double) | 'z' (complex double). */
COSMA_B200_API int cosma_b200_plan_create(void* comm, int rank, int nranks, int m, int n, int k, const char* steps, char dtype,
                                          void** plan_out);
COSMA_B200_API int cosma_b200_plan_destroy(void* plan);
/* Arena sizes in elements. matrix: 0 = A, 1 = B, 2 = C. The first initial_elements of an arena are the rank's local
 * matrix in the reference's layout (CosmaMatrix::matrix_pointer(), matrix_size(); src/cosma/matrix.hpp:26-213); the
 * rest is communication workspace. */
COSMA_B200_API int64_t cosma_b200_plan_arena_elements(void* plan, int matrix);
COSMA_B200_API int64_t cosma_b200_plan_initial_elements(void* plan, int matrix);
COSMA_B200_API int cosma_b200_plan_strategy(void* plan, char* out, int out_len, int* P_used);
COSMA_B200_API double cosma_b200_plan_gemm_flops(void* plan);                  /* real flops of this rank's local GEMMs */
COSMA_B200_API int cosma_b200_plan_export(void* plan, int64_t* buf, int64_t cap, int64_t* len); /* see schedule.cpp */
COSMA_B200_API int cosma_b200_plan_local_blocks(void* plan, int matrix, int rank, int* out, int cap, int* n_blocks);
/* C = alpha*A*B + beta*C on the plan's layout. A, B, C: DEVICE arenas of at least plan_arena_elements each; alpha,
 * beta: host pointers (1 double, or 2 for 'z'). Asynchronous on `stream`. Idle ranks return immediately. */
COSMA_B200_API int cosma_b200_multiply(void* plan, const double* alpha, const double* beta, void* A, void* B, void* C,
                                       void* stream);
/* Same with HOST local matrices (pinned for asynchrony): H2D of local A, B (C too when beta != 0) into arenas owned by
 * the plan, run, D2H of local C. The reference's calling convention (host-resident CosmaMatrix buffers). */
COSMA_B200_API int cosma_b200_multiply_host(void* plan, const double* alpha, const double* beta, const void* A, const void* B,
                                            void* C, void* stream);
COSMA_B200_API int cosma_b200_plan_last_launches(void* plan);                  /* GEMM kernels launched by the last run */
COSMA_B200_API int cosma_b200_plan_time_gemms(void* plan, int enable);         /* record CUDA events around each GEMM */
COSMA_B200_API int cosma_b200_plan_gemm_times(void* plan, float* out_ms, int cap, int* n);

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
