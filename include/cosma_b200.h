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
/* K3/K4: single precision on the tcgen05 tensor cores with the 3xTF32 split (FP32 accumulation in TMEM); same contract
 * as above with float data (cblas_sgemm / cblas_cgemm, cublasSgemm / cublasCgemm in the reference: src/cosma/blas.cpp:
 * 24-130, libs/Tiled-MM/src/Tiled-MM/tiled_mm.cpp:181-268). Normwise error ~1e-6 or better against an FP32 GEMM.
 * CGEMM runs on the tensor cores for transa = transb = 'N' (the only form cosma::multiply issues); other complex
 * transposes use the generic kernel. */
COSMA_B200_API int cosma_b200_sgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha,
                     const float* A, int64_t lda, const float* B, int64_t ldb, const float* beta, float* C, int64_t ldc);
COSMA_B200_API int cosma_b200_cgemm(void* stream, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha,
                     const float* A, int64_t lda, const float* B, int64_t ldb, const float* beta, float* C, int64_t ldc);
/* Which kernel the last ?gemm call on this thread used: 0 none, 1 TMA + tensor-pipe persistent kernel (DMMA for d/z,
 * tcgen05 for s/c), 2 generic. */
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
 * "pm2,pn2,pk2". dtype: 's' | 'd' | 'c' | 'z' (float, double, complex float, complex double). */
COSMA_B200_API int cosma_b200_plan_create(void* comm, int rank, int nranks, int m, int n, int k, const char* steps, char dtype,
                                          void** plan_out);
/* Same for an explicit cosma::Strategy (reference multiply(A, B, C, strategy, comm, alpha, beta), multiply.hpp:47-54): the
 * strategy's own rank count P (<= ranks of comm; ranks >= P idle, multiply.cpp:258-260) and its step list, which may be
 * empty when P == 1. `steps` is never completed or replaced by the automatic strategy here. */
COSMA_B200_API int cosma_b200_plan_create_for_strategy(void* comm, int rank, int nranks, int m, int n, int k, int P, const char* steps,
                                                       char dtype, void** plan_out);
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
 * beta: host pointers to doubles for every dtype (1 value, or 2 = (re, im) for 'c'/'z'; converted to float for 's'/'c').
 * Asynchronous on `stream`. Idle ranks return immediately. */
COSMA_B200_API int cosma_b200_multiply(void* plan, const double* alpha, const double* beta, void* A, void* B, void* C,
                                       void* stream);
/* Planning only (host-resident operands at N > 1): column panel j of c of this rank's local matrices. The column ranges
 * of all ranks' B and C blocks are refined into elementary ranges; panel j of all ranks -- the j-th c-th of every elementary range,
 * taken from local B where it lies inside the rank's B columns and from local C where inside its C columns -- is the native layout of
 * the problem (m, n / c, k) under the same strategy, so the host-pointer entry point can run c sub-problems with A uploaded and
 * gathered once while panel j + 1 travels up and panel j - 1 down (COSMA_B200_HOST_PANELS=c, opt-in). b_pieces / c_pieces: (src_off,
 * len, dst_off) element triples (local buffer -> panel buffer); *eligible = 0 when the layout cannot be cut this way. */
COSMA_B200_API int cosma_b200_plan_host_panel(void* plan, int c, int j, int64_t* b_pieces, int b_cap, int* n_b, int64_t* c_pieces, int c_cap,
                                              int* n_c, int* eligible);
/* Same with HOST local matrices (pinned for asynchrony): H2D of local A, B (C too when beta != 0) into arenas owned by
 * the plan, run, D2H of local C. The reference's calling convention (host-resident CosmaMatrix buffers). */
COSMA_B200_API int cosma_b200_multiply_host(void* plan, const double* alpha, const double* beta, const void* A, const void* B,
                                            void* C, void* stream);
COSMA_B200_API int cosma_b200_plan_last_launches(void* plan);                  /* GEMM kernels launched by the last run */
COSMA_B200_API int cosma_b200_plan_time_gemms(void* plan, int enable);         /* record CUDA events around each GEMM */
COSMA_B200_API int cosma_b200_plan_gemm_times(void* plan, float* out_ms, int cap, int* n);
/* With timing enabled, per op of the compiled schedule: kind (0 GEMM, 1 allgather, 2 reduce-scatter), device ms, and for the
 * collectives the bytes this rank exchanges, (d-1)/d of the gathered / reduced buffer (ring size d). */
COSMA_B200_API int cosma_b200_plan_op_times(void* plan, int* kinds, float* ms, int64_t* wire_bytes, int cap, int* n);
/* The plan's communication / computation overlap (reference one_sided_communicator.cpp:417-1016, gate strategy.cpp:851-901; here a
 * plan-time lowering, include/cosma/overlap.hpp): *enabled = 0 when the ops run serially (why = a one-line reason; controlled by
 * COSMA_OVERLAP_COMM_AND_COMP = ON | OFF | FORCE). buf receives the micro-op program, est_ms[3] the planner's estimates
 * {serial, overlapped, communication}. */
COSMA_B200_API int cosma_b200_plan_overlap_export(void* plan, int64_t* buf, int64_t cap, int64_t* len, int* enabled, char* why, int why_len,
                                                  double* est_ms);
/* Binds the DEVICE arenas the plan will run on -- collective over the plan's communicator, idle ranks included. An overlapped plan then
 * moves its ring-of-two transfers with copy engines straight into the ring mates' arenas (CUDA IPC mappings, ordered by stream memory
 * operations on epoch flags: no SM is spent on communication) and re-plans its GEMM panels for the whole device. *active = 1 when that
 * transport is in place (every rank alike); 0: plan not overlapped, COSMA_B200_PEER_COPY=OFF (the overlapped transfers stay NCCL kernels
 * beside narrow GEMMs), or a rank could not map its mate's memory (the plan then runs its serial schedule over NCCL, as it does on
 * arenas that were never bound). cosma_b200_multiply must afterwards be called with exactly these arenas. The library binds
 * the arenas it owns itself (cosma_b200_multiply_host, ?multiply_using_layout, p?gemm). */
COSMA_B200_API int cosma_b200_plan_bind_arenas(void* plan, void* A, void* B, void* C, int* active);

/* ---- COSTA relayout (R3/R4 + exchange) ------------------------------------------------------------------
 * Layout description in the shape of the reference's C interface (src/cosma/cinterface.hpp:16-41: struct block,
 * struct layout; libs/COSTA/src/costa/layout.hpp:14-48: block_t, custom_layout): a grid of blocks given by split
 * points, the owner rank of every block (row-major: owners[i*colblocks + j]) and the blocks of the calling rank.
 * `data` points at element (0,0) of the block in DEVICE memory; `ld` is its leading dimension in elements. */
typedef struct cosma_b200_block {
    void* data;
    int ld;
    int row;  /* block row index in the grid */
    int col;  /* block column index in the grid */
} cosma_b200_block;
typedef struct cosma_b200_layout {
    int rowblocks;
    int colblocks;
    const int* rowsplit;  /* rowblocks + 1 entries: block i covers rows [rowsplit[i], rowsplit[i+1]) */
    const int* colsplit;
    const int* owners;
    int nlocalblocks;
    cosma_b200_block* localblocks;
} cosma_b200_layout;

/* One batched launch of COSTA's kernel: dest = beta*dest + alpha*op(src) for every piece.
 * Replaces costa::memory::copy_and_transform (libs/COSTA/src/costa/grid2grid/memory_utils.hpp:287-346), same
 * argument meaning: n_rows x n_cols is the SOURCE block, orderings 'C' (column-major) | 'R' (row-major), ld = 0 means
 * tight. alpha/beta: (re, im) pairs (im ignored for real dtypes). dtype: 's','d','c','z'. DEVICE pointers. With
 * alpha = 1, beta = 0 the result is a bit-exact move; beta == 0 never reads dest. */
typedef struct cosma_b200_piece {
    const void* src;
    void* dst;
    int64_t src_ld, dst_ld;
    int n_rows, n_cols;
    char src_ordering, dst_ordering;
    char transpose, conjugate;  /* 0 | 1 */
    double alpha[2], beta[2];
} cosma_b200_piece;
COSMA_B200_API int cosma_b200_relayout_batch(void* stream, char dtype, int n, const cosma_b200_piece* pieces);

/* costa::transform(from[], to[], trans[], alpha[], beta[], comm) (libs/COSTA/src/costa/grid2grid/transform.cpp:231-282)
 * and costa::transformer<T>::schedule/transform (transformer.hpp:8-63), split into plan and run:
 *   to[i] = beta[i]*to[i] + alpha[i]*op_i(from[i]),  op_i = trans[i] in 'N' | 'T' | 'C', for all i in one exchange.
 * ordering_from/ordering_to: one char per layout ('C' | 'R'), NULL = all 'C'. alpha/beta: 2 doubles per transform.
 * comm == NULL plans for (rank, nranks) without the ability to exchange (single rank, or tests that only export the
 * plan). Collective over comm when executed: pack kernel -> one NCCL group of send/recv -> unpack kernel. */
COSMA_B200_API int cosma_b200_transform_plan_create(void* comm, int rank, int nranks, char dtype, int n,
                                                    const cosma_b200_layout* from, const cosma_b200_layout* to,
                                                    const char* ordering_from, const char* ordering_to, const char* trans,
                                                    const double* alpha, const double* beta, void** plan_out);
COSMA_B200_API int cosma_b200_transform_run(void* plan, void* stream);
COSMA_B200_API int cosma_b200_transform_plan_destroy(void* plan);
/* Flat int64 dump of the plan for tests/tools (format: transform_exec.cu). Two-call pattern: buf == NULL -> *len. */
COSMA_B200_API int cosma_b200_transform_plan_export(void* plan, int64_t* buf, int64_t cap, int64_t* len);
/* Elements this rank moves per run: [0] staying on the rank, [1] sent to peers; kernels launched by the last run. */
COSMA_B200_API int cosma_b200_transform_plan_stats(void* plan, int64_t* local_elements, int64_t* remote_elements, int* launches);

/* costa::get_scalapack_layout (libs/COSTA/src/costa/grid2grid/scalapack_layout.cpp:178-285): grid, owners and the
 * local blocks (as element offsets into the rank's local array) of sub(A) = A(ia:ia+sub_m-1, ja:ja+sub_n-1) in a
 * block-cyclic distribution. Two-call pattern: pass NULL arrays to obtain *rowblocks, *colblocks, *nlocal. */
COSMA_B200_API int cosma_b200_scalapack_layout(int lld, int mat_rows, int mat_cols, int ia, int ja, int sub_m, int sub_n, int mb,
                                               int nb, int nprow, int npcol, char grid_order, int rsrc, int csrc,
                                               char data_ordering, int rank, int* rowblocks, int* colblocks, int* rowsplit,
                                               int* colsplit, int* owners, int* nlocal, int* local_row, int* local_col,
                                               int64_t* local_offset);
/* costa::communication_volume (libs/COSTA/src/costa/grid2grid/transform.cpp:9-44): elements exchanged between every pair of
 * ranks when a matrix moves from grid a (transposed first when trans != 'N') to grid b; volume is n_ranks x n_ranks,
 * entry [u * n_ranks + v], u <= v, counts both directions, [u * n_ranks + u] is what stays on rank u.
 * costa::optimal_reordering (ranks_reordering.cpp:4-61): the greedy matching on that graph; permutation[r] = new label. */
COSMA_B200_API int cosma_b200_comm_volume(int rowblocks_a, int colblocks_a, const int* rowsplit_a, const int* colsplit_a, const int* owners_a,
                                          int rowblocks_b, int colblocks_b, const int* rowsplit_b, const int* colsplit_b, const int* owners_b,
                                          char trans, int n_ranks, long long* volume);
COSMA_B200_API int cosma_b200_optimal_reordering(int n_ranks, const long long* volume, int* permutation, int* reordered);
/* cosma::adapt_strategy_to_block_cyclic_grid (src/cosma/cosma_pxgemm.cpp:517-650): the strategy prefix that reproduces the
 * block-cyclic grid of the largest operand of a p?gemm call ("" when the reference's conditions do not hold). */
COSMA_B200_API int cosma_b200_adapt_strategy(int m, int n, int k, int P, const int* desca, int ia, int ja, const int* descb, int ib, int jb,
                                             const int* descc, int ic, int jc, char transa, char transb, int nprow, int npcol, char order,
                                             char* out, int out_len);
/* The automatic strategy tightened with sequential steps until the device arenas of its COMPILED schedule (local matrices +
 * communication workspace) fit budget_bytes per rank. The same is applied inside cosma_b200_plan_create (steps = ""),
 * ?multiply_using_layout and p?gemm when COSMA_B200_DEVICE_MEMORY_MB is set; COSMA_CPU_MAX_MEMORY [MB] keeps its reference meaning
 * (limit of the Strategy's own memory model). */
COSMA_B200_API int cosma_b200_fit_strategy(int m, int n, int k, int P, const char* prefix, int elem_bytes, long long budget_bytes, char* out,
                                           int out_len, int* P_out, long long* footprint_bytes);
/* ScaLAPACK NUMROC. */
COSMA_B200_API int cosma_b200_numroc(int n, int nb, int iproc, int isrcproc, int nprocs);

/* ---- multiply on caller-defined layouts ------------------------------------------------------------------
 * {d,z}multiply_using_layout(MPI_Comm, transa, transb, alpha, layout A, layout B, beta, layout C)
 * (reference src/cosma/cinterface.hpp:42-76, cinterface.cpp:54-150 -> multiply_using_layout, multiply.cpp:78-213):
 * C = alpha*op(A)*op(B) + beta*C with every matrix in its own block layout (column-major blocks, DEVICE pointers).
 * op(A) and op(B) are moved into COSMA's native layout by the relayout kernels, multiplied with the automatic
 * Strategy(m,n,k,P), and the result is moved into C's layout with (alpha, beta) applied on the way. Collective over
 * `comm` (a cosma_b200 communicator handle in place of MPI_Comm). alpha/beta: 1 double ('d') or 2 ('z'). Strategy,
 * ring communicators, arenas and transform plans are cached in the communicator across calls. */
COSMA_B200_API int cosma_b200_dmultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                                     const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                                     const cosma_b200_layout* C, void* stream);
COSMA_B200_API int cosma_b200_zmultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                                     const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                                     const cosma_b200_layout* C, void* stream);
/* single precision: float / complex-float blocks; alpha and beta are still passed as doubles (converted to float) */
COSMA_B200_API int cosma_b200_smultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                                     const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                                     const cosma_b200_layout* C, void* stream);
COSMA_B200_API int cosma_b200_cmultiply_using_layout(void* comm, const char* transa, const char* transb, const double* alpha,
                                                     const cosma_b200_layout* A, const cosma_b200_layout* B, const double* beta,
                                                     const cosma_b200_layout* C, void* stream);

/* After a synchronised ?multiply_using_layout / p?gemm on `comm`: device milliseconds of its three phases (relayout
 * of A and B in, multiply, relayout of C out), elements moved by this rank's relayouts (in: staying on the rank, sent to
 * peers; out: same), the strategy used ("pm2,pn2,pk2") and the number of kernels launched. */
COSMA_B200_API int cosma_b200_last_layout_multiply_stats(void* comm, float* ms3, int64_t* elements4, char* strategy, int strategy_len,
                                                         int* launches);

/* ---- ScaLAPACK p?gemm ------------------------------------------------------------------------------------
 * Process grid = what BLACS answers for the context id in desc[1] (Cblacs_gridinfo / Cblacs_get; reference
 * src/cosma/blacs.hpp:5-35, scalapack.cpp:3-46, cosma_pxgemm.cpp:57-69). BLACS does not exist on this box, so the grid
 * is an explicit handle: order 'R' (row-major rank numbering) | 'C'; rank r of comm sits at (r / npcol, r % npcol) or
 * (r % nprow, r / nprow). */
COSMA_B200_API int cosma_b200_grid_create(void* comm, char order, int nprow, int npcol, void** grid_out);
COSMA_B200_API int cosma_b200_grid_destroy(void* grid);
COSMA_B200_API int cosma_b200_grid_info(void* grid, int* nprow, int* npcol, int* myrow, int* mycol);
/* pdgemm_ / pzgemm_ (reference src/cosma/pxgemm.h:6-107 -> cosma::pxgemm<T>, cosma_pxgemm.cpp:16-388):
 * sub(C) = alpha*op(sub(A))*op(sub(B)) + beta*sub(C) on 2D block-cyclic matrices. desc = the 9-int ScaLAPACK descriptor
 * ([2] M, [3] N, [4] MB, [5] NB, [6] RSRC, [7] CSRC, [8] LLD; [1], the BLACS context, is ignored in favour of `grid`);
 * ia, ja, ... 1-based. a, b, c: the rank's local arrays, DEVICE or HOST pointers (host arrays are staged through HBM
 * on `stream`; use pinned memory for asynchrony). beta == 0 never reads C. Corner cases as the reference
 * (m|n = 0: return; k = 0 or alpha = 0: sub(C) *= beta). */
COSMA_B200_API int cosma_b200_pdgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const double* a,
                                     int ia, int ja, const int* desca, const double* b, int ib, int jb, const int* descb,
                                     const double* beta, double* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pzgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const double* a,
                                     int ia, int ja, const int* desca, const double* b, int ib, int jb, const int* descb,
                                     const double* beta, double* c, int ic, int jc, const int* descc, void* stream);

/* psgemm_ / pcgemm_: float / complex-float local arrays; alpha and beta passed as doubles (converted to float). */
COSMA_B200_API int cosma_b200_psgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const float* a,
                                     int ia, int ja, const int* desca, const float* b, int ib, int jb, const int* descb,
                                     const double* beta, float* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pcgemm(void* grid, char transa, char transb, int m, int n, int k, const double* alpha, const float* a,
                                     int ia, int ja, const int* desca, const float* b, int ib, int jb, const int* descb,
                                     const double* beta, float* c, int ic, int jc, const int* descc, void* stream);

/* ---- ScaLAPACK p?tran / p?tranu / p?tranc and p?gemr2d (COSTA's wrappers) ---------------------------------
 * p?tran*: sub(C) = beta*sub(C) + alpha*op(sub(A)); sub(C) = C(ic:ic+m-1, jc:jc+n-1), sub(A) = A(ia:ia+n-1, ja:ja+m-1);
 *   op = transpose (p{s,d}tran, p{c,z}tranu) or conjugate transpose (p{c,z}tranc). Reference: libs/COSTA/src/costa/pxtran/
 *   pxtran.h:7-20, pxtranu/pxtranu.h:7-20, pxtranc/pxtranc.h:7-20 -> costa::pxtran_op (pxtran_op/costa_pxtran_op.cpp:14-172).
 * p?gemr2d: sub(C) = sub(A) (m x n) between two block-cyclic distributions, possibly on different process grids of the same
 *   communicator. Reference: pxgemr2d/pxgemr2d.h:7-41 -> costa::pxgemr2d (pxgemr2d/costa_pxgemr2d.cpp:14-168); the trailing
 *   BLACS context argument of the reference is replaced by the two grid handles.
 * Same conventions as p?gemm: `grid` from cosma_b200_grid_create, 9-int descriptors, 1-based sub-matrix origins, local arrays
 * as DEVICE or HOST pointers, beta == 0 never reads C, pure data movement (alpha = 1, beta = 0) is bit-exact. One
 * costa::transform plan per (descriptors, pointers, op, scalars) is cached in the grid. */
COSMA_B200_API int cosma_b200_pxtran(void* grid, char dtype, char op, int m, int n, const double* alpha, const void* a, int ia, int ja,
                                     const int* desca, const double* beta, void* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pxgemr2d(void* grid_a, void* grid_c, char dtype, int m, int n, const void* a, int ia, int ja, const int* desca,
                                       void* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pstran(void* grid, int m, int n, const float* alpha, const float* a, int ia, int ja, const int* desca,
                                     const float* beta, float* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pdtran(void* grid, int m, int n, const double* alpha, const double* a, int ia, int ja, const int* desca,
                                     const double* beta, double* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pctranu(void* grid, int m, int n, const float* alpha, const float* a, int ia, int ja, const int* desca,
                                      const float* beta, float* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pztranu(void* grid, int m, int n, const double* alpha, const double* a, int ia, int ja, const int* desca,
                                      const double* beta, double* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pctranc(void* grid, int m, int n, const float* alpha, const float* a, int ia, int ja, const int* desca,
                                      const float* beta, float* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pztranc(void* grid, int m, int n, const double* alpha, const double* a, int ia, int ja, const int* desca,
                                      const double* beta, double* c, int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_psgemr2d(void* grid_a, void* grid_c, int m, int n, const float* a, int ia, int ja, const int* desca, float* c,
                                       int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pdgemr2d(void* grid_a, void* grid_c, int m, int n, const double* a, int ia, int ja, const int* desca, double* c,
                                       int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pcgemr2d(void* grid_a, void* grid_c, int m, int n, const float* a, int ia, int ja, const int* desca, float* c,
                                       int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_pzgemr2d(void* grid_a, void* grid_c, int m, int n, const double* a, int ia, int ja, const int* desca, double* c,
                                       int ic, int jc, const int* descc, void* stream);
COSMA_B200_API int cosma_b200_grid_last_launches(void* grid);                  /* kernels launched by the last p?tran / p?gemr2d */

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

/* ---- host memory and device housekeeping for C++ / Fortran hosts that do not link the CUDA runtime ----------------
 * The reference pins the buffers of its memory pool so that Tiled-MM's copies are asynchronous (src/cosma/
 * pinned_buffers.cpp:10-40, memory_pool.cpp:153-170, local_multiply.cpp:341-347). host_alloc returns page-locked memory
 * (cudaHostAlloc, portable); host_register / host_unregister pin and unpin caller memory in place. */
COSMA_B200_API int cosma_b200_host_alloc(void** ptr, uint64_t bytes);
COSMA_B200_API int cosma_b200_host_free(void* ptr);
COSMA_B200_API int cosma_b200_host_register(void* ptr, uint64_t bytes);
COSMA_B200_API int cosma_b200_host_unregister(void* ptr);
COSMA_B200_API int cosma_b200_device_count(int* count);
COSMA_B200_API int cosma_b200_set_device(int device);
/* "domain:bus:device.function" of a CUDA device, lower case, as under /sys/bus/pci/devices (cudaDeviceGetPCIBusId): what a host needs
 * to place a rank on the NUMA node of its GPU (an MPI launcher's binding does this for the reference). */
COSMA_B200_API int cosma_b200_device_pci_bus_id(int device, char* out, int out_len);
/* Blocks until everything queued on `stream` (NULL: the default stream) has finished. */
COSMA_B200_API int cosma_b200_stream_synchronize(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COSMA_B200_H */
