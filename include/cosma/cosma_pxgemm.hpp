// cosma::pxgemm<T> -- p?gemm on 2D block-cyclic matrices, reference signature (src/cosma/cosma_pxgemm.hpp:15-33,
// cosma_pxgemm.cpp:16-388): sub(C) = alpha * op(sub(A)) * op(sub(B)) + beta * sub(C). The process grid and communicator
// come from BLACS for the context in desc[1]. a, b, c: the rank's local arrays in host or device memory.
#pragma once
#include <cosma/scalapack.hpp>

#include <complex>

namespace cosma {
using zdouble_t = std::complex<double>;
using zfloat_t = std::complex<float>;

template <typename T>
void pxgemm(const char trans_a, const char trans_b, const int m, const int n, const int k, const T alpha, const T* a, const int ia, const int ja,
            const int* desca, const T* b, const int ib, const int jb, const int* descb, const T beta, T* c, const int ic, const int jc,
            const int* descc);

// min(m, n, k) < COSMA_DIM_THRESHOLD (reference is_problem_too_small, cosma_pxgemm.cpp:652-655); the threshold defaults to 0
bool problem_below_dim_threshold(int m, int n, int k);

// releases the grid handles cached per BLACS context (call before Cblacs_gridexit / MPI_Finalize)
void pxgemm_release_grids();

namespace b200 {
// the cosma_b200 process-grid handle of a BLACS grid context (shape, numbering and communicator asked from BLACS as in
// cosma_pxgemm.cpp:57-69), created on first use and cached; shared by cosma::pxgemm and costa::pxgemr2d / pxtran_op
void* grid_for_blacs_context(int ctxt);
}  // namespace b200
}  // namespace cosma
