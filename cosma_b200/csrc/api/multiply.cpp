// cosma::multiply / cosma::multiply_using_layout (reference src/cosma/multiply.cpp:78-314) over the C ABI.
#include "c_layout.hpp"

#include <cosma/b200_runtime.hpp>
#include <cosma/multiply.hpp>

#include <complex>

namespace cosma {

template <typename Scalar>
void multiply(cosma_context<Scalar>* ctx, CosmaMatrix<Scalar>& A, CosmaMatrix<Scalar>& B, CosmaMatrix<Scalar>& C, const Strategy& strategy,
              MPI_Comm comm, Scalar alpha, Scalar beta) {
    if (strategy.m == 0 || strategy.n == 0 || strategy.k == 0) return;  // multiply.cpp:252-254
    if (comm == MPI_COMM_NULL) return;
    int rank = 0;
    MPI_Comm_rank(comm, &rank);
    if (rank >= static_cast<int>(strategy.P)) return;  // idle ranks take no part at all (multiply.cpp:258-260)
    // everything below is collective over the first strategy.P ranks only, as in the reference (communicator.cpp:282-343)
    ctx->register_state(b200::active_comm(comm, static_cast<int>(strategy.P)), strategy);
    if (A.m() != strategy.m || A.n() != strategy.k || B.m() != strategy.k || B.n() != strategy.n || C.m() != strategy.m || C.n() != strategy.n)
        throw std::runtime_error("cosma::multiply: matrix dimensions do not match the strategy");
    double a2[2], b2[2];
    b200::to_pair(alpha, a2);
    b200::to_pair(beta, b2);
    b200::trace("multiply: queue");
    b200::check(cosma_b200_multiply_host(ctx->plan(), a2, b2, A.matrix_pointer(), B.matrix_pointer(), C.matrix_pointer(), ctx->stream()),
                "cosma::multiply");
    b200::trace("multiply: synchronize");
    b200::check(cosma_b200_stream_synchronize(ctx->stream()), "cosma::multiply (synchronize)");
    b200::trace("multiply: done");
}

template <typename Scalar>
void multiply(CosmaMatrix<Scalar>& A, CosmaMatrix<Scalar>& B, CosmaMatrix<Scalar>& C, const Strategy& strategy, MPI_Comm comm, Scalar alpha,
              Scalar beta) {
    multiply(A.get_context(), A, B, C, strategy, comm, alpha, beta);
}

template <typename Scalar>
void multiply_using_layout(costa::grid_layout<Scalar>& A, costa::grid_layout<Scalar>& B, costa::grid_layout<Scalar>& C, Scalar alpha, Scalar beta,
                           char transa, char transb, MPI_Comm comm) {
    if (comm == MPI_COMM_NULL) return;
    if (A.ordering != 'C' || B.ordering != 'C' || C.ordering != 'C')
        throw std::runtime_error("cosma::multiply_using_layout: blocks must be column-major (the C interface's convention, cinterface.hpp:16-41)");
    void* handle = b200::comm_handle(comm);
    b200::c_layout cA(A.erased()), cB(B.erased()), cC(C.erased());
    double a2[2], b2[2];
    b200::to_pair(alpha, a2);
    b200::to_pair(beta, b2);
    int st;
    switch (b200::type_code<Scalar>::value) {
        case 's': st = cosma_b200_smultiply_using_layout(handle, &transa, &transb, a2, &cA.c, &cB.c, b2, &cC.c, nullptr); break;
        case 'd': st = cosma_b200_dmultiply_using_layout(handle, &transa, &transb, a2, &cA.c, &cB.c, b2, &cC.c, nullptr); break;
        case 'c': st = cosma_b200_cmultiply_using_layout(handle, &transa, &transb, a2, &cA.c, &cB.c, b2, &cC.c, nullptr); break;
        default: st = cosma_b200_zmultiply_using_layout(handle, &transa, &transb, a2, &cA.c, &cB.c, b2, &cC.c, nullptr); break;
    }
    const int sy = cosma_b200_stream_synchronize(nullptr);
    b200::check(st, "cosma::multiply_using_layout");
    b200::check(sy, "cosma::multiply_using_layout (synchronize)");
}

#define COSMA_B200_INSTANTIATE(T)                                                                                                       \
    template void multiply<T>(cosma_context<T>*, CosmaMatrix<T>&, CosmaMatrix<T>&, CosmaMatrix<T>&, const Strategy&, MPI_Comm, T, T);     \
    template void multiply<T>(CosmaMatrix<T>&, CosmaMatrix<T>&, CosmaMatrix<T>&, const Strategy&, MPI_Comm, T, T);                        \
    template void multiply_using_layout<T>(costa::grid_layout<T>&, costa::grid_layout<T>&, costa::grid_layout<T>&, T, T, char, char, MPI_Comm);
COSMA_B200_INSTANTIATE(float)
COSMA_B200_INSTANTIATE(double)
COSMA_B200_INSTANTIATE(std::complex<float>)
COSMA_B200_INSTANTIATE(std::complex<double>)
#undef COSMA_B200_INSTANTIATE

}  // namespace cosma
