// cosma::multiply / cosma::multiply_using_layout with the reference's signatures (src/cosma/multiply.hpp:31-54,
// multiply.cpp:78-314): C = alpha * op(A) * op(B) + beta * C.
#pragma once
#include <cosma/context.hpp>
#include <cosma/matrix.hpp>
#include <cosma/mpi_compat.hpp>
#include <cosma/strategy.hpp>
#include <costa/grid2grid/transform.hpp>

namespace cosma {

// Matrices in arbitrary grid-like layouts (host or device blocks): relayout into COSMA's layout, multiply, relayout back
// with (alpha, beta). Collective over comm.
template <typename Scalar>
void multiply_using_layout(costa::grid_layout<Scalar>& A_layout, costa::grid_layout<Scalar>& B_layout, costa::grid_layout<Scalar>& C_layout,
                           Scalar alpha, Scalar beta, char transa, char transb, MPI_Comm comm);

// Matrices in COSMA's native layout for `strategy`. Ranks >= strategy.P return at once; m, n or k == 0 returns at once
// (multiply.cpp:252-260). Collective over the first strategy.P ranks of comm only.
template <typename Scalar>
void multiply(CosmaMatrix<Scalar>& A, CosmaMatrix<Scalar>& B, CosmaMatrix<Scalar>& C, const Strategy& strategy, MPI_Comm comm, Scalar alpha,
              Scalar beta);
template <typename Scalar>
void multiply(cosma_context<Scalar>* ctx, CosmaMatrix<Scalar>& A, CosmaMatrix<Scalar>& B, CosmaMatrix<Scalar>& C, const Strategy& strategy,
              MPI_Comm comm, Scalar alpha, Scalar beta);

}  // namespace cosma
