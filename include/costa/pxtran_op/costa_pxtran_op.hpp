// costa::pxtran_op<T> -- sub(C) = beta * sub(C) + alpha * op(sub(A)), sub(C) m x n and sub(A) n x m, op = 'T' or 'C': what
// p?tran / p?tranu / p?tranc compute (reference libs/COSTA/src/costa/pxtran_op/costa_pxtran_op.hpp:12-27,
// costa_pxtran_op.cpp:14-172). Local arrays in host or device memory; beta == 0 never reads C.
#pragma once
#include <complex>

namespace costa {
template <typename T>
void pxtran_op(const int m, const int n, const T alpha, const T* a, const int ia, const int ja, const int* desca, const T beta, T* c, const int ic,
               const int jc, const int* descc, char op);
}  // namespace costa
