// costa::pxgemr2d<T> -- sub(C) = sub(A) (m x n) between two block-cyclic distributions, possibly on different process grids,
// with the reference's signature (libs/COSTA/src/costa/pxgemr2d/costa_pxgemr2d.hpp:12-24, costa_pxgemr2d.cpp:14-168). The
// grids of desca[1] and descc[1] must be built over the same communicator; ictxt (a context spanning both) is accepted for
// compatibility. Local arrays in host or device memory. One relayout: pack kernel -> NCCL exchange -> unpack kernel.
#pragma once
#include <complex>

namespace costa {
using zdouble_t = std::complex<double>;
using zfloat_t = std::complex<float>;

template <typename T>
void pxgemr2d(const int m, const int n, const T* a, const int ia, const int ja, const int* desca, T* c, const int ic, const int jc, const int* descc,
              const int ictxt);
}  // namespace costa
