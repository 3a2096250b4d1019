// costa::pxgemr2d<T> and costa::pxtran_op<T> (reference libs/COSTA/src/costa/pxgemr2d/costa_pxgemr2d.cpp:14-168,
// pxtran_op/costa_pxtran_op.cpp:14-172): BLACS context -> cached grid handle -> cosma_b200_pxgemr2d / cosma_b200_pxtran.
#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <costa/pxgemr2d/costa_pxgemr2d.hpp>
#include <costa/pxtran_op/costa_pxtran_op.hpp>

#include <cctype>

namespace costa {

template <typename T>
void pxgemr2d(const int m, const int n, const T* a, const int ia, const int ja, const int* desca, T* c, const int ic, const int jc, const int* descc,
              const int /*ictxt*/) {
    if (m == 0 || n == 0) return;
    void* grid_a = cosma::b200::grid_for_blacs_context(cosma::scalapack::get_grid_context(desca));
    void* grid_c = cosma::b200::grid_for_blacs_context(cosma::scalapack::get_grid_context(descc));
    const int st = cosma_b200_pxgemr2d(grid_a, grid_c, cosma::b200::type_code<T>::value, m, n, a, ia, ja, desca, c, ic, jc, descc, nullptr);
    const int sy = cosma_b200_stream_synchronize(nullptr);
    cosma::b200::check(st, "costa::pxgemr2d");
    cosma::b200::check(sy, "costa::pxgemr2d (synchronize)");
}

template <typename T>
void pxtran_op(const int m, const int n, const T alpha, const T* a, const int ia, const int ja, const int* desca, const T beta, T* c, const int ic,
               const int jc, const int* descc, char op) {
    if (m == 0 || n == 0) return;
    if (cosma::scalapack::get_grid_context(desca) != cosma::scalapack::get_grid_context(descc))
        throw std::runtime_error("costa::pxtran_op: A and C must live in the same BLACS context");
    void* grid = cosma::b200::grid_for_blacs_context(cosma::scalapack::get_grid_context(descc));
    double a2[2], b2[2];
    cosma::b200::to_pair(alpha, a2);
    cosma::b200::to_pair(beta, b2);
    const int st = cosma_b200_pxtran(grid, cosma::b200::type_code<T>::value, static_cast<char>(std::toupper(op)), m, n, a2, a, ia, ja, desca, b2, c, ic, jc,
                                     descc, nullptr);
    const int sy = cosma_b200_stream_synchronize(nullptr);
    cosma::b200::check(st, "costa::pxtran_op");
    cosma::b200::check(sy, "costa::pxtran_op (synchronize)");
}

#define COSTA_B200_INSTANTIATE(T)                                                                                              \
    template void pxgemr2d<T>(const int, const int, const T*, const int, const int, const int*, T*, const int, const int, const int*, const int); \
    template void pxtran_op<T>(const int, const int, const T, const T*, const int, const int, const int*, const T, T*, const int, const int,    \
                               const int*, char);
COSTA_B200_INSTANTIATE(float)
COSTA_B200_INSTANTIATE(double)
COSTA_B200_INSTANTIATE(zfloat_t)
COSTA_B200_INSTANTIATE(zdouble_t)
#undef COSTA_B200_INSTANTIATE

}  // namespace costa
