// costa::transform<T> (reference libs/COSTA/src/costa/grid2grid/transform.cpp:162-282) over the C ABI of libcosma_b200.so.
#include "c_layout.hpp"

#include <cosma/b200_runtime.hpp>
#include <costa/grid2grid/transform.hpp>

#include <complex>
#include <memory>

namespace costa {
namespace {

template <typename T>
void transform_impl(std::vector<layout_ref<T>>& from, std::vector<layout_ref<T>>& to, const char* trans, const T* alpha, const T* beta, MPI_Comm comm) {
    using cosma::b200::check;
    if (from.size() != to.size()) throw std::runtime_error("costa::transform: different numbers of initial and final layouts");
    const int n = static_cast<int>(from.size());
    if (n == 0) return;
    void* handle = cosma::b200::comm_handle(comm);
    std::vector<std::unique_ptr<cosma::b200::c_layout>> F, G;
    std::vector<cosma_b200_layout> cf(n), cg(n);
    std::string of(n, 'C'), og(n, 'C'), ops(n, 'N');
    std::vector<double> a(2 * n), b(2 * n);
    for (int i = 0; i < n; ++i) {
        F.emplace_back(new cosma::b200::c_layout(from[i].get().erased()));
        G.emplace_back(new cosma::b200::c_layout(to[i].get().erased()));
        cf[i] = F.back()->c;
        cg[i] = G.back()->c;
        of[i] = from[i].get().ordering;
        og[i] = to[i].get().ordering;
        ops[i] = trans ? static_cast<char>(std::toupper(trans[i])) : 'N';
        cosma::b200::to_pair(alpha ? alpha[i] : T{1}, &a[2 * i]);
        cosma::b200::to_pair(beta ? beta[i] : T{0}, &b[2 * i]);
    }
    int rank = 0, size = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    void* plan = nullptr;
    check(cosma_b200_transform_plan_create(handle, rank, size, cosma::b200::type_code<T>::value, n, cf.data(), cg.data(), of.c_str(), og.c_str(),
                                           ops.c_str(), a.data(), b.data(), &plan),
          "costa::transform (plan)");
    const int st = cosma_b200_transform_run(plan, nullptr);
    const int sy = cosma_b200_stream_synchronize(nullptr);
    cosma_b200_transform_plan_destroy(plan);
    check(st, "costa::transform (run)");
    check(sy, "costa::transform (synchronize)");
}

}  // namespace

template <typename T>
void transform(grid_layout<T>& initial_layout, grid_layout<T>& final_layout, MPI_Comm comm) {
    std::vector<layout_ref<T>> f{initial_layout}, g{final_layout};
    transform_impl<T>(f, g, nullptr, nullptr, nullptr, comm);
}
template <typename T>
void transform(grid_layout<T>& initial_layout, grid_layout<T>& final_layout, const char trans, const T alpha, const T beta, MPI_Comm comm) {
    std::vector<layout_ref<T>> f{initial_layout}, g{final_layout};
    transform_impl<T>(f, g, &trans, &alpha, &beta, comm);
}
template <typename T>
void transform(std::vector<layout_ref<T>>& initial_layouts, std::vector<layout_ref<T>>& final_layouts, MPI_Comm comm) {
    transform_impl<T>(initial_layouts, final_layouts, nullptr, nullptr, nullptr, comm);
}
template <typename T>
void transform(std::vector<layout_ref<T>>& initial_layouts, std::vector<layout_ref<T>>& final_layouts, const char* trans, const T* alpha,
               const T* beta, MPI_Comm comm) {
    transform_impl<T>(initial_layouts, final_layouts, trans, alpha, beta, comm);
}

#define COSTA_B200_INSTANTIATE(T)                                                                                          \
    template void transform<T>(grid_layout<T>&, grid_layout<T>&, MPI_Comm);                                               \
    template void transform<T>(grid_layout<T>&, grid_layout<T>&, const char, const T, const T, MPI_Comm);                 \
    template void transform<T>(std::vector<layout_ref<T>>&, std::vector<layout_ref<T>>&, MPI_Comm);                       \
    template void transform<T>(std::vector<layout_ref<T>>&, std::vector<layout_ref<T>>&, const char*, const T*, const T*, MPI_Comm);
COSTA_B200_INSTANTIATE(float)
COSTA_B200_INSTANTIATE(double)
COSTA_B200_INSTANTIATE(std::complex<float>)
COSTA_B200_INSTANTIATE(std::complex<double>)
#undef COSTA_B200_INSTANTIATE

}  // namespace costa
