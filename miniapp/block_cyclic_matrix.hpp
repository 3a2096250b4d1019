// Shared by pxgemr2d_miniapp and pxtran_miniapp: a block-cyclic matrix set up the ScaLAPACK way (numroc_, descinit_, host-resident
// local array, page-locked as the reference pins its buffers) whose elements are an analytic function of their GLOBAL coordinates, so
// that --test can check a redistribution on every rank without gathering anything and without a second library.
#pragma once
#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <cosma/memory_pool.hpp>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <string>
#include <vector>

extern "C" {
void descinit_(int* desc, const int* m, const int* n, const int* mb, const int* nb, const int* irsrc, const int* icsrc, const int* ictxt,
               const int* lld, int* info);
int numroc_(const int* n, const int* nb, const int* iproc, const int* isrcproc, const int* nprocs);
}

namespace miniapp {

template <typename T> struct real_of { using type = T; };
template <typename T> struct real_of<std::complex<T>> { using type = T; };

// element (i, j) of matrix number `which`: small integers (exact in every type, also after alpha/beta with integer values)
template <typename T> inline T element(int which, long long i, long long j) { return static_cast<T>((7 * i + 13 * j + 5 * which) % 19 - 9); }
template <> inline std::complex<double> element<std::complex<double>>(int which, long long i, long long j) {
    return {static_cast<double>((7 * i + 13 * j + 5 * which) % 19 - 9), static_cast<double>((3 * i + 11 * j + which) % 17 - 8)};
}
template <> inline std::complex<float> element<std::complex<float>>(int which, long long i, long long j) {
    return std::complex<float>(element<std::complex<double>>(which, i, j));
}
template <typename T> inline T conj_if(const T& v, bool) { return v; }
template <typename T> inline std::complex<T> conj_if(const std::complex<T>& v, bool c) { return c ? std::conj(v) : v; }

inline void pair_of(const std::string& s, int out[2]) {
    const auto c = s.find(',');
    out[0] = std::atoi(s.substr(0, c).c_str());
    out[1] = c == std::string::npos ? out[0] : std::atoi(s.substr(c + 1).c_str());
}

// the most square nprow x npcol = P (nprow <= npcol)
inline void square_grid(int P, int grid[2]) {
    int r = 1;
    for (int d = 1; d * d <= P; ++d)
        if (P % d == 0) r = d;
    grid[0] = r;
    grid[1] = P / r;
}

template <typename T>
class block_cyclic_matrix {
  public:
    int desc[9];
    block_cyclic_matrix(cosma::memory_pool<T>& pool, int ctxt, int rows, int cols, int mb, int nb) : pool_(pool), mb_(mb), nb_(nb) {
        cosma::blacs::Cblacs_gridinfo(ctxt, &nprow_, &npcol_, &myrow_, &mycol_);
        const int zero = 0;
        lrows_ = myrow_ >= 0 ? numroc_(&rows, &mb, &myrow_, &zero, &nprow_) : 0;
        lcols_ = myrow_ >= 0 ? numroc_(&cols, &nb, &mycol_, &zero, &npcol_) : 0;
        lld_ = std::max(1, lrows_);
        int info = 0;
        descinit_(desc, &rows, &cols, &mb, &nb, &zero, &zero, &ctxt, &lld_, &info);
        elements_ = static_cast<size_t>(lld_) * std::max(1, lcols_);
        data_ = pool_.allocate(elements_);
    }
    ~block_cyclic_matrix() { pool_.deallocate(data_); }
    block_cyclic_matrix(const block_cyclic_matrix&) = delete;
    block_cyclic_matrix& operator=(const block_cyclic_matrix&) = delete;

    T* data() { return data_; }
    size_t local_elements() const { return static_cast<size_t>(lrows_) * lcols_; }
    template <typename F> void fill(F f) {
        for (int lj = 0; lj < lcols_; ++lj)
            for (int li = 0; li < lrows_; ++li) data_[static_cast<size_t>(lj) * lld_ + li] = f(global(li, mb_, myrow_, nprow_), global(lj, nb_, mycol_, npcol_));
    }
    // number of local elements that differ from want(global row, global column)
    template <typename F> long long mismatches(F want) const {
        long long bad = 0;
        for (int lj = 0; lj < lcols_; ++lj)
            for (int li = 0; li < lrows_; ++li)
                bad += data_[static_cast<size_t>(lj) * lld_ + li] != want(global(li, mb_, myrow_, nprow_), global(lj, nb_, mycol_, npcol_));
        return bad;
    }

  private:
    static long long global(int l, int b, int p, int np) { return (static_cast<long long>(l / b) * np + p) * b + l % b; }
    cosma::memory_pool<T>& pool_;
    int mb_, nb_, nprow_ = 0, npcol_ = 0, myrow_ = -1, mycol_ = -1, lrows_ = 0, lcols_ = 0, lld_ = 1;
    size_t elements_ = 0;
    T* data_ = nullptr;
};

}  // namespace miniapp
