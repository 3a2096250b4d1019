// cosma::pxgemm<T> (reference src/cosma/cosma_pxgemm.cpp:16-388): the BLACS context of the descriptors gives the process
// grid, its numbering and the communicator; everything else -- corner cases, relayout of op(sub(A)), op(sub(B)) into
// COSMA's layout, the multiply, the relayout into sub(C) with (alpha, beta) -- happens behind cosma_b200_p?gemm.
#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <cosma/environment_variables.hpp>

#include <algorithm>

#include <map>
#include <mutex>

namespace cosma {
namespace {
std::mutex g_mu;
struct cached_grid {
    void* handle = nullptr;
    int nprow = 0, npcol = 0;
    char order = 'R';
    unsigned long long comm = 0;
};
std::map<int, cached_grid> g_grids;  // BLACS grid context -> cosma_b200 grid handle and what it was built from

}  // namespace

void* b200::grid_for_blacs_context(int ctxt) {
    std::lock_guard<std::mutex> lock(g_mu);
    // BLACS answers are cheap; asking every time lets a context id that was released and handed out again (real BLACS
    // libraries recycle them) be recognised instead of serving a stale grid
    int nprow = 0, npcol = 0, myrow = 0, mycol = 0;
    blacs::Cblacs_gridinfo(ctxt, &nprow, &npcol, &myrow, &mycol);
    MPI_Comm comm = scalapack::get_communicator(ctxt);
    int P = 1;
    MPI_Comm_size(comm, &P);
    const char order = scalapack::rank_ordering(ctxt, P) == costa::scalapack::ordering::row_major ? 'R' : 'C';
    auto it = g_grids.find(ctxt);
    if (it != g_grids.end()) {
        const cached_grid& g = it->second;
        if (g.nprow == nprow && g.npcol == npcol && g.order == order && g.comm == comm_key(comm)) return g.handle;
        cosma_b200_grid_destroy(g.handle);
        g_grids.erase(it);
    }
    cached_grid g;
    g.nprow = nprow; g.npcol = npcol; g.order = order; g.comm = comm_key(comm);
    b200::check(cosma_b200_grid_create(b200::comm_handle(comm), order, nprow, npcol, &g.handle), "pxgemm (process grid)");
    g_grids[ctxt] = g;
    return g.handle;
}

bool problem_below_dim_threshold(int m, int n, int k) {
    static const int threshold = get_cosma_dim_threshold();
    return std::min(m, std::min(n, k)) < threshold;
}

void pxgemm_release_grids() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_grids) cosma_b200_grid_destroy(kv.second.handle);
    g_grids.clear();
}

template <typename T>
void pxgemm(const char trans_a, const char trans_b, const int m, const int n, const int k, const T alpha, const T* a, const int ia, const int ja,
            const int* desca, const T* b, const int ib, const int jb, const int* descb, const T beta, T* c, const int ic, const int jc,
            const int* descc) {
    if (m == 0 || n == 0) return;
    void* grid = b200::grid_for_blacs_context(scalapack::get_grid_context(desca, descb, descc));
    double a2[2], b2[2];
    b200::to_pair(alpha, a2);
    b200::to_pair(beta, b2);
    int st;
    switch (b200::type_code<T>::value) {
        case 's': st = cosma_b200_psgemm(grid, trans_a, trans_b, m, n, k, a2, reinterpret_cast<const float*>(a), ia, ja, desca, reinterpret_cast<const float*>(b), ib, jb, descb, b2, reinterpret_cast<float*>(c), ic, jc, descc, nullptr); break;
        case 'd': st = cosma_b200_pdgemm(grid, trans_a, trans_b, m, n, k, a2, reinterpret_cast<const double*>(a), ia, ja, desca, reinterpret_cast<const double*>(b), ib, jb, descb, b2, reinterpret_cast<double*>(c), ic, jc, descc, nullptr); break;
        case 'c': st = cosma_b200_pcgemm(grid, trans_a, trans_b, m, n, k, a2, reinterpret_cast<const float*>(a), ia, ja, desca, reinterpret_cast<const float*>(b), ib, jb, descb, b2, reinterpret_cast<float*>(c), ic, jc, descc, nullptr); break;
        default: st = cosma_b200_pzgemm(grid, trans_a, trans_b, m, n, k, a2, reinterpret_cast<const double*>(a), ia, ja, desca, reinterpret_cast<const double*>(b), ib, jb, descb, b2, reinterpret_cast<double*>(c), ic, jc, descc, nullptr); break;
    }
    const int sy = cosma_b200_stream_synchronize(nullptr);
    b200::check(st, "cosma::pxgemm");
    b200::check(sy, "cosma::pxgemm (synchronize)");
}

template void pxgemm<float>(const char, const char, const int, const int, const int, const float, const float*, const int, const int, const int*,
                            const float*, const int, const int, const int*, const float, float*, const int, const int, const int*);
template void pxgemm<double>(const char, const char, const int, const int, const int, const double, const double*, const int, const int, const int*,
                             const double*, const int, const int, const int*, const double, double*, const int, const int, const int*);
template void pxgemm<zfloat_t>(const char, const char, const int, const int, const int, const zfloat_t, const zfloat_t*, const int, const int,
                               const int*, const zfloat_t*, const int, const int, const int*, const zfloat_t, zfloat_t*, const int, const int,
                               const int*);
template void pxgemm<zdouble_t>(const char, const char, const int, const int, const int, const zdouble_t, const zdouble_t*, const int, const int,
                                const int*, const zdouble_t*, const int, const int, const int*, const zdouble_t, zdouble_t*, const int, const int,
                                const int*);

}  // namespace cosma
