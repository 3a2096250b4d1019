// Interposition semantics of libcosma_pxgemm.so (reference src/cosma/pxgemm.cpp:8-136, interpose.h:33-51, CMakeLists.txt:71-87): the
// library defines the ScaLAPACK names itself, so an application linked against it AND a ScaLAPACK gets COSMA's p?gemm; problems with
// min(m, n, k) below COSMA_DIM_THRESHOLD are handed on to the next pdgemm_ in link order (here tests/cpp/fake_scalapack.c, which only
// stamps a marker). Run with COSMA_DIM_THRESHOLD=64 on one rank.
#include "check.hpp"

#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <cosma/pxgemm.h>

#include <vector>

extern "C" void descinit_(int*, const int*, const int*, const int*, const int*, const int*, const int*, const int*, const int*, int*);

static double run(int dim) {
    int ctxt = 0, zero = 0, one = 1, info = 0, blk = 8, lld = dim;
    char R = 'R', N = 'N';
    cosma::blacs::Cblacs_get(0, 0, &ctxt);
    cosma::blacs::Cblacs_gridinit(&ctxt, &R, 1, 1);
    int desc[9];
    descinit_(desc, &dim, &dim, &blk, &blk, &zero, &zero, &ctxt, &lld, &info);
    std::vector<double> a(static_cast<size_t>(dim) * dim, 1.0), b(a), c(a.size(), 0.0);
    const double alpha = 1.0, beta = 0.0;
    pdgemm_(&N, &N, &dim, &dim, &dim, &alpha, a.data(), &one, &one, desc, b.data(), &one, &one, desc, &beta, c.data(), &one, &one, desc);
    cosma::pxgemm_release_grids();
    cosma::blacs::Cblacs_gridexit(ctxt);
    return c[0];
}

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    CHECK_TRUE(cosma::problem_below_dim_threshold(8, 100, 100));
    CHECK_TRUE(!cosma::problem_below_dim_threshold(64, 64, 64));
    CHECK_TRUE(run(8) == 424242.0);    // below the threshold: served by the next pdgemm_ in link order
    CHECK_TRUE(run(96) == 96.0);       // above: served here (ones x ones = dim)
    cosma::b200::release_all_comms();
    const int rc = check::finish("test_interpose");
    MPI_Finalize();
    return rc;
}
