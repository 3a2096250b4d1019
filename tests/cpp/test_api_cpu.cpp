// CPU-only test of the host-side C++ API (no GPU needed): Strategy / CosmaMatrix in dry-run mode -- the coordinate maps the
// reference's tests/mapper.cpp:408-562 pin for (m, n, k, P) = (8, 4, 2, 4), steps pm2, sm2, pn2 -- the native layout as a
// COSTA grid, block_cyclic_layout's owners / local blocks, and the loud failure of every compute entry point without a GPU.
#include "check.hpp"

#include <cosma/multiply.hpp>
#include <cosma_b200.h>
#include <costa/layout.hpp>

#include <set>

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    std::vector<int> divs = {2, 2, 2};
    std::string dims = "mmn", types = "psp";
    cosma::Strategy strategy(8, 4, 2, 4, divs, dims, types);
    CHECK_TRUE(strategy.to_string() == "pm2,sm2,pn2");
    auto ctx = cosma::make_context<double>();
    // sizes per rank: A {4,4,4,4}, B {2,2,2,2}, C {8,8,8,8} (tests/mapper.cpp:470-476)
    const size_t want_size[3] = {4, 2, 8};
    const char labels[3] = {'A', 'B', 'C'};
    for (int x = 0; x < 3; ++x) {
        for (int rank = 0; rank < 4; ++rank) {
            cosma::CosmaMatrix<double> M(ctx, labels[x], strategy, rank, /*dry_run=*/true);
            CHECK_TRUE(M.matrix_pointer() == nullptr);
            CHECK_TRUE(M.matrix_size() == want_size[x]);
            // local -> global -> local round trip, and every global element has exactly one home
            for (size_t l = 0; l < M.matrix_size(); ++l) {
                int gi, gj, li, lr;
                std::tie(gi, gj) = M.global_coordinates(static_cast<int>(l));
                std::tie(li, lr) = M.local_coordinates(gi, gj);
                CHECK_TRUE(gi >= 0 && gi < M.m() && gj >= 0 && gj < M.n() && lr == rank && li == static_cast<int>(l));
            }
            auto layout = M.get_grid_layout();
            CHECK_TRUE(layout.num_rows() == M.m() && layout.num_cols() == M.n() && layout.num_ranks() == 4);
            size_t covered = 0;
            for (size_t b = 0; b < layout.blocks.num_blocks(); ++b) covered += layout.blocks.get_block(b).total_size();
            CHECK_TRUE(covered == M.matrix_size());
        }
    }
    // ranks beyond strategy.P own nothing
    {
        cosma::Strategy two(64, 64, 64, 2);
        cosma::CosmaMatrix<float> idle(nullptr, 'A', two, 1, true);
        CHECK_TRUE(idle.matrix_size(5) == 0);
    }
    // block-cyclic layout: 10 x 7 in 4 x 3 blocks on a 2 x 2 row-major grid, sub-matrix = everything
    {
        std::vector<double> local(100, 0.0);
        auto L = costa::block_cyclic_layout<double>(10, 7, 4, 3, 1, 1, 10, 7, 2, 2, 'R', 0, 0, local.data(), 6, 'C', 0);
        CHECK_TRUE(L.num_blocks_row() == 3 && L.num_blocks_col() == 3);
        CHECK_TRUE(L.grid.owner(0, 0) == 0 && L.grid.owner(0, 1) == 1 && L.grid.owner(1, 0) == 2 && L.grid.owner(1, 1) == 3 && L.grid.owner(2, 2) == 0);
        CHECK_TRUE(L.blocks.num_blocks() == 4);  // (0,0), (0,2), (2,0), (2,2)
        // rank 0 owns rows {0..3, 8..9} and cols {0..2, 6}: 6 x 4 elements, column-major with lld 6
        size_t elems = 0;
        for (size_t b = 0; b < L.blocks.num_blocks(); ++b) elems += L.blocks.get_block(b).total_size();
        CHECK_TRUE(elems == 24);
        L.initialize([](int i, int j) { return 100.0 * i + j; });
        CHECK_TRUE(local[0] == 0.0 && local[3] == 300.0 && local[4] == 800.0 && local[6] == 1.0 && local[3 * 6 + 5] == 906.0);
        CHECK_TRUE(L.validate([](int i, int j) { return 100.0 * i + j; }, 0.0));
        // the transposed view of the same memory: element (i, j) is the old (j, i)
        L.transpose();
        CHECK_TRUE(L.num_rows() == 7 && L.num_cols() == 10 && L.ordering == 'R' && L.grid.owner(1, 0) == 1 && L.grid.owner(0, 1) == 2);
        CHECK_TRUE(L.validate([](int i, int j) { return 100.0 * j + i; }, 0.0));
        L.transpose();
        CHECK_TRUE(L.ordering == 'C' && L.validate([](int i, int j) { return 100.0 * i + j; }, 0.0));
    }
    // no GPU on this box: compute entry points must fail loudly, never fall back
    {
        int ndev = 0;
        const bool have_gpu = cosma_b200_device_count(&ndev) == COSMA_B200_OK && ndev > 0;
        if (!have_gpu) {
            bool threw = false;
            try {
                cosma::Strategy s1(16, 16, 16, 1);
                cosma::CosmaMatrix<double> A(ctx, 'A', s1, 0);  // page-locked allocation needs CUDA
                (void)A;
            } catch (const std::runtime_error&) { threw = true; }
            CHECK_TRUE(threw);
        }
    }
    const int rc = check::finish("test_api_cpu");
    MPI_Finalize();
    return rc;
}
