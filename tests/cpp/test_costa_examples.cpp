// The two usage examples of the reference's COSTA (libs/COSTA/examples/example0.cpp, example1.cpp) re-told against this API, as a
// test on exactly 4 ranks: (0) a 4 x 4 matrix in 2 x 2 blocks moves from a row-major to a column-major process grid, with the
// target stored ROW-major inside its blocks; (1) a layout described by hand with costa::custom_layout (grid lines, owners, one
// local block) moves into a block-cyclic one; then the scaled, transposed form final = beta*final + alpha*initial^T. Element
// values are a function of the global coordinates, so validate() can check every rank's share.
#include "check.hpp"

#include <cosma/b200_runtime.hpp>
#include <costa/grid2grid/transform.hpp>
#include <costa/layout.hpp>

#include <vector>

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    MPI_Comm comm = MPI_COMM_WORLD;
    int P = 0, rank = 0;
    MPI_Comm_size(comm, &P);
    MPI_Comm_rank(comm, &rank);
    if (P != 4) {
        if (rank == 0) std::printf("test_costa_examples needs exactly 4 ranks (has %d): skipped\n", P);
        ++check::skipped();
        const int rc = check::finish("test_costa_examples");
        MPI_Finalize();
        return rc;
    }
    auto f = [](int i, int j) -> double { return i + 10.0 * j; };

    {  // example 0
        const int mat_dim = 4, block_size = 2;
        std::vector<double> initial_data(block_size * block_size), final_data(block_size * block_size, -1.0);
        auto init_layout = costa::block_cyclic_layout(mat_dim, mat_dim, block_size, block_size, 1, 1, mat_dim, mat_dim, 2, 2, 'R', 0, 0,
                                                      &initial_data[0], block_size, 'C', rank);
        init_layout.initialize(f);
        auto final_layout = costa::block_cyclic_layout(mat_dim, mat_dim, block_size, block_size, 1, 1, mat_dim, mat_dim, 2, 2, 'C', 0, 0,
                                                       &final_data[0], block_size, 'R', rank);
        costa::transform<double>(init_layout, final_layout, comm);
        CHECK_TRUE(final_layout.validate(f, 0.0));
        // rank r of the column-major grid sits at (r % 2, r / 2): its block starts at global (2 * (r % 2), 2 * (r / 2)), row-major
        const int gi = 2 * (rank % 2), gj = 2 * (rank / 2);
        CHECK_TRUE(final_data[0] == f(gi, gj) && final_data[1] == f(gi, gj + 1) && final_data[2] == f(gi + 1, gj));
    }
    {  // example 1 and the scaled, transposed variant
        const int mat_dim = 10;
        std::vector<int> rowsplit = {0, mat_dim / 2, mat_dim}, colsplit = {0, mat_dim / 2, mat_dim};
        std::vector<int> owners = {0, 1, 2, 3};  // row-major: block (i, j) belongs to rank 2 * i + j
        const int half = mat_dim / 2;
        std::vector<double> initial_data(half * half);
        costa::block_t local_block{&initial_data[0], half, rank / 2, rank % 2};
        auto init_layout = costa::custom_layout<double>(2, 2, &rowsplit[0], &colsplit[0], &owners[0], 1, &local_block, 'C');
        init_layout.initialize(f);
        CHECK_TRUE(init_layout.num_rows() == mat_dim && init_layout.num_blocks_row() == 2 && init_layout.num_ranks() == 4);
        // target: 3 x 2 blocks, block-cyclic on a 2 x 2 column-major grid
        const int bm = 3, bn = 2;
        const int lr = costa::numroc(mat_dim, bm, rank % 2, 0, 2), lc = costa::numroc(mat_dim, bn, rank / 2, 0, 2);
        std::vector<double> final_data(static_cast<size_t>(lr + 1) * lc, 7.0);
        auto final_layout = costa::block_cyclic_layout(mat_dim, mat_dim, bm, bn, 1, 1, mat_dim, mat_dim, 2, 2, 'C', 0, 0, &final_data[0], lr + 1, 'C', rank);
        costa::transform<double>(init_layout, final_layout, comm);
        CHECK_TRUE(final_layout.validate(f, 0.0));
        // final = beta * final + alpha * initial^T
        const double alpha = 0.5, beta = 2.0;
        costa::transform<double>(init_layout, final_layout, 'T', alpha, beta, comm);
        CHECK_TRUE(final_layout.validate([&](int i, int j) { return beta * f(i, j) + alpha * f(j, i); }, 1e-13));
        bool padding_intact = true;
        for (int j = 0; j < lc; ++j) padding_intact = padding_intact && final_data[static_cast<size_t>(j) * (lr + 1) + lr] == 7.0;
        CHECK_TRUE(padding_intact);
    }
    cosma::b200::release_all_comms();
    const int rc = check::finish("test_costa_examples");
    MPI_Finalize();
    return rc;
}
