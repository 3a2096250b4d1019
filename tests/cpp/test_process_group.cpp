// CPU-only test of the MPI-name subset over cosma::pg (include/cosma/mpi_compat.hpp): run on N >= 1 processes by
// python -m cosma_b200.launch. No GPU, no CUDA library.
#include "check.hpp"

#include <complex>
#include <numeric>
#include <vector>

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int rank = -1, size = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    CHECK_TRUE(rank >= 0 && rank < size);

    // broadcast from every root
    for (int root = 0; root < size; ++root) {
        unsigned char id[128];
        for (int i = 0; i < 128; ++i) id[i] = rank == root ? static_cast<unsigned char>(i * 3 + root) : 0;
        MPI_Bcast(id, 128, MPI_BYTE, root, MPI_COMM_WORLD);
        bool ok = true;
        for (int i = 0; i < 128; ++i) ok = ok && id[i] == static_cast<unsigned char>(i * 3 + root);
        CHECK_TRUE(ok);
    }
    // gather / allgather
    std::vector<int> all(size, -1);
    int mine = 10 * rank + 1;
    MPI_Allgather(&mine, 1, MPI_INT, all.data(), 1, MPI_INT, MPI_COMM_WORLD);
    for (int r = 0; r < size; ++r) CHECK_TRUE(all[r] == 10 * r + 1);
    // reductions
    double x = rank + 0.5, sum = 0, mx = 0;
    MPI_Allreduce(&x, &sum, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    MPI_Allreduce(&x, &mx, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    CHECK_TRUE(sum == size * (size - 1) / 2.0 + 0.5 * size);
    CHECK_TRUE(mx == size - 0.5);
    std::complex<double> z(rank, -rank), zs;
    MPI_Reduce(&z, &zs, 1, MPI_C_DOUBLE_COMPLEX, MPI_SUM, 0, MPI_COMM_WORLD);
    if (rank == 0) CHECK_TRUE(zs == std::complex<double>(size * (size - 1) / 2.0, -size * (size - 1) / 2.0));
    // point to point with tags arriving out of order
    if (size > 1) {
        if (rank == 1) {
            std::vector<double> a(1000, 1.0), b(3, 2.0);
            MPI_Send(a.data(), 1000, MPI_DOUBLE, 0, 7, MPI_COMM_WORLD);
            MPI_Ssend(b.data(), 3, MPI_DOUBLE, 0, 8, MPI_COMM_WORLD);
        } else if (rank == 0) {
            std::vector<double> a(1000, 0.0), b(3, 0.0);
            MPI_Recv(b.data(), 3, MPI_DOUBLE, 1, 8, MPI_COMM_WORLD, MPI_STATUS_IGNORE);  // the later message first
            MPI_Recv(a.data(), 1000, MPI_DOUBLE, 1, 7, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            CHECK_TRUE(b[2] == 2.0 && a[999] == 1.0);
        }
    }
    // split: even / odd ranks, reversed order inside; sub-communicator traffic must not mix with the world's
    MPI_Comm half = MPI_COMM_NULL;
    MPI_Comm_split(MPI_COMM_WORLD, rank % 2, -rank, &half);
    int hr = -1, hs = 0;
    MPI_Comm_rank(half, &hr);
    MPI_Comm_size(half, &hs);
    CHECK_TRUE(hs == (size + (rank % 2 == 0 ? 1 : 0)) / 2);
    int top = rank;  // the highest world rank of my parity is rank 0 of the half
    MPI_Bcast(&top, 1, MPI_INT, 0, half);
    int expect_top = (size - 1) - (((size - 1) % 2) != (rank % 2) ? 1 : 0);
    CHECK_TRUE(top == expect_top);
    CHECK_TRUE(cosma::comm_key(half) != cosma::comm_key(MPI_COMM_WORLD));
    // MPI_UNDEFINED -> MPI_COMM_NULL (the reference tests cut the first P ranks out of the world this way)
    MPI_Comm first = MPI_COMM_NULL;
    MPI_Comm_split(MPI_COMM_WORLD, rank < 1 ? 0 : MPI_UNDEFINED, rank, &first);
    CHECK_TRUE((rank < 1) == (first != MPI_COMM_NULL));
    MPI_Comm dup = MPI_COMM_NULL;
    MPI_Comm_dup(half, &dup);
    CHECK_TRUE(cosma::comm_key(dup) != cosma::comm_key(half));
    MPI_Barrier(dup);
    MPI_Comm_free(&dup);
    MPI_Comm_free(&half);
    if (first != MPI_COMM_NULL) MPI_Comm_free(&first);
    MPI_Barrier(MPI_COMM_WORLD);
    const int rc = check::finish("test_process_group");
    MPI_Finalize();
    return rc;
}
