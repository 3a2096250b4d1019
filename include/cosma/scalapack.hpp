// ScaLAPACK descriptor helpers with the reference's names (src/cosma/scalapack.hpp:11-75, scalapack.cpp:3-140).
#pragma once
#include <cosma/blacs.hpp>
#include <costa/grid2grid/scalapack_layout.hpp>

namespace cosma {
namespace scalapack {
struct block_size {
    int rows = 0, cols = 0;
    block_size() = default;
    block_size(int r, int c) : rows(r), cols(c) {}
    explicit block_size(const int* desc) : rows(desc[4]), cols(desc[5]) {}
};
struct global_matrix_size {
    int rows = 0, cols = 0;
    global_matrix_size() = default;
    global_matrix_size(int r, int c) : rows(r), cols(c) {}
    explicit global_matrix_size(const int* desc) : rows(desc[2]), cols(desc[3]) {}
};
struct rank_src {
    int row_src = 0, col_src = 0;
    rank_src() = default;
    rank_src(int r, int c) : row_src(r), col_src(c) {}
    explicit rank_src(const int* desc) : row_src(desc[6]), col_src(desc[7]) {}
};

// row- or column-major numbering of the process grid: where BLACS puts rank 1 (scalapack.cpp:3-16)
costa::scalapack::ordering rank_ordering(int ctxt, int P);
int get_grid_context(const int* desca, const int* descb, const int* descc);
int get_grid_context(const int* desc);
int get_comm_context(const int grid_context);      // Cblacs_get(ctxt, 10, ...)
MPI_Comm get_communicator(const int grid_context);  // Cblacs2sys_handle of the above
int leading_dimension(const int* desc);
int numroc(int n, int nb, int proc_coord, int proc_src, int n_procs);
int min_leading_dimension(int n, int nb, int rank_grid_dim);
int max_leading_dimension(int n, int nb, int rank_grid_dim);
int local_buffer_size(const int* desc);
}  // namespace scalapack
}  // namespace cosma
