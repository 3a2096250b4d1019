// costa::get_scalapack_layout<T> with the reference's argument list (libs/COSTA/src/costa/grid2grid/scalapack_layout.hpp,
// scalapack_layout.cpp:178-285): the layout of sub(A) = A(ia:ia+sub_m-1, ja:ja+sub_n-1) of a block-cyclic matrix.
#pragma once
#include <costa/layout.hpp>

namespace costa {
namespace scalapack {
enum ordering { row_major, column_major };
struct int_pair {
    int row = 0, col = 0;
    int_pair() = default;
    int_pair(int r, int c) : row(r), col(c) {}
};
using matrix_dim = int_pair;
using block_dim = int_pair;
using rank_grid_coord = int_pair;
using rank_decomposition = int_pair;
using elem_grid_coord = int_pair;
}  // namespace scalapack

template <typename T>
grid_layout<T> get_scalapack_layout(int lld, scalapack::matrix_dim m_dim, scalapack::elem_grid_coord ij, scalapack::matrix_dim subm_dim,
                                    scalapack::block_dim b_dim, scalapack::rank_decomposition r_grid, scalapack::ordering rank_grid_ordering,
                                    scalapack::rank_grid_coord rank_src, T* ptr, const int rank, const char data_ordering = 'C') {
    return block_cyclic_layout<T>(m_dim.row, m_dim.col, b_dim.row, b_dim.col, ij.row, ij.col, subm_dim.row, subm_dim.col, r_grid.row, r_grid.col,
                                  rank_grid_ordering == scalapack::row_major ? 'R' : 'C', rank_src.row, rank_src.col, ptr, lld, data_ordering, rank);
}
}  // namespace costa
