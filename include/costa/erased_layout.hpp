// costa:: layouts -- how a distributed matrix is cut into blocks, who owns each block and where the blocks of the
// calling rank sit in (device) memory. Restates the data model of the reference's COSTA
// (libs/COSTA/src/costa/grid2grid/grid2D.hpp, grid_layout.hpp:9-181, block.hpp:64-137, layout.hpp:34-86,
// scalapack_layout.cpp:152-285) in type-erased form: the element type only matters to the kernels, so a layout
// carries byte pointers and leading dimensions in ELEMENTS, and the transform is told the element size.
#pragma once
#include <cstdint>
#include <vector>

namespace costa {

// grid lines: block (i, j) covers rows [rows_split[i], rows_split[i+1]) x cols [cols_split[j], cols_split[j+1])
struct grid2D {
    std::vector<int> rows_split{0};
    std::vector<int> cols_split{0};
    int n_rows() const { return static_cast<int>(rows_split.size()) - 1; }
    int n_cols() const { return static_cast<int>(cols_split.size()) - 1; }
    int total_rows() const { return rows_split.back(); }
    int total_cols() const { return cols_split.back(); }
};

// half-open index range [start, end) (reference grid2grid/interval.hpp)
struct interval {
    int start = 0, end = 0;
    interval() = default;
    interval(int s, int e) : start(s), end(e) {}
    int length() const { return end - start; }
    bool contains(int v) const { return v >= start && v < end; }
};

struct assigned_grid2D {
    grid2D grid;
    std::vector<int> owners;  // row-major: owners[i * n_cols + j] (reference cinterface.cpp:40-45)
    int n_ranks = 1;
    int owner(int i, int j) const { return owners[static_cast<size_t>(i) * grid.n_cols() + j]; }
    // the reference's accessor names (grid2D.hpp)
    int num_ranks() const { return n_ranks; }
    int num_blocks_row() const { return grid.n_rows(); }
    int num_blocks_col() const { return grid.n_cols(); }
    int num_rows() const { return grid.total_rows(); }
    int num_cols() const { return grid.total_cols(); }
    interval rows_interval(int i) const { return interval(grid.rows_split[i], grid.rows_split[i + 1]); }
    interval cols_interval(int j) const { return interval(grid.cols_split[j], grid.cols_split[j + 1]); }
    // the grid of the transposed matrix (reference assigned_grid2D::transpose, grid2D.cpp)
    assigned_grid2D transposed() const;
    // relabel ranks: owner r becomes perm[r] (grid_layout::reorder_ranks)
    void reorder_ranks(const std::vector<int>& perm);
};

// a block of the calling rank: grid coordinates + where its (bi, bj) storage starts
struct local_block {
    int bi = 0, bj = 0;
    void* data = nullptr;      // device (or host, for planning-only tests) address of element (0, 0) of the block
    std::int64_t ld = 0;       // leading dimension in elements (column stride if ordering 'C', row stride if 'R')
};

struct erased_layout {
    assigned_grid2D grid;
    std::vector<local_block> blocks;  // blocks owned by the calling rank
    char ordering = 'C';              // storage order of every local block
    int num_rows() const { return grid.grid.total_rows(); }
    int num_cols() const { return grid.grid.total_cols(); }
};

// erased_custom_layout (reference custom_layout, layout.hpp:34-48): arrays in the shape of the C interface's struct layout
erased_layout erased_custom_layout(int rowblocks, int colblocks, const int* rowsplit, const int* colsplit, const int* owners,
                          int nlocalblocks, const int* block_rows, const int* block_cols, void* const* block_data,
                          const std::int64_t* block_ld, char ordering);

// split points of [begin, end) cut at multiples of blk_len, shifted to start at 0 (scalapack_layout.cpp:152-177)
std::vector<int> line_split(int begin, int end, int blk_len);

// rank <-> coordinates in a process grid ordered 'R' (row-major) or 'C' (column-major) (scalapack_layout.cpp:11-57)
int rank_from_grid(int prow, int pcol, int nprow, int npcol, char order);
void rank_to_grid(int rank, int nprow, int npcol, char order, int* prow, int* pcol);

// Block-cyclic (ScaLAPACK) layout of the sub-matrix sub(A) = A(ia : ia+sub_m-1, ja : ja+sub_n-1), 1-based ia/ja
// (reference get_scalapack_layout, scalapack_layout.cpp:178-285). ptr = local array of the WHOLE matrix A on `rank`
// with leading dimension lld; elem_bytes turns element offsets into addresses.
erased_layout erased_scalapack_layout(int lld, int mat_rows, int mat_cols, int ia, int ja, int sub_m, int sub_n, int mb, int nb,
                                 int nprow, int npcol, char grid_order, int rsrc, int csrc, void* ptr, int elem_bytes,
                                 char data_ordering, int rank);

// ScaLAPACK NUMROC: rows/cols of a block-cyclic dimension that land on process coordinate iproc
int numroc(int n, int nb, int iproc, int isrcproc, int nprocs);

}  // namespace costa
