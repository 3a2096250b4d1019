// Public layout constructors of COSTA with the reference's signatures (libs/COSTA/src/costa/layout.hpp:14-100):
// custom_layout<T>, custom_grid, block_cyclic_layout<T>, block_cyclic_grid. Header-only over the planner in
// <costa/erased_layout.hpp>. Pointers may be host or device memory.
#pragma once
#include <costa/grid2grid/grid_layout.hpp>

#include <stdexcept>

namespace costa {

// a local block as the C / Fortran interfaces describe it
struct block_t {
    void* data;
    int ld;
    int row;
    int col;
};

inline assigned_grid2D custom_grid(const int rowblocks, const int colblocks, const int* rowsplit, const int* colsplit, const int* owners) {
    erased_layout e = erased_custom_layout(rowblocks, colblocks, rowsplit, colsplit, owners, 0, nullptr, nullptr, nullptr, nullptr, 'C');
    int n_ranks = 1;
    for (int o : e.grid.owners) n_ranks = o + 1 > n_ranks ? o + 1 : n_ranks;
    e.grid.n_ranks = n_ranks;
    return e.grid;
}

template <typename T>
grid_layout<T> custom_layout(const int rowblocks, const int colblocks, const int* rowsplit, const int* colsplit, const int* owners,
                             const int nlocalblocks, const block_t* localblocks, const char ordering) {
    assigned_grid2D g = custom_grid(rowblocks, colblocks, rowsplit, colsplit, owners);
    std::vector<block<T>> loc;
    loc.reserve(nlocalblocks);
    for (int b = 0; b < nlocalblocks; ++b) {
        const block_t& cb = localblocks[b];
        if (cb.row < 0 || cb.row >= rowblocks || cb.col < 0 || cb.col >= colblocks) throw std::runtime_error("custom_layout: block coordinates outside the grid");
        loc.emplace_back(g, cb.row, cb.col, static_cast<T*>(cb.data), cb.ld);
    }
    return grid_layout<T>(std::move(g), local_blocks<T>(std::move(loc)), ordering);
}

inline assigned_grid2D block_cyclic_grid(const int m, const int n, const int block_m, const int block_n, const int i, const int j,
                                         const int sub_m, const int sub_n, const int proc_m, const int proc_n, const char rank_grid_ordering,
                                         const int rsrc, const int csrc) {
    erased_layout e = erased_scalapack_layout(/*lld=*/1, m, n, i, j, sub_m, sub_n, block_m, block_n, proc_m, proc_n, rank_grid_ordering, rsrc, csrc,
                                              nullptr, 1, 'C', /*rank=*/-1);
    e.grid.n_ranks = proc_m * proc_n;
    return e.grid;
}

template <typename T>
grid_layout<T> block_cyclic_layout(const int m, const int n, const int block_m, const int block_n, const int i, const int j, const int sub_m,
                                   const int sub_n, const int p_m, const int p_n, const char order, const int rsrc, const int csrc, T* ptr,
                                   const int lld, const char ordering, const int rank) {
    erased_layout e = erased_scalapack_layout(lld, m, n, i, j, sub_m, sub_n, block_m, block_n, p_m, p_n, order, rsrc, csrc, ptr, static_cast<int>(sizeof(T)),
                                              ordering, rank);
    e.grid.n_ranks = p_m * p_n;
    std::vector<block<T>> loc;
    loc.reserve(e.blocks.size());
    for (const auto& b : e.blocks) loc.emplace_back(e.grid, b.bi, b.bj, static_cast<T*>(b.data), static_cast<int>(b.ld));
    assigned_grid2D g = e.grid;
    return grid_layout<T>(std::move(g), local_blocks<T>(std::move(loc)), ordering);
}

}  // namespace costa
