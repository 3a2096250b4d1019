#include <costa/erased_layout.hpp>

#include <stdexcept>

namespace costa {

assigned_grid2D assigned_grid2D::transposed() const {
    assigned_grid2D t;
    t.grid.rows_split = grid.cols_split;
    t.grid.cols_split = grid.rows_split;
    t.n_ranks = n_ranks;
    const int nr = grid.n_rows(), nc = grid.n_cols();
    t.owners.resize(owners.size());
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) t.owners[static_cast<size_t>(j) * nr + i] = owners[static_cast<size_t>(i) * nc + j];
    return t;
}

void assigned_grid2D::reorder_ranks(const std::vector<int>& perm) {
    for (auto& o : owners) o = perm[o];
}

erased_layout erased_custom_layout(int rowblocks, int colblocks, const int* rowsplit, const int* colsplit, const int* owners,
                          int nlocalblocks, const int* block_rows, const int* block_cols, void* const* block_data,
                          const std::int64_t* block_ld, char ordering) {
    if (rowblocks < 0 || colblocks < 0 || nlocalblocks < 0) throw std::runtime_error("erased_custom_layout: negative block count");
    if (ordering != 'C' && ordering != 'R') throw std::runtime_error("erased_custom_layout: ordering must be 'C' or 'R'");
    erased_layout l;
    l.ordering = ordering;
    l.grid.grid.rows_split.assign(rowsplit, rowsplit + rowblocks + 1);
    l.grid.grid.cols_split.assign(colsplit, colsplit + colblocks + 1);
    for (int i = 0; i < rowblocks; ++i)
        if (rowsplit[i + 1] < rowsplit[i]) throw std::runtime_error("erased_custom_layout: rowsplit must be non-decreasing");
    for (int j = 0; j < colblocks; ++j)
        if (colsplit[j + 1] < colsplit[j]) throw std::runtime_error("erased_custom_layout: colsplit must be non-decreasing");
    l.grid.owners.assign(owners, owners + static_cast<size_t>(rowblocks) * colblocks);
    int max_owner = 0;
    for (int o : l.grid.owners) {
        if (o < 0) throw std::runtime_error("erased_custom_layout: negative owner");
        if (o > max_owner) max_owner = o;
    }
    l.grid.n_ranks = max_owner + 1;
    l.blocks.resize(nlocalblocks);
    for (int b = 0; b < nlocalblocks; ++b) {
        if (block_rows[b] < 0 || block_rows[b] >= rowblocks || block_cols[b] < 0 || block_cols[b] >= colblocks)
            throw std::runtime_error("erased_custom_layout: local block coordinates outside the grid");
        l.blocks[b] = local_block{block_rows[b], block_cols[b], block_data[b], block_ld[b]};
    }
    return l;
}

std::vector<int> line_split(int begin, int end, int blk_len) {
    const int len = end - begin;
    const int rem = blk_len - begin % blk_len;
    std::vector<int> splits{0};
    if (rem >= len) {
        splits.push_back(len);
        return splits;
    }
    if (rem != 0) splits.push_back(rem);
    const int num_blocks = (len - rem) / blk_len;
    for (int i = 0; i < num_blocks; ++i) splits.push_back(splits.back() + blk_len);
    if (splits.back() != len) splits.push_back(len);
    return splits;
}

int rank_from_grid(int prow, int pcol, int nprow, int npcol, char order) {
    if (prow < 0 || prow >= nprow || pcol < 0 || pcol >= npcol)
        throw std::runtime_error("rank_from_grid: coordinates outside the process grid");
    return (order == 'C' || order == 'c') ? pcol * nprow + prow : prow * npcol + pcol;
}

void rank_to_grid(int rank, int nprow, int npcol, char order, int* prow, int* pcol) {
    if (rank < 0 || rank >= nprow * npcol) throw std::runtime_error("rank_to_grid: rank outside the process grid");
    if (order == 'C' || order == 'c') {
        *prow = rank % nprow;
        *pcol = rank / nprow;
    } else {
        *prow = rank / npcol;
        *pcol = rank % npcol;
    }
}

erased_layout erased_scalapack_layout(int lld, int mat_rows, int mat_cols, int ia, int ja, int sub_m, int sub_n, int mb, int nb,
                                 int nprow, int npcol, char grid_order, int rsrc, int csrc, void* ptr, int elem_bytes,
                                 char data_ordering, int rank) {
    (void)mat_rows;
    (void)mat_cols;
    if (ia < 1 || ja < 1) throw std::runtime_error("erased_scalapack_layout: ia, ja are 1-based");
    if (mb < 1 || nb < 1 || nprow < 1 || npcol < 1) throw std::runtime_error("erased_scalapack_layout: bad block or grid size");
    const int r0 = ia - 1, c0 = ja - 1;
    erased_layout l;
    l.ordering = data_ordering;
    l.grid.n_ranks = nprow * npcol;
    l.grid.grid.rows_split = line_split(r0, r0 + sub_m, mb);
    l.grid.grid.cols_split = line_split(c0, c0 + sub_n, nb);
    const auto& rs = l.grid.grid.rows_split;
    const auto& cs = l.grid.grid.cols_split;
    const int gr = l.grid.grid.n_rows(), gc = l.grid.grid.n_cols();
    l.grid.owners.assign(static_cast<size_t>(gr) * gc, 0);
    // first block row / column of the matrix touched by the sub-matrix, and the process coordinates owning it
    const int first_br = r0 / mb, first_bc = c0 / nb;
    const int src_prow = (first_br % nprow + rsrc) % nprow;
    const int src_pcol = (first_bc % npcol + csrc) % npcol;
    for (int j = 0; j < gc; ++j) {
        const int pcol = (j % npcol + src_pcol) % npcol;
        for (int i = 0; i < gr; ++i) {
            const int prow = (i % nprow + src_prow) % nprow;
            const int owner = rank_from_grid(prow, pcol, nprow, npcol, grid_order);
            l.grid.owners[static_cast<size_t>(i) * gc + j] = owner;
            if (owner != rank) continue;
            // position of matrix block (first_br + i, first_bc + j) inside the owner's local array, plus the offset of
            // the sub-matrix inside that block (non-zero only for border blocks)
            const std::int64_t loc_br = (first_br + i) / nprow, loc_bc = (first_bc + j) / npcol;
            const std::int64_t in_r = r0 + rs[i] - static_cast<std::int64_t>(first_br + i) * mb;
            const std::int64_t in_c = c0 + cs[j] - static_cast<std::int64_t>(first_bc + j) * nb;
            const std::int64_t lr = loc_br * mb + in_r, lc = loc_bc * nb + in_c;
            const std::int64_t off = data_ordering == 'R' ? lc + static_cast<std::int64_t>(lld) * lr
                                                          : lr + static_cast<std::int64_t>(lld) * lc;
            l.blocks.push_back(local_block{i, j, static_cast<char*>(ptr) + off * elem_bytes, lld});
        }
    }
    return l;
}

int numroc(int n, int nb, int iproc, int isrcproc, int nprocs) {
    const int mydist = (nprocs + iproc - isrcproc) % nprocs;
    const int nblocks = n / nb;
    int res = (nblocks / nprocs) * nb;
    const int extra = nblocks % nprocs;
    if (mydist < extra) res += nb;
    else if (mydist == extra) res += n % nb;
    return res;
}

}  // namespace costa
