// CosmaMatrix<T> (reference src/cosma/matrix.cpp:9-470): Mapper + the rank's local storage.
#include <cosma/b200_runtime.hpp>
#include <cosma/matrix.hpp>

#include <algorithm>
#include <complex>

namespace cosma {

template <typename T>
CosmaMatrix<T>::CosmaMatrix(cosma_context<T>* ctxt, char label, const Strategy& strategy, int rank, bool dry_run)
    : ctxt_(ctxt), mapper_(label, strategy, rank), rank_(rank), label_(mapper_.label()), m_(mapper_.m()), n_(mapper_.n()),
      P_(static_cast<size_t>(mapper_.P())) {
    if (!dry_run) allocate();
}

template <typename T>
CosmaMatrix<T>::CosmaMatrix(cosma_context<T>* ctxt, Mapper&& mapper, int rank, bool dry_run)
    : ctxt_(ctxt), mapper_(std::move(mapper)), rank_(rank), label_(mapper_.label()), m_(mapper_.m()), n_(mapper_.n()),
      P_(static_cast<size_t>(mapper_.P())) {
    mapper_.reorder_rank(rank);
    if (!dry_run) allocate();
}

template <typename T>
CosmaMatrix<T>::CosmaMatrix(std::unique_ptr<cosma_context<T>>& ctxt, char label, const Strategy& strategy, int rank, bool dry_run)
    : CosmaMatrix(ctxt.get(), label, strategy, rank, dry_run) {}
template <typename T>
CosmaMatrix<T>::CosmaMatrix(std::unique_ptr<cosma_context<T>>& ctxt, Mapper&& mapper, int rank, bool dry_run)
    : CosmaMatrix(ctxt.get(), std::move(mapper), rank, dry_run) {}
template <typename T>
CosmaMatrix<T>::CosmaMatrix(char label, const Strategy& strategy, int rank, bool dry_run)
    : CosmaMatrix(get_context_instance<T>(), label, strategy, rank, dry_run) {}
template <typename T>
CosmaMatrix<T>::CosmaMatrix(Mapper&& mapper, int rank, bool dry_run) : CosmaMatrix(get_context_instance<T>(), std::move(mapper), rank, dry_run) {}

template <typename T>
CosmaMatrix<T>::~CosmaMatrix() {
    if (data_ && ctxt_) ctxt_->get_memory_pool().deallocate(data_);
}

template <typename T>
void CosmaMatrix<T>::allocate() {
    if (data_ || rank_ < 0 || static_cast<size_t>(rank_) >= P_) return;  // idle ranks own nothing (matrix.cpp:24-30)
    const size_t n = matrix_size();
    data_ = ctxt_->get_memory_pool().allocate(n);
    std::fill(data_, data_ + n, T{0});
}

template <typename T>
size_t CosmaMatrix<T>::matrix_size() const {
    return matrix_size(rank_);
}
template <typename T>
size_t CosmaMatrix<T>::matrix_size(int rank) const {
    if (rank < 0 || static_cast<size_t>(rank) >= P_) return 0;
    return mapper_.initial_size(rank);
}

template <typename T>
std::pair<int, int> CosmaMatrix<T>::local_coordinates(int gi, int gj) {
    const auto lr = mapper_.local_coordinates(gi, gj);
    return {static_cast<int>(lr.first), lr.second};
}
template <typename T>
std::pair<int, int> CosmaMatrix<T>::global_coordinates(int local_index, int rank) {
    return mapper_.global_coordinates(local_index, rank);
}
template <typename T>
std::pair<int, int> CosmaMatrix<T>::global_coordinates(int local_index) {
    return mapper_.global_coordinates(local_index);
}

template <typename T>
costa::grid_layout<T> CosmaMatrix<T>::get_grid_layout() {
    costa::assigned_grid2D g;
    g.grid.rows_split = mapper_.row_split();
    g.grid.cols_split = mapper_.col_split();
    g.n_ranks = static_cast<int>(P_);
    const auto owners = mapper_.grid_owners();
    const int nr = g.grid.n_rows(), nc = g.grid.n_cols();
    g.owners.resize(static_cast<size_t>(nr) * nc);
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) g.owners[static_cast<size_t>(i) * nc + j] = owners[i][j];
    std::vector<costa::block<T>> loc;
    if (rank_ >= 0 && static_cast<size_t>(rank_) < P_) {
        const auto blocks = mapper_.local_blocks();
        const auto& offs = mapper_.local_blocks_offsets();
        for (size_t b = 0; b < blocks.size(); ++b) {
            costa::interval rows(blocks[b].rows.first(), blocks[b].rows.last() + 1), cols(blocks[b].cols.first(), blocks[b].cols.last() + 1);
            loc.emplace_back(g, rows, cols, data_ ? data_ + offs[b] : nullptr, rows.length());
        }
    }
    return costa::grid_layout<T>(std::move(g), costa::local_blocks<T>(std::move(loc)), 'C');
}

template <typename T>
std::vector<size_t> CosmaMatrix<T>::required_memory() {
    std::vector<size_t> out;
    if (rank_ < 0 || static_cast<size_t>(rank_) >= P_) return out;
    out.push_back(matrix_size());
    if (ctxt_ && ctxt_->plan() && ctxt_->registered_strategy() == mapper_.strategy()) {
        const int x = label_ == 'A' ? 0 : (label_ == 'B' ? 1 : 2);
        out.push_back(static_cast<size_t>(cosma_b200_plan_arena_elements(ctxt_->plan(), x)));
    }
    return out;
}

template class CosmaMatrix<float>;
template class CosmaMatrix<double>;
template class CosmaMatrix<std::complex<float>>;
template class CosmaMatrix<std::complex<double>>;

}  // namespace cosma
