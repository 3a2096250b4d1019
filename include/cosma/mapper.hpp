// cosma::Mapper -- the initial (and, for C, final) data layout induced by a Strategy: which rank owns which
// column-major blocks of A, B or C, in which order they sit in the rank's local buffer, and the global <-> local
// coordinate maps. Same public surface and identical block lists as the reference (src/cosma/mapper.hpp:21-126,
// mapper.cpp:5-443); local offsets are 64-bit (the reference's int overflows at 2^31 elements per rank).
#pragma once
#include <cosma/interval.hpp>
#include <cosma/strategy.hpp>

#include <cstdint>
#include <unordered_map>
#include <utility>
#include <vector>

namespace cosma {

class Mapper {
  public:
    Mapper() = default;
    Mapper(char label, const Strategy& strategy, int rank);

    size_t initial_size(int rank) const;
    size_t initial_size() const { return initial_size(rank_); }
    std::vector<size_t> all_initial_sizes() const { return initial_buffer_size_; }

    // rank -> blocks it owns, in local-buffer order
    const std::vector<Interval2D>& initial_layout(int rank) const { return rank_to_range_[rank]; }
    const std::vector<Interval2D>& initial_layout() const { return rank_to_range_[rank_]; }
    std::vector<std::vector<Interval2D>>& complete_layout() { return rank_to_range_; }
    const std::vector<std::vector<Interval2D>>& complete_layout() const { return rank_to_range_; }

    // (gi, gj) -> (local index, rank)
    std::pair<std::int64_t, int> local_coordinates(int gi, int gj) const;
    // (local index, rank) -> (gi, gj); {-1,-1} if out of range
    std::pair<int, int> global_coordinates(std::int64_t local_index, int rank) const;
    std::pair<int, int> global_coordinates(std::int64_t local_index) const { return global_coordinates(local_index, rank_); }

    char which_matrix() const { return label_; }
    // offsets of this rank's blocks inside its local buffer (n_blocks + 1 entries)
    const std::vector<std::size_t>& local_blocks_offsets() const { return range_offset_[rank_]; }
    const std::vector<std::size_t>& blocks_offsets(int rank) const { return range_offset_[rank]; }
    std::vector<Interval2D> local_blocks() const;
    int owner(const Interval2D& block) const;

    // the grid lines (cumulative split points, starting at 0) of the block grid, and the owner of each grid cell
    const std::vector<int>& row_split() const { return row_split_; }
    const std::vector<int>& col_split() const { return col_split_; }
    std::vector<std::vector<int>> grid_owners() const;

    int m() const { return m_; }
    int n() const { return n_; }
    int P() const { return static_cast<int>(P_); }
    int rank() const { return rank_; }
    char label() const { return label_; }
    const Strategy& strategy() const { return *strategy_; }
    void reorder_rank(int new_rank) { rank_ = new_rank; }

  private:
    char label_ = 'A';
    int m_ = 0, n_ = 0;
    size_t P_ = 0;
    int rank_ = 0;
    const Strategy* strategy_ = nullptr;

    std::vector<std::vector<Interval2D>> rank_to_range_;
    std::unordered_map<Interval2D, std::pair<int, std::size_t>> range_to_rank_;  // block -> (rank, local offset)
    std::vector<size_t> initial_buffer_size_;
    std::vector<std::vector<std::size_t>> range_offset_;
    std::vector<int> row_split_, col_split_;
    std::vector<int> fixed_blocks_;  // per rank: blocks frozen by an enclosing sequential step

    void assign(Interval rows, Interval cols, Interval ranks, size_t step);
};

}  // namespace cosma
