// Restatement of the reference data-layout rule (src/cosma/mapper.cpp:106-206): replay the strategy; a step that
// splits this matrix hands each sub-range to a sub-group of ranks (parallel) or to all ranks in turn (sequential);
// a parallel step that does NOT split this matrix ("copy case") leaves the matrix with the first sub-group and then
// deals the COLUMNS of every block that group owns round the `div` ring members, so that an allgather over the ring
// reproduces the block.
#include <cosma/mapper.hpp>

#include <algorithm>
#include <iostream>
#include <set>
#include <stdexcept>

namespace cosma {

Mapper::Mapper(char label, const Strategy& strategy, int rank)
    : label_(label), m_(strategy.n_rows(label)), n_(strategy.n_cols(label)), P_(strategy.P), rank_(rank),
      strategy_(&strategy) {
    rank_to_range_.assign(P_, {});
    fixed_blocks_.assign(P_, 0);
    assign(Interval(0, m_ - 1), Interval(0, n_ - 1), Interval(0, static_cast<int>(P_) - 1), 0);

    initial_buffer_size_.assign(P_, 0);
    range_offset_.assign(P_, {});
    std::set<int> row_ends, col_ends;
    for (size_t r = 0; r < P_; ++r) {
        size_t off = 0;
        for (const auto& block : rank_to_range_[r]) {
            range_offset_[r].push_back(off);
            range_to_rank_.insert({block, {static_cast<int>(r), off}});
            row_ends.insert(block.rows.last());
            col_ends.insert(block.cols.last());
            off += block.size();
        }
        range_offset_[r].push_back(off);
        initial_buffer_size_[r] = off;
        if (rank_to_range_[r].empty()) std::cout << "RANK " << r << " DOES NOT OWN ANYTHING" << std::endl;
    }
    row_split_.push_back(0);
    for (int e : row_ends) row_split_.push_back(e + 1);
    col_split_.push_back(0);
    for (int e : col_ends) col_split_.push_back(e + 1);
}

void Mapper::assign(Interval rows, Interval cols, Interval ranks, size_t step) {
    const Strategy& st = *strategy_;
    if (st.final_step(step) || st.empty()) {
        rank_to_range_[ranks.first()].emplace_back(rows, cols);
        return;
    }
    const int div = st.divisor(step);
    const int div_rows = st.divisor_row(label_, step), div_cols = st.divisor_col(label_, step);
    const bool split_here = div_rows * div_cols > 1;

    if (st.sequential_step(step)) {
        std::vector<int> saved(fixed_blocks_.begin() + ranks.first(), fixed_blocks_.begin() + ranks.last() + 1);
        for (int i = 0; i < div; ++i) {
            assign(rows.subinterval(div_rows, div_rows > 1 ? i : 0), cols.subinterval(div_cols, div_cols > 1 ? i : 0), ranks,
                   step + 1);
            // blocks produced by this sub-problem must not be re-dealt by a copy step of the next one
            for (int r = ranks.first(); r <= ranks.last(); ++r) fixed_blocks_[r] = static_cast<int>(rank_to_range_[r].size());
            if (!split_here) break;  // a dimension this matrix does not have: one pass describes all sub-problems
        }
        std::copy(saved.begin(), saved.end(), fixed_blocks_.begin() + ranks.first());
        return;
    }

    if (split_here) {
        for (int i = 0; i < div; ++i)
            assign(rows.subinterval(div_rows, div_rows > 1 ? i : 0), cols.subinterval(div_cols, div_cols > 1 ? i : 0),
                   ranks.subinterval(div, i), step + 1);
        return;
    }

    // copy case
    const Interval group0 = ranks.subinterval(div, 0);
    assign(rows, cols, group0, step + 1);
    const int group_size = static_cast<int>(group0.length());
    for (int r = group0.first(); r <= group0.last(); ++r) {
        auto& mine = rank_to_range_[r];
        for (size_t b = fixed_blocks_[r]; b < mine.size(); ++b) {
            const Interval2D whole = mine[b];
            for (int part = 1; part < div; ++part) rank_to_range_[part * group_size + r].push_back(whole.submatrix(div, part));
            mine[b] = whole.submatrix(div, 0);
        }
    }
}

size_t Mapper::initial_size(int rank) const { return rank < static_cast<int>(P_) ? initial_buffer_size_[rank] : 0; }

std::pair<std::int64_t, int> Mapper::local_coordinates(int gi, int gj) const {
    // the blocks tile the matrix as a Cartesian grid of the row and column split points
    auto cell = [](const std::vector<int>& split, int x) {
        auto it = std::upper_bound(split.begin(), split.end(), x);  // first split point > x
        const int hi = static_cast<int>(it - split.begin());
        return std::make_pair(split[hi - 1], split[hi] - 1);
    };
    const auto r = cell(row_split_, gi), c = cell(col_split_, gj);
    const Interval2D block(r.first, r.second, c.first, c.second);
    const auto it = range_to_rank_.find(block);
    if (it == range_to_rank_.end()) {
        std::cout << "Error in local_coordinates(" << gi << ", " << gj << ") does not belong to the range " << block << std::endl;
        return {-1, -1};
    }
    return {static_cast<std::int64_t>(it->second.second) + block.local_index(gi, gj), it->second.first};
}

std::pair<int, int> Mapper::global_coordinates(std::int64_t local_index, int rank) const {
    if (rank < 0 || rank >= static_cast<int>(P_) || local_index < 0) return {-1, -1};
    const auto& offs = range_offset_[rank];
    if (local_index >= static_cast<std::int64_t>(offs.back())) return {-1, -1};
    const size_t b = std::upper_bound(offs.begin(), offs.end(), static_cast<std::size_t>(local_index)) - offs.begin() - 1;
    return rank_to_range_[rank][b].global_index(local_index - static_cast<std::int64_t>(offs[b]));
}

std::vector<Interval2D> Mapper::local_blocks() const {
    if (rank_ < static_cast<int>(strategy_->P)) return rank_to_range_[rank_];
    return {};
}

int Mapper::owner(const Interval2D& block) const {
    const auto it = range_to_rank_.find(block);
    if (it == range_to_rank_.end())
        throw std::runtime_error("ERROR in mapper.cpp: the owner cannot be determined, the block not found.");
    return it->second.first;
}

std::vector<std::vector<int>> Mapper::grid_owners() const {
    const int nr = static_cast<int>(row_split_.size()) - 1, nc = static_cast<int>(col_split_.size()) - 1;
    std::vector<std::vector<int>> owners(nr, std::vector<int>(nc));
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j)
            owners[i][j] = owner(Interval2D(row_split_[i], row_split_[i + 1] - 1, col_split_[j], col_split_[j + 1] - 1));
    return owners;
}

}  // namespace cosma
