// costa::grid_layout<T> -- the typed, public face of a distributed matrix layout, with the reference's names
// (libs/COSTA/src/costa/grid2grid/grid_layout.hpp:9-181, block.hpp:64-137): a grid of blocks with owners plus views on the
// blocks of the calling rank. The views may point at HOST memory (the reference's convention; staged through HBM by the
// library) or at DEVICE memory (no staging). The element-wise helpers (initialize / apply / validate / accumulate /
// local_element) touch the memory from the host and are therefore only meaningful for host-resident blocks.
#pragma once
#include <costa/erased_layout.hpp>

#include <cassert>
#include <cctype>
#include <cmath>
#include <complex>
#include <functional>
#include <iostream>
#include <tuple>
#include <utility>
#include <vector>

namespace costa {

// view on one local block: rows x cols of the global matrix, stored with `stride` between columns ('C') or rows ('R')
template <typename T>
struct block {
    T* data = nullptr;
    int stride = 0;
    interval rows_interval, cols_interval;   // global index ranges
    std::pair<int, int> coordinates{0, 0};   // block coordinates in the grid
    char _ordering = 'C';
    bool transposed = false;

    block() = default;
    block(const assigned_grid2D& g, interval r, interval c, T* ptr, int stride_ = 0) : data(ptr), stride(stride_), rows_interval(r), cols_interval(c) {
        coordinates = {locate(g.grid.rows_split, r.start), locate(g.grid.cols_split, c.start)};
        if (stride == 0) stride = r.length();
    }
    // block (bi, bj) of the grid
    block(const assigned_grid2D& g, int bi, int bj, T* ptr, int stride_ = 0)
        : block(g, g.rows_interval(bi), g.cols_interval(bj), ptr, stride_) {}

    int n_rows() const { return rows_interval.length(); }
    int n_cols() const { return cols_interval.length(); }
    bool non_empty() const { return data != nullptr && n_rows() > 0 && n_cols() > 0; }
    std::size_t total_size() const { return static_cast<std::size_t>(n_rows()) * n_cols(); }
    void set_ordering(char o) { _ordering = static_cast<char>(std::toupper(o)); }
    T& local_element(int li, int lj) {
        return _ordering == 'C' ? data[static_cast<std::size_t>(lj) * stride + li] : data[static_cast<std::size_t>(li) * stride + lj];
    }
    const T& local_element(int li, int lj) const { return const_cast<block*>(this)->local_element(li, lj); }
    std::pair<int, int> local_to_global(int li, int lj) const { return {rows_interval.start + li, cols_interval.start + lj}; }
    std::pair<int, int> global_to_local(int gi, int gj) const {
        if (!rows_interval.contains(gi) || !cols_interval.contains(gj)) return {-1, -1};
        return {gi - rows_interval.start, gj - cols_interval.start};
    }
    void scale_by(T beta) {
        if (beta == T{1}) return;
        for (int lj = 0; lj < n_cols(); ++lj)
            for (int li = 0; li < n_rows(); ++li) local_element(li, lj) = beta == T{0} ? T{0} : beta * local_element(li, lj);
    }
    void fill(T value) {
        for (int lj = 0; lj < n_cols(); ++lj)
            for (int li = 0; li < n_rows(); ++li) local_element(li, lj) = value;
    }

  private:
    static int locate(const std::vector<int>& split, int v) {
        int i = 0;
        while (i + 1 < static_cast<int>(split.size()) && split[i + 1] <= v) ++i;
        return i;
    }
};

template <typename T>
class local_blocks {
  public:
    local_blocks() = default;
    explicit local_blocks(std::vector<block<T>>&& b) : blocks_(std::move(b)) {}
    std::size_t num_blocks() const { return blocks_.size(); }
    block<T>& get_block(std::size_t i) { return blocks_[i]; }
    const block<T>& get_block(std::size_t i) const { return blocks_[i]; }
    std::size_t size() const {
        std::size_t s = 0;
        for (const auto& b : blocks_) s += b.total_size();
        return s;
    }
    std::vector<block<T>>& blocks() { return blocks_; }
    const std::vector<block<T>>& blocks() const { return blocks_; }

  private:
    std::vector<block<T>> blocks_;
};

template <typename T>
class grid_layout {
  public:
    grid_layout() = default;
    grid_layout(assigned_grid2D&& g, local_blocks<T>&& b, char ordering_) : grid(std::move(g)), blocks(std::move(b)) {
        ordering = static_cast<char>(std::toupper(ordering_));
        assert(ordering == 'R' || ordering == 'C');
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) blocks.get_block(i).set_ordering(ordering);
    }

    int num_ranks() const { return grid.num_ranks(); }
    int num_cols() const noexcept { return grid.num_cols(); }
    int num_rows() const noexcept { return grid.num_rows(); }
    int num_blocks_col() const noexcept { return grid.num_blocks_col(); }
    int num_blocks_row() const noexcept { return grid.num_blocks_row(); }

    void reorder_ranks(std::vector<int>& reordering) { grid.reorder_ranks(reordering); }

    // view the same memory as the transposed matrix (reference grid_layout::transpose, grid_layout.hpp:27-30): grid lines and
    // block coordinates swap, and a column-major block read as its transpose is a row-major block with the same stride
    void transpose() {
        grid = grid.transposed();
        ordering = ordering == 'C' ? 'R' : 'C';
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) {
            auto& b = blocks.get_block(i);
            std::swap(b.rows_interval, b.cols_interval);
            std::swap(b.coordinates.first, b.coordinates.second);
            b.transposed = !b.transposed;
            b.set_ordering(ordering);
        }
    }

    // host-side element-wise helpers (host-resident blocks only)
    void scale_by(const T beta) {
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) blocks.get_block(i).scale_by(beta);
    }
    void fill(const T value) {
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) blocks.get_block(i).fill(value);
    }
    template <typename Function>
    void initialize(Function f) {
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) {
            auto& b = blocks.get_block(i);
            for (int li = 0; li < b.n_rows(); ++li)
                for (int lj = 0; lj < b.n_cols(); ++lj) {
                    int gi, gj;
                    std::tie(gi, gj) = b.local_to_global(li, lj);
                    b.local_element(li, lj) = static_cast<T>(f(gi, gj));
                }
        }
    }
    template <typename Function>
    void apply(Function f) {
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) {
            auto& b = blocks.get_block(i);
            for (int li = 0; li < b.n_rows(); ++li)
                for (int lj = 0; lj < b.n_cols(); ++lj) {
                    int gi, gj;
                    std::tie(gi, gj) = b.local_to_global(li, lj);
                    b.local_element(li, lj) = static_cast<T>(f(gi, gj, b.local_element(li, lj)));
                }
        }
    }
    template <typename Function>
    bool validate(Function f, double tolerance = 1e-12) {
        bool ok = true;
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) {
            auto& b = blocks.get_block(i);
            for (int li = 0; li < b.n_rows(); ++li)
                for (int lj = 0; lj < b.n_cols(); ++lj) {
                    int gi, gj;
                    std::tie(gi, gj) = b.local_to_global(li, lj);
                    const T want = static_cast<T>(f(gi, gj));
                    if (!(std::abs(b.local_element(li, lj) - want) <= tolerance)) {
                        if (ok) std::cout << "[ERROR] mat(" << gi << ", " << gj << ") = " << b.local_element(li, lj) << " instead of " << want << std::endl;
                        ok = false;
                    }
                }
        }
        return ok;
    }
    template <typename Function>
    T accumulate(Function f, T initial_value) {
        T result = initial_value;
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) {
            auto& b = blocks.get_block(i);
            for (int li = 0; li < b.n_rows(); ++li)
                for (int lj = 0; lj < b.n_cols(); ++lj) result = f(result, b.local_element(li, lj));
        }
        return result;
    }

    // the type-erased description the planner and the C ABI consume
    erased_layout erased() const {
        erased_layout e;
        e.grid = grid;
        e.ordering = ordering;
        for (std::size_t i = 0; i < blocks.num_blocks(); ++i) {
            const auto& b = blocks.get_block(i);
            e.blocks.push_back(local_block{b.coordinates.first, b.coordinates.second, const_cast<T*>(b.data), b.stride});
        }
        return e;
    }

    assigned_grid2D grid;
    local_blocks<T> blocks;
    char ordering = 'C';
};

template <typename T>
using layout_ref = std::reference_wrapper<grid_layout<T>>;

}  // namespace costa
