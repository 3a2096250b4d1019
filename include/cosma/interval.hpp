// cosma::Interval / cosma::Interval2D -- closed integer ranges used for matrix dimensions and rank sets.
// Same interface and arithmetic as the reference (src/cosma/interval.hpp:7-80, interval.cpp:5-205):
// subinterval(div, i) = [len*i/div, len*(i+1)/div - 1] (interleaves larger and smaller pieces).
// Products of lengths are 64-bit here (the reference's `int` sizes overflow at 2^31 elements).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <iosfwd>
#include <utility>
#include <vector>

namespace cosma {

class Interval {
  public:
    int start_ = 0;
    int end_ = 0;

    Interval() = default;
    Interval(int start, int end);  // throws std::runtime_error unless 0 <= start <= end

    int first() const { return start_; }
    int last() const { return end_; }
    std::size_t length() const { return static_cast<std::size_t>(end_ - start_ + 1); }
    bool empty() const { return start_ == end_; }  // (sic) reference semantics, interval.cpp:26
    bool only_one() const { return length() == 1; }

    std::vector<Interval> divide_by(int divisor) const;
    int subinterval_index(int divisor, int elem) const;
    int subinterval_offset(int divisor, int elem) const;
    std::pair<int, int> locate_in_subinterval(int divisor, int elem) const;
    int locate_in_interval(int divisor, int subint_index, int subint_offset) const;
    Interval subinterval_containing(int divisor, int elem) const;
    Interval subinterval(int divisor, int box_index) const;
    int largest_subinterval_length(int divisor) const;
    int smallest_subinterval_length(int divisor) const;

    bool contains(int num) const { return num >= start_ && num <= end_; }
    bool contains(const Interval& other) const { return start_ <= other.start_ && end_ >= other.end_; }
    bool before(const Interval& other) const { return end_ < other.start_; }
    bool operator==(const Interval& other) const { return start_ == other.start_ && end_ == other.end_; }
    friend std::ostream& operator<<(std::ostream& os, const Interval& inter);
};

class Interval2D {
  public:
    Interval rows;
    Interval cols;

    Interval2D() = default;
    Interval2D(Interval row, Interval col) : rows(row), cols(col) {}
    Interval2D(int row_start, int row_end, int col_start, int col_end)
        : rows(row_start, row_end), cols(col_start, col_end) {}

    // size of the index-th of `divisor` column slices
    std::size_t split_by(int divisor, int index) const;
    std::size_t size() const { return rows.length() * cols.length(); }
    bool contains(int row, int col) const { return rows.contains(row) && cols.contains(col); }
    bool contains(const Interval2D& other) const { return rows.contains(other.rows) && cols.contains(other.cols); }
    bool before(const Interval2D& other) const;
    // column-major position inside this block (-1 if outside) and its inverse
    std::int64_t local_index(int row, int col) const;
    std::pair<int, int> global_index(std::int64_t local_index) const;
    Interval2D submatrix(int divisor, int index) const { return Interval2D(rows, cols.subinterval(divisor, index)); }
    bool operator==(const Interval2D& other) const { return rows == other.rows && cols == other.cols; }
    friend std::ostream& operator<<(std::ostream& os, const Interval2D& inter);
};

}  // namespace cosma

namespace std {
template <>
struct hash<cosma::Interval2D> {
    std::size_t operator()(const cosma::Interval2D& b) const noexcept {
        std::uint64_t h = 0xcbf29ce484222325ull;
        for (int v : {b.rows.start_, b.rows.end_, b.cols.start_, b.cols.end_}) {
            h ^= static_cast<std::uint32_t>(v);
            h *= 0x100000001b3ull;
        }
        return static_cast<std::size_t>(h);
    }
};
}  // namespace std
