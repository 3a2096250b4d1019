#include <cosma/interval.hpp>

#include <iostream>
#include <stdexcept>

namespace cosma {

Interval::Interval(int start, int end) : start_(start), end_(end) {
    if (start < 0 || end < 0)
        throw std::runtime_error("ERROR: in class interval (COSMA): start, end > 0 must be satisfied.");
    if (start > end) throw std::runtime_error("ERROR: in class interval (COSMA): start<=end must be satisfied.");
}

std::vector<Interval> Interval::divide_by(int divisor) const {
    if (length() < static_cast<std::size_t>(divisor)) return {*this};
    std::vector<Interval> parts;
    parts.reserve(divisor);
    for (int i = 0; i < divisor; ++i) parts.push_back(subinterval(divisor, i));
    return parts;
}

// NB (reference semantics, interval.cpp:48-76): these use the FLOOR piece size len/divisor, i.e. they describe
// the regular ring arithmetic on rank intervals (always evenly divisible there).
int Interval::subinterval_index(int divisor, int elem) const {
    const int piece = static_cast<int>(length()) / divisor;
    return (elem - start_) / piece;
}
int Interval::subinterval_offset(int divisor, int elem) const {
    const int piece = static_cast<int>(length()) / divisor;
    return (elem - start_) % piece;
}
std::pair<int, int> Interval::locate_in_subinterval(int divisor, int elem) const {
    return {subinterval_index(divisor, elem), subinterval_offset(divisor, elem)};
}
int Interval::locate_in_interval(int divisor, int subint_index, int subint_offset) const {
    const int piece = static_cast<int>(length()) / divisor;
    return subint_index * piece + subint_offset;
}
Interval Interval::subinterval_containing(int divisor, int elem) const {
    return subinterval(divisor, subinterval_index(divisor, elem));
}

Interval Interval::subinterval(int divisor, int box_index) const {
    const std::int64_t len = static_cast<std::int64_t>(length());
    if (len < divisor) return *this;
    const int lo = static_cast<int>(len * box_index / divisor);
    const int hi = static_cast<int>(len * (box_index + 1) / divisor - 1);
    return Interval(start_ + lo, start_ + hi);
}

int Interval::largest_subinterval_length(int divisor) const {
    const int len = static_cast<int>(length());
    return len / divisor + (len % divisor == 0 ? 0 : 1);
}
int Interval::smallest_subinterval_length(int divisor) const { return static_cast<int>(length()) / divisor; }

std::ostream& operator<<(std::ostream& os, const Interval& inter) {
    return os << '[' << inter.start_ << ", " << inter.end_ << ']';
}

std::size_t Interval2D::split_by(int divisor, int index) const {
    if (index >= divisor) {
        std::cout << "Error in Interval2D.split_by: trying to access " << index << "-subinterval, out of " << divisor
                  << " total subintervals\n";
        return static_cast<std::size_t>(-1);
    }
    if (cols.length() < static_cast<std::size_t>(divisor)) {
        std::cout << "Error in Interval2D.split_by: trying to divide the subinterval of length " << cols.length()
                  << " into " << divisor << " many subintervals\n";
        return static_cast<std::size_t>(-1);
    }
    return rows.length() * cols.subinterval(divisor, index).length();
}

bool Interval2D::before(const Interval2D& other) const {
    return (rows.before(other.rows) && other.cols.contains(cols)) || (cols.before(other.cols) && other.rows.contains(rows));
}

std::int64_t Interval2D::local_index(int row, int col) const {
    if (!contains(row, col)) return -1;
    return static_cast<std::int64_t>(col - cols.first()) * static_cast<std::int64_t>(rows.length()) + (row - rows.first());
}

std::pair<int, int> Interval2D::global_index(std::int64_t local_index) const {
    const std::int64_t nrows = static_cast<std::int64_t>(rows.length());
    return {rows.first() + static_cast<int>(local_index % nrows), cols.first() + static_cast<int>(local_index / nrows)};
}

std::ostream& operator<<(std::ostream& os, const Interval2D& inter) {
    return os << "rows " << inter.rows << "; columns: " << inter.cols;
}

}  // namespace cosma
