// cosma::CosmaMatrix<Scalar> -- one of A, B, C in COSMA's native layout on this rank, with the reference's public
// interface (src/cosma/matrix.hpp:26-213): a Mapper (who owns which block, local <-> global coordinates) plus the rank's
// local storage, `matrix_size()` elements at `matrix_pointer()`: the rank's Mapper blocks in order, each column-major with
// ld = rows (matrix.cpp:393-429). The storage is page-locked HOST memory from the context's pool, exactly what a caller of
// the reference fills and reads; multiply() moves it through HBM. The reference's buffer/bucket bookkeeping methods
// (shift, seq_bucket, advance_buffer, ...) are internals of its recursion and have no counterpart: the compiled schedule
// owns the device arenas.
#pragma once
#include <cosma/context.hpp>
#include <cosma/interval.hpp>
#include <cosma/mapper.hpp>
#include <cosma/strategy.hpp>
#include <costa/grid2grid/grid_layout.hpp>

#include <iostream>
#include <memory>
#include <tuple>
#include <vector>

namespace cosma {

template <typename Scalar>
class CosmaMatrix {
  public:
    using scalar_t = Scalar;

    CosmaMatrix(cosma_context<Scalar>* ctxt, char label, const Strategy& strategy, int rank, bool dry_run = false);
    CosmaMatrix(cosma_context<Scalar>* ctxt, Mapper&& mapper, int rank, bool dry_run = false);
    CosmaMatrix(std::unique_ptr<cosma_context<Scalar>>& ctxt, char label, const Strategy& strategy, int rank, bool dry_run = false);
    CosmaMatrix(std::unique_ptr<cosma_context<Scalar>>& ctxt, Mapper&& mapper, int rank, bool dry_run = false);
    CosmaMatrix(char label, const Strategy& strategy, int rank, bool dry_run = false);  // global context
    CosmaMatrix(Mapper&& mapper, int rank, bool dry_run = false);
    ~CosmaMatrix();
    CosmaMatrix(const CosmaMatrix&) = delete;
    CosmaMatrix& operator=(const CosmaMatrix&) = delete;

    int m() const { return m_; }
    int n() const { return n_; }
    char label() const { return label_; }
    char which_matrix() const { return label_; }
    int rank() const { return rank_; }

    // (gi, gj) -> (local index, rank); (local index[, rank]) -> (gi, gj)
    std::pair<int, int> local_coordinates(int gi, int gj);
    std::pair<int, int> global_coordinates(int local_index, int rank);
    std::pair<int, int> global_coordinates(int local_index);
    const std::vector<Interval2D>& initial_layout(int rank) const { return mapper_.initial_layout(rank); }
    const std::vector<Interval2D>& initial_layout() const { return mapper_.initial_layout(); }

    scalar_t& operator[](std::size_t index) { return data_[index]; }
    scalar_t operator[](std::size_t index) const { return data_[index]; }
    scalar_t* matrix_pointer() { return data_; }
    const scalar_t* matrix_pointer() const { return data_; }
    size_t matrix_size() const;
    size_t matrix_size(int rank) const;

    // the native layout as a COSTA layout whose blocks view matrix_pointer() (matrix.cpp:393-429)
    costa::grid_layout<scalar_t> get_grid_layout();

    // elements this matrix needs on this rank: [0] the local matrix (host), then the plan's device arena when a plan for
    // the same strategy is registered in the context
    std::vector<size_t> required_memory();
    void allocate();  // turns off dry-run mode
    cosma_context<scalar_t>* get_context() { return ctxt_; }
    const Mapper& mapper() const { return mapper_; }

  protected:
    cosma_context<scalar_t>* ctxt_;
    Mapper mapper_;
    int rank_;
    char label_;
    int m_, n_;
    size_t P_;
    scalar_t* data_ = nullptr;
};

template <typename Scalar>
std::ostream& operator<<(std::ostream& os, CosmaMatrix<Scalar>& mat) {
    for (size_t local = 0; local < mat.matrix_size(); ++local) {
        int row, col;
        std::tie(row, col) = mat.global_coordinates(static_cast<int>(local));
        os << row << " " << col << " " << mat[local] << std::endl;
    }
    return os;
}

}  // namespace cosma
