// costa::transformer<T>: collect (from, to[, op, alpha, beta]) pairs, then move them all in ONE exchange
// (the batching interface of the reference, libs/COSTA/src/costa/grid2grid/transformer.hpp:8-63: schedule() / transform()).
// A pair scheduled without scalars is the plain copy to = from, i.e. op 'N', alpha 1, beta 0.
#pragma once
#include <costa/grid2grid/transform.hpp>

#include <vector>

namespace costa {
template <typename T>
struct transformer {
    MPI_Comm comm = MPI_COMM_NULL;
    int P = 0;     // ranks of comm
    int rank = 0;  // the caller's rank in comm

    transformer() = default;
    explicit transformer(MPI_Comm communicator) : comm(communicator) {
        MPI_Comm_rank(comm, &rank);
        MPI_Comm_size(comm, &P);
    }

    void schedule(grid_layout<T>& source, grid_layout<T>& target, const char op, const T alpha, const T beta) {
        sources_.push_back(source);
        targets_.push_back(target);
        ops_.push_back(op);
        alphas_.push_back(alpha);
        betas_.push_back(beta);
    }
    void schedule(grid_layout<T>& source, grid_layout<T>& target) { schedule(source, target, 'N', T{1}, T{0}); }

    std::size_t pending() const { return sources_.size(); }

    // runs everything scheduled since the last call, then forgets it
    void transform() {
        if (!sources_.empty()) costa::transform<T>(sources_, targets_, ops_.data(), alphas_.data(), betas_.data(), comm);
        clear();
    }
    void clear() {
        sources_.clear();
        targets_.clear();
        ops_.clear();
        alphas_.clear();
        betas_.clear();
    }

  private:
    std::vector<layout_ref<T>> sources_, targets_;
    std::vector<char> ops_;
    std::vector<T> alphas_, betas_;
};
}  // namespace costa
