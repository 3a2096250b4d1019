// costa::transformer<T>: collect (from, to[, op, alpha, beta]) pairs, then move them all in one exchange
// (reference libs/COSTA/src/costa/grid2grid/transformer.hpp:8-63).
#pragma once
#include <costa/grid2grid/transform.hpp>

#include <cassert>
#include <vector>

namespace costa {
template <typename T>
struct transformer {
    std::vector<layout_ref<T>> from;
    std::vector<layout_ref<T>> to;
    std::vector<T> alpha;
    std::vector<T> beta;
    std::vector<char> transpose;
    MPI_Comm comm = MPI_COMM_NULL;
    int P = 0;
    int rank = 0;

    transformer() = default;
    explicit transformer(MPI_Comm c) : comm(c) {
        MPI_Comm_size(comm, &P);
        MPI_Comm_rank(comm, &rank);
    }

    void schedule(grid_layout<T>& from_layout, grid_layout<T>& to_layout) {
        from.push_back(from_layout);
        to.push_back(to_layout);
    }
    void schedule(grid_layout<T>& from_layout, grid_layout<T>& to_layout, const char trans, const T a, const T b) {
        alpha.push_back(a);
        beta.push_back(b);
        transpose.push_back(trans);
        schedule(from_layout, to_layout);
    }
    void transform() {
        assert(alpha.size() == beta.size() && alpha.size() == transpose.size());
        if (!alpha.empty())
            costa::transform<T>(from, to, &transpose[0], &alpha[0], &beta[0], comm);
        else
            costa::transform<T>(from, to, comm);
        clear();
    }
    void clear() {
        from.clear(); to.clear(); alpha.clear(); beta.clear(); transpose.clear();
    }
};
}  // namespace costa
