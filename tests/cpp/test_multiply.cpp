// C++ API test of cosma::multiply: the 40 distributed cases of the reference's tests/multiply.cpp:142-321 and the mixed
// sequential/parallel case of tests/scalar_matmul.cpp:7-39, each in all four element types, on however many ranks the job has
// (cases needing more ranks than the job has are skipped and counted).
#include "cosma_test_utils.hpp"

#include <cosma/b200_runtime.hpp>

#include <map>

#include <initializer_list>
#include <string>

using cosma::Strategy;

struct multiply_state {
    int m, n, k, P;
    std::vector<int> divs;
    std::string dims, step_types;
    multiply_state(int mm, int nn, int kk, int PP, std::vector<int> d = {}, std::string dim = "", std::string steps = "")
        : m(mm), n(nn), k(kk), P(PP), divs(std::move(d)), dims(std::move(dim)), step_types(std::move(steps)) {}
};

static const std::vector<multiply_state>& cases() {
    static const std::vector<multiply_state> c = {
        {4, 4, 4, 1}, {3, 4, 5, 1},
        {8, 4, 2, 4, {2, 2, 2}, "mmn", "psp"}, {8, 4, 2, 4},
        {4, 4, 4, 2, {2}, "m", "p"}, {4, 4, 4, 2},
        {4, 4, 4, 4, {2, 2, 2}, "mnn", "spp"},
        {30, 35, 40, 4},
        {8, 8, 2, 2, {2, 2, 2}, "mmn", "ssp"}, {8, 8, 2, 2},
        {16, 4, 4, 4, {2, 2}, "mm", "pp"}, {16, 4, 4, 4},
        {20, 20, 20, 3, {2, 3}, "km", "sp"}, {20, 20, 20, 3},
        {16, 16, 16, 16, {2, 2, 2, 2}, "mnkm", "pppp"}, {16, 16, 16, 16},
        {20, 30, 25, 4, {2, 2, 2, 2}, "mnkm", "sspp"}, {20, 30, 25, 4},
        {100, 100, 100, 10, {2, 2, 2, 5}, "mnkm", "spsp"}, {100, 100, 100, 10},
        {4, 4, 5, 4, {2, 2, 2, 2}, "mnkm", "spsp"}, {4, 4, 5, 4},
        {10, 10, 10, 12, {2, 2, 3}, "mnk", "ppp"},
        {100, 100, 100, 12, {2, 2, 3}, "mnk", "ppp"}, {100, 100, 100, 12},
        {100, 100, 100, 4},
        {100, 100, 100, 7, {7}, "m", "p"}, {100, 100, 100, 7},
        {100, 100, 100, 8, {2, 2, 2, 2, 2, 2}, "mnkmnk", "spspsp"}, {100, 100, 100, 8},
        {100, 100, 100, 4, {2, 2}, "mk", "pp"},
        {100, 100, 100, 8, {2, 2}, "mk", "pp"},
        {100, 100, 100, 8, {2, 2, 2, 2, 2, 2}, "mknnmk", "sssppp"},
        {100, 100, 100, 8, {2, 2, 2, 2, 2, 2}, "kmnkmn", "spspsp"},
        {200, 200, 200, 8, {3, 3, 3, 2, 2, 2}, "kmnknm", "sssppp"}, {200, 200, 200, 8},
        {200, 200, 200, 8, {3, 2, 3, 2, 3, 2}, "mnkmnk", "spspsp"},
        {512, 32, 736, 8, {2, 2, 2}, "kmk", "ppp"},
        // tests/scalar_matmul.cpp: one strategy, every type
        {100, 100, 100, 8, {2, 2, 2, 2, 2, 2}, "mnkmnk", "spspsp"},
        // sequential-only strategies run on a single rank too
        {96, 80, 64, 1, {2, 2, 2}, "mnk", "sss"}, {300, 260, 200, 1, {3}, "k", "s"},
    };
    return c;
}

template <typename T>
bool run_typed(const multiply_state& st, MPI_Comm comm, int& tag) {
    auto ctx = cosma::make_context<T>();
    std::vector<int> divs = st.divs;
    std::string dims = st.dims, types = st.step_types;
    Strategy strategy = divs.empty() && st.P > 1 ? Strategy(st.m, st.n, st.k, st.P) : Strategy(st.m, st.n, st.k, st.P, divs, dims, types);
    int rank = 0;
    MPI_Comm_rank(comm, &rank);
    if (rank >= static_cast<int>(strategy.P)) { ++tag; return true; }
    // the automatic strategy may use fewer ranks than asked for: the communicator must then be cut again by the caller
    return testutil::test_cosma<T>(strategy, ctx, comm, 1e-8, tag++);
}

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int rank = 0, world = 1;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &world);
    int tag = 0, index = 0;
    std::map<int, MPI_Comm> comm_of;  // one sub-communicator (and so one NCCL communicator) per rank count
    for (const auto& st : cases()) {
        ++index;
        if (st.P > world) { ++check::skipped(); continue; }
        // automatic strategies can idle ranks; size the communicator to the ranks the strategy really uses
        int P_used = st.P;
        if (st.divs.empty() && st.P > 1) P_used = static_cast<int>(Strategy(st.m, st.n, st.k, st.P).P);
        if (!comm_of.count(P_used)) comm_of[P_used] = testutil::subcommunicator(P_used);
        MPI_Comm comm = comm_of[P_used];
        if (rank < P_used) {
            multiply_state eff = st;
            if (st.divs.empty() && st.P > 1) {  // hand the explicit form of the automatic strategy to every type
                Strategy autos(st.m, st.n, st.k, st.P);
                eff.P = static_cast<int>(autos.P);
                eff.divs = autos.divisors; eff.dims = autos.split_dimension; eff.step_types = autos.step_type;
            }
            const bool d = run_typed<double>(eff, comm, tag), s = run_typed<float>(eff, comm, tag);
            const bool z = run_typed<std::complex<double>>(eff, comm, tag), c = run_typed<std::complex<float>>(eff, comm, tag);
            if (rank == 0)
                std::printf("case %2d  (%d x %d x %d, P = %d, steps '%s' '%s')  d:%d s:%d z:%d c:%d\n", index, st.m, st.n, st.k, st.P, eff.dims.c_str(),
                            eff.step_types.c_str(), d, s, z, c);
            CHECK_TRUE(d); CHECK_TRUE(s); CHECK_TRUE(z); CHECK_TRUE(c);
        } else {
            tag += 4;  // keep the message tags of the ranks that sat this case out in step (the reference: tests/multiply.cpp:118-121)
        }
        MPI_Barrier(MPI_COMM_WORLD);
    }
    // idle ranks: every rank of the world calls multiply() with the WORLD communicator and an automatic strategy that uses fewer
    // ranks than the world has (small matrices; reference miniapp/cosma_miniapp.cpp:35-81, multiply.cpp:258-260). Ranks >= P own
    // nothing, return at once and take no part in any collective.
    {
        Strategy small(260, 240, 220, world);
        auto ctx = cosma::make_context<double>();
        const bool ok = testutil::test_cosma<double>(small, ctx, MPI_COMM_WORLD, 1e-8, 1000);
        if (rank == 0) std::printf("idle ranks: strategy uses %d of %d ranks: %d\n", static_cast<int>(small.P), world, ok);
        CHECK_TRUE(ok);
        MPI_Barrier(MPI_COMM_WORLD);
    }
    cosma::b200::release_all_comms();
    const int rc = check::finish("test_multiply");
    MPI_Finalize();
    return rc;
}
