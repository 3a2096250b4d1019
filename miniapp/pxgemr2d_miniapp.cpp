// pxgemr2d_miniapp for cosma_b200: C (block-cyclic on grid q, blocks block_c) <- A (block-cyclic on grid p, blocks block_a) through
// costa::pxgemr2d<T>, set up the way a ScaLAPACK application does (two BLACS grids, descinit_, numroc_, host-resident local arrays).
// Options follow the reference's miniapp (libs/COSTA/miniapps/pxgemr2d_miniapp.cpp:19-53: -m -n, --block_a/--block_c "r,c",
// -p/--p_grid_a, -q/--p_grid_c "r,c" (a grid that does not use all ranks becomes 1 x P, as there), -r/--n_rep, -t/--type
// float|double|zfloat|zdouble, --test). The reference's --test compares with a vendor ScaLAPACK; here every rank checks its part of
// C against the definition on analytically generated matrices, bit for bit. Prints "COSTA TIMES [ms] = ..." and the bytes moved.
#include "block_cyclic_matrix.hpp"

#include <costa/pxgemr2d/costa_pxgemr2d.hpp>

#include <chrono>
#include <iostream>

using namespace miniapp;

struct params {
    int m = 1000, n = 1000, n_rep = 2;
    int ba[2] = {128, 128}, bc[2] = {128, 128}, ga[2] = {1, 1}, gc[2] = {1, 1};
    std::string type = "double";
    bool test = false;
};

template <typename T>
static bool run(const params& p, int ctxt_a, int ctxt_c, int ctxt_all, std::vector<double>& times) {
    cosma::memory_pool<T> pool;
    block_cyclic_matrix<T> A(pool, ctxt_a, p.m, p.n, p.ba[0], p.ba[1]), C(pool, ctxt_c, p.m, p.n, p.bc[0], p.bc[1]);
    A.fill([](long long i, long long j) { return element<T>(0, i, j); });
    C.fill([](long long i, long long j) { return element<T>(1, i, j); });
    for (int r = 0; r < p.n_rep; ++r) {
        MPI_Barrier(MPI_COMM_WORLD);
        const auto t0 = std::chrono::steady_clock::now();
        costa::pxgemr2d<T>(p.m, p.n, A.data(), 1, 1, A.desc, C.data(), 1, 1, C.desc, ctxt_all);
        MPI_Barrier(MPI_COMM_WORLD);
        times.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    if (!p.test) return true;
    return C.mismatches([](long long i, long long j) { return element<T>(0, i, j); }) == 0;
}

int main(int argc, char** argv) {
    params p;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
        if (a == "-m" || a == "--m_dim") p.m = std::atoi(next().c_str());
        else if (a == "-n" || a == "--n_dim") p.n = std::atoi(next().c_str());
        else if (a == "--block_a") pair_of(next(), p.ba);
        else if (a == "--block_c") pair_of(next(), p.bc);
        else if (a == "-p" || a == "--p_grid_a") pair_of(next(), p.ga);
        else if (a == "-q" || a == "--p_grid_c") pair_of(next(), p.gc);
        else if (a == "-r" || a == "--n_rep") p.n_rep = std::atoi(next().c_str());
        else if (a == "-t" || a == "--type") p.type = next();
        else if (a == "--test") p.test = true;
        else if (a == "--algorithm") next();  // accepted for command-line compatibility: there is only one algorithm here
        else if (a == "-h" || a == "--help") {
            std::cout << "usage: pxgemr2d_miniapp -m M -n N [--block_a r,c] [--block_c r,c] [-p r,c] [-q r,c] [-r reps] "
                         "[-t float|double|zfloat|zdouble] [--test]" << std::endl;
            return 0;
        }
    }
    std::transform(p.type.begin(), p.type.end(), p.type.begin(), [](unsigned char c) { return std::tolower(c); });
    if (p.test) p.n_rep = 1;
    MPI_Init(&argc, &argv);
    int rank = 0, P = 1;
    cosma::blacs::Cblacs_pinfo(&rank, &P);
    if (std::max(p.ga[0] * p.ga[1], p.gc[0] * p.gc[1]) != P) {
        if (rank == 0) std::cout << "pxgemr2d_miniapp: the larger grid must use all " << P << " ranks; using 1 x " << P << " for both" << std::endl;
        p.ga[0] = p.gc[0] = 1;
        p.ga[1] = p.gc[1] = P;
    }
    char order = 'R';
    int ctxt_a = 0, ctxt_c = 0, ctxt_all = 0;
    cosma::blacs::Cblacs_get(0, 0, &ctxt_a);
    cosma::blacs::Cblacs_gridinit(&ctxt_a, &order, p.ga[0], p.ga[1]);
    cosma::blacs::Cblacs_get(0, 0, &ctxt_c);
    cosma::blacs::Cblacs_gridinit(&ctxt_c, &order, p.gc[0], p.gc[1]);
    cosma::blacs::Cblacs_get(0, 0, &ctxt_all);  // the context that contains every rank of both grids (the last argument of p?gemr2d)
    cosma::blacs::Cblacs_gridinit(&ctxt_all, &order, 1, P);
    std::vector<double> times;
    bool ok = true;
    try {
        if (p.type == "double") ok = run<double>(p, ctxt_a, ctxt_c, ctxt_all, times);
        else if (p.type == "float") ok = run<float>(p, ctxt_a, ctxt_c, ctxt_all, times);
        else if (p.type == "zdouble") ok = run<std::complex<double>>(p, ctxt_a, ctxt_c, ctxt_all, times);
        else if (p.type == "zfloat") ok = run<std::complex<float>>(p, ctxt_a, ctxt_c, ctxt_all, times);
        else throw std::runtime_error("--type must be one of float, double, zfloat, zdouble");
    } catch (const std::exception& e) {
        std::cerr << "pxgemr2d_miniapp: " << e.what() << std::endl;
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    int bad = ok ? 0 : 1, bad_all = 0;
    MPI_Allreduce(&bad, &bad_all, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    std::sort(times.begin(), times.end());
    if (rank == 0) {
        const double bytes = static_cast<double>(p.m) * p.n * (p.type == "double" ? 8 : p.type == "float" ? 4 : p.type == "zdouble" ? 16 : 8);
        std::cout << "P" << (p.type == "double" ? "D" : p.type == "float" ? "S" : p.type == "zdouble" ? "Z" : "C") << "GEMR2D: " << p.m << " x " << p.n
                  << ", blocks " << p.ba[0] << " x " << p.ba[1] << " on grid " << p.ga[0] << " x " << p.ga[1] << " -> blocks " << p.bc[0] << " x " << p.bc[1]
                  << " on grid " << p.gc[0] << " x " << p.gc[1] << std::endl;
        std::cout << "COSTA TIMES [ms] = ";
        for (double t : times) std::cout << t << " ";
        std::cout << std::endl;
        std::cout << "COSTA BEST [GB/s, matrix bytes / time] = " << bytes / (times.front() * 1e-3) * 1e-9 << std::endl;
        if (p.test) std::cout << "Result is" << (bad_all == 0 ? "" : " NOT") << " CORRECT!" << std::endl;
    }
    cosma::pxgemm_release_grids();
    cosma::blacs::Cblacs_gridexit(ctxt_a);
    cosma::blacs::Cblacs_gridexit(ctxt_c);
    cosma::blacs::Cblacs_gridexit(ctxt_all);
    cosma::b200::release_all_comms();
    MPI_Finalize();
    return p.test && bad_all != 0 ? 1 : 0;
}
