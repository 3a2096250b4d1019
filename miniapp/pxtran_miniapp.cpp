// pxtran_miniapp for cosma_b200: C (m x n, blocks block_c) = beta * C + alpha * op(A) (A: n x m, blocks block_a) on one BLACS grid through
// costa::pxtran_op<T>, set up the way a ScaLAPACK application does. Options follow the reference's miniapp
// (libs/COSTA/miniapps/pxtran_miniapp.cpp:19-53: -m -n, --block_a/--block_c "r,c", -p/--p_grid "r,c" (default: the most square grid
// of all ranks), --alpha, --beta (integers, as there), -r/--n_rep, -t/--type float|double|zfloat|zdouble, --test) plus --op T|C
// (p?tran / p?tranu vs p?tranc; the reference's miniapp times p?tran only). The reference's --test compares with a vendor
// ScaLAPACK; here every rank checks its part of C against the definition on analytically generated integer-valued matrices, exactly.
#include "block_cyclic_matrix.hpp"

#include <costa/pxtran_op/costa_pxtran_op.hpp>

#include <chrono>
#include <iostream>

using namespace miniapp;

struct params {
    int m = 1000, n = 1000, n_rep = 2, alpha = 1, beta = 0;
    int ba[2] = {128, 128}, bc[2] = {128, 128}, grid[2] = {0, 0};
    std::string type = "double";
    char op = 'T';
    bool test = false;
};

template <typename T>
static bool run(const params& p, int ctxt, std::vector<double>& times) {
    cosma::memory_pool<T> pool;
    block_cyclic_matrix<T> A(pool, ctxt, p.n, p.m, p.ba[0], p.ba[1]), C(pool, ctxt, p.m, p.n, p.bc[0], p.bc[1]);
    const T alpha = static_cast<T>(p.alpha), beta = static_cast<T>(p.beta);
    const char op = sizeof(T) == sizeof(typename real_of<T>::type) ? 'T' : p.op;  // real types: p?tran
    A.fill([](long long i, long long j) { return element<T>(0, i, j); });
    for (int r = 0; r < p.n_rep; ++r) {
        C.fill([](long long i, long long j) { return element<T>(1, i, j); });
        MPI_Barrier(MPI_COMM_WORLD);
        const auto t0 = std::chrono::steady_clock::now();
        costa::pxtran_op<T>(p.m, p.n, alpha, A.data(), 1, 1, A.desc, beta, C.data(), 1, 1, C.desc, op);
        MPI_Barrier(MPI_COMM_WORLD);
        times.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    if (!p.test) return true;
    const bool conj = op == 'C';
    return C.mismatches([&](long long i, long long j) { return beta * element<T>(1, i, j) + alpha * conj_if(element<T>(0, j, i), conj); }) == 0;
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
        else if (a == "-p" || a == "--p_grid") pair_of(next(), p.grid);
        else if (a == "--alpha") p.alpha = std::atoi(next().c_str());
        else if (a == "--beta") p.beta = std::atoi(next().c_str());
        else if (a == "--op") p.op = static_cast<char>(std::toupper(next()[0]));
        else if (a == "-r" || a == "--n_rep") p.n_rep = std::atoi(next().c_str());
        else if (a == "-t" || a == "--type") p.type = next();
        else if (a == "--test") p.test = true;
        else if (a == "--algorithm") next();  // accepted for command-line compatibility: there is only one algorithm here
        else if (a == "-h" || a == "--help") {
            std::cout << "usage: pxtran_miniapp -m M -n N [--block_a r,c] [--block_c r,c] [-p r,c] [--alpha a] [--beta b] [--op T|C] [-r reps] "
                         "[-t float|double|zfloat|zdouble] [--test]" << std::endl;
            return 0;
        }
    }
    std::transform(p.type.begin(), p.type.end(), p.type.begin(), [](unsigned char c) { return std::tolower(c); });
    if (p.op != 'C') p.op = 'T';
    if (p.test) p.n_rep = 1;
    MPI_Init(&argc, &argv);
    int rank = 0, P = 1;
    cosma::blacs::Cblacs_pinfo(&rank, &P);
    if (p.grid[0] * p.grid[1] != P) {
        if (p.grid[0] > 0 && rank == 0) std::cout << "pxtran_miniapp: the grid must use all " << P << " ranks; using the most square one" << std::endl;
        square_grid(P, p.grid);
    }
    char order = 'R';
    int ctxt = 0;
    cosma::blacs::Cblacs_get(0, 0, &ctxt);
    cosma::blacs::Cblacs_gridinit(&ctxt, &order, p.grid[0], p.grid[1]);
    std::vector<double> times;
    bool ok = true;
    try {
        if (p.type == "double") ok = run<double>(p, ctxt, times);
        else if (p.type == "float") ok = run<float>(p, ctxt, times);
        else if (p.type == "zdouble") ok = run<std::complex<double>>(p, ctxt, times);
        else if (p.type == "zfloat") ok = run<std::complex<float>>(p, ctxt, times);
        else throw std::runtime_error("--type must be one of float, double, zfloat, zdouble");
    } catch (const std::exception& e) {
        std::cerr << "pxtran_miniapp: " << e.what() << std::endl;
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    int bad = ok ? 0 : 1, bad_all = 0;
    MPI_Allreduce(&bad, &bad_all, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    std::sort(times.begin(), times.end());
    if (rank == 0) {
        const double bytes = static_cast<double>(p.m) * p.n * (p.type == "double" ? 8 : p.type == "float" ? 4 : p.type == "zdouble" ? 16 : 8);
        std::cout << "P" << (p.type == "double" ? "D" : p.type == "float" ? "S" : p.type == "zdouble" ? "Z" : "C") << "TRAN" << (p.type[0] == 'z' ? (p.op == 'C' ? "C" : "U") : "")
                  << ": C " << p.m << " x " << p.n << " (blocks " << p.bc[0] << " x " << p.bc[1] << ") = " << p.beta << " * C + " << p.alpha << " * op(A), A blocks "
                  << p.ba[0] << " x " << p.ba[1] << ", grid " << p.grid[0] << " x " << p.grid[1] << std::endl;
        std::cout << "COSTA TIMES [ms] = ";
        for (double t : times) std::cout << t << " ";
        std::cout << std::endl;
        std::cout << "COSTA BEST [GB/s, matrix bytes / time] = " << bytes / (times.front() * 1e-3) * 1e-9 << std::endl;
        if (p.test) std::cout << "Result is" << (bad_all == 0 ? "" : " NOT") << " CORRECT!" << std::endl;
    }
    cosma::pxgemm_release_grids();
    cosma::blacs::Cblacs_gridexit(ctxt);
    cosma::b200::release_all_comms();
    MPI_Finalize();
    return p.test && bad_all != 0 ? 1 : 0;
}
