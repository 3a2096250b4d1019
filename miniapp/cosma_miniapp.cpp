// cosma_miniapp for cosma_b200: C = A * B with dim(A) = m x k, dim(B) = k x n in COSMA's native layout, one rank per GPU.
// Same options, inputs and output line as the reference's miniapp (miniapp/cosma_miniapp.cpp: -m -n -k -s/--steps -r/--n_rep
// -t/--type float|double|zfloat|zdouble; per-rank srand48(rank) and 10*drand48() fills of A then B; alpha = 1, beta = 0;
// "COSMA TIMES [ms] = ..." sorted ascending), so the two can be driven by the same scripts. The reference's --test mode
// (compare with a naive CPU GEMM) lives in tests/cpp/test_multiply.cpp: the product never links a CPU GEMM.
//     python -m cosma_b200.launch -np 8 miniapp/cosma_miniapp -m 32768 -n 32768 -k 32768 -r 5
// The times are wall-clock around multiply() between barriers, host operands in and out (what the reference measures).
#include <cosma/multiply.hpp>

#include <cosma/b200_runtime.hpp>
#include <cosma/environment_variables.hpp>

#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

using namespace cosma;

template <typename T>
static void fill_int(T* ptr, size_t size) {
    for (size_t i = 0; i < size; ++i) ptr[i] = static_cast<T>(10 * drand48());
}

template <typename T>
static bool run(int m, int n, int k, const std::string& steps, double& ms, MPI_Comm comm) {
    int rank = 0, size = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    const long long memory_limit = get_cpu_max_memory<T>();
    Strategy strategy = parse_strategy(m, n, k, static_cast<size_t>(size), steps, memory_limit, false);
    if (rank == 0) std::cout << "Strategy = " << strategy << std::endl;
    CosmaMatrix<T> A('A', strategy, rank), B('B', strategy, rank), C('C', strategy, rank);
    srand48(rank);
    fill_int(A.matrix_pointer(), A.matrix_size());
    fill_int(B.matrix_pointer(), B.matrix_size());
    MPI_Barrier(comm);
    const auto start = std::chrono::steady_clock::now();
    multiply(A, B, C, strategy, comm, T{1}, T{0});
    MPI_Barrier(comm);
    ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count();
    return true;
}

int main(int argc, char** argv) {
    int m = 1000, n = 1000, k = 1000, n_rep = 2;
    std::string steps, type = "double";
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
        if (a == "-m" || a == "--m_dim") m = std::atoi(next().c_str());
        else if (a == "-n" || a == "--n_dim") n = std::atoi(next().c_str());
        else if (a == "-k" || a == "--k_dim") k = std::atoi(next().c_str());
        else if (a == "-r" || a == "--n_rep") n_rep = std::atoi(next().c_str());
        else if (a == "-t" || a == "--type") type = next();
        else if (a == "-s" || a == "--steps") steps = next();
        else if (a == "--test") { std::cout << "cosma_miniapp: correctness is checked by tests/cpp/test_multiply (same cases, naive GEMM oracle)" << std::endl; return 0; }
        else if (a == "-h" || a == "--help") {
            std::cout << "usage: cosma_miniapp -m M -n N -k K [-s pm2,sn2,...] [-r repetitions] [-t float|double|zfloat|zdouble]" << std::endl;
            return 0;
        }
    }
    std::transform(type.begin(), type.end(), type.begin(), [](unsigned char c) { return std::tolower(c); });
    MPI_Init(&argc, &argv);
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    std::vector<double> times;
    bool ok = true;
    int rc = 0;
    try {
        for (int i = 0; i < n_rep; ++i) {
            double ms = 0;
            if (type == "double") ok = run<double>(m, n, k, steps, ms, MPI_COMM_WORLD) && ok;
            else if (type == "float") ok = run<float>(m, n, k, steps, ms, MPI_COMM_WORLD) && ok;
            else if (type == "zdouble") ok = run<std::complex<double>>(m, n, k, steps, ms, MPI_COMM_WORLD) && ok;
            else if (type == "zfloat") ok = run<std::complex<float>>(m, n, k, steps, ms, MPI_COMM_WORLD) && ok;
            else throw std::runtime_error("--type must be one of float, double, zfloat, zdouble");
            times.push_back(ms);
        }
    } catch (const std::exception& e) {
        std::cerr << "cosma_miniapp: " << e.what() << std::endl;
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    std::sort(times.begin(), times.end());
    if (rank == 0) {
        std::cout << "COSMA TIMES [ms] = ";
        for (double t : times) std::cout << t << " ";
        std::cout << std::endl;
        const double flops = (type[0] == 'z' ? 8.0 : 2.0) * m * n * k;
        std::cout << "COSMA BEST [TFLOP/s] = " << flops / (times.front() * 1e-3) * 1e-12 << std::endl;
    }
    if (!ok) rc = 1;
    b200::release_all_comms();
    MPI_Finalize();
    return rc;
}
