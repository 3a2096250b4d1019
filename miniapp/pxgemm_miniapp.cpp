// pxgemm_miniapp for cosma_b200: p?gemm on 2D block-cyclic matrices through cosma::pxgemm<T>, set up the way a ScaLAPACK
// application does (BLACS grid, descinit_, numroc_, local arrays in host memory). Options follow the reference's miniapp
// (miniapp/pxgemm_miniapp.cpp:12-63: -m -n -k, --block_a/b/c "r,c", -p/--p_grid "r,c" (default: the most square grid of all
// ranks), --transpose "NN"|"TN"|..., --alpha, --beta, -r/--n_rep, -t/--type float|double|zfloat|zdouble); inputs are
// U[0,1) with seed 1234 + rank, C pre-filled with NaN when beta == 0. Prints "COSMA TIMES [ms] = ..." sorted.
#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <cosma/memory_pool.hpp>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <vector>

extern "C" {
void descinit_(int* desc, const int* m, const int* n, const int* mb, const int* nb, const int* irsrc, const int* icsrc, const int* ictxt,
               const int* lld, int* info);
int numroc_(const int* n, const int* nb, const int* iproc, const int* isrcproc, const int* nprocs);
}

template <typename T> struct real_of { using type = T; };
template <typename T> struct real_of<std::complex<T>> { using type = T; };

template <typename T> static T draw(std::mt19937_64& g) { return static_cast<T>(std::uniform_real_distribution<double>(0.0, 1.0)(g)); }
template <> std::complex<double> draw<std::complex<double>>(std::mt19937_64& g) { std::uniform_real_distribution<double> d(0.0, 1.0); const double r = d(g); return {r, d(g)}; }
template <> std::complex<float> draw<std::complex<float>>(std::mt19937_64& g) { std::uniform_real_distribution<float> d(0.0f, 1.0f); const float r = d(g); return {r, d(g)}; }

struct params {
    int m = 1000, n = 1000, k = 1000, n_rep = 2;
    int ba[2] = {128, 128}, bb[2] = {128, 128}, bc[2] = {128, 128}, grid[2] = {0, 0};
    std::string trans = "NN", type = "double";
    double alpha = 1, beta = 0;
};

static void pair_of(const std::string& s, int out[2]) {
    const auto c = s.find(',');
    out[0] = std::atoi(s.substr(0, c).c_str());
    out[1] = c == std::string::npos ? out[0] : std::atoi(s.substr(c + 1).c_str());
}

template <typename T>
static std::vector<double> run(const params& p, int ctxt) {
    using R = typename real_of<T>::type;
    int nprow, npcol, myrow, mycol, rank = 0;
    cosma::blacs::Cblacs_gridinfo(ctxt, &nprow, &npcol, &myrow, &mycol);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const char ta = p.trans[0], tb = p.trans[1];
    const int dims[3][2] = {{ta == 'N' ? p.m : p.k, ta == 'N' ? p.k : p.m}, {tb == 'N' ? p.k : p.n, tb == 'N' ? p.n : p.k}, {p.m, p.n}};
    const int* blk[3] = {p.ba, p.bb, p.bc};
    int desc[3][9];
    std::vector<T*> local(3, nullptr);
    std::vector<size_t> elems(3, 0);
    const int zero = 0;
    std::mt19937_64 gen(1234 + rank);
    cosma::memory_pool<T> pool;  // page-locked local arrays, as the reference pins its buffers
    for (int x = 0; x < 3; ++x) {
        const int lr = myrow >= 0 ? numroc_(&dims[x][0], &blk[x][0], &myrow, &zero, &nprow) : 0;
        const int lc = myrow >= 0 ? numroc_(&dims[x][1], &blk[x][1], &mycol, &zero, &npcol) : 0;
        const int lld = std::max(1, lr);
        int info = 0;
        descinit_(desc[x], &dims[x][0], &dims[x][1], &blk[x][0], &blk[x][1], &zero, &zero, &ctxt, &lld, &info);
        elems[x] = static_cast<size_t>(lld) * std::max(1, lc);
        local[x] = pool.allocate(elems[x]);
        for (size_t i = 0; i < elems[x]; ++i)
            local[x][i] = (x == 2 && p.beta == 0) ? T(std::numeric_limits<R>::quiet_NaN()) : draw<T>(gen);
    }
    const int one = 1;
    std::vector<double> times;
    for (int r = 0; r < p.n_rep; ++r) {
        MPI_Barrier(MPI_COMM_WORLD);
        const auto t0 = std::chrono::steady_clock::now();
        cosma::pxgemm<T>(ta, tb, p.m, p.n, p.k, static_cast<T>(p.alpha), local[0], one, one, desc[0], local[1], one, one, desc[1], static_cast<T>(p.beta),
                         local[2], one, one, desc[2]);
        MPI_Barrier(MPI_COMM_WORLD);
        times.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    // the result must be finite everywhere (NaN in C with beta == 0 must not have been read)
    bool finite = true;
    for (size_t i = 0; i < elems[2] && myrow >= 0; ++i) finite = finite && std::isfinite(std::abs(local[2][i]));
    if (!finite) throw std::runtime_error("pxgemm_miniapp: non-finite values in C");
    for (int x = 0; x < 3; ++x) pool.deallocate(local[x]);
    return times;
}

int main(int argc, char** argv) {
    params p;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
        if (a == "-m" || a == "--m_dim") p.m = std::atoi(next().c_str());
        else if (a == "-n" || a == "--n_dim") p.n = std::atoi(next().c_str());
        else if (a == "-k" || a == "--k_dim") p.k = std::atoi(next().c_str());
        else if (a == "--block_a") pair_of(next(), p.ba);
        else if (a == "--block_b") pair_of(next(), p.bb);
        else if (a == "--block_c") pair_of(next(), p.bc);
        else if (a == "-p" || a == "--p_grid") pair_of(next(), p.grid);
        else if (a == "--transpose") p.trans = next();
        else if (a == "--trans_a") p.trans[0] = next()[0];
        else if (a == "--trans_b") p.trans[1] = next()[0];
        else if (a == "--alpha") p.alpha = std::atof(next().c_str());
        else if (a == "--beta") p.beta = std::atof(next().c_str());
        else if (a == "-r" || a == "--n_rep") p.n_rep = std::atoi(next().c_str());
        else if (a == "-t" || a == "--type") p.type = next();
        else if (a == "-h" || a == "--help") {
            std::cout << "usage: pxgemm_miniapp -m M -n N -k K [--block_a r,c] [--block_b r,c] [--block_c r,c] [-p r,c] [--transpose NN] [--alpha a] "
                         "[--beta b] [-r reps] [-t float|double|zfloat|zdouble]" << std::endl;
            return 0;
        }
    }
    for (auto& c : p.trans) c = static_cast<char>(std::toupper(c));
    std::transform(p.type.begin(), p.type.end(), p.type.begin(), [](unsigned char c) { return std::tolower(c); });
    MPI_Init(&argc, &argv);
    int rank = 0, P = 1;
    cosma::blacs::Cblacs_pinfo(&rank, &P);
    if (p.grid[0] <= 0 || p.grid[1] <= 0) {
        int r = 1;
        for (int d = 1; d * d <= P; ++d)
            if (P % d == 0) r = d;
        p.grid[0] = r;
        p.grid[1] = P / r;
    }
    int ctxt = 0;
    char order = 'R';
    cosma::blacs::Cblacs_get(0, 0, &ctxt);
    cosma::blacs::Cblacs_gridinit(&ctxt, &order, p.grid[0], p.grid[1]);
    std::vector<double> times;
    try {
        if (p.type == "double") times = run<double>(p, ctxt);
        else if (p.type == "float") times = run<float>(p, ctxt);
        else if (p.type == "zdouble") times = run<std::complex<double>>(p, ctxt);
        else if (p.type == "zfloat") times = run<std::complex<float>>(p, ctxt);
        else throw std::runtime_error("--type must be one of float, double, zfloat, zdouble");
    } catch (const std::exception& e) {
        std::cerr << "pxgemm_miniapp: " << e.what() << std::endl;
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    std::sort(times.begin(), times.end());
    if (rank == 0) {
        std::cout << "grid " << p.grid[0] << " x " << p.grid[1] << ", " << p.trans << ", m n k = " << p.m << " " << p.n << " " << p.k << std::endl;
        std::cout << "COSMA TIMES [ms] = ";
        for (double t : times) std::cout << t << " ";
        std::cout << std::endl;
        const double flops = (p.type[0] == 'z' ? 8.0 : 2.0) * p.m * p.n * p.k;
        std::cout << "COSMA BEST [TFLOP/s] = " << flops / (times.front() * 1e-3) * 1e-12 << std::endl;
    }
    cosma::pxgemm_release_grids();
    cosma::blacs::Cblacs_gridexit(ctxt);
    cosma::b200::release_all_comms();
    MPI_Finalize();
    return 0;
}
