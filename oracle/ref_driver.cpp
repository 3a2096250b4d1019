// TEST INFRASTRUCTURE ONLY -- multi-rank driver of the UNMODIFIED reference (oracle/_ref/libcosma_ref.so), started on R
// ranks by oracle/minirun.py over the minimpi / miniblacs stand-ins. Inputs and outputs are raw little-endian files in a
// scratch directory so that tests (and only tests / bench.py's reference arm) can compare the reference's per-rank
// results with ours:
//
//   ref_driver op=multiply  dtype=d m= n= k= steps=auto|pm2,pn2,.. alpha=re,im beta=re,im dir=D
//       D/A.bin D/B.bin D/C.bin : dense global matrices, column-major.  Each rank fills its CosmaMatrix through
//       global_coordinates(), runs cosma::multiply (src/cosma/multiply.cpp:222-314) and writes its raw local C buffer
//       (matrix_pointer(), matrix_size()) to D/Cout.<rank>.bin.
//   ref_driver op=pxgemm    dtype= ta= tb= m= n= k= alpha= beta= ia= ja= ib= jb= ic= jc= order=R|C nprow= npcol=
//                           desca=M,N,MB,NB,RSRC,CSRC,LLD descb=.. descc=.. [llda=l0,l1,.. per rank] dir=D
//       D/a.<rank>.bin, b.<rank>.bin, c.<rank>.bin : the rank's local ScaLAPACK arrays (LLD x local columns).
//       Runs cosma::pxgemm<T> (src/cosma/cosma_pxgemm.cpp:16-388) and writes D/cout.<rank>.bin.
//   ref_driver op=pxgemr2d  dtype= m= n= ia= ja= ic= jc= order= nprow= npcol= orderc= desca= descc= dir=D
//       costa::pxgemr2d<T> (libs/COSTA/src/costa/pxgemr2d/costa_pxgemr2d.cpp:14-168); a.<rank>.bin -> cout.<rank>.bin
//   ref_driver op=pxtran    dtype= trans=T|C m= n= alpha= beta= ia= ja= ic= jc= order= nprow= npcol= desca= descc= dir=D
//       costa::pxtran_op<T> (libs/COSTA/src/costa/pxtran_op/costa_pxtran_op.cpp:14-172): sub(C) = beta*sub(C) +
//       alpha*op(sub(A)); a.<rank>.bin, c.<rank>.bin -> cout.<rank>.bin
//
// Prints "REF_TIME_MS <t>" per rank-0 repetition when reps= is given (op=multiply|pxgemm).
#include <cosma/cosma_pxgemm.hpp>
#include <cosma/multiply.hpp>
#include <cosma/strategy.hpp>
#include <costa/pxgemr2d/costa_pxgemr2d.hpp>
#include <costa/pxtran_op/costa_pxtran_op.hpp>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

extern "C" {
void Cblacs_get(int, int, int*);
void Cblacs_gridinit(int*, char*, int, int);
void Cblacs_gridexit(int);
}

namespace {
using Args = std::map<std::string, std::string>;

std::string need(const Args& a, const std::string& k) {
    auto it = a.find(k);
    if (it == a.end()) {
        std::cerr << "ref_driver: missing argument " << k << "=\n";
        std::exit(2);
    }
    return it->second;
}
int geti(const Args& a, const std::string& k) { return std::stoi(need(a, k)); }
int geti(const Args& a, const std::string& k, int dflt) { return a.count(k) ? std::stoi(a.at(k)) : dflt; }
std::vector<double> getv(const Args& a, const std::string& k) {
    std::vector<double> v;
    std::stringstream ss(need(a, k));
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(tok));
    return v;
}
template <typename T> struct scalar_of {
    static T make(const std::vector<double>& v) { return T(v.at(0)); }
};
template <typename R> struct scalar_of<std::complex<R>> {
    static std::complex<R> make(const std::vector<double>& v) { return std::complex<R>(R(v.at(0)), R(v.size() > 1 ? v[1] : 0)); }
};

template <typename T>
std::vector<T> read_file(const std::string& path, size_t min_elems = 0) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) {
        std::cerr << "ref_driver: cannot open " << path << "\n";
        std::exit(3);
    }
    const size_t bytes = size_t(f.tellg());
    f.seekg(0);
    std::vector<T> v(std::max(bytes / sizeof(T), min_elems));
    f.read(reinterpret_cast<char*>(v.data()), bytes);
    return v;
}
template <typename T>
void write_file(const std::string& path, const T* p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(p), n * sizeof(T));
}

cosma::Strategy make_strategy(int m, int n, int k, int P, const std::string& steps) {
    if (steps.empty() || steps == "auto") return cosma::Strategy(m, n, k, P);
    std::vector<int> divs;
    std::string dims, types, tok;
    std::stringstream ss(steps);
    while (std::getline(ss, tok, ',')) {
        types += tok.at(0);
        dims += tok.at(1);
        divs.push_back(std::stoi(tok.substr(2)));
    }
    return cosma::Strategy(m, n, k, P, divs, dims, types);
}

template <typename T>
int run_multiply(const Args& a, int rank, int P) {
    const int m = geti(a, "m"), n = geti(a, "n"), k = geti(a, "k"), reps = geti(a, "reps", 1);
    const std::string dir = need(a, "dir");
    const T alpha = scalar_of<T>::make(getv(a, "alpha")), beta = scalar_of<T>::make(getv(a, "beta"));
    cosma::Strategy strategy = make_strategy(m, n, k, P, need(a, "steps"));
    if (rank == 0) std::cout << "REF_STRATEGY " << strategy.n_steps() << " P=" << strategy.P << std::endl;
    const auto Ag = read_file<T>(dir + "/A.bin"), Bg = read_file<T>(dir + "/B.bin"), Cg = read_file<T>(dir + "/C.bin");
    cosma::CosmaMatrix<T> A('A', strategy, rank), B('B', strategy, rank), C('C', strategy, rank);
    auto fill = [&](cosma::CosmaMatrix<T>& M, const std::vector<T>& G, int rows) {
        if (rank >= (int)strategy.P) return;
        for (size_t l = 0; l < M.matrix_size(); ++l) {
            auto ij = M.global_coordinates((int)l);
            M.matrix_pointer()[l] = G[size_t(ij.second) * rows + ij.first];
        }
    };
    for (int r = 0; r < reps; ++r) {
        fill(A, Ag, m);
        fill(B, Bg, k);
        fill(C, Cg, m);
        MPI_Barrier(MPI_COMM_WORLD);
        auto t0 = std::chrono::steady_clock::now();
        cosma::multiply(A, B, C, strategy, MPI_COMM_WORLD, alpha, beta);
        MPI_Barrier(MPI_COMM_WORLD);
        auto t1 = std::chrono::steady_clock::now();
        if (rank == 0) std::cout << "REF_TIME_MS " << std::chrono::duration<double, std::milli>(t1 - t0).count() << std::endl;
    }
    if (rank < (int)strategy.P) write_file(dir + "/Cout." + std::to_string(rank) + ".bin", C.matrix_pointer(), C.matrix_size());
    return 0;
}

// 9-int descriptor from "M,N,MB,NB,RSRC,CSRC,LLD"
// (LLD may differ per rank: "lld" + last letter of key, e.g. llda=l0,l1,..., overrides it with the rank's own)
std::vector<int> make_desc(const Args& a, const std::string& key, int ctxt, int rank) {
    const auto v = getv(a, key);
    int lld = int(v.at(6));
    const std::string per_rank = std::string("lld") + key.back();
    if (a.count(per_rank)) lld = int(getv(a, per_rank).at(rank));
    return {1, ctxt, int(v.at(0)), int(v.at(1)), int(v.at(2)), int(v.at(3)), int(v.at(4)), int(v.at(5)), lld};
}
int make_grid(const Args& a, const std::string& order_key) {
    int sys = 0, ctxt = 0;
    Cblacs_get(0, 0, &sys);
    ctxt = sys;
    char order = need(a, order_key).at(0);
    Cblacs_gridinit(&ctxt, &order, geti(a, "nprow"), geti(a, "npcol"));
    return ctxt;
}
std::string rank_file(const std::string& dir, const char* name, int rank) { return dir + "/" + name + "." + std::to_string(rank) + ".bin"; }

template <typename T>
int run_pxgemm(const Args& a, int rank) {
    const std::string dir = need(a, "dir");
    const int reps = geti(a, "reps", 1);
    const int ctxt = make_grid(a, "order");
    auto da = make_desc(a, "desca", ctxt, rank), db = make_desc(a, "descb", ctxt, rank), dc = make_desc(a, "descc", ctxt, rank);
    auto la = read_file<T>(rank_file(dir, "a", rank), 1), lb = read_file<T>(rank_file(dir, "b", rank), 1);
    const auto lc0 = read_file<T>(rank_file(dir, "c", rank), 1);
    auto lc = lc0;
    const T alpha = scalar_of<T>::make(getv(a, "alpha")), beta = scalar_of<T>::make(getv(a, "beta"));
    for (int r = 0; r < reps; ++r) {
        lc = lc0;
        MPI_Barrier(MPI_COMM_WORLD);
        auto t0 = std::chrono::steady_clock::now();
        cosma::pxgemm<T>(need(a, "ta").at(0), need(a, "tb").at(0), geti(a, "m"), geti(a, "n"), geti(a, "k"), alpha, la.data(), geti(a, "ia"),
                         geti(a, "ja"), da.data(), lb.data(), geti(a, "ib"), geti(a, "jb"), db.data(), beta, lc.data(), geti(a, "ic"),
                         geti(a, "jc"), dc.data());
        MPI_Barrier(MPI_COMM_WORLD);
        auto t1 = std::chrono::steady_clock::now();
        if (rank == 0) std::cout << "REF_TIME_MS " << std::chrono::duration<double, std::milli>(t1 - t0).count() << std::endl;
    }
    write_file(rank_file(dir, "cout", rank), lc.data(), lc.size());
    return 0;
}

template <typename T>
int run_pxgemr2d(const Args& a, int rank) {
    const std::string dir = need(a, "dir");
    const int ctxt_a = make_grid(a, "order");
    const int ctxt_c = make_grid(a, a.count("orderc") ? "orderc" : "order");
    auto da = make_desc(a, "desca", ctxt_a, rank), dc = make_desc(a, "descc", ctxt_c, rank);
    auto la = read_file<T>(rank_file(dir, "a", rank), 1), lc = read_file<T>(rank_file(dir, "c", rank), 1);
    costa::pxgemr2d<T>(geti(a, "m"), geti(a, "n"), la.data(), geti(a, "ia"), geti(a, "ja"), da.data(), lc.data(), geti(a, "ic"), geti(a, "jc"),
                       dc.data(), ctxt_a);
    write_file(rank_file(dir, "cout", rank), lc.data(), lc.size());
    return 0;
}

template <typename T>
int run_pxtran(const Args& a, int rank) {
    const std::string dir = need(a, "dir");
    const int ctxt = make_grid(a, "order");
    auto da = make_desc(a, "desca", ctxt, rank), dc = make_desc(a, "descc", ctxt, rank);
    auto la = read_file<T>(rank_file(dir, "a", rank), 1), lc = read_file<T>(rank_file(dir, "c", rank), 1);
    const T alpha = scalar_of<T>::make(getv(a, "alpha")), beta = scalar_of<T>::make(getv(a, "beta"));
    costa::pxtran_op<T>(geti(a, "m"), geti(a, "n"), alpha, la.data(), geti(a, "ia"), geti(a, "ja"), da.data(), beta, lc.data(), geti(a, "ic"),
                        geti(a, "jc"), dc.data(), need(a, "trans").at(0));
    write_file(rank_file(dir, "cout", rank), lc.data(), lc.size());
    return 0;
}

template <typename T>
int dispatch(const Args& a, int rank, int P) {
    const std::string op = need(a, "op");
    if (op == "multiply") return run_multiply<T>(a, rank, P);
    if (op == "pxgemm") return run_pxgemm<T>(a, rank);
    if (op == "pxgemr2d") return run_pxgemr2d<T>(a, rank);
    if (op == "pxtran") return run_pxtran<T>(a, rank);
    std::cerr << "ref_driver: unknown op " << op << "\n";
    return 2;
}
}  // namespace

// a crash inside the reference (e.g. SIGFPE) should say where
static void crash_handler(int sig) {
    void* frames[48];
    const int n = backtrace(frames, 48);
    const char msg[] = "ref_driver: fatal signal inside the reference; backtrace:\n";
    (void)!write(2, msg, sizeof msg - 1);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

int main(int argc, char** argv) {
    signal(SIGFPE, crash_handler);
    signal(SIGSEGV, crash_handler);
    MPI_Init(&argc, &argv);
    int rank, P;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &P);
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s(argv[i]);
        const size_t eq = s.find('=');
        if (eq != std::string::npos) a[s.substr(0, eq)] = s.substr(eq + 1);
    }
    int rc = 2;
    try {
        const char dt = need(a, "dtype").at(0);
        if (dt == 's') rc = dispatch<float>(a, rank, P);
        else if (dt == 'd') rc = dispatch<double>(a, rank, P);
        else if (dt == 'c') rc = dispatch<std::complex<float>>(a, rank, P);
        else if (dt == 'z') rc = dispatch<std::complex<double>>(a, rank, P);
    } catch (const std::exception& e) {
        std::cerr << "ref_driver[" << rank << "]: exception: " << e.what() << "\n";
        rc = 4;
    }
    MPI_Finalize();
    return rc;
}
