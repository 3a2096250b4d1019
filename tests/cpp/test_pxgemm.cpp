// C++ / Fortran-ABI test of p?gemm: the ScaLAPACK symbols (pdgemm_, pzgemm_, psgemm, PCGEMM_ ...) called the way a ScaLAPACK
// application calls them -- BLACS grid, descinit_, numroc_, the rank's local arrays in HOST memory -- against the dense
// definition sub(C) = alpha*op(sub(A))*op(sub(B)) + beta*sub(C). Follows the reference's tests/pdgemm.cpp + utils/
// pxgemm_utils.hpp:100-189,603-637 (descriptor cases with sub-matrix offsets, different block sizes per matrix, rsrc/csrc,
// and the rule that C is not read when beta == 0: it is pre-filled with NaN). There the expected values come from the
// vendor ScaLAPACK, which this image does not have (parity unpinned against ScaLAPACK); here from the naive GEMM.
#include "cosma_test_utils.hpp"

#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <cosma/prefixed_pxgemm.h>
#include <cosma/pxgemm.h>

#include <cstdlib>
#include <fstream>
#include <limits>
#include <string>
#include <unistd.h>

using testutil::real_of;
extern "C" {
void descinit_(int* desc, const int* m, const int* n, const int* mb, const int* nb, const int* irsrc, const int* icsrc, const int* ictxt,
               const int* lld, int* info);
int numroc_(const int* n, const int* nb, const int* iproc, const int* isrcproc, const int* nprocs);
}

template <typename T> T gen(int which, int i, int j) { return static_cast<T>(std::sin(0.3 * which + 0.37 * i + 1.1 * j)); }
template <> std::complex<double> gen<std::complex<double>>(int which, int i, int j) { return {std::sin(0.3 * which + 0.37 * i + 1.1 * j), std::cos(0.7 * which + 0.2 * i - 0.9 * j)}; }
template <> std::complex<float> gen<std::complex<float>>(int which, int i, int j) { return std::complex<float>(gen<std::complex<double>>(which, i, j)); }
template <typename T> struct wide_of { using type = double; };
template <typename T> struct wide_of<std::complex<T>> { using type = std::complex<double>; };
template <typename T> const char* type_name();
template <> const char* type_name<float>() { return "float"; }
template <> const char* type_name<double>() { return "double"; }
template <> const char* type_name<std::complex<float>>() { return "complex<float>"; }
template <> const char* type_name<std::complex<double>>() { return "complex<double>"; }
template <typename T> T conj_if(const T& v, bool) { return v; }
template <typename T> std::complex<T> conj_if(const std::complex<T>& v, bool c) { return c ? std::conj(v) : v; }

template <typename T>
struct dist_matrix {
    int M, N, mb, nb, rsrc, csrc, lld, lrows, lcols;
    int desc[9];
    std::vector<T> local;
    // lld_ <= 0: local rows + 2 rows of padding
    dist_matrix(int ctxt, int M_, int N_, int mb_, int nb_, int rsrc_, int csrc_, int nprow, int npcol, int myrow, int mycol, int lld_ = 0)
        : M(M_), N(N_), mb(mb_), nb(nb_), rsrc(rsrc_), csrc(csrc_) {
        lrows = myrow >= 0 ? numroc_(&M, &mb, &myrow, &rsrc, &nprow) : 0;
        lcols = myrow >= 0 ? numroc_(&N, &nb, &mycol, &csrc, &npcol) : 0;
        lld = lld_ > 0 ? std::max(lld_, std::max(1, lrows)) : std::max(1, lrows) + 2;
        int info = 0;
        descinit_(desc, &M, &N, &mb, &nb, &rsrc, &csrc, &ctxt, &lld, &info);
        local.assign(static_cast<size_t>(lld) * std::max(1, lcols), T{-555});
    }
    // global index of local (li, lj) on process (myrow, mycol)
    static int l2g(int l, int b, int p, int src, int np) { return (l / b * np + (np + p - src) % np) * b + l % b; }
    template <typename F>
    void fill(int myrow, int mycol, int nprow, int npcol, F f) {
        for (int lj = 0; lj < lcols; ++lj)
            for (int li = 0; li < lrows; ++li)
                local[static_cast<size_t>(lj) * lld + li] = f(l2g(li, mb, myrow, rsrc, nprow), l2g(lj, nb, mycol, csrc, npcol));
    }
};

typedef void (*real_entry_d)(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*, const int*, const int*,
                             const double*, const int*, const int*, const int*, const double*, double*, const int*, const int*, const int*);
typedef void (*real_entry_s)(const char*, const char*, const int*, const int*, const int*, const float*, const float*, const int*, const int*, const int*,
                             const float*, const int*, const int*, const int*, const float*, float*, const int*, const int*, const int*);

template <typename T> struct entry;
template <> struct entry<double> { static real_entry_d get(int v) { real_entry_d e[] = {pdgemm_, pdgemm, PDGEMM_, cosma_pdgemm_}; return e[v % 4]; } };
template <> struct entry<float> { static real_entry_s get(int v) { real_entry_s e[] = {psgemm_, PSGEMM, cosma_psgemm, COSMA_PSGEMM_}; return e[v % 4]; } };
template <> struct entry<std::complex<double>> { static real_entry_d get(int v) { real_entry_d e[] = {pzgemm_, PZGEMM_, cosma_pzgemm, pzgemm}; return e[v % 4]; } };
template <> struct entry<std::complex<float>> { static real_entry_s get(int v) { real_entry_s e[] = {pcgemm_, PCGEMM, COSMA_PCGEMM, cosma_pcgemm_}; return e[v % 4]; } };

struct px_case {
    char ta, tb;
    int m, n, k;
    int a_blk[2], b_blk[2], c_blk[2];
    int ia, ja, ib, jb, ic, jc;
    int extra;  // rows/cols of the global matrices beyond the sub-matrix
    int src;    // 1: rsrc/csrc = last process row/col, 0: (0, 0)
    double alpha, beta;
};

// every argument of a p?gemm call spelled out: the parameter set of the reference's tests/pdgemm.cpp (cosma::pxgemm_params,
// src/cosma/pxgemm_params.hpp:227-290)
struct full_case {
    int ma, na, mb, nb, mc, nc;              // global matrix sizes
    int bma, bna, bmb, bnb, bmc, bnc;        // block sizes
    int ia, ja, ib, jb, ic, jc;              // sub-matrix origins (1-based)
    int m, n, k;
    char ta, tb;
    double alpha, beta;
    int lld_a, lld_b, lld_c;                 // <= 0: chosen by the test
    int p_rows, p_cols;
    char order;
    int src_ma, src_na, src_mb, src_nb, src_mc, src_nc;
};

static full_case expand(const px_case& pc, int nprow, int npcol, char order) {
    const bool tA = pc.ta != 'N', tB = pc.tb != 'N';
    const int am = tA ? pc.k : pc.m, an = tA ? pc.m : pc.k, bm = tB ? pc.n : pc.k, bn = tB ? pc.k : pc.n;
    const int rs = pc.src ? nprow - 1 : 0, cs = pc.src ? npcol - 1 : 0;
    return full_case{am + pc.ia - 1 + pc.extra, an + pc.ja - 1 + pc.extra, bm + pc.ib - 1 + pc.extra, bn + pc.jb - 1 + pc.extra,
                     pc.m + pc.ic - 1 + pc.extra, pc.n + pc.jc - 1 + pc.extra,
                     pc.a_blk[0], pc.a_blk[1], pc.b_blk[0], pc.b_blk[1], pc.c_blk[0], pc.c_blk[1],
                     pc.ia, pc.ja, pc.ib, pc.jb, pc.ic, pc.jc, pc.m, pc.n, pc.k, pc.ta, pc.tb, pc.alpha, pc.beta, 0, 0, 0, nprow, npcol, order,
                     rs, cs, 0, cs, rs, 0};
}

template <typename T>
static void run_case(int ctxt, const full_case& pc, int variant) {
    using R = typename real_of<T>::type;
    int nprow, npcol, myrow, mycol;
    cosma::blacs::Cblacs_gridinfo(ctxt, &nprow, &npcol, &myrow, &mycol);
    const bool tA = pc.ta != 'N', tB = pc.tb != 'N';
    dist_matrix<T> A(ctxt, pc.ma, pc.na, pc.bma, pc.bna, pc.src_ma, pc.src_na, nprow, npcol, myrow, mycol, pc.lld_a);
    dist_matrix<T> B(ctxt, pc.mb, pc.nb, pc.bmb, pc.bnb, pc.src_mb, pc.src_nb, nprow, npcol, myrow, mycol, pc.lld_b);
    dist_matrix<T> C(ctxt, pc.mc, pc.nc, pc.bmc, pc.bnc, pc.src_mc, pc.src_nc, nprow, npcol, myrow, mycol, pc.lld_c);
    const T alpha = static_cast<T>(pc.alpha), beta = static_cast<T>(pc.beta);
    const T nan = T(std::numeric_limits<R>::quiet_NaN());
    auto in_sub_c = [&](int gi, int gj) { return gi >= pc.ic - 1 && gi < pc.ic - 1 + pc.m && gj >= pc.jc - 1 && gj < pc.jc - 1 + pc.n; };
    if (myrow >= 0) {
        A.fill(myrow, mycol, nprow, npcol, [](int i, int j) { return gen<T>(0, i, j); });
        B.fill(myrow, mycol, nprow, npcol, [](int i, int j) { return gen<T>(1, i, j); });
        C.fill(myrow, mycol, nprow, npcol, [&](int i, int j) { return (pc.beta == 0.0 && in_sub_c(i, j)) ? nan : gen<T>(2, i, j); });
    }
    // dense expectation of sub(C), evaluated in DOUBLE precision whatever T is (a 4-byte naive loop is itself off by ~k*eps, more than
    // the library under test), together with the magnitude sum_l |a_il||b_lj| that bounds the rounding error of any summation order
    using D = typename wide_of<T>::type;
    const D alpha_d = static_cast<D>(alpha), beta_d = static_cast<D>(beta);
    std::vector<D> dA(static_cast<size_t>(pc.m) * std::max(pc.k, 1)), dB(static_cast<size_t>(std::max(pc.k, 1)) * pc.n), dC(static_cast<size_t>(pc.m) * pc.n);
    std::vector<double> bound(static_cast<size_t>(pc.m) * pc.n, 0.0);
    for (int j = 0; j < pc.k; ++j)
        for (int i = 0; i < pc.m; ++i)
            dA[static_cast<size_t>(j) * pc.m + i] = static_cast<D>(tA ? conj_if(gen<T>(0, pc.ia - 1 + j, pc.ja - 1 + i), pc.ta == 'C') : gen<T>(0, pc.ia - 1 + i, pc.ja - 1 + j));
    for (int j = 0; j < pc.n; ++j)
        for (int i = 0; i < pc.k; ++i)
            dB[static_cast<size_t>(j) * pc.k + i] = static_cast<D>(tB ? conj_if(gen<T>(1, pc.ib - 1 + j, pc.jb - 1 + i), pc.tb == 'C') : gen<T>(1, pc.ib - 1 + i, pc.jb - 1 + j));
    for (int j = 0; j < pc.n; ++j)
        for (int i = 0; i < pc.m; ++i) {
            const D c0 = pc.beta == 0.0 ? D{0} : static_cast<D>(gen<T>(2, pc.ic - 1 + i, pc.jc - 1 + j));
            D acc{0};
            double mag = 0.0;
            if (pc.alpha != 0.0)
                for (int l = 0; l < pc.k; ++l) {
                    const D a = dA[static_cast<size_t>(l) * pc.m + i], b = dB[static_cast<size_t>(j) * pc.k + l];
                    acc += a * b;
                    mag += std::abs(a) * std::abs(b);
                }
            dC[static_cast<size_t>(j) * pc.m + i] = alpha_d * acc + beta_d * c0;
            bound[static_cast<size_t>(j) * pc.m + i] = std::abs(alpha_d) * mag + std::abs(beta_d) * std::abs(c0);
        }

    entry<T>::get(variant)(&pc.ta, &pc.tb, &pc.m, &pc.n, &pc.k, reinterpret_cast<const R*>(&alpha), reinterpret_cast<const R*>(A.local.data()), &pc.ia, &pc.ja,
                           A.desc, reinterpret_cast<const R*>(B.local.data()), &pc.ib, &pc.jb, B.desc, reinterpret_cast<const R*>(&beta),
                           reinterpret_cast<R*>(C.local.data()), &pc.ic, &pc.jc, C.desc);

    bool ok = true, untouched = true, pad = true;
    double worst = 0.0;
    // component-wise: |got - want| <= c * (|alpha| sum|a||b| + |beta||c|), c = 2e-6 for 4-byte reals (3xTF32 split: ~2^-21 per product
    // plus the result's own rounding), 1e-13 for 8-byte ones (north_star's tolerances, stated per element instead of per norm)
    const double rel = sizeof(R) == 4 ? 2e-6 : 1e-13;
    if (myrow >= 0) {
        for (int lj = 0; lj < C.lcols; ++lj) {
            for (int li = 0; li < C.lld; ++li) {
                const T got = C.local[static_cast<size_t>(lj) * C.lld + li];
                if (li >= C.lrows) { pad = pad && got == T{-555}; continue; }
                const int gi = dist_matrix<T>::l2g(li, C.mb, myrow, C.rsrc, nprow), gj = dist_matrix<T>::l2g(lj, C.nb, mycol, C.csrc, npcol);
                if (in_sub_c(gi, gj)) {
                    const size_t at = static_cast<size_t>(gj - pc.jc + 1) * pc.m + (gi - pc.ic + 1);
                    const double err = std::abs(static_cast<D>(got) - dC[at]), tol = rel * bound[at] + std::numeric_limits<R>::min();
                    if (!(err <= tol)) {
                        ok = false;
                        worst = std::max(worst, bound[at] > 0 ? err / bound[at] : err);
                    }
                }
                else untouched = untouched && got == gen<T>(2, gi, gj);
            }
        }
    }
    CHECK_MSG(ok, "p?gemm " << type_name<T>() << " " << pc.ta << pc.tb << " " << pc.m << "x" << pc.n << "x" << pc.k << " variant " << variant << " grid " << nprow << "x" << npcol
                            << " worst err/bound " << worst << " (allowed " << rel << ")");
    CHECK_TRUE(untouched);
    CHECK_TRUE(pad);
}

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int rank = 0, P = 1;
    cosma::blacs::Cblacs_pinfo(&rank, &P);
    int nprow = 1;
    for (int d = 1; d * d <= P; ++d)
        if (P % d == 0) nprow = d;
    const int npcol = P / nprow;
    const px_case cases[] = {
        {'N', 'N', 64, 48, 40, {8, 8}, {8, 8}, {8, 8}, 1, 1, 1, 1, 1, 1, 0, 0, 1.0, 0.0},
        {'T', 'N', 37, 53, 29, {5, 7}, {4, 9}, {6, 3}, 3, 2, 2, 6, 4, 5, 11, 1, 2.0, -1.0},
        {'N', 'T', 45, 30, 61, {16, 4}, {8, 32}, {7, 7}, 1, 5, 7, 1, 2, 2, 3, 0, 1.0, 1.0},
        {'C', 'C', 33, 41, 27, {6, 6}, {5, 5}, {9, 4}, 2, 3, 4, 1, 1, 8, 5, 1, -0.5, 0.0},
        {'N', 'N', 50, 50, 0, {8, 8}, {8, 8}, {8, 8}, 1, 1, 1, 1, 3, 3, 4, 0, 1.0, 2.0},   // k == 0: sub(C) *= beta
        {'N', 'N', 20, 24, 16, {8, 8}, {8, 8}, {8, 8}, 1, 1, 1, 1, 1, 1, 0, 0, 0.0, 0.0},  // alpha == 0, beta == 0: zeros, NaN not read
        {'T', 'T', 128, 96, 200, {32, 32}, {64, 16}, {32, 8}, 1, 1, 1, 1, 1, 1, 0, 0, 1.0, 0.0},
    };
    for (const char order : {'R', 'C'}) {
        int ctxt = 0;
        cosma::blacs::Cblacs_get(0, 0, &ctxt);
        char ord = order;
        cosma::blacs::Cblacs_gridinit(&ctxt, &ord, nprow, npcol);
        int v = 0;
        for (const auto& pc : cases) {
            const full_case fc = expand(pc, nprow, npcol, order);
            run_case<double>(ctxt, fc, v);
            run_case<std::complex<double>>(ctxt, fc, v);
            run_case<float>(ctxt, fc, v);
            run_case<std::complex<float>>(ctxt, fc, v);
            ++v;
        }
        cosma::pxgemm_release_grids();
        cosma::blacs::Cblacs_gridexit(ctxt);
    }
    // the reference's own 50 parameter sets (tests/pdgemm.cpp, extracted by tests/golden/make_pdgemm_cases.py): each on the grid
    // it names, made of the first p_rows x p_cols ranks of the job; sets needing more ranks than the job has are skipped
    {
        const char* env = std::getenv("COSMA_B200_PDGEMM_CASES");
        // the fixture is found relative to this BINARY (tests/cpp/bin/ -> tests/golden/), not to the working directory
        std::string fixture = env && *env ? env : "";
        if (fixture.empty()) {
            char self[4096];
            const ssize_t len = ::readlink("/proc/self/exe", self, sizeof(self) - 1);
            std::string dir = len > 0 ? std::string(self, static_cast<size_t>(len)) : std::string(argv[0]);
            const size_t slash = dir.rfind('/');
            dir = slash == std::string::npos ? "." : dir.substr(0, slash);
            fixture = dir + "/../../golden/pdgemm_cases.txt";
        }
        std::ifstream in(fixture);
        CHECK_MSG(in.good(), "cannot open the reference parameter sets: " << fixture);
        full_case fc;
        int v = 0, ran = 0;
        while (in >> fc.ma >> fc.na >> fc.mb >> fc.nb >> fc.mc >> fc.nc >> fc.bma >> fc.bna >> fc.bmb >> fc.bnb >> fc.bmc >> fc.bnc >> fc.ia >> fc.ja >>
               fc.ib >> fc.jb >> fc.ic >> fc.jc >> fc.m >> fc.n >> fc.k >> fc.ta >> fc.tb >> fc.alpha >> fc.beta >> fc.lld_a >> fc.lld_b >> fc.lld_c >>
               fc.p_rows >> fc.p_cols >> fc.order >> fc.src_ma >> fc.src_na >> fc.src_mb >> fc.src_nb >> fc.src_mc >> fc.src_nc) {
            ++v;
            if (fc.p_rows * fc.p_cols > P) { ++check::skipped(); continue; }
            int ctxt = 0;
            cosma::blacs::Cblacs_get(0, 0, &ctxt);
            cosma::blacs::Cblacs_gridinit(&ctxt, &fc.order, fc.p_rows, fc.p_cols);
            run_case<double>(ctxt, fc, v);
            if (fc.k <= 2000) run_case<std::complex<float>>(ctxt, fc, v);
            cosma::pxgemm_release_grids();
            cosma::blacs::Cblacs_gridexit(ctxt);
            ++ran;
        }
        CHECK_MSG(v == 50, "expected the reference's 50 parameter sets in " << fixture << ", read " << v);
        if (rank == 0) std::printf("reference pdgemm parameter sets: %d read, %d run on %d rank(s)\n", v, ran, P);
    }
    cosma::b200::release_all_comms();
    const int rc = check::finish("test_pxgemm");
    MPI_Finalize();
    return rc;
}
