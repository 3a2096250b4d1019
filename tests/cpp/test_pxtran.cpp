// C++ / Fortran-ABI test of COSTA's ScaLAPACK wrappers: pdtran_, pstran, pztranu_, pctranc, costa_pztranc_ ... and
// p?gemr2d[_] called the ScaLAPACK way (BLACS grids, descriptors, host-resident local arrays) and checked, rank by rank, against
// the definition on analytically generated matrices -- pure moves bit for bit. The reference checks these wrappers only in its
// miniapps against a vendor ScaLAPACK (libs/COSTA/miniapps/pxtran_miniapp.cpp, pxgemr2d_miniapp.cpp), absent here; the unmodified
// reference wrappers themselves are compared with OUR plans per rank in tests/test_pxtran_cpu.py.
#include "check.hpp"

#include <cosma/b200_runtime.hpp>
#include <cosma/cosma_pxgemm.hpp>
#include <costa/pxgemr2d/prefixed_pxgemr2d.h>
#include <costa/pxgemr2d/pxgemr2d.h>
#include <costa/pxtran/prefixed_pxtran.h>
#include <costa/pxtran/pxtran.h>
#include <costa/pxtranc/prefixed_pxtranc.h>
#include <costa/pxtranc/pxtranc.h>
#include <costa/pxtranu/pxtranu.h>

#include <cmath>
#include <complex>
#include <vector>

extern "C" {
void descinit_(int* desc, const int* m, const int* n, const int* mb, const int* nb, const int* irsrc, const int* icsrc, const int* ictxt,
               const int* lld, int* info);
int numroc_(const int* n, const int* nb, const int* iproc, const int* isrcproc, const int* nprocs);
// the Fortran entry points of BLACS (libcosma_blacs_lite.so supplies them when no BLACS is linked)
void blacs_pinfo_(int* mypnum, int* nprocs);
void blacs_get_(const int* ictxt, const int* what, int* val);
void blacs_gridinit_(int* ictxt, const char* order, const int* nprow, const int* npcol);
void blacs_gridinfo_(const int* ictxt, int* nprow, int* npcol, int* myrow, int* mycol);
int blacs_pnum_(const int* ictxt, const int* prow, const int* pcol);
void blacs_pcoord_(const int* ictxt, const int* pnum, int* prow, int* pcol);
void blacs_gridexit_(const int* ictxt);
}

template <typename T> struct real_of { using type = T; };
template <typename T> struct real_of<std::complex<T>> { using type = T; };
template <typename T> T gen(int which, int i, int j) { return static_cast<T>(std::sin(0.3 * which + 0.37 * i + 1.1 * j)); }
template <> std::complex<double> gen<std::complex<double>>(int which, int i, int j) { return {std::sin(0.3 * which + 0.37 * i + 1.1 * j), std::cos(0.7 * which + 0.2 * i - 0.9 * j)}; }
template <> std::complex<float> gen<std::complex<float>>(int which, int i, int j) { return std::complex<float>(gen<std::complex<double>>(which, i, j)); }
template <typename T> T conj_if(const T& v, bool) { return v; }
template <typename T> std::complex<T> conj_if(const std::complex<T>& v, bool c) { return c ? std::conj(v) : v; }

template <typename T>
struct dist_matrix {
    int M, N, mb, nb, lld, lrows, lcols, myrow, mycol, nprow, npcol;
    int desc[9];
    std::vector<T> local;
    dist_matrix(int ctxt, int M_, int N_, int mb_, int nb_, int pad) : M(M_), N(N_), mb(mb_), nb(nb_) {
        cosma::blacs::Cblacs_gridinfo(ctxt, &nprow, &npcol, &myrow, &mycol);
        const int zero = 0;
        lrows = myrow >= 0 ? numroc_(&M, &mb, &myrow, &zero, &nprow) : 0;
        lcols = myrow >= 0 ? numroc_(&N, &nb, &mycol, &zero, &npcol) : 0;
        lld = std::max(1, lrows) + pad;
        int info = 0;
        descinit_(desc, &M, &N, &mb, &nb, &zero, &zero, &ctxt, &lld, &info);
        local.assign(static_cast<size_t>(lld) * std::max(1, lcols), T{-555});
    }
    static int l2g(int l, int b, int p, int np) { return (l / b * np + p) * b + l % b; }
    int grow(int li) const { return l2g(li, mb, myrow, nprow); }
    int gcol(int lj) const { return l2g(lj, nb, mycol, npcol); }
    template <typename F> void fill(F f) {
        for (int lj = 0; lj < lcols; ++lj)
            for (int li = 0; li < lrows; ++li) local[static_cast<size_t>(lj) * lld + li] = f(grow(li), gcol(lj));
    }
    // every local element equals want(gi, gj) within tol (0: bit for bit); padding rows untouched
    template <typename F> bool holds(F want, double tol) const {
        bool ok = true;
        for (int lj = 0; lj < lcols; ++lj)
            for (int li = 0; li < lld; ++li) {
                const T got = local[static_cast<size_t>(lj) * lld + li];
                if (li >= lrows) ok = ok && got == T{-555};
                else ok = ok && std::abs(got - want(grow(li), gcol(lj))) <= tol;
            }
        return ok;
    }
};

typedef void (*tran_d)(const int*, const int*, const double*, const double*, const int*, const int*, const int*, const double*, double*, const int*, const int*, const int*);
typedef void (*tran_s)(const int*, const int*, const float*, const float*, const int*, const int*, const int*, const float*, float*, const int*, const int*, const int*);

// sub(C) = C(ic:ic+m-1, jc:jc+n-1) = beta * sub(C) + alpha * op(A(ia:ia+n-1, ja:ja+m-1))
template <typename T, typename Entry>
static void tran_case(int ctxt, Entry entry, bool conj, int m, int n, int ia, int ja, int ic, int jc, T alpha, T beta, const char* what) {
    using R = typename real_of<T>::type;
    dist_matrix<T> A(ctxt, n + ia - 1 + 3, m + ja - 1 + 2, 5, 7, 2), C(ctxt, m + ic - 1 + 1, n + jc - 1 + 4, 4, 9, 1);
    A.fill([](int i, int j) { return gen<T>(0, i, j); });
    C.fill([](int i, int j) { return gen<T>(1, i, j); });
    entry(&m, &n, reinterpret_cast<const R*>(&alpha), reinterpret_cast<const R*>(A.local.data()), &ia, &ja, A.desc, reinterpret_cast<const R*>(&beta),
          reinterpret_cast<R*>(C.local.data()), &ic, &jc, C.desc);
    const bool pure = alpha == T{1} && beta == T{0};
    auto want = [&](int gi, int gj) -> T {
        const int i = gi - (ic - 1), j = gj - (jc - 1);
        if (i < 0 || i >= m || j < 0 || j >= n) return gen<T>(1, gi, gj);  // outside sub(C): untouched
        const T s = conj_if(gen<T>(0, ia - 1 + j, ja - 1 + i), conj);
        if (pure) return s;
        return beta == T{0} ? alpha * s : beta * gen<T>(1, gi, gj) + alpha * s;
    };
    CHECK_MSG(C.holds(want, pure ? 0.0 : (sizeof(R) == 4 ? 1e-5 : 1e-14)), what);
}

typedef void (*gemr2d_d)(const int*, const int*, const double*, const int*, const int*, const int*, double*, const int*, const int*, const int*, const int*);
typedef void (*gemr2d_s)(const int*, const int*, const float*, const int*, const int*, const int*, float*, const int*, const int*, const int*, const int*);

template <typename T, typename Entry>
static void gemr2d_case(int ctxt_a, int ctxt_c, Entry entry, int m, int n, int ia, int ja, int ic, int jc, const char* what) {
    using R = typename real_of<T>::type;
    dist_matrix<T> A(ctxt_a, m + ia - 1 + 2, n + ja - 1 + 5, 6, 4, 3), C(ctxt_c, m + ic - 1 + 3, n + jc - 1 + 1, 3, 8, 0);
    A.fill([](int i, int j) { return gen<T>(2, i, j); });
    C.fill([](int i, int j) { return gen<T>(3, i, j); });
    entry(&m, &n, reinterpret_cast<const R*>(A.local.data()), &ia, &ja, A.desc, reinterpret_cast<R*>(C.local.data()), &ic, &jc, C.desc, &ctxt_a);
    auto want = [&](int gi, int gj) -> T {
        const int i = gi - (ic - 1), j = gj - (jc - 1);
        if (i < 0 || i >= m || j < 0 || j >= n) return gen<T>(3, gi, gj);
        return gen<T>(2, ia - 1 + i, ja - 1 + j);
    };
    CHECK_MSG(C.holds(want, 0.0), what);
}

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int rank = 0, P = 1;
    cosma::blacs::Cblacs_pinfo(&rank, &P);
    int nprow = 1;
    for (int d = 1; d * d <= P; ++d)
        if (P % d == 0) nprow = d;
    const int npcol = P / nprow;
    int row_major = 0, col_major = 0;
    char R = 'R', C = 'C';
    cosma::blacs::Cblacs_get(0, 0, &row_major);
    cosma::blacs::Cblacs_gridinit(&row_major, &R, nprow, npcol);
    cosma::blacs::Cblacs_get(0, 0, &col_major);
    cosma::blacs::Cblacs_gridinit(&col_major, &C, npcol, nprow);  // the transposed grid shape, numbered column-major

    using zd = std::complex<double>;
    using zf = std::complex<float>;
    for (int ctxt : {row_major, col_major}) {
        tran_case<double, tran_d>(ctxt, pdtran_, false, 40, 56, 1, 1, 1, 1, 1.0, 0.0, "pdtran_ (move)");
        tran_case<double, tran_d>(ctxt, PDTRAN, false, 37, 53, 3, 2, 2, 6, 2.0, -1.0, "PDTRAN (scaled, offsets)");
        tran_case<double, tran_d>(ctxt, costa_pdtran, false, 64, 16, 1, 5, 7, 1, 1.0, 1.0, "costa_pdtran (accumulate)");
        tran_case<float, tran_s>(ctxt, pstran, false, 33, 29, 2, 2, 1, 3, 1.0f, 0.0f, "pstran (move)");
        tran_case<zd, tran_d>(ctxt, pztranu_, false, 21, 34, 1, 1, 4, 2, zd(1.0, 0.0), zd(0.0, 0.0), "pztranu_ (move)");
        tran_case<zd, tran_d>(ctxt, pztranc, true, 21, 34, 2, 3, 1, 1, zd(1.0, 0.0), zd(0.0, 0.0), "pztranc (move, conjugated)");
        tran_case<zd, tran_d>(ctxt, costa_pztranc_, true, 30, 30, 1, 1, 1, 1, zd(0.5, -1.5), zd(1.0, 2.0), "costa_pztranc_ (scaled)");
        tran_case<zf, tran_s>(ctxt, PCTRANC_, true, 18, 27, 1, 2, 3, 1, zf(1.0f, 0.0f), zf(0.0f, 0.0f), "PCTRANC_ (move, conjugated)");
        tran_case<zf, tran_s>(ctxt, pctranu, false, 18, 27, 1, 2, 3, 1, zf(2.0f, 1.0f), zf(0.0f, 0.0f), "pctranu (scaled)");
    }
    // redistribution between the two grids (and within one)
    gemr2d_case<double, gemr2d_d>(row_major, col_major, pdgemr2d_, 45, 38, 1, 1, 1, 1, "pdgemr2d_ R -> C");
    gemr2d_case<double, gemr2d_d>(col_major, row_major, PDGEMR2D, 45, 38, 2, 4, 3, 1, "PDGEMR2D C -> R (offsets)");
    gemr2d_case<float, gemr2d_s>(row_major, row_major, psgemr2d, 30, 41, 1, 3, 2, 2, "psgemr2d R -> R");
    gemr2d_case<zd, gemr2d_d>(row_major, col_major, costa_pzgemr2d_, 27, 27, 3, 3, 1, 2, "costa_pzgemr2d_ R -> C");
    gemr2d_case<zf, gemr2d_s>(col_major, row_major, pcgemr2d_, 19, 50, 1, 1, 5, 1, "pcgemr2d_ C -> R");

    {   // a grid set up the way a Fortran application does it: same numbering as the C calls, and the wrappers work on it
        int me = -1, np = 0, ctxt = 0;
        const int zero = 0;
        blacs_pinfo_(&me, &np);
        CHECK_MSG(me == rank && np == P, "blacs_pinfo_");
        blacs_get_(&zero, &zero, &ctxt);
        blacs_gridinit_(&ctxt, "Col-major", &nprow, &npcol);
        int r1, c1, pr, pc, r2, c2, pr2, pc2;
        blacs_gridinfo_(&ctxt, &r1, &c1, &pr, &pc);
        cosma::blacs::Cblacs_gridinfo(ctxt, &r2, &c2, &pr2, &pc2);
        CHECK_MSG(r1 == nprow && c1 == npcol && r1 == r2 && c1 == c2 && pr == pr2 && pc == pc2, "blacs_gridinfo_ == Cblacs_gridinfo");
        CHECK_MSG(blacs_pnum_(&ctxt, &pr, &pc) == rank, "blacs_pnum_ of my coordinates");
        int qr = -1, qc = -1;
        blacs_pcoord_(&ctxt, &rank, &qr, &qc);
        CHECK_MSG(qr == pr && qc == pc && rank == pc * nprow + pr, "blacs_pcoord_ (column-major numbering)");
        tran_case<double, tran_d>(ctxt, pdtran_, false, 26, 31, 2, 1, 1, 3, 1.0, 0.0, "pdtran_ on a grid from blacs_gridinit_");
        blacs_gridexit_(&ctxt);
    }
    cosma::pxgemm_release_grids();
    cosma::blacs::Cblacs_gridexit(row_major);
    cosma::blacs::Cblacs_gridexit(col_major);
    cosma::b200::release_all_comms();
    const int rc = check::finish("test_pxtran");
    MPI_Finalize();
    return rc;
}
