// C++ API tests of the layout-based entry points:
//   (1) the reference's own test (tests/multiply_using_layout.cpp:46-125): multiply_using_layout on the native COSMA grids of
//       A, B, C gives what multiply() gives (beta = 1), compared with ASSERT_DOUBLE_EQ there, to 4 ULP here as well;
//   (2) block-cyclic layouts built with costa::block_cyclic_layout<T>, op(A) in {N, T, C}, host-resident blocks, all four
//       types, validated against the naive GEMM of analytically defined matrices;
//   (3) the C interface dmultiply_using_layout / zmultiply_using_layout on the same data (cinterface.hpp);
//   (4) costa::transform between two block-cyclic layouts with transposition, alpha and beta: bit-exact for pure moves.
#include "cosma_test_utils.hpp"

#include <cosma/b200_runtime.hpp>
#include <cosma/cinterface.hpp>
#include <costa/grid2grid/transformer.hpp>
#include <costa/layout.hpp>

#include <cstdint>
#include <cstring>
#include <limits>

using testutil::real_of;

template <typename T> T value_a(int i, int j) { return static_cast<T>(std::sin(0.37 * i + 1.3 * j)); }
template <typename T> T value_b(int i, int j) { return static_cast<T>(std::cos(0.91 * i - 0.53 * j)); }
template <typename T> T value_c(int i, int j) { return static_cast<T>(0.25 * std::sin(0.11 * i * j + 0.7)); }
template <> std::complex<double> value_a<std::complex<double>>(int i, int j) { return {std::sin(0.37 * i + 1.3 * j), std::cos(0.2 * i - j)}; }
template <> std::complex<double> value_b<std::complex<double>>(int i, int j) { return {std::cos(0.91 * i - 0.53 * j), std::sin(0.4 * j + i)}; }
template <> std::complex<double> value_c<std::complex<double>>(int i, int j) { return {0.25 * std::sin(0.11 * i * j + 0.7), 0.5 * std::cos(0.3 * i + j)}; }
template <> std::complex<float> value_a<std::complex<float>>(int i, int j) { return std::complex<float>(value_a<std::complex<double>>(i, j)); }
template <> std::complex<float> value_b<std::complex<float>>(int i, int j) { return std::complex<float>(value_b<std::complex<double>>(i, j)); }
template <> std::complex<float> value_c<std::complex<float>>(int i, int j) { return std::complex<float>(value_c<std::complex<double>>(i, j)); }

template <typename T> T conj_if(const T& v, bool) { return v; }
template <typename T> std::complex<T> conj_if(const std::complex<T>& v, bool c) { return c ? std::conj(v) : v; }

// process grid for P ranks: the most square nprow x npcol with nprow <= npcol
static void grid_shape(int P, int* nprow, int* npcol) {
    int r = 1;
    for (int d = 1; d * d <= P; ++d)
        if (P % d == 0) r = d;
    *nprow = r;
    *npcol = P / r;
}

static bool within_ulps(double a, double b, int ulps) {
    if (a == b) return true;
    const double diff = std::abs(a - b), scale = std::max(std::abs(a), std::abs(b));
    return diff <= ulps * std::numeric_limits<double>::epsilon() * scale;
}

// (1)
static void native_grids_match_multiply(MPI_Comm world) {
    using scalar_t = double;
    int rank = 0, size = 1;
    MPI_Comm_rank(world, &rank);
    MPI_Comm_size(world, &size);
    const int nprocs = std::min(4, size), m = 20, n = 20, k = 80;
    MPI_Comm comm = testutil::subcommunicator(nprocs, world);
    if (rank >= nprocs) return;
    cosma::Strategy strategy(m, n, k, nprocs);
    auto ctx = cosma::make_context<scalar_t>();
    cosma::CosmaMatrix<scalar_t> A(ctx, 'A', strategy, rank), B(ctx, 'B', strategy, rank), C(ctx, 'C', strategy, rank), C_act(ctx, 'C', strategy, rank);
    for (size_t i = 0; i < A.matrix_size(); ++i) A.matrix_pointer()[i] = std::sin(static_cast<double>(i));
    for (size_t i = 0; i < B.matrix_size(); ++i) B.matrix_pointer()[i] = std::sin(static_cast<double>(i));
    for (size_t i = 0; i < C.matrix_size(); ++i) C_act.matrix_pointer()[i] = C.matrix_pointer()[i] = std::sin(static_cast<double>(i));
    auto A_grid = A.get_grid_layout(), B_grid = B.get_grid_layout(), C_grid = C.get_grid_layout();
    cosma::multiply_using_layout(A_grid, B_grid, C_grid, scalar_t{1}, scalar_t{1}, 'N', 'N', comm);
    cosma::multiply(A, B, C_act, strategy, comm, scalar_t{1}, scalar_t{1});
    CHECK_TRUE(C.matrix_size() == C_act.matrix_size());
    bool same = true;
    for (size_t i = 0; i < C_act.matrix_size(); ++i) same = same && within_ulps(C.matrix_pointer()[i], C_act.matrix_pointer()[i], 4);
    CHECK_TRUE(same);
    cosma::b200::release_comm(comm);
    MPI_Comm_free(&comm);
}

template <typename T>
struct cyclic {  // one block-cyclic matrix: local array + layout
    int rows, cols, bm, bn, nprow, npcol, lld;
    std::vector<T> local;
    costa::grid_layout<T> layout;
    cyclic(int rows_, int cols_, int bm_, int bn_, int nprow_, int npcol_, char order, int rank, int pad = 3)
        : rows(rows_), cols(cols_), bm(bm_), bn(bn_), nprow(nprow_), npcol(npcol_) {
        int myrow = 0, mycol = 0;
        costa::rank_to_grid(rank, nprow, npcol, order, &myrow, &mycol);
        const int lr = costa::numroc(rows, bm, myrow, 0, nprow), lc = costa::numroc(cols, bn, mycol, 0, npcol);
        lld = std::max(lr, 1) + pad;
        local.assign(static_cast<size_t>(lld) * std::max(lc, 1), T{-777});
        layout = costa::block_cyclic_layout<T>(rows, cols, bm, bn, 1, 1, rows, cols, nprow, npcol, order, 0, 0, local.data(), lld, 'C', rank);
    }
};

// (2) + (3)
template <typename T>
static void block_cyclic_case(MPI_Comm comm, char ta, char tb, int m, int n, int k, T alpha, T beta, bool c_interface) {
    int rank = 0, P = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &P);
    int pr, pc;
    grid_shape(P, &pr, &pc);
    const bool tA = ta != 'N', tB = tb != 'N';
    cyclic<T> A(tA ? k : m, tA ? m : k, 5, 7, pr, pc, 'R', rank), B(tB ? n : k, tB ? k : n, 6, 4, pc, pr, 'C', rank), C(m, n, 8, 3, pr, pc, 'R', rank);
    A.layout.initialize([](int i, int j) { return value_a<T>(i, j); });
    B.layout.initialize([](int i, int j) { return value_b<T>(i, j); });
    if (beta == T{0}) C.layout.initialize([](int, int) { return T(std::numeric_limits<typename real_of<T>::type>::quiet_NaN()); });
    else C.layout.initialize([](int i, int j) { return value_c<T>(i, j); });
    // dense expectation (every rank computes it; the sizes are small)
    std::vector<T> dA(static_cast<size_t>(m) * k), dB(static_cast<size_t>(k) * n), dC(static_cast<size_t>(m) * n);
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < m; ++i) dA[static_cast<size_t>(j) * m + i] = tA ? conj_if(value_a<T>(j, i), ta == 'C') : value_a<T>(i, j);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < k; ++i) dB[static_cast<size_t>(j) * k + i] = tB ? conj_if(value_b<T>(j, i), tb == 'C') : value_b<T>(i, j);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) dC[static_cast<size_t>(j) * m + i] = beta == T{0} ? T{0} : value_c<T>(i, j);
    testutil::naive_gemm('N', 'N', m, n, k, alpha, dA.data(), m, dB.data(), k, beta, dC.data(), m);

    if (!c_interface) {
        cosma::multiply_using_layout(A.layout, B.layout, C.layout, alpha, beta, ta, tb, comm);
    } else {
        auto to_c = [](costa::grid_layout<T>& L, std::vector<block>& blocks) {
            for (size_t b = 0; b < L.blocks.num_blocks(); ++b) {
                auto& v = L.blocks.get_block(b);
                blocks.push_back(block{v.data, v.stride, v.coordinates.first, v.coordinates.second});
            }
            return layout{L.num_blocks_row(), L.num_blocks_col(), L.grid.grid.rows_split.data(), L.grid.grid.cols_split.data(), L.grid.owners.data(),
                          static_cast<int>(blocks.size()), blocks.data()};
        };
        std::vector<block> ba, bb, bc;
        layout la = to_c(A.layout, ba), lb = to_c(B.layout, bb), lc = to_c(C.layout, bc);
        using R = typename real_of<T>::type;
        if (std::is_same<T, double>::value) dmultiply_using_layout(comm, &ta, &tb, reinterpret_cast<const double*>(&alpha), &la, &lb, reinterpret_cast<const double*>(&beta), &lc);
        else if (std::is_same<T, float>::value) smultiply_using_layout(comm, &ta, &tb, reinterpret_cast<const float*>(&alpha), &la, &lb, reinterpret_cast<const float*>(&beta), &lc);
        else if (std::is_same<R, double>::value) zmultiply_using_layout(comm, &ta, &tb, reinterpret_cast<const double*>(&alpha), &la, &lb, reinterpret_cast<const double*>(&beta), &lc);
        else cmultiply_using_layout(comm, &ta, &tb, reinterpret_cast<const float*>(&alpha), &la, &lb, reinterpret_cast<const float*>(&beta), &lc);
    }
    const double tol = sizeof(typename real_of<T>::type) == 4 ? 2e-4 : 1e-11;
    const bool ok = C.layout.validate([&](int i, int j) { return dC[static_cast<size_t>(j) * m + i]; }, tol);
    CHECK_MSG(ok, "multiply_using_layout " << ta << tb << " " << m << "x" << n << "x" << k << (c_interface ? " (C interface)" : ""));
    // the padding rows of the local array (between the local rows and lld) still hold the sentinel
    bool pad_ok = true;
    int myrow = 0, mycol = 0;
    costa::rank_to_grid(rank, pr, pc, 'R', &myrow, &mycol);
    const int lr = costa::numroc(m, 8, myrow, 0, pr), lc_ = costa::numroc(n, 3, mycol, 0, pc);
    for (int j = 0; j < lc_; ++j)
        for (int i = lr; i < C.lld; ++i) pad_ok = pad_ok && C.local[static_cast<size_t>(j) * C.lld + i] == T{-777};
    CHECK_TRUE(pad_ok);
}

// (4)
template <typename T>
static void transform_case(MPI_Comm comm, char op, T alpha, T beta) {
    int rank = 0, P = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &P);
    int pr, pc;
    grid_shape(P, &pr, &pc);
    const int m = 53, n = 38;  // target is m x n; source is n x m when transposed
    const bool t = op != 'N';
    cyclic<T> src(t ? n : m, t ? m : n, 4, 9, pr, pc, 'R', rank), dst(m, n, 7, 5, pc, pr, 'C', rank, 1);
    src.layout.initialize([](int i, int j) { return value_a<T>(i, j); });
    dst.layout.initialize([](int i, int j) { return value_c<T>(i, j); });
    costa::transformer<T> tr(comm);
    tr.schedule(src.layout, dst.layout, op, alpha, beta);
    tr.transform();
    // the reference's operation order: beta * dst + alpha * op(src), products rounded separately
    auto expect = [&](int i, int j) {
        const T s = t ? conj_if(value_a<T>(j, i), op == 'C') : value_a<T>(i, j);
        if (alpha == T{1} && beta == T{0}) return s;
        return beta == T{0} ? alpha * s : beta * value_c<T>(i, j) + alpha * s;
    };
    const bool exact = alpha == T{1} && beta == T{0};
    const bool ok = dst.layout.validate(expect, exact ? 0.0 : (sizeof(typename real_of<T>::type) == 4 ? 1e-6 : 1e-14));
    CHECK_MSG(ok, "costa::transform op " << op << (exact ? " (bit-exact move)" : " (scaled)"));
}

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    MPI_Comm world = MPI_COMM_WORLD;
    native_grids_match_multiply(world);
    using zd = std::complex<double>;
    using zf = std::complex<float>;
    for (int ci = 0; ci < 2; ++ci) {
        block_cyclic_case<double>(world, 'N', 'N', 61, 47, 39, 1.0, 0.0, ci == 1);
        block_cyclic_case<double>(world, 'T', 'N', 40, 52, 77, 0.5, 2.0, ci == 1);
        block_cyclic_case<double>(world, 'N', 'T', 33, 65, 20, -1.0, 1.0, ci == 1);
        block_cyclic_case<float>(world, 'T', 'T', 45, 31, 58, 1.0f, 0.0f, ci == 1);
        block_cyclic_case<zd>(world, 'C', 'N', 37, 41, 29, zd(1.0, -0.5), zd(0.0, 0.0), ci == 1);
        block_cyclic_case<zd>(world, 'N', 'C', 30, 30, 64, zd(0.3, 0.2), zd(1.0, 1.0), ci == 1);
        block_cyclic_case<zf>(world, 'C', 'T', 26, 35, 44, zf(1.0f, 0.0f), zf(0.0f, 0.0f), ci == 1);
    }
    transform_case<double>(world, 'N', 1.0, 0.0);
    transform_case<double>(world, 'T', 1.0, 0.0);
    transform_case<double>(world, 'T', 2.0, -1.0);
    transform_case<float>(world, 'T', 1.0f, 0.0f);
    transform_case<zd>(world, 'C', zd(1.0, 0.0), zd(0.0, 0.0));
    transform_case<zd>(world, 'C', zd(0.5, 1.5), zd(1.0, -1.0));
    transform_case<zf>(world, 'N', zf(1.0f, 0.0f), zf(0.0f, 0.0f));
    cosma::b200::release_all_comms();
    const int rc = check::finish("test_multiply_using_layout");
    MPI_Finalize();
    return rc;
}
