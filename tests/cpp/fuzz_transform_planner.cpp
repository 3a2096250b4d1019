// TEST INFRASTRUCTURE: costa::plan_transform (host/costa_transform.cpp, the message-list planner that runs inside libcosma_b200.so for
// every costa::transform / multiply_using_layout / p?gemm / p?tran / p?gemr2d) under AddressSanitizer + UBSan with random layouts:
// random grids and owners ("custom" layouts) and block-cyclic sub-matrices (erased_scalapack_layout), op N / T / C, both storage
// orders, 1-9 ranks. Every rank's plan is built and checked structurally:
//   * every piece lies inside the block it reads / writes (real allocations of exactly the blocks' sizes back the layouts, so a piece
//     outside its block would also be an address ASan knows nothing about) and inside the send / receive buffer;
//   * what rank r sends to p is what p expects from r, byte for byte;  * the elements moved add up to the matrix.
//   fuzz_transform_planner SEED N
#include <costa/erased_layout.hpp>
#include <costa/transform_plan.hpp>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <stdexcept>
#include <vector>

namespace {

std::mt19937 rng;
int pick(int lo, int hi) { return lo + static_cast<int>(rng() % static_cast<unsigned>(hi - lo + 1)); }

struct Dist {  // one distributed matrix: the layout of every rank + the storage behind it
    std::vector<costa::erased_layout> of_rank;
    std::vector<std::unique_ptr<char[]>> storage;
    std::vector<std::pair<char*, size_t>> extents;  // (base, bytes) of every allocation
};

std::vector<int> random_split(int total, int parts) {
    std::vector<int> cut{0, total};
    for (int i = 1; i < parts; ++i) cut.push_back(pick(0, total));
    std::sort(cut.begin(), cut.end());
    cut.erase(std::unique(cut.begin(), cut.end()), cut.end());
    return cut;
}

Dist custom(int rows, int cols, int P, int eb) {
    Dist d;
    const std::vector<int> rs = random_split(rows, pick(1, 5)), cs = random_split(cols, pick(1, 5));
    const int nr = static_cast<int>(rs.size()) - 1, nc = static_cast<int>(cs.size()) - 1;
    std::vector<int> owners(static_cast<size_t>(nr) * nc);
    for (auto& o : owners) o = pick(0, P - 1);
    const char ordering = rng() % 2 ? 'C' : 'R';
    for (int r = 0; r < P; ++r) {
        std::vector<int> bi, bj;
        std::vector<void*> data;
        std::vector<std::int64_t> ld;
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc; ++j)
                if (owners[static_cast<size_t>(i) * nc + j] == r) {
                    const int h = rs[i + 1] - rs[i], w = cs[j + 1] - cs[j];
                    const std::int64_t lead = (ordering == 'C' ? h : w) + pick(0, 3);
                    const size_t bytes = static_cast<size_t>(lead) * (ordering == 'C' ? w : h) * eb + 1;
                    d.storage.emplace_back(new char[bytes]);
                    d.extents.emplace_back(d.storage.back().get(), bytes - 1);
                    bi.push_back(i); bj.push_back(j); data.push_back(d.storage.back().get()); ld.push_back(std::max<std::int64_t>(lead, 1));
                }
        d.of_rank.push_back(costa::erased_custom_layout(nr, nc, rs.data(), cs.data(), owners.data(), static_cast<int>(bi.size()), bi.data(), bj.data(),
                                                        data.data(), ld.data(), ordering));
        d.of_rank.back().grid.n_ranks = P;
    }
    return d;
}

Dist block_cyclic(int rows, int cols, int P, int eb) {
    Dist d;
    int nprow = 1;
    for (int f = 1; f <= P; ++f)
        if (P % f == 0 && rng() % 2) nprow = f;
    const int npcol = P / nprow;
    const int mb = pick(1, 9), nb = pick(1, 9);
    // sub(A) of a larger matrix: A is (ia - 1 + rows + pad) x (ja - 1 + cols + pad)
    const int ia = pick(1, 6), ja = pick(1, 6);
    const int M = ia - 1 + rows + pick(0, 4), N = ja - 1 + cols + pick(0, 4);
    const int rsrc = pick(0, nprow - 1), csrc = pick(0, npcol - 1);
    const char order = rng() % 2 ? 'R' : 'C';
    for (int r = 0; r < P; ++r) {
        int pr, pc;
        costa::rank_to_grid(r, nprow, npcol, order, &pr, &pc);
        const int lr = costa::numroc(M, mb, pr, rsrc, nprow), lc = costa::numroc(N, nb, pc, csrc, npcol);
        const int lld = std::max(lr, 1) + pick(0, 2);
        const size_t bytes = static_cast<size_t>(lld) * std::max(lc, 1) * eb + 1;
        d.storage.emplace_back(new char[bytes]);
        d.extents.emplace_back(d.storage.back().get(), bytes - 1);
        d.of_rank.push_back(costa::erased_scalapack_layout(lld, M, N, ia, ja, rows, cols, mb, nb, nprow, npcol, order, rsrc, csrc, d.storage.back().get(), eb, 'C', r));
    }
    return d;
}

bool inside(const Dist& d, const char* p, size_t bytes) {
    for (const auto& e : d.extents)
        if (p >= e.first && p + bytes <= e.first + e.second) return true;
    return bytes == 0;
}

// bytes spanned by a piece stored with leading dimension ld (elements)
size_t span(int rows, int cols, char ordering, std::int64_t ld, int eb) {
    if (rows <= 0 || cols <= 0) return 0;
    const std::int64_t major = ordering == 'C' ? cols : rows, minor = ordering == 'C' ? rows : cols;
    return static_cast<size_t>((major - 1) * ld + minor) * eb;
}

}  // namespace

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? static_cast<unsigned>(std::atoi(argv[1])) : 1u;
    const int N = argc > 2 ? std::atoi(argv[2]) : 200;
    rng.seed(seed);
    long long plans = 0, pieces = 0, bad = 0, refused = 0;
    for (int it = 0; it < N; ++it) {
        const int P = pick(1, 9), eb = 4 << (rng() % 3);
        const int rows = pick(1, 60), cols = pick(1, 60);
        const char op = "NTC"[rng() % 3];
        // the source holds op's argument: (cols x rows) when transposed
        const int srows = op == 'N' ? rows : cols, scols = op == 'N' ? cols : rows;
        Dist from = rng() % 2 ? custom(srows, scols, P, eb) : block_cyclic(srows, scols, P, eb);
        Dist to = rng() % 2 ? custom(rows, cols, P, eb) : block_cyclic(rows, cols, P, eb);
        std::vector<costa::transform_plan> plan(P);
        try {
            for (int r = 0; r < P; ++r) {
                costa::transform_spec sp;
                sp.from = &from.of_rank[r]; sp.to = &to.of_rank[r]; sp.op = op;
                sp.alpha[0] = 2.0; sp.beta[0] = rng() % 2 ? 0.0 : -1.0;
                plan[r] = costa::plan_transform({sp}, r, P, eb);
                ++plans;
            }
        } catch (const std::exception& e) {
            ++refused;
            continue;
        }
        std::int64_t moved = 0;
        for (int r = 0; r < P; ++r) {
            const auto& pl = plan[r];
            for (const auto& pc : pl.pack) {
                ++pieces;
                const size_t tight = static_cast<size_t>(pc.n_rows) * pc.n_cols * eb;
                if (!inside(from, static_cast<const char*>(pc.src), span(pc.n_rows, pc.n_cols, pc.src_ordering, pc.src_ld, eb))) { ++bad; std::printf("pack piece outside its source block\n"); }
                if (reinterpret_cast<std::int64_t>(pc.dst) < 0 || reinterpret_cast<std::int64_t>(pc.dst) + static_cast<std::int64_t>(tight) > pl.total_send) { ++bad; std::printf("pack piece outside the send buffer\n"); }
            }
            for (const auto& pc : pl.local) {
                ++pieces;
                const int dr = pc.transpose ? pc.n_cols : pc.n_rows, dc = pc.transpose ? pc.n_rows : pc.n_cols;
                if (!pc.scale_only && !inside(from, static_cast<const char*>(pc.src), span(pc.n_rows, pc.n_cols, pc.src_ordering, pc.src_ld, eb))) { ++bad; std::printf("local piece outside its source block\n"); }
                if (!inside(to, static_cast<const char*>(pc.dst), span(dr, dc, pc.dst_ordering, pc.dst_ld, eb))) { ++bad; std::printf("local piece outside its target block\n"); }
                if (!pc.scale_only) moved += static_cast<std::int64_t>(pc.n_rows) * pc.n_cols;
            }
            for (const auto& pc : pl.unpack) {
                ++pieces;
                const int dr = pc.transpose ? pc.n_cols : pc.n_rows, dc = pc.transpose ? pc.n_rows : pc.n_cols;
                const size_t tight = static_cast<size_t>(pc.n_rows) * pc.n_cols * eb;
                if (reinterpret_cast<std::int64_t>(pc.src) < 0 || reinterpret_cast<std::int64_t>(pc.src) + static_cast<std::int64_t>(tight) > pl.total_recv) { ++bad; std::printf("unpack piece outside the receive buffer\n"); }
                if (!inside(to, static_cast<const char*>(pc.dst), span(dr, dc, pc.dst_ordering, pc.dst_ld, eb))) { ++bad; std::printf("unpack piece outside its target block\n"); }
                moved += static_cast<std::int64_t>(pc.n_rows) * pc.n_cols;
            }
            for (int p = 0; p < P; ++p)
                if (pl.send_bytes[p] != plan[p].recv_bytes[r]) { ++bad; std::printf("rank %d sends %lld bytes to %d, which expects %lld\n", r, (long long)pl.send_bytes[p], p, (long long)plan[p].recv_bytes[r]); }
        }
        if (moved != static_cast<std::int64_t>(rows) * cols) { ++bad; std::printf("it %d: %lld elements moved, matrix has %d\n", it, (long long)moved, rows * cols); }
    }
    std::printf("fuzz_transform_planner seed %u: %d problems, %lld plans, %lld pieces checked, %lld refused, %lld problems\n", seed, N, plans, pieces, refused, bad);
    return bad ? 1 : 0;
}
