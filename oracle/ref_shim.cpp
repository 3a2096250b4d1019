// TEST INFRASTRUCTURE ONLY -- extern "C" accessors over the UNMODIFIED reference classes, compiled
// together with the reference sources into oracle/_ref/libcosma_ref.so (see oracle/Makefile).
// Used by tests (to pin our restatement against the real reference) and by bench.py --impl reference.
#include <cosma/strategy.hpp>
#include <cosma/mapper.hpp>
#include <cosma/blas.hpp>
#include <cblas.h>

#include <cstring>
#include <string>
#include <vector>

extern "C" {

// Strategy(m,n,k,P,mem_limit) -> "pm2,sn4,..." written to out; returns #steps or -1 on exception.
// strat != "" : use it as (possibly incomplete) prefix, as utils/parse_strategy.hpp does.
int ref_strategy(int m, int n, int k, int P, long long mem_limit, const char* prefix, char* out, int out_len,
                 int* P_out, long long* mem_used) {
    try {
        std::vector<int> divs;
        std::string dims, types;
        std::string s(prefix ? prefix : "");
        size_t pos = 0;
        while (pos < s.size()) {
            size_t comma = s.find(',', pos);
            if (comma == std::string::npos) comma = s.size();
            std::string step = s.substr(pos, comma - pos);
            if (step.size() >= 3) {
                types += step[0];
                dims += step[1];
                divs.push_back(std::stoi(step.substr(2)));
            }
            pos = comma + 1;
        }
        if (mem_limit <= 0) mem_limit = std::numeric_limits<long long>::max();
        cosma::Strategy st = divs.empty() ? cosma::Strategy(m, n, k, P, mem_limit)
                                          : cosma::Strategy(m, n, k, P, divs, dims, types, mem_limit);
        std::string res;
        for (size_t i = 0; i < st.n_steps(); ++i) {
            if (i) res += ",";
            res += st.step_type[i];
            res += st.split_dimension[i];
            res += std::to_string(st.divisors[i]);
        }
        if ((int)res.size() + 1 > out_len) return -2;
        std::strcpy(out, res.c_str());
        if (P_out) *P_out = (int)st.P;
        if (mem_used) *mem_used = st.memory_used;
        return (int)st.n_steps();
    } catch (...) {
        return -1;
    }
}

static cosma::Strategy make_strategy(int m, int n, int k, int P, const char* steps) {
    std::vector<int> divs;
    std::string dims, types;
    std::string s(steps ? steps : "");
    size_t pos = 0;
    while (pos < s.size()) {
        size_t comma = s.find(',', pos);
        if (comma == std::string::npos) comma = s.size();
        std::string step = s.substr(pos, comma - pos);
        if (step.size() >= 3) {
            types += step[0];
            dims += step[1];
            divs.push_back(std::stoi(step.substr(2)));
        }
        pos = comma + 1;
    }
    if (divs.empty()) return cosma::Strategy(m, n, k, P);
    return cosma::Strategy(m, n, k, P, divs, dims, types);
}

// Mapper(label, strategy, rank=0).complete_layout(): for every rank, its list of blocks
// (row_start,row_end,col_start,col_end inclusive). out = flat int array [rank][block][4];
// counts[rank] = number of blocks. Returns total number of blocks or -1.
int ref_mapper_layout(char label, int m, int n, int k, int P, const char* steps, int* counts, int* out, int out_cap) {
    try {
        cosma::Strategy st = make_strategy(m, n, k, P, steps);
        cosma::Mapper mapper(label, st, 0);
        auto& layout = mapper.complete_layout();
        int total = 0;
        for (size_t r = 0; r < layout.size(); ++r) {
            counts[r] = (int)layout[r].size();
            for (auto& b : layout[r]) {
                if (4 * (total + 1) > out_cap) return -2;
                out[4 * total + 0] = b.rows.first();
                out[4 * total + 1] = b.rows.last();
                out[4 * total + 2] = b.cols.first();
                out[4 * total + 3] = b.cols.last();
                ++total;
            }
        }
        return total;
    } catch (...) {
        return -1;
    }
}

// global <-> local coordinate tables of the reference Mapper
int ref_mapper_local_coordinates(char label, int m, int n, int k, int P, const char* steps, int gi, int gj, int* local_idx,
                                 int* rank) {
    try {
        cosma::Strategy st = make_strategy(m, n, k, P, steps);
        cosma::Mapper mapper(label, st, 0);
        auto res = mapper.local_coordinates(gi, gj);
        *local_idx = res.first;
        *rank = res.second;
        return 0;
    } catch (...) {
        return -1;
    }
}

// the reference's CPU base-case GEMM (src/cosma/blas.cpp -> cblas_dgemm/zgemm of OpenBLAS)
void ref_dgemm(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
               double* C, int ldc) {
    cosma::gemm(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
void ref_zgemm(int m, int n, int k, const double* alpha, const double* A, int lda, const double* B, int ldb,
               const double* beta, double* C, int ldc) {
    cosma::gemm(m, n, k, std::complex<double>(alpha[0], alpha[1]), reinterpret_cast<const std::complex<double>*>(A), lda,
                reinterpret_cast<const std::complex<double>*>(B), ldb, std::complex<double>(beta[0], beta[1]),
                reinterpret_cast<std::complex<double>*>(C), ldc);
}

}  // extern "C"

extern "C" void ref_set_blas_threads(int n) { openblas_set_num_threads(n); }

// ---- timing the reference's own cosma::multiply at P = 1 (bench.py --impl reference / cpu_baseline) ----
// Mirrors run<T>() of the reference miniapp (miniapp/cosma_miniapp.cpp:35-81): Strategy(m,n,k,1), CosmaMatrix
// A/B/C in the singleton context, srand48(rank), 10*drand48() fill, alpha = 1, beta = 0, steady_clock around
// multiply(). Differences: per-repetition times are returned unsorted so warm-up can be dropped, and the
// matrices are built once.
#include <cosma/multiply.hpp>
#include <chrono>
#include <cstdlib>

extern "C" int ref_multiply_time_d(int m, int n, int k, int reps, double* times_ms, double* checksum) {
    try {
        using namespace cosma;
        Strategy strategy(m, n, k, 1);
        CosmaMatrix<double> A('A', strategy, 0);
        CosmaMatrix<double> B('B', strategy, 0);
        CosmaMatrix<double> C('C', strategy, 0);
        srand48(0);
        for (size_t i = 0; i < A.matrix_size(); ++i) A.matrix_pointer()[i] = 10 * drand48();
        for (size_t i = 0; i < B.matrix_size(); ++i) B.matrix_pointer()[i] = 10 * drand48();
        for (int r = 0; r < reps; ++r) {
            auto start = std::chrono::steady_clock::now();
            multiply(A, B, C, strategy, MPI_COMM_WORLD, 1.0, 0.0);
            auto end = std::chrono::steady_clock::now();
            times_ms[r] = std::chrono::duration<double, std::milli>(end - start).count();
        }
        if (checksum) {
            double s = 0;
            for (size_t i = 0; i < C.matrix_size(); ++i) s += C.matrix_pointer()[i];
            *checksum = s;
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

// ---- COSTA (unmodified reference) ------------------------------------------------------------------
#include <costa/grid2grid/memory_utils.hpp>
#include <costa/grid2grid/scalapack_layout.hpp>
#include <costa/grid2grid/transform.hpp>
#include <costa/layout.hpp>

namespace {
template <typename T>
void cat_typed(int n_rows, int n_cols, const void* src, int src_ld, int src_cm, void* dst, int dst_ld, int dst_cm, int transpose,
               int conj, T alpha, T beta) {
    auto& workspace = *costa::memory::get_costa_context_instance<T>();
    costa::memory::copy_and_transform<T>(n_rows, n_cols, static_cast<const T*>(src), src_ld, src_cm != 0, static_cast<T*>(dst), dst_ld,
                                         dst_cm != 0, transpose != 0, conj != 0, alpha, beta, workspace);
}
}  // namespace

extern "C" {

// costa::memory::copy_and_transform (memory_utils.hpp:287-346) for dtype 'i' | 's' | 'd' | 'c' | 'z'
int ref_copy_and_transform(char dtype, int n_rows, int n_cols, const void* src, int src_ld, int src_cm, void* dst, int dst_ld, int dst_cm,
                           int transpose, int conj, const double* alpha, const double* beta) {
    try {
        switch (dtype) {
            case 'i': cat_typed<int>(n_rows, n_cols, src, src_ld, src_cm, dst, dst_ld, dst_cm, transpose, conj, (int)alpha[0], (int)beta[0]); break;
            case 's': cat_typed<float>(n_rows, n_cols, src, src_ld, src_cm, dst, dst_ld, dst_cm, transpose, conj, (float)alpha[0], (float)beta[0]); break;
            case 'd': cat_typed<double>(n_rows, n_cols, src, src_ld, src_cm, dst, dst_ld, dst_cm, transpose, conj, alpha[0], beta[0]); break;
            case 'c': cat_typed<std::complex<float>>(n_rows, n_cols, src, src_ld, src_cm, dst, dst_ld, dst_cm, transpose, conj,
                                                     std::complex<float>((float)alpha[0], (float)alpha[1]), std::complex<float>((float)beta[0], (float)beta[1])); break;
            case 'z': cat_typed<std::complex<double>>(n_rows, n_cols, src, src_ld, src_cm, dst, dst_ld, dst_cm, transpose, conj,
                                                      std::complex<double>(alpha[0], alpha[1]), std::complex<double>(beta[0], beta[1])); break;
            default: return -2;
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

// costa::get_scalapack_layout<double> (scalapack_layout.cpp:178-285) for an arbitrary `rank`: split points, owners
// (row-major) and the local blocks as (block row, block col, element offset into the local array).
int ref_scalapack_layout(int lld, int mat_rows, int mat_cols, int ia, int ja, int sub_m, int sub_n, int mb, int nb, int nprow, int npcol,
                         char grid_order, int rsrc, int csrc, char data_ordering, int rank, int* rowblocks, int* colblocks, int* rowsplit,
                         int* colsplit, int* owners, int* nlocal, int* local_row, int* local_col, long long* local_offset) {
    try {
        double* base = reinterpret_cast<double*>(std::uintptr_t(1) << 40);
        auto ord = (grid_order == 'C' || grid_order == 'c') ? costa::scalapack::ordering::column_major : costa::scalapack::ordering::row_major;
        auto l = costa::get_scalapack_layout<double>(lld, {mat_rows, mat_cols}, {ia, ja}, {sub_m, sub_n}, {mb, nb}, {nprow, npcol}, ord,
                                                     {rsrc, csrc}, base, data_ordering, rank);
        const int nr = l.grid.grid().n_rows, nc = l.grid.grid().n_cols;
        if (rowblocks) *rowblocks = nr;
        if (colblocks) *colblocks = nc;
        if (nlocal) *nlocal = l.blocks.num_blocks();
        if (rowsplit) for (int i = 0; i <= nr; ++i) rowsplit[i] = l.grid.grid().rows_split[i];
        if (colsplit) for (int j = 0; j <= nc; ++j) colsplit[j] = l.grid.grid().cols_split[j];
        if (owners) for (int i = 0; i < nr; ++i) for (int j = 0; j < nc; ++j) owners[i * nc + j] = l.grid.owner(i, j);
        for (int b = 0; b < l.blocks.num_blocks(); ++b) {
            const auto& blk = l.blocks.get_block(b);
            if (local_row) local_row[b] = blk.coordinates.row;
            if (local_col) local_col[b] = blk.coordinates.col;
            if (local_offset) local_offset[b] = blk.data - base;
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

// One rank (P = 1) run of costa::transform (transform.cpp:162-200): to = beta*to + alpha*op(from) between two custom
// block layouts over the same host memory model as the C interface. Exercises the reference's grid overlay, block
// decomposition and copy_local_blocks path. dtype 'd' | 'z'.
struct ref_block { void* data; int ld; int row; int col; };
struct ref_layout { int rowblocks, colblocks; const int* rowsplit; const int* colsplit; const int* owners; int nlocalblocks; ref_block* localblocks; };

}  // extern "C"

namespace {
template <typename T>
int transform_p1(const ref_layout* from, char from_ord, const ref_layout* to, char to_ord, char trans, T alpha, T beta) {
    auto mk = [](const ref_layout* l, char ord) {
        std::vector<costa::block_t> blocks(l->nlocalblocks);
        for (int i = 0; i < l->nlocalblocks; ++i) blocks[i] = costa::block_t{l->localblocks[i].data, l->localblocks[i].ld, l->localblocks[i].row, l->localblocks[i].col};
        return costa::custom_layout<T>(l->rowblocks, l->colblocks, l->rowsplit, l->colsplit, l->owners, l->nlocalblocks, blocks.data(), ord);
    };
    auto F = mk(from, from_ord);
    auto G = mk(to, to_ord);
    costa::transform<T>(F, G, trans, alpha, beta, MPI_COMM_WORLD);
    return 0;
}
}  // namespace

extern "C" int ref_transform_p1(char dtype, const ref_layout* from, char from_ord, const ref_layout* to, char to_ord, char trans,
                                const double* alpha, const double* beta) {
    try {
        if (dtype == 'd') return transform_p1<double>(from, from_ord, to, to_ord, trans, alpha[0], beta[0]);
        if (dtype == 'z') return transform_p1<std::complex<double>>(from, from_ord, to, to_ord, trans, std::complex<double>(alpha[0], alpha[1]),
                                                                   std::complex<double>(beta[0], beta[1]));
        return -2;
    } catch (...) {
        return -1;
    }
}

// ---- communication volume and rank relabelling of the unmodified reference ------------------------------------------
// costa::communication_volume (transform.cpp:9-44) between two custom grids and costa::optimal_reordering
// (ranks_reordering.cpp:4-61). Volumes are returned as a dense n_ranks x n_ranks matrix indexed [min(u,v)][max(u,v)].
#include <costa/grid2grid/ranks_reordering.hpp>

extern "C" int ref_comm_volume(int rb_a, int cb_a, const int* rs_a, const int* cs_a, const int* own_a, int rb_b, int cb_b, const int* rs_b,
                               const int* cs_b, const int* own_b, char trans, int n_ranks, long long* out) {
    try {
        auto ga = costa::custom_grid(rb_a, cb_a, rs_a, cs_a, own_a);
        auto gb = costa::custom_grid(rb_b, cb_b, rs_b, cs_b, own_b);
        auto vol = costa::communication_volume(ga, gb, trans);
        for (long long i = 0; i < (long long)n_ranks * n_ranks; ++i) out[i] = 0;
        for (const auto& kv : vol.volume) {
            const int u = std::min(kv.first.src, kv.first.dest), v = std::max(kv.first.src, kv.first.dest);
            out[(long long)u * n_ranks + v] += (long long)kv.second;
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

extern "C" int ref_optimal_reordering(int n_ranks, const long long* volume, int* permutation, int* reordered) {
    try {
        costa::comm_volume vol;
        for (int u = 0; u < n_ranks; ++u)
            for (int v = u; v < n_ranks; ++v)
                if (volume[(long long)u * n_ranks + v] > 0) vol.volume[costa::edge_t{u, v}] = (size_t)volume[(long long)u * n_ranks + v];
        bool re = false;
        auto perm = costa::optimal_reordering(vol, n_ranks, re);
        for (int i = 0; i < n_ranks; ++i) permutation[i] = perm[i];
        *reordered = re ? 1 : 0;
        return 0;
    } catch (...) {
        return -1;
    }
}

// ---- cosma::adapt_strategy_to_block_cyclic_grid of the unmodified reference (cosma_pxgemm.cpp:517-650) ---------------------------
#include <cosma/cosma_pxgemm.hpp>
extern "C" int ref_adapt_strategy(int m, int n, int k, int P, const int* desca, int ia, int ja, const int* descb, int ib, int jb, const int* descc, int ic,
                                  int jc, char transa, char transb, int nprow, int npcol, char order, char* out, int out_len) {
    try {
        std::vector<int> divisors;
        std::string dims, types;
        cosma::scalapack::global_matrix_size ma(desca), mb(descb), mc(descc);
        cosma::scalapack::block_size ba(desca), bb(descb), bc(descc);
        cosma::adapt_strategy_to_block_cyclic_grid(divisors, dims, types, m, n, k, P, ma, mb, mc, ba, bb, bc, ia, ja, ib, jb, ic, jc, transa, transb, nprow,
                                                   npcol, order);
        std::string s;
        for (size_t i = 0; i < divisors.size(); ++i) {
            if (i) s += ',';
            s += types[i];
            s += dims[i];
            s += std::to_string(divisors[i]);
        }
        if ((int)s.size() + 1 > out_len) return -2;
        std::strcpy(out, s.c_str());
        return 0;
    } catch (...) {
        return -1;
    }
}
