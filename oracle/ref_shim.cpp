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
