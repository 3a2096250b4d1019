// Shared helpers of the C++ API tests. test_cosma<T>() has the role of the reference's helper of the same name
// (utils/cosma_utils.hpp:80-420): fill A, B, C on every rank, assemble the global matrices on rank 0 through
// global_coordinates(), multiply with the library, assemble global C again and compare it element-wise with the naive
// triple loop (oracle/gemm_oracle.c, the restatement of the reference's own local_multiply_cpu) under the reference's
// criterion (relative error < epsilon, 1e-5 for 4-byte reals; cosma_utils.hpp:366-377).
#pragma once
#include "check.hpp"

#include <cosma/multiply.hpp>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

extern "C" {
void oracle_dgemm(char, char, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t);
void oracle_sgemm(char, char, int64_t, int64_t, int64_t, float, const float*, int64_t, const float*, int64_t, float, float*, int64_t);
void oracle_zgemm(char, char, int64_t, int64_t, int64_t, const double*, const double*, int64_t, const double*, int64_t, const double*, double*, int64_t);
void oracle_cgemm(char, char, int64_t, int64_t, int64_t, const float*, const float*, int64_t, const float*, int64_t, const float*, float*, int64_t);
}

namespace testutil {

template <typename T> struct mpi_type;
template <> struct mpi_type<float> { static MPI_Datatype get() { return MPI_FLOAT; } };
template <> struct mpi_type<double> { static MPI_Datatype get() { return MPI_DOUBLE; } };
template <> struct mpi_type<std::complex<float>> { static MPI_Datatype get() { return MPI_C_FLOAT_COMPLEX; } };
template <> struct mpi_type<std::complex<double>> { static MPI_Datatype get() { return MPI_C_DOUBLE_COMPLEX; } };

template <typename T> struct real_of { using type = T; };
template <typename T> struct real_of<std::complex<T>> { using type = T; };

inline void naive_gemm(char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc) {
    oracle_dgemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
inline void naive_gemm(char ta, char tb, int m, int n, int k, float alpha, const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc) {
    oracle_sgemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
inline void naive_gemm(char ta, char tb, int m, int n, int k, std::complex<double> alpha, const std::complex<double>* A, int lda,
                       const std::complex<double>* B, int ldb, std::complex<double> beta, std::complex<double>* C, int ldc) {
    oracle_zgemm(ta, tb, m, n, k, reinterpret_cast<const double*>(&alpha), reinterpret_cast<const double*>(A), lda, reinterpret_cast<const double*>(B), ldb,
                 reinterpret_cast<const double*>(&beta), reinterpret_cast<double*>(C), ldc);
}
inline void naive_gemm(char ta, char tb, int m, int n, int k, std::complex<float> alpha, const std::complex<float>* A, int lda,
                       const std::complex<float>* B, int ldb, std::complex<float> beta, std::complex<float>* C, int ldc) {
    oracle_cgemm(ta, tb, m, n, k, reinterpret_cast<const float*>(&alpha), reinterpret_cast<const float*>(A), lda, reinterpret_cast<const float*>(B), ldb,
                 reinterpret_cast<const float*>(&beta), reinterpret_cast<float*>(C), ldc);
}

// entries in [1, 10), positive like the reference's fill (utils/cosma_utils.hpp:14-34), so that its element-wise RELATIVE
// criterion is meaningful (no cancellation in the dot products)
template <typename T>
inline T make_value(std::mt19937& gen) {
    std::uniform_real_distribution<double> d(1.0, 10.0);
    return static_cast<T>(d(gen));
}
template <>
inline std::complex<double> make_value<std::complex<double>>(std::mt19937& gen) {
    std::uniform_real_distribution<double> d(1.0, 10.0);
    const double re = d(gen);
    return {re, d(gen)};
}
template <>
inline std::complex<float> make_value<std::complex<float>>(std::mt19937& gen) {
    std::uniform_real_distribution<float> d(1.0f, 10.0f);
    const float re = d(gen);
    return {re, d(gen)};
}

// the first new_P ranks of comm (the reference builds it with MPI groups, tests/multiply.cpp:7-36)
inline MPI_Comm subcommunicator(int new_P, MPI_Comm comm = MPI_COMM_WORLD) {
    int P = 0;
    MPI_Comm_size(comm, &P);
    MPI_Group all, kept;
    MPI_Comm_group(comm, &all);
    std::vector<int> excluded;
    for (int i = new_P; i < P; ++i) excluded.push_back(i);
    MPI_Group_excl(all, static_cast<int>(excluded.size()), excluded.data(), &kept);
    MPI_Comm out = MPI_COMM_NULL;
    MPI_Comm_create_group(comm, kept, 0, &out);  // MPI_COMM_NULL on the excluded ranks
    MPI_Group_free(&all);
    MPI_Group_free(&kept);
    return out;
}

template <typename T>
void fill_matrix(cosma::CosmaMatrix<T>& M, unsigned seed) {
    std::mt19937 gen(seed);
    for (size_t i = 0; i < M.matrix_size(); ++i) M.matrix_pointer()[i] = make_value<T>(gen);
}

// rank 0 of comm receives every rank's local matrix and scatters it into a column-major global matrix
template <typename T>
std::vector<T> gather_global(cosma::CosmaMatrix<T>& M, int P, MPI_Comm comm, int tag) {
    int rank = 0;
    MPI_Comm_rank(comm, &rank);
    std::vector<T> global;
    if (rank == 0) {
        global.assign(static_cast<size_t>(M.m()) * M.n(), T{0});
        for (int r = 0; r < P; ++r) {
            const size_t sz = M.matrix_size(r);
            std::vector<T> part(sz);
            if (r == 0) std::memcpy(part.data(), M.matrix_pointer(), sz * sizeof(T));
            else MPI_Recv(part.data(), static_cast<int>(sz), mpi_type<T>::get(), r, tag, comm, MPI_STATUS_IGNORE);
            for (size_t j = 0; j < sz; ++j) {
                int gi, gj;
                std::tie(gi, gj) = M.global_coordinates(static_cast<int>(j), r);
                if (gi >= 0 && gj >= 0) global.at(static_cast<size_t>(gj) * M.m() + gi) = part[j];
            }
        }
    } else if (rank < P) {
        MPI_Ssend(M.matrix_pointer(), static_cast<int>(M.matrix_size()), mpi_type<T>::get(), 0, tag, comm);
    }
    return global;
}

template <typename T>
bool close_enough(const T& got, const T& want, double epsilon) {
    const double err = std::abs(got - want), scale = std::max(std::abs(got), std::abs(want));
    const double rel = scale > 1e-10 ? err / scale : err;
    const double tol = sizeof(typename real_of<T>::type) == 4 ? 1e-5 : epsilon;
    return rel < tol;
}

// one distributed multiply checked against the naive GEMM on rank 0; every rank of comm returns the verdict
template <typename T>
bool test_cosma(const cosma::Strategy& s, cosma::context<T>& ctx, MPI_Comm comm, double epsilon = 1e-8, int tag = 0, T alpha = T{1}, T beta = T{1}) {
    int rank = 0;
    MPI_Comm_rank(comm, &rank);
    const int m = s.m, n = s.n, k = s.k, P = static_cast<int>(s.P);
    cosma::CosmaMatrix<T> A(ctx, 'A', s, rank), B(ctx, 'B', s, rank), C(ctx, 'C', s, rank);
    fill_matrix(A, 100 + rank);
    fill_matrix(B, 200 + rank);
    fill_matrix(C, 300 + rank);
    std::vector<T> gA = gather_global(A, P, comm, 5 * tag), gB = gather_global(B, P, comm, 5 * tag + 1), want = gather_global(C, P, comm, 5 * tag + 2);
    if (rank == 0) naive_gemm('N', 'N', m, n, k, alpha, gA.data(), m, gB.data(), k, beta, want.data(), m);

    cosma::multiply(A, B, C, s, comm, alpha, beta);

    std::vector<T> got = gather_global(C, P, comm, 5 * tag + 3);
    int ok = 1;
    if (rank == 0) {
        ok = got.size() == want.size();
        int shown = 0;
        for (size_t i = 0; ok && i < got.size(); ++i) {
            if (!close_enough(got[i], want[i], epsilon)) {
                if (shown++ < 5) {
                    int li, lr;
                    std::tie(li, lr) = C.local_coordinates(static_cast<int>(i % m), static_cast<int>(i / m));
                    std::cout << "global(" << i % m << ", " << i / m << ") = (loc " << li << ", rank " << lr << ") = " << got[i] << " and should be " << want[i] << std::endl;
                }
                ok = 0;
            }
        }
    }
    MPI_Bcast(&ok, 1, MPI_INT, 0, comm);
    return ok != 0;
}

}  // namespace testutil
