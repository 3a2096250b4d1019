// Minimal test harness for the C++ API tests (the image has no gtest): CHECK macros, a per-process failure count and a
// summary that is reduced over the ranks so that every rank exits with the same code.
#pragma once
#include <cosma/mpi_compat.hpp>

#include <cstdio>
#include <iostream>
#include <string>

namespace check {
inline int& failures() { static int f = 0; return f; }
inline int& passed() { static int p = 0; return p; }
inline int& skipped() { static int s = 0; return s; }

inline int finish(const char* suite) {
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    int local[3] = {failures(), passed(), skipped()}, total[3] = {0, 0, 0};
    MPI_Allreduce(local, total, 3, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    if (rank == 0) std::printf("[%s] checks passed (all ranks) = %d, failed = %d, cases skipped on rank 0 = %d\n", suite, total[1], total[0], local[2]);
    std::fflush(stdout);
    return total[0] == 0 ? 0 : 1;
}
}  // namespace check

#define CHECK_TRUE(cond)                                                                          \
    do {                                                                                          \
        if (cond) { ++check::passed(); }                                                          \
        else { ++check::failures(); std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::fflush(stdout); } \
    } while (0)
#define CHECK_MSG(cond, msg)                                                                      \
    do {                                                                                          \
        if (cond) { ++check::passed(); }                                                          \
        else { ++check::failures(); std::cout << "CHECK FAILED " << __FILE__ << ":" << __LINE__ << ": " << #cond << " -- " << msg << std::endl; } \
    } while (0)
