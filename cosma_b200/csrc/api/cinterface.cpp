// {s,d,c,z}multiply_using_layout of the C / Fortran interface (reference src/cosma/cinterface.cpp:10-150): the caller's
// `layout` structs go to the C ABI unchanged in meaning. As in the reference, errors are not caught at this boundary.
#include <cosma/b200_runtime.hpp>
#include <cosma/cinterface.hpp>

#include <vector>

namespace {

struct converted {
    std::vector<cosma_b200_block> blocks;
    cosma_b200_layout c;
    explicit converted(const layout* l) {
        if (!l) throw std::runtime_error("multiply_using_layout: null layout");
        blocks.reserve(l->nlocalblocks);
        for (int b = 0; b < l->nlocalblocks; ++b)
            blocks.push_back(cosma_b200_block{l->localblocks[b].data, l->localblocks[b].ld, l->localblocks[b].row, l->localblocks[b].col});
        c.rowblocks = l->rowblocks;
        c.colblocks = l->colblocks;
        c.rowsplit = l->rowsplit;
        c.colsplit = l->colsplit;
        c.owners = l->owners;
        c.nlocalblocks = l->nlocalblocks;
        c.localblocks = blocks.data();
    }
};

typedef int (*entry_t)(void*, const char*, const char*, const double*, const cosma_b200_layout*, const cosma_b200_layout*, const double*,
                       const cosma_b200_layout*, void*);

template <typename Real>
void run(entry_t entry, bool cplx, MPI_Comm comm, const char* transa, const char* transb, const Real* alpha, const layout* A, const layout* B,
         const Real* beta, const layout* C) {
    void* handle = cosma::b200::comm_handle(comm);
    converted a(A), b(B), c(C);
    const double a2[2] = {static_cast<double>(alpha[0]), cplx ? static_cast<double>(alpha[1]) : 0.0};
    const double b2[2] = {static_cast<double>(beta[0]), cplx ? static_cast<double>(beta[1]) : 0.0};
    const int st = entry(handle, transa, transb, a2, &a.c, &b.c, b2, &c.c, nullptr);
    const int sy = cosma_b200_stream_synchronize(nullptr);
    cosma::b200::check(st, "multiply_using_layout");
    cosma::b200::check(sy, "multiply_using_layout (synchronize)");
}

}  // namespace

extern "C" {
void smultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const float* alpha, const layout* A, const layout* B,
                            const float* beta, const layout* C) {
    run<float>(cosma_b200_smultiply_using_layout, false, comm, transa, transb, alpha, A, B, beta, C);
}
void dmultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const double* alpha, const layout* A, const layout* B,
                            const double* beta, const layout* C) {
    run<double>(cosma_b200_dmultiply_using_layout, false, comm, transa, transb, alpha, A, B, beta, C);
}
void cmultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const float* alpha, const layout* A, const layout* B,
                            const float* beta, const layout* C) {
    run<float>(cosma_b200_cmultiply_using_layout, true, comm, transa, transb, alpha, A, B, beta, C);
}
void zmultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const double* alpha, const layout* A, const layout* B,
                            const double* beta, const layout* C) {
    run<double>(cosma_b200_zmultiply_using_layout, true, comm, transa, transb, alpha, A, B, beta, C);
}
}
