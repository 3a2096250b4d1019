// C / Fortran interface with the reference's declarations (src/cosma/cinterface.hpp:16-76): matrices described by
// `layout` structs (grid of blocks, owners row-major, the caller's local blocks column-major with leading dimension ld).
// Block data may be host or device memory. Complex alpha / beta are (re, im) pairs.
#pragma once
#include <cosma/mpi_compat.hpp>

#ifdef __cplusplus
extern "C" {
#endif

struct block {
    void* data;
    const int ld;
    const int row;
    const int col;
};

struct layout {
    int rowblocks;
    int colblocks;
    const int* rowsplit;
    const int* colsplit;
    const int* owners;
    int nlocalblocks;
    block* localblocks;
};

void smultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const float* alpha, const layout* layout_a,
                            const layout* layout_b, const float* beta, const layout* layout_c);
void dmultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const double* alpha, const layout* layout_a,
                            const layout* layout_b, const double* beta, const layout* layout_c);
void cmultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const float* alpha, const layout* layout_a,
                            const layout* layout_b, const float* beta, const layout* layout_c);
void zmultiply_using_layout(MPI_Comm comm, const char* transa, const char* transb, const double* alpha, const layout* layout_a,
                            const layout* layout_b, const double* beta, const layout* layout_c);

#ifdef __cplusplus
}
#endif
