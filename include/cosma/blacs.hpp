// The BLACS entry points cosma::pxgemm needs (reference src/cosma/blacs.hpp:5-35). In a ScaLAPACK application they come
// from the ScaLAPACK library the application already links; on a box without one, libcosma_blacs_lite.so (csrc/api/
// blacs_lite.cpp) provides process grids over the process group.
#pragma once
#include <cosma/mpi_compat.hpp>

namespace cosma {
namespace blacs {
extern "C" {
void Cblacs_pinfo(int* mypnum, int* nprocs);
void Cblacs_get(int ictxt, int what, int* val);
void Cblacs_gridinit(int* ictxt, char* order, int nprow, int npcol);
void Cblacs_gridexit(int ictxt);
void Cblacs_exit(int NotDone);
void Cblacs_gridinfo(int ictxt, int* nprow, int* npcol, int* myrow, int* mycol);
int Cblacs_pnum(int ictxt, int prow, int pcol);
void Cblacs_pcoord(int ictxt, int nodenum, int* prow, int* pcol);
void Cblacs_barrier(int ictxt, char* scope);
MPI_Comm Cblacs2sys_handle(int ictxt);
int Csys2blacs_handle(MPI_Comm mpi_comm);
void Cfree_blacs_system_handle(int i_sys_ctxt);
}
}  // namespace blacs
}  // namespace cosma
