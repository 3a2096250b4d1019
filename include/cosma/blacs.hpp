// The BLACS entry points cosma::pxgemm needs (reference src/cosma/blacs.hpp:5-35). In a ScaLAPACK application they come
// from the ScaLAPACK library the application already links; on a box without one, libcosma_blacs_lite.so (csrc/api/
// blacs_lite.cpp) provides process grids over the process group.
#pragma once
#include <cosma/mpi_compat.hpp>

namespace cosma {
namespace blacs {
extern "C" {
// who am I among how many processes
void Cblacs_pinfo(int* my_process, int* n_processes);
// what == 0: default system context; what == 10: the system context a grid context was created from
void Cblacs_get(int context, int what, int* value);
// in: a system context, out: the grid context of an nprow x npcol grid numbered row- ("R") or column-major ("C")
void Cblacs_gridinit(int* context, char* numbering, int nprow, int npcol);
void Cblacs_gridexit(int context);
void Cblacs_exit(int keep_mpi);
// shape of the grid and the caller's coordinates (-1, -1 outside the grid)
void Cblacs_gridinfo(int context, int* nprow, int* npcol, int* my_row, int* my_col);
int Cblacs_pnum(int context, int process_row, int process_col);
void Cblacs_pcoord(int context, int process, int* process_row, int* process_col);
void Cblacs_barrier(int context, char* scope);
// system context <-> communicator
MPI_Comm Cblacs2sys_handle(int system_context);
int Csys2blacs_handle(MPI_Comm communicator);
void Cfree_blacs_system_handle(int system_context);
}
}  // namespace blacs
}  // namespace cosma
