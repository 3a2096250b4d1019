// TEST INFRASTRUCTURE ONLY -- the handful of BLACS entry points the reference's ScaLAPACK wrappers call
// (src/cosma/blacs.hpp:5-35, libs/COSTA/src/costa/blacs.hpp; call sites cosma_pxgemm.cpp:57-69, scalapack.cpp:3-52,
// costa_pxgemr2d.cpp:40-60, costa_pxtran_op.cpp:45-56), over the minimpi stand-in. There is no BLACS/ScaLAPACK in the
// image; this lets the UNMODIFIED cosma::pxgemm / costa::pxgemr2d / costa::pxtran_op run here as the oracle of
// our p?gemm / p?gemr2d / p?tran entry points. Process grids always span MPI_COMM_WORLD ranks 0..nprow*npcol-1.
#include "mpi.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {
struct Grid {
    bool live = false;
    char order = 'R';
    int nprow = 0, npcol = 0;
};
std::vector<Grid> g_grids;

Grid& grid(int ictxt) {
    if (ictxt < 0 || ictxt >= (int)g_grids.size() || !g_grids[ictxt].live) {
        std::fprintf(stderr, "miniblacs: invalid context %d\n", ictxt);
        std::abort();
    }
    return g_grids[ictxt];
}
void coords(const Grid& g, int pnum, int* prow, int* pcol) {
    if (pnum < 0 || pnum >= g.nprow * g.npcol) {
        *prow = *pcol = -1;
    } else if (g.order == 'C' || g.order == 'c') {
        *prow = pnum % g.nprow;
        *pcol = pnum / g.nprow;
    } else {
        *prow = pnum / g.npcol;
        *pcol = pnum % g.npcol;
    }
}
}  // namespace

extern "C" {
void Cblacs_pinfo(int* mypnum, int* nprocs) {
    MPI_Comm_rank(MPI_COMM_WORLD, mypnum);
    MPI_Comm_size(MPI_COMM_WORLD, nprocs);
}
void Cblacs_setup(int* mypnum, int* nprocs) { Cblacs_pinfo(mypnum, nprocs); }
void Cblacs_set(int, int, int*) {}
// what = 0: default system context; what = 10: the system context a grid was built on. Both are handle 0 = WORLD.
void Cblacs_get(int, int, int* val) { *val = 0; }
void Cblacs_gridinit(int* ictxt, char* order, int nprow, int npcol) {
    Grid g;
    g.live = true;
    g.order = *order;
    g.nprow = nprow;
    g.npcol = npcol;
    g_grids.push_back(g);
    *ictxt = (int)g_grids.size() - 1;
}
void Cblacs_gridmap(int*, int*, int, int, int) { std::fprintf(stderr, "miniblacs: Cblacs_gridmap unsupported\n"); std::abort(); }
void Cblacs_freebuff(int, int) {}
void Cblacs_gridexit(int ictxt) { grid(ictxt).live = false; }
void Cblacs_exit(int) {}
void Cblacs_abort(int, int err) { MPI_Abort(MPI_COMM_WORLD, err); }
void Cblacs_gridinfo(int ictxt, int* nprow, int* npcol, int* myrow, int* mycol) {
    const Grid& g = grid(ictxt);
    int me;
    MPI_Comm_rank(MPI_COMM_WORLD, &me);
    *nprow = g.nprow;
    *npcol = g.npcol;
    coords(g, me, myrow, mycol);
}
int Cblacs_pnum(int ictxt, int prow, int pcol) {
    const Grid& g = grid(ictxt);
    return (g.order == 'C' || g.order == 'c') ? pcol * g.nprow + prow : prow * g.npcol + pcol;
}
void Cblacs_pcoord(int ictxt, int nodenum, int* prow, int* pcol) { coords(grid(ictxt), nodenum, prow, pcol); }
void Cblacs_barrier(int, char*) { MPI_Barrier(MPI_COMM_WORLD); }
MPI_Comm Cblacs2sys_handle(int) { return MPI_COMM_WORLD; }
int Csys2blacs_handle(MPI_Comm) { return 0; }
void Cfree_blacs_system_handle(int) {}
}
