// libcosma_blacs_lite.so -- process grids for boxes without ScaLAPACK/BLACS (this image has neither; the reference gets
// them from MKL / Cray LibSci / netlib-scalapack, CMakeLists.txt:21). Only what cosma::pxgemm and its miniapp / tests call:
// grid creation and queries (reference call sites: src/cosma/blacs.hpp:5-35, scalapack.cpp:3-46, cosma_pxgemm.cpp:57-69,
// utils/pxgemm_utils.hpp:100-189) plus descinit_ and numroc_. A real BLACS, when linked, simply takes precedence: every
// symbol here is weak.
#include <cosma/blacs.hpp>
#include <costa/erased_layout.hpp>

#include <cctype>
#include <mutex>
#include <vector>

namespace {
struct grid_t {
    bool live = false;
    int sys = 0;
    char order = 'R';
    int nprow = 0, npcol = 0;
};
std::mutex g_mu;
std::vector<MPI_Comm>& systems() {
    static std::vector<MPI_Comm> s{MPI_COMM_WORLD};  // system context 0 = the world
    return s;
}
std::vector<grid_t>& grids() {
    static std::vector<grid_t> g;
    return g;
}
const grid_t* grid(int ictxt) {
    auto& g = grids();
    return (ictxt >= 0 && ictxt < static_cast<int>(g.size()) && g[ictxt].live) ? &g[ictxt] : nullptr;
}
}  // namespace

#define COSMA_B200_WEAK __attribute__((weak, visibility("default")))

extern "C" {

COSMA_B200_WEAK void Cblacs_pinfo(int* mypnum, int* nprocs) {
    MPI_Comm_rank(MPI_COMM_WORLD, mypnum);
    MPI_Comm_size(MPI_COMM_WORLD, nprocs);
}

COSMA_B200_WEAK void Cblacs_get(int ictxt, int what, int* val) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (what == 10) {  // system context of a grid
        const grid_t* g = grid(ictxt);
        *val = g ? g->sys : 0;
    } else {
        *val = 0;  // what == 0: the default system context
    }
}

COSMA_B200_WEAK int Csys2blacs_handle(MPI_Comm comm) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto& s = systems();
    for (size_t i = 0; i < s.size(); ++i)
        if (s[i] == comm) return static_cast<int>(i);
    s.push_back(comm);
    return static_cast<int>(s.size()) - 1;
}

COSMA_B200_WEAK MPI_Comm Cblacs2sys_handle(int ictxt) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto& s = systems();
    return (ictxt >= 0 && ictxt < static_cast<int>(s.size())) ? s[ictxt] : MPI_COMM_NULL;
}

COSMA_B200_WEAK void Cfree_blacs_system_handle(int) {}

// in: *ictxt = system context; out: the grid context
COSMA_B200_WEAK void Cblacs_gridinit(int* ictxt, char* order, int nprow, int npcol) {
    std::lock_guard<std::mutex> lock(g_mu);
    grid_t g;
    g.live = true;
    g.sys = (*ictxt >= 0 && *ictxt < static_cast<int>(systems().size())) ? *ictxt : 0;
    g.order = (order && std::toupper(*order) == 'C') ? 'C' : 'R';
    g.nprow = nprow;
    g.npcol = npcol;
    grids().push_back(g);
    *ictxt = static_cast<int>(grids().size()) - 1;
}

COSMA_B200_WEAK void Cblacs_gridinfo(int ictxt, int* nprow, int* npcol, int* myrow, int* mycol) {
    std::lock_guard<std::mutex> lock(g_mu);
    const grid_t* g = grid(ictxt);
    if (!g) { *nprow = *npcol = *myrow = *mycol = -1; return; }
    int rank = 0;
    MPI_Comm_rank(systems()[g->sys], &rank);
    *nprow = g->nprow;
    *npcol = g->npcol;
    if (rank >= g->nprow * g->npcol) { *myrow = *mycol = -1; return; }
    costa::rank_to_grid(rank, g->nprow, g->npcol, g->order, myrow, mycol);
}

COSMA_B200_WEAK int Cblacs_pnum(int ictxt, int prow, int pcol) {
    std::lock_guard<std::mutex> lock(g_mu);
    const grid_t* g = grid(ictxt);
    return g ? costa::rank_from_grid(prow, pcol, g->nprow, g->npcol, g->order) : -1;
}

COSMA_B200_WEAK void Cblacs_pcoord(int ictxt, int nodenum, int* prow, int* pcol) {
    std::lock_guard<std::mutex> lock(g_mu);
    const grid_t* g = grid(ictxt);
    if (!g || nodenum < 0 || nodenum >= g->nprow * g->npcol) { *prow = *pcol = -1; return; }
    costa::rank_to_grid(nodenum, g->nprow, g->npcol, g->order, prow, pcol);
}

COSMA_B200_WEAK void Cblacs_barrier(int ictxt, char*) {
    MPI_Comm comm;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        const grid_t* g = grid(ictxt);
        comm = systems()[g ? g->sys : 0];
    }
    MPI_Barrier(comm);
}

COSMA_B200_WEAK void Cblacs_gridexit(int ictxt) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (grid(ictxt)) grids()[ictxt].live = false;
}

COSMA_B200_WEAK void Cblacs_exit(int) {}

// The Fortran entry points of the same calls (all arguments by reference, trailing underscore; character arguments carry a hidden
// length that is not read) -- what a Fortran ScaLAPACK application calls before p?gemm_ when this library stands in for BLACS.
COSMA_B200_WEAK void blacs_pinfo_(int* mypnum, int* nprocs) { Cblacs_pinfo(mypnum, nprocs); }
COSMA_B200_WEAK void blacs_get_(const int* ictxt, const int* what, int* val) { Cblacs_get(*ictxt, *what, val); }
COSMA_B200_WEAK void blacs_gridinit_(int* ictxt, const char* order, const int* nprow, const int* npcol) {
    char o = order ? *order : 'R';
    Cblacs_gridinit(ictxt, &o, *nprow, *npcol);
}
COSMA_B200_WEAK void blacs_gridinfo_(const int* ictxt, int* nprow, int* npcol, int* myrow, int* mycol) { Cblacs_gridinfo(*ictxt, nprow, npcol, myrow, mycol); }
COSMA_B200_WEAK int blacs_pnum_(const int* ictxt, const int* prow, const int* pcol) { return Cblacs_pnum(*ictxt, *prow, *pcol); }
COSMA_B200_WEAK void blacs_pcoord_(const int* ictxt, const int* pnum, int* prow, int* pcol) { Cblacs_pcoord(*ictxt, *pnum, prow, pcol); }
COSMA_B200_WEAK void blacs_barrier_(const int* ictxt, const char* scope) {
    char sc = scope ? *scope : 'A';
    Cblacs_barrier(*ictxt, &sc);
}
COSMA_B200_WEAK void blacs_gridexit_(const int* ictxt) { Cblacs_gridexit(*ictxt); }
COSMA_B200_WEAK void blacs_exit_(const int* cont) { Cblacs_exit(*cont); }

// ScaLAPACK tools
COSMA_B200_WEAK void descinit_(int* desc, const int* m, const int* n, const int* mb, const int* nb, const int* irsrc, const int* icsrc,
                               const int* ictxt, const int* lld, int* info) {
    desc[0] = 1;  // dense matrix
    desc[1] = *ictxt;
    desc[2] = *m; desc[3] = *n; desc[4] = *mb; desc[5] = *nb; desc[6] = *irsrc; desc[7] = *icsrc; desc[8] = *lld;
    if (info) *info = (*m < 0 || *n < 0 || *mb < 1 || *nb < 1 || *lld < 1) ? -1 : 0;
}
COSMA_B200_WEAK int numroc_(const int* n, const int* nb, const int* iproc, const int* isrcproc, const int* nprocs) {
    return costa::numroc(*n, *nb, *iproc, *isrcproc, *nprocs);
}
}
