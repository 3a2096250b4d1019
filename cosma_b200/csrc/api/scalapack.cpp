// ScaLAPACK descriptor helpers (reference src/cosma/scalapack.cpp:3-140).
#include <cosma/scalapack.hpp>
#include <costa/erased_layout.hpp>

#include <stdexcept>

namespace cosma {
namespace scalapack {

costa::scalapack::ordering rank_ordering(int ctxt, int P) {
    // BLACS numbers a row-major grid 0 -> (0,0), 1 -> (0,1): look where rank 1 sits
    if (P > 1) {
        int prow = 0, pcol = 0;
        blacs::Cblacs_pcoord(ctxt, 1, &prow, &pcol);
        if (prow == 0 && pcol == 1) return costa::scalapack::ordering::row_major;
    }
    return costa::scalapack::ordering::column_major;
}

int get_grid_context(const int* desca, const int* descb, const int* descc) {
    if (desca[1] != descb[1] || descb[1] != descc[1]) throw std::runtime_error("pxgemm: A, B and C must live in the same BLACS context");
    return desca[1];
}
int get_grid_context(const int* desc) { return desc[1]; }

int get_comm_context(const int grid_context) {
    int comm_context = 0;
    blacs::Cblacs_get(grid_context, 10, &comm_context);
    return comm_context;
}
MPI_Comm get_communicator(const int grid_context) { return blacs::Cblacs2sys_handle(get_comm_context(grid_context)); }

int leading_dimension(const int* desc) { return desc[8]; }
int numroc(int n, int nb, int proc_coord, int proc_src, int n_procs) { return costa::numroc(n, nb, proc_coord, proc_src, n_procs); }

// bounds on a legal lld: the fewest / most rows any process of the grid dimension can own
int min_leading_dimension(int n, int nb, int rank_grid_dim) { return (n / nb) / rank_grid_dim * nb; }
int max_leading_dimension(int n, int nb, int rank_grid_dim) {
    const int whole = n / nb;
    return min_leading_dimension(n, nb, rank_grid_dim) + ((whole % rank_grid_dim == 0) ? n % nb : nb);
}

int local_buffer_size(const int* desc) {
    int nprow = 0, npcol = 0, myrow = 0, mycol = 0;
    blacs::Cblacs_gridinfo(desc[1], &nprow, &npcol, &myrow, &mycol);
    return desc[8] * numroc(desc[3], desc[5], mycol, desc[7], npcol);
}

}  // namespace scalapack
}  // namespace cosma
