// MPI names for hosts without MPI.
//
// The reference's public signatures carry MPI_Comm (multiply.hpp:47-54, cinterface.hpp:42-76) and its tests and miniapps
// call a dozen MPI functions around them. With a real MPI, compile with -DCOSMA_B200_WITH_MPI and this header only includes
// <mpi.h>. Without one (this image has none) it supplies exactly the subset that cosma's headers, tests and miniapps use,
// as inline wrappers over cosma::pg (process_group.hpp) -- no MPI_* symbol is exported from any library, so linking a real
// MPI next to it later cannot clash.
#pragma once

#if defined(COSMA_B200_WITH_MPI)
#include <mpi.h>
#else
#include <cosma/process_group.hpp>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#define COSMA_B200_MPI_COMPAT 1

typedef cosma::pg::group* MPI_Comm;
typedef cosma::pg::dtype MPI_Datatype;
typedef cosma::pg::op MPI_Op;
typedef int MPI_Status;

#define MPI_COMM_WORLD (cosma::pg::world())
#define MPI_COMM_NULL (static_cast<MPI_Comm>(nullptr))
#define MPI_SUCCESS 0
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-2)
#define MPI_STATUS_IGNORE (static_cast<MPI_Status*>(nullptr))
#define MPI_STATUSES_IGNORE (static_cast<MPI_Status*>(nullptr))
#define MPI_IN_PLACE (reinterpret_cast<void*>(-1))
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

#define MPI_BYTE cosma::pg::dtype::byte_
#define MPI_CHAR cosma::pg::dtype::char_
#define MPI_INT cosma::pg::dtype::int_
#define MPI_LONG_LONG cosma::pg::dtype::long_long_
#define MPI_UNSIGNED_LONG_LONG cosma::pg::dtype::unsigned_long_long_
#define MPI_FLOAT cosma::pg::dtype::float_
#define MPI_DOUBLE cosma::pg::dtype::double_
#define MPI_C_FLOAT_COMPLEX cosma::pg::dtype::complex_float_
#define MPI_C_DOUBLE_COMPLEX cosma::pg::dtype::complex_double_
#define MPI_C_BOOL cosma::pg::dtype::bool_
#define MPI_SUM cosma::pg::op::sum
#define MPI_MIN cosma::pg::op::min
#define MPI_MAX cosma::pg::op::max

inline int MPI_Init(int*, char***) { cosma::pg::init(); return MPI_SUCCESS; }
inline int MPI_Init_thread(int*, char***, int required, int* provided) {
    cosma::pg::init();
    if (provided) *provided = required;
    return MPI_SUCCESS;
}
inline int MPI_Initialized(int* flag) { *flag = cosma::pg::initialized() ? 1 : 0; return MPI_SUCCESS; }
inline int MPI_Finalize() { cosma::pg::finalize(); return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::_Exit(code); }
inline double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int MPI_Comm_rank(MPI_Comm c, int* r) { *r = cosma::pg::rank(c); return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm c, int* s) { *s = cosma::pg::size(c); return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm c) { cosma::pg::barrier(c); return MPI_SUCCESS; }
inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
    cosma::pg::bcast(c, buf, cosma::pg::dtype_size(t) * static_cast<std::size_t>(count), root);
    return MPI_SUCCESS;
}
inline int MPI_Send(const void* buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    cosma::pg::send(c, buf, cosma::pg::dtype_size(t) * static_cast<std::size_t>(count), dst, tag);
    return MPI_SUCCESS;
}
inline int MPI_Ssend(const void* buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    return MPI_Send(buf, count, t, dst, tag, c);
}
inline int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* status) {
    const std::size_t bytes = cosma::pg::dtype_size(t) * static_cast<std::size_t>(count);
    if (src == MPI_ANY_SOURCE) {
        const int from = cosma::pg::recv_any(c, buf, bytes, tag);
        if (status) *status = from;  // MPI_Status is the source rank in this subset
    } else {
        cosma::pg::recv(c, buf, bytes, src, tag);
        if (status) *status = src;
    }
    return MPI_SUCCESS;
}
inline int MPI_Gather(const void* send, int scount, MPI_Datatype st, void* recv, int, MPI_Datatype, int root, MPI_Comm c) {
    cosma::pg::gather(c, send, cosma::pg::dtype_size(st) * static_cast<std::size_t>(scount), recv, root);
    return MPI_SUCCESS;
}
inline int MPI_Allgather(const void* send, int scount, MPI_Datatype st, void* recv, int, MPI_Datatype, MPI_Comm c) {
    cosma::pg::allgather(c, send, cosma::pg::dtype_size(st) * static_cast<std::size_t>(scount), recv);
    return MPI_SUCCESS;
}
inline int MPI_Reduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) {
    cosma::pg::reduce(c, send == MPI_IN_PLACE ? recv : send, recv, count, t, o, root);
    return MPI_SUCCESS;
}
inline int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op o, MPI_Comm c) {
    cosma::pg::allreduce(c, send == MPI_IN_PLACE ? recv : send, recv, count, t, o);
    return MPI_SUCCESS;
}
inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* out) {
    *out = cosma::pg::split(c, color == MPI_UNDEFINED ? -1 : color, key);
    return MPI_SUCCESS;
}
// groups: ordered lists of ranks of the communicator they were taken from
struct cosma_b200_mpi_group {
    MPI_Comm parent;
    std::vector<int> ranks;
};
typedef cosma_b200_mpi_group* MPI_Group;
#define MPI_GROUP_NULL (static_cast<MPI_Group>(nullptr))
inline int MPI_Comm_group(MPI_Comm c, MPI_Group* g) {
    *g = new cosma_b200_mpi_group{c, {}};
    for (int r = 0; r < cosma::pg::size(c); ++r) (*g)->ranks.push_back(r);
    return MPI_SUCCESS;
}
inline int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out) {
    *out = new cosma_b200_mpi_group{g->parent, {}};
    for (int i = 0; i < n; ++i) (*out)->ranks.push_back(g->ranks[ranks[i]]);
    return MPI_SUCCESS;
}
inline int MPI_Group_excl(MPI_Group g, int n, const int* ranks, MPI_Group* out) {
    *out = new cosma_b200_mpi_group{g->parent, {}};
    for (std::size_t i = 0; i < g->ranks.size(); ++i) {
        bool drop = false;
        for (int j = 0; j < n; ++j) drop = drop || ranks[j] == static_cast<int>(i);
        if (!drop) (*out)->ranks.push_back(g->ranks[i]);
    }
    return MPI_SUCCESS;
}
inline int MPI_Group_size(MPI_Group g, int* n) { *n = static_cast<int>(g->ranks.size()); return MPI_SUCCESS; }
inline int MPI_Group_free(MPI_Group* g) {
    if (g) { delete *g; *g = MPI_GROUP_NULL; }
    return MPI_SUCCESS;
}
// collective over the members of `g` only (the reference cuts the active ranks out of a communicator this way:
// communicator.cpp:282-343, tests/multiply.cpp:7-36)
inline int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* out) {
    *out = cosma::pg::create_group(c, g->ranks, tag);
    return MPI_SUCCESS;
}
inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* out) { *out = cosma::pg::dup(c); return MPI_SUCCESS; }
inline int MPI_Comm_free(MPI_Comm* c) {
    if (c) { cosma::pg::free(*c); *c = MPI_COMM_NULL; }
    return MPI_SUCCESS;
}
#endif  // COSMA_B200_WITH_MPI

namespace cosma {
// a key that identifies a communicator across calls (the reference compares communicators with MPI_Comm_compare in
// cosma_context::register_state, context.cpp:80-125)
inline unsigned long long comm_key(MPI_Comm comm) {
#if defined(COSMA_B200_WITH_MPI)
    return static_cast<unsigned long long>(MPI_Comm_c2f(comm));
#else
    return comm ? cosma::pg::id(comm) : 0ull;
#endif
}
}  // namespace cosma
