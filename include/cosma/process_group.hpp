// cosma::pg -- the handful of process-group operations the COSMA host layer needs from MPI, for boxes without MPI.
//
// The reference takes an MPI_Comm everywhere (src/cosma/multiply.hpp:47-54, cinterface.hpp:42-76, pxgemm.h) and uses it
// for (a) rank / size, (b) one broadcast of the ncclUniqueId per communicator (src/cosma/gpu/nccl_utils.cpp:21-42) and,
// in its tests and miniapps, (c) barriers, point-to-point gathers to rank 0 and a few reductions. All matrix traffic of
// the hot path goes over NCCL, never through this group. With a real MPI (define COSMA_B200_WITH_MPI) none of this is
// compiled; without one, <cosma/mpi_compat.hpp> maps the MPI names onto this group.
//
// Transport: one process per GPU started torchrun-style (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR and
// COSMA_B200_PG_PORT, default MASTER_PORT + 1); rank 0 listens, peers register, then every pair keeps one TCP socket.
// Collectives are rooted (gather to the first member, broadcast back): a few hundred bytes per call on this path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cosma {
namespace pg {

struct group;  // a communicator: an ordered list of world ranks

enum class dtype : int { byte_ = 0, char_, int_, long_long_, unsigned_long_long_, float_, double_, complex_float_, complex_double_, bool_ };
enum class op : int { sum = 0, min, max };
std::size_t dtype_size(dtype t);

// Idempotent. Reads the environment; WORLD_SIZE unset or 1 -> a single-process world without sockets.
// Throws std::runtime_error when the rendezvous fails.
void init();
bool initialized();
void finalize();
group* world();
int local_rank();  // LOCAL_RANK (the GPU to use), 0 if unset

int rank(const group* g);
int size(const group* g);
std::uint64_t id(const group* g);  // equal on all members, unique among live groups

void barrier(group* g);
void bcast(group* g, void* buf, std::size_t bytes, int root);
void send(group* g, const void* buf, std::size_t bytes, int dst, int tag);
void recv(group* g, void* buf, std::size_t bytes, int src, int tag);
// receive from whichever member sends a message with this tag first (MPI_ANY_SOURCE); returns the sender's rank in g
int recv_any(group* g, void* buf, std::size_t bytes, int tag);
// recv must hold size(g) * bytes on the root
void gather(group* g, const void* send, std::size_t bytes, void* recv, int root);
void allgather(group* g, const void* send, std::size_t bytes, void* recv);
void reduce(group* g, const void* send, void* recv, int count, dtype t, op o, int root);
void allreduce(group* g, const void* send, void* recv, int count, dtype t, op o);

// MPI_Comm_split: members with the same color form a group ordered by (key, old rank); color < 0 -> nullptr.
group* split(group* g, int color, int key);
// MPI_Comm_create_group: the listed ranks of g (in this order) form a new group. Collective over the LISTED ranks only -- in
// fact purely local here: membership and identity follow from (g, ranks, tag), which every member passes identically.
// Returns nullptr on a rank that is not listed.
group* create_group(group* g, const std::vector<int>& ranks, int tag);
group* dup(group* g);
void free(group* g);  // the world group is never freed

}  // namespace pg
}  // namespace cosma
