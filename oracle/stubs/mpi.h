/* TEST INFRASTRUCTURE ONLY -- single-rank MPI stand-in ("unimpi") used to compile the UNMODIFIED
 * reference sources under /root/reference into oracle/_ref/ (there is no MPI in this image).
 * One process, rank 0 of 1: collectives degenerate to memcpy, communicator constructors hand out
 * fresh handles. RMA entry points exist so one_sided_communicator.cpp links, and abort if reached
 * (the reference only uses them with COSMA_OVERLAP_COMM_AND_COMP=ON, default OFF).
 * Nothing under cosma_b200/ may include or link this. */
#ifndef ORACLE_UNIMPI_H
#define ORACLE_UNIMPI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Group;
typedef int MPI_Win;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_GROUP_NULL 0
#define MPI_GROUP_EMPTY 1
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL 0
#define MPI_WIN_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_SIMILAR 2
#define MPI_UNEQUAL 3
#define MPI_COMM_TYPE_SHARED 1
#define MPI_MODE_NOCHECK 1
#define MPI_MODE_NOSTORE 2
#define MPI_MODE_NOPUT 4
#define MPI_MODE_NOPRECEDE 8
#define MPI_MODE_NOSUCCEED 16
#define MPI_LOCK_EXCLUSIVE 1
#define MPI_LOCK_SHARED 2
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_MAX_PROCESSOR_NAME 256

/* datatype handle = size in bytes in the low byte, a tag in the next */
#define MPI_CHAR 0x0101
#define MPI_UNSIGNED_CHAR 0x0201
#define MPI_BYTE 0x0301
#define MPI_C_BOOL 0x0401
#define MPI_SHORT 0x0502
#define MPI_INT 0x0604
#define MPI_UINT32_T 0x0704
#define MPI_FLOAT 0x0804
#define MPI_UNSIGNED 0x0904
#define MPI_LONG 0x0a08
#define MPI_UNSIGNED_LONG 0x0b08
#define MPI_LONG_LONG 0x0c08
#define MPI_LONG_LONG_INT 0x0c08
#define MPI_UNSIGNED_LONG_LONG 0x0d08
#define MPI_DOUBLE 0x0e08
#define MPI_C_FLOAT_COMPLEX 0x0f08
#define MPI_C_DOUBLE_COMPLEX 0x1010
#define MPI_CXX_FLOAT_COMPLEX 0x0f08
#define MPI_CXX_DOUBLE_COMPLEX 0x1010
#define MPI_DATATYPE_NULL 0

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_REPLACE 4

int MPI_Init(int*, char***);
int MPI_Init_thread(int*, char***, int, int*);
int MPI_Finalize(void);
int MPI_Finalized(int*);
int MPI_Initialized(int*);
int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
int MPI_Get_processor_name(char*, int*);

int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_dup(MPI_Comm, MPI_Comm*);
int MPI_Comm_free(MPI_Comm*);
/* communicator attributes (used by the B200 host layer to tie its NCCL communicators to the life of the MPI communicator) */
#define MPI_KEYVAL_INVALID (-1)
typedef int MPI_Comm_copy_attr_function(MPI_Comm, int, void*, void*, void*, int*);
typedef int MPI_Comm_delete_attr_function(MPI_Comm, int, void*, void*);
#define MPI_COMM_NULL_COPY_FN ((MPI_Comm_copy_attr_function*)0)
int MPI_Comm_create_keyval(MPI_Comm_copy_attr_function*, MPI_Comm_delete_attr_function*, int*, void*);
int MPI_Comm_set_attr(MPI_Comm, int, void*);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm*);
int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm*);
int MPI_Comm_group(MPI_Comm, MPI_Group*);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm*);
int MPI_Comm_create_group(MPI_Comm, MPI_Group, int, MPI_Comm*);
int MPI_Comm_compare(MPI_Comm, MPI_Comm, int*);
int MPI_Dist_graph_create(MPI_Comm, int, const int[], const int[], const int[], const int[], MPI_Info, int, MPI_Comm*);

int MPI_Group_incl(MPI_Group, int, const int[], MPI_Group*);
int MPI_Group_excl(MPI_Group, int, const int[], MPI_Group*);
int MPI_Group_free(MPI_Group*);
int MPI_Group_union(MPI_Group, MPI_Group, MPI_Group*);
int MPI_Group_intersection(MPI_Group, MPI_Group, MPI_Group*);
int MPI_Group_compare(MPI_Group, MPI_Group, int*);
int MPI_Group_translate_ranks(MPI_Group, int, const int[], MPI_Group, int[]);
int MPI_Group_size(MPI_Group, int*);
int MPI_Group_rank(MPI_Group, int*);

int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int[], const int[], MPI_Datatype, MPI_Comm);
int MPI_Gather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce_scatter(const void*, void*, const int[], MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce_scatter_block(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);

int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Ssend(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitany(int, MPI_Request[], int*, MPI_Status*);
int MPI_Waitall(int, MPI_Request[], MPI_Status[]);
int MPI_Test(MPI_Request*, int*, MPI_Status*);
int MPI_Startall(int, MPI_Request[]);
int MPI_Probe(int, int, MPI_Comm, MPI_Status*);
int MPI_Get_count(const MPI_Status*, MPI_Datatype, int*);
int MPI_Get_elements(const MPI_Status*, MPI_Datatype, int*);

int MPI_Info_create(MPI_Info*);
int MPI_Info_set(MPI_Info, const char*, const char*);
int MPI_Info_free(MPI_Info*);

int MPI_Win_create(void*, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win*);
int MPI_Win_free(MPI_Win*);
int MPI_Win_fence(int, MPI_Win);
int MPI_Win_lock(int, int, int, MPI_Win);
int MPI_Win_unlock(int, MPI_Win);
int MPI_Win_lock_all(int, MPI_Win);
int MPI_Win_unlock_all(MPI_Win);
int MPI_Win_flush_local(int, MPI_Win);
int MPI_Get(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win);
int MPI_Rget(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win, MPI_Request*);
int MPI_Accumulate(const void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Op, MPI_Win);

#ifdef __cplusplus
}
#endif
#endif
